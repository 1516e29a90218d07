"""GPU parity: patch assembly through the C ABI against the oracle and the real torch CPU interpolate."""
import numpy as np
import pytest
import torch

from oracle import patches_oracle as po

pytestmark = pytest.mark.gpu
TOL = 3e-5      # fp32 bicubic weights differ by ~1 ulp between implementations (see tests/test_oracle_patches.py)


def rand_db(rng, shape):
    x = -60 * rng.random(shape)
    x[rng.random(shape) < 0.4] = -120.0
    return x.astype(np.float32)


@pytest.mark.parametrize("t_in", [5, 9, 7])
@pytest.mark.parametrize("size", [(224, 224), (96, 160), (30, 50)])
def test_vit_patches(lib, t_in, size):
    from gtc_b200 import ops
    db = rand_db(np.random.default_rng(t_in), (37, 96, t_in))
    got = ops.patches(torch.from_numpy(db).cuda(), img_size=size).cpu().numpy()
    assert got.shape == (37, 3) + size
    for i in range(0, 37, 9):
        assert np.abs(got[i] - po.vit_patch_torch(db[i], size)).max() < TOL
        assert np.abs(got[i] - po.vit_patch(db[i], size)).max() < TOL
    assert (got[:, 0] == got[:, 1]).all() and (got[:, 0] == got[:, 2]).all()


def test_index_selects_and_orders(lib):
    from gtc_b200 import ops
    rng = np.random.default_rng(0)
    db = rand_db(rng, (300, 96, 5))
    idx = rng.permutation(300)[:128].astype(np.int64)
    got = ops.patches(torch.from_numpy(db).cuda(), index=torch.from_numpy(idx).cuda()).cpu().numpy()
    full = ops.patches(torch.from_numpy(db).cuda()).cpu().numpy()
    assert np.array_equal(got, full[idx])
    assert np.abs(got[5] - po.vit_patch(db[idx[5]])).max() < TOL


def test_cnn_patches(lib):
    from gtc_b200 import ops, _lib
    db = rand_db(np.random.default_rng(2), (20, 96, 5))
    got = ops.patches(torch.from_numpy(db).cuda(), mode=_lib.GTC_PATCH_CNN).cpu().numpy()
    for i in (0, 7, 19):
        assert np.abs(got[i] - po.cnn_patch(db[i])).max() < 1e-4     # values up to ~2.6 after /std
    db9 = rand_db(np.random.default_rng(3), (4, 96, 9))
    got9 = ops.patches(torch.from_numpy(db9).cuda(), mode=_lib.GTC_PATCH_CNN).cpu().numpy()
    assert np.abs(got9[2] - po.cnn_patch(db9[2])).max() < 1e-4


def test_empty_batch(lib):
    from gtc_b200 import ops
    out = ops.patches(torch.zeros((0, 96, 5), dtype=torch.float32, device="cuda"))
    assert out.shape == (0, 3, 224, 224)
