"""The patch oracle is pinned against the real torch.nn.functional.interpolate (the reference's own call)."""
import numpy as np
import pytest
import torch

from oracle import patches_oracle as po


def rand_db(rng, shape):
    x = -60 * rng.random(shape)
    x[rng.random(shape) < 0.4] = -120.0
    return x.astype(np.float32)


@pytest.mark.parametrize("shape", [(96, 5), (96, 9), (84, 130), (7, 3)])
@pytest.mark.parametrize("size", [(224, 224), (96, 160)])
def test_bicubic_restatement_matches_torch(shape, size):
    db = rand_db(np.random.default_rng(shape[1]), shape)
    a, b = po.vit_patch(db, size), po.vit_patch_torch(db, size)
    assert a.shape == (3,) + size and a.dtype == np.float32
    assert np.abs(a - b).max() < 2e-5
    assert (a[0] == a[1]).all() and (a[0] == a[2]).all()


def test_overshoot_is_not_clipped():
    db = np.full((96, 5), -120, np.float32)
    db[40:44, 2] = 0.0
    p = po.vit_patch(db)
    assert p.max() > 1.0 and p.min() < 0.0


def test_bilinear_matches_torch():
    g = po.vit_normalize(rand_db(np.random.default_rng(0), (96, 5)))
    t = torch.nn.functional.interpolate(torch.tensor(g)[None, None], size=(224, 224), mode='bilinear',
                                        align_corners=False)[0, 0].numpy()
    assert np.abs(po.bilinear_resize(g, 224, 224) - t).max() < 1e-5


def test_cnn_contract():
    db = rand_db(np.random.default_rng(1), (96, 5))
    p = po.cnn_patch(db)
    assert p.shape == (3, 224, 224) and p.dtype == np.float32
    # undo the normalisation: all three channels carry the same grey picture in [0, 1]
    mean = np.array(po.IMAGENET_MEAN, np.float32)[:, None, None]
    std = np.array(po.IMAGENET_STD, np.float32)[:, None, None]
    grey = p * std + mean
    assert np.abs(grey[0] - grey[1]).max() < 1e-6 and grey.min() > -1e-6 and grey.max() < 1 + 1e-6
    # top image row is the highest CQT bin
    flat = np.full((96, 5), -120, np.float32)
    flat[95] = 0
    assert (po.cnn_patch(flat)[0, 0] > po.cnn_patch(flat)[0, -1]).all()
