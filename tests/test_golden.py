"""The committed oracle fixture (tests/golden/make_golden.py) pins the oracle against drift; GPU parity vs the same file."""
import os

import numpy as np
import pytest

from oracle import cqt_oracle as co, labels_oracle as lo

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_cqt_seed0.npz"))


def test_oracle_reproduces_fixture(basis_cache):
    y = GOLD["audio"]
    assert np.abs(co.halfband_taps() - GOLD["taps"]).max() == 0
    for i in range(9):
        f, p, C = co.segment_features(y[i * 2205: i * 2205 + 4410], 22050, fmin=co.note_to_hz_C(1), _basis_cache=basis_cache,
                                      return_pre_cut=True)
        assert np.abs(C - GOLD["cqt"][i]).max() <= 1e-6 * np.abs(C).max()
        assert np.abs(p - GOLD["pre_cut"][i]).max() < 1e-3
    assert np.array_equal(lo.rasterize_events_numpy(GOLD["onset"], GOLD["dur"], GOLD["pitch"], GOLD["times"]), GOLD["labels"])


@pytest.mark.gpu
def test_gpu_matches_fixture(lib, recipe):
    import torch
    from gtc_b200 import ops
    dev = torch.device("cuda")
    plan = ops.CqtPlan(recipe)
    y = GOLD["audio"]
    clip_off, seg_off = plan.offsets([len(y)])
    got = plan.segments_db(torch.from_numpy(y).to(dev), torch.from_numpy(clip_off).to(dev), torch.from_numpy(seg_off).to(dev), 9).cpu().numpy()
    pre = GOLD["pre_cut"]
    ok = np.abs(pre + 60) > 0.02
    assert np.abs(got - GOLD["features"])[ok].max() < 0.01
    t_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tabs, _ = ops.rasterize_tabs(t_(GOLD["onset"]), t_(GOLD["dur"]), t_(GOLD["pitch"]), t_(np.array([0, 40])), t_(GOLD["times"]),
                                 t_(np.array([0, 25])))
    assert np.array_equal(tabs.cpu().numpy(), GOLD["labels"])
