"""The product's operator design (float64 linear algebra) against the procedural float32 oracle."""
import numpy as np
import pytest

from gtc_b200 import cqt_design as cd
from oracle import cqt_oracle as co
from conftest import make_test_audio


def test_decimator_designs_agree():
    assert np.abs(cd.decimator_taps() - co.halfband_taps()).max() < 1e-15


def test_recipe_defaults_follow_reference():
    r = cd.CqtRecipe()
    assert (r.seg_len, r.seg_hop, r.n_octaves, cd.n_frames_of(r)) == (4410, 2205, 8, 5)
    assert abs(r.fmin_hz - 32.7032) < 1e-4
    r44 = cd.CqtRecipe(sr=44100.0)
    assert (r44.seg_len, r44.seg_hop, cd.n_frames_of(r44)) == (8820, 4410, 9)


def test_operator_reproduces_oracle_cqt(basis_cache):
    r = cd.CqtRecipe()
    A = cd.get_operator(r).astype(np.float64)
    assert A.shape == (960, 4410)
    for seed in range(3):
        x = make_test_audio(4410, seed)
        C = co.cqt(x, sr=22050, fmin=co.note_to_hz_C(1), _basis_cache=basis_cache)
        y = (A @ x.astype(np.float64)).reshape(5, 96, 2)
        Cg = (y[..., 0] + 1j * y[..., 1]).T
        assert np.abs(Cg - C).max() < 2e-6 * np.abs(C).max()


def test_non_power_hop_is_rejected():
    with pytest.raises(ValueError):
        cd.build_operator(cd.CqtRecipe(hop_length=1000))
