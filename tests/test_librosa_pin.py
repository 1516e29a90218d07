"""CQT pinning kit (scripts/pin_with_librosa.py): runs against the real librosa / soxr wherever they import, against a
committed librosa golden file where one exists, and always against an oracle stand-in so the kit itself cannot rot.

In the build image librosa and soxr are not installable: the first two tests then SKIP with that reason and
the CQT row of DESIGN.md stays "parity unpinned"; profiles/r02_tap_sensitivity.md bounds what that can cost."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import pin_with_librosa as pin  # noqa: E402

HAVE_LIBROSA = importlib.util.find_spec("librosa") is not None and importlib.util.find_spec("soxr") is not None


@pytest.mark.skipif(not HAVE_LIBROSA, reason="librosa / soxr not importable here: CQT parity stays unpinned (restated oracle only)")
def test_oracle_and_operator_match_real_librosa():
    rep = pin.run(n_segments=32, with_operator=True, use_gpu=False)
    assert rep["status"] == "pinned", rep
    assert rep["oracle_vs_reference"]["max_db_delta_above_cut"] <= pin.DB_GATE, rep
    assert rep["oracle_vs_reference"]["max_rel_magnitude_error_above_cut"] <= pin.REL_GATE, rep
    assert rep["operator"]["rel_fro_diff"] < 1e-4, rep
    assert max(rep["resample_vs_reference"].values()) < 1e-5, rep


@pytest.mark.skipif(not os.path.exists(pin.GOLDEN), reason="tests/golden/librosa_pin.npz not generated yet "
                    "(python scripts/pin_with_librosa.py --write-golden on a machine with librosa + soxr)")
def test_oracle_matches_committed_librosa_golden():
    from oracle import cqt_oracle as o
    g = np.load(pin.GOLDEN)
    cache, db, mag = {}, [], []
    for s in g["segments"]:
        _, pre, C = o.segment_features(s, int(g["sr"]), fmin=o.note_to_hz_C(1), _basis_cache=cache, return_pre_cut=True)
        db.append(pre)
        mag.append(np.abs(C))
    d = pin.diff_features(g["db_pre_cut"], g["magnitude"], np.stack(db), np.stack(mag))
    assert d["max_db_delta_above_cut"] <= pin.DB_GATE and d["max_rel_magnitude_error_above_cut"] <= pin.REL_GATE, d
    for key in ("soxr_4410", "soxr_8820"):
        x = g["segments"][0] if key.endswith("4410") else np.concatenate([g["segments"][0], g["segments"][min(2, len(g["segments"]) - 1)]])
        assert np.abs(o.resample_2to1(x) - g[key]).max() < 1e-5


def test_kit_machinery_with_oracle_stand_in():
    """The kit's own code paths (tap measurement by impulse response, recipe diff, measured-taps re-evaluation) driven by a
    stand-in backend built from the oracle: everything must come back identical.  Pins nothing -- it keeps the kit alive."""
    rep = pin.run(pin.oracle_backend(), n_segments=4, with_operator=False, use_gpu=False)
    assert rep["status"] == "pinned" and "NOT librosa" in rep["backend"]
    assert rep["taps"]["measured_len"] == rep["taps"]["restated_len"] == 389
    assert rep["taps"]["max_abs_diff"] < 1e-15
    assert rep["oracle_vs_reference"]["max_db_delta_above_cut"] == 0.0
    assert rep["oracle_with_measured_taps_vs_reference"]["max_db_delta_above_cut"] < 1e-4


def test_measured_operator_equals_designed_operator_on_stand_in():
    """Unit impulses through the (stand-in) cqt rebuild the segment operator; it must equal cqt_design.build_operator, the
    matrix libgtc evaluates, to fp32 rounding -- the same check the kit makes against real librosa."""
    from gtc_b200 import cqt_design
    be = pin.oracle_backend()
    A_ref = pin.measured_operator(be, 4410, 22050, columns=np.arange(0, 4410, 63))      # every 63rd column keeps the test short
    A = cqt_design.build_operator(cqt_design.CqtRecipe())[:, ::63]
    assert A_ref.shape == A.shape == (960, 70)
    assert np.linalg.norm(A - A_ref) / np.linalg.norm(A_ref) < 2e-6


def test_unpinned_status_without_librosa():
    if HAVE_LIBROSA:
        pytest.skip("librosa is importable here")
    rep = pin.run(None)
    assert rep["status"] == "unpinned"
