"""GPU parity: segment CQT + dB features through the C ABI against the CPU oracle (tolerances of BASELINE.json:
dB within 0.01 dB, relative magnitude error within 1e-4)."""
import numpy as np
import pytest
import torch

from conftest import make_test_audio
from oracle import cqt_oracle as co

pytestmark = pytest.mark.gpu

SR = 22050
ENGINES = {"simt": 1, "tcgen05": 0, "fp16x2": 2}


def oracle_segments(clips, recipe, cache):
    out = []
    for y in clips:
        n = co.num_segments(len(y), recipe.seg_len, recipe.seg_hop)
        for i in range(n):
            seg = y[i * recipe.seg_hop: i * recipe.seg_hop + recipe.seg_len]
            out.append(co.segment_features(seg, SR, fmin=co.note_to_hz_C(1), _basis_cache=cache, return_pre_cut=True))
    return out


def run_gpu(plan, clips, complex_out=False):
    dev = torch.device("cuda")
    lens = [len(c) for c in clips]
    clip_off, seg_off = plan.offsets(lens)
    flat = torch.from_numpy(np.concatenate(clips) if clips else np.zeros(0, np.float32)).to(dev)
    if flat.numel() == 0:
        flat = torch.zeros(1, dtype=torch.float32, device=dev)
    n_seg = int(seg_off[-1])
    co_t, so_t = torch.from_numpy(clip_off).to(dev), torch.from_numpy(seg_off).to(dev)
    if complex_out:
        return plan.segments_complex(flat, co_t, so_t, n_seg).cpu().numpy()
    return plan.segments_db(flat, co_t, so_t, n_seg).cpu().numpy()


@pytest.fixture(scope="module", params=list(ENGINES))
def plan(request, recipe, lib):
    from gtc_b200 import ops
    p = ops.CqtPlan(recipe, engine=ENGINES[request.param])
    yield p
    p.close()


@pytest.fixture(scope="module")
def clips():
    # ragged lengths incl. a clip shorter than one window (0 segments) and an exact-fit clip
    lens = [SR * 2, 4410, 3000, SR + 777, 2205 * 9]
    return [make_test_audio(n, seed=10 + i) for i, n in enumerate(lens)]


def test_complex_matches_oracle(plan, clips, recipe, basis_cache):
    ref = oracle_segments(clips, recipe, basis_cache)
    got = run_gpu(plan, clips, complex_out=True)
    assert got.shape == (len(ref), 96, 5)
    for i, (_, pre, C) in enumerate(ref):
        peak = np.abs(C).max()
        assert np.abs(got[i] - C).max() < 2e-5 * peak, f"segment {i}"
        keep = np.abs(C) > peak * 10 ** (-15.5 / 20)          # elements that survive the -60 dB (|C|^4) cut
        rel = np.abs(np.abs(got[i][keep]) - np.abs(C[keep])) / np.abs(C[keep])
        assert rel.max() < 1e-4, f"segment {i}: relative magnitude error {rel.max()}"


def test_db_features_match_oracle(plan, clips, recipe, basis_cache):
    ref = oracle_segments(clips, recipe, basis_cache)
    got = run_gpu(plan, clips)
    assert got.shape == (len(ref), 96, 5) and got.dtype == np.float32
    tol = 0.01
    for i, (cut, pre, _) in enumerate(ref):
        above = pre > -60 + 2 * tol
        below = pre < -60 - 2 * tol
        assert np.abs(got[i][above] - pre[above]).max() < tol, f"segment {i}"
        assert (got[i][below] == -120).all(), f"segment {i}"
        edge = ~(above | below)                                # within tolerance of the threshold: either side is fine
        assert ((got[i][edge] == -120) | (np.abs(got[i][edge] - pre[edge]) < tol)).all()
        assert got[i].max() == 0.0


def test_value_set(plan, clips):
    got = run_gpu(plan, clips)
    assert ((got == -120) | ((got >= -60) & (got <= 0))).all()


def test_silence_and_quiet_segments(plan, recipe):
    z = np.zeros(4410 * 2, np.float32)
    q = (1e-4 * make_test_audio(4410 * 2, 3)).astype(np.float32)     # peak |C|^4 < amin -> all 0 dB (A.3)
    got = run_gpu(plan, [z, q])
    assert got.shape[0] == 6
    assert (got[:3] == 0).all()
    want = np.stack([co.segment_features(q[i * 2205: i * 2205 + 4410], SR, fmin=co.note_to_hz_C(1)) for i in range(3)])
    assert np.abs(got[3:] - want).max() < 0.01


def test_amin_floor_visible(plan):
    t = np.arange(4410) / SR
    y = (0.011 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)     # peak |C| ~ 0.16 -> floor 20log10(1e-5/ref) in (-60,0)
    got = run_gpu(plan, [y])[0]
    want, pre, _ = co.segment_features(y, SR, fmin=co.note_to_hz_C(1), return_pre_cut=True)
    floor = pre.min()
    assert -60 < floor < -1
    assert np.abs(got - want).max() < 0.01 and abs(got.min() - floor) < 0.01


def test_no_clips_and_no_segments(plan):
    assert run_gpu(plan, [np.zeros(100, np.float32)]).shape == (0, 96, 5)


def test_linearity_property_at_scale(plan):
    """Size-independent property on a larger batch: C(a + 2b) == C(a) + 2 C(b)."""
    n = SR * 20
    a, b = make_test_audio(n, 1), make_test_audio(n, 2)
    Ca, Cb = run_gpu(plan, [a], True), run_gpu(plan, [b], True)
    Cab = run_gpu(plan, [(a + 2 * b).astype(np.float32)], True)
    assert Cab.shape[0] == 199
    scale = np.abs(Cab).max(axis=(1, 2), keepdims=True)
    assert (np.abs(Cab - (Ca + 2 * Cb)) / scale).max() < 2e-5


def test_engines_agree(recipe, lib, clips):
    from gtc_b200 import ops
    p0, p1, p2 = ops.CqtPlan(recipe, engine=0), ops.CqtPlan(recipe, engine=1), ops.CqtPlan(recipe, engine=2)
    a, b, c = run_gpu(p0, clips, True), run_gpu(p1, clips, True), run_gpu(p2, clips, True)
    scale = np.abs(b).max(axis=(1, 2), keepdims=True)
    assert (np.abs(a - b) / scale).max() < 2e-5      # fp32 FMA chain (K = 4410) vs 3xTF32 + split accumulation
    assert (np.abs(c - b) / scale).max() < 2e-5      # ... vs fp16x2 + split accumulation
    p0.close(); p1.close(); p2.close()


def test_non_overlapping_windows_44k(lib):
    """new_cqt.py recipe: sr literal 44100, non-overlapping 0.2 s windows -> (96, 9) features, P = 1."""
    from gtc_b200 import ops, CqtRecipe
    r = CqtRecipe(sr=44100.0, hop_size=0.2)
    plan = ops.CqtPlan(r)
    y = make_test_audio(8820 * 3 + 100, 4, sr=44100.0)
    got = run_gpu(plan, [y])
    assert got.shape == (3, 96, 9)
    cache = {}
    for i in range(3):
        want, pre, _ = co.segment_features(y[i * 8820:(i + 1) * 8820], 44100, _basis_cache=cache, return_pre_cut=True)
        ok = np.abs(pre + 60) > 0.02
        assert np.abs(got[i] - want)[ok].max() < 0.01
    plan.close()


def test_pcm16_input_is_bit_identical_to_librosa_fp32(plan, clips):
    """int16 PCM uploaded as-is and converted on the device (x/32768) == the fp32 array librosa.load returns (cqt.py:23)."""
    dev = torch.device("cuda")
    pcm = [np.clip(np.round(c.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16) for c in clips]
    as_f32 = [p.astype(np.float32) / np.float32(32768.0) for p in pcm]
    want = run_gpu(plan, as_f32)
    clip_off, seg_off = plan.offsets([len(c) for c in pcm])
    got = plan.segments_db(torch.from_numpy(np.concatenate(pcm)).to(dev), torch.from_numpy(clip_off).to(dev),
                           torch.from_numpy(seg_off).to(dev), int(seg_off[-1])).cpu().numpy()
    assert np.array_equal(got, want)


def test_fused_finish_option_is_bit_identical(recipe, lib, clips):
    """GTC_OPT_FUSE_FINISH: the dB finish done by the tcgen05 epilogue (the CTA completing a 128-row block converts it)
    gives the same bits as the separate finish_db_kernel, for ragged clips and for a chunk of several tile waves."""
    from gtc_b200 import ops, _lib
    big = [make_test_audio(SR * 3, seed=90 + i) for i in range(40)]            # 40 x 29 segments = 1160 rows = 10 row blocks
    for engine in (ENGINES["fp16x2"], ENGINES["tcgen05"]):
        p = ops.CqtPlan(recipe, engine=engine)
        for cl in (clips, big):
            p.configure(_lib.GTC_OPT_FUSE_FINISH, 0)
            a = run_gpu(p, cl)
            p.configure(_lib.GTC_OPT_FUSE_FINISH, 1)
            b = run_gpu(p, cl)
            c = run_gpu(p, cl)                                                    # counters are left clean for the next call
            assert np.array_equal(a, b) and np.array_equal(b, c)
        p.close()


def test_zero_skipping_schedule_equals_dense_loops(recipe, lib, clips, monkeypatch):
    """The tensor-core engine multiplies a k-block only with the frames whose operator rows are non-zero on it (the plan's
    schedule, cqt_gemm_tc.cu); GTC_TC_DENSE=1 at plan creation restores the dense loops.  Same products, K splits cut at
    other places: complex outputs agree to fp32 rounding, dB features to 1e-3 dB away from the cut."""
    from gtc_b200 import ops
    big = clips + [make_test_audio(SR * 3, seed=190 + i) for i in range(40)]      # several row blocks, rotated tile order
    p_s = ops.CqtPlan(recipe)
    monkeypatch.setenv("GTC_TC_DENSE", "1")
    p_d = ops.CqtPlan(recipe)
    monkeypatch.delenv("GTC_TC_DENSE")
    a, b = run_gpu(p_s, big, complex_out=True), run_gpu(p_d, big, complex_out=True)
    peak = np.abs(b).max(axis=(1, 2), keepdims=True)
    assert (np.abs(a - b) / np.maximum(peak, 1e-20)).max() < 3e-6
    da, db = run_gpu(p_s, big), run_gpu(p_d, big)
    both = (da > -59.9) & (db > -59.9)
    assert np.abs(da - db)[both].max() < 1e-3 and ((da == -120) != (db == -120)).mean() < 1e-4
    p_s.close(); p_d.close()


def test_plan_from_a_caller_supplied_operator(recipe, lib, clips):
    """The pinning hook (scripts/pin_with_librosa.py): CqtPlan(operator=A) evaluates the matrix the caller hands over -- e.g.
    one measured from the real librosa.cqt -- instead of the designed one.  The designed matrix gives the default plan's
    bits; a scaled copy scales the complex output."""
    from gtc_b200 import ops
    from gtc_b200.cqt_design import get_operator
    A = np.array(get_operator(recipe), dtype=np.float32)
    p0, p1, p2 = ops.CqtPlan(recipe), ops.CqtPlan(recipe, operator=A), ops.CqtPlan(recipe, operator=0.5 * A)
    c0, c1, c2 = (run_gpu(p, clips, complex_out=True) for p in (p0, p1, p2))
    assert np.array_equal(c0, c1)
    assert np.abs(c2 - 0.5 * c0).max() <= 1e-6 * np.abs(c0).max()
    with pytest.raises(Exception):
        ops.CqtPlan(recipe, operator=A[:, :100])
    for p in (p0, p1, p2):
        p.close()
