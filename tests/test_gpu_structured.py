"""GPU parity of the structured (multirate) CQT path -- decimation chain + per-octave filters -- through the C ABI
against the CPU oracle, for the recipes of cqt.py:55, tablature_generator.py:616-620 and whole clips.
Tolerances (BASELINE.json): relative magnitude error within 1e-4, dB within 0.01 dB (away from the -60 dB cut)."""
import numpy as np
import pytest
import torch

from conftest import make_test_audio
from oracle import cqt_oracle as co

pytestmark = pytest.mark.gpu

C2 = 65.40639132514966


def seg_tables(starts, valids, lens, dev):
    return (torch.tensor(starts, dtype=torch.int64, device=dev), torch.tensor(valids, dtype=torch.int32, device=dev),
            torch.tensor(lens, dtype=torch.int32, device=dev))


def check_complex(got, ref, tol=1e-4):
    peak = np.abs(ref).max()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= tol * peak, (np.abs(got - ref).max() / peak)


@pytest.fixture(scope="module")
def splan(lib):
    from gtc_b200 import ops, CqtRecipe
    p = ops.StructuredCqtPlan(CqtRecipe())
    yield p
    p.close()


def test_segments_of_ragged_clips_match_oracle(splan, recipe, basis_cache):
    """cqt.py:26-58 windows of ragged clips: complex CQT and dB features."""
    dev = torch.device("cuda")
    clips = [make_test_audio(n, seed=20 + i) for i, n in enumerate([22050 * 2, 4410, 3000, 22050 + 777])]
    off = np.concatenate([[0], np.cumsum([len(c) for c in clips])])
    starts, ref = [], []
    for c, y in enumerate(clips):
        for i in range(co.num_segments(len(y), recipe.seg_len, recipe.seg_hop)):
            starts.append(int(off[c]) + i * recipe.seg_hop)
            seg = y[i * recipe.seg_hop: i * recipe.seg_hop + recipe.seg_len]
            ref.append(co.segment_features(seg, 22050, fmin=co.note_to_hz_C(1), _basis_cache=basis_cache, return_pre_cut=True))
    n = len(starts)
    audio = torch.from_numpy(np.concatenate(clips)).to(dev)
    st, va, le = seg_tables(starts, [recipe.seg_len] * n, [recipe.seg_len] * n, dev)
    got_c = splan.segments_complex(audio, st, va, le, recipe.seg_len).cpu().numpy()
    got_db = splan.segments_db(audio, st, va, le, recipe.seg_len).cpu().numpy()
    assert got_c.shape == (n, 96, 5) and got_db.shape == (n, 96, 5)
    for i, (cut, pre, C) in enumerate(ref):
        check_complex(got_c[i], C)
        away = np.abs(pre + 60.0) > 0.02                      # elements not sitting on the -60 dB threshold
        assert np.abs(got_db[i] - cut)[away].max() <= 0.01


def test_structured_equals_operator_path(splan, recipe, lib):
    """Property at a larger size: the multirate evaluation and the collapsed tensor-core operator agree."""
    from gtc_b200 import ops
    dev = torch.device("cuda")
    n_clips, n = 24, 22050 * 10
    from gtc_b200 import synth
    audio = synth.pluck_clips(n_clips, n, sr=22050, seed=5, device=dev).reshape(-1)
    plan = ops.CqtPlan(recipe)
    clip_off, seg_off = plan.offsets([n] * n_clips)
    n_seg = int(seg_off[-1])
    C_op = plan.segments_complex(audio, torch.from_numpy(clip_off).to(dev), torch.from_numpy(seg_off).to(dev), n_seg)
    per = n_seg // n_clips
    starts = (np.arange(n_clips)[:, None] * n + np.arange(per)[None, :] * recipe.seg_hop).reshape(-1)
    st, va, le = seg_tables(starts.tolist(), [recipe.seg_len] * n_seg, [recipe.seg_len] * n_seg, dev)
    C_st = splan.segments_complex(audio, st, va, le, recipe.seg_len)
    peak = C_op.abs().amax(dim=(1, 2), keepdim=True)
    err = ((C_st - C_op).abs() / peak).max().item()
    assert err < 5e-5, err
    db_op = plan.segments_db(audio, torch.from_numpy(clip_off).to(dev), torch.from_numpy(seg_off).to(dev), n_seg)
    db_st = splan.segments_db(audio, st, va, le, recipe.seg_len)
    pre_cut_far = (db_op > -59.9) & (db_st > -59.9)
    assert (db_op - db_st).abs()[pre_cut_far].max().item() < 0.01
    assert ((db_op == -120.0) != (db_st == -120.0)).float().mean().item() < 1e-3       # only threshold straddlers differ
    plan.close()


def test_pcm16_input_is_bit_identical(splan, recipe):
    dev = torch.device("cuda")
    y = make_test_audio(22050, seed=31)
    pcm = np.clip(np.round(y * 32768.0), -32768, 32767).astype(np.int16)
    f32 = pcm.astype(np.float32) / 32768.0
    starts = list(range(0, len(y) - recipe.seg_len + 1, recipe.seg_hop))
    n = len(starts)
    st, va, le = seg_tables(starts, [recipe.seg_len] * n, [recipe.seg_len] * n, dev)
    a = splan.segments_db(torch.from_numpy(f32).to(dev), st, va, le, recipe.seg_len)
    b = splan.segments_db(torch.from_numpy(pcm).to(dev), st, va, le, recipe.seg_len)
    assert torch.equal(a, b)


def test_inference_recipe_three_second_segments(lib):
    """tablature_generator.py:616-620 + segment_audio :637-666: sr 22 050, hop 512, C2, 84 bins, |C| (power 1),
    amplitude_to_db(ref=np.max) with top_db 80 and no cut; 3 s segments, 50 % overlap, zero-padded tail."""
    from gtc_b200 import ops, CqtRecipe
    dev = torch.device("cuda")
    sr, seg_len = 22050, 66150
    hop = int(seg_len * 0.5)
    y = make_test_audio(int(sr * 7.3), seed=41)
    r = CqtRecipe(sr=float(sr), hop_length=512, n_bins=84, fmin=C2, power=1.0, cut_db=-np.inf)
    p = ops.StructuredCqtPlan(r)
    starts = list(range(0, len(y), hop))
    valids = [min(seg_len, len(y) - s) for s in starts]
    st, va, le = seg_tables(starts, valids, [seg_len] * len(starts), dev)
    audio = torch.from_numpy(y).to(dev)
    got_c = p.segments_complex(audio, st, va, le, seg_len).cpu().numpy()
    got_db = p.segments_db(audio, st, va, le, seg_len).cpu().numpy()
    assert got_c.shape == (len(starts), 84, 130)
    cache = {}
    for i, s in enumerate(starts):
        seg = y[s: s + seg_len]
        seg = np.pad(seg, (0, seg_len - len(seg)))
        C = co.cqt(seg, sr=sr, hop_length=512, fmin=C2, n_bins=84, _basis_cache=cache)
        check_complex(got_c[i], C)
        db = co.amplitude_to_db_amax(np.abs(C))
        assert db.min() >= -80.0 - 1e-4
        # 0.01 dB where the magnitude is resolved (|C| above 1e-4 of the peak sits 4e-5 relative error away at most)
        big = np.abs(C) > 3e-3 * np.abs(C).max()
        assert np.abs(got_db[i] - db)[big].max() <= 0.01
        assert np.abs(got_db[i] - db).max() <= 0.5
    p.close()


def test_whole_clips_of_different_lengths(lib):
    """librosa.cqt of whole clips (new_cqt.py:25 semantics at the clip level), variable lengths in one batch, 44.1 kHz."""
    from gtc_b200 import ops, CqtRecipe
    dev = torch.device("cuda")
    r = CqtRecipe(sr=44100.0)
    p = ops.StructuredCqtPlan(r)
    assert p.n_fft == 256
    lens = [44100, 30001, 8820, 1500]
    clips = [make_test_audio(n, seed=50 + i, sr=44100.0) for i, n in enumerate(lens)]
    off = np.concatenate([[0], np.cumsum(lens)])
    st, va, le = seg_tables(off[:-1].tolist(), lens, lens, dev)
    got = p.segments_complex(torch.from_numpy(np.concatenate(clips)).to(dev), st, va, le, max(lens)).cpu().numpy()
    for i, y in enumerate(clips):
        C = co.cqt(y, sr=44100, fmin=co.note_to_hz_C(1))
        T = C.shape[1]
        assert T == p.frames(len(y))
        check_complex(got[i, :, :T], C)
        assert np.all(got[i, :, T:] == 0)
    p.close()


def test_silent_and_quiet_segments(splan, recipe):
    """A.3 corner cases through the structured path: silence -> 0 dB everywhere; amin floor visible."""
    dev = torch.device("cuda")
    y = np.zeros(4410 * 2, dtype=np.float32)
    y[4410:] = 1e-3 * make_test_audio(4410, seed=7)
    st, va, le = seg_tables([0, 4410], [4410, 4410], [4410, 4410], dev)
    got = splan.segments_db(torch.from_numpy(y).to(dev), st, va, le, 4410).cpu().numpy()
    assert np.all(got[0] == 0.0)
    ref = co.segment_features(y[4410:], 22050, fmin=co.note_to_hz_C(1))
    pre = co.segment_features(y[4410:], 22050, fmin=co.note_to_hz_C(1), return_pre_cut=True)[1]
    away = np.abs(pre + 60.0) > 0.02
    assert np.abs(got[1] - ref)[away].max() <= 0.01


def test_argument_errors(splan):
    from gtc_b200 import _lib
    dev = torch.device("cuda")
    a = torch.zeros(100, device=dev)
    st, va, le = seg_tables([0], [100], [100], dev)
    with pytest.raises(_lib.GtcError):
        splan.segments_db(a.cpu(), st, va, le, 100)                  # host tensor: no CPU fallback
    small = torch.empty(16, dtype=torch.uint8, device=dev)
    import ctypes as C
    rc = _lib.load().gtc_scqt_segments_db(splan._h, C.c_void_p(a.data_ptr()), 0, C.c_void_p(st.data_ptr()), C.c_void_p(va.data_ptr()),
                                          C.c_void_p(le.data_ptr()), 1, 100, C.c_void_p(a.data_ptr()), C.c_void_p(small.data_ptr()), 16,
                                          4.0, 1e-5, 80.0, -60.0, -120.0, None)
    assert rc == _lib.load().gtc_version() * 0 - 3               # GTC_E_NOMEM


def test_tensor_core_path_equals_fp32_simt_path(lib, monkeypatch):
    """dB features: the tcgen05 evaluation (Toeplitz decimator + slotted response GEMMs on fp16 hi/lo planes, the default)
    against the fp32 CUDA-core evaluation of the same plan tables (GTC_SCQT_SIMT=1 at plan creation), on ragged whole
    clips at 44.1 kHz (n_fft 256, 8 octaves), 17 clips so that the 16-segment tile groups have a ragged tail."""
    from gtc_b200 import ops, CqtRecipe
    dev = torch.device("cuda")
    r = CqtRecipe(sr=44100.0)
    lens = [44100, 30001, 8820, 1500, 256, 70000] + [12000 + 977 * i for i in range(11)]
    clips = [make_test_audio(n, seed=150 + i, sr=44100.0) for i, n in enumerate(lens)]
    off = np.concatenate([[0], np.cumsum(lens)])
    st, va, le = seg_tables(off[:-1].tolist(), lens, lens, dev)
    audio = torch.from_numpy(np.concatenate(clips)).to(dev)
    p_tc = ops.StructuredCqtPlan(r)
    a = p_tc.segments_db(audio, st, va, le, max(lens)).cpu().numpy()
    monkeypatch.setenv("GTC_SCQT_SIMT", "1")
    p_simt = ops.StructuredCqtPlan(r)
    monkeypatch.delenv("GTC_SCQT_SIMT")
    b = p_simt.segments_db(audio, st, va, le, max(lens)).cpu().numpy()
    assert a.shape == b.shape == (len(lens), 96, p_tc.frames(max(lens)))
    for i, n in enumerate(lens):
        T = p_tc.frames(n)
        x, y = a[i, :, :T], b[i, :, :T]
        both = (x > -59.9) & (y > -59.9)
        assert both.sum() > 0 and np.abs(x - y)[both].max() < 0.01, f"clip {i}"
        assert ((x == -120) != (y == -120)).mean() < 2e-3            # elements within rounding of the -60 dB cut
        assert np.array_equal(a[i, :, T:], b[i, :, T:])                # frames past the clip's own length: same filler
    # 16-bit PCM input takes the same path
    pcm = torch.clamp(torch.round(audio * 32768.0), -32768, 32767).to(torch.int16)
    c = p_tc.segments_db(pcm, st, va, le, max(lens)).cpu().numpy()
    d = p_tc.segments_db(pcm.to(torch.float32) / 32768.0, st, va, le, max(lens)).cpu().numpy()
    assert np.array_equal(c, d)
    p_tc.close(); p_simt.close()


def test_resident_toeplitz_operator_equals_streamed_operator(lib, monkeypatch):
    """The decimator GEMM keeps ONE master tile of its banded Toeplitz operator in shared memory and addresses k-block kb
    as rows 16 (nkb-1-kb) .. +128 of it (cqt_gemm_tc.cu: RES); GTC_SCQT_STREAM_OP=1 at plan creation streams the operator's
    own k-blocks through the TMA ring instead.  Same MMAs on the same operands in the same order: bit-identical features,
    on ragged clips at both sample rates (different slot geometries: 16 x 8, 32 x 4 and 64 x 2 TMA boxes)."""
    from gtc_b200 import ops, CqtRecipe
    dev = torch.device("cuda")
    for sr, lens in ((22050.0, [66150, 4410, 3000, 40000, 977] + [5000 + 613 * i for i in range(30)]),
                     (44100.0, [44100, 30001, 256] + [12000 + 977 * i for i in range(15)])):
        r = CqtRecipe(sr=sr)
        clips = [make_test_audio(n, seed=300 + i, sr=sr) for i, n in enumerate(lens)]
        off = np.concatenate([[0], np.cumsum(lens)])
        st, va, le = seg_tables(off[:-1].tolist(), lens, lens, dev)
        audio = torch.from_numpy(np.concatenate(clips)).to(dev)
        p_res = ops.StructuredCqtPlan(r)
        a = p_res.segments_db(audio, st, va, le, max(lens)).cpu().numpy()
        monkeypatch.setenv("GTC_SCQT_STREAM_OP", "1")
        p_str = ops.StructuredCqtPlan(r)
        monkeypatch.delenv("GTC_SCQT_STREAM_OP")
        b = p_str.segments_db(audio, st, va, le, max(lens)).cpu().numpy()
        assert np.isfinite(a).all() and np.array_equal(a, b), f"sr {sr}"
        # the workspace is reused: a second call with SHORTER segments must not see the first call's samples in the slot gaps
        short = [max(64, n // 3) for n in lens]
        st2, va2, le2 = seg_tables(off[:-1].tolist(), short, short, dev)
        c = p_res.segments_db(audio, st2, va2, le2, max(lens)).cpu().numpy()
        d = p_str.segments_db(audio, st2, va2, le2, max(short)).cpu().numpy()
        for i, n in enumerate(short):
            T = p_res.frames(n)
            x, y = c[i, :, :T], d[i, :, :T]
            both = (x > -59.9) & (y > -59.9)
            assert both.sum() > 0 and np.abs(x - y)[both].max() < 1e-3, f"sr {sr} clip {i}: slot gaps"
        p_res.close(); p_str.close()
