"""Known-answer tests that pin the CQT oracle -- there are no reference golden vectors."""
import numpy as np
import pytest

from oracle import cqt_oracle as co

SR = 22050


def test_soxr_design_constants():
    h = co.halfband_taps()
    assert len(h) % 4 == 1 and 380 <= len(h) <= 400          # ~385-393 taps, == 1 (mod 4)
    assert np.allclose(h, h[::-1])                            # linear phase
    assert abs(h.sum() - 1.0) < 1e-6                          # unity DC gain (not re-normalised by libsoxr)
    H = np.abs(np.fft.rfft(h, 1 << 16))
    f = np.arange(len(H)) / (len(H) - 1)                      # 1.0 = input Nyquist
    assert np.abs(H[f <= 0.4568] - 1).max() < 1e-5            # pass-band up to 0.91363 of the new Nyquist
    assert 20 * np.log10(H[f >= 0.5].max()) < -115            # full rejection from the new Nyquist on


def test_basis_matches_survey_counts(basis_cache):
    freqs = co.note_to_hz_C(1) * 2.0 ** (np.arange(96) / 12)
    alpha = co.relative_bandwidth(freqs)
    assert np.allclose(alpha, 0.057698, atol=1e-6)
    lengths, cutoff = co.wavelet_lengths(freqs, SR, alpha=alpha)
    assert abs(lengths[0] - 11685.8) < 0.1 and abs(lengths[-1] - 48.36) < 0.01 and abs(cutoff - 8244.1) < 0.1
    fb, n_fft, _ = co.vqt_filter_fft(float(SR), freqs[84:], alpha[84:])
    assert n_fft == 128
    assert list((fb != 0).sum(1)) == [9, 10, 10, 10, 11, 11, 13, 13, 14, 14, 15, 15]
    fb2, _, _ = co.vqt_filter_fft(SR / 2.0, freqs[72:84], alpha[72:84])
    assert np.abs(fb - fb2).max() < 1e-6                       # identical for every octave


def test_pure_tone_peaks_at_its_bin(basis_cache):
    t = np.arange(SR) / SR
    y = np.sin(2 * np.pi * 440.0 * t).astype(np.float32)
    C = co.cqt(y, sr=SR, _basis_cache=basis_cache)
    assert C.shape == (96, 1 + SR // 1024) and C.dtype == np.complex64
    mid = np.abs(C)[:, 5:15]
    assert (mid.argmax(0) == 45).all()
    assert np.allclose(mid[45], 14.73, atol=0.02)             # 0.5 * sqrt(Q * sr / f)


def test_linearity(basis_cache):
    rng = np.random.default_rng(0)
    a = rng.standard_normal(4410).astype(np.float32) * 0.1
    b = rng.standard_normal(4410).astype(np.float32) * 0.1
    Ca, Cb = co.cqt(a, sr=SR, _basis_cache=basis_cache), co.cqt(b, sr=SR, _basis_cache=basis_cache)
    Cab = co.cqt((a + 2 * b).astype(np.float32), sr=SR, _basis_cache=basis_cache)
    assert Cab.shape == (96, 5)
    assert np.abs(Cab - (Ca + 2 * Cb)).max() < 2e-5 * np.abs(Cab).max()


def test_db_edge_cases():
    # silence -> every element 0 dB, survives the cut (A.3)
    z = np.zeros((96, 5), np.float32)
    assert (co.cqt_lim(co.amplitude_to_db_amax(z)) == 0).all()
    # quiet: peak |C|^4 < amin -> all 0 dB
    q = np.full((96, 5), 1e-6, np.float32)
    assert (co.amplitude_to_db_amax(q) == 0).all()
    # amin floor visible: peak in (amin, ...) and zeros elsewhere -> floor = 20*log10(amin/ref) in (-60, 0)
    s = np.zeros((96, 5), np.float32)
    s[3, 2] = 1e-3
    d = co.amplitude_to_db_amax(s)
    assert d[3, 2] == 0 and np.allclose(d[0, 0], 20 * np.log10(1e-5 / 1e-3), atol=1e-4)
    assert co.cqt_lim(d)[0, 0] == pytest.approx(-40.0, abs=1e-4)
    # top_db clamp then cut: anything more than 60 dB down becomes -120
    s[0, 0] = 1e-7
    s[3, 2] = 10.0
    assert co.cqt_lim(co.amplitude_to_db_amax(s))[0, 0] == -120


def test_segment_recipe_shapes_and_values(basis_cache):
    rng = np.random.default_rng(1)
    seg = (0.2 * rng.standard_normal(4410)).astype(np.float32)
    f = co.segment_features(seg, SR, fmin=co.note_to_hz_C(1), _basis_cache=basis_cache)
    assert f.shape == (96, 5) and f.dtype == np.float32
    assert f.max() == 0.0 and set(np.unique(f[f < -60])) <= {-120.0}


def test_window_arithmetic():
    w, h = co.window_params(22050)
    assert (w, h) == (4410, 2205)
    assert co.num_segments(661500, w, h) == 299
    assert co.num_segments(4409, w, h) == 0 and co.num_segments(4410, w, h) == 1
    assert co.window_params(44100) == (8820, 4410)


def test_decimator_meets_the_published_soxr_hq_specification():
    """Spec-level check of the restated 2:1 stage, independent of HOW its taps are computed: libsoxr's documented HQ quality is
    20-bit precision (soxr.h: SOXR_HQ "20-bit"), linear phase, pass-band end 0.913 of the new Nyquist (1 - 0.05 / TO_3dB), stop-band
    from the new Nyquist on; its DFT stage is designed for (20 + 1) x 6.02 = 126.4 dB.  The frequency response of the
    table shared by the product and the oracle must show exactly that (it cannot tell WHICH Kaiser-class filter meets the
    spec -- profiles/r02_tap_sensitivity.md bounds that freedom)."""
    from gtc_b200 import cqt_design
    h = co.soxr_hq_halfband_taps()
    assert np.abs(h - cqt_design.decimator_taps()).max() < 1e-16             # one table for product and checker (4e-18 apart)
    assert len(h) % 4 == 1 and np.abs(h - h[::-1]).max() == 0.0              # linear phase, libsoxr's tap-count rule
    n = 1 << 18
    H = np.abs(np.fft.rfft(h, n))
    f = np.arange(len(H)) / (n / 2)                                           # in units of the INPUT Nyquist
    lin2db = 20 * np.log10(2.0)
    passband_end = (1 - 0.05 / ((1.6e-6 * 20 * lin2db - 7.5e-4) * 20 * lin2db + 0.646)) / 2
    assert abs(passband_end - 0.45682) < 1e-5
    assert np.abs(20 * np.log10(H[f <= passband_end])).max() < 1e-5         # flat to 1e-5 dB over the pass-band
    assert 20 * np.log10(H[f >= 0.5].max()) < -(20 + 1) * lin2db             # >= 126.4 dB from the new Nyquist on: no aliasing
    assert abs(f[np.argmin(np.abs(H - 0.5))] - 0.47841) < 2e-4               # -6 dB at Fc = Fs - tr_bw
    assert abs(H[0] - 1.0) < 1e-7
