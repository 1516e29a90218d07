"""CPU oracle: the reference's CQT feature recipe, restated with NumPy/SciPy.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: librosa / soxr are not installable in this image and the reference holds no golden
vectors.  This file restates, step by step and in the same order and dtypes,

* ``librosa.cqt`` -> ``librosa.vqt(gamma=0)``                      (called at /root/reference/cqt.py:55, new_cqt.py:25)
* ``librosa.filters.wavelet`` / ``wavelet_lengths`` / ``util.sparsify_rows``
* ``librosa.stft(window='ones', center=True, pad_mode='constant')``
* ``librosa.resample(orig_sr=2, target_sr=1, res_type='soxr_hq', scale=True)`` -> libsoxr 0.1.3 HQ 2:1 stage
* ``np.abs(C)**4`` ; ``librosa.amplitude_to_db(ref=np.amax)`` ; ``cqt_lim``          (cqt.py:10-13, 56-58)
* the sliding-window driver of ``process_all_audio``                                (cqt.py:26-49)

as published for librosa 0.10.2 / 0.11.0 and libsoxr 0.1.3 (Appendix A of SURVEY.md).  The libsoxr
Kaiser design (``lsx_design_lpf`` / ``lsx_kaiser_beta`` / ``lsx_make_lpf``) is restated from the published
source; its agreement with a real soxr build could not be verified here.
"""
from __future__ import annotations

import math
import numpy as np
import scipy.fft
import scipy.signal

# ----------------------------------------------------------------------------------------------------
# libsoxr 0.1.3 "HQ" decimate-by-2 stage  
# ----------------------------------------------------------------------------------------------------

_KAISER_COEFS = (
    (-6.784957e-10, 1.02856e-05, 0.1087556, -0.8988365 + .001),
    (-6.897885e-10, 1.027433e-05, 0.10876, -0.8994658 + .002),
    (-1.000683e-09, 1.030092e-05, 0.1087677, -0.9007898 + .003),
    (-3.654474e-10, 1.040631e-05, 0.1087085, -0.8977766 + .006),
    (8.106988e-09, 6.983091e-06, 0.1091387, -0.9172048 + .015),
    (9.519571e-09, 7.272678e-06, 0.1090068, -0.9140768 + .025),
    (-5.626821e-09, 1.342186e-05, 0.1083999, -0.9065452 + .05),
    (-9.965946e-08, 5.073548e-05, 0.1040967, -0.7672778 + .085),
    (1.604808e-07, -5.856462e-05, 0.1185998, -1.34824 + .1),
    (-1.511964e-07, 6.363034e-05, 0.1064627, -0.9876665 + .18),
)


def lsx_kaiser_beta(att: float, tr_bw: float) -> float:
    """libsoxr filter.c:lsx_kaiser_beta -- Kaiser beta from attenuation (dB) and relative transition width."""
    if att >= 60:
        realm = math.log(tr_bw / .0005) / math.log(2.)
        i0 = min(max(int(realm), 0), len(_KAISER_COEFS) - 1)
        i1 = min(max(1 + int(realm), 0), len(_KAISER_COEFS) - 1)
        c0, c1 = _KAISER_COEFS[i0], _KAISER_COEFS[i1]
        b0 = ((c0[0] * att + c0[1]) * att + c0[2]) * att + c0[3]
        b1 = ((c1[0] * att + c1[1]) * att + c1[2]) * att + c1[3]
        return b0 + (b1 - b0) * (realm - int(realm))
    if att > 50:
        return .1102 * (att - 8.7)
    if att > 20.96:
        return .58417 * (att - 20.96) ** .4 + .07886 * (att - 20.96)
    return 0.


def soxr_hq_halfband_taps(beta_scale: float = 1.0, taps_delta: int = 0, fc_scale: float = 1.0, rho: float = .5) -> np.ndarray:
    """Taps of the single 2:1 DFT stage libsoxr builds for quality SOXR_HQ, io_ratio 2, linear phase.
    The keyword arguments perturb the design (Kaiser beta x beta_scale, tap count + taps_delta, cut-off x fc_scale, the
    `rho` of the window argument) for scripts/tap_sensitivity.py; the defaults are libsoxr's values.

    soxr.c:soxr_quality_spec -> precision 20 bits, passband_end = 1 - .05/TO_3dB(rej), stopband_begin = 1;
    cr.c:_soxr_init        -> one pre-stage L=1, M=2, att = (20+1)*6.0206 dB, dft_stage_init(Fp, Fs, Fn=2, k=-4)
    filter.c:lsx_design_lpf / lsx_kaiser_params / lsx_make_lpf (rho = .5, scale = 1, no DC normalisation).
    """
    lin2db = math.log10(2.) * 20
    rej = 20 * lin2db
    to3db = (1.6e-6 * rej - 7.5e-4) * rej + .646
    Fp, Fs, Fn = 1 - .05 / to3db, 1.0, 2.0
    att = (20 + 1) * lin2db
    Fp /= Fn
    Fs /= Fn
    tr_bw = .5 * (Fs - Fp)
    tr_bw = min(tr_bw, .5 * Fs)
    Fc = Fs - tr_bw
    beta = lsx_kaiser_beta(att, tr_bw * .5 / Fc)
    a = ((.0007528358 - 1.577737e-05 * beta) * beta + .6248022) * beta + .06186902
    n = int(math.ceil(a / tr_bw + 1))
    modulo = 4                                   # k = -4  ->  num_taps == 1 (mod 4)
    n = (n + modulo - 2) // modulo * modulo + 1
    n += int(taps_delta)
    beta *= beta_scale
    Fc *= fc_scale
    m = n - 1
    i = np.arange(n, dtype=np.float64)
    z = i - .5 * m
    x = z * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        h = np.where(x != 0, np.sin(Fc * x) / x, Fc)
    y = z * (1 / (.5 * m + rho))
    h = h * np.i0(beta * np.sqrt(np.maximum(0.0, 1 - y * y))) / np.i0(beta)
    half = m // 2
    h[m - np.arange(half + 1)] = h[: half + 1]   # lsx_make_lpf computes i <= m/2 and mirrors h[m-i] = h[i]
    return h


_TAPS = None


def halfband_taps() -> np.ndarray:
    global _TAPS
    if _TAPS is None:
        _TAPS = soxr_hq_halfband_taps()
    return _TAPS


class taps_override:
    """``with taps_override(h): ...`` -- evaluate the oracle with another 2:1 decimator (odd length, centre tap in the
    middle): scripts/tap_sensitivity.py (perturbed designs) and scripts/pin_with_librosa.py (taps measured from soxr)."""

    def __init__(self, h):
        self.h = np.asarray(h, dtype=np.float64)
        assert self.h.ndim == 1 and len(self.h) % 2 == 1

    def __enter__(self):
        global _TAPS
        self.prev, _TAPS = _TAPS, self.h
        return self

    def __exit__(self, *a):
        global _TAPS
        _TAPS = self.prev


def resample_2to1(y: np.ndarray) -> np.ndarray:
    """librosa.resample(y, orig_sr=2, target_sr=1, res_type='soxr_hq', scale=True) along the last axis.

    Zero-extended signal, group delay compensated (output k <-> input time 2k), length fixed to
    ceil(n/2) (librosa util.fix_length), then ``y_hat /= sqrt(ratio)`` i.e. * sqrt(2).  float32 in -> float32 out.
    """
    h = halfband_taps()
    c = (len(h) - 1) // 2
    n = y.shape[-1]
    n_out = (n + 1) // 2
    full = scipy.signal.fftconvolve(y.astype(np.float64), h.reshape((1,) * (y.ndim - 1) + (-1,)), axes=-1) \
        if y.ndim > 1 else np.convolve(y.astype(np.float64), h)
    # full[j] = sum_i h[i] x[j-i];  want out[k] = sum_i h[i] x[2k + c - i] = full[2k + c]
    out = full[..., c: c + 2 * n_out: 2]
    out = out / np.sqrt(0.5)
    return np.asarray(out, dtype=y.dtype)


# ----------------------------------------------------------------------------------------------------
# librosa.filters / util pieces  
# ----------------------------------------------------------------------------------------------------

HANN_BANDWIDTH = 1.50018310546875           # librosa.filters.WINDOW_BANDWIDTHS['hann']


def note_to_hz_C(octave: int) -> float:
    """librosa.note_to_hz('C<octave>') = 440 * 2**((midi-69)/12), midi = 12*(octave+1)."""
    midi = 12 * (octave + 1)
    return 440.0 * (2.0 ** ((midi - 69.0) / 12.0))


def relative_bandwidth(freqs: np.ndarray) -> np.ndarray:
    """librosa.filters._relative_bandwidth (0.10.1+)."""
    if len(freqs) <= 1:
        raise ValueError("2 or more frequencies are required to compute bandwidths")
    bpo = np.empty_like(freqs)
    logf = np.log2(freqs)
    bpo[0] = 1 / (logf[1] - logf[0])
    bpo[-1] = 1 / (logf[-1] - logf[-2])
    bpo[1:-1] = 2 / (logf[2:] - logf[:-2])
    alpha = (2.0 ** (2 / bpo) - 1) / (2.0 ** (2 / bpo) + 1)
    return alpha


def wavelet_lengths(freqs, sr, filter_scale=1.0, gamma=0.0, alpha=None):
    """librosa.filters.wavelet_lengths(window='hann')."""
    freqs = np.asarray(freqs, dtype=np.float64)
    alpha = relative_bandwidth(freqs) if alpha is None else np.asarray(alpha)
    Q = float(filter_scale) / alpha
    f_cutoff = np.max(freqs * (1 + 0.5 * HANN_BANDWIDTH / Q) + 0.5 * gamma)
    lengths = Q * sr / (freqs + gamma / alpha)
    return lengths, f_cutoff


def wavelet(freqs, sr, alpha, filter_scale=1.0, norm=1, gamma=0.0):
    """librosa.filters.wavelet(window='hann', pad_fft=True, dtype=complex64)."""
    lengths, _ = wavelet_lengths(freqs, sr, filter_scale, gamma, alpha)
    filts = []
    for ilen, freq in zip(lengths, freqs):
        t = np.arange(-ilen // 2, ilen // 2, dtype=float)
        ang = t * 2 * np.pi * freq / sr
        sig = np.cos(ang) + 1j * np.sin(ang)                       # util.phasor
        sig = sig * scipy.signal.get_window("hann", len(sig), fftbins=True)   # __float_window on an integer n
        assert norm == 1
        sig = sig / np.sum(np.abs(sig))                            # util.normalize(norm=1)
        filts.append(sig)
    max_len = int(2.0 ** (np.ceil(np.log2(max(lengths)))))
    out = np.zeros((len(filts), max_len), dtype=np.complex64)
    for k, f in enumerate(filts):
        lpad = int((max_len - len(f)) // 2)                        # util.pad_center
        out[k, lpad: lpad + len(f)] = f
    return out, lengths


def sparsify_rows(x: np.ndarray, quantile: float = 0.01) -> np.ndarray:
    """librosa.util.sparsify_rows, returned dense (zeros where librosa's CSR matrix has no entry)."""
    mags = np.abs(x)
    norms = np.sum(mags, axis=1, keepdims=True)
    mag_sort = np.sort(mags, axis=1)
    cumulative_mag = np.cumsum(mag_sort / norms, axis=1)
    threshold_idx = np.argmin(cumulative_mag < quantile, axis=1)
    out = np.zeros_like(x)
    for i, j in enumerate(threshold_idx):
        idx = np.where(mags[i] >= mag_sort[i, j])
        out[i, idx] = x[i, idx]
    return out


def vqt_filter_fft(sr, freqs, alpha, filter_scale=1.0, norm=1, sparsity=0.01):
    """librosa.core.constantq.__vqt_filter_fft (hop_length=None as vqt calls it)."""
    basis, lengths = wavelet(freqs, sr, alpha, filter_scale, norm)
    n_fft = basis.shape[1]
    basis *= (lengths[:, np.newaxis] / float(n_fft))
    fft_basis = scipy.fft.fft(basis, n=n_fft, axis=1)[:, : (n_fft // 2) + 1]
    fft_basis = sparsify_rows(fft_basis, quantile=sparsity).astype(np.complex64)
    return fft_basis, n_fft, lengths


def stft_ones(y: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """librosa.stft(window='ones', center=True, pad_mode='constant', dtype=complex64): (..., 1+n_fft//2, frames)."""
    n = y.shape[-1]
    pad = n_fft // 2
    yp = np.zeros(y.shape[:-1] + (n + 2 * pad,), dtype=y.dtype)
    yp[..., pad: pad + n] = y
    n_frames = 1 + n // hop
    idx = (np.arange(n_frames) * hop)[:, None] + np.arange(n_fft)[None, :]
    frames = yp[..., idx]                                          # (..., frames, n_fft)
    D = np.fft.rfft(frames.astype(np.float64), axis=-1)            # window is float64 ones -> double FFT
    return np.swapaxes(D, -1, -2).astype(np.complex64)


def early_downsample_count(nyquist, filter_cutoff, hop_length, n_octaves):
    c1 = max(0, int(np.ceil(np.log2(nyquist / filter_cutoff)) - 1) - 1)
    num_twos = 0
    h = hop_length
    while h % 2 == 0 and h > 0:
        num_twos += 1
        h //= 2
    c2 = max(0, num_twos - n_octaves + 1)
    return min(c1, c2)


# ----------------------------------------------------------------------------------------------------
# librosa.cqt
# ----------------------------------------------------------------------------------------------------

def cqt(y, sr=22050, hop_length=1024, fmin=None, n_bins=96, bins_per_octave=12,
        filter_scale=1.0, norm=1, sparsity=0.01, scale=True, _basis_cache=None):
    """librosa.cqt(y, sr=..., hop_length=..., fmin=..., n_bins=..., bins_per_octave=...) -> complex64 (..., n_bins, T).

    ``y`` float32 (..., n).  Follows librosa.vqt for gamma=0, intervals='equal', tuning=0.0, window='hann',
    pad_mode='constant', res_type='soxr_hq'.  ``_basis_cache`` (dict) lets tests skip the per-call basis rebuild
    librosa performs (its own cache is off unless LIBROSA_CACHE_DIR is set).
    """
    y = np.asarray(y)
    assert y.dtype == np.float32
    n_octaves = int(np.ceil(float(n_bins) / bins_per_octave))
    n_filters = min(bins_per_octave, n_bins)
    if fmin is None:
        fmin = note_to_hz_C(1)
    freqs = fmin * (2.0 ** (np.arange(n_bins, dtype=np.float64) / bins_per_octave))   # interval_frequencies('equal')
    alpha = relative_bandwidth(freqs)
    lengths, filter_cutoff = wavelet_lengths(freqs, sr, filter_scale, 0.0, alpha)
    nyquist = sr / 2.0
    if filter_cutoff > nyquist:
        raise ValueError(f"Wavelet basis with max frequency={np.max(freqs)} would exceed the Nyquist frequency={nyquist}.")
    if early_downsample_count(nyquist, filter_cutoff, hop_length, n_octaves) > 0:
        raise NotImplementedError("early down-sampling (a non 2:1 soxr ratio) is not restated")

    resp = []
    my_y, my_sr, my_hop = y, float(sr), hop_length
    for i in range(n_octaves):
        sl = slice(-n_filters, None) if i == 0 else slice(-n_filters * (i + 1), -n_filters * i)
        key = (my_sr, i)
        if _basis_cache is not None and key in _basis_cache:
            fft_basis, n_fft = _basis_cache[key]
        else:
            fft_basis, n_fft, _ = vqt_filter_fft(my_sr, freqs[sl], alpha[sl], filter_scale, norm, sparsity)
            fft_basis = (fft_basis * np.sqrt(sr / my_sr)).astype(np.complex64)
            if _basis_cache is not None:
                _basis_cache[key] = (fft_basis, n_fft)
        D = stft_ones(my_y, n_fft, my_hop)                          # (..., 65, frames) complex64
        resp.append(np.matmul(fft_basis, D))                       # __cqt_response: fft_basis.dot(D), complex64
        if my_hop % 2 == 0:
            my_hop //= 2
            my_sr /= 2.0
            my_y = resample_2to1(my_y)
    # __trim_stack
    max_col = min(c.shape[-1] for c in resp)
    V = np.empty(y.shape[:-1] + (n_bins, max_col), dtype=np.complex64)
    end = n_bins
    for c in resp:
        n_oct = c.shape[-2]
        if end < n_oct:
            V[..., :end, :] = c[..., -end:, :max_col]
        else:
            V[..., end - n_oct: end, :] = c[..., :max_col]
        end -= n_oct
    if scale:
        V /= np.sqrt(lengths)[:, None]
    return V


# ----------------------------------------------------------------------------------------------------
# |C|**4 -> dB -> cut   (cqt.py:56-58, new_cqt.py:26-30)
# ----------------------------------------------------------------------------------------------------

def amplitude_to_db_amax(S: np.ndarray, amin: float = 1e-5, top_db: float | None = 80.0) -> np.ndarray:
    """librosa.amplitude_to_db(S, ref=np.amax, amin=1e-5, top_db=80.0) on one (n_bins, T) float32 array (A.3)."""
    magnitude = np.abs(S)
    ref_value = np.amax(magnitude)
    power = np.square(magnitude)
    amin2 = amin ** 2
    ref2 = np.abs(ref_value ** 2)
    log_spec = 10.0 * np.log10(np.maximum(amin2, power))
    log_spec -= 10.0 * np.log10(np.maximum(amin2, ref2))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def cqt_lim(CQT: np.ndarray) -> np.ndarray:
    """cqt.py:10-13."""
    new_CQT = np.copy(CQT)
    new_CQT[new_CQT < -60] = -120
    return new_CQT


def segment_features(segment, sr, hop_length=1024, n_bins=96, bins_per_octave=12, fmin=None,
                     _basis_cache=None, return_pre_cut=False):
    """One iteration of the hot loop at cqt.py:55-58: float32 (n,) -> float32 (n_bins, T) dB features."""
    C = cqt(segment, sr=sr, hop_length=hop_length, n_bins=n_bins, bins_per_octave=bins_per_octave, fmin=fmin,
            _basis_cache=_basis_cache)
    mag = np.abs(C) ** 4
    db = amplitude_to_db_amax(mag)
    out = cqt_lim(db)
    if return_pre_cut:
        return out, db, C
    return out


def window_params(sr, window_size=0.2, hop_size=0.1):
    """cqt.py:26-27."""
    return int(window_size * sr), int(hop_size * sr)


def num_segments(n_samples, window_samples, hop_samples):
    """cqt.py:30 (negative -> the range() at :36 is empty)."""
    return max(0, (n_samples - window_samples) // hop_samples + 1)


def process_clip(data, sr, window_size=0.2, hop_size=0.1, _basis_cache=None):
    """The per-clip body of process_all_audio (cqt.py:26-65) without file I/O: list of (n_bins, T) float32."""
    w, h = window_params(sr, window_size, hop_size)
    n = num_segments(len(data), w, h)
    out = []
    for i in range(n):
        s = i * h
        e = s + w
        if e > len(data):
            break
        seg = data[s:e]
        if len(seg) < w:
            continue
        fmin = note_to_hz_C(1) if len(seg) >= 256 else None
        out.append(segment_features(seg, sr, fmin=fmin, _basis_cache=_basis_cache))
    return out
