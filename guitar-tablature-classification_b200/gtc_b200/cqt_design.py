"""Host-side design of the CQT segment operator (float64 linear algebra, done once per recipe).

For a fixed segment length the whole ``librosa.cqt`` call of the reference (/root/reference/cqt.py:55,
new_cqt.py:25) is a linear map of the segment's samples.  This module builds
that map explicitly -- wavelet basis (hann, L1-normalised, sparsified in the FFT domain), rectangular-window
STFT framing with zero padding, the recursive soxr-HQ 2:1 decimation chain and the ``sqrt(length)`` scaling --
as one dense matrix ``A`` of shape ``(2 * n_bins * n_frames, seg_len)``.  libgtc.so then evaluates all segments
of a shard as a single GEMM on the tensor cores.

Row order: ``row = (t * n_bins + bin) * 2 + c`` with ``c = 0`` real part, ``c = 1`` imaginary part.

This is product code: it does not import ``oracle/``.  The oracle restates the same published algorithms
procedurally (per call, float32) and the tests compare the two.
"""
from __future__ import annotations

import hashlib
import math
import os
from dataclasses import dataclass, asdict
from functools import lru_cache

import numpy as np
import scipy.fft
import scipy.sparse

HANN_BANDWIDTH = 1.50018310546875


@dataclass(frozen=True)
class CqtRecipe:
    """Constants of the reference's feature recipe (cqt.py:5,55-58; BASELINE.json configs[0])."""
    sr: float = 22050.0
    window_size: float = 0.2        # cqt.py:5
    hop_size: float = 0.1           # cqt.py:5
    hop_length: int = 1024          # cqt.py:55
    n_bins: int = 96
    bins_per_octave: int = 12
    fmin: float | None = None       # None -> C1 (librosa default; cqt.py:52 passes note_to_hz('C1'))
    filter_scale: float = 1.0
    sparsity: float = 0.01
    power: float = 4.0              # cqt.py:56
    amin: float = 1e-5              # librosa.amplitude_to_db default
    top_db: float = 80.0
    cut_db: float = -60.0           # cqt.py:12
    floor_db: float = -120.0

    @property
    def seg_len(self) -> int:
        return int(self.window_size * self.sr)      # cqt.py:26

    @property
    def seg_hop(self) -> int:
        return int(self.hop_size * self.sr)         # cqt.py:27

    @property
    def fmin_hz(self) -> float:
        return 440.0 * (2.0 ** ((24 - 69.0) / 12.0)) if self.fmin is None else float(self.fmin)

    @property
    def n_octaves(self) -> int:
        return int(math.ceil(self.n_bins / self.bins_per_octave))


# --------------------------------------------------------------------------- soxr HQ 2:1 stage (libsoxr 0.1.3)

def _kaiser_beta(att: float, tr_bw: float) -> float:
    table = (
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
    assert att >= 60
    realm = math.log2(tr_bw / .0005)
    lo = min(max(int(realm), 0), len(table) - 1)
    hi = min(max(int(realm) + 1, 0), len(table) - 1)

    def poly(c):
        return ((c[0] * att + c[1]) * att + c[2]) * att + c[3]

    return poly(table[lo]) + (poly(table[hi]) - poly(table[lo])) * (realm - int(realm))


@lru_cache(maxsize=1)
def decimator_taps() -> np.ndarray:
    """Kaiser-windowed sinc of libsoxr's quality-HQ 2:1 DFT stage (20-bit precision, linear phase); SURVEY.md A.2."""
    db_per_bit = 20 * math.log10(2.)
    rej = 20 * db_per_bit
    passband_end = 1 - .05 / ((1.6e-6 * rej - 7.5e-4) * rej + .646)
    att = 21 * db_per_bit
    fp, fs = passband_end / 2, 0.5                      # normalised to the input Nyquist
    tr_bw = min(.5 * (fs - fp), .5 * fs)
    fc = fs - tr_bw
    beta = _kaiser_beta(att, tr_bw * .5 / fc)
    width = ((.0007528358 - 1.577737e-05 * beta) * beta + .6248022) * beta + .06186902
    n = int(math.ceil(width / tr_bw + 1))
    n = (n + 2) // 4 * 4 + 1                            # num_taps = 1 (mod 4)
    m = n - 1
    half = np.arange(m // 2 + 1, dtype=np.float64)
    z = half - .5 * m
    x = z * math.pi
    sinc = np.where(x != 0, np.sin(fc * x) / np.where(x != 0, x, 1.0), fc)
    win = np.i0(beta * np.sqrt(np.maximum(0.0, 1 - (z / (.5 * m + .5)) ** 2))) / np.i0(beta)
    h = np.empty(n, dtype=np.float64)
    h[: m // 2 + 1] = sinc * win
    h[m - np.arange(m // 2 + 1)] = h[: m // 2 + 1]
    return h


def decimation_matrix(n_in: int) -> scipy.sparse.csr_matrix:
    """(ceil(n_in/2), n_in) matrix of librosa.resample(orig_sr=2, target_sr=1, 'soxr_hq', scale=True):
    zero-extended signal, zero-phase alignment (output k <-> input 2k), gain sqrt(2)."""
    h = decimator_taps() * math.sqrt(2.0)
    c = (len(h) - 1) // 2
    n_out = (n_in + 1) // 2
    k = np.repeat(np.arange(n_out), len(h))
    j = np.tile(np.arange(len(h)), n_out)
    col = 2 * k + c - j                                  # out[k] = sum_j h[j] x[2k + c - j]
    ok = (col >= 0) & (col < n_in)
    vals = np.tile(h, n_out)
    return scipy.sparse.csr_matrix((vals[ok], (k[ok], col[ok])), shape=(n_out, n_in))


# --------------------------------------------------------------------------- wavelet basis (librosa.filters.wavelet)

def _relative_bandwidth(freqs: np.ndarray) -> np.ndarray:
    logf = np.log2(freqs)
    bpo = np.empty_like(freqs)
    bpo[0] = 1 / (logf[1] - logf[0])
    bpo[-1] = 1 / (logf[-1] - logf[-2])
    bpo[1:-1] = 2 / (logf[2:] - logf[:-2])
    r = 2.0 ** (2 / bpo)
    return (r - 1) / (r + 1)


def octave_time_filters(freqs_oct, alpha_oct, octave_sr, sr, filter_scale, sparsity):
    """Effective time-domain filters of one octave: (n_filters, n_fft) complex128 ``W`` such that the
    octave's CQT response at frame t is ``sum_j W[b, j] * ypad[t*hop + j]``  ( = fft_basis @ rfft(frame) )."""
    Q = filter_scale / alpha_oct
    lengths = Q * octave_sr / freqs_oct
    n_fft = int(2.0 ** math.ceil(math.log2(lengths.max())))
    basis = np.zeros((len(freqs_oct), n_fft), dtype=np.complex64)
    for b, (ilen, f) in enumerate(zip(lengths, freqs_oct)):
        t = np.arange(-ilen // 2, ilen // 2, dtype=float)
        n = len(t)
        win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / n)          # periodic hann
        sig = np.exp(1j * (t * 2 * np.pi * f / octave_sr)) * win
        sig = sig / np.abs(sig).sum()
        lpad = (n_fft - n) // 2
        basis[b, lpad: lpad + n] = sig
    basis *= (lengths[:, None] / float(n_fft))
    fb = scipy.fft.fft(basis, n=n_fft, axis=1)[:, : n_fft // 2 + 1]
    # librosa.util.sparsify_rows(quantile=sparsity): drop the smallest entries holding < `sparsity` of the row's L1 mass
    mags = np.abs(fb)
    srt = np.sort(mags, axis=1)
    cum = np.cumsum(srt / mags.sum(axis=1, keepdims=True), axis=1)
    thr = srt[np.arange(len(fb)), np.argmin(cum < sparsity, axis=1)]
    fb = np.where(mags >= thr[:, None], fb, 0).astype(np.complex64)
    fb = (fb * np.sqrt(sr / octave_sr)).astype(np.complex64)
    k = np.arange(n_fft // 2 + 1)[:, None] * np.arange(n_fft)[None, :]
    dft = np.exp(-2j * np.pi * k / n_fft)                               # rfft matrix (65, 128)
    return fb.astype(np.complex128) @ dft, n_fft


def n_frames_of(recipe: CqtRecipe, seg_len: int | None = None) -> int:
    n = recipe.seg_len if seg_len is None else seg_len
    hop, frames = recipe.hop_length, []
    for i in range(recipe.n_octaves):
        frames.append(1 + n // hop)
        if hop % 2 == 0:
            hop //= 2
            n = (n + 1) // 2
    return min(frames)


def build_operator(recipe: CqtRecipe, seg_len: int | None = None) -> np.ndarray:
    """Dense float32 operator (2*n_bins*T, seg_len), row = (t*n_bins + bin)*2 + {re, im}."""
    seg_len = recipe.seg_len if seg_len is None else int(seg_len)
    n_bins, bpo, n_oct = recipe.n_bins, recipe.bins_per_octave, recipe.n_octaves
    sr = float(recipe.sr)
    if recipe.hop_length % (2 ** (n_oct - 1)) != 0:
        raise ValueError(f"hop_length must be a positive integer multiple of 2^{n_oct - 1} for {n_oct}-octave CQT")
    freqs = recipe.fmin_hz * (2.0 ** (np.arange(n_bins, dtype=np.float64) / bpo))
    alpha = _relative_bandwidth(freqs)
    lengths = (recipe.filter_scale / alpha) * sr / freqs
    cutoff = np.max(freqs * (1 + 0.5 * HANN_BANDWIDTH / (recipe.filter_scale / alpha)))
    if cutoff > sr / 2:
        raise ValueError(f"Wavelet basis with max frequency={freqs.max()} would exceed the Nyquist frequency={sr / 2}")
    num_twos = (recipe.hop_length & -recipe.hop_length).bit_length() - 1
    early = min(max(0, int(math.ceil(math.log2((sr / 2) / cutoff)) - 1) - 1), max(0, num_twos - n_oct + 1))
    if early > 0:
        raise NotImplementedError("recipes that trigger librosa's early down-sampling are not supported")

    T = n_frames_of(recipe, seg_len)
    A = np.zeros((T, n_bins, 2, seg_len), dtype=np.float64)
    n_filters = min(bpo, n_bins)
    Y = None                                         # (len_i, seg_len): input -> octave-i signal; None = identity
    cur_len, hop, octave_sr = seg_len, recipe.hop_length, sr
    for i in range(n_oct):
        hi = n_bins - n_filters * i
        lo = max(0, hi - n_filters)
        W, n_fft = octave_time_filters(freqs[lo:hi], alpha[lo:hi], octave_sr, sr, recipe.filter_scale, recipe.sparsity)
        W = W / np.sqrt(lengths[lo:hi])[:, None]
        for t in range(T):
            start = t * hop - n_fft // 2             # first signal index under the frame (centre padding)
            j0, j1 = max(0, -start), min(n_fft, cur_len - start)
            if j1 <= j0:
                continue
            Wt = W[:, j0:j1]
            if Y is None:
                A[t, lo:hi, 0, start + j0: start + j1] = Wt.real
                A[t, lo:hi, 1, start + j0: start + j1] = Wt.imag
            else:
                rows = Y[start + j0: start + j1]
                A[t, lo:hi, 0] = Wt.real @ rows
                A[t, lo:hi, 1] = Wt.imag @ rows
        if hop % 2 == 0 and i + 1 < n_oct:
            D = decimation_matrix(cur_len)
            Y = D.toarray() if Y is None else D @ Y
            cur_len, hop, octave_sr = (cur_len + 1) // 2, hop // 2, octave_sr / 2.0
    return np.ascontiguousarray(A.reshape(T * n_bins * 2, seg_len).astype(np.float32))


def _recipe_grid(recipe: CqtRecipe):
    """freqs, alpha, full-rate lengths of a recipe, with librosa.vqt's own argument checks."""
    n_bins, bpo, n_oct = recipe.n_bins, recipe.bins_per_octave, recipe.n_octaves
    sr = float(recipe.sr)
    if recipe.hop_length % (2 ** (n_oct - 1)) != 0:
        raise ValueError(f"hop_length must be a positive integer multiple of 2^{n_oct - 1} for {n_oct}-octave CQT")
    freqs = recipe.fmin_hz * (2.0 ** (np.arange(n_bins, dtype=np.float64) / bpo))
    alpha = _relative_bandwidth(freqs)
    lengths = (recipe.filter_scale / alpha) * sr / freqs
    cutoff = np.max(freqs * (1 + 0.5 * HANN_BANDWIDTH / (recipe.filter_scale / alpha)))
    if cutoff > sr / 2:
        raise ValueError(f"Wavelet basis with max frequency={freqs.max()} would exceed the Nyquist frequency={sr / 2}")
    num_twos = (recipe.hop_length & -recipe.hop_length).bit_length() - 1
    early = min(max(0, int(math.ceil(math.log2((sr / 2) / cutoff)) - 1) - 1), max(0, num_twos - n_oct + 1))
    if early > 0:
        raise NotImplementedError("recipes that trigger librosa's early down-sampling are not supported")
    return freqs, alpha, lengths


def structured_filters(recipe: CqtRecipe):
    """Tables of the structured (multirate) evaluation, libgtc's gtc_scqt_plan_create inputs:
    ``filters`` float32 [n_octaves, 2*filters_per_octave, n_fft] (row = filter*2 + {re, im}; octave 0 = top octave; every
    scaling of librosa.vqt folded in: lengths/n_fft, sqrt(sr/octave_sr), 1/sqrt(full-rate length)), ``n_fft`` and the
    2:1 decimator ``taps`` float32 with librosa.resample's sqrt(2) gain folded in.
    Octaves whose own n_fft is smaller are zero-padded symmetrically, which leaves the centred frames unchanged."""
    freqs, alpha, lengths = _recipe_grid(recipe)
    n_bins, n_oct, sr = recipe.n_bins, recipe.n_octaves, float(recipe.sr)
    n_filters = min(recipe.bins_per_octave, n_bins)
    per_oct, octave_sr = [], sr
    for i in range(n_oct):
        hi = n_bins - n_filters * i
        lo = max(0, hi - n_filters)
        W, n_fft = octave_time_filters(freqs[lo:hi], alpha[lo:hi], octave_sr, sr, recipe.filter_scale, recipe.sparsity)
        per_oct.append((W / np.sqrt(lengths[lo:hi])[:, None], n_fft))
        octave_sr /= 2.0
    n_fft_max = max(n for _, n in per_oct)
    filters = np.zeros((n_oct, 2 * n_filters, n_fft_max), dtype=np.float32)
    for i, (W, n_fft) in enumerate(per_oct):
        pad = (n_fft_max - n_fft) // 2
        filters[i, 0: 2 * len(W): 2, pad: pad + n_fft] = W.real
        filters[i, 1: 2 * len(W): 2, pad: pad + n_fft] = W.imag
    taps = (decimator_taps() * math.sqrt(2.0)).astype(np.float32)
    return filters, n_fft_max, taps


def _cache_dir() -> str:
    d = os.environ.get("GTC_CACHE_DIR", os.path.join(os.path.expanduser("~"), ".cache", "gtc_b200"))
    os.makedirs(d, exist_ok=True)
    return d


_MEM_CACHE: dict = {}


def get_operator(recipe: CqtRecipe, seg_len: int | None = None) -> np.ndarray:
    """build_operator with an in-process and an on-disk cache (the design takes a few seconds)."""
    seg_len = recipe.seg_len if seg_len is None else int(seg_len)
    measured = os.environ.get("GTC_OPERATOR_FILE")
    if measured:
        # an operator measured from the real librosa.cqt (scripts/pin_with_librosa.py --operator-out); used only where
        # its shape fits the recipe, so inference / new_cqt plans of other geometries keep the designed operator
        key = ("file", measured, seg_len, recipe.n_bins)
        if key not in _MEM_CACHE:
            op = np.load(measured)
            _MEM_CACHE[key] = op if op.shape == (2 * recipe.n_bins * n_frames_of(recipe, seg_len), seg_len) else None
        if _MEM_CACHE[key] is not None:
            return _MEM_CACHE[key]
    design = {k: v for k, v in asdict(recipe).items() if k in
              ("sr", "hop_length", "n_bins", "bins_per_octave", "filter_scale", "sparsity")}
    design.update(fmin=recipe.fmin_hz, seg_len=seg_len, v=3)
    key = hashlib.sha1(repr(sorted(design.items())).encode()).hexdigest()[:16]
    if key in _MEM_CACHE:
        return _MEM_CACHE[key]
    path = os.path.join(_cache_dir(), f"segop_{key}.npy")
    op = None
    if os.path.exists(path):
        try:
            op = np.load(path)
        except Exception:
            op = None
    if op is None:
        op = build_operator(recipe, seg_len)
        tmp = f"{path}.{os.getpid()}.tmp.npy"
        np.save(tmp, op)
        os.replace(tmp, path)
    _MEM_CACHE[key] = op
    return op
