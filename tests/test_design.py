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


def emulate_structured(y, filters, n_fft, taps, hop, n_bins):
    """What libgtc's structured kernels compute (csrc/cqt_structured.cu), in float64 NumPy from the float32 tables."""
    n_oct, n_real, _ = filters.shape
    nf = n_real // 2
    c = (len(taps) - 1) // 2
    resp, x, frames = [], y.astype(np.float64), []
    for i in range(n_oct):
        T = 1 + len(x) // hop
        frames.append(T)
        xp = np.concatenate([np.zeros(n_fft // 2), x, np.zeros(n_fft // 2 + hop)])
        F = np.stack([xp[t * hop: t * hop + n_fft] for t in range(T)])            # (T, n_fft)
        R = F @ filters[i].astype(np.float64).T                                   # (T, 2*nf)
        resp.append((R[:, 0::2] + 1j * R[:, 1::2]).T)                             # (nf, T)
        if hop % 2 == 0:
            full = np.convolve(x, taps.astype(np.float64))
            x = full[c: c + 2 * ((len(x) + 1) // 2): 2]
            hop //= 2
    T = min(frames)
    V = np.zeros((n_bins, T), dtype=np.complex128)
    for i, Ri in enumerate(resp):
        hi = n_bins - nf * i
        lo = max(0, hi - nf)
        V[lo:hi] = Ri[: hi - lo, :T]
    return V


@pytest.mark.parametrize("kw,n", [
    (dict(), 4410),                                                               # cqt.py:55 at 22.05 kHz
    (dict(sr=44100.0), 8820),                                                     # GuitarSet native rate, n_fft 256
    (dict(hop_length=512, n_bins=84, fmin=65.40639132514966), 22050),             # tablature_generator.py:616-617
    (dict(n_bins=90), 5000),                                                      # partial lowest octave
])
def test_structured_tables_reproduce_oracle_cqt(kw, n):
    r = cd.CqtRecipe(**kw)
    filters, n_fft, taps = cd.structured_filters(r)
    x = make_test_audio(n, seed=3, sr=r.sr)
    C = co.cqt(x, sr=r.sr, hop_length=r.hop_length, fmin=r.fmin_hz, n_bins=r.n_bins)
    V = emulate_structured(x, filters, n_fft, taps, r.hop_length, r.n_bins)
    assert V.shape == C.shape
    assert np.abs(V - C).max() < 3e-6 * np.abs(C).max()


def test_operator_zero_structure_at_mma_granularity():
    """The numbers DESIGN.md 3.1 quotes for the zero-skipping schedule of the tensor-core engine: fill of the segment
    operator per (N tile of 24 bins) x (k-block of 32 samples) x (frame), in the library's frame-major tile order."""
    from gtc_b200.cqt_design import CqtRecipe, get_operator
    A = get_operator(CqtRecipe())
    T, NB = 5, 96
    nz = (A.reshape(T, NB, 2, 4410) != 0)
    assert abs(nz.mean() - 0.634) < 0.005

    def kblocks(mask):                                   # two audio rows of 2205 samples, each padded to 69 k-blocks of 32
        out = []
        for p in range(2):
            m = np.zeros(2208, bool)
            m[:2205] = mask[p * 2205:(p + 1) * 2205]
            out.append(m.reshape(69, 32).any(1))
        return np.concatenate(out)

    fill = []
    for tile in range(4):                                # library tile j = bins 24j .. 24j+23 (tile 3 = the two top octaves)
        bins = slice(24 * tile, 24 * tile + 24)
        fill.append(np.mean([kblocks(nz[t, bins].any(axis=(0, 1))).mean() for t in range(T)]))
    assert fill[0] == 1.0 and fill[1] == 1.0             # low octaves: filters longer than the segment
    assert abs(fill[2] - 0.635) < 0.01 and abs(fill[3] - 0.139) < 0.01
    assert abs(np.mean(fill) - 0.694) < 0.01             # 31 % of the tensor work multiplied zeros
    # whole-k-block skipping alone would reach far less: the union over a tile's five frames covers most of the segment
    assert kblocks(nz[:, 72:96].any(axis=(0, 1, 2))).mean() > 0.69


def test_toeplitz_decimator_operator_structure():
    """The structure the resident-operator decimator GEMM relies on (csrc/cqt_structured.cu builds the operator, csrc/cqt_gemm_tc.cu
    `RES` / `band_*` use it), restated on the product's tap table:
      * Op[n][i] = h[2n + c + left - i] is Toeplitz: k-block kb + 1 (32 samples) is k-block kb moved down 16 rows;
      * one master tile of 128 + 16 (nkb - 1) rows, filled by 128-row boxes of k-blocks nkb-1, nkb-9, ... and, once those run out,
        of k-block 0 from row r - 16 (nkb - 1) on, holds every k-block kb at rows 16 (nkb - 1 - kb) .. + 128;
      * per k-block the non-zero rows are one contiguous range of 16-row groups, a full k-block exists (it initialises the
        accumulator), and the band is 120 of the 176 group-blocks (the tensor work the schedule keeps);
      * the GEMM of overlapping windows against it is the decimator: y[k] = sum_m h[m] x[2k + c - m]."""
    h = cd.decimator_taps().astype(np.float32)
    taps, c = len(h), (len(h) - 1) // 2
    E, sh, rows = 32, 16, 128
    left = -(-c // 32) * 32
    K = -(-(left + 254 + c + 1) // 32) * 32
    assert (taps, c, left, K) == (389, 194, 224, 704)
    n, i = np.arange(rows)[:, None], np.arange(K)[None, :]
    m = 2 * n + c + left - i
    op = np.where((m >= 0) & (m < taps), h[np.clip(m, 0, taps - 1)], 0.0).astype(np.float32)
    nkb = K // E
    assert np.array_equal(op[sh:, E:], op[:-sh, :-E])                                   # the shift property
    master = np.full((rows + sh * (nkb - 1) + rows, E), np.nan, np.float32)               # box-sized slack at the end
    need = rows + sh * (nkb - 1)
    for r in range(0, need, rows):
        kb = nkb - 1 - r // sh
        if kb >= 0:
            master[r:r + rows] = op[:, kb * E:(kb + 1) * E]
        else:
            n0 = r - sh * (nkb - 1)
            master[r:r + rows - n0] = op[n0:, :E]
            master[r + rows - n0:r + rows] = 0.0                                           # TMA zero-fills rows past the operator
    assert not np.isnan(master[:need]).any()
    for kb in range(nkb):
        at = sh * (nkb - 1 - kb)
        assert np.array_equal(master[at:at + rows], op[:, kb * E:(kb + 1) * E]), kb
    work, full = 0, 0
    for kb in range(nkb):
        nz = np.flatnonzero(np.any(op[:, kb * E:(kb + 1) * E] != 0, axis=1))
        groups = np.unique(nz // 16)
        assert len(groups) and np.array_equal(groups, np.arange(groups[0], groups[-1] + 1))   # contiguous
        work += len(groups)
        full += len(groups) == rows // 16
    assert full >= 1 and work == 120 and nkb * (rows // 16) == 176
    # windows x operator == the 2:1 decimator (zero-extended signal, group delay c)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(3000)
    xz = np.concatenate([np.zeros(left), x, np.zeros(K)])
    y = np.array([op.astype(np.float64) @ xz[256 * j: 256 * j + K] for j in range(4)]).reshape(-1)
    ref = np.array([sum(h[mm] * (x[2 * k + c - mm] if 0 <= 2 * k + c - mm < len(x) else 0.0) for mm in range(taps)) for k in range(0, 512, 37)])
    assert np.abs(y[0:512:37] - ref).max() < 1e-5 * np.abs(ref).max()
