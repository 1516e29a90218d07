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
