"""GPU parity: label rasterisation through the C ABI, bit-exact against the reference's loops (oracle)."""
import numpy as np
import pytest
import torch

from oracle import labels_oracle as lo
from oracle.labels_oracle import Annotation, Jam, Observation

pytestmark = pytest.mark.gpu


def gpu_rasterize(jams, times_per_clip):
    from gtc_b200 import events, ops
    dev = torch.device("cuda")
    notes = [events.marshal_notes(j) for j in jams]
    cons = [events.marshal_contours(j) for j in jams]
    (on, du, pi), eoff = events.pack_clips(notes)
    (ct, cm, cc, ck), coff = events.pack_clips(cons)
    soff = np.concatenate([[0], np.cumsum([len(t) for t in times_per_clip])]).astype(np.int64)
    times = np.concatenate(times_per_clip).astype(np.float64) if len(times_per_clip) else np.zeros(0)
    t_ = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    tabs, stats = ops.rasterize_tabs(t_(on, np.float64), t_(du, np.float64), t_(pi, np.float64), t_(eoff, np.int64),
                                     t_(times, np.float64), t_(soff, np.int64),
                                     contour=(t_(ct, np.float64), t_(cm, np.float64), t_(cc, np.float64), t_(ck, np.int8),
                                              t_(coff, np.int64)))
    return tabs.cpu().numpy(), stats.cpu().numpy()


def oracle_rasterize(jams, times_per_clip):
    labs, tot = [], {'total': 0, 'with_notes': 0, 'with_first_string': 0}
    for j, t in zip(jams, times_per_clip):
        l, s = lo.process_segments(j, list(t))
        labs.append(l)
        for k in tot:
            tot[k] += s[k]
    return np.concatenate(labs), np.array([tot['total'], tot['with_notes'], tot['with_first_string']])


def random_jam(rng, duration, dict_frac=0.05, with_contour=True):
    notes = []
    for s in range(6):
        for _ in range(rng.poisson(3 * duration)):
            p = [40, 45, 50, 55, 59, 64][s] + int(rng.integers(0, 19)) + rng.normal(0, 0.15)
            if rng.random() < 0.1:
                p = np.round(p * 2) / 2
            v = {'pitch': p} if rng.random() < dict_frac else ({'value': p} if rng.random() < dict_frac else p)
            notes.append(Observation(rng.uniform(0, duration), float(np.clip(rng.exponential(0.4), 0.05, 4)), v))
    anns = [Annotation('note_midi', notes[: len(notes) // 2]), Annotation('tempo', [Observation(0, duration, 120.0)]),
            Annotation('note_midi', notes[len(notes) // 2:])]
    if with_contour:
        obs = []
        for k in range(int(duration * 100)):
            f = 440.0 * 2 ** ((rng.uniform(38, 84) - 69) / 12) if rng.random() > 0.3 else 0.0
            v = {'frequency': f, 'voiced': f > 0} if rng.random() < 0.5 else f
            obs.append(Observation(k / 100.0, 0.0, v, float(rng.random())))
        anns.append(Annotation('pitch_contour', obs))
    return Jam(anns)


def test_random_clips_bit_exact(lib):
    rng = np.random.default_rng(2)
    durs = [22.3, 14.6, 30.0, 0.5]
    jams = [random_jam(rng, d) for d in durs]
    times = [np.asarray(lo.segment_times(d, max(1, int(d / 0.2)))) for d in durs]
    got, gstats = gpu_rasterize(jams, times)
    want, wstats = oracle_rasterize(jams, times)
    assert got.dtype == np.int8 and got.shape == want.shape
    assert np.array_equal(got, want)
    assert np.array_equal(gstats, wstats)


def test_known_answers(lib):
    notes = [Observation(1.0, 0.5, 45.0), Observation(2.0, 1.0, 40.5), Observation(2.0, 1.0, 41.5),
             Observation(4.0, 1.0, 39.5), Observation(5.0, 1.0, 82.5), Observation(6.0, 1.0, 82.51),
             Observation(7.0, 1.0, {'pitch': 50.2}), Observation(7.0, 1.0, {'nope': 1}), Observation(7.0, 1.0, 'abc'),
             Observation(8.0, 1.0, float('nan')), Observation(8.0, 1.0, float('inf')), Observation(8.0, 1.0, 1e300),
             Observation(9.0, 1.0, 41.0), Observation(9.2, 1.0, 43.0), Observation(10.0, 0.0, 45.0)]
    jam = Jam([Annotation('note_midi', notes)])
    t = np.array([1.0, 1.5, 1.4999999, 0.9999999, 2.5, 4.5, 5.5, 6.5, 7.5, 8.5, 9.5, 10.0, 20.0])
    got, gstats = gpu_rasterize([jam], [t])
    want, wstats = oracle_rasterize([jam], [t])
    assert np.array_equal(got, want) and np.array_equal(gstats, wstats)
    assert got[0, 1, 0] == 1 and got[1].sum() == 0 and got[2].sum() == 1 and got[3].sum() == 0
    assert got[4, 0, 0] == 1 and got[4, 0, 2] == 1          # 40.5 -> 0 ; 41.5 -> 2 (half-to-even), same string
    assert got[6, 5, 18] == 1 and got[7].sum() == 0
    assert got[10, 0, 1] == 1 and got[10, 0, 3] == 1


def test_contour_fallback_and_poison(lib):
    con = [Observation(1.00, 0, {'frequency': 110.0}, 0.9), Observation(1.04, 0, 0.0, 0.9),
           Observation(1.049, 0, 220.0, 0.4), Observation(1.2, 0, 440.0, 0.9), Observation(3.0, 0, 330.0, None),
           Observation(3.01, 0, 330.0, 0.9), Observation(5.0, 0, 'x', 0.9), Observation(6.0, 0, 196.0, float('nan'))]
    jam = Jam([Annotation('note_midi', [Observation(1.15, 0.1, 64.0)]), Annotation('pitch_contour', con)])
    t = np.array([1.0, 1.2, 3.0, 4.0, 5.0, 6.0])
    got, gstats = gpu_rasterize([jam], [t])
    want, wstats = oracle_rasterize([jam], [t])
    assert np.array_equal(got, want) and np.array_equal(gstats, wstats)
    assert got[0, 1, 0] == 1                                  # fallback used
    assert got[1, 5, 0] == 1 and got[1].sum() == 1            # a note is active -> no fallback
    assert got[2].sum() == 0                                  # None confidence -> exception swallowed -> zeros


def test_empty_inputs(lib):
    got, gstats = gpu_rasterize([Jam([])], [np.array([0.1, 0.3])])
    assert got.shape == (2, 6, 19) and got.sum() == 0 and list(gstats) == [2, 0, 0]
    got, gstats = gpu_rasterize([Jam([])], [np.zeros(0)])
    assert got.shape == (0, 6, 19)


def test_large_batch_property(lib):
    """BASELINE-size property run: 360 clips; vectorised oracle on plain-number events; stats equal serial sums."""
    from gtc_b200 import ops, synth
    dev = torch.device("cuda")
    rng = np.random.default_rng(1)
    durs = rng.uniform(14.6, 30.0, 360)
    on, du, pi, eoff = synth.note_events(durs, seed=2)
    times = [np.asarray(lo.segment_times(d, int(d / 0.2))) for d in durs]
    soff = np.concatenate([[0], np.cumsum([len(t) for t in times])]).astype(np.int64)
    t_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tabs, stats = ops.rasterize_tabs(t_(on), t_(du), t_(pi), t_(eoff), t_(np.concatenate(times)), t_(soff))
    want = np.concatenate([lo.rasterize_events_numpy(on[eoff[c]:eoff[c + 1]], du[eoff[c]:eoff[c + 1]], pi[eoff[c]:eoff[c + 1]],
                                                      times[c]) for c in range(360)])
    got = tabs.cpu().numpy()
    assert np.array_equal(got, want)
    s = stats.cpu().numpy()
    assert s[0] == len(want) and s[1] == (want.sum(axis=(1, 2)) > 0).sum() and s[2] == (want[:, 0].sum(axis=1) > 0).sum()


def test_label_views(lib):
    from gtc_b200 import ops
    rng = np.random.default_rng(0)
    tabs = (rng.random((257, 6, 19)) < 0.08).astype(np.int8)
    d = torch.from_numpy(tabs).cuda()
    assert np.array_equal(ops.labels_argmax(d).cpu().numpy(), np.stack([lo.labels_argmax(t) for t in tabs]))
    idx = torch.from_numpy(rng.permutation(257)[:50].astype(np.int64)).cuda()
    assert np.array_equal(ops.labels_argmax(d, idx).cpu().numpy(), np.stack([lo.labels_argmax(tabs[i]) for i in idx.cpu()]))
    heads = ops.labels_vit_heads(d, idx)
    assert len(heads) == 6 and heads[0].shape == (50, 19) and heads[0].dtype == torch.int64
    for s in range(6):
        assert np.array_equal(heads[s].cpu().numpy(), tabs[idx.cpu().numpy(), s].astype(np.int64))
