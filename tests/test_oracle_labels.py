"""KATs for the label oracle and loop-vs-vectorised agreement."""
import numpy as np

from oracle import labels_oracle as lo
from oracle.labels_oracle import Annotation, Jam, Observation


def tab_of(*positions):
    t = np.zeros((6, 19), np.int8)
    for s, f in positions:
        t[s, f] = 1
    return t


def test_round_half_even_and_range():
    assert np.array_equal(lo.midi_to_tablature([40.5]), tab_of((0, 0)))      # round(0.5) = 0
    assert np.array_equal(lo.midi_to_tablature([41.5]), tab_of((0, 2)))      # round(1.5) = 2
    assert np.array_equal(lo.midi_to_tablature([39.5]), tab_of((0, 0)))      # round(-0.5) = -0 -> 0
    assert np.array_equal(lo.midi_to_tablature([82.5]), tab_of((5, 18)))     # 18.5 -> 18 kept
    assert lo.midi_to_tablature([82.51]).sum() == 0                          # 19 -> invalid everywhere
    assert lo.midi_to_tablature([39.4]).sum() == 0
    assert lo.midi_to_tablature([float('nan')]).sum() == 0
    assert lo.midi_to_tablature([float('inf')]).sum() == 0


def test_lowest_fret_rule_and_multi_hot():
    assert np.array_equal(lo.midi_to_tablature([64]), tab_of((5, 0)))        # open high e, not fret 24.. of low E
    assert np.array_equal(lo.midi_to_tablature([45]), tab_of((1, 0)))
    assert np.array_equal(lo.midi_to_tablature([41, 43]), tab_of((0, 1), (0, 3)))   # two 1s in one row
    assert np.array_equal(lo.midi_to_tablature([{'pitch': 50.2}, {'value': 57}, {'x': 1}, 'abc', '52']),
                          tab_of((2, 0), (3, 2), (2, 2)))
    assert lo.midi_to_tablature([50], [0.49]).sum() == 0
    assert lo.midi_to_tablature([50], [0.5]).sum() == 1


def test_interval_is_half_open():
    jam = Jam([Annotation('note_midi', [Observation(1.0, 0.5, 45.0)])])
    assert lo.extract_tablature_from_jams(jam, 1.0).sum() == 1                # t == onset included
    assert lo.extract_tablature_from_jams(jam, 1.5).sum() == 0                # t == onset + duration excluded
    assert lo.extract_tablature_from_jams(jam, 1.4999999).sum() == 1
    other = Jam([Annotation('pitch_contour', [Observation(1.0, 0.5, 45.0)])])
    assert lo.extract_tablature_from_jams(other, 1.2).sum() == 0              # namespace filter


def test_contour_fallback_and_none_confidence():
    jam = Jam([Annotation('note_midi', []),
               Annotation('pitch_contour', [Observation(1.00, 0, {'frequency': 110.0}, 0.9),
                                            Observation(1.04, 0, {'frequency': 0.0}, 0.9),
                                            Observation(1.049, 0, 220.0, 0.4),
                                            Observation(1.2, 0, 440.0, 0.9)])])
    labels, stats = lo.process_segments(jam, [1.0, 1.2, 3.0])
    assert np.array_equal(labels[0], tab_of((1, 0)))                          # 110 Hz = A2 = MIDI 45
    assert np.array_equal(labels[1], tab_of((5, 5)))                          # 440 Hz = MIDI 69 -> string 5 fret 5
    assert labels[2].sum() == 0
    assert stats == {'total': 3, 'with_notes': 2, 'with_first_string': 0}
    bad = Jam([Annotation('pitch_contour', [Observation(1.0, 0, 110.0, None)])])
    labels, stats = lo.process_segments(bad, [1.0])
    assert labels.sum() == 0 and stats['with_notes'] == 0                     # TypeError swallowed, zeros kept


def test_segment_times():
    t = lo.segment_times(30.0, 150)
    assert len(t) == 150 and t[0] == 0.1 and abs(t[-1] - 29.9) < 1e-12


def test_vectorised_matches_loops():
    rng = np.random.default_rng(5)
    for trial in range(5):
        n = int(rng.integers(0, 200))
        onset = rng.uniform(0, 10, n)
        dur = rng.uniform(0.05, 2.0, n)
        pitch = rng.uniform(35, 90, n)
        half = rng.random(n) < 0.2
        pitch[half] = np.round(pitch[half] * 2) / 2              # exercise the .5 rounding boundary
        times = lo.segment_times(10.0, 50)
        jam = Jam([Annotation('note_midi', [Observation(o, d, p) for o, d, p in zip(onset, dur, pitch)])])
        want = np.stack([lo.extract_tablature_from_jams(jam, t) for t in times])
        got = lo.rasterize_events_numpy(onset, dur, pitch, times)
        assert np.array_equal(got, want)


def test_label_views():
    tab = tab_of((0, 3), (0, 7), (4, 1))
    assert list(lo.labels_argmax(tab)) == [3, 0, 0, 0, 1, 0]
    heads = lo.labels_vit_heads(tab)
    assert len(heads) == 6 and heads[0].dtype == np.int64 and heads[0].sum() == 2
