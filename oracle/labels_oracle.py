"""CPU oracle: JAMS note events -> (6, 19) int8 tablature labels.  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/jam_to_tablature.py:
  * midi_to_tablature                      :55-109
  * extract_tablature_from_jams            :110-143
  * extract_tablature_from_pitch_contour   :145-178   (librosa.hz_to_midi = 12*(log2(f)-log2(440))+69)
  * the per-segment body of process_file   :269-331   (time grid, fallback rule, stats)
and the label side of my_dataloader.py:38-44 / ViT_dataloader.py:54.

PARITY UNPINNED for real-data semantics (``jams`` is not installable here); the arithmetic itself is plain
Python/NumPy in the reference and is followed line by line, so these loops *are* the reference algorithm.
A ``Jam`` here is the minimal stand-in for ``jams.JAMS``: ``.annotations`` -> objects with ``.namespace`` and
``.data`` (iterable of observations with ``.time .duration .value .confidence``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List

import numpy as np

NUM_STRINGS = 6
NUM_FRETS = 19
OPEN_STRING_PITCHES = [40, 45, 50, 55, 59, 64]           # jam_to_tablature.py:38


@dataclass
class Observation:
    time: float
    duration: float
    value: Any
    confidence: Any = None


@dataclass
class Annotation:
    namespace: str
    data: List[Observation] = field(default_factory=list)


@dataclass
class Jam:
    annotations: List[Annotation] = field(default_factory=list)


def midi_to_tablature(midi_pitches, confidence=None):
    """jam_to_tablature.py:55-109, loop for loop."""
    tablature = np.zeros((NUM_STRINGS, NUM_FRETS), dtype=np.int8)
    if len(midi_pitches) == 0:
        return tablature
    for i, pitch in enumerate(midi_pitches):
        conf = confidence[i] if confidence is not None else 1.0
        if conf < 0.5:
            continue
        if isinstance(pitch, dict):
            if 'pitch' in pitch:
                pitch = pitch['pitch']
            elif 'value' in pitch:
                pitch = pitch['value']
            else:
                continue
        try:
            pitch_value = float(pitch)
        except (ValueError, TypeError):
            continue
        possible_positions = []
        for string_idx, open_pitch in enumerate(OPEN_STRING_PITCHES):
            try:
                fret = int(round(pitch_value - open_pitch))
                if 0 <= fret < NUM_FRETS:
                    possible_positions.append((string_idx, fret))
            except Exception:
                continue
        if possible_positions:
            possible_positions.sort(key=lambda x: x[1])
            string_idx, fret = possible_positions[0]
            tablature[string_idx, fret] = 1
    return tablature


def extract_tablature_from_jams(jam, segment_time):
    """jam_to_tablature.py:110-143."""
    midi_notes = []
    midi_conf = []
    for ann in jam.annotations:
        if ann.namespace == 'note_midi':
            for note in ann.data:
                start_time = note.time
                end_time = start_time + note.duration
                if start_time <= segment_time < end_time:
                    if isinstance(note.value, dict):
                        if 'pitch' in note.value:
                            midi_notes.append(note.value['pitch'])
                        elif 'value' in note.value:
                            midi_notes.append(note.value['value'])
                        else:
                            continue
                    else:
                        midi_notes.append(note.value)
                    midi_conf.append(1.0)
    return midi_to_tablature(midi_notes, midi_conf)


def hz_to_midi(f):
    """librosa.hz_to_midi."""
    return 12 * (np.log2(np.asanyarray(f)) - np.log2(440.0)) + 69


def extract_tablature_from_pitch_contour(jam, segment_time):
    """jam_to_tablature.py:145-178."""
    pitches = []
    confidences = []
    for ann in jam.annotations:
        if ann.namespace == 'pitch_contour':
            for pitch_obs in ann.data:
                if abs(pitch_obs.time - segment_time) < 0.05:
                    pitch_val = None
                    if isinstance(pitch_obs.value, dict):
                        if 'frequency' in pitch_obs.value:
                            pitch_val = pitch_obs.value['frequency']
                        elif 'value' in pitch_obs.value:
                            pitch_val = pitch_obs.value['value']
                    else:
                        pitch_val = pitch_obs.value
                    if pitch_val is not None and pitch_val > 0:
                        try:
                            midi_pitch = hz_to_midi(float(pitch_val))
                            pitches.append(midi_pitch)
                            confidences.append(pitch_obs.confidence)
                        except (ValueError, TypeError):
                            continue
    return midi_to_tablature(pitches, confidences)


def segment_times(duration, num_images):
    """jam_to_tablature.py:273-274."""
    adjusted_segment_duration = duration / num_images
    return [(i + 0.5) * adjusted_segment_duration for i in range(num_images)]


def process_segments(jam, times):
    """Per-segment body of process_file (jam_to_tablature.py:303-331) without file I/O.

    Returns (labels (n,6,19) int8, stats dict).  Exceptions inside a segment are swallowed exactly as
    :314-320 does (the tablature computed so far for that segment is kept).
    """
    stats = {'total': 0, 'with_notes': 0, 'with_first_string': 0}
    out = np.zeros((len(times), NUM_STRINGS, NUM_FRETS), dtype=np.int8)
    for i, t in enumerate(times):
        tablature = np.zeros((NUM_STRINGS, NUM_FRETS), dtype=np.int8)
        try:
            tablature = extract_tablature_from_jams(jam, t)
            if np.sum(tablature) == 0:
                tablature = extract_tablature_from_pitch_contour(jam, t)
        except Exception:
            pass
        out[i] = tablature
        stats['total'] += 1
        if np.sum(tablature) > 0:
            stats['with_notes'] += 1
        if np.sum(tablature[0, :]) > 0:
            stats['with_first_string'] += 1
    return out, stats


def labels_argmax(annotation):
    """my_dataloader.py:40-44: (6,19) -> int64 (6,) class indices (first 1 wins, all-zero row -> 0)."""
    annotation = np.asarray(annotation)
    if len(annotation.shape) == 2 and annotation.shape[1] == 19:
        annotation = np.argmax(annotation, axis=1)
    return annotation.astype(np.int64)


def labels_vit_heads(annotation):
    """ViT_dataloader.py:28,54: (6,19) int8 -> float32 -> six int64 (19,) heads."""
    annotation = np.asarray(annotation).astype(np.float32)
    return [annotation[i].astype(np.int64) for i in range(6)]


# ---- a vectorised restatement used only to check the loops above on large synthetic inputs -------------

def rasterize_events_numpy(onset, dur, pitch, times):
    """Plain-number note events (no dicts) -> (n,6,19) int8; same arithmetic as the loops, vectorised."""
    onset = np.asarray(onset, dtype=np.float64)
    end = onset + np.asarray(dur, dtype=np.float64)
    pitch = np.asarray(pitch, dtype=np.float64)
    times = np.asarray(times, dtype=np.float64)
    out = np.zeros((len(times), NUM_STRINGS, NUM_FRETS), dtype=np.int8)
    if len(onset) == 0 or len(times) == 0:
        return out
    active = (onset[None, :] <= times[:, None]) & (times[:, None] < end[None, :])        # (n_seg, n_evt)
    opens = np.asarray(OPEN_STRING_PITCHES, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        fr = np.rint(pitch[:, None] - opens[None, :])                                      # round-half-even
    valid = np.isfinite(fr) & (fr >= 0) & (fr < NUM_FRETS)
    frk = np.where(valid, fr, np.inf)
    best_s = np.argmin(frk, axis=1)                                                       # first minimum = stable sort
    has = valid.any(axis=1)
    best_f = np.where(has, frk[np.arange(len(pitch)), best_s], 0).astype(np.int64)
    si, ei = np.nonzero(active & has[None, :])
    out[si, best_s[ei], best_f[ei]] = 1
    return out
