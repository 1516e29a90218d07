"""Host-side marshalling of JAMS annotations into the flat fp64 event arrays libgtc's rasteriser consumes.

The object model mirrors what /root/reference/jam_to_tablature.py touches of ``jams.JAMS``:
``jam.annotations`` -> items with ``.namespace`` and ``.data`` (iterable of observations with
``.time .duration .value .confidence``).  A real ``jams.JAMS`` object works unchanged; ``load_jams`` parses the
``.jams`` JSON directly so the ``jams`` package is not required.

Value handling follows the reference line by line:
  * note values   : dict -> 'pitch' | 'value' | skip (jam_to_tablature.py:128-136), then ``float()`` or skip (:74-89)
  * contour values: dict -> 'frequency' | 'value' (:159-165), ``> 0`` (:168), hz_to_midi(float()) or skip (:169-175)
  * a ``None`` confidence makes ``conf < 0.5`` raise inside midi_to_tablature (:70); process_file swallows the
    exception and keeps zeros (:319-320).  Such observations are flagged kind=1 ("poison").
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Any, List, Sequence

import numpy as np


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
    duration: float | None = None


def load_jams(path) -> Jam:
    """Minimal .jams (JSON) reader: supports the list-of-observations layout (JAMS >= 0.3) and the
    dict-of-lists layout of older files."""
    with open(path, "r") as fh:
        doc = json.load(fh)
    jam = Jam(duration=(doc.get("file_metadata") or {}).get("duration"))
    for ann in doc.get("annotations", []):
        data = ann.get("data", [])
        obs = []
        if isinstance(data, dict):
            n = len(data.get("time", []))
            for i in range(n):
                obs.append(Observation(data["time"][i], data.get("duration", [0.0] * n)[i], data["value"][i],
                                       data.get("confidence", [None] * n)[i]))
        else:
            for d in data:
                obs.append(Observation(d.get("time"), d.get("duration", 0.0), d.get("value"), d.get("confidence")))
        jam.annotations.append(Annotation(ann.get("namespace", ""), obs))
    return jam


def _resolve_pitch(value):
    """Nested dict handling of :128-136 and :74-82, then float() (:85-89).  Returns float or None (skip)."""
    for _ in range(2):
        if isinstance(value, dict):
            if 'pitch' in value:
                value = value['pitch']
            elif 'value' in value:
                value = value['value']
            else:
                return None
    try:
        return float(value)
    except (ValueError, TypeError):
        return None


def marshal_notes(jam):
    """All ``note_midi`` observations of a clip -> (onset, dur, pitch) fp64 arrays (annotation order kept)."""
    on, du, pi = [], [], []
    for ann in jam.annotations:
        if ann.namespace == 'note_midi':
            for note in ann.data:
                p = _resolve_pitch(note.value)
                if p is None:
                    continue
                on.append(float(note.time)); du.append(float(note.duration)); pi.append(p)
    return (np.asarray(on, dtype=np.float64), np.asarray(du, dtype=np.float64), np.asarray(pi, dtype=np.float64))


def hz_to_midi(f):
    """librosa.hz_to_midi, computed on the host in fp64 so the device never evaluates a transcendental."""
    return 12 * (np.log2(np.asanyarray(f, dtype=np.float64)) - np.log2(440.0)) + 69


def marshal_contours(jam):
    """All ``pitch_contour`` observations -> (time, midi, conf fp64, kind int8)."""
    tt, mm, cc, kk = [], [], [], []
    for ann in jam.annotations:
        if ann.namespace == 'pitch_contour':
            for o in ann.data:
                v = o.value
                if isinstance(v, dict):
                    v = v['frequency'] if 'frequency' in v else (v['value'] if 'value' in v else None)
                if v is None:
                    continue
                try:
                    if not (v > 0):
                        continue
                except TypeError:
                    tt.append(float(o.time)); mm.append(np.nan); cc.append(np.nan); kk.append(1)
                    continue
                try:
                    midi = float(hz_to_midi(float(v)))
                except (ValueError, TypeError):
                    continue
                conf = o.confidence
                poison = False
                try:
                    _ = conf < 0.5
                    conf = float(conf)
                except TypeError:
                    poison = True
                tt.append(float(o.time)); mm.append(midi); cc.append(np.nan if poison else conf); kk.append(1 if poison else 0)
    return (np.asarray(tt, dtype=np.float64), np.asarray(mm, dtype=np.float64), np.asarray(cc, dtype=np.float64),
            np.asarray(kk, dtype=np.int8))


def pack_clips(per_clip: Sequence[Sequence[np.ndarray]]):
    """[(a0, b0, ...), (a1, b1, ...)] per clip -> concatenated arrays + offsets int64 [n_clips+1]."""
    n_fields = len(per_clip[0]) if per_clip else 0
    off = np.zeros(len(per_clip) + 1, dtype=np.int64)
    for i, fields in enumerate(per_clip):
        off[i + 1] = off[i] + len(fields[0])
    cat = [np.concatenate([c[k] for c in per_clip]) if per_clip else np.zeros(0) for k in range(n_fields)]
    return cat, off


def segment_times(duration: float, num_images: int) -> np.ndarray:
    """jam_to_tablature.py:273-274 in fp64: (i + 0.5) * (duration / num_images)."""
    adj = duration / num_images
    return (np.arange(num_images, dtype=np.float64) + 0.5) * adj
