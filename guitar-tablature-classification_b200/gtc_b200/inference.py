"""Inference-time feature front-ends of the reference, on the structured CQT path.

* ``TabCnnFrontEnd``  -- /root/reference/tablature_generator.py:599-666 (``audio_to_cqt_image`` numeric part +
  ``segment_audio``): audio at 22 050 Hz, 3 s segments with 50 % overlap and a zero-padded tail, each segment
  ``librosa.cqt(hop 512, fmin C2, 84 bins)`` -> ``amplitude_to_db(np.abs(C), ref=np.max)`` -> (84, 130) float32.
  The reference then *draws* that array with matplotlib and feeds the PNG to the CNN; rendering is out of scope
   -- this module stops at the array ``specshow`` receives.
* ``vit_preprocess``   -- "/root/reference/tablature-generator (1).py":282-340 (``preprocess_audio``): 44.1 kHz audio,
  0.2 s windows every 0.1 s, the cqt.py recipe, ``cqt_lim``, ``(x+120)/120`` clipped; and ``prepare_for_vit`` :349-372
  up to the HuggingFace processor call (bicubic 224 x 224, 3 channels).
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np
import torch

from . import audio_io, ops
from .cqt_design import CqtRecipe

C2_HZ = 440.0 * 2.0 ** ((36 - 69) / 12.0)          # librosa.note_to_hz('C2'), tablature_generator.py:616


def segment_table(n_samples: int, segment_length: int, hop_length: int) -> Tuple[np.ndarray, np.ndarray]:
    """tablature_generator.py:655-664: starts ``range(0, n, hop)``; every segment is padded to ``segment_length``."""
    starts = np.arange(0, n_samples, hop_length, dtype=np.int64)
    valid = np.minimum(segment_length, n_samples - starts).astype(np.int32)
    return starts, valid


class TabCnnFrontEnd:
    """Feature side of TablatureImageGenerator (tablature_generator.py:474-666)."""

    def __init__(self, sr: int = 22050, hop_length: int = 512, device: int | None = None):
        self.sr, self.hop_length = int(sr), int(hop_length)
        self.recipe = CqtRecipe(sr=float(sr), hop_length=hop_length, n_bins=84, bins_per_octave=12, fmin=C2_HZ,
                                power=1.0, cut_db=-math.inf)                       # :616-620, no cqt_lim on this path
        self.plan = ops.StructuredCqtPlan(self.recipe, device=device)
        self.device = torch.device("cuda", self.plan.device)

    def load(self, audio_file) -> np.ndarray:
        """librosa.load(audio_file, sr=self.sr) (:613,650) for files at sr or 2*sr (the 2:1 soxr-HQ stage on the device)."""
        y, native = audio_io.load_wav(audio_file)
        if native == self.sr:
            return y
        if native == 2 * self.sr:
            return self.plan.halve_rate(torch.from_numpy(y).to(self.device)).cpu().numpy()
        raise ValueError(f"{audio_file}: {native} Hz -> {self.sr} Hz is not a 2:1 ratio; resample the file first")

    def segment_audio(self, audio_file, segment_duration=3.0, sr=None, overlap=0.5):
        """Same return value as the reference's method (:637-666): ([(segment float32, start seconds), ...], sr)."""
        y = self.load(audio_file) if isinstance(audio_file, (str, bytes)) or hasattr(audio_file, "__fspath__") else np.asarray(audio_file, np.float32)
        sr = self.sr if sr is None else int(sr)
        segment_length = int(segment_duration * sr)
        hop = int(segment_length * (1 - overlap))
        starts, _ = segment_table(len(y), segment_length, hop)
        segments = []
        for s in starts:
            seg = y[s: s + segment_length]
            if len(seg) < segment_length:
                seg = np.pad(seg, (0, segment_length - len(seg)))
            segments.append((seg, s / sr))
        return segments, sr

    def cqt_db_segments(self, y, segment_duration=3.0, overlap=0.5, pcm16_roundtrip: bool = True):
        """What audio_to_cqt_image hands to specshow for every segment of generate_tablature_from_mp3 (:871-884):
        (features float32 device tensor [n_seg, 84, T], start times in seconds).
        ``pcm16_roundtrip`` reproduces the temp-file hop of :878-882 (``sf.write`` stores PCM_16, ``librosa.load`` reads it
        back): samples become int16 on the host and are converted x/32768 on the device.  [3P] libsndfile scales
        float -> int16 by 32767 and int16 -> float by 1/32768; restated, not verified against a real soundfile build."""
        y = np.ascontiguousarray(y, dtype=np.float32)
        segment_length = int(segment_duration * self.sr)
        hop = int(segment_length * (1 - overlap))
        starts, valid = segment_table(len(y), segment_length, hop)
        if len(starts) == 0:
            return torch.zeros((0, 84, self.plan.frames(segment_length)), device=self.device), starts / self.sr
        if pcm16_roundtrip:
            q = np.clip(np.rint(y.astype(np.float64) * 32767.0), -32768, 32767).astype(np.int16)
            audio = torch.from_numpy(q).to(self.device)
        else:
            audio = torch.from_numpy(y).to(self.device)
        st = torch.from_numpy(starts).to(self.device)
        va = torch.from_numpy(valid).to(self.device)
        le = torch.full((len(starts),), segment_length, dtype=torch.int32, device=self.device)
        db = self.plan.segments_db(audio, st, va, le, segment_length)
        return db, starts / self.sr

    def audio_to_cqt_db(self, audio_file):
        """Numeric part of audio_to_cqt_image (:612-620) for one file: (84, T) float32 of the whole file."""
        y = self.load(audio_file)
        n = len(y)
        audio = torch.from_numpy(y).to(self.device)
        st = torch.zeros(1, dtype=torch.int64, device=self.device)
        ln = torch.full((1,), n, dtype=torch.int32, device=self.device)
        return self.plan.segments_db(audio, st, ln, ln, n)[0].cpu().numpy()


def vit_window_table(n_samples: int, sr: int, segment_duration=0.2, hop_duration=0.1):
    """"tablature-generator (1).py":299-323: (starts, valid) of the windows it keeps."""
    segment_length = int(segment_duration * sr)
    hop_length = int(hop_duration * sr)
    num_segments = max(1, int((n_samples - segment_length) / hop_length) + 1)
    starts, valid = [], []
    for i in range(num_segments):
        s = i * hop_length
        e = min(s + segment_length, n_samples)
        if e - s < segment_length // 2:
            continue
        starts.append(s)
        valid.append(e - s)
    return np.asarray(starts, np.int64), np.asarray(valid, np.int32), segment_length


_VIT_PLANS: dict = {}


def vit_preprocess(data: np.ndarray, sr: int = 44100, segment_duration=0.2, hop_duration=0.1):
    """preprocess_audio (:282-340) on already-loaded mono audio at ``sr``: returns (normalised CQT segments as one
    device tensor [n, 96, T] in [0, 1], timestamps list)."""
    from . import augment
    data = np.ascontiguousarray(data, dtype=np.float32)
    starts, valid, segment_length = vit_window_table(len(data), sr, segment_duration, hop_duration)
    key = (int(sr), torch.cuda.current_device())
    plan = _VIT_PLANS.get(key)
    if plan is None:
        plan = _VIT_PLANS[key] = ops.StructuredCqtPlan(CqtRecipe(sr=float(sr)))
    dev = torch.device("cuda", plan.device)
    if len(starts) == 0:
        return torch.zeros((0, 96, plan.frames(segment_length)), device=dev), []
    audio = torch.from_numpy(data).to(dev) if len(data) else torch.zeros(1, device=dev)
    le = torch.full((len(starts),), segment_length, dtype=torch.int32, device=dev)
    db = plan.segments_db(audio, torch.from_numpy(starts).to(dev), torch.from_numpy(valid).to(dev), le, segment_length)
    return augment.db_normalize(db, -120.0), [float(s) / sr for s in starts]


def prepare_for_vit(normalised: torch.Tensor) -> torch.Tensor:
    """prepare_for_vit (:349-368) before the HuggingFace processor call: [n, 96, T] in [0,1] -> [n, 3, 224, 224] bicubic."""
    from . import _lib
    return ops.patches(normalised.contiguous(), mode=_lib.GTC_PATCH_VIT_PRENORM)
