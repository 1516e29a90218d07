"""WAV decode and .npy writers with the semantics the reference gets from librosa / numpy.

* ``load_wav(path)``      == ``librosa.load(path, sr=None, mono=True)`` for PCM/float WAV files: float32 in [-1, 1),
  channels averaged (cqt.py:23, new_cqt.py:22; GuitarSet hex-pickup files are 6-channel).
* ``wav_duration(path)``  == ``librosa.get_duration(y=y, sr=sr)`` without decoding the samples (jam_to_tablature.py:269-270).
* ``save_feature``        writes what ``np.save(path, new_CQT)`` writes at cqt.py:63: '<f4', shape (n_bins, T),
  Fortran order (librosa's __trim_stack allocates order="F" and every later op preserves it).
* ``save_label``          writes the (6, 19) '|i1' C-order file of jam_to_tablature.py:323-324 (242 bytes).
"""
from __future__ import annotations

import warnings

import numpy as np
import scipy.io.wavfile


def _to_float32(x: np.ndarray) -> np.ndarray:
    if x.dtype == np.int16:
        return x.astype(np.float32) / np.float32(32768.0)
    if x.dtype == np.int32:                       # 24-bit files arrive left-justified in int32
        return (x.astype(np.float64) / 2147483648.0).astype(np.float32)
    if x.dtype == np.uint8:
        return (x.astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
    if x.dtype in (np.float32, np.float64):
        return x.astype(np.float32)
    raise ValueError(f"unsupported WAV sample type {x.dtype}")


def load_wav(path, offset: float = 0.0, duration: float | None = None):
    """(y float32 mono, sr).  offset/duration in seconds select frames [int(offset*sr), +int(duration*sr))."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", scipy.io.wavfile.WavFileWarning)
        sr, data = scipy.io.wavfile.read(path, mmap=True)
    start = int(offset * sr) if offset else 0
    stop = data.shape[0] if duration is None else min(data.shape[0], start + int(duration * sr))
    y = _to_float32(np.asarray(data[start:stop]))
    if y.ndim > 1:
        y = np.mean(y, axis=1, dtype=np.float32)
    return np.ascontiguousarray(y, dtype=np.float32), int(sr)


def wav_info(path):
    """(n_frames, sr) from the header."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", scipy.io.wavfile.WavFileWarning)
        sr, data = scipy.io.wavfile.read(path, mmap=True)
    return int(data.shape[0]), int(sr)


def wav_duration(path) -> float:
    n, sr = wav_info(path)
    return float(n) / float(sr)


def write_wav_pcm16(path, y: np.ndarray, sr: int) -> None:
    """Test/bench helper: float [-1,1) -> 16-bit PCM WAV."""
    q = np.clip(np.round(np.asarray(y, dtype=np.float64) * 32768.0), -32768, 32767).astype(np.int16)
    scipy.io.wavfile.write(path, int(sr), q)


def save_feature(path, feat: np.ndarray) -> None:
    np.save(path, np.asfortranarray(feat, dtype=np.float32))


def save_label(path, tab: np.ndarray) -> None:
    np.save(path, np.ascontiguousarray(tab, dtype=np.int8))
