"""WAV decode and .npy writers with the semantics the reference gets from librosa / numpy.

* ``load_wav(path)``      == ``librosa.load(path, sr=None, mono=True)`` for PCM/float WAV files: float32 in [-1, 1),
  channels averaged (cqt.py:23, new_cqt.py:22; GuitarSet hex-pickup files are 6-channel).
* ``wav_duration(path)``  == ``librosa.get_duration(y=y, sr=sr)`` without decoding the samples (jam_to_tablature.py:269-270).
* ``save_feature``        writes what ``np.save(path, new_CQT)`` writes at cqt.py:63: '<f4', shape (n_bins, T),
  Fortran order (librosa's __trim_stack allocates order="F" and every later op preserves it).
* ``save_label``          writes the (6, 19) '|i1' C-order file of jam_to_tablature.py:323-324 (242 bytes).
* packed files            one ``.npy`` per clip holding all its segments ([n_seg, n_bins, T] '<f4' or [n_seg, 6, 19] '|i1')
  instead of ~10 files per second of audio; ``explode_features`` / ``explode_labels`` turn a
  packed file back into exactly the per-segment files the reference writes, ``packed_item_names`` gives the names those
  files would have so that loaders keep the reference's sorted-listdir pairing order.
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


_HEADER_CACHE: dict = {}


def _npy_header(shape, dtype, fortran: bool) -> bytes:
    """The exact header bytes np.save writes for an array of this shape / dtype / order (format 1.0)."""
    key = (tuple(shape), np.dtype(dtype).str, bool(fortran))
    h = _HEADER_CACHE.get(key)
    if h is None:
        import io
        buf = io.BytesIO()
        np.save(buf, np.zeros(shape, dtype=dtype, order="F" if fortran else "C"))
        raw = buf.getvalue()
        h = raw[: len(raw) - int(np.prod(shape)) * np.dtype(dtype).itemsize]
        _HEADER_CACHE[key] = h
    return h


def save_features_exploded(out_dir, base_name: str, feats: np.ndarray, first: int = 0) -> int:
    """The per-segment files of cqt.py:61-63 for one clip -- ``{base}_segment_{k}.npy``, (n_bins, T) '<f4' Fortran order --
    byte for byte what ``save_feature`` / ``np.save`` writes, without building 10 array objects and headers per second of
    audio: one transposed copy of the clip's block, one cached header, one write per file."""
    import os
    feats = np.asarray(feats, dtype=np.float32)
    if feats.ndim != 3 or len(feats) == 0:
        return 0
    n, nb, T = feats.shape
    header = _npy_header((nb, T), np.float32, True)
    data = np.ascontiguousarray(feats.transpose(0, 2, 1))            # C order of (T, n_bins) == Fortran order of (n_bins, T)
    raw = memoryview(data).cast("B")
    step = nb * T * 4
    prefix = os.path.join(out_dir, base_name + "_segment_")
    flags = os.O_WRONLY | os.O_CREAT | os.O_TRUNC
    for k in range(n):                                               # open + one gathered write + close: 13 us per file on tmpfs
        fd = os.open(f"{prefix}{first + k}.npy", flags, 0o666)
        try:
            os.writev(fd, (header, raw[k * step: (k + 1) * step]))
        finally:
            os.close(fd)
    return n


# ---------------------------------------------------------------------------------------------------- packed files

FEATURE_PACK_SUFFIX = "_segments.npy"      # {base}_segments.npy       <->  {base}_segment_{k}.npy      (cqt.py:62)
LABEL_PACK_SUFFIX = "_tabs.npy"            # {base}_tabs.npy           <->  {base}/{base}_{i:04d}.npy   (jam_to_tablature.py:323)


def save_features_packed(path, feats: np.ndarray) -> None:
    """[n_seg, n_bins, T] float32, C order: item k is the array cqt.py:63 would save as segment k."""
    feats = np.ascontiguousarray(feats, dtype=np.float32)
    assert feats.ndim == 3
    np.save(path, feats)


def save_labels_packed(path, tabs: np.ndarray, indices=None) -> None:
    """[n, 6, 19] int8; ``indices`` (the segment numbers kept, jam_to_tablature.py:303-309) travel in a side file."""
    tabs = np.ascontiguousarray(tabs, dtype=np.int8)
    assert tabs.ndim == 3 and tabs.shape[1:] == (6, 19)
    np.save(path, tabs)
    if indices is not None:
        np.save(str(path)[: -len(".npy")] + ".idx.npy", np.asarray(indices, dtype=np.int64))


def packed_item_names(packed_name: str, n: int, indices=None):
    """File names the reference would have written for the items of a packed file."""
    import os
    name = os.path.basename(str(packed_name))
    if name.endswith(FEATURE_PACK_SUFFIX):
        base = name[: -len(FEATURE_PACK_SUFFIX)]
        return [f"{base}_segment_{k}.npy" for k in range(n)]
    if name.endswith(LABEL_PACK_SUFFIX):
        base = name[: -len(LABEL_PACK_SUFFIX)]
        idx = range(n) if indices is None else indices
        return [f"{base}_{int(i):04d}.npy" for i in idx]
    raise ValueError(f"{packed_name} is not a packed feature/label file")


def explode_features(packed_path, out_dir) -> int:
    """Write the reference's per-segment feature files (Fortran order, cqt.py:62-63) from a packed file."""
    import os
    arr = np.load(packed_path)
    os.makedirs(out_dir, exist_ok=True)
    for name, a in zip(packed_item_names(packed_path, len(arr)), arr):
        save_feature(os.path.join(out_dir, name), a)
    return len(arr)


def explode_labels(packed_path, out_dir) -> int:
    """Write {out_dir}/{base}/{base}_{i:04d}.npy (jam_to_tablature.py:280,323-324) from a packed label file."""
    import os
    arr = np.load(packed_path)
    idx_path = str(packed_path)[: -len(".npy")] + ".idx.npy"
    idx = np.load(idx_path) if os.path.exists(idx_path) else None
    base = os.path.basename(str(packed_path))[: -len(LABEL_PACK_SUFFIX)]
    os.makedirs(os.path.join(out_dir, base), exist_ok=True)
    for name, a in zip(packed_item_names(packed_path, len(arr), idx), arr):
        save_label(os.path.join(out_dir, base, name), a)
    return len(arr)
