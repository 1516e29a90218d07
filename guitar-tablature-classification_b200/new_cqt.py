"""Drop-in for the reference's ``new_cqt.py``: per-offset CQT pictures of every audio file, GPU-batched.

    audio_CQT_parallel(file_num, start, dur=0.2)                       (reference new_cqt.py:8)
    process_all_files_parallel(start, dur=0.2, max_images=45000)       (reference new_cqt.py:46)

Kept from the reference: the window is the file's frames [int(start*sr), +int(dur*sr)) (librosa.load offset/duration,
new_cqt.py:22); the CQT is designed for ``sr=44100`` regardless of the file (new_cqt.py:25); ``|C|**4`` -> dB(ref=max)
-> ``< -60 -> -120`` (new_cqt.py:26-30); output name ``{audio_name}_segment_{file_num}_{start:.2f}`` (new_cqt.py:40);
``max_images // n_files`` windows per file at offsets ``start + j*dur`` (new_cqt.py:50-58).
Changed on purpose: the hard-coded Windows directory becomes ``AUDIO_DIR`` / ``OUTPUT_DIR`` (module attributes or the
GTC_AUDIO_DIR / GTC_CQT_IMAGES_DIR environment variables); ``os.listdir`` is sorted so file_num is stable
(SURVEY.md 8g.13); the ProcessPoolExecutor fan-out is replaced by one GEMM per file batch; besides the picture a
``.npy`` with the numeric (96, T) dB array is written, because the matplotlib rendering (colormap, figure geometry) is
not a numeric contract -- the PNG here is a plain grey image, top row = highest bin.
"""
from __future__ import annotations

import os

import numpy as np

from gtc_b200 import audio_io, features
from gtc_b200.cqt_design import CqtRecipe

AUDIO_DIR = os.environ.get("GTC_AUDIO_DIR", "audio_hex-pickup_debleeded")
OUTPUT_DIR = os.environ.get("GTC_CQT_IMAGES_DIR", "cqt_images")
CQT_SR = 44100.0            # literal at new_cqt.py:25


def _audio_files():
    return sorted(f for f in os.listdir(AUDIO_DIR) if f.lower().endswith('.wav'))


def _windows(path, offsets, dur):
    """Frames [int(off*sr), +int(dur*sr)) for each offset; incomplete windows are dropped (and reported)."""
    n_frames, sr = audio_io.wav_info(path)
    y, _ = audio_io.load_wav(path)
    w = int(dur * sr)
    segs, kept = [], []
    for off in offsets:
        s = int(off * sr)
        if w > 0 and s + w <= n_frames:
            segs.append(y[s:s + w])
            kept.append(off)
    return segs, kept, w


def _save_picture(path_noext, db):
    audio_io.save_feature(path_noext + ".npy", db)
    try:
        from PIL import Image
        grey = np.clip((db[::-1] + 120.0) / 120.0, 0.0, 1.0)
        Image.fromarray((grey * 255.0 + 0.5).astype(np.uint8)).save(path_noext + ".png")
    except ImportError:
        pass


def picture_stems(audio_name, file_num, offsets):
    """Output names of one file's windows, new_cqt.py:40: ``{audio_name}_segment_{file_num}_{start:.2f}`` -- the names
    jam_to_tablature.py later gives its label files (tests/test_reference_golden.py checks them against the 43 188 label
    files the reference repository ships)."""
    return [f"{audio_name}_segment_{file_num}_{off:.2f}" for off in offsets]


def window_offsets(start, dur, max_images, total_files):
    """Window starts of every file, new_cqt.py:53-61: ``max_images // total_files`` windows, ``dur`` apart."""
    return [start + j * dur for j in range(max_images // total_files)]


def _process_file(file_num, offsets, dur):
    files = _audio_files()
    name = files[file_num]
    audio_name = os.path.splitext(name)[0]
    segs, kept, w = _windows(os.path.join(AUDIO_DIR, name), offsets, dur)
    if len(kept) < len(offsets):
        print(f"{name}: {len(offsets) - len(kept)} window(s) run past the end of the file and were skipped")
    if not segs:
        return 0
    recipe = CqtRecipe(sr=CQT_SR, window_size=dur, hop_size=dur)
    feats = features.clips_features(segs, recipe, seg_len=w, seg_hop=w)       # every window is its own one-segment clip
    os.makedirs(OUTPUT_DIR, exist_ok=True)
    for stem, f in zip(picture_stems(audio_name, file_num, kept), feats):
        out = os.path.join(OUTPUT_DIR, stem)
        _save_picture(out, f[0])
        print(f"Saved: {out}.png")
    return len(kept)


def audio_CQT_parallel(file_num, start, dur=0.2):  # start and dur in seconds
    return _process_file(file_num, [start], dur)


def process_all_files_parallel(start, dur=0.2, max_images=45000):
    total_files = len(_audio_files())
    if total_files == 0:
        print(f"No .wav files in {AUDIO_DIR}")
        return 0
    offsets = window_offsets(start, dur, max_images, total_files)
    done = 0
    for i in range(total_files):
        done += _process_file(i, offsets, dur)
    return done


def main():
    process_all_files_parallel(start=0, dur=0.2, max_images=45000)


if __name__ == "__main__":
    main()
