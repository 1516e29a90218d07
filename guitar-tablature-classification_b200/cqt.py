"""Drop-in for the reference's ``cqt.py``: same function, same signature, same ``.npy`` files -- computed on a B200.

    process_all_audio(dataset_path, window_size=0.2, hop_size=0.1, save_path='output')      (reference cqt.py:5)

Differences that are deliberate: importing this module does not start processing (the reference runs
with hard-coded D:\\ paths at import, cqt.py:69-72); all complete windows of a batch of files are evaluated by one
tensor-core GEMM instead of one librosa.cqt call per window.  File naming ({base}_segment_{k}.npy with an un-padded running
counter, cqt.py:62), window arithmetic (cqt.py:26-30), the per-file native sample rate (cqt.py:23) and the array layout
((96, T) float32, Fortran order) are the reference's.
"""
from __future__ import annotations

import os

import numpy as np

from gtc_b200 import audio_io, features
from gtc_b200.cqt_design import CqtRecipe


def _load_batch(dataset_path, names):
    loaded = []
    for name in names:
        try:
            y, sr = audio_io.load_wav(os.path.join(dataset_path, name))
        except Exception as exc:                        # the reference would raise; keep going and report
            print(f'Could not read {name}: {exc}')
            continue
        loaded.append((name, y, sr))
    return loaded


def _write_clip(save_path, name, f, packed):
    base_name = os.path.splitext(name)[0]
    if packed:      # one file per clip instead of ~10 per second of audio; audio_io.explode_features undoes it
        audio_io.save_features_packed(os.path.join(save_path, base_name + audio_io.FEATURE_PACK_SUFFIX), f)
    else:           # {base}_segment_{k}.npy, un-padded running counter (cqt.py:62)
        audio_io.save_features_exploded(save_path, base_name, f)
    print(f'Saved {len(f)} valid segments for {name} in {save_path}')
    return len(f)


def process_all_audio(dataset_path, window_size=0.2, hop_size=0.1, save_path='output', files_per_batch=64, packed=False,
                      io_threads=2):
    """Three overlapped stages: a reader thread decodes the next batch of WAV files while the GPU evaluates the current one
    (one GEMM for all its windows) and ``io_threads`` writer threads put the previous batches' ``.npy`` files on disk
    (scripts/file_path_bench.py measures wav-dir -> feature-dir throughput)."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(save_path, exist_ok=True)
    audio_files = [f for f in os.listdir(dataset_path) if f.endswith('.wav')]
    batches = [audio_files[b0:b0 + files_per_batch] for b0 in range(0, len(audio_files), files_per_batch)]
    pending = []
    with ThreadPoolExecutor(1) as reader, ThreadPoolExecutor(max(1, int(io_threads))) as writer:
        nxt = reader.submit(_load_batch, dataset_path, batches[0]) if batches else None
        for i in range(len(batches)):
            loaded = nxt.result()
            nxt = reader.submit(_load_batch, dataset_path, batches[i + 1]) if i + 1 < len(batches) else None
            for sr in sorted({sr for _, _, sr in loaded}):      # one operator per native sample rate
                group = [(n, y) for n, y, s in loaded if s == sr]
                recipe = CqtRecipe(sr=float(sr), window_size=window_size, hop_size=hop_size)
                feats = features.clips_features([y for _, y in group], recipe)
                for (name, y), f in zip(group, feats):
                    print(f'Processing {len(f)} valid segments for: {name}')
                    pending.append(writer.submit(_write_clip, save_path, name, f, packed))
        written = sum(p.result() for p in pending)
    return written


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="CQT dB features of every 0.2 s window of every .wav in a directory")
    ap.add_argument("dataset_path")
    ap.add_argument("--save-path", default="output")
    ap.add_argument("--window-size", type=float, default=0.2)
    ap.add_argument("--hop-size", type=float, default=0.1)
    a = ap.parse_args()
    process_all_audio(a.dataset_path, window_size=a.window_size, hop_size=a.hop_size, save_path=a.save_path)
