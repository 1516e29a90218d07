"""Corpus streaming: clips sharded over ranks -> training batches in the dataloaders' tensor contracts, never stored.

BASELINE.json configs[3] (my_dataloader.py batches ``(128, 3, 224, 224)`` + ``(128, 6)`` into the CNN of bestengine.py)
and configs[4] (a 10k-clip corpus sharded ``clip % world_size`` feeding ViT_dataloader.py batches ``(50, 3, 224, 224)`` +
six ``(50, 19)`` label heads).  The patch tensors of a corpus are far larger than HBM (1.8 TB for 10k clips), so batches are
views into the FrontEnd's patch ring, handed to a consumer callback on the compute stream and then overwritten.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np
import torch

from . import _lib, ops, shard, synth
from .cqt_design import CqtRecipe
from .pipeline import FrontEnd, ShardInputs


@dataclass
class StreamReport:
    n_clips: int = 0
    n_segments: int = 0
    n_batches: int = 0
    n_full_batches: int = 0
    seconds_of_audio: float = 0.0
    device_ms: float = 0.0
    label_stats: Optional[np.ndarray] = None


def device_corpus_block(clip_ids: np.ndarray, n_samples: int, sr: float, device, plucks_per_s: float = 2.0) -> torch.Tensor:
    """Synthetic clips generated ON the device from torch's counter-based CUDA generator (Philox), one seed per clip id, so
    any rank can regenerate any clip: [len(clip_ids) * n_samples] fp32.  (26.5 GB of audio for 10k clips is never stored.)"""
    out = torch.empty((len(clip_ids), n_samples), dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    t = torch.arange(n_samples, dtype=torch.float32, device=device) / sr
    dur = n_samples / sr
    n_pl = max(1, int(round(plucks_per_s * dur)))
    for i, cid in enumerate(clip_ids):
        g.manual_seed(1_000_003 * int(cid) + 17)
        par = torch.rand((4, n_pl), generator=g, device=device)
        y = torch.randn(n_samples, generator=g, device=device) * 0.01
        f = 440.0 * 2 ** ((40 + 42 * par[0] - 69) / 12)
        onset, amp, tau = dur * par[1], 0.3 + 0.7 * par[2], 0.15 + 0.6 * par[3]
        for p0 in range(0, n_pl, 16):                           # 16 plucks per pass keeps the temporaries at 42 MB
            rel = t[None, :] - onset[p0:p0 + 16, None]
            env = torch.where(rel >= 0, torch.exp(-rel.clamp(min=0) / tau[p0:p0 + 16, None]), torch.zeros_like(rel))
            ph = 6.283185307179586 * f[p0:p0 + 16, None] * rel
            y = y + (amp[p0:p0 + 16, None] * env * (torch.sin(ph) + 0.5 * torch.sin(2 * ph))).sum(0)
        out[i] = y * (0.5 / y.abs().max().clamp(min=1e-12))
    return out.reshape(-1)


def stream_corpus(n_clips: int, clip_seconds: float, rank: int = 0, world_size: int = 1, batch_size: int = 50,
                  mode: str = "vit", consumer: Optional[Callable] = None, recipe: CqtRecipe = CqtRecipe(),
                  clips_per_block: int = 360, device: Optional[int] = None, plucks_per_s: float = 2.0) -> StreamReport:
    """Process this rank's clips (``clip % world_size == rank``) block by block and hand every training batch to
    ``consumer(inputs, labels)``: ViT mode -> inputs (B,3,224,224) fp32, labels = list of six (B,19) int64;
    CNN mode -> inputs ImageNet-normalised, labels (B,6) int64.  Batches never straddle a patch-ring boundary, so a block
    ends with at most a few short batches (the DataLoader's last batch is short in the same way)."""
    dev_index = torch.cuda.current_device() if device is None else int(device)
    dev = torch.device("cuda", dev_index)
    sr = float(recipe.sr)
    n_samples = int(sr * clip_seconds)
    mine = shard.partition_round_robin(n_clips, rank, world_size)
    patch_mode = _lib.GTC_PATCH_VIT if mode == "vit" else _lib.GTC_PATCH_CNN
    chunk_segments = 28400                                          # FrontEnd default: six full GEMM tile waves of 30 s clips
    ring = max(batch_size, (chunk_segments // batch_size) * batch_size)   # whole batches per patch launch
    fe = FrontEnd(recipe, device=dev_index, patch_mode=patch_mode, chunk_segments=chunk_segments, patch_batch=ring)
    rep = StreamReport(label_stats=np.zeros(3, dtype=np.int64))

    def on_ring(patches, tabs, first_segment):
        n = patches.shape[0]
        heads = ops.labels_vit_heads(tabs) if mode == "vit" else ops.labels_argmax(tabs)
        for b0 in range(0, n, batch_size):
            b1 = min(n, b0 + batch_size)
            rep.n_batches += 1
            rep.n_full_batches += int(b1 - b0 == batch_size)
            if consumer is not None:
                labels = [h[b0:b1] for h in heads] if mode == "vit" else heads[b0:b1]
                consumer(patches[b0:b1], labels)

    for c0 in range(0, len(mine), clips_per_block):
        ids = mine[c0:c0 + clips_per_block]
        audio = device_corpus_block(ids, n_samples, sr, dev, plucks_per_s)
        # note events are a function of the CLIP id alone (like the audio), so any sharding of the corpus rasterises the
        # same labels and the gathered stats of N ranks equal the serial sums (jam_to_tablature.py:376-378)
        per_clip = [synth.note_events([clip_seconds], seed=7919 * int(cid) + 2) for cid in ids]
        on, du, pi = (np.concatenate([e[j] for e in per_clip]) for j in range(3))
        evt_off = np.concatenate([[0], np.cumsum([len(e[0]) for e in per_clip])]).astype(np.int64)
        events = torch.from_numpy(np.stack([on, du, pi])).to(dev)
        inp = ShardInputs(audio, np.full(len(ids), n_samples, dtype=np.int64), events, evt_off, sr=sr)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fe.run(inp, device_inputs=True, consumer=on_ring)
        e1.record()
        torch.cuda.synchronize(dev)
        rep.device_ms += e0.elapsed_time(e1)
        rep.n_clips += len(ids)
        rep.n_segments += out.n_seg
        rep.seconds_of_audio += out.seconds_of_audio
        rep.label_stats += fe.stats()
        del audio, events, inp
    return rep
