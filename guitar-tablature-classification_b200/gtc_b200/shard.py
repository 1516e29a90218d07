"""Multi-GPU sharding: one process per GPU, clips partitioned across ranks, no data-path collective.

The reference parallelises over independent (file, offset) tasks with a ProcessPoolExecutor (new_cqt.py:53-61) and sums
its per-file stats serially (jam_to_tablature.py:376-378).  Here clip ``c`` goes to rank ``c % world_size`` (or a greedy
duration balance) and the only communication is one all-gather of an 8 x int64 stats vector per shard at the end.
Works with backend "nccl" (CUDA tensors) and "gloo" (CPU tensors, used by the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

STAT_FIELDS = ("n_clips", "n_segments", "n_samples", "total", "with_notes", "with_first_string", "n_skipped", "elapsed_ns")


def partition_round_robin(n_clips: int, rank: int, world_size: int) -> np.ndarray:
    """Indices of the clips owned by ``rank``: c % world_size == rank."""
    return np.arange(rank, n_clips, world_size, dtype=np.int64)


def partition_balanced(durations: Sequence[float], rank: int, world_size: int) -> np.ndarray:
    """Greedy longest-first balance by duration (deterministic: ties broken by clip index)."""
    order = sorted(range(len(durations)), key=lambda i: (-float(durations[i]), i))
    loads = [0.0] * world_size
    owner = [0] * len(durations)
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        owner[i] = r
        loads[r] += float(durations[i])
    return np.asarray([i for i in range(len(durations)) if owner[i] == rank], dtype=np.int64)


@dataclass
class ShardStats:
    n_clips: int = 0
    n_segments: int = 0
    n_samples: int = 0
    total: int = 0
    with_notes: int = 0
    with_first_string: int = 0
    n_skipped: int = 0
    elapsed_ns: int = 0

    def as_tensor(self, device="cpu") -> torch.Tensor:
        return torch.tensor([getattr(self, f) for f in STAT_FIELDS], dtype=torch.int64, device=device)


def gather_stats(local: torch.Tensor) -> torch.Tensor:
    """All-gather the 8 x int64 stats vector of every rank -> [world_size, 8].  Identity when not distributed."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local.reshape(1, -1).clone()
    out = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(out, local)
    return torch.stack(out)


def reduce_stats(gathered: torch.Tensor) -> dict:
    """Job totals the way jam_to_tablature.py:376-378 sums them; elapsed is the max over ranks."""
    g = gathered.cpu().numpy()
    tot = {f: int(g[:, i].sum()) for i, f in enumerate(STAT_FIELDS)}
    tot["elapsed_ns"] = int(g[:, STAT_FIELDS.index("elapsed_ns")].max())
    return tot


def merge_sharded(outputs: List[np.ndarray], owners: List[np.ndarray], counts: np.ndarray) -> np.ndarray:
    """Re-assemble per-rank per-segment outputs into clip order (tests: N-way shard == 1-way run, bit for bit).
    ``owners[r]`` lists rank r's clip indices (in its processing order), ``counts[c]`` segments of clip c."""
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    first = next(o for o in outputs if o is not None and len(o))
    full = np.empty((int(off[-1]),) + first.shape[1:], dtype=first.dtype)
    for out, own in zip(outputs, owners):
        at = 0
        for c in own:
            n = int(counts[c])
            full[off[c]:off[c] + n] = out[at:at + n]
            at += n
    return full
