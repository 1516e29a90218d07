"""Chunk planning for the shard pipeline: pure integer arithmetic, no CUDA (tests/test_host_logic.py runs it on CPU).

A chunk is a run of whole clips.  Its size decides two costs on the GPU:
  * the persistent tcgen05 GEMM deals ceil(rows / 128) x ceil(n_out / tile_n) tiles round-robin to one CTA per SM, so a
    chunk costs ceil(tiles / SMs) full waves -- 3.4 waves cost as much as 4.0;
  * with host inputs the first chunks must be small, or the compute stream sits idle while their audio is on the bus.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

GEMM_TILE_M = 128                           # cqt_gemm_tc.cu TBM
GEMM_TILE_WIDTHS = (256, 240, 192, 128, 64)  # cqt_gemm_tc.cu kTileWidths: the first that divides n_out
# host-input head in GEMM waves: chunks of 1, 1, 2, 2, 3, 3 FULL waves (15, 15, 31, 31, 47, 47 clips of 30 s on 148 SMs), so
# the small chunks of the ramp do not pay for half-empty waves.  GTC_RAMP="w0,w1,..." overrides it for experiments.
RAMP_WAVES = (1, 1, 2, 2, 3, 3)


def rows_per_wave(n_out: int, sm_count: int) -> int:
    """Operand rows one full wave of GEMM tiles covers: sm_count tiles of GEMM_TILE_M rows, ceil(n_out / width) per row block."""
    width = next((w for w in GEMM_TILE_WIDTHS if n_out % w == 0), GEMM_TILE_WIDTHS[0])
    return int(sm_count) * GEMM_TILE_M // -(-int(n_out) // width)


def gemm_tiles(n_seg: int, n_clips: int, parts: int, n_out: int) -> int:
    """Tiles of the segment-operator GEMM for a chunk: clip c owns n_seg_c + (parts - 1) operand rows."""
    rows = int(n_seg) + int(n_clips) * (int(parts) - 1)
    width = next((w for w in GEMM_TILE_WIDTHS if n_out % w == 0), GEMM_TILE_WIDTHS[0])
    return -(-rows // GEMM_TILE_M) * -(-int(n_out) // width)


def wave_efficiency(n_seg: int, n_clips: int, parts: int, n_out: int, sm_count: int) -> float:
    """Filled fraction of the waves the chunk's GEMM occupies (1.0 = the last wave is full)."""
    tiles = gemm_tiles(n_seg, n_clips, parts, n_out)
    if tiles == 0:
        return 1.0
    waves = -(-tiles // int(sm_count))
    return tiles / float(waves * int(sm_count))


def plan_bounds(n_seg_per_clip: Sequence[int], limit: int, ramp: bool = False,
                efficiency: Optional[Callable[[int, int], float]] = None, wave_rows: Optional[int] = None) -> List[Tuple[int, int]]:
    """[(first clip, end clip)] of every chunk: whole clips, at most ``limit`` segments (a single longer clip is its own
    chunk).  ``ramp``: the first chunks hold RAMP_WAVES waves of ``wave_rows`` rows each (an eighth of the limit, growing,
    when ``wave_rows`` is not given) -- the copy engine delivers ~37 clips/ms of int16 PCM (PCIe ~50 GB/s) against ~31
    clips/ms of kernels, so a chunk may be at most ~1.2 x + a head start larger than its predecessor if it is to find its
    audio resident; there is no ramp-down because the last chunk's device->host copy hides under its own patch stores.
    ``efficiency(n_seg, n_clips)``: when given, a chunk that is followed by more clips ends, within the last 20 % of its
    greedy size, where the efficiency is highest (ties: the larger chunk)."""
    nseg = np.asarray(n_seg_per_clip, dtype=np.int64)
    seg_off = np.concatenate([[0], np.cumsum(nseg)])
    n_clips, total = len(nseg), int(seg_off[-1])
    waves = RAMP_WAVES
    if os.environ.get("GTC_RAMP"):
        waves = tuple(float(x) for x in os.environ["GTC_RAMP"].split(","))
    unit = wave_rows if wave_rows else max(1, limit // 8)
    sizes = [min(limit, int(w * unit)) for w in waves] if ramp and total > 3 * limit else []
    bounds, c0, k = [], 0, 0
    while c0 < n_clips:
        cap = sizes[k] if k < len(sizes) else limit
        c1 = c0 + 1
        while c1 < n_clips and seg_off[c1 + 1] - seg_off[c0] <= cap:
            c1 += 1
        if efficiency is not None and c1 < n_clips and c1 - c0 > 4:
            lo = c0 + max(1, int(0.8 * (c1 - c0)))
            c1 = max(range(lo, c1 + 1), key=lambda c: (round(efficiency(int(seg_off[c] - seg_off[c0]), c - c0), 2), c))
        bounds.append((c0, c1))
        c0, k = c1, k + 1
    return bounds
