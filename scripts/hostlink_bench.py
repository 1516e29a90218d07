#!/usr/bin/env python
"""Host-link ceiling of the box: concurrent pinned host<->device copies on N GPUs, no kernels.

    python scripts/hostlink_bench.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/hostlink_bench.py [--out gpurun_out/hostlink_nN.jsonl]

What it measures is the denominator of bench.py's `e2e.hostlink_frac`: the end-to-end arm uploads one shard's int16 PCM
(360 clips x 30 s = 476 MB) and reads dB features + labels back (219 MB) every step; the reference fans the same data
out with a process pool over host memory (/root/reference/new_cqt.py:53-61).  Every rank copies the SAME byte counts
through its own PCIe link at the same time (barrier + device events, max over ranks):

    directions : h2d only, d2h only, both at once (two streams: the GPU's two copy engines)
    granularity: one copy, 64 MB pieces, 8-clip pieces (10.6 MB = FrontEnd.stage_piece_clips) and, for d2h, the
                 per-chunk pairs (features + labels of a 94-clip chunk) the pipeline issues

One JSON line per case on rank 0: per-rank GB/s (bytes of one rank / max-over-ranks time) and aggregate GB/s.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

SR = 22050
CLIP_SAMPLES = int(SR * 30.0)
SEG_PER_CLIP = (CLIP_SAMPLES - 4410) // 2205 + 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=360)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--out", default=None, help="append the JSON lines to this file (rank 0)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)

    n_in = args.clips * CLIP_SAMPLES                         # int16 samples of the shard
    n_seg = args.clips * SEG_PER_CLIP
    n_db, n_tab = n_seg * 480, n_seg * 114                   # fp32 features, int8 labels
    h_in = torch.empty(n_in, dtype=torch.int16, pin_memory=True)
    h_in.zero_()
    d_in = torch.empty(n_in, dtype=torch.int16, device=dev)
    d_db = torch.zeros(n_db, dtype=torch.float32, device=dev)
    d_tab = torch.zeros(n_tab, dtype=torch.int8, device=dev)
    h_db = torch.empty(n_db, dtype=torch.float32, pin_memory=True)
    h_tab = torch.empty(n_tab, dtype=torch.int8, pin_memory=True)
    h_db.zero_(); h_tab.zero_()
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    in_bytes, out_bytes = n_in * 2, n_db * 4 + n_tab

    def pieces(n, step):
        return [(a, min(n, a + step)) for a in range(0, n, step)]

    def up(step_elems):
        def go():
            with torch.cuda.stream(s_up):
                for a, b in pieces(n_in, step_elems):
                    d_in[a:b].copy_(h_in[a:b], non_blocking=True)
        return go

    def down(step_segs, pairs=True):
        def go():
            with torch.cuda.stream(s_dn):
                for a, b in pieces(n_seg, step_segs):
                    h_db[a * 480:b * 480].copy_(d_db[a * 480:b * 480], non_blocking=True)
                    if pairs:
                        h_tab[a * 114:b * 114].copy_(d_tab[a * 114:b * 114], non_blocking=True)
                if not pairs:
                    h_tab.copy_(d_tab, non_blocking=True)
        return go

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(name, fns, nbytes):
        for _ in range(args.warmup):
            for f in fns:
                f()
        barrier()
        times = []
        for _ in range(args.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            s_up.wait_event(e0); s_dn.wait_event(e0)
            for f in fns:
                f()
            torch.cuda.current_stream().wait_stream(s_up)
            torch.cuda.current_stream().wait_stream(s_dn)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            times.append(float(ms.item()))
        best, mean = min(times), float(np.mean(times))
        line = {"case": name, "n_gpus": world, "bytes_per_rank": nbytes, "ms_best": round(best, 3), "ms_mean": round(mean, 3),
                "per_rank_GBs": round(nbytes / (mean * 1e-3) / 1e9, 2), "aggregate_GBs": round(world * nbytes / (mean * 1e-3) / 1e9, 2),
                "how": "barrier; device events around the copies of every rank; max over ranks; mean of %d" % args.iters}
        if rank == 0:
            print(json.dumps(line), flush=True)
            if args.out:
                with open(args.out, "a") as f:
                    f.write(json.dumps(line) + "\n")
        return mean

    clip8 = 8 * CLIP_SAMPLES
    mb64 = 64 << 20
    measure("h2d one copy", [up(n_in)], in_bytes)
    measure("h2d 64MB pieces", [up(mb64 // 2)], in_bytes)
    measure("h2d 8-clip pieces (10.6 MB)", [up(clip8)], in_bytes)
    measure("d2h one copy each (features, labels)", [down(n_seg, pairs=False)], out_bytes)
    measure("d2h per 94-clip chunk, features+labels pairs", [down(94 * SEG_PER_CLIP)], out_bytes)
    measure("d2h 8-clip pieces, pairs", [down(8 * SEG_PER_CLIP)], out_bytes)
    t_both = measure("both: h2d 8-clip pieces + d2h per-chunk pairs", [up(clip8), down(94 * SEG_PER_CLIP)], in_bytes + out_bytes)
    measure("both: one copy each", [up(n_in), down(n_seg, pairs=False)], in_bytes + out_bytes)
    measure("both: 64MB pieces + d2h per-chunk", [up(mb64 // 2), down(94 * SEG_PER_CLIP)], in_bytes + out_bytes)
    if rank == 0:
        s_audio = args.clips * 30.0
        line = {"case": "e2e ceiling (both directions, pipeline granularity)", "n_gpus": world,
                "step_ms_floor": round(t_both, 3), "s_audio_per_s_ceiling": round(world * s_audio / (t_both * 1e-3), 1)}
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(line) + "\n")
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
