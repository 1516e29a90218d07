#!/usr/bin/env python
"""bench.py -- seconds-of-audio/sec of the hot path (CQT + dB, labels, patches) on N B200s, one JSON line.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (default N=1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (restated oracle)

A "step" is one pass of the hot path over BASELINE.json configs[1]: 360 synthetic clips x 30 s @ 22.05 kHz mono
(107 640 segments) per GPU (weak scaling).  `value` = kernel-only throughput with inputs resident in HBM;
`e2e` = the same through the public FrontEnd API with pinned HOST inputs (audio + events H2D, features + labels D2H in
the timed region; patches stay on the device because the training engines consume them there).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.join(ROOT, "guitar-tablature-classification_b200")
for _p in (ROOT, PKG_ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "seconds-of-audio/sec CQT+labels+patches"
UNIT = "s_audio/s"
SR = 22050
N_CLIPS = 360
CLIP_SECONDS = 30.0
ALGO_BYTES_PER_S_AUDIO = 88200 + 19200 + 1600 + 19200 + 6021120          # SURVEY.md 8d / BASELINE.md section 4
PATCH_BYTES_PER_SEGMENT = 3 * 224 * 224 * 4 + 96 * 5 * 4                  # written + read per segment by the patch kernel


def workload_config(n_clips=N_CLIPS):
    n = int(SR * CLIP_SECONDS)
    segs = (n - 4410) // 2205 + 1
    return {"workload": "BASELINE.json configs[1]: GuitarSet-shaped batch, %d synthetic clips x %.0f s @ 22.05 kHz mono -> "
                        "cqt.py CQT+|.|^4+dB+cut, jam_to_tablature labels, ViT_dataloader (3,224,224) patches" % (n_clips, CLIP_SECONDS),
            "clips_per_gpu": n_clips, "segments_per_gpu": n_clips * segs, "sr": SR, "seg_len": 4410, "seg_hop": 2205,
            "n_bins": 96, "frames": 5, "patch": [3, 224, 224],
            "cache": "per-step inputs (0.95 GB audio) and outputs (65 GB of patches through a 17 GB ring) exceed the 126 MB L2; no flush needed"}


# ======================================================================================================================
# CPU reference arm: the reference's control flow (cqt.py loop, jam_to_tablature.process_file loop,
# ViT_dataloader.__getitem__) executed with the restated oracle -- librosa/soxr/jams are not installable here.
# ======================================================================================================================

def _cpu_clip_job(args):
    """One clip through the reference's three scripts, file I/O on tmpfs included."""
    seed, n_samples, outdir = args
    import torch
    torch.set_num_threads(1)
    from oracle import cqt_oracle, labels_oracle, patches_oracle
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples) / SR
    y = 0.003 * rng.standard_normal(n_samples)
    for _ in range(int(6 * n_samples / SR)):
        f = 440.0 * 2 ** ((rng.uniform(40, 82) - 69) / 12)
        on = rng.uniform(0, n_samples / SR)
        rel = np.maximum(t - on, 0)
        y += rng.uniform(0.3, 1.0) * np.where(t >= on, np.exp(-rel / 0.4), 0.0) * np.sin(2 * np.pi * f * rel)
    y = (0.5 * y / np.abs(y).max()).astype(np.float32)
    dur = n_samples / SR
    notes = []
    for s in range(6):
        for _ in range(rng.poisson(3 * dur)):
            notes.append(labels_oracle.Observation(rng.uniform(0, dur), float(np.clip(rng.exponential(0.4), 0.05, 4)),
                                                   [40, 45, 50, 55, 59, 64][s] + int(rng.integers(0, 19)) + rng.normal(0, 0.15)))
    jam = labels_oracle.Jam([labels_oracle.Annotation('note_midi', notes)])
    t0 = time.perf_counter()
    feats = cqt_oracle.process_clip(y, SR)                                   # cqt.py:19-65 (basis rebuilt per call, like librosa)
    for k, f in enumerate(feats):
        np.save(os.path.join(outdir, f"clip{seed}_segment_{k}.npy"), np.asfortranarray(f))
    times = labels_oracle.segment_times(dur, len(feats))                     # jam_to_tablature.py:259-274
    tabs, _ = labels_oracle.process_segments(jam, times)
    for i, tab in enumerate(tabs):
        np.save(os.path.join(outdir, f"clip{seed}_{i:04d}.npy"), tab)
    for k in range(len(feats)):                                              # ViT_dataloader.py:22-56
        a = np.load(os.path.join(outdir, f"clip{seed}_segment_{k}.npy"))
        lab = np.load(os.path.join(outdir, f"clip{seed}_{k:04d}.npy")).astype(np.float32)
        patches_oracle.vit_patch_torch(a)
        [lab[i].astype(np.int64) for i in range(6)]
    return dur, time.perf_counter() - t0


def cpu_reference_pass(clips: int, clip_seconds: float, workers: int, seed0: int = 0):
    """ProcessPoolExecutor fan-out over clips as new_cqt.py:53-61 does.  Returns (audio seconds, seconds of work):
    the jobs run concurrently (one per worker per round) and each reports its own busy time, so interpreter start-up
    and imports are excluded while memory-bandwidth contention between the workers is included."""
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    outdir = tempfile.mkdtemp(prefix="gtc_cpu_", dir=base)
    try:
        jobs = [(seed0 + i, int(SR * clip_seconds), outdir) for i in range(clips)]
        if workers <= 1:
            res = [_cpu_clip_job(j) for j in jobs]
        else:
            with ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("fork")) as ex:
                res = list(ex.map(_cpu_clip_job, jobs))
        rounds = -(-clips // max(1, workers))
        wall = rounds * max(r[1] for r in res) if workers > 1 else sum(r[1] for r in res)
        return sum(r[0] for r in res), wall
    finally:
        shutil.rmtree(outdir, ignore_errors=True)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    workers = max(1, cores)
    clip_s = 10.0                                       # bounded sample: `workers` clips x 10 s per step
    for _ in range(args.warmup):
        cpu_reference_pass(min(workers, 2), 2.0, min(workers, 2))
    t_audio = t_wall = 0.0
    for s in range(args.steps):
        a, w = cpu_reference_pass(workers, clip_s, workers, seed0=1000 * s)
        t_audio += a
        t_wall += w
    value = t_audio / t_wall
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_wall / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                             "sample": f"{workers} clips x {clip_s:.0f} s per step on {workers} processes "
                                       "(restated NumPy/SciPy oracle of cqt.py + jam_to_tablature.py + ViT_dataloader.py, "
                                       ".npy I/O on tmpfs; librosa/soxr/jams are not installable, so this is a port, not librosa; "
                                       "time = rounds x the slowest worker's busy time, process start-up and imports excluded)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ======================================================================================================================
# CUDA arm
# ======================================================================================================================

class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML, in-process thread)."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index: int, period: float = 0.02):
        self.period = period
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None and self.period > 0:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def hostlink_floor(fe, inp_host, chunk_plan, out, world, dev, iters=5, d2h=True):
    """Copies only: the end-to-end arm's bytes per step (int16 PCM + events up in `stage_piece_clips`-clip pieces, dB
    features + labels down once per chunk) with no kernel in between, every rank at the same time.  The fastest step the
    box's host links allow -> `e2e.hostlink_frac` = this floor / the measured e2e step (scripts/hostlink_bench.py sweeps
    the granularities; profiles/r02_hostlink.md)."""
    import torch
    import torch.distributed as dist
    n_in = inp_host.audio.numel()
    d_in = torch.empty(n_in, dtype=inp_host.audio.dtype, device=dev)
    d_ev = torch.empty(inp_host.events.shape, dtype=torch.float64, device=dev)
    n_seg = out.n_seg
    d_db = torch.zeros((n_seg, 96, 5), dtype=torch.float32, device=dev) if d2h else None
    d_tab = torch.zeros((n_seg, 6, 19), dtype=torch.int8, device=dev) if d2h else None
    h_db = fe._buf("db_host", (n_seg, 96, 5), torch.float32, pinned=True) if d2h else None
    h_tab = fe._buf("tabs_host", (n_seg, 6, 19), torch.int8, pinned=True) if d2h else None
    s_up, s_dn = fe.s_stage, fe.s_out
    piece = fe.stage_piece_clips
    clip_off = np.concatenate([[0], np.cumsum(inp_host.clip_lens)]).astype(np.int64)

    def one():
        with torch.cuda.stream(s_up):
            for c in range(0, len(inp_host.clip_lens), piece):
                a, b = int(clip_off[c]), int(clip_off[min(len(inp_host.clip_lens), c + piece)])
                d_in[a:b].copy_(inp_host.audio[a:b], non_blocking=True)
                if c == 0:
                    d_ev.copy_(inp_host.events, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s_dn):
                for ch in chunk_plan:
                    h_db[ch.g0:ch.g1].copy_(d_db[ch.g0:ch.g1], non_blocking=True)
                    h_tab[ch.g0:ch.g1].copy_(d_tab[ch.g0:ch.g1], non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        one()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s_up.wait_event(e0); s_dn.wait_event(e0)
    for _ in range(iters):
        one()
    torch.cuda.current_stream().wait_stream(s_up)
    torch.cuda.current_stream().wait_stream(s_dn)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    nbytes = n_in * inp_host.audio.element_size() + inp_host.events.numel() * 8 + (n_seg * (1920 + 114) if d2h else 0)
    return {"ms_per_step_floor": float(ms.item()), "bytes_per_rank": int(nbytes),
            "aggregate_GBs": world * nbytes / (float(ms.item()) * 1e-3) / 1e9}


def ragged_corpus(n_clips_total, seed=1):
    """clip durations uniform in [14.6, 30.0] s (the committed label fixtures show 73..125 segments
    of 0.2 s per clip), numpy default_rng(seed)."""
    rng = np.random.default_rng(seed)
    return (rng.uniform(14.6, 30.0, n_clips_total) * SR).astype(np.int64)


def ragged_arm(args, fe, timed, rank, world, dev):
    """Device-resident arm on ragged clips.  A corpus of 360 x world clips is dealt to the ranks (a) `c % W` as north_star
    words it (shard.partition_round_robin) and (b) longest-first greedy by duration (shard.partition_balanced); value =
    all ranks' seconds of audio / the slowest rank's time."""
    import torch
    import torch.distributed as dist
    from gtc_b200 import synth, shard
    from gtc_b200.pipeline import ShardInputs
    lens_all = ragged_corpus(args.clips * world)
    res = {"durations": "uniform [14.6, 30.0] s, numpy default_rng(1), %d clips over %d rank(s)" % (len(lens_all), world)}
    n_max = int(SR * CLIP_SECONDS)
    parts = [("round_robin", shard.partition_round_robin(len(lens_all), rank, world))]
    if world > 1:
        parts.append(("balanced", shard.partition_balanced((lens_all / SR).tolist(), rank, world)))
    for name, ids in parts:
        lens = lens_all[ids]
        full = synth.pluck_clips(len(ids), n_max, sr=SR, seed=101 + rank, device=dev, block=24)
        audio = torch.cat([full[i, : int(lens[i])] for i in range(len(ids))]).contiguous()
        del full
        on, du, pi, evt_off = synth.note_events((lens / SR).tolist(), seed=202 + rank)
        events = torch.from_numpy(np.stack([on, du, pi])).to(dev)
        inp = ShardInputs(audio, lens, events, evt_off, sr=SR)
        plan = fe.plan_chunks(inp)
        ms, out = timed(inp, True, args.steps, args.warmup, plans=(plan, plan, plan), events=False)
        secs = torch.tensor([float(lens.sum()) / SR], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(secs, op=dist.ReduceOp.SUM)
        res[name] = {"value": float(secs.item()) * args.steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / args.steps,
                     "segments_this_rank": int(out.n_seg), "chunks": [int(c.g1 - c.g0) for c in plan]}
        del audio, events, inp
    return res



def run_cuda_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # CPU baseline first (rank 0, N=1 only), before CUDA is initialised in this process
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        workers = max(1, min(cores, 16))
        a, w = cpu_reference_pass(workers, 10.0, workers)
        cpu_baseline = {"value": a / w, "unit": UNIT, "cores": workers, "kind": "port",
                        "sample": f"{workers} clips x 10 s, one process per core ({cores} host cores), restated reference "
                                  "(cqt.py + jam_to_tablature.py + ViT_dataloader.py loops on the NumPy/SciPy oracle, .npy I/O on tmpfs); "
                                  "not librosa -- it is not installable here; time = the slowest worker's busy time, process start-up "
                                  "and imports excluded"}

    import torch
    import torch.distributed as dist
    from gtc_b200 import CqtRecipe, synth, shard
    from gtc_b200.pipeline import FrontEnd, ShardInputs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this framework has no CPU fallback")
    numa = None
    if args.numa_bind:
        # run this rank's host thread (and first-touch its pinned staging buffers) on the CPUs NVML reports as local to
        # the rank's GPU: the host-input arm moves 0.7 GB per step per GPU through the socket's PCIe root
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(physical_gpu_index(local_rank))
            pynvml.nvmlDeviceSetCpuAffinity(h)
            numa = sorted(os.sched_getaffinity(0))
            numa = f"{len(numa)} cpus {numa[0]}-{numa[-1]}"
        except Exception as e:                      # no NVML / not permitted: run unbound
            numa = f"unbound ({type(e).__name__})"
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)

    recipe = CqtRecipe()
    n_clips = args.clips
    n = int(SR * CLIP_SECONDS)
    # synthetic shard of this rank (clip ids rank, rank+world, ... of a 360*world-clip corpus): weak scaling
    audio_dev = synth.pluck_clips(n_clips, n, sr=SR, seed=1 + rank, device=dev, block=24).reshape(-1)
    on, du, pi, evt_off = synth.note_events([CLIP_SECONDS] * n_clips, seed=2 + rank)
    events_host = torch.from_numpy(np.stack([on, du, pi])).pin_memory()
    events_dev = events_host.to(dev)
    # The clips are 16-bit PCM WAV files in the reference's world (GuitarSet; librosa.load converts x/32768 to fp32,
    # cqt.py:23).  Quantise the synthetic audio to PCM once; the device-resident arm gets the fp32 array librosa would
    # return, the end-to-end arm uploads the file's int16 samples and converts on the device (identical values).
    pcm_dev = torch.clamp(torch.round(audio_dev * 32768.0), -32768, 32767).to(torch.int16)
    audio_dev = pcm_dev.to(torch.float32) / 32768.0
    host_dtype = torch.float32 if args.host_audio == "f32" else torch.int16
    audio_host = torch.empty(audio_dev.shape, dtype=host_dtype, pin_memory=True)
    audio_host.copy_(audio_dev if args.host_audio == "f32" else pcm_dev)
    del pcm_dev
    lens = np.full(n_clips, n, dtype=np.int64)
    from gtc_b200 import ops as _ops
    for kv in args.opt:
        k, v = kv.split("=")
        _ops.set_option(int(k), int(v))
    fe = FrontEnd(recipe, device=local_rank, engine=args.engine, chunk_segments=args.chunk_segments, patch_batch=args.patch_batch,
                  overlap=args.overlap, gemm_ctas=args.gemm_ctas, patch_ctas_per_sm=args.patch_ctas_per_sm,
                  coresident=args.coresident, wave_aware=not args.no_wave_aware)
    inp_dev = ShardInputs(audio_dev, lens, events_dev, evt_off, sr=SR)
    inp_host = ShardInputs(audio_host, lens, events_host, evt_off, sr=SR)
    chunks = fe.plan_chunks(inp_dev)
    chunks_host = fe.plan_chunks(inp_host, ramp=not args.no_ramp)     # small first chunks: short pipeline fill
    chunks_flat_host = fe.plan_chunks(inp_host, ramp=False)
    seconds_per_step = n_clips * CLIP_SECONDS
    stats_vec = torch.zeros(8, dtype=torch.int64, device=dev)

    # shard constants of the stats vector: staged once in pinned memory.  (`stats_vec[0] = n_clips` would be a copy from
    # PAGEABLE host memory, which synchronises the stream first -- one hidden host sync per step.)
    stats_head = torch.tensor([n_clips, 0, int(lens.sum())], dtype=torch.int64).pin_memory()

    def step(inp, device_inputs, last=False, first=False, host_outputs=True, plans=None):
        # host-input arm: the next step's shard (the same pinned buffers stand in for it) is prefetched behind this
        # step's copies, as a training loop would do with the next shard of the corpus; the last step prefetches nothing
        # a cold step ramps its chunks up from 15 clips (its audio is still on the bus); a prefetched step finds ~90 clips
        # resident and uses the same full-wave chunks as the device-resident arm
        pref = args.prefetch and not device_inputs
        p_dev, p_ramp, p_flat = plans if plans is not None else (chunks, chunks_host, chunks_flat_host)
        mine = p_dev if device_inputs else (p_flat if (pref and not first and args.flat_after_first) else p_ramp)
        nxt = p_flat if args.flat_after_first else p_ramp
        out = fe.run(inp, device_inputs=device_inputs, chunks=mine, want_host_outputs=host_outputs,
                     next_inp=inp if (pref and not last) else None, next_chunks=nxt)
        # tiny per-shard stats gather (the path's only collective), jam_to_tablature.py:376-378
        stats_head[1] = out.n_seg
        stats_vec[:3].copy_(stats_head, non_blocking=True)
        stats_vec[3:6] = fe._last_stats[0]
        if world > 1:
            shard.gather_stats(stats_vec)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(inp, device_inputs, steps, warmup, host_outputs=True, plans=None, events=True):
        for _ in range(warmup):
            step(inp, device_inputs, last=True, first=True, host_outputs=host_outputs, plans=plans)   # warm-up steps prefetch nothing: the timed region starts cold
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fe.patch_events = [] if (device_inputs and events) else None      # per-launch CUDA events on the launching stream, timed region only
        fe.gemm_events = [] if (device_inputs and events) else None
        e0.record()
        out = None
        t_host = time.perf_counter()
        for i in range(steps):
            out = step(inp, device_inputs, last=(i == steps - 1), first=(i == 0), host_outputs=host_outputs, plans=plans)
        timed.host_ms = 1e3 * (time.perf_counter() - t_host) / max(1, steps)     # CPU time to enqueue one step
        e1.record()
        barrier()
        if device_inputs and events:
            timed.patch_launches = [(a.elapsed_time(b), n) for a, b, n in fe.patch_events]
            timed.gemm_launches = [(a.elapsed_time(b), n) for a, b, n in fe.gemm_events]
        fe.patch_events = fe.gemm_events = None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    with ClockSampler(physical_gpu_index(local_rank), period=args.clock_period) as clocks:
        ms_dev, out_dev = timed(inp_dev, True, args.steps, args.warmup)
    host_ms_dev = timed.host_ms
    ms_e2e, out_e2e = timed(inp_host, False, args.steps, args.warmup)
    # training flow (BASELINE configs[3]/[4]): the engines consume features, labels and patches ON the device
    # (bestengine.py:899-901, ViT_engine.py:277-278), so only the label stats return to the host
    ms_train, out_train = timed(inp_host, False, args.steps, args.warmup, host_outputs=False)
    # what the box's host links can move: the same bytes per step, same copy granularity, no kernels, all ranks at once
    link = hostlink_floor(fe, inp_host, chunks_flat_host, out_e2e, world, dev, iters=max(3, min(args.steps, 10)))
    link_train = hostlink_floor(fe, inp_host, chunks_flat_host, out_train, world, dev, iters=max(3, min(args.steps, 10)), d2h=False)
    ragged = None
    if not args.no_ragged:
        ragged = ragged_arm(args, fe, timed, rank, world, dev)

    # dominant kernel (patch store stream) timed live, per launch, on its own stream
    n_seg = out_dev.n_seg
    pb = min(args.patch_batch, n_seg)
    db_all = fe._bufs[("db_all", False)][: n_seg * 480].view(n_seg, 96, 5)
    pb = min(pb, fe._bufs[("patch0", False)].numel() // (3 * 224 * 224))
    ring = fe._bufs[("patch0", False)][: pb * 3 * 224 * 224].view(pb, 3, 224, 224)
    from gtc_b200 import ops, _lib
    ops.set_option(_lib.GTC_OPT_PATCH_MAX_CTAS, 0)              # time the patch kernel alone on the whole GPU
    evs = []
    torch.cuda.synchronize()
    with torch.cuda.stream(fe.s_comp):
        for _ in range(2):
            ops.patches(db_all[:pb], out=ring)
        for i in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(fe.s_comp)
            ops.patches(db_all[(i * pb) % max(1, n_seg - pb + 1):][:pb], out=ring)
            b.record(fe.s_comp)
            evs.append((a, b))
        # reference point: the fastest pure store stream this GPU produces over the same buffer (torch.fill_)
        fill_evs = []
        for i in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(fe.s_comp)
            ring.fill_(0.5)
            b.record(fe.s_comp)
            fill_evs.append((a, b))
    torch.cuda.synchronize()
    patch_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    fill_gbs = ring.numel() * 4 / (min(a.elapsed_time(b) for a, b in fill_evs[1:]) * 1e-3) / 1e9

    value = world * seconds_per_step * args.steps / (ms_dev * 1e-3)
    e2e_value = world * seconds_per_step * args.steps / (ms_e2e * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    live = getattr(timed, "patch_launches", [])
    live_ms = float(np.mean([t for t, _ in live])) if live else patch_ms
    live_seg = float(np.mean([n for _, n in live])) if live else pb
    achieved = live_seg * PATCH_BYTES_PER_SEGMENT / (live_ms * 1e-3) / 1e9          # inside the timed steps
    isolated = pb * PATCH_BYTES_PER_SEGMENT / (patch_ms * 1e-3) / 1e9                # the kernel alone on the GPU
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "patch_kernel_traffic.json")))
        # ncu --set full capture of one launch (dram__bytes_read.sum + dram__bytes_write.sum), scaled to this run's launch size
        traffic = tj["dram_bytes_per_launch"] / tj.get("segments_per_launch", 4096) * live_seg
    except Exception:
        pass
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(n_clips),
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": out_e2e.h2d_bytes, "d2h_bytes_per_step": out_e2e.d2h_bytes,
                        "host_audio": "int16 PCM (the WAV files' samples; x/32768 on the device == librosa.load)" if args.host_audio == "pcm16" else "fp32",
                        "note": "pinned host audio+events in, dB features + labels + stats back; patches stay in HBM for the engines",
                        "prefetch": bool(args.prefetch),
                        "cpu_affinity": numa,
                        "hostlink": link,
                        "hostlink_frac": link["ms_per_step_floor"] / (ms_e2e / args.steps),
                        "hostlink_note": "hostlink = the same bytes per step moved with NO kernels (same copy granularity, all ranks at once, device "
                                         "events, max over ranks): the floor the box's host links set for this arm; frac = floor / measured step"},
                "e2e_train": {"value": world * seconds_per_step * args.steps / (ms_train * 1e-3), "unit": UNIT, "ms_per_step": ms_train / args.steps,
                              "h2d_bytes_per_step": out_train.h2d_bytes, "d2h_bytes_per_step": out_train.d2h_bytes,
                              "note": "training flow of BASELINE configs[3]/[4]: pinned host audio+events in; features, labels and patches stay in HBM "
                                      "for the engines (bestengine.py:899-901), only the label stats return",
                              "hostlink": link_train, "hostlink_frac": link_train["ms_per_step_floor"] / (ms_train / args.steps)},
                "ragged": ragged,
                "gpu_launches": out_dev.launches * args.steps, "host_enqueue_ms_per_step": host_ms_dev,
                "roofline": {"bound": "hbm", "kernel": "patch_kernel<5> (gtc_patches)", "achieved": achieved, "peak": peak_hbm,
                             "unit": "GB/s", "frac": achieved / peak_hbm, "traffic": traffic,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                             "note": "frac > 1 because the peak is a measured COPY rate (read + write) while this kernel is a pure store "
                                     "stream; pure_store_gbs is torch.fill_ over the same buffer in this run",
                             "pure_store_gbs": fill_gbs, "frac_of_pure_store": achieved / fill_gbs,
                             "launch_ms": live_ms, "segments_per_launch": live_seg, "launches_timed": len(live),
                             "how": "mean of per-launch CUDA events on the launching stream over every patch launch inside the timed steps",
                             "per_launch_ms": [round(t, 3) for t, _ in live], "per_launch_segments": [n for _, n in live],
                             "isolated": {"achieved": isolated, "frac": isolated / peak_hbm, "launch_ms": patch_ms, "segments_per_launch": pb}},
                "gemm_live": {"kernel": "gemm_tc_kernel + finish_db_kernel", "launch_ms": float(np.mean([t for t, _ in getattr(timed, "gemm_launches", [(0.0, 0)])])),
                              "segments_per_launch": float(np.mean([n for _, n in getattr(timed, "gemm_launches", [(0.0, 0)])])),
                              "how": "per-launch CUDA events on the GEMM stream inside the timed steps"},
                "path_roofline": {"algorithmic_bytes_per_s_audio": ALGO_BYTES_PER_S_AUDIO,
                                  "frac_of_hbm_peak": (value / world) * ALGO_BYTES_PER_S_AUDIO / (peak_hbm * 1e9)},
                "cpu_baseline": cpu_baseline,
                "clocks": clocks.summary()}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_inference_workload(args):
    """`--workload inference`: the structured (multirate) CQT on the inference recipe of tablature_generator.py:599-666 --
    64 songs x 60 s @ 22.05 kHz cut into 3 s segments with 50 % overlap (2 560 segments), C2, 84 bins, hop 512, |C| -> dB
    (ref = max).  Both contractions run on the tcgen05 engine (cqt_structured.cu).  value = seconds of (unique) audio per
    second with the audio resident in HBM; e2e = pinned host int16 PCM in, dB features back."""
    import torch
    import torch.distributed as dist
    from gtc_b200 import synth
    from gtc_b200.inference import TabCnnFrontEnd
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)
    songs, L, seg_len, hop = 64, SR * 60, 66150, 33075
    y = synth.pluck_clips(8, L, sr=SR, seed=2 + rank, device=dev).repeat(8, 1).contiguous().reshape(-1)
    pcm = torch.clamp(torch.round(y * 32768.0), -32768, 32767).to(torch.int16)
    y = pcm.to(torch.float32) / 32768.0
    s1 = np.arange(0, L, hop)
    starts = (np.arange(songs)[:, None] * L + s1[None, :]).reshape(-1)
    valid = np.tile(np.minimum(seg_len, L - s1), songs).astype(np.int32)
    n_seg = len(starts)
    st, va = torch.from_numpy(starts).to(dev), torch.from_numpy(valid).to(dev)
    le = torch.full((n_seg,), seg_len, dtype=torch.int32, device=dev)
    plan = TabCnnFrontEnd().plan
    out = plan.segments_db(y, st, va, le, seg_len)
    h_pcm = torch.empty(pcm.shape, dtype=torch.int16, pin_memory=True)
    h_pcm.copy_(pcm)
    h_out = torch.empty(out.shape, dtype=torch.float32, pin_memory=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    join = []

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        for s_ in join:                                   # side streams of the end-to-end arm: their copies are inside the region
            torch.cuda.current_stream().wait_stream(s_)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # end to end, double-buffered: call i's upload and call i-1's read-back run on their own streams beside call i's kernels
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    join += [s_up, s_dn]
    bufs = [(torch.empty_like(pcm), torch.empty_like(out)) for _ in range(2)]
    state = {"i": 0, "ev_dn": [None, None]}

    def step_e2e():
        cur = torch.cuda.current_stream()
        k = state["i"] & 1
        d_in, d_out = bufs[k]
        if state["ev_dn"][k] is not None:
            s_up.wait_event(state["ev_dn"][k])          # the read-back of call i-2 has left this buffer pair
            cur.wait_event(state["ev_dn"][k])
        s_up.wait_stream(cur)                            # ... and call i-2's kernels have read d_in
        with torch.cuda.stream(s_up):
            d_in.copy_(h_pcm, non_blocking=True)
        cur.wait_stream(s_up)
        plan.segments_db(d_in, st, va, le, seg_len, out=d_out)
        s_dn.wait_stream(cur)
        with torch.cuda.stream(s_dn):
            h_out.copy_(d_out, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s_dn)
        state["ev_dn"][k] = ev
        state["i"] += 1

    with ClockSampler(physical_gpu_index(local_rank), period=args.clock_period) as clocks:
        ms_dev = timed(lambda: plan.segments_db(y, st, va, le, seg_len, out=out))
    ms_e2e = timed(step_e2e)            # (timed() ends with a device synchronise: the last read-back is inside the region)
    secs = songs * 60.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    algo = n_seg * (seg_len * 4 + out.shape[1] * out.shape[2] * 4)          # read every segment's samples once, write its features
    achieved = algo / (ms_dev / args.steps * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps({
            "metric": "seconds-of-audio/sec structured CQT + dB (inference recipe)", "value": world * secs * args.steps / (ms_dev * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "tablature_generator.py:599-666 recipe: %d songs x 60 s @ 22.05 kHz -> %d segments of 3 s (50 %% overlap), "
                                   "librosa.cqt(hop 512, fmin C2, 84 bins) -> amplitude_to_db(ref=max); structured multirate evaluation on tcgen05" % (songs, n_seg),
                       "segments": n_seg, "segment_seconds_per_s": world * n_seg * 3.0 * args.steps / (ms_dev * 1e-3),
                       "cache": "fp16 hi/lo planes of one call (1.4 GB) exceed the 126 MB L2; no flush needed"},
            "e2e": {"value": world * secs * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h_pcm.numel() * 2, "d2h_bytes_per_step": h_out.numel() * 4,
                    "note": "pinned int16 PCM of the 64 songs in, [2560, 84, 130] fp32 dB features back"},
            "gpu_launches": (2 + (plan.n_octaves - 1) + plan.n_octaves + 1) * args.steps,
            "roofline": {"bound": "hbm", "kernel": "whole call (split + 6 decimator GEMMs + 7 response GEMMs + dB pass)", "achieved": achieved,
                         "peak": peak_hbm, "unit": "GB/s", "frac": achieved / peak_hbm, "traffic": None,
                         "algorithmic_bytes_per_call": algo,
                         "note": "algorithmic bytes = each segment's fp32 samples read once + its dB features written; the decimator GEMMs are bound "
                                 "by L2 -> SM operand delivery (ncu: lts 72 %, tensor pipe 30 %), profiles/r02_structured.md"},
            "clocks": clocks.summary()}))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=N_CLIPS)
    ap.add_argument("--chunk-segments", type=int, default=28400, help="upper limit of segments per chunk; the planner ends chunks where the GEMM tile waves are full")
    ap.add_argument("--patch-batch", type=int, default=28400, help="segments per patch launch (one launch per chunk measured fastest)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ragged", action="store_true", help="skip the ragged-duration arm (SURVEY 8d config 2 durations)")
    ap.add_argument("--no-ramp", action="store_true", help="e2e arm: equal chunks instead of the ramped first/last chunks")
    ap.add_argument("--host-audio", default="pcm16", choices=["pcm16", "f32"], help="sample type of the pinned host audio of the e2e arm")
    ap.add_argument("--engine", type=int, default=None, help="GEMM engine: 0 tcgen05 3xTF32, 1 SIMT fp32, 2 tcgen05 fp16x2 (default: library default)")
    ap.add_argument("--coresident", action="store_true", help="experiment: patch kernels on their own stream under the next chunk GEMM (needs -DTC_MAXNREG=152; slower, see profiles/r01j_coresident.md)")
    ap.add_argument("--no-wave-aware", action="store_true", help="plain greedy chunks (largest that fit) instead of full GEMM tile waves")
    ap.add_argument("--no-prefetch", dest="prefetch", action="store_false", help="e2e arm: do not start the next step's host->device copies under the current step")
    ap.add_argument("--ramp-always", dest="flat_after_first", action="store_false", help="e2e arm: prefetched steps also use the ramped chunk plan of a cold step")
    ap.add_argument("--numa-bind", action="store_true", help="bind each rank to the CPUs local to its GPU before allocating pinned memory")
    ap.add_argument("--clock-period", type=float, default=0.02, help="seconds between NVML clock samples during the timed region (0 = off)")
    ap.add_argument("--opt", action="append", default=[], help="library option id=value (gtc_set_option), for experiments")
    ap.add_argument("--overlap", action="store_true", help="run each chunk's patch kernel beside the next chunk's GEMM (slower on B200, see profiles/)")
    ap.add_argument("--patch-ctas-per-sm", type=int, default=4, help="0 = do not limit the patch grid while overlapping")
    ap.add_argument("--gemm-ctas", type=int, default=56, help="SMs given to the persistent tcgen05 GEMM while patches overlap")
    ap.add_argument("--workload", default="front_end", choices=["front_end", "inference"],
                    help="front_end = the headline (BASELINE configs[1] + labels + patches); inference = the structured CQT on the inference recipe")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "inference":
        return run_inference_workload(args)
    return run_cuda_arm(args)


if __name__ == "__main__":
    sys.exit(main())
