"""Shard pipeline: audio + note events of many clips -> dB features, tablature labels and patch batches.

This is the index-aligned in-memory path: clip ``c`` yields ``n_seg[c]`` feature segments
(cqt.py:26-45), one label per segment at ``t_i = (i + 0.5) * duration / n_seg`` (jam_to_tablature.py:273-274 with
``num_images = n_seg``) and one (3, H, W) patch per segment (ViT_dataloader.py:27-51 or the CNN contract).

Clips are processed in chunks of whole clips (gtc_b200/chunks.py: each chunk fills whole waves of GEMM tiles).  Streams:
``s_stage`` host->device staging of whole shards (two slots; never joined, so the next shard can be prefetched under the
current one), ``s_pre`` framing + label rasterisation (beside the previous chunk's GEMM), ``s_comp`` GEMM -> dB finish ->
patches back to back, ``s_out`` device->host copies of features and labels.  With ``device_inputs=True`` the audio and
events are already resident in HBM and no copies are issued (the kernel-only number of bench.py).  DESIGN.md section 4.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np
import torch

from . import _lib, ops
from .chunks import plan_bounds, rows_per_wave
from .cqt_design import CqtRecipe


@dataclass
class ShardInputs:
    """One shard.  ``audio`` is the concatenation of all clips (1-D; fp32 as librosa.load returns it, or the WAV files'
    own int16 PCM, converted on the device exactly as librosa does): pinned host memory for the end-to-end path or a
    CUDA tensor for the device-resident path.  Events are fp64 arrays concatenated per clip."""
    audio: torch.Tensor            # [n_samples] fp32 or int16
    clip_lens: np.ndarray          # [n_clips] int64 (host)
    events: torch.Tensor           # [3, n_evt] fp64 rows = onset, duration, pitch (pinned host or CUDA, like audio)
    evt_off: np.ndarray            # [n_clips+1] int64 (host)
    sr: float = 22050.0


@dataclass
class ShardOutputs:
    """Results of one ``FrontEnd.run``.  ``db`` / ``tabs`` are views into buffers the FrontEnd owns and REUSES: they are valid
    (in stream order on the caller's current stream; after a synchronize for the pinned host copies) until the next ``run``
    of the same FrontEnd -- copy what must outlive it.  Patch batches are only ever handed to ``consumer``."""
    db: Optional[torch.Tensor] = None          # [n_seg, n_bins, T] fp32 (host pinned for e2e, device otherwise)
    tabs: Optional[torch.Tensor] = None        # [n_seg, 6, 19] int8
    stats: Optional[np.ndarray] = None         # total, with_notes, with_first_string
    n_seg: int = 0
    seconds_of_audio: float = 0.0
    launches: int = 0
    h2d_bytes: int = 0
    d2h_bytes: int = 0


@dataclass
class _Chunk:
    c0: int
    c1: int
    s0: int          # first sample
    s1: int
    g0: int          # first segment
    g1: int
    e0: int          # first event
    e1: int
    clip_off: np.ndarray = field(default=None)
    seg_off: np.ndarray = field(default=None)
    evt_off: np.ndarray = field(default=None)
    seg_time: np.ndarray = field(default=None)


class _Staging:
    """One whole-shard staging slot of the host-input path: the shard's audio and note events are copied into HBM by a
    train of piece copies on a stream of their own; consumers wait for the event of the piece they need."""

    def __init__(self, fe: "FrontEnd", slot: int, inp: ShardInputs, chunks: List["_Chunk"], prev: Optional["_Staging"]):
        n_clips = chunks[-1].c1
        self.fe, self.slot, self.inp, self.n_clips, self.chunks = fe, slot, inp, int(n_clips), chunks
        self.key = FrontEnd._stage_key(inp, chunks)
        self.consumed = False
        self.clip_off = np.concatenate([[0], np.cumsum(np.asarray(inp.clip_lens[:n_clips], dtype=np.int64))])
        self.n_samples, self.n_evt = int(self.clip_off[-1]), int(inp.evt_off[n_clips])
        self.audio = fe._buf(f"audio_all{slot}", (self.n_samples,), inp.audio.dtype)
        self.events = fe._buf(f"ev_all{slot}", (3, max(1, self.n_evt)), torch.float64)
        self.piece_ev, self.piece_end = [], []
        # chunk metadata (offsets, label times) travels at the head of the train: a copy queued later on another stream
        # would wait behind the whole train in the copy engine's queue
        meta_np, time_np = FrontEnd._chunk_metadata(chunks)
        if prev is not None and prev.meta_ev is not None:
            prev.meta_ev.synchronize()           # the slot's previous upload has read the pinned staging buffers
        h_meta = fe._buf(f"meta_host{slot}", (meta_np.size,), torch.int64, pinned=True)
        h_time = fe._buf(f"time_host{slot}", (max(1, time_np.size),), torch.float64, pinned=True)
        h_meta.numpy()[:] = meta_np
        h_time.numpy()[: time_np.size] = time_np
        self.d_meta = fe._buf(f"meta_dev{slot}", (meta_np.size,), torch.int64)
        self.d_time = fe._buf(f"time_dev{slot}", (max(1, time_np.size),), torch.float64)
        # the staging stream is never joined into the caller's stream: tell the caching allocator who else uses the blocks
        for t in (self.audio, self.events, self.d_meta, self.d_time):
            t.record_stream(fe.s_stage)
        with torch.cuda.stream(fe.s_stage):
            self.d_meta.copy_(h_meta, non_blocking=True)
            self.d_time.copy_(h_time, non_blocking=True)
            self.meta_ev = torch.cuda.Event()
            self.meta_ev.record(fe.s_stage)
        self.nbytes = self.n_samples * inp.audio.element_size() + self.n_evt * 24 + meta_np.nbytes + time_np.nbytes

    def stage_until(self, c_goal: int) -> None:
        """Enqueue piece copies until clip ``c_goal`` is covered (a few chunks ahead of the kernels that read them, so
        the first chunk's kernels are not queued behind 45 copy calls)."""
        fe, inp = self.fe, self.inp
        c = self.piece_end[-1] if self.piece_end else 0
        with torch.cuda.stream(fe.s_stage):
            while c < min(c_goal, self.n_clips):
                c_next = min(self.n_clips, c + fe.stage_piece_clips)
                a0, a1 = int(self.clip_off[c]), int(self.clip_off[c_next])
                if c == 0:
                    fe._mark("h2d<", fe.s_stage, "copy")
                self.audio[a0:a1].copy_(inp.audio[a0:a1], non_blocking=True)
                if c == 0:                                        # all note events ride behind the first piece
                    for j in range(3):
                        self.events[j, :self.n_evt].copy_(inp.events[j, :self.n_evt], non_blocking=True)
                e = torch.cuda.Event()
                e.record(fe.s_stage)
                self.piece_ev.append(e)
                self.piece_end.append(c_next)
                c = c_next
                if c == self.n_clips:
                    fe._mark("h2d>", fe.s_stage, "copy")

    def event_for_clip_end(self, c1: int):
        return self.piece_ev[int(np.searchsorted(np.asarray(self.piece_end), c1))]   # first piece ending at or after c1


class FrontEnd:
    def __init__(self, recipe: CqtRecipe = CqtRecipe(), device: Optional[int] = None, engine: Optional[int] = None,
                 patch_mode: int = _lib.GTC_PATCH_VIT, img_size=(224, 224), chunk_segments: int = 28400,
                 patch_batch: int = 28400, overlap: bool = False, gemm_ctas: int = 64,
                 patch_ctas_per_sm: int = 4, coresident: bool = False, wave_aware: bool = True):
        """``coresident=True``: the patch kernels run on their own lower-priority stream, gated only by their chunk's dB
        features, while the GEMM stream goes on with the next chunks, so that (with libgtc built -DTC_MAXNREG=152) one
        patch CTA per SM runs beside the GEMM CTA.  ``overlap=True`` is the older experiment with the two kernels on
        DISJOINT SMs.  Both were measured on B200 and are SLOWER than running the kernels back to back
        (profiles/r01j_coresident.md, profiles/r01_overlap_sweep.md): the GEMM re-reads its operator tiles from L2 for
        every M tile and the patch kernel pushes 7 TB/s of stores through the same L2, so they do not hide each other.
        The default keeps them sequential."""
        self.recipe = recipe
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.dev = torch.device(f"cuda:{self.device}")
        self.plan = ops.CqtPlan(recipe, device=self.device, engine=engine)
        self.patch_mode = patch_mode
        self.img_size = (int(img_size[0]), int(img_size[1]))
        self.chunk_segments = int(chunk_segments)
        self.patch_batch = int(patch_batch)
        self.overlap = bool(overlap)
        self.coresident = bool(coresident) and not self.overlap
        self.wave_aware = bool(wave_aware)
        self.gemm_ctas = int(gemm_ctas) if self.overlap else 0
        if self.gemm_ctas > 0:
            self.plan.configure(_lib.GTC_OPT_GEMM_MAX_CTAS, self.gemm_ctas)
            if patch_ctas_per_sm > 0:
                # the patch CTAs must not back-fill the SMs the persistent GEMM runs on (see run())
                ops.set_option(_lib.GTC_OPT_PATCH_MAX_CTAS, (self.plan.sm_count - self.gemm_ctas) * int(patch_ctas_per_sm))
        with torch.cuda.device(self.device):
            self.s_copy, self.s_out, self.s_pre = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
            self.s_stage = torch.cuda.Stream()                        # whole-shard input staging; never joined by run()
            self.s_comp = torch.cuda.Stream(priority=-1)              # GEMM path: scheduled ahead of pending patch CTAs
            self.s_patch = torch.cuda.Stream(priority=0)
        self._bufs = {}
        self.patch_events = None       # set to a list to collect (start_event, end_event, n_segments) of every patch launch
        self.gemm_events = None        # same for every GEMM + dB-finish pair
        self._meta_plan, self._meta_ev, self._meta_sizes = None, None, None
        self.stage_bytes_limit = 4 << 30   # host-input shards up to this size are staged whole in HBM (else per-chunk double buffers)
        self.stage_piece_clips = 8         # clips per host->device copy of the staging train
        self._stages = [None, None]        # staging slots: the current shard's and the prefetched next one's
        self._stage_free = [None, None]    # event after which a slot's buffers are no longer read
        self._stage_turn = 0
        self.trace = None              # set to a list to collect (label, stream name, event) marks: scripts/timeline.py

    # ------------------------------------------------------------------ host-side planning (integer arithmetic only)
    def plan_chunks(self, inp: ShardInputs, ramp: bool = False) -> List[_Chunk]:
        """Cut the shard into chunks of whole clips, at most ``chunk_segments`` segments each.  ``ramp=True`` (the host-input
        path) makes the first chunks small (1/8 of the limit, growing x1.5), so the un-overlapped head of the pipeline
        (first host->device copy) is short and the compute stream never waits for audio that is still on the bus."""
        r = self.recipe
        lens = np.asarray(inp.clip_lens, dtype=np.int64)
        nseg = ops.segment_counts(lens, self.plan.seg_len, self.plan.seg_hop)
        clip_off = np.concatenate([[0], np.cumsum(lens)])
        seg_off = np.concatenate([[0], np.cumsum(nseg)])
        full = self.chunk_segments
        eff = (lambda n, c: self.plan.gemm_wave_efficiency(n, c)) if self.wave_aware else None
        chunks = []
        wave = rows_per_wave(2 * self.plan.n_bins * self.plan.n_frames, self.plan.sm_count)
        for c0, c1 in plan_bounds(nseg, full, ramp=ramp, efficiency=eff, wave_rows=wave):
            ch = _Chunk(c0, c1, int(clip_off[c0]), int(clip_off[c1]), int(seg_off[c0]), int(seg_off[c1]),
                        int(inp.evt_off[c0]), int(inp.evt_off[c1]))
            assert ch.g1 - ch.g0 <= max(self.chunk_segments, int(nseg[c0]))
            ch.clip_off = (clip_off[c0:c1 + 1] - clip_off[c0]).astype(np.int64)
            ch.seg_off = (seg_off[c0:c1 + 1] - seg_off[c0]).astype(np.int64)
            ch.evt_off = (np.asarray(inp.evt_off[c0:c1 + 1]) - inp.evt_off[c0]).astype(np.int64)
            times = []
            for c in range(c0, c1):
                n = int(nseg[c])
                if n:
                    duration = float(lens[c]) / float(inp.sr)                 # librosa.get_duration(y, sr)
                    times.append((np.arange(n, dtype=np.float64) + 0.5) * (duration / n))
            ch.seg_time = np.concatenate(times) if times else np.zeros(0, dtype=np.float64)
            chunks.append(ch)
        return chunks

    @staticmethod
    def _stage_key(inp: ShardInputs, chunks: List[_Chunk]):
        return (inp.audio.data_ptr(), inp.audio.dtype, inp.events.data_ptr(), tuple((c.c0, c.c1, c.s1, c.e1) for c in chunks))

    @staticmethod
    def _chunk_metadata(chunks: List[_Chunk]):
        meta_np = np.concatenate([np.concatenate([c.clip_off, c.seg_off, c.evt_off]) for c in chunks]).astype(np.int64) \
            if chunks else np.zeros(1, np.int64)
        time_np = np.concatenate([c.seg_time for c in chunks]) if chunks else np.zeros(1)
        return meta_np, time_np

    def _stage_slot_for(self, inp: ShardInputs, chunks: List[_Chunk], after_current_stream: bool) -> _Staging:
        """The staging slot that holds (or is receiving) ``inp`` cut into ``chunks``; a new train is started when there
        is none."""
        key = self._stage_key(inp, chunks)
        for st in self._stages:
            if st is not None and st.key == key and not st.consumed:
                return st
        slot = self._stage_turn
        self._stage_turn ^= 1
        if self._stage_free[slot] is not None:
            self.s_stage.wait_event(self._stage_free[slot])      # the run that last read this slot has finished
        if after_current_stream:
            self.s_stage.wait_stream(torch.cuda.current_stream())
        st = _Staging(self, slot, inp, chunks, self._stages[slot])
        self._stages[slot] = st
        return st

    def _stageable(self, inp: ShardInputs, chunks: List[_Chunk]) -> bool:
        return bool(chunks) and chunks[-1].s1 * inp.audio.element_size() <= self.stage_bytes_limit

    def prefetch(self, inp: ShardInputs, chunks: Optional[List[_Chunk]] = None) -> None:
        """Start copying the NEXT shard's pinned host inputs (and its chunk metadata) into the other staging slot now, on
        the staging stream: the copies run under the current shard's kernels and the next ``run(inp, chunks=chunks)``
        finds its audio resident.  ``inp``'s host buffers must not change until that run.
        (``run(..., next_inp=inp, next_chunks=chunks)`` calls this after its own copies are queued.)"""
        chunks = self.plan_chunks(inp, ramp=True) if chunks is None else chunks
        if not self._stageable(inp, chunks):
            return
        with torch.cuda.device(self.device):
            st = self._stage_slot_for(inp, chunks, after_current_stream=False)
            st.stage_until(st.n_clips)

    def _mark(self, label, stream, name):
        if self.trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            self.trace.append((label, name, e, time.perf_counter()))

    def _buf(self, name, shape, dtype, pinned=False):
        key = (name, pinned)
        n = int(np.prod(shape))
        t = self._bufs.get(key)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(n, dtype=dtype, pin_memory=True) if pinned else torch.empty(n, dtype=dtype, device=self.dev)
            self._bufs[key] = t
        return t[:n].view(shape)

    # ------------------------------------------------------------------ execution
    def run(self, inp: ShardInputs, device_inputs: bool = False, want_host_outputs: bool = True,
            emit_patches: bool = True, consumer: Optional[Callable] = None, chunks: Optional[List[_Chunk]] = None,
            next_inp: Optional[ShardInputs] = None, next_chunks: Optional[List[_Chunk]] = None) -> ShardOutputs:
        """Process one shard.  ``consumer(patches, tabs_batch, first_segment_index)`` is called on the compute stream
        for every patch batch (the training engine's input); without it patches are produced into a ring and dropped.
        ``next_inp`` (host inputs): the shard the NEXT call will process; its host->device copies are queued behind this
        shard's, so the copy engine runs on under this shard's last kernels (see ``prefetch``)."""
        plan, dev = self.plan, self.dev
        chunks = self.plan_chunks(inp, ramp=not device_inputs) if chunks is None else chunks
        n_seg = chunks[-1].g1 if chunks else 0
        out = ShardOutputs(n_seg=n_seg, seconds_of_audio=float(np.sum(inp.clip_lens)) / float(inp.sr))
        nb, T = plan.n_bins, plan.n_frames
        host_out = want_host_outputs and not device_inputs
        if host_out:
            out.db = self._buf("db_host", (n_seg, nb, T), torch.float32, pinned=True)
            out.tabs = self._buf("tabs_host", (n_seg, 6, 19), torch.int8, pinned=True)
        else:
            out.db = self._buf("db_all", (n_seg, nb, T), torch.float32)
            out.tabs = self._buf("tabs_all", (n_seg, 6, 19), torch.int8)
        stats = self._buf("stats", (3,), torch.int64)
        max_samples = max((c.s1 - c.s0 for c in chunks), default=1)
        max_seg = max((c.g1 - c.g0 for c in chunks), default=1)
        max_evt = max((c.e1 - c.e0 for c in chunks), default=1)
        max_clips = max((c.c1 - c.c0 for c in chunks), default=1)
        ws_bytes = plan.workspace_bytes(max_seg, max_clips)
        ws2 = [self._buf(f"ws{j}", (ws_bytes,), torch.uint8) for j in range(2)]      # framed operands, double-buffered
        ev_pre = [None, None]
        ev_ws = [None, None]                                                          # GEMM that last read ws2[b]
        pb = min(self.patch_batch, max(1, max_seg))
        ev_free = [[], []]
        # host inputs: the whole shard is staged in HBM by ONE train of copies in pieces of a few clips (chunk metadata at
        # its head), on a stream of its own and independent of the compute chunks (a 360-clip shard is 0.48 GB of int16
        # PCM).  The copy engine never waits for a staging buffer to be released by a framing kernel, which with two
        # per-chunk buffers it did whenever compute lagged (profiles/r01k_timeline_host.log).
        staged = (not device_inputs) and self._stageable(inp, chunks)
        # otherwise all chunk metadata (offsets, label times) goes up in one copy from pinned memory; with device-resident
        # inputs a chunk plan that is passed in again (same list object: epochs over the same shard) keeps its device copy
        reuse_meta = staged or (device_inputs and self._meta_plan is not None and self._meta_plan is chunks)
        if not reuse_meta:
            meta_np, time_np = self._chunk_metadata(chunks)
            if self._meta_ev is not None:
                self._meta_ev.synchronize()      # the previous run's async upload has read the pinned staging buffers
            h_meta = self._buf("meta_host", (meta_np.size,), torch.int64, pinned=True)
            h_time = self._buf("time_host", (max(1, time_np.size),), torch.float64, pinned=True)
            h_meta.numpy()[:] = meta_np
            h_time.numpy()[: time_np.size] = time_np
            self._meta_sizes = (meta_np.size, max(1, time_np.size), meta_np.nbytes + time_np.nbytes)
        if not staged:
            d_meta_all = self._buf("meta_dev", (self._meta_sizes[0],), torch.int64)
            d_time_all = self._buf("time_dev", (self._meta_sizes[1],), torch.float64)
        with torch.cuda.device(self.device):
            stats.zero_()
            self.s_copy.wait_stream(torch.cuda.current_stream())
            self.s_pre.wait_stream(torch.cuda.current_stream())
            self.s_comp.wait_stream(torch.cuda.current_stream())
            self.s_out.wait_stream(torch.cuda.current_stream())
            self.s_patch.wait_stream(torch.cuda.current_stream())
            if not reuse_meta:
                with torch.cuda.stream(self.s_copy):
                    d_meta_all.copy_(h_meta, non_blocking=True)
                    d_time_all.copy_(h_time, non_blocking=True)
                    self._meta_ev = torch.cuda.Event()
                    self._meta_ev.record(self.s_copy)
                self._meta_plan = chunks
            if not device_inputs and not staged:
                out.h2d_bytes += self._meta_sizes[2]      # host-input runs upload their metadata every time
            meta_ev = self._meta_ev
            stage = None
            if staged:
                stage = self._stage_slot_for(inp, chunks, after_current_stream=True)
                d_meta_all, d_time_all, meta_ev = stage.d_meta, stage.d_time, stage.meta_ev
                out.h2d_bytes += stage.nbytes
            def emit(job, gate, s_p):
                jb, jch, j_db, j_tabs, _ = job
                jng = jch.g1 - jch.g0
                with torch.cuda.stream(s_p):
                    s_p.wait_event(gate)
                    for j, p0 in enumerate(range(0, jng, pb)):
                        p1 = min(jng, p0 + pb)
                        ring = self._buf(f"patch{j & 1}", (pb, 3) + self.img_size, torch.float32)[: p1 - p0]
                        if self.patch_events is not None:
                            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            t0.record(s_p)
                        self._mark(f"patch{jch.c0}<", s_p, "patch")
                        ops.patches(j_db[p0:p1], img_size=self.img_size, mode=self.patch_mode, out=ring)
                        self._mark(f"patch{jch.c0}>", s_p, "patch")
                        if self.patch_events is not None:
                            t1.record(s_p)
                            self.patch_events.append((t0, t1, p1 - p0))
                        out.launches += 1
                        if consumer is not None:
                            consumer(ring, j_tabs[p0:p1], jch.g0 + p0)
                    ev_p = torch.cuda.Event()
                    ev_p.record(s_p)
                ev_free[jb].append(ev_p)

            pending = None
            m_at = 0
            for k, ch in enumerate(chunks):
                b = k & 1
                ns, ne, nc, ng = ch.s1 - ch.s0, ch.e1 - ch.e0, ch.c1 - ch.c0, ch.g1 - ch.g0
                d_meta = d_meta_all[m_at: m_at + 3 * (nc + 1)]
                m_at += 3 * (nc + 1)
                d_time = d_time_all[ch.g0:ch.g1]
                # ---- stage inputs
                if staged:
                    stage.stage_until(chunks[min(k + 3, len(chunks) - 1)].c1)
                    d_audio = stage.audio[ch.s0:ch.s1]
                    d_on, d_du, d_pi = (stage.events[j, ch.e0:ch.e1] for j in range(3))
                    ev_h2d = stage.event_for_clip_end(ch.c1)
                else:
                  with torch.cuda.stream(self.s_copy):
                    if ev_pre[b] is not None:
                        self.s_copy.wait_event(ev_pre[b])             # chunk k-2's framing / label kernels have read audio{b}, ev{b}
                    self._mark(f"h2d{k}<", self.s_copy, "copy")
                    if device_inputs:
                        d_audio = inp.audio[ch.s0:ch.s1]
                        d_ev = inp.events[:, ch.e0:ch.e1]
                        d_on, d_du, d_pi = d_ev[0], d_ev[1], d_ev[2]
                    else:
                        d_audio = self._buf(f"audio{b}", (max_samples,), inp.audio.dtype)[:ns]
                        d_audio.copy_(inp.audio[ch.s0:ch.s1], non_blocking=True)
                        d_evb = self._buf(f"ev{b}", (3, max_evt), torch.float64)
                        for j in range(3):                            # contiguous row slices -> plain async memcpys
                            d_evb[j, :ne].copy_(inp.events[j, ch.e0:ch.e1], non_blocking=True)
                        d_on, d_du, d_pi = d_evb[0, :ne], d_evb[1, :ne], d_evb[2, :ne]
                        out.h2d_bytes += ns * inp.audio.element_size() + ne * 24
                    self._mark(f"h2d{k}>", self.s_copy, "copy")
                    ev_h2d = torch.cuda.Event()
                    ev_h2d.record(self.s_copy)
                with torch.cuda.stream(self.s_pre):
                    self.s_pre.wait_event(ev_h2d)
                    if meta_ev is not None:
                        self.s_pre.wait_event(meta_ev)                # chunk offsets / label times
                    d_clip_off, d_seg_off, d_evt_off = d_meta[: nc + 1], d_meta[nc + 1: 2 * nc + 2], d_meta[2 * nc + 2:]
                    d_db = out.db[ch.g0:ch.g1] if not host_out else self._buf(f"db{b}", (max_seg, nb, T), torch.float32)[:ng]
                    d_tabs = out.tabs[ch.g0:ch.g1] if not host_out else self._buf(f"tabs{b}", (max_seg, 6, 19), torch.int8)[:ng]
                    # ---- framing (HBM-bound) and label rasterisation (latency-bound) run on their own stream and start
                    #      when chunk k-2's patches are done, i.e. together with chunk k-1's GEMM.  Measured on B200
                    #      (profiles/r01j_timelines.md): beside the GEMM the framing kernel costs 0.035 ms per chunk;
                    #      released earlier (right after the GEMM of chunk k-2, under the patch stores) it slows the
                    #      store stream by 0.09 ms per chunk and more, so the patch-completion wait below stays.
                    if ng:
                        if ev_ws[b] is not None:
                            self.s_pre.wait_event(ev_ws[b])           # the GEMM of chunk k-2 has consumed ws2[b]
                        for e in ev_free[b]:
                            self.s_pre.wait_event(e)                  # chunk k-2's patches / D2H released tabs{b}
                        self._mark(f"frame{k}<", self.s_pre, "pre")
                        plan.frame(d_audio, d_clip_off, d_seg_off, ng, ws2[b])
                        self._mark(f"frame{k}>", self.s_pre, "pre")
                        ev_fr = torch.cuda.Event()
                        ev_fr.record(self.s_pre)
                        ops.rasterize_tabs(d_on, d_du, d_pi, d_evt_off, d_time, d_seg_off, out=d_tabs, stats=stats)
                        self._mark(f"labels{k}>", self.s_pre, "pre")
                        out.launches += 2
                    ev_in = torch.cuda.Event()
                    ev_in.record(self.s_pre)
                ev_pre[b] = ev_in
                # ---- tensor-core contraction + dB finish
                with torch.cuda.stream(self.s_comp):
                    self.s_comp.wait_event(ev_fr if ng else ev_in)    # the GEMM needs the framed rows, not the labels
                    for e in ev_free[b]:
                        self.s_comp.wait_event(e)                     # chunk k-2's patches / D2H released db{b}
                    ev_g = ev_in if self.overlap else None
                    if ng:
                        if self.gemm_events is not None:
                            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            t0.record(self.s_comp)
                        self._mark(f"gemm{k}<", self.s_comp, "comp")
                        plan.contract_db(d_clip_off, d_seg_off, ng, d_db, ws2[b])
                        self._mark(f"gemm{k}>", self.s_comp, "comp")
                        if self.gemm_events is not None:
                            t1.record(self.s_comp)
                            self.gemm_events.append((t0, t1, ng))
                        out.launches += 2
                    self.s_comp.wait_event(ev_in)                     # labels of this chunk: patch consumers and D2H read them
                    ev_k = torch.cuda.Event()
                    ev_k.record(self.s_comp)
                ev_ws[b] = ev_k
                ev_free[b] = []
                # ---- patches.  Overlap mode: the store-bound patch kernel of chunk k-1 is released on its own
                #      (low-priority) stream at the moment the tensor-core GEMM of chunk k becomes runnable on the
                #      high-priority stream: the GEMM takes its `gemm_ctas` SMs first, the patch CTAs fill the rest.
                if emit_patches and self.overlap:
                    if pending is not None:
                        emit(pending, ev_g if ev_g is not None else ev_k, self.s_patch)
                    pending = (b, ch, d_db, d_tabs, ev_k) if ng else None
                elif emit_patches and ng:
                    emit((b, ch, d_db, d_tabs, ev_k), ev_k, self.s_patch if self.coresident else self.s_comp)
                # ---- results back to the host
                if host_out:
                    with torch.cuda.stream(self.s_out):
                        self.s_out.wait_event(ev_k)
                        self._mark(f"d2h{k}<", self.s_out, "out")
                        out.db[ch.g0:ch.g1].copy_(d_db, non_blocking=True)
                        out.tabs[ch.g0:ch.g1].copy_(d_tabs, non_blocking=True)
                        self._mark(f"d2h{k}>", self.s_out, "out")
                        out.d2h_bytes += ng * (nb * T * 4 + 114)
                        ev_o = torch.cuda.Event()
                        ev_o.record(self.s_out)
                    ev_free[b].append(ev_o)
            if pending is not None:                                   # last chunk's patches: nothing left to overlap with
                emit(pending, pending[4], self.s_patch)
            # ---- stats (tiny) and join
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_stream(self.s_comp)
                if not device_inputs:                               # host-input runs always return the label stats
                    h_stats = self._buf("stats_host", (3,), torch.int64, pinned=True)
                    h_stats.copy_(stats, non_blocking=True)
                    out.d2h_bytes += 24
            torch.cuda.current_stream().wait_stream(self.s_out)
            torch.cuda.current_stream().wait_stream(self.s_comp)
            torch.cuda.current_stream().wait_stream(self.s_copy)
            torch.cuda.current_stream().wait_stream(self.s_pre)
            torch.cuda.current_stream().wait_stream(self.s_patch)
            if stage is not None:
                # every kernel that read this slot is ordered before this point of the current stream; the staging
                # stream itself is NOT joined, so a prefetch of the next shard keeps copying past the end of this call
                stage.consumed = True
                self._stage_free[stage.slot] = torch.cuda.Event()
                self._stage_free[stage.slot].record(torch.cuda.current_stream())
            if next_inp is not None and not device_inputs:
                self.prefetch(next_inp, next_chunks)
        self._last_stats = (stats, self._bufs.get(("stats_host", True)) if not device_inputs else None)
        return out

    def stats(self) -> np.ndarray:
        """Synchronise and return (total, with_notes, with_first_string) of the last run."""
        torch.cuda.synchronize(self.device)
        dev_stats, host_stats = self._last_stats
        return host_stats[:3].numpy().copy() if host_stats is not None else dev_stats.cpu().numpy()
