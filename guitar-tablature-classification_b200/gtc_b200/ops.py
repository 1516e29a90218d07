"""Torch-facing wrappers over the C ABI: device tensors in, device tensors out, everything on the current stream.

PyTorch is used here only for device memory and streams (plumbing); all arithmetic runs in libgtc.so's kernels.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .cqt_design import CqtRecipe, get_operator, n_frames_of, structured_filters


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts: torch.Tensor) -> None:
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.GtcError("libgtc kernels take CUDA tensors only (there is no CPU fallback)")
        if not t.is_contiguous():
            raise _lib.GtcError("libgtc kernels take contiguous tensors")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            # the launch goes to the CURRENT device's stream: a tensor of another GPU would be read through a foreign pointer
            raise _lib.GtcError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur} "
                                "(wrap the call in torch.cuda.device(tensor.device))")


def segment_counts(clip_lens: Sequence[int], seg_len: int, seg_hop: int) -> np.ndarray:
    """cqt.py:30 -- number of complete windows per clip (never negative)."""
    n = np.asarray(clip_lens, dtype=np.int64)
    return np.maximum(0, (n - seg_len) // seg_hop + 1)


class CqtPlan:
    """Device-resident segment operator for one recipe (replaces the per-call basis rebuild of librosa.cqt)."""

    def __init__(self, recipe: CqtRecipe = CqtRecipe(), device: Optional[int] = None, engine: Optional[int] = None,
                 seg_len: Optional[int] = None, seg_hop: Optional[int] = None, operator: Optional[np.ndarray] = None):
        """``operator``: a (2 * n_bins * T, seg_len) float32 matrix to evaluate instead of the designed one, row =
        (t * n_bins + bin) * 2 + {re, im} -- e.g. measured from the real ``librosa.cqt`` by scripts/pin_with_librosa.py
        (exact soxr behaviour without restating it).  ``GTC_OPERATOR_FILE`` does the same for every plan of that shape."""
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.GtcError("CqtPlan needs a CUDA device (there is no CPU fallback)")
        self.recipe = recipe
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.seg_len = recipe.seg_len if seg_len is None else int(seg_len)
        self.seg_hop = recipe.seg_hop if seg_hop is None else int(seg_hop)
        self.n_bins = recipe.n_bins
        self.n_frames = n_frames_of(recipe, self.seg_len)
        self.engine = _lib.GTC_GEMM_TCGEN05_FP16X2 if engine is None else int(engine)
        op = np.ascontiguousarray(get_operator(recipe, self.seg_len) if operator is None else operator, dtype=np.float32)
        if op.shape != (2 * self.n_bins * self.n_frames, self.seg_len):
            raise _lib.GtcError(f"operator of shape {op.shape}, expected {(2 * self.n_bins * self.n_frames, self.seg_len)}")
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.gtc_cqt_plan_create(C.byref(handle), self.device, self.seg_len, self.seg_hop, self.n_bins,
                                               self.n_frames, op.ctypes.data_as(C.c_void_p), self.engine),
                       "gtc_cqt_plan_create")
        self._h = handle
        self.sm_count = torch.cuda.get_device_properties(self.device).multi_processor_count
        self._ws: Optional[torch.Tensor] = None

    @property
    def parts(self) -> int:
        """Audio rows per segment (P): a segment of seg_len = P * seg_hop samples is P consecutive operand rows."""
        return int(_lib.load().gtc_cqt_plan_parts(self._h))

    def gemm_wave_efficiency(self, n_seg: int, n_clips: int) -> float:
        """Fraction of the persistent tcgen05 GEMM's last wave that is filled for a chunk of ``n_seg`` segments in
        ``n_clips`` clips: tiles = ceil(rows / 128) x ceil(n_out / tile width) are dealt round-robin to one CTA per SM
        (cqt_gemm_tc.cu), so a chunk of 3.4 waves costs as much as one of 4.0.  Used by the chunk planner."""
        from .chunks import wave_efficiency
        return wave_efficiency(n_seg, n_clips, self.parts, 2 * self.n_bins * self.n_frames, self.sm_count)

    def configure(self, option: int, value: int) -> None:
        """Tuning knobs (GTC_OPT_TC_KSPLIT, GTC_OPT_GEMM_MAX_CTAS)."""
        _lib.check(_lib.load().gtc_cqt_plan_configure(self._h, int(option), int(value)), "gtc_cqt_plan_configure")

    def close(self) -> None:
        if getattr(self, "_h", None):
            _lib.load().gtc_cqt_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def workspace_bytes(self, n_seg: int, n_clips: int) -> int:
        need = C.c_size_t()
        _lib.check(_lib.load().gtc_cqt_workspace_bytes(self._h, n_seg, n_clips, C.byref(need)), "gtc_cqt_workspace_bytes")
        return int(need.value)

    def workspace(self, n_seg: int, n_clips: int) -> torch.Tensor:
        need = C.c_size_t()
        _lib.check(_lib.load().gtc_cqt_workspace_bytes(self._h, n_seg, n_clips, C.byref(need)), "gtc_cqt_workspace_bytes")
        if self._ws is None or self._ws.numel() < need.value:
            self._ws = None
            self._ws = torch.empty(need.value, dtype=torch.uint8, device=f"cuda:{self.device}")
        return self._ws

    def offsets(self, clip_lens: Sequence[int]):
        """Host-side window arithmetic (cqt.py:26-30): (clip_off, seg_off) int64 arrays of n_clips+1 entries."""
        lens = np.asarray(clip_lens, dtype=np.int64)
        clip_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        seg_off = np.concatenate([[0], np.cumsum(segment_counts(lens, self.seg_len, self.seg_hop))]).astype(np.int64)
        return clip_off, seg_off

    def segments_db(self, audio: torch.Tensor, clip_off: torch.Tensor, seg_off: torch.Tensor, n_seg: int,
                    out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
        """audio: concatenated mono fp32 clips (device); returns [n_seg, n_bins, T] fp32 dB features
        (== np.stack of what cqt.py:58 saves per segment)."""
        _need_cuda(audio, clip_off, seg_off)
        assert audio.dtype in (torch.float32, torch.int16) and clip_off.dtype == torch.int64 and seg_off.dtype == torch.int64
        n_clips = clip_off.numel() - 1
        if out is None:
            out = torch.empty((n_seg, self.n_bins, self.n_frames), dtype=torch.float32, device=audio.device)
        ws = self.workspace(n_seg, n_clips) if workspace is None else workspace
        r = self.recipe
        if audio.dtype == torch.int16:                       # 16-bit PCM straight from the WAV file
            if n_seg:
                self.frame(audio, clip_off, seg_off, n_seg, ws)
                self.contract_db(clip_off, seg_off, n_seg, out, ws)
            return out
        _lib.check(_lib.load().gtc_cqt_segments_db(self._h, _ptr(audio), _ptr(clip_off), _ptr(seg_off), n_clips, n_seg,
                                                   _ptr(out), _ptr(ws), ws.numel(), r.power, r.amin, r.top_db, r.cut_db,
                                                   r.floor_db, _stream()), "gtc_cqt_segments_db")
        return out

    def frame(self, audio: torch.Tensor, clip_off: torch.Tensor, seg_off: torch.Tensor, n_seg: int, workspace: torch.Tensor) -> None:
        """Stage 1 of segments_db: audio (fp32, or the file's int16 PCM) -> hi/lo operand row matrix in the workspace."""
        _need_cuda(audio, clip_off, seg_off, workspace)
        if audio.dtype == torch.int16:
            _lib.check(_lib.load().gtc_cqt_frame_pcm16(self._h, _ptr(audio), _ptr(clip_off), _ptr(seg_off), clip_off.numel() - 1,
                                                       n_seg, _ptr(workspace), workspace.numel(), _stream()), "gtc_cqt_frame_pcm16")
            return
        assert audio.dtype == torch.float32
        _lib.check(_lib.load().gtc_cqt_frame(self._h, _ptr(audio), _ptr(clip_off), _ptr(seg_off), clip_off.numel() - 1, n_seg,
                                             _ptr(workspace), workspace.numel(), _stream()), "gtc_cqt_frame")

    def contract_db(self, clip_off: torch.Tensor, seg_off: torch.Tensor, n_seg: int, out: torch.Tensor, workspace: torch.Tensor) -> torch.Tensor:
        """Stage 2 of segments_db: tensor-core contraction + dB finish from a framed workspace."""
        _need_cuda(clip_off, seg_off, out, workspace)
        r = self.recipe
        _lib.check(_lib.load().gtc_cqt_contract_db(self._h, _ptr(clip_off), _ptr(seg_off), clip_off.numel() - 1, n_seg, _ptr(out),
                                                   _ptr(workspace), workspace.numel(), r.power, r.amin, r.top_db, r.cut_db,
                                                   r.floor_db, _stream()), "gtc_cqt_contract_db")
        return out

    def segments_complex(self, audio: torch.Tensor, clip_off: torch.Tensor, seg_off: torch.Tensor, n_seg: int) -> torch.Tensor:
        """[n_seg, n_bins, T] complex64 (== librosa.cqt of every segment)."""
        _need_cuda(audio, clip_off, seg_off)
        n_clips = clip_off.numel() - 1
        out = torch.empty((n_seg, self.n_bins, self.n_frames, 2), dtype=torch.float32, device=audio.device)
        ws = self.workspace(n_seg, n_clips)
        _lib.check(_lib.load().gtc_cqt_segments_complex(self._h, _ptr(audio), _ptr(clip_off), _ptr(seg_off), n_clips, n_seg,
                                                        _ptr(out), _ptr(ws), ws.numel(), _stream()), "gtc_cqt_segments_complex")
        return torch.view_as_complex(out)


class StructuredCqtPlan:
    """Multirate evaluation of librosa.cqt for variable-length segments (decimation chain + per-octave filters),
    the way librosa computes it.  Used where the collapsed operator of CqtPlan does not apply: the 3 s inference
    segments of tablature_generator.py:616-620, whole clips, any sample rate."""

    def __init__(self, recipe: CqtRecipe = CqtRecipe(), device: Optional[int] = None):
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.GtcError("StructuredCqtPlan needs a CUDA device (there is no CPU fallback)")
        self.recipe = recipe
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.n_bins, self.n_octaves, self.hop_length = recipe.n_bins, recipe.n_octaves, recipe.hop_length
        filters, n_fft, taps = structured_filters(recipe)
        self.n_fft = n_fft
        filters = np.ascontiguousarray(filters, dtype=np.float32)
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.gtc_scqt_plan_create(C.byref(handle), self.device, self.n_octaves, n_fft, recipe.hop_length,
                                                self.n_bins, filters.shape[1] // 2, filters.ctypes.data_as(C.c_void_p),
                                                taps.ctypes.data_as(C.c_void_p), len(taps)), "gtc_scqt_plan_create")
        self._h = handle
        self._ws: Optional[torch.Tensor] = None

    def close(self) -> None:
        if getattr(self, "_h", None):
            _lib.load().gtc_scqt_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def frames(self, seg_len: int) -> int:
        """librosa's frame count for a signal of ``seg_len`` samples (min over octaves of 1 + len_i // hop_i)."""
        return int(_lib.load().gtc_scqt_frames(int(seg_len), self.hop_length, self.n_octaves))

    def workspace(self, n_seg: int, max_len: int) -> torch.Tensor:
        need = C.c_size_t()
        _lib.check(_lib.load().gtc_scqt_workspace_bytes(self._h, n_seg, max_len, C.byref(need)), "gtc_scqt_workspace_bytes")
        if self._ws is None or self._ws.numel() < need.value:
            self._ws = None
            self._ws = torch.empty(max(need.value, 1024), dtype=torch.uint8, device=f"cuda:{self.device}")
        return self._ws

    def _run(self, audio, seg_start, seg_valid, seg_len, max_len, complex_out, out):
        _need_cuda(audio, seg_start, seg_valid, seg_len)
        assert audio.dtype in (torch.float32, torch.int16)
        assert seg_start.dtype == torch.int64 and seg_valid.dtype == torch.int32 and seg_len.dtype == torch.int32
        n_seg = seg_start.numel()
        assert seg_valid.numel() == n_seg and seg_len.numel() == n_seg
        t_max = self.frames(max_len)
        shape = (n_seg, self.n_bins, t_max, 2) if complex_out else (n_seg, self.n_bins, t_max)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=audio.device)
        assert out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == shape
        ws = self.workspace(n_seg, max_len)
        fmt = _lib.GTC_SAMPLES_PCM16 if audio.dtype == torch.int16 else _lib.GTC_SAMPLES_F32
        r = self.recipe
        if complex_out:
            _lib.check(_lib.load().gtc_scqt_segments_complex(self._h, _ptr(audio), fmt, _ptr(seg_start), _ptr(seg_valid), _ptr(seg_len),
                                                             n_seg, max_len, _ptr(out), _ptr(ws), ws.numel(), _stream()),
                       "gtc_scqt_segments_complex")
            return torch.view_as_complex(out)
        _lib.check(_lib.load().gtc_scqt_segments_db(self._h, _ptr(audio), fmt, _ptr(seg_start), _ptr(seg_valid), _ptr(seg_len), n_seg,
                                                    max_len, _ptr(out), _ptr(ws), ws.numel(), r.power, r.amin, r.top_db, r.cut_db,
                                                    r.floor_db, _stream()), "gtc_scqt_segments_db")
        return out

    def halve_rate(self, audio: torch.Tensor, scale: bool = False) -> torch.Tensor:
        """librosa.resample(y, orig_sr=2k, target_sr=k, res_type='soxr_hq', scale=scale) of one device signal
        (what librosa.load(path, sr=native/2) applies, tablature_generator.py:613,650): fp32 [ceil(n/2)]."""
        _need_cuda(audio)
        assert audio.dim() == 1 and audio.dtype in (torch.float32, torch.int16)
        n = audio.numel()
        out = torch.empty((n + 1) // 2, dtype=torch.float32, device=audio.device)
        if n == 0:
            return out
        st = torch.zeros(1, dtype=torch.int64, device=audio.device)
        ln = torch.full((1,), n, dtype=torch.int32, device=audio.device)
        fmt = _lib.GTC_SAMPLES_PCM16 if audio.dtype == torch.int16 else _lib.GTC_SAMPLES_F32
        _lib.check(_lib.load().gtc_scqt_decimate(self._h, _ptr(audio), fmt, _ptr(st), _ptr(ln), _ptr(ln), 1, n, _ptr(out),
                                                 out.numel(), 1.0 if scale else 0.7071067811865476, _stream()), "gtc_scqt_decimate")
        return out

    def segments_db(self, audio: torch.Tensor, seg_start: torch.Tensor, seg_valid: torch.Tensor, seg_len: torch.Tensor,
                    max_len: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """audio: device fp32 (or int16 PCM) samples; segment s = audio[seg_start[s] : +seg_valid[s]] zero-padded to
        seg_len[s] samples.  Returns dB features [n_seg, n_bins, frames(max_len)] (recipe's power/amin/top_db/cut)."""
        return self._run(audio, seg_start, seg_valid, seg_len, int(max_len), False, out)

    def segments_complex(self, audio: torch.Tensor, seg_start: torch.Tensor, seg_valid: torch.Tensor, seg_len: torch.Tensor,
                         max_len: int) -> torch.Tensor:
        """[n_seg, n_bins, frames(max_len)] complex64 (== librosa.cqt of every segment)."""
        return self._run(audio, seg_start, seg_valid, seg_len, int(max_len), True, None)


def set_option(option: int, value: int) -> None:
    """Process-wide tunables of libgtc (GTC_OPT_PATCH_MAX_CTAS)."""
    _lib.check(_lib.load().gtc_set_option(int(option), int(value)), "gtc_set_option")


def rasterize_tabs(onset, dur, pitch, evt_off, seg_time, seg_off, contour=None, stats: Optional[torch.Tensor] = None,
                   out: Optional[torch.Tensor] = None):
    """Device label rasteriser.  onset/dur/pitch fp64 [n_evt], evt_off int64 [n_clips+1], seg_time fp64 [n_seg],
    seg_off int64 [n_clips+1]; contour = (time, midi, conf, kind, off) or None.
    Returns (labels int8 [n_seg,6,19], stats int64 [3] = total, with_notes, with_first_string)."""
    _need_cuda(onset, dur, pitch, evt_off, seg_time, seg_off)
    n_seg = seg_time.numel()
    n_clips = seg_off.numel() - 1
    if out is None:
        out = torch.empty((n_seg, 6, 19), dtype=torch.int8, device=seg_time.device)
    assert out.dtype == torch.int8 and out.numel() == n_seg * 114 and out.is_contiguous()
    if stats is None:
        stats = torch.zeros(3, dtype=torch.int64, device=seg_time.device)
    ct = cm = cc = ck = co = None
    if contour is not None:
        ct, cm, cc, ck, co = contour
        _need_cuda(ct, cm, cc, ck, co)
        assert ck.dtype == torch.int8 and co.dtype == torch.int64
    _lib.check(_lib.load().gtc_rasterize_tabs(_ptr(onset), _ptr(dur), _ptr(pitch), _ptr(evt_off), _ptr(ct), _ptr(cm),
                                              _ptr(cc), _ptr(ck), _ptr(co), _ptr(seg_time), _ptr(seg_off), n_clips, n_seg,
                                              _ptr(out), _ptr(stats), _stream()), "gtc_rasterize_tabs")
    return out, stats


def labels_argmax(tabs: torch.Tensor, index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(n_total,6,19) int8 [+ index] -> (n,6) int64 (my_dataloader.py:40-44)."""
    _need_cuda(tabs, index)
    assert tabs.dtype == torch.int8 and (index is None or index.dtype == torch.int64)
    n = tabs.shape[0] if index is None else index.numel()
    out = torch.empty((n, 6), dtype=torch.int64, device=tabs.device)
    _lib.check(_lib.load().gtc_labels_argmax(_ptr(tabs), _ptr(index), n, _ptr(out), _stream()), "gtc_labels_argmax")
    return out


def labels_vit_heads(tabs: torch.Tensor, index: Optional[torch.Tensor] = None):
    """(n_total,6,19) int8 [+ index] -> list of six (n,19) int64 tensors (ViT_dataloader.py:54 after default collate)."""
    _need_cuda(tabs, index)
    assert tabs.dtype == torch.int8 and (index is None or index.dtype == torch.int64)
    n = tabs.shape[0] if index is None else index.numel()
    out = torch.empty((6, n, 19), dtype=torch.int64, device=tabs.device)
    _lib.check(_lib.load().gtc_labels_vit_heads(_ptr(tabs), _ptr(index), n, _ptr(out), _stream()), "gtc_labels_vit_heads")
    return [out[i] for i in range(6)]


def patches(db: torch.Tensor, index: Optional[torch.Tensor] = None, img_size=(224, 224), mode: int = _lib.GTC_PATCH_VIT,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """db [n_total, n_bins, T] fp32 dB features -> [n, 3, H, W] fp32 patches (ViT or CNN contract)."""
    _need_cuda(db, index)
    assert db.dtype == torch.float32 and db.dim() == 3
    n = db.shape[0] if index is None else index.numel()
    if index is not None:
        assert index.dtype == torch.int64
    h, w = int(img_size[0]), int(img_size[1])
    if out is None:
        out = torch.empty((n, 3, h, w), dtype=torch.float32, device=db.device)
    if tuple(out.shape) != (n, 3, h, w) or out.dtype != torch.float32 or out.device != db.device:
        raise _lib.GtcError(f"patches: out must be float32 {(n, 3, h, w)} on {db.device}, got {out.dtype} {tuple(out.shape)} on {out.device}")
    _need_cuda(out)
    _lib.check(_lib.load().gtc_patches(_ptr(db), _ptr(index), n, db.shape[1], db.shape[2], h, w, int(mode), _ptr(out),
                                       _stream()), "gtc_patches")
    return out


IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)        # my_dataloader.py:20


def patches_rgb8(rgb: torch.Tensor, index: Optional[torch.Tensor] = None, mean=IMAGENET_MEAN, std=IMAGENET_STD,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rgb [n_total, H, W, 3] uint8 (decoded + resized pictures, PIL's byte order) -> [n, 3, H, W] fp32
    ((x / 255) - mean) / std : ToTensor + Normalize of my_dataloader.py:19-20, bit for bit."""
    _need_cuda(rgb, index)
    assert rgb.dtype == torch.uint8 and rgb.dim() == 4 and rgb.shape[3] == 3 and rgb.is_contiguous()
    n = rgb.shape[0] if index is None else index.numel()
    if index is not None:
        assert index.dtype == torch.int64
    h, w = int(rgb.shape[1]), int(rgb.shape[2])
    if out is None:
        out = torch.empty((n, 3, h, w), dtype=torch.float32, device=rgb.device)
    _lib.check(_lib.load().gtc_patches_rgb8(_ptr(rgb), _ptr(index), n, h, w, *[float(v) for v in mean], *[float(v) for v in std],
                                            _ptr(out), _stream()), "gtc_patches_rgb8")
    return out
