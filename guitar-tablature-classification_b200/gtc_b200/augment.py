"""Device batch augmentation / normalisation with the reference's function names and signatures
(/root/reference/ViT_engine.py:28-117).  Every function is one launch of libgtc's fused streaming kernel
(csrc/augment.cu); ``augment_batch`` draws its decisions from Python's ``random`` in exactly the order the reference
does (so ``random.seed(k)`` selects the same ops, shift and masks) and then applies the whole composition in ONE pass
over the batch instead of one torch op chain per augmentation.

Tensors are CUDA float32 ``(B, C, dim2, dim3)``; the reference calls dim2 "time" and dim3 "frequency"
(ViT_engine.py:30,51) whatever the caller put there.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import random
from typing import Optional, Sequence

import torch

from . import _lib
from .ops import _need_cuda, _ptr, _stream

TIME_SHIFT, NOISE, FREQ_MASK, TIME_MASK = (_lib.GTC_AUG_TIME_SHIFT, _lib.GTC_AUG_NOISE, _lib.GTC_AUG_FREQ_MASK,
                                           _lib.GTC_AUG_TIME_MASK)


def _noise_seed(like: torch.Tensor) -> int:
    """Philox seed of the noise op, taken from the CUDA default generator of ``like``'s device: the reference's add_noise
    calls ``torch.randn_like`` on the device tensor (ViT_engine.py:46), i.e. it consumes the CUDA generator and leaves the CPU
    generator -- which orders the DataLoader's shuffles (loaders.sampler_order) -- alone.  No host synchronisation: the
    generator's (seed, offset) pair is read and the offset advanced by what randn_like would consume."""
    g = torch.cuda.default_generators[like.device.index]
    off = int(g.get_offset())
    g.set_offset(off + 4 * ((like.numel() + 3) // 4))
    return (int(g.initial_seed()) * 0x9E3779B97F4A7C15 + off * 0xD1B54A32D192ED03 + 1) & (2 ** 62 - 1)


def apply_ops(batch: torch.Tensor, ops: Sequence[int], shift: int = 0, freq: tuple = (0, 0), time: tuple = (0, 0),
              noise_level: float = 0.0, noise_seed: int = 0, normalize_ref_db: Optional[float] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The fused kernel: ``ops`` (distinct GTC_AUG_* codes, application order) then optional db_normalize."""
    _need_cuda(batch)
    if batch.dtype != torch.float32 or batch.dim() != 4:
        raise _lib.GtcError("augmentation kernels take float32 (B, C, dim2, dim3) tensors")
    b, c, h, w = batch.shape
    if out is None:
        out = torch.empty_like(batch)
    _need_cuda(out)
    arr = (C.c_int * max(1, len(ops)))(*ops)
    _lib.check(_lib.load().gtc_augment_batch(_ptr(batch), _ptr(out), b, c, h, w, C.cast(arr, C.c_void_p), len(ops), int(shift),
                                             int(freq[0]), int(freq[1]), int(time[0]), int(time[1]), float(noise_level),
                                             int(noise_seed), 0 if normalize_ref_db is None else 1,
                                             -120.0 if normalize_ref_db is None else float(normalize_ref_db), _stream()),
               "gtc_augment_batch")
    return out


# ---------------------------------------------------------------------------------------------------- the reference's API

def time_shift(audio, shift_range=0.1):
    """ViT_engine.py:28-42."""
    time_dim = audio.shape[2]
    if time_dim < 2:
        return audio
    shift = int(random.uniform(-shift_range, shift_range) * time_dim)
    if shift == 0:
        return audio
    return apply_ops(audio, [TIME_SHIFT], shift=shift)


def add_noise(audio, noise_level=0.005):
    """ViT_engine.py:44-47."""
    return apply_ops(audio, [NOISE], noise_level=noise_level, noise_seed=_noise_seed(audio))


def _draw_mask(dim, max_width):
    max_width = min(max_width, dim)
    if max_width < 1:
        return None, max_width
    w = random.randint(1, max_width)
    return (random.randint(0, dim - w), w), max_width


def frequency_mask(audio, num_masks=1, max_width=5):
    """ViT_engine.py:49-63 (in place, like the reference)."""
    freq_dim = audio.shape[3]
    if freq_dim < 2:
        return audio
    for _ in range(num_masks):
        m, max_width = _draw_mask(freq_dim, max_width)
        if m is not None:
            apply_ops(audio, [FREQ_MASK], freq=m, out=audio)
    return audio


def time_mask(audio, num_masks=1, max_width=10):
    """ViT_engine.py:65-79 (in place, like the reference)."""
    time_dim = audio.shape[2]
    if time_dim < 2:
        return audio
    for _ in range(num_masks):
        m, max_width = _draw_mask(time_dim, max_width)
        if m is not None:
            apply_ops(audio, [TIME_MASK], time=m, out=audio)
    return audio


def draw_augmentation(shape, augment_prob=0.5):
    """The random decisions of augment_batch (ViT_engine.py:81-93) for a batch of ``shape``, drawn from ``random`` in
    the reference's order.  Returns kwargs for ``apply_ops`` (``ops`` may be empty)."""
    plan = dict(ops=[], shift=0, freq=(0, 0), time=(0, 0), noise_level=0.0)
    if random.random() < augment_prob:
        names = ["time_shift", "add_noise", "frequency_mask", "time_mask"]
        num_augs = random.randint(1, 3)
        for name in random.sample(names, num_augs):
            if name == "time_shift":
                if shape[2] >= 2:
                    shift = int(random.uniform(-0.1, 0.1) * shape[2])
                    if shift != 0:
                        plan["ops"].append(TIME_SHIFT)
                        plan["shift"] = shift
            elif name == "add_noise":
                plan["ops"].append(NOISE)
                plan["noise_level"] = 0.005
            elif name == "frequency_mask":
                if shape[3] >= 2:
                    m, _ = _draw_mask(shape[3], 5)
                    if m is not None:
                        plan["ops"].append(FREQ_MASK)
                        plan["freq"] = m
            else:
                if shape[2] >= 2:
                    m, _ = _draw_mask(shape[2], 10)
                    if m is not None:
                        plan["ops"].append(TIME_MASK)
                        plan["time"] = m
    return plan


def augment_batch(batch, augment_prob=0.5, normalize_ref_db: Optional[float] = None):
    """ViT_engine.py:81-93, one fused pass; ``normalize_ref_db`` additionally fuses the db_normalize the engine applies
    next (ViT_engine.py:284-287).  Returns a NEW tensor: the reference's frequency_mask / time_mask write into the caller's
    batch in place and return it (ViT_engine.py:57-79); the returned values are the same, the side effect on ``batch`` is
    deliberately not reproduced."""
    plan = draw_augmentation(tuple(batch.shape), augment_prob)
    if not plan["ops"] and normalize_ref_db is None:
        return batch
    seed = _noise_seed(batch) if NOISE in plan["ops"] else 0
    return apply_ops(batch, noise_seed=seed, normalize_ref_db=normalize_ref_db, **plan)


def db_normalize(batch, ref_db=-120.0):
    """ViT_engine.py:112-117: map [ref_db, 0] dB to [0, 1] with clamping."""
    _need_cuda(batch)
    if batch.dtype != torch.float32:
        raise _lib.GtcError("db_normalize takes float32 tensors")
    out = torch.empty_like(batch)
    _lib.check(_lib.load().gtc_db_normalize(_ptr(batch), batch.numel(), float(ref_db), _ptr(out), _stream()), "gtc_db_normalize")
    return out
