"""Host-facing helpers shared by the drop-in modules: cached device plans and batched feature extraction."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from . import ops
from .cqt_design import CqtRecipe

_PLANS: Dict[tuple, ops.CqtPlan] = {}


def get_plan(recipe: CqtRecipe, seg_len: int | None = None, seg_hop: int | None = None) -> ops.CqtPlan:
    """One device-resident plan per (recipe, segment geometry, device); the operator design is done once."""
    key = (recipe, seg_len, seg_hop, torch.cuda.current_device())
    plan = _PLANS.get(key)
    if plan is None:
        plan = ops.CqtPlan(recipe, seg_len=seg_len, seg_hop=seg_hop)
        _PLANS[key] = plan
    return plan


def clips_features(clips: Sequence[np.ndarray], recipe: CqtRecipe, seg_len: int | None = None, seg_hop: int | None = None,
                   max_segments_per_call: int = 32768) -> List[np.ndarray]:
    """dB features of every complete window of every clip (the hot loop of cqt.py:36-58), batched on the GPU.
    Returns one float32 array [n_seg_c, n_bins, T] per clip (possibly empty)."""
    plan = get_plan(recipe, seg_len, seg_hop)
    dev = torch.device("cuda", torch.cuda.current_device())
    lens = np.asarray([len(c) for c in clips], dtype=np.int64)
    counts = ops.segment_counts(lens, plan.seg_len, plan.seg_hop)
    out: List[np.ndarray] = [None] * len(clips)
    c0 = 0
    while c0 < len(clips):
        c1, segs = c0, 0
        while c1 < len(clips) and (c1 == c0 or segs + counts[c1] <= max_segments_per_call):
            segs += int(counts[c1])
            c1 += 1
        batch = [np.ascontiguousarray(c, dtype=np.float32) for c in clips[c0:c1]]
        clip_off, seg_off = plan.offsets([len(b) for b in batch])
        n_seg = int(seg_off[-1])
        if n_seg:
            flat = torch.from_numpy(np.concatenate(batch) if len(batch) > 1 else batch[0]).to(dev)
            db = plan.segments_db(flat, torch.from_numpy(clip_off).to(dev), torch.from_numpy(seg_off).to(dev), n_seg).cpu().numpy()
        else:
            db = np.zeros((0, plan.n_bins, plan.n_frames), dtype=np.float32)
        for i in range(c1 - c0):
            out[c0 + i] = db[seg_off[i]: seg_off[i + 1]]
        c0 = c1
    return out
