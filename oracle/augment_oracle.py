"""CPU oracle: the batch augmentations and normalisation of /root/reference/ViT_engine.py:28-117, restated with plain
torch CPU ops and EXPLICIT parameters.  TEST INFRASTRUCTURE ONLY.

Pinned: tests/golden/ref_augment.npz holds outputs of the reference's own functions (extracted from ViT_engine.py and
executed in the build container by tests/golden/make_reference_golden_aug.py) together with the decisions they drew;
tests/test_oracle_augment.py replays them through this file.
"""
from __future__ import annotations

import torch

TIME_SHIFT, NOISE, FREQ_MASK, TIME_MASK = 1, 2, 3, 4


def time_shift(audio: torch.Tensor, shift: int) -> torch.Tensor:
    """ViT_engine.py:35-41 with the drawn ``shift``."""
    if shift > 0:
        return torch.cat([audio[:, :, shift:, :], torch.zeros_like(audio[:, :, :shift, :])], dim=2)
    if shift < 0:
        s = -shift
        return torch.cat([torch.zeros_like(audio[:, :, :s, :]), audio[:, :, :-s, :]], dim=2)
    return audio


def add_noise(audio: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """ViT_engine.py:46-47 with the noise tensor (already scaled) given."""
    return audio + noise


def frequency_mask(audio: torch.Tensor, f0: int, f: int) -> torch.Tensor:
    """ViT_engine.py:62."""
    audio = audio.clone()
    audio[:, :, :, f0:f0 + f] = 0
    return audio


def time_mask(audio: torch.Tensor, t0: int, t: int) -> torch.Tensor:
    """ViT_engine.py:78."""
    audio = audio.clone()
    audio[:, :, t0:t0 + t, :] = 0
    return audio


def db_normalize(batch: torch.Tensor, ref_db: float = -120.0) -> torch.Tensor:
    """ViT_engine.py:112-117."""
    return torch.clamp((batch - ref_db) / (-ref_db), 0, 1)


def apply_ops(batch, ops, shift=0, freq=(0, 0), time=(0, 0), noise=None, normalize_ref_db=None):
    """The composition augment_batch applies (ViT_engine.py:90-91), ops in order."""
    x = batch
    for op in ops:
        if op == TIME_SHIFT:
            x = time_shift(x, shift)
        elif op == NOISE:
            x = add_noise(x, noise)
        elif op == FREQ_MASK:
            x = frequency_mask(x, *freq)
        elif op == TIME_MASK:
            x = time_mask(x, *time)
    if normalize_ref_db is not None:
        x = db_normalize(x, normalize_ref_db)
    return x
