"""Deterministic synthetic inputs shaped like the reference's data.

* ``pluck_clips``  -- mono fp32 audio: decaying harmonic plucks (6 per second, MIDI 40..82) + white noise at -40 dBFS,
                      peak-normalised to 0.5.  Written with torch ops so the 360-clip batch can be generated on the GPU
                      for the benchmark and on the CPU for tests (same function, the caller picks the device).
* ``note_events``  -- JAMS-like note events per clip: 6 strings x Poisson(3 notes/s), fractional MIDI pitches.
* ``contours``     -- pitch-contour observations for the fallback path of jam_to_tablature.py:145-178.
"""
from __future__ import annotations

import numpy as np
import torch

OPEN_STRINGS = np.array([40, 45, 50, 55, 59, 64], dtype=np.float64)


def pluck_clips(n_clips: int, n_samples: int, sr: float = 22050.0, seed: int = 0, device="cpu",
                plucks_per_s: float = 6.0, block: int = 8) -> torch.Tensor:
    """[n_clips, n_samples] fp32.  Each pluck: sum_h a_h sin(2 pi h f t) exp(-t/tau), 4 harmonics."""
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    dur = n_samples / sr
    n_pl = max(1, int(round(plucks_per_s * dur)))
    out = torch.empty((n_clips, n_samples), dtype=torch.float32, device=device)
    t = torch.arange(n_samples, dtype=torch.float32, device=device) / sr
    for c0 in range(0, n_clips, block):
        c1 = min(n_clips, c0 + block)
        b = c1 - c0
        midi = 40 + 42 * torch.rand((b, n_pl), generator=g)
        onset = dur * torch.rand((b, n_pl), generator=g)
        amp = 0.3 + 0.7 * torch.rand((b, n_pl), generator=g)
        tau = 0.15 + 0.6 * torch.rand((b, n_pl), generator=g)
        noise = torch.randn((b, n_samples), generator=g) * (10 ** (-40 / 20))
        f = (440.0 * 2 ** ((midi - 69) / 12)).to(device)
        onset, amp, tau = onset.to(device), amp.to(device), tau.to(device)
        y = noise.to(device)
        for p in range(n_pl):
            rel = t[None, :] - onset[:, p, None]
            env = torch.where(rel >= 0, torch.exp(-rel.clamp(min=0) / tau[:, p, None]), torch.zeros_like(rel))
            ph = 2 * np.pi * f[:, p, None] * rel
            s = torch.sin(ph) + 0.5 * torch.sin(2 * ph) + 0.25 * torch.sin(3 * ph) + 0.125 * torch.sin(4 * ph)
            y = y + amp[:, p, None] * env * s
        y = y * (0.5 / y.abs().amax(dim=1, keepdim=True).clamp(min=1e-12))
        out[c0:c1] = y
    return out


def note_events(durations, seed: int = 2, notes_per_s: float = 3.0):
    """Per clip: arrays (onset, dur, pitch) fp64, concatenated, plus evt_off [n_clips+1] (config 3 of SURVEY 8d)."""
    rng = np.random.default_rng(seed)
    on, du, pi, off = [], [], [], [0]
    for d in durations:
        o_c, d_c, p_c = [], [], []
        for s in range(6):
            n = rng.poisson(notes_per_s * d)
            o = rng.uniform(0, d, n)
            dd = np.clip(rng.exponential(0.4, n), 0.05, 4.0)
            p = OPEN_STRINGS[s] + rng.integers(0, 19, n) + rng.normal(0, 0.15, n)
            o_c.append(o); d_c.append(dd); p_c.append(p)
        o_c, d_c, p_c = np.concatenate(o_c), np.concatenate(d_c), np.concatenate(p_c)
        order = np.argsort(o_c, kind="stable")
        on.append(o_c[order]); du.append(d_c[order]); pi.append(p_c[order])
        off.append(off[-1] + len(order))
    cat = lambda xs: np.concatenate(xs).astype(np.float64) if xs else np.zeros(0)
    return cat(on), cat(du), cat(pi), np.asarray(off, dtype=np.int64)


def contours(durations, seed: int = 3, obs_per_s: float = 100.0):
    """Per clip: (time, frequency_hz, confidence) fp64 + con_off; ~30 % unvoiced (frequency 0)."""
    rng = np.random.default_rng(seed)
    tt, ff, cc, off = [], [], [], [0]
    for d in durations:
        n = int(d * obs_per_s)
        t = np.arange(n) / obs_per_s
        midi = 40 + 40 * rng.random(n)
        f = 440.0 * 2 ** ((midi - 69) / 12)
        f[rng.random(n) < 0.3] = 0.0
        c = rng.random(n)
        tt.append(t); ff.append(f); cc.append(c)
        off.append(off[-1] + n)
    return np.concatenate(tt), np.concatenate(ff), np.concatenate(cc), np.asarray(off, dtype=np.int64)
