#!/usr/bin/env python
"""How far can a wrong soxr-HQ tap table move the dB features?  (CPU only; writes profiles/r02_tap_sensitivity.md)

The CQT arithmetic of the reference (/root/reference/cqt.py:55 -> librosa.cqt -> soxr_hq 2:1 decimation) is restated, not
diffed against libsoxr (parity unpinned).  This script bounds that risk: it perturbs every
free parameter of the restated Kaiser design -- beta, tap count, cut-off, the `rho` of the window argument -- plus two
independent textbook designs, pushes the seed-0 clip of SURVEY.md 8d config 1 through the oracle with each table, and
reports the largest change of the dB features above the -60 dB cut and of the magnitudes.

    python scripts/tap_sensitivity.py [--seconds 30] [--out profiles/r02_tap_sensitivity.md]
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import time

import numpy as np
import scipy.signal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "guitar-tablature-classification_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import cqt_oracle  # noqa: E402


def seed0_clip(seconds: float, sr: int = 22050) -> np.ndarray:
    from gtc_b200 import synth
    return synth.pluck_clips(1, int(sr * seconds), sr=sr, seed=0)[0].numpy()


def features(y: np.ndarray, sr: int, taps) -> tuple:
    """(pre-cut dB [n_seg, 96, 5], |C| [n_seg, 96, 5]) of every cqt.py window of the clip, all segments in one batch."""
    w, h = cqt_oracle.window_params(sr)
    n = cqt_oracle.num_segments(len(y), w, h)
    segs = np.stack([y[i * h: i * h + w] for i in range(n)]).astype(np.float32)
    with cqt_oracle.taps_override(taps):
        C = cqt_oracle.cqt(segs, sr=sr, fmin=cqt_oracle.note_to_hz_C(1), _basis_cache={})
    mag = np.abs(C)
    db = np.stack([cqt_oracle.amplitude_to_db_amax(m ** 4) for m in mag])
    return db, mag


def compare(base, other, margin=0.02):
    db0, m0 = base
    db1, m1 = other
    keep = db0 > -60.0 + margin                       # elements the reference keeps (cqt_lim), away from the discontinuity
    d_db = float(np.abs(db1 - db0)[keep].max())
    rel = float((np.abs(m1 - m0) / np.maximum(m0, 1e-30))[keep].max())
    flips = int(((db0 >= -60.0) != (db1 >= -60.0)).sum())
    return d_db, rel, flips, int(keep.sum())


def variants():
    t = cqt_oracle.soxr_hq_halfband_taps
    yield "restated libsoxr design (389 taps, beta 13.04) -- baseline", t()
    for s in (0.98, 1.02):
        yield f"beta x {s:.2f}", t(beta_scale=s)
    for d in (-8, 8):
        yield f"tap count {d:+d}", t(taps_delta=d)
    for s in (0.995, 1.005):
        yield f"cut-off Fc x {s:.3f}", t(fc_scale=s)
    yield "window argument without the rho term (rho = 0)", t(rho=0.0)
    yield "window argument rho = 1", t(rho=1.0)
    # DC-normalised copy (libsoxr's lsx_design_lpf can be called with or without normalisation; the sum is 1 + 4.6e-8)
    h = t()
    yield "taps normalised to unit DC gain", h / h.sum()
    # float32 taps (python-soxr runs float32 input through the single-precision engine)
    yield "taps rounded to float32", h.astype(np.float32).astype(np.float64)
    # two independent designs for scale: SciPy's Kaiser estimate for the same band edges, and a remez half-band
    # (frequencies in units of the INPUT Nyquist: pass-band end 0.45682, stop-band 0.5, cut-off 0.47841; firwin is DC-normalised)
    n_k, beta_k = scipy.signal.kaiserord(120.41, 0.5 - 0.45682)
    n_k += 1 - n_k % 2
    yield f"SciPy kaiserord stand-in ({n_k} taps, beta {beta_k:.2f})", scipy.signal.firwin(n_k, 0.47841, window=("kaiser", beta_k), fs=2.0)
    yield "SciPy firwin, 397 taps, beta 13.4", scipy.signal.firwin(397, 0.47841, window=("kaiser", 13.4), fs=2.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_tap_sensitivity.md"))
    args = ap.parse_args()
    sr = 22050
    y = seed0_clip(args.seconds, sr)
    rows, base = [], None
    t0 = time.time()
    for name, h in variants():
        f = features(y, sr, h)
        if base is None:
            base = f
            rows.append((name, len(h), 0.0, 0.0, 0, int((f[0] > -59.98).sum())))
            continue
        d_db, rel, flips, kept = compare(base, f)
        rows.append((name, len(h), d_db, rel, flips, kept))
        print(f"{name:60s} taps {len(h):4d}  max |d dB| {d_db:.4f}  max rel |C| {rel:.2e}  cut flips {flips}", flush=True)
    n_seg = base[0].shape[0]
    lines = [
        "# Sensitivity of the dB features to the soxr-HQ 2:1 tap table (round 2, CPU, `scripts/tap_sensitivity.py`)",
        "",
        f"Input: seed-0 clip of SURVEY.md 8d config 1 ({args.seconds:.0f} s @ 22 050 Hz, `gtc_b200.synth.pluck_clips(seed=0)`), all {n_seg} windows of",
        "`cqt.py:26-49` (4410 samples, hop 2205), evaluated by `oracle/cqt_oracle.py` with each tap table in turn.",
        "`max |d dB|` and `max rel |C|` are taken over the elements the reference keeps (pre-cut dB > -59.98, i.e. above `cqt_lim`'s",
        f"-60 dB cut, `cqt.py:10-13`): {rows[0][5]} of {n_seg * 480} elements.  `cut flips` = elements that change side of the cut.",
        "Gates (north_star): 0.01 dB and 1e-4 relative magnitude.",
        "",
        "| tap table | taps | max &#124;d dB&#124; | max rel &#124;C&#124; | cut flips |",
        "|---|---|---|---|---|",
    ]
    for name, n, d_db, rel, flips, _ in rows:
        lines.append(f"| {name} | {n} | {d_db:.4f} | {rel:.2e} | {flips} |")
    worst_shape = max(r[2] for r in rows[1:5] + rows[7:11])
    worst_fc = max(r[2] for r in rows[5:7])
    lines += [
        "",
        f"Reading.  The window SHAPE parameters (beta +-2 %, +-8 taps, rho 0..1, DC normalisation, fp32 taps) move the kept dB features by at most",
        f"**{worst_shape:.4f} dB** -- the same order as the 0.01 dB gate, as SURVEY.md section 7 measured for close Kaiser designs.  The CUT-OFF is the",
        f"sensitive parameter: Fc +-0.5 % moves them by **{worst_fc:.2f} dB** (it sets how much of the 0.914-1.0 x Nyquist transition band aliases",
        "into the next octave's sparsified filters, and every frame of an isolated 0.2 s segment is an edge frame).  `Fc`, beta and the tap count are",
        "not free recollections: they follow from libsoxr's published `passband_end = 1 - .05 / TO_3dB(rej)`, `stopband_begin = 1`,",
        "`lsx_kaiser_beta` table and `lsx_kaiser_params` polynomial (oracle/cqt_oracle.py:59-91 restates them line by line).  But the table",
        "says plainly what an error there would cost: the 0.01 dB gate holds against librosa only if the tap table is libsoxr's to ~0.1 % in",
        "Fc and ~1 % in beta.  The GPU path and the oracle share ONE table (so GPU-vs-oracle parity is exact to fp32 rounding) and the risk",
        "is confined to that table; `scripts/pin_with_librosa.py` measures the real table (impulse response of `soxr.resample`) and the",
        "real operator wherever `librosa` + `soxr` import, and `tests/test_librosa_pin.py` runs it automatically there.",
        "",
        f"(generated in {time.time() - t0:.0f} s on the build container's CPU)",
    ]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", args.out)


if __name__ == "__main__":
    main()
