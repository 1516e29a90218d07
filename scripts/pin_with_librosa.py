#!/usr/bin/env python
"""Pin the CQT path to the real librosa / soxr wherever they are importable (they are NOT in the build image).

    python scripts/pin_with_librosa.py [--segments 64] [--write-golden] [--operator-out op.npy]

What the reference computes per segment (/root/reference/cqt.py:55-58):

    C   = librosa.cqt(segment, sr=sr, hop_length=1024, n_bins=96, bins_per_octave=12, fmin=librosa.note_to_hz('C1'))
    out = cqt_lim(librosa.amplitude_to_db(np.abs(C)**4, ref=np.amax))

This kit, with `import librosa, soxr` working:
  1. measures the true 2:1 tap table as the impulse response of ``soxr.resample(x, 2, 1, 'HQ')`` and diffs it against the
     restated libsoxr design shared by the product (gtc_b200/cqt_design.py:decimator_taps) and the oracle;
  2. runs the reference recipe above on the seed-0 clip of SURVEY.md 8d config 1 and diffs oracle/cqt_oracle.py against it
     (worst dB delta above the -60 dB cut, worst relative magnitude error; gates 0.01 dB / 1e-4);
  3. builds the segment operator by pushing unit impulses through the real ``librosa.cqt`` (the call is linear in the
     samples) and diffs it against ``cqt_design.build_operator``; ``--operator-out`` saves it, and
     ``GTC_OPERATOR_FILE=<that file>`` (or ``ops.CqtPlan(operator=...)`` -> ``gtc_cqt_plan_create(h_operator)``) makes libgtc
     evaluate the MEASURED operator, i.e. exact soxr behaviour without restating it;
  4. on a CUDA machine, evaluates that operator on the GPU and diffs the dB features against librosa's directly;
  5. ``--write-golden`` stores segments + librosa outputs + soxr outputs in tests/golden/librosa_pin.npz, after which
     tests/test_librosa_pin.py pins the oracle (and the GPU path) to them on machines WITHOUT librosa.

Without librosa/soxr it prints {"status": "unpinned", ...} and exits 3.  ``backend=`` lets the tests drive the same code with
a stand-in built from the oracle, so the kit itself is exercised in the build image.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "guitar-tablature-classification_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

GOLDEN = os.path.join(ROOT, "tests", "golden", "librosa_pin.npz")
SR = 22050
DB_GATE, REL_GATE = 0.01, 1e-4


def real_backend():
    """(librosa, soxr) wrapped in the four calls the kit needs, or None."""
    try:
        import librosa
        import soxr
    except Exception:
        return None
    return SimpleNamespace(
        name=f"librosa {librosa.__version__} / soxr {soxr.__version__} (libsoxr {getattr(soxr, '__libsoxr_version__', '?')})",
        cqt=lambda y, sr: librosa.cqt(y, sr=sr, hop_length=1024, n_bins=96, bins_per_octave=12, fmin=librosa.note_to_hz("C1")),
        amplitude_to_db=lambda m: librosa.amplitude_to_db(m, ref=np.amax),
        resample2=lambda x: soxr.resample(x, 2, 1, "HQ"),
        librosa_resample2=lambda x: librosa.resample(x, orig_sr=2, target_sr=1, res_type="soxr_hq", scale=True))


def oracle_backend():
    """Stand-in with the same four calls, made of the oracle: drives the kit's own code in images without librosa."""
    from oracle import cqt_oracle as o

    def resample2(x):
        x = np.asarray(x)
        return (o.resample_2to1(x) * np.sqrt(0.5)).astype(x.dtype)
    return SimpleNamespace(name="oracle stand-in (NOT librosa)",
                           cqt=lambda y, sr: o.cqt(np.asarray(y, dtype=np.float32), sr=sr, fmin=o.note_to_hz_C(1)),
                           amplitude_to_db=lambda m: o.amplitude_to_db_amax(m),
                           resample2=resample2,
                           librosa_resample2=lambda x: o.resample_2to1(np.asarray(x, dtype=np.float32)))


def cqt_lim(x):
    y = np.copy(x)
    y[y < -60] = -120
    return y


def seed0_segments(n_segments: int, sr: int = SR):
    from gtc_b200 import synth
    y = synth.pluck_clips(1, sr * 30, sr=sr, seed=0)[0].numpy()
    w, h = int(0.2 * sr), int(0.1 * sr)
    n = (len(y) - w) // h + 1
    pick = np.unique(np.linspace(0, n - 1, min(n, n_segments)).astype(int))
    return np.stack([y[i * h: i * h + w] for i in pick]).astype(np.float32), pick


def measure_taps(be, n: int = 2048):
    """Impulse response of the backend's 2:1 resampler: an impulse at an even / odd input position gives the even / odd
    phase of h (out[k] = sum_i h[i] x[2k + c - i]); returned centred, trimmed to the span above 1e-13.  float64 input
    selects libsoxr's double-precision engine (same design, no fp32 noise floor above the 3e-8 edge taps)."""
    outs = []
    for p in (n // 2, n // 2 + 1):
        x = np.zeros(n, dtype=np.float64)
        x[p] = 1.0
        outs.append(np.asarray(be.resample2(x), dtype=np.float64))
    # out_even[k] = h[2k + c - p0], out_odd[k] = h[2k + c - p0 - 1]  ->  interleave on the grid j = 2k - p0
    h = np.zeros(2 * len(outs[0]) + 2)
    h[1::2][: len(outs[0])] = outs[0]         # index 2k+1  <->  j = 2k - p0 (+1 shift keeps everything non-negative)
    h[0::2][: len(outs[1])] = outs[1]         # index 2k    <->  j = 2k - p0 - 1
    nz = np.nonzero(np.abs(h) > 1e-13)[0]
    h = h[nz[0]: nz[-1] + 1]
    return h


def features_of(be, segs, sr):
    feats, mags = [], []
    for s in segs:
        C = be.cqt(s, sr)
        m = np.abs(C)
        db = be.amplitude_to_db(m ** 4)
        feats.append(np.asarray(db, dtype=np.float32))
        mags.append(np.asarray(m, dtype=np.float32))
    return np.stack(feats), np.stack(mags)


def diff_features(db_ref, mag_ref, db_got, mag_got, margin=0.02):
    keep = db_ref > -60.0 + margin
    return {"max_db_delta_above_cut": float(np.abs(db_got - db_ref)[keep].max()),
            "max_rel_magnitude_error_above_cut": float((np.abs(mag_got - mag_ref) / np.maximum(mag_ref, 1e-30))[keep].max()),
            "cut_flips": int(((db_ref >= -60.0) != (db_got >= -60.0)).sum()), "elements_compared": int(keep.sum())}


def measured_operator(be, seg_len: int, sr: int, block: int = 490, columns=None):
    """(2 * n_bins * T, len(columns)) float32, row = (t * n_bins + bin) * 2 + {re, im}: column j = cqt(unit impulse at j);
    ``columns`` defaults to all seg_len sample positions."""
    cols = []
    columns = np.arange(seg_len) if columns is None else np.asarray(columns)
    for j0 in range(0, len(columns), block):
        js = columns[j0: j0 + block]
        eye = np.zeros((len(js), seg_len), dtype=np.float32)
        eye[np.arange(len(js)), js] = 1.0
        C = np.asarray(be.cqt(eye, sr))                               # multichannel call: (b, n_bins, T) complex64
        b, nb, T = C.shape
        cols.append(np.stack([C.real, C.imag], axis=-1).transpose(0, 2, 1, 3).reshape(b, T * nb * 2))
    return np.ascontiguousarray(np.concatenate(cols, axis=0).T.astype(np.float32))


def run(be=None, n_segments: int = 64, with_operator: bool = True, operator_out: str | None = None,
        write_golden: bool = False, golden_path: str = GOLDEN, use_gpu: bool | None = None) -> dict:
    be = real_backend() if be is None else be
    if be is None:
        return {"status": "unpinned", "why": "librosa / soxr are not importable here; CQT parity stays restated-only "
                                             "(profiles/r02_tap_sensitivity.md bounds the risk)"}
    from oracle import cqt_oracle as o
    from gtc_b200 import cqt_design
    rep = {"status": "pinned", "backend": be.name, "gates": {"db": DB_GATE, "rel_magnitude": REL_GATE}}

    # 1. tap table
    h_true = measure_taps(be)
    h_rest = cqt_design.decimator_taps()
    rep["taps"] = {"measured_len": int(len(h_true)), "restated_len": int(len(h_rest))}
    if len(h_true) % 2 == 1:
        n = max(len(h_true), len(h_rest))
        pad = lambda h: np.pad(h, ((n - len(h)) // 2, (n - len(h)) // 2))
        rep["taps"]["max_abs_diff"] = float(np.abs(pad(h_true) - pad(h_rest)).max())
        rep["taps"]["dc_gain_measured"] = float(h_true.sum())
    # 2. oracle vs reference recipe
    segs, pick = seed0_segments(n_segments)
    db_ref, mag_ref = features_of(be, segs, SR)
    cache = {}
    db_o, mag_o = [], []
    for s in segs:
        _, pre, C = o.segment_features(s, SR, fmin=o.note_to_hz_C(1), _basis_cache=cache, return_pre_cut=True)
        db_o.append(pre)
        mag_o.append(np.abs(C))
    rep["oracle_vs_reference"] = diff_features(db_ref, mag_ref, np.stack(db_o), np.stack(mag_o))
    if len(h_true) % 2 == 1:
        # the oracle evaluated with the MEASURED table: separates "wrong taps" from "wrong anything else"
        with o.taps_override(h_true):
            db_t, mag_t = [], []
            for s in segs:
                _, pre, C = o.segment_features(s, SR, fmin=o.note_to_hz_C(1), _basis_cache=cache, return_pre_cut=True)
                db_t.append(pre)
                mag_t.append(np.abs(C))
        rep["oracle_with_measured_taps_vs_reference"] = diff_features(db_ref, mag_ref, np.stack(db_t), np.stack(mag_t))
    # soxr on fixed segments (the first 4410 / 8820 samples of the clip), what ADVICE asks to keep as golden
    x1, x2 = segs[0], np.concatenate([segs[0], segs[min(2, len(segs) - 1)]])
    sox = {f"soxr_{len(x)}": np.asarray(be.librosa_resample2(x), dtype=np.float32) for x in (x1, x2)}
    rep["resample_vs_reference"] = {k: float(np.abs(o.resample_2to1(x) - v).max()) for (k, v), x in zip(sox.items(), (x1, x2))}
    # 3. operator
    A_ref = None
    if with_operator:
        recipe = cqt_design.CqtRecipe()
        A_ref = measured_operator(be, recipe.seg_len, SR)
        A = cqt_design.build_operator(recipe)
        rep["operator"] = {"shape": list(A_ref.shape), "max_abs_diff": float(np.abs(A - A_ref).max()),
                           "rel_fro_diff": float(np.linalg.norm(A - A_ref) / np.linalg.norm(A_ref))}
        mag_lin = np.abs((segs.astype(np.float64) @ A_ref.T.astype(np.float64)).reshape(len(segs), -1, 96, 2) @ np.array([1, 1j]))
        mag_lin = mag_lin.transpose(0, 2, 1)                            # (seg, bin, t)
        rep["operator"]["linearity_rel_err"] = float(np.abs(mag_lin - mag_ref).max() / mag_ref.max())
        if operator_out:
            np.save(operator_out, A_ref)
            rep["operator"]["saved"] = operator_out
    # 4. GPU evaluation of the measured operator
    try:
        import torch
        gpu = torch.cuda.is_available() if use_gpu is None else use_gpu
    except Exception:
        gpu = False
    if gpu and A_ref is not None:
        import torch
        from gtc_b200 import ops
        plan = ops.CqtPlan(cqt_design.CqtRecipe(), operator=A_ref)
        flat = torch.from_numpy(segs.reshape(-1)).cuda()
        lens = [segs.shape[1]] * len(segs)
        clip_off, seg_off = plan.offsets(lens)
        got = plan.segments_db(flat, torch.from_numpy(clip_off).cuda(), torch.from_numpy(seg_off).cuda(), len(segs)).cpu().numpy()
        want = np.stack([cqt_lim(d) for d in db_ref])
        safe = np.abs(db_ref + 60.0) > 0.02
        rep["gpu_measured_operator_vs_reference"] = {"max_db_delta": float(np.abs(got - want)[safe].max())}
    ok = rep["oracle_vs_reference"]["max_db_delta_above_cut"] <= DB_GATE and \
        rep["oracle_vs_reference"]["max_rel_magnitude_error_above_cut"] <= REL_GATE
    rep["verdict"] = "restated design matches the reference within the gates" if ok else \
        "restated design is OUTSIDE the gates: use the measured operator (GTC_OPERATOR_FILE) and fix decimator_taps()"
    # 5. golden
    if write_golden:
        np.savez_compressed(golden_path, backend=np.array(be.name), sr=np.array(SR), segment_index=pick, segments=segs,
                            db_pre_cut=db_ref, magnitude=mag_ref, taps=h_true, **sox)
        rep["golden"] = golden_path
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--segments", type=int, default=64)
    ap.add_argument("--no-operator", action="store_true")
    ap.add_argument("--operator-out", default=None)
    ap.add_argument("--write-golden", action="store_true")
    ap.add_argument("--stand-in", action="store_true", help="drive the kit with the oracle stand-in (self-test, pins nothing)")
    a = ap.parse_args()
    rep = run(oracle_backend() if a.stand_in else None, a.segments, not a.no_operator, a.operator_out,
              a.write_golden and not a.stand_in)
    print(json.dumps(rep, indent=1))
    return 3 if rep["status"] == "unpinned" else 0


if __name__ == "__main__":
    sys.exit(main())
