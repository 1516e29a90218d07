"""Regenerates tests/golden/.  Run in the build container (needs /root/reference for the label fixtures):

    python tests/golden/make_golden.py

* ref_label_*.npy  -- three label files copied byte for byte from /root/reference/tablatures/ (data fixtures, not source):
                     they pin the on-disk label LAYOUT only (their content was produced by a script
                     that is not in the reference repo).
* oracle_cqt_seed0.npz -- the CPU oracle's own output on a seeded input.  The reference has no golden vectors and
                     librosa/soxr cannot be installed, so this fixture pins the ORACLE against silent drift
                     ("parity unpinned" with respect to real librosa, see oracle/__init__.py).
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import cqt_oracle as co          # noqa: E402
from oracle import labels_oracle as lo       # noqa: E402
from conftest import make_test_audio         # noqa: E402

REF = "/root/reference/tablatures"
if os.path.isdir(REF):
    for name in ("00_BN1-129-Eb_comp_segment_0_0.00.npy", "00_BN1-129-Eb_solo_segment_1_10.20.npy", "01_BN3-119-G_comp_segment_68_11.00.npy"):
        shutil.copyfile(os.path.join(REF, name), os.path.join(HERE, "ref_label_" + name))

y = make_test_audio(22050, seed=0)
feats, pre, cplx = [], [], []
for i in range(9):
    f, p, C = co.segment_features(y[i * 2205: i * 2205 + 4410], 22050, fmin=co.note_to_hz_C(1), return_pre_cut=True)
    feats.append(f); pre.append(p); cplx.append(C)
rng = np.random.default_rng(0)
onset, dur, pitch = rng.uniform(0, 5, 40), rng.uniform(0.05, 1.5, 40), rng.uniform(38, 84, 40)
times = np.asarray(lo.segment_times(5.0, 25))
np.savez_compressed(os.path.join(HERE, "oracle_cqt_seed0.npz"), audio=y, features=np.stack(feats), pre_cut=np.stack(pre),
                    cqt=np.stack(cplx), taps=co.halfband_taps(), onset=onset, dur=dur, pitch=pitch, times=times,
                    labels=lo.rasterize_events_numpy(onset, dur, pitch, times))
print("golden fixtures written to", HERE)
