"""The structured inference recipe (tablature_generator.py:599-666) on 1 / 4 / 16 / 64 songs of 60 s, one call each: what a single
song costs, and what programmatic dependent launch (GTC_SCQT_NO_PDL=1 turns it off) is worth at each size."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "guitar-tablature-classification_b200"))
import numpy as np, torch
from gtc_b200 import synth
from gtc_b200.inference import TabCnnFrontEnd
dev = torch.device("cuda:0")
plan = TabCnnFrontEnd().plan
res = {}
for songs in (1, 4, 16, 64):
    L = 22050 * 60
    y = synth.pluck_clips(songs, L, sr=22050, seed=2, device=dev).contiguous().reshape(-1)
    seg_len, hop = 66150, 33075
    s1 = np.arange(0, L, hop)
    starts = (np.arange(songs)[:, None] * L + s1[None, :]).reshape(-1)
    valid = np.tile(np.minimum(seg_len, L - s1), songs).astype(np.int32)
    st, va = torch.from_numpy(starts).to(dev), torch.from_numpy(valid).to(dev)
    le = torch.full((len(starts),), seg_len, dtype=torch.int32, device=dev)
    out = plan.segments_db(y, st, va, le, seg_len)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.segments_db(y, st, va, le, seg_len, out=out); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res[songs] = {"segments": len(starts), "ms_best": round(min(ts), 4), "s_audio_per_s": round(songs * 60.0 / (min(ts) * 1e-3))}
print(json.dumps(res))
