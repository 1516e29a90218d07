"""The structured (multirate) CQT alone, for ncu: inference recipe (3 s segments, 50 % overlap) of 64 songs x 60 s, or with
--recipe cqt the cqt.py recipe (0.2 s windows) through the same path.  usage: prof_structured.py [--recipe inference|cqt] [--reps 3]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "guitar-tablature-classification_b200"))
import numpy as np, torch
from gtc_b200 import ops, synth, CqtRecipe
from gtc_b200.inference import TabCnnFrontEnd
ap = argparse.ArgumentParser(); ap.add_argument("--recipe", default="inference"); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
if a.recipe == "inference":
    plan = TabCnnFrontEnd().plan
    songs, L = 64, 22050 * 60
    y = synth.pluck_clips(8, L, sr=22050, seed=2, device=dev).repeat(8, 1).contiguous().reshape(-1)
    seg_len, hop = 66150, 33075
    s1 = np.arange(0, L, hop)
    starts = (np.arange(songs)[:, None] * L + s1[None, :]).reshape(-1)
    valid = np.tile(np.minimum(seg_len, L - s1), songs).astype(np.int32)
    secs = songs * 60.0
else:
    r = CqtRecipe(); plan = ops.StructuredCqtPlan(r)
    n_clips, n = 55, 22050 * 30
    y = synth.pluck_clips(8, n, sr=22050, seed=1, device=dev).repeat(7, 1)[:n_clips].contiguous().reshape(-1)
    per = (n - r.seg_len) // r.seg_hop + 1
    starts = (np.arange(n_clips)[:, None] * n + np.arange(per)[None, :] * r.seg_hop).reshape(-1)
    seg_len = r.seg_len; valid = np.full(len(starts), seg_len, np.int32); secs = len(starts) * 0.1
st, va = torch.from_numpy(starts).to(dev), torch.from_numpy(valid).to(dev)
le = torch.full((len(starts),), seg_len, dtype=torch.int32, device=dev)
out = plan.segments_db(y, st, va, le, seg_len)
torch.cuda.synchronize()
ts = []
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.segments_db(y, st, va, le, seg_len, out=out); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(json.dumps({"recipe": a.recipe, "n_seg": len(starts), "ms_best": min(ts), "s_audio_per_s": secs / (min(ts) * 1e-3)}))
