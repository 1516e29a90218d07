"""Per-kernel durations of one structured CQT call in a normal (not serialised, warm) run, via torch.profiler (CUPTI).
usage: prof_structured_kernels.py [--recipe inference|cqt]   -> one line per launch: name, grid, us"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "guitar-tablature-classification_b200"))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from gtc_b200 import ops, synth, CqtRecipe
from gtc_b200.inference import TabCnnFrontEnd
ap = argparse.ArgumentParser(); ap.add_argument("--recipe", default="inference")
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
else:
    r = CqtRecipe(); plan = ops.StructuredCqtPlan(r)
    n_clips, n = 55, 22050 * 30
    y = synth.pluck_clips(8, n, sr=22050, seed=1, device=dev).repeat(7, 1)[:n_clips].contiguous().reshape(-1)
    per = (n - r.seg_len) // r.seg_hop + 1
    starts = (np.arange(n_clips)[:, None] * n + np.arange(per)[None, :] * r.seg_hop).reshape(-1)
    seg_len = r.seg_len; valid = np.full(len(starts), seg_len, np.int32)
st, va = torch.from_numpy(starts).to(dev), torch.from_numpy(valid).to(dev)
le = torch.full((len(starts),), seg_len, dtype=torch.int32, device=dev)
out = plan.segments_db(y, st, va, le, seg_len)
for _ in range(3):
    plan.segments_db(y, st, va, le, seg_len, out=out)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    plan.segments_db(y, st, va, le, seg_len, out=out)
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
tot = 0.0
for e in evs:
    d = e.time_range.end - e.time_range.start
    tot += d
    print(f"{e.time_range.start - t0:9.1f} us  +{d:7.1f} us  {e.name[:90]}")
print(json.dumps({"recipe": a.recipe, "kernels": len(evs), "sum_us": tot, "span_us": evs[-1].time_range.end - t0}))
