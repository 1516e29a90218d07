"""Event timeline of one FrontEnd.run() step (no nsys in this image): every kernel group is bracketed by CUDA events on
its own stream and printed relative to the first mark.  usage: python scripts/timeline.py [--host] [--steps 2]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "guitar-tablature-classification_b200"))
import numpy as np, torch
from gtc_b200 import CqtRecipe, synth
from gtc_b200.pipeline import FrontEnd, ShardInputs
ap = argparse.ArgumentParser()
ap.add_argument("--host", action="store_true"); ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--clips", type=int, default=360); ap.add_argument("--chunk-segments", type=int, default=28400)
a = ap.parse_args()
dev = torch.device("cuda:0"); SR = 22050; n = SR * 30
audio = synth.pluck_clips(a.clips, n, sr=SR, seed=1, device=dev, block=24).reshape(-1)
on, du, pi, evt_off = synth.note_events([30.0] * a.clips, seed=2)
ev_host = torch.from_numpy(np.stack([on, du, pi])).pin_memory()
pcm = torch.clamp(torch.round(audio * 32768.0), -32768, 32767).to(torch.int16)
lens = np.full(a.clips, n, dtype=np.int64)
fe = FrontEnd(CqtRecipe(), device=0, chunk_segments=a.chunk_segments, patch_batch=a.chunk_segments)
if a.host:
    h = torch.empty(pcm.shape, dtype=torch.int16, pin_memory=True); h.copy_(pcm)
    inp = ShardInputs(h, lens, ev_host, evt_off, sr=SR)
else:
    inp = ShardInputs(pcm.float() / 32768.0, lens, ev_host.to(dev), evt_off, sr=SR)
for _ in range(3):
    fe.run(inp, device_inputs=not a.host)
torch.cuda.synchronize()
fe.trace = []
import time
base = torch.cuda.Event(enable_timing=True); base.record()
h0 = time.perf_counter()
for i in range(a.steps):
    fe.run(inp, device_inputs=not a.host)
    print(f"host: run() {i} returned at {1e3 * (time.perf_counter() - h0):.3f} ms")
end = torch.cuda.Event(enable_timing=True); end.record()
torch.cuda.synchronize()
rows = sorted(((base.elapsed_time(e), lab, st, 1e3 * (h - h0)) for lab, st, e, h in fe.trace))
print("total ms", base.elapsed_time(end))
open_at = {}
for t, lab, st, h in rows:      # device time of the mark, host time at which it was enqueued
    if lab.endswith("<"):
        open_at[lab[:-1]] = t
        print(f"{t:9.3f}  (host {h:8.3f})  {st:6s} {lab}")
    else:
        k = lab[:-1]; t0 = open_at.get(k)
        print(f"{t:9.3f}  (host {h:8.3f})  {st:6s} {lab}" + (f"   dur {t - t0:.3f}" if t0 is not None else ""))
