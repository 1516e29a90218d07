"""GEMM + dB finish alone (plan.contract_db) on chunks of whole 30 s clips, and the pure-store bandwidth of the GPU
(cudaMemset through torch) as the upper bound of what the patch kernel can reach.  usage: gemm_bench.py [clips ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "guitar-tablature-classification_b200"))
import numpy as np, torch
from gtc_b200 import ops, CqtRecipe
dev = torch.device("cuda:0")
r = CqtRecipe(); n = int(r.sr) * 30
plan = ops.CqtPlan(r)
def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.min(ts)), float(np.median(ts))
for n_clips in [int(a) for a in sys.argv[1:]] or [54, 63]:
    audio = (torch.rand(n_clips * n, device=dev) - 0.5)
    clip_off, seg_off = plan.offsets([n] * n_clips)
    n_seg = int(seg_off[-1])
    co, so = torch.from_numpy(clip_off).to(dev), torch.from_numpy(seg_off).to(dev)
    ws = plan.workspace(n_seg, n_clips)
    db = torch.empty((n_seg, 96, 5), device=dev)
    plan.frame(audio, co, so, n_seg, ws)
    best, med = timeit(lambda: plan.contract_db(co, so, n_seg, db, ws))
    rows = n_seg + n_clips
    print(json.dumps({"kernel": "gemm+finish", "clips": n_clips, "rows": rows, "ms_best": round(best, 4), "ms_med": round(med, 4),
                      "wave_eff": round(plan.gemm_wave_efficiency(n_seg, n_clips), 3),
                      "fp16_issued_PFLOPs": round(3 * 2.0 * rows * 960 * 4416 / (best * 1e-3) / 1e15, 3)}), flush=True)
buf = torch.empty(10 * 2**30, dtype=torch.uint8, device=dev)
best, med = timeit(lambda: buf.zero_())
print(json.dumps({"kernel": "memset 10 GiB", "ms_best": round(best, 4), "GBs": round(buf.numel() / (best * 1e-3) / 1e9, 1)}))
f = buf.view(torch.float32)
best, med = timeit(lambda: f.fill_(1.5))
print(json.dumps({"kernel": "fill_ fp32 10 GiB", "ms_best": round(best, 4), "GBs": round(buf.numel() / (best * 1e-3) / 1e9, 1)}))
