"""Patch kernel bandwidth against the number of resident CTAs (GTC_OPT_PATCH_MAX_CTAS): how many CTAs per SM the HBM store
stream really needs -- the register budget a co-resident GEMM CTA would have to fit beside."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "guitar-tablature-classification_b200"))
import numpy as np, torch
from gtc_b200 import ops
dev = torch.device("cuda:0")
n = 8192
db = (torch.rand((n, 96, 5), device=dev) * 120 - 120)
out = torch.empty((n, 3, 224, 224), device=dev)
byts = n * (3 * 224 * 224 * 4 + 1920)
for ahead, ctas in [(a, c) for c in (148 * 2, 148, 148 * 3) for a in (0, 1, 0, 1)]:
    ops.set_option(16, ctas)
    ops.set_option(17, 4)
    try:
        ops.set_option(18, ahead)
    except Exception:
        if ahead: continue
    for _ in range(3): ops.patches(db, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.patches(db, out=out); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts))
    print(json.dumps({"ahead": ahead, "max_ctas": ctas, "ms": round(ms, 4), "GBs": round(byts / ms / 1e6, 1)}), flush=True)
ops.set_option(16, 0); ops.set_option(17, 0)
