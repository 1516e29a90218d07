"""Patch kernel alone at bench.py's launch size, for A/B builds of libgtc.so (GTC_LIB_PATH): python scripts/patch_ab.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "guitar-tablature-classification_b200"))
import numpy as np, torch
from gtc_b200 import ops
dev = torch.device("cuda:0")
n = 18837
db = (torch.rand((n, 96, 5), device=dev) * 120 - 120)
out = torch.empty((n, 3, 224, 224), device=dev)
byts = n * (3 * 224 * 224 * 4 + 1920)
if os.environ.get('PATCH_CTAS'): ops.set_option(17, int(os.environ['PATCH_CTAS']))
for _ in range(3): ops.patches(db, out=out)
torch.cuda.synchronize()
ts = []
for _ in range(15):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.patches(db, out=out); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print(json.dumps({"lib": os.environ.get("GTC_LIB_PATH", "libgtc.so").split("/")[-1], "ctas_per_sm": os.environ.get("PATCH_CTAS", "default"), "ms_med": round(float(np.median(ts)), 4),
                  "ms_min": round(min(ts), 4), "GBs_med": round(byts / np.median(ts) / 1e6, 1)}))
