"""tcgen05 3xTF32 vs fp32 SIMT engine on noise-like segments, for several K-split sizes (GTC_TC_KSPLIT)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "guitar-tablature-classification_b200"))
import numpy as np, torch
from gtc_b200 import ops, CqtRecipe
dev = torch.device("cuda:0")
r = CqtRecipe()
n = 22050 * 60
g = torch.Generator().manual_seed(0)
audio = (0.1 * torch.randn(n, generator=g)).to(dev)
ref_plan = ops.CqtPlan(r, engine=1)
clip_off, seg_off = ref_plan.offsets([n])
co, so = torch.from_numpy(clip_off).to(dev), torch.from_numpy(seg_off).to(dev)
n_seg = int(seg_off[-1])
ref = ref_plan.segments_complex(audio, co, so, n_seg).cpu().numpy().astype(np.complex128)
peak = np.abs(ref).max(axis=(1, 2), keepdims=True)
for eng, ks in ((0, 1000), (0, 8), (2, 1000), (2, 16), (2, 8), (2, 4)):
    os.environ["GTC_TC_KSPLIT"] = str(ks)
    p = ops.CqtPlan(r, engine=eng)
    got = p.segments_complex(audio, co, so, n_seg).cpu().numpy()
    err = np.abs(got - ref) / peak
    keep = np.abs(ref) > peak * 10 ** (-15.5 / 20)
    rel = (np.abs(np.abs(got) - np.abs(ref)) / np.abs(ref))[keep]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    db = torch.empty((n_seg, 96, 5), dtype=torch.float32, device=dev)
    for _ in range(3): p.segments_db(audio, co, so, n_seg, out=db)
    a.record()
    for _ in range(5): p.segments_db(audio, co, so, n_seg, out=db)
    b.record(); torch.cuda.synchronize()
    print(json.dumps({"engine": eng, "ksplit": ks, "max_err_rel_peak": float(err.max()), "rms_err_rel_peak": float(np.sqrt((err ** 2).mean())),
                      "max_rel_mag_err_above_cut": float(rel.max()), "ms": a.elapsed_time(b) / 5, "n_seg": n_seg}))
    p.close()

# quiet audio (amplitude 2e-3): fp16 lo parts go subnormal -- check the scaling keeps precision
audio_q = audio * 0.02
refq = ref_plan.segments_complex(audio_q, co, so, n_seg).cpu().numpy().astype(np.complex128)
peakq = np.abs(refq).max(axis=(1, 2), keepdims=True)
os.environ["GTC_TC_KSPLIT"] = "8"
for eng in (0, 2):
    p = ops.CqtPlan(r, engine=eng)
    got = p.segments_complex(audio_q, co, so, n_seg).cpu().numpy()
    print(json.dumps({"quiet_engine": eng, "max_err_rel_peak": float((np.abs(got - refq) / peakq).max())}))
    p.close()
