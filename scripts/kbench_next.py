"""Per-kernel timing (CUDA events, L2-exceeding working sets) of the round-1 "next rows": structured CQT, augmentation.
    python scripts/kbench_next.py            # JSON lines; run under ncu for the launch list (see profiles/)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "guitar-tablature-classification_b200"))
import numpy as np, torch
from gtc_b200 import ops, synth, CqtRecipe, augment, _lib
from gtc_b200.inference import TabCnnFrontEnd

dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def timeit(fn, reps=7, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), min(ts)


def report(name, ms, best, algo_bytes=None, **kw):
    d = {"kernel": name, "ms_mean": round(ms, 4), "ms_best": round(best, 4)}
    if algo_bytes:
        d["algo_GBs"] = round(algo_bytes / (ms * 1e-3) / 1e9, 1)
        d["frac_of_measured_hbm_peak"] = round(d["algo_GBs"] / PEAK, 3)
    d.update(kw)
    print(json.dumps(d), flush=True)


# ---- augmentation: (B,3,224,224) read once + written once -------------------------------------------------------------
for B in (128, 2048):
    x = torch.rand((B, 3, 224, 224), device=dev) * 120 - 120
    out = torch.empty_like(x)
    byts = 2 * x.numel() * 4
    ms, best = timeit(lambda: augment.apply_ops(x, [1, 3, 4], shift=11, freq=(50, 5), time=(30, 10), out=out))
    report(f"augment shift+masks B={B}", ms, best, byts)
    ms, best = timeit(lambda: augment.apply_ops(x, [2, 4, 3], freq=(50, 5), time=(30, 10), noise_level=0.005, noise_seed=1, normalize_ref_db=-120.0, out=out))
    report(f"augment noise+masks+db_normalize B={B}", ms, best, byts)
    ms, best = timeit(lambda: augment.db_normalize(x))
    report(f"db_normalize B={B} (incl. torch.empty_like)", ms, best, byts)
    # what the reference's torch op chain costs on the same GPU (time_shift + frequency_mask + time_mask + db_normalize)
    def torch_chain():
        y = torch.cat([x[:, :, 11:, :], torch.zeros_like(x[:, :, :11, :])], dim=2)
        y[:, :, :, 50:55] = 0
        y[:, :, 30:40, :] = 0
        return torch.clamp((y + 120.0) / 120.0, 0, 1)
    ms, best = timeit(torch_chain)
    report(f"torch op chain (ViT_engine.py ops on GPU) B={B}", ms, best, byts)
    del x, out

# ---- structured CQT at the cqt.py recipe: 16384 segments of 4410 samples -----------------------------------------------
r = CqtRecipe()
n_clips, n = 55, 22050 * 30
audio = synth.pluck_clips(8, n, sr=22050, seed=1, device=dev).repeat(7, 1)[:n_clips].contiguous().reshape(-1)
per = (n - r.seg_len) // r.seg_hop + 1
starts = (np.arange(n_clips)[:, None] * n + np.arange(per)[None, :] * r.seg_hop).reshape(-1)
n_seg = len(starts)
st = torch.from_numpy(starts).to(dev)
le = torch.full((n_seg,), r.seg_len, dtype=torch.int32, device=dev)
sp = ops.StructuredCqtPlan(r)
out = torch.empty((n_seg, 96, 5), device=dev)
ms, best = timeit(lambda: sp.segments_db(audio, st, le, le, r.seg_len, out=out))
report("structured cqt.py recipe (7 decimations + 8 responses + finish)", ms, best, None, n_seg=n_seg,
       s_audio_per_s=round(n_seg * 0.1 / (ms * 1e-3)))
cp = ops.CqtPlan(r)
clip_off, seg_off = cp.offsets([n] * n_clips)
co_, so_ = torch.from_numpy(clip_off).to(dev), torch.from_numpy(seg_off).to(dev)
ms, best = timeit(lambda: cp.segments_db(audio, co_, so_, n_seg, out=out))
report("collapsed operator (frame + tcgen05 fp16x2 GEMM + finish), same input", ms, best, None, n_seg=n_seg,
       s_audio_per_s=round(n_seg * 0.1 / (ms * 1e-3)))

# ---- inference recipe: 3 s segments, 50 % overlap, of 64 songs x 60 s --------------------------------------------------
fe = TabCnnFrontEnd()
songs, L = 64, 22050 * 60
y = synth.pluck_clips(8, L, sr=22050, seed=2, device=dev).repeat(8, 1).contiguous().reshape(-1)
seg_len, hop = 66150, 33075
s1 = np.arange(0, L, hop)
starts = (np.arange(songs)[:, None] * L + s1[None, :]).reshape(-1)
valid = np.tile(np.minimum(seg_len, L - s1), songs).astype(np.int32)
n_seg = len(starts)
st, va = torch.from_numpy(starts).to(dev), torch.from_numpy(valid).to(dev)
le = torch.full((n_seg,), seg_len, dtype=torch.int32, device=dev)
out = torch.empty((n_seg, 84, 130), device=dev)
ms, best = timeit(lambda: fe.plan.segments_db(y, st, va, le, seg_len, out=out))
report("structured inference recipe (3 s segments, C2, 84 bins, hop 512)", ms, best, None, n_seg=n_seg,
       s_audio_per_s=round(songs * 60.0 / (ms * 1e-3)), segment_s_per_s=round(n_seg * 3.0 / (ms * 1e-3)))
# whole songs as single segments (60 s each)
st = torch.from_numpy(np.arange(songs) * L).to(dev)
ln = torch.full((songs,), L, dtype=torch.int32, device=dev)
ms, best = timeit(lambda: fe.plan.segments_db(y, st, ln, ln, L))
report("structured whole-clip CQT (64 x 60 s)", ms, best, None, s_audio_per_s=round(songs * 60.0 / (ms * 1e-3)))
