"""Per-kernel timing (CUDA events) at GuitarSet-batch scale: python scripts/kbench.py [n_clips] [engine]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "guitar-tablature-classification_b200"))
import numpy as np, torch
from gtc_b200 import ops, synth, CqtRecipe, _lib

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 60
engine = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
r = CqtRecipe()
sr = int(r.sr)
n = sr * 30
t0 = time.time()
audio = synth.pluck_clips(min(n_clips, 8), n, sr=sr, seed=1, device=dev)
audio = audio.repeat((n_clips + 7) // 8, 1)[:n_clips].contiguous().reshape(-1)
torch.cuda.synchronize(); print("synth", time.time() - t0)
plan = ops.CqtPlan(r, engine=engine)
clip_off, seg_off = plan.offsets([n] * n_clips)
n_seg = int(seg_off[-1])
co, so = torch.from_numpy(clip_off).to(dev), torch.from_numpy(seg_off).to(dev)

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

db = torch.empty((n_seg, 96, 5), dtype=torch.float32, device=dev)
best, med = timeit(lambda: plan.segments_db(audio, co, so, n_seg, out=db))
secs = n_clips * 30.0
flop = 2.0 * (n_seg + n_clips) * 960 * 4416
print(json.dumps({"kernel": "cqt_segments_db(frame+gemm+finish)", "engine": engine, "n_seg": n_seg, "ms_best": best, "ms_med": med,
                  "s_audio_per_s": secs / (best * 1e-3), "fp32eq_TFLOPs": flop / (best * 1e-3) / 1e12}))
chunk = min(n_seg, 8192)
pt = torch.empty((chunk, 3, 224, 224), dtype=torch.float32, device=dev)
best, med = timeit(lambda: ops.patches(db[:chunk], out=pt))
byts = chunk * (3 * 224 * 224 * 4 + 1920)
print(json.dumps({"kernel": "patches_vit", "n": chunk, "ms_best": best, "ms_med": med, "GBs": byts / (best * 1e-3) / 1e9,
                  "frac_of_6547": byts / (best * 1e-3) / 1e9 / 6547.2}))
best, med = timeit(lambda: ops.patches(db[:chunk], out=pt, mode=_lib.GTC_PATCH_CNN))
print(json.dumps({"kernel": "patches_cnn", "n": chunk, "ms_best": best, "GBs": byts / (best * 1e-3) / 1e9}))
# labels
durs = [30.0] * n_clips
on, du, pi, eoff = synth.note_events(durs, seed=2)
times = np.concatenate([(np.arange(150) + 0.5) * (30.0 / 150)] * n_clips)
soff = (np.arange(n_clips + 1) * 150).astype(np.int64)
t_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
args = (t_(on), t_(du), t_(pi), t_(eoff), t_(times), t_(soff))
best, med = timeit(lambda: ops.rasterize_tabs(*args))
print(json.dumps({"kernel": "rasterize_tabs", "n_seg": len(times), "n_evt": len(on), "ms_best": best}))
