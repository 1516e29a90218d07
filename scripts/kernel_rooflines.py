"""Every kernel of the main path alone on the GPU at the chunk size bench.py uses (63 clips x 30 s = 18 837 segments):
CUDA-event time, algorithmic bytes or flops, and the fraction of the measured peak (MEASURED_PEAKS.json).
    python scripts/kernel_rooflines.py > profiles/<tag>_kernel_rooflines.md"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "guitar-tablature-classification_b200"))
import numpy as np, torch
from gtc_b200 import ops, synth, CqtRecipe, _lib

dev = torch.device("cuda:0")
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    PEAKS = {}
HBM, TC = float(PEAKS.get("hbm_gbs", 6650.0)), float(PEAKS.get("bf16_tflops", 1665.0))
r = CqtRecipe(); SR = int(r.sr); n = SR * 30; n_clips = 63


def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))


rows = []
def hbm_row(name, ms, byts, note=""):
    gbs = byts / (ms * 1e-3) / 1e9
    rows.append(f"| `{name}` | HBM | {ms * 1e3:.1f} | {byts / 1e6:.1f} MB | {gbs:.0f} GB/s | {gbs / HBM:.2f} | {note} |")

audio = synth.pluck_clips(8, n, sr=SR, seed=1, device=dev).repeat(8, 1)[:n_clips].contiguous().reshape(-1)
pcm = torch.clamp(torch.round(audio * 32768.0), -32768, 32767).to(torch.int16)
plan = ops.CqtPlan(r)
clip_off, seg_off = plan.offsets([n] * n_clips)
n_seg = int(seg_off[-1]); n_rows = n_seg + n_clips * (plan.parts - 1)
co, so = torch.from_numpy(clip_off).to(dev), torch.from_numpy(seg_off).to(dev)
ws = plan.workspace(n_seg, n_clips)
db = torch.empty((n_seg, 96, 5), device=dev)
kp = 2208                                                     # fp16 operand row: 2205 samples padded to a 64-byte multiple
rows_alloc = n_rows + 128
ms = timeit(lambda: plan.frame(audio, co, so, n_seg, ws))
hbm_row("frame_kernel<half,float>", ms, n_rows * 2205 * 4 + 2 * rows_alloc * kp * 2, "fp32 audio in, fp16 hi+lo rows out")
ms = timeit(lambda: plan.frame(pcm, co, so, n_seg, ws))
hbm_row("frame_kernel<half,short>", ms, n_rows * 2205 * 2 + 2 * rows_alloc * kp * 2, "int16 PCM in (the e2e arm)")
plan.frame(audio, co, so, n_seg, ws)
ms_c = timeit(lambda: plan.contract_db(co, so, n_seg, db, ws))
flop = 3 * 2.0 * n_rows * 960 * 2 * kp
rows.append(f"| `gemm_tc_kernel<240,0,1,5>` + `finish_db_kernel` | tensor | {ms_c * 1e3:.1f} | {flop / 1e9:.0f} GFLOP fp16 issued "
            f"({flop / 3e9:.0f} fp32-equivalent) | {flop / (ms_c * 1e-3) / 1e12:.0f} TFLOP/s | {flop / (ms_c * 1e-3) / 1e12 / TC:.2f} of cuBLAS bf16 burst | "
            f"two launches timed together; finish alone is 19.5 us per 28 200-row chunk in the ncu launch list (profiles/r02e_gemm_finish_ab.md) |")
on, du, pi, eoff = synth.note_events([30.0] * n_clips, seed=2)
t_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
times = np.concatenate([(np.arange(299) + 0.5) * (30.0 / 299)] * n_clips)
args = (t_(on), t_(du), t_(pi), t_(eoff), t_(times), so)
tabs = torch.empty((n_seg, 6, 19), dtype=torch.int8, device=dev); stats = torch.zeros(3, dtype=torch.int64, device=dev)
ms = timeit(lambda: ops.rasterize_tabs(*args, out=tabs, stats=stats))
rows.append(f"| `rasterize_kernel` | latency | {ms * 1e3:.1f} | {(len(on) * 24 + n_seg * 122) / 1e6:.1f} MB | - | - | "
            f"{n_seg} segments x {len(on) // n_clips} notes per clip scanned by one warp each; {n_seg * (len(on) // n_clips) / (ms * 1e-3) / 1e9:.0f} G interval tests/s |")
out = torch.empty((n_seg, 3, 224, 224), device=dev)
for mode, name in ((_lib.GTC_PATCH_VIT, "patch_kernel<5> ViT bicubic"), (_lib.GTC_PATCH_CNN, "patch_kernel<5> CNN bilinear+ImageNet")):
    ms = timeit(lambda: ops.patches(db, out=out, mode=mode))
    hbm_row(name, ms, n_seg * (3 * 224 * 224 * 4 + 1920), "pure store stream")
f = out.view(-1)
ms = timeit(lambda: f.fill_(1.5))
hbm_row("torch fill_ (reference point)", ms, f.numel() * 4, "pure-store ceiling of this GPU")
g = torch.empty_like(f)
ms = timeit(lambda: g.copy_(f))
hbm_row("torch copy_ (reference point)", ms, 2 * f.numel() * 4, "what MEASURED_PEAKS.json hbm_gbs measures")

print(f"# Per-kernel rooflines, main path, one 63-clip chunk ({n_seg} segments, {n_rows} operand rows) alone on the GPU\n")
print(f"Peaks: HBM {HBM:.0f} GB/s (measured copy), tensor {TC:.0f} TFLOP/s (cuBLAS bf16 burst) -- MEASURED_PEAKS.json.\n")
print("| kernel | bound | us | algorithmic work | achieved | fraction of peak | note |\n|---|---|---|---|---|---|---|")
print("\n".join(rows))
