"""Every kernel of libgtc.so once, at sizes small enough for compute-sanitizer (scripts/sanitize.sh):
    compute-sanitizer --tool memcheck  python scripts/sanitize_run.py
    compute-sanitizer --tool racecheck python scripts/sanitize_run.py --no-tma      (the TMA/mbarrier GEMM is skipped)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "guitar-tablature-classification_b200"))
import numpy as np, torch
from gtc_b200 import ops, synth, augment, CqtRecipe, _lib
from gtc_b200.pipeline import FrontEnd, ShardInputs

no_tma = "--no-tma" in sys.argv
dev = torch.device("cuda:0")
r = CqtRecipe(); SR = int(r.sr)
rng = np.random.default_rng(0)
lens = np.array([SR * 2, 4000, SR + 777, 4410], dtype=np.int64)            # incl. a clip without a complete window
audio = (rng.standard_normal(int(lens.sum())) * 0.1).astype(np.float32)
a_dev = torch.from_numpy(audio).to(dev)
pcm = torch.clamp(torch.round(a_dev * 32768), -32768, 32767).to(torch.int16)
done = []

engines = [_lib.GTC_GEMM_SIMT_FP32] if no_tma else [_lib.GTC_GEMM_SIMT_FP32, _lib.GTC_GEMM_TCGEN05_3XTF32, _lib.GTC_GEMM_TCGEN05_FP16X2]
for eng in engines:
    plan = ops.CqtPlan(r, engine=eng)
    co, so = plan.offsets(lens)
    n_seg = int(so[-1])
    co_t, so_t = torch.from_numpy(co).to(dev), torch.from_numpy(so).to(dev)
    db = plan.segments_db(a_dev, co_t, so_t, n_seg)
    db2 = plan.segments_db(pcm, co_t, so_t, n_seg)
    cx = plan.segments_complex(a_dev, co_t, so_t, n_seg)
    if eng != _lib.GTC_GEMM_SIMT_FP32:
        plan.configure(_lib.GTC_OPT_FUSE_FINISH, 1)
        db3 = plan.segments_db(a_dev, co_t, so_t, n_seg)
        assert torch.equal(db, db3)
    torch.cuda.synchronize(); plan.close(); done.append(f"cqt engine {eng}: {n_seg} segments")

if no_tma:
    os.environ["GTC_SCQT_SIMT"] = "1"            # structured CQT on its fp32 CUDA-core kernels (no TMA / tcgen05 under racecheck)
sp = ops.StructuredCqtPlan(r)                     # default: decimator + response GEMMs on the slotted tcgen05 kernels
starts = torch.tensor([0, 2205, int(lens[0]), int(lens[0]) + 100], dtype=torch.int64, device=dev)
valid = torch.tensor([4410, 4410, 3000, 1234], dtype=torch.int32, device=dev)
seglen = torch.tensor([4410, 4410, 4410, 4410], dtype=torch.int32, device=dev)
sdb = sp.segments_db(a_dev, starts, valid, seglen, 4410)
scx = sp.segments_complex(pcm, starts, valid, seglen, 4410)
half = sp.halve_rate(a_dev[:10001])
torch.cuda.synchronize(); sp.close(); done.append("structured cqt + decimator")
if not no_tma:
    os.environ["GTC_TC_DENSE"] = "1"              # the dense loops of the tensor-core engine (default plans use the zero-skipping schedule)
    plan = ops.CqtPlan(r)
    co, so = plan.offsets(lens)
    plan.segments_db(a_dev, torch.from_numpy(co).to(dev), torch.from_numpy(so).to(dev), int(so[-1]))
    torch.cuda.synchronize(); plan.close(); del os.environ["GTC_TC_DENSE"]; done.append("dense tensor-core loops")
    os.environ["GTC_SCQT_SIMT"] = "1"
    sp = ops.StructuredCqtPlan(r)
    sp.segments_db(a_dev, starts, valid, seglen, 4410)
    torch.cuda.synchronize(); sp.close(); del os.environ["GTC_SCQT_SIMT"]; done.append("structured cqt, fp32 SIMT kernels")

on, du, pi, eoff = synth.note_events([n / SR for n in lens], seed=3)
plan = ops.CqtPlan(r, engine=_lib.GTC_GEMM_SIMT_FP32)
co, so = plan.offsets(lens); n_seg = int(so[-1])
t_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
times = np.concatenate([(np.arange(int(so[c + 1] - so[c])) + 0.5) * (lens[c] / SR / max(1, int(so[c + 1] - so[c]))) for c in range(len(lens))])
tabs, stats = ops.rasterize_tabs(t_(on), t_(du), t_(pi), t_(eoff), t_(times), t_(so))
idx = torch.tensor([5, 0, 3, 3, n_seg - 1], dtype=torch.int64, device=dev)
ops.labels_argmax(tabs, idx); ops.labels_vit_heads(tabs, idx); ops.labels_argmax(tabs); ops.labels_vit_heads(tabs)
torch.cuda.synchronize(); done.append(f"labels: {n_seg} segments, {len(on)} notes")

db = plan.segments_db(a_dev, t_(co), t_(so), n_seg)
for mode in (_lib.GTC_PATCH_VIT, _lib.GTC_PATCH_CNN, _lib.GTC_PATCH_VIT_PRENORM):
    ops.patches(db, mode=mode); ops.patches(db, index=idx, mode=mode)
ops.patches(db, img_size=(37, 50))                                           # generic gather path
rgb = torch.randint(0, 256, (7, 224, 224, 3), dtype=torch.uint8, device=dev)
ops.patches_rgb8(rgb); ops.patches_rgb8(rgb, index=torch.tensor([6, 1], dtype=torch.int64, device=dev))
torch.cuda.synchronize(); done.append("patches (fast, generic, rgb8)")

x = torch.rand((6, 3, 224, 224), device=dev) * 120 - 120
augment.apply_ops(x, [1, 3, 4], shift=11, freq=(50, 5), time=(30, 10))
augment.apply_ops(x, [2, 4, 3], freq=(50, 5), time=(30, 10), noise_level=0.005, noise_seed=1, normalize_ref_db=-120.0)
augment.db_normalize(x); augment.db_normalize(x.reshape(-1)[:1001])
torch.cuda.synchronize(); done.append("augmentation + db_normalize")

if not no_tma:
    fe = FrontEnd(r, chunk_segments=12, patch_batch=8)
    ev = np.stack([on, du, pi])
    host = ShardInputs(torch.from_numpy(audio).pin_memory(), lens, torch.from_numpy(ev).pin_memory(), eoff, sr=SR)
    o1 = fe.run(host, next_inp=host); o2 = fe.run(host)
    o3 = fe.run(ShardInputs(a_dev, lens, torch.from_numpy(ev).to(dev), eoff, sr=SR), device_inputs=True)
    torch.cuda.synchronize(); done.append(f"pipeline host/prefetch/device: {o3.n_seg} segments")
print("sanitize_run ok:", "; ".join(done))
