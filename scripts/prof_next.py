"""Short driver for ncu --set full captures of the structured-CQT and augmentation kernels (see profiles/r01e_*)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "guitar-tablature-classification_b200"))
import numpy as np, torch
from gtc_b200 import synth, augment
from gtc_b200.inference import TabCnnFrontEnd
dev = torch.device("cuda:0")
fe = TabCnnFrontEnd()
songs, L = 32, 22050 * 60
y = synth.pluck_clips(4, L, sr=22050, seed=2, device=dev).repeat(8, 1).contiguous().reshape(-1)
seg_len, hop = 66150, 33075
s1 = np.arange(0, L, hop)
starts = (np.arange(songs)[:, None] * L + s1[None, :]).reshape(-1)
valid = np.tile(np.minimum(seg_len, L - s1), songs).astype(np.int32)
st, va = torch.from_numpy(starts).to(dev), torch.from_numpy(valid).to(dev)
le = torch.full((len(starts),), seg_len, dtype=torch.int32, device=dev)
for _ in range(2):
    out = fe.plan.segments_db(y, st, va, le, seg_len)
x = torch.rand((1024, 3, 224, 224), device=dev) * 120 - 120
o = torch.empty_like(x)
for _ in range(2):
    augment.apply_ops(x, [1, 3, 4], shift=11, freq=(50, 5), time=(30, 10), out=o)
    augment.apply_ops(x, [2, 4, 3], freq=(50, 5), time=(30, 10), noise_level=0.005, noise_seed=1, normalize_ref_db=-120.0, out=o)
torch.cuda.synchronize()
print("ok", float(out.mean()), float(o.mean()))
