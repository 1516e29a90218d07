"""Golden vectors for the batch augmentations, produced by RUNNING THE REFERENCE'S OWN FUNCTIONS (build container only).

    python tests/golden/make_reference_golden_aug.py        # needs /root/reference; writes tests/golden/ref_augment.npz

The function definitions time_shift / add_noise / frequency_mask / time_mask / augment_batch / db_normalize are taken
from /root/reference/ViT_engine.py with ``ast`` at run time (the module itself cannot be imported: it pulls seaborn,
matplotlib and a ViT checkpoint at import) and executed unmodified with the real ``torch`` and ``random``.  Thin
recording wrappers placed in the same namespace note which ops augment_batch selected and what they drew; they call
the reference functions and do not change results.  Nothing from the reference is copied into the repo.
"""
from __future__ import annotations

import ast
import os
import random

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/ViT_engine.py"
NAMES = ["time_shift", "add_noise", "frequency_mask", "time_mask", "augment_batch", "db_normalize"]
SHAPE = (2, 3, 24, 16)
N_SEEDS = 64


def input_batch():
    g = torch.Generator().manual_seed(1234)
    return (torch.rand(SHAPE, generator=g) * 130.0 - 125.0).float()          # dB-like values in [-125, 5]


def load_reference_functions():
    tree = ast.parse(open(REF).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in NAMES]
    assert sorted(n.name for n in body) == sorted(NAMES)
    ns = {"torch": torch, "random": random}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


def main():
    ns = load_reference_functions()
    log = []
    orig = {k: ns[k] for k in ("time_shift", "add_noise", "frequency_mask", "time_mask")}

    def wrap(name, code):
        def f(x, *a, **kw):
            before = x.clone()
            out = orig[name](x, *a, **kw)
            log.append((code, before, out.clone()))
            return out
        return f

    for code, name in enumerate(("time_shift", "add_noise", "frequency_mask", "time_mask"), start=1):
        ns[name] = wrap(name, code)
    x0 = input_batch()
    outs, ops, shifts, freqs, times, noises = [], [], [], [], [], []
    for seed in range(N_SEEDS):
        random.seed(seed)
        torch.manual_seed(seed)
        log.clear()
        y = ns["augment_batch"](x0.clone())
        o = np.zeros(4, np.int32)
        sh, fr, tm, nz = 0, (0, 0), (0, 0), np.zeros(SHAPE, np.float32)
        k = 0
        for code, before, after in log:
            if code == 1:                                   # recover the drawn shift from the rows that moved
                if torch.equal(before, after):
                    continue
                for s in range(-SHAPE[2] + 1, SHAPE[2]):
                    if s != 0 and torch.equal(torch.roll(before, -s, 2)[:, :, max(0, -s):SHAPE[2] - max(0, s)],
                                              after[:, :, max(0, -s):SHAPE[2] - max(0, s)]):
                        sh = s
                        break
            elif code == 2:
                nz = (after - before).numpy()
            elif code == 3:
                cols = np.where((after != before).any(dim=0).any(dim=0).any(dim=0).numpy())[0]   # columns the mask changed
                fr = (int(cols.min()), int(cols.max() - cols.min() + 1)) if len(cols) else (0, 0)
            elif code == 4:
                rows = np.where((after != before).any(dim=0).any(dim=0).any(dim=1).numpy())[0]   # rows the mask changed
                tm = (int(rows.min()), int(rows.max() - rows.min() + 1)) if len(rows) else (0, 0)
            o[k] = code
            k += 1
        outs.append(y.numpy()); ops.append(o); shifts.append(sh); freqs.append(fr); times.append(tm); noises.append(nz)
    norm = ns["db_normalize"](x0.clone()).numpy()
    np.savez_compressed(os.path.join(HERE, "ref_augment.npz"), x0=x0.numpy(), out=np.stack(outs), ops=np.stack(ops),
                        shift=np.array(shifts, np.int32), freq=np.array(freqs, np.int32), time=np.array(times, np.int32),
                        noise=np.stack(noises).astype(np.float32), db_normalize=norm)
    print("wrote ref_augment.npz:", np.stack(ops)[:10].tolist())


if __name__ == "__main__":
    main()
