"""GPU: the ctypes stub printed in INTEGRATION.md (what a maintainer of the reference would paste next to cqt.py) is
executed verbatim -- only the library path is made absolute -- and must reproduce the package's own result."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, make_test_audio

pytestmark = pytest.mark.gpu


def test_integration_md_stub_runs_and_matches(lib, recipe):
    from gtc_b200 import _lib, ops
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# cqt_gtc\.py.*?)```", text, re.S).group(1)
    assert 'C.CDLL("libgtc.so")' in block
    ns = {}
    exec(block.replace('C.CDLL("libgtc.so")', f'C.CDLL({_lib.LIB_PATH!r})'), ns)
    sr = 22050
    y = make_test_audio(sr * 2 + 123, seed=99)
    operator = np.ascontiguousarray(ops.get_operator(recipe, recipe.seg_len), dtype=np.float32)
    plan = ns["make_plan"](operator, recipe.seg_len, recipe.seg_hop)
    got = ns["segments_db"](plan, y, sr, recipe.seg_len, recipe.seg_hop)
    mine = ops.CqtPlan(recipe, engine=0)                       # the stub passes gemm_engine = 0 (tcgen05 3xTF32)
    co, so = mine.offsets([len(y)])
    dev = torch.device("cuda")
    want = mine.segments_db(torch.from_numpy(y).to(dev), torch.from_numpy(co).to(dev), torch.from_numpy(so).to(dev), int(so[-1]))
    assert got.shape == (int(so[-1]), 96, 5) and np.array_equal(got, want.cpu().numpy())
    ns["lib"].gtc_cqt_plan_destroy.argtypes = [__import__("ctypes").c_void_p]
    ns["lib"].gtc_cqt_plan_destroy(plan)
