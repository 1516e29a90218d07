"""The C-ABI library loads on a CPU-only box and exports every symbol include/gtc.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gtc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gtc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for name in ("gtc_cqt_plan_create", "gtc_cqt_segments_db", "gtc_rasterize_tabs", "gtc_patches",
                 "gtc_labels_argmax", "gtc_labels_vit_heads", "gtc_last_error", "gtc_version",
                 "gtc_scqt_plan_create", "gtc_scqt_segments_db", "gtc_scqt_segments_complex"):
        assert name in syms


def test_library_exports_every_declared_symbol(lib):
    from gtc_b200 import _lib
    syms = declared_symbols()
    for name in syms:
        assert hasattr(lib, name), f"libgtc.so does not export {name}"
    assert sorted(_lib.PROTOTYPES) == syms, "ctypes prototypes and include/gtc.h disagree"
    assert lib.gtc_version() == 102


def test_no_torch_types_in_abi():
    text = open(os.path.join(ROOT, "include", "gtc.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)             # comments cite the reference's torch calls; signatures must not
    assert "torch" not in code and "at::" not in code and "Tensor" not in code


def test_compute_call_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        return
    handle = ctypes.c_void_p()
    op = (ctypes.c_float * 16)()
    rc = lib.gtc_cqt_plan_create(ctypes.byref(handle), 0, 8, 4, 1, 1, ctypes.cast(op, ctypes.c_void_p), 1)
    assert rc < 0 and handle.value is None
    assert lib.gtc_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "guitar-tablature-classification_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
