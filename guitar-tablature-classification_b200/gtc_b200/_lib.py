"""ctypes binding of libgtc.so (C ABI declared in include/gtc.h).

There is no CPU fallback: if the shared library is missing or fails to load, importing the compute ops raises.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C guitar-tablature-classification_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GTC_LIB_PATH") or os.path.join(_HERE, "libgtc.so")   # the override is for A/B kernel experiments

GTC_GEMM_TCGEN05_3XTF32 = 0
GTC_GEMM_SIMT_FP32 = 1
GTC_GEMM_TCGEN05_FP16X2 = 2
GTC_OPT_TC_KSPLIT = 1
GTC_OPT_GEMM_MAX_CTAS = 2
GTC_OPT_FUSE_FINISH = 3
GTC_OPT_PATCH_MAX_CTAS = 16
GTC_OPT_PATCH_CTAS_PER_SM = 17
GTC_PATCH_VIT = 0
GTC_PATCH_CNN = 1
GTC_PATCH_VIT_PRENORM = 2
GTC_SAMPLES_F32 = 0
GTC_SAMPLES_PCM16 = 1
GTC_AUG_TIME_SHIFT, GTC_AUG_NOISE, GTC_AUG_FREQ_MASK, GTC_AUG_TIME_MASK = 1, 2, 3, 4

# every symbol include/gtc.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
PROTOTYPES = {
    "gtc_version": (_i, []),
    "gtc_last_error": (C.c_char_p, []),
    "gtc_device_info": (_i, [_i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz)]),
    "gtc_cqt_plan_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _vp, _i]),
    "gtc_cqt_plan_destroy": (_i, [_vp]),
    "gtc_cqt_plan_parts": (_i, [_vp]),
    "gtc_cqt_plan_configure": (_i, [_vp, _i, _i]),
    "gtc_cqt_workspace_bytes": (_i, [_vp, _i64, _i64, C.POINTER(_sz)]),
    "gtc_cqt_segments_db": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _sz, _f, _f, _f, _f, _f, _vp]),
    "gtc_cqt_frame": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _sz, _vp]),
    "gtc_cqt_frame_pcm16": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _sz, _vp]),
    "gtc_cqt_contract_db": (_i, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _sz, _f, _f, _f, _f, _f, _vp]),
    "gtc_set_option": (_i, [_i, _i]),
    "gtc_cqt_segments_complex": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "gtc_scqt_frames": (_i, [_i64, _i, _i]),
    "gtc_scqt_plan_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _i, _vp, _vp, _i]),
    "gtc_scqt_plan_destroy": (_i, [_vp]),
    "gtc_scqt_workspace_bytes": (_i, [_vp, _i64, _i64, C.POINTER(_sz)]),
    "gtc_scqt_decimate": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _f, _vp]),
    "gtc_scqt_segments_db": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _sz, _f, _f, _f, _f, _f, _vp]),
    "gtc_scqt_segments_complex": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "gtc_rasterize_tabs": (_i, [_vp] * 11 + [_i64, _i64, _vp, _vp, _vp]),
    "gtc_labels_argmax": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "gtc_labels_vit_heads": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "gtc_augment_batch": (_i, [_vp, _vp, _i64, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _f, C.c_uint64, _i, _f, _vp]),
    "gtc_db_normalize": (_i, [_vp, _i64, _f, _vp, _vp]),
    "gtc_patches": (_i, [_vp, _vp, _i64, _i, _i, _i, _i, _i, _vp, _vp]),
    "gtc_patches_rgb8": (_i, [_vp, _vp, _i64, _i, _i, _f, _f, _f, _f, _f, _f, _vp, _vp]),
}


class GtcError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load libgtc.so once; raise loudly if it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GtcError(f"{LIB_PATH} not found: build the CUDA extension first (__graft_entry__.build()); "
                       "this package has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gtc_last_error()
        raise GtcError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
