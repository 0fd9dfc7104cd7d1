"""ctypes binding of libmft_gnn.so (C ABI declared in include/mft_gnn.h).

There is no other compute path: if the shared library is missing the first call
raises ``LibraryMissing`` -- build it with ``python -c "import __graft_entry__ as
g; g.build()"`` or ``make -C meta-fine-tuning_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libmft_gnn.so"

PREC_FP32 = 0
PREC_TF32 = 1
MAX_LAYERS = 3

_fp = C.c_void_p   # device pointers travel as plain addresses


class WcomputeParams(C.Structure):
    _fields_ = [("conv_w", _fp * 4), ("bn_g", _fp * 4), ("bn_b", _fp * 4), ("last_w", _fp), ("last_b", _fp)]


class WcomputeGrads(C.Structure):
    _fields_ = [("conv_w", _fp * 4), ("conv_b", _fp * 4), ("bn_g", _fp * 4), ("bn_b", _fp * 4),
                ("last_w", _fp), ("last_b", _fp)]


class GconvParams(C.Structure):
    _fields_ = [("fc_w", _fp), ("fc_b", _fp), ("bn_g", _fp), ("bn_b", _fp)]


class GconvGrads(C.Structure):
    _fields_ = [("fc_w", _fp), ("fc_b", _fp), ("bn_g", _fp), ("bn_b", _fp)]


class GnnParams(C.Structure):
    _fields_ = [("w", WcomputeParams * MAX_LAYERS), ("l", GconvParams * MAX_LAYERS)]


class GnnGrads(C.Structure):
    _fields_ = [("w", WcomputeGrads * MAX_LAYERS), ("l", GconvGrads * MAX_LAYERS)]


class LibraryMissing(RuntimeError):
    pass


# name -> (restype, argtypes); every symbol include/mft_gnn.h declares
_i, _sz, _vp = C.c_int, C.c_size_t, C.c_void_p
SIGNATURES = {
    "mft_last_error": (C.c_char_p, []),
    "mft_version": (_i, []),
    "mft_device_check": (_i, [_i]),
    "mft_tf32_supported": (_i, [_i, _i]),
    "mft_wcompute_saved_bytes": (_sz, [_i, _i, _i, _i]),
    "mft_wcompute_saved_bytes_for": (_sz, [_i, _i, _i, _i, _i]),
    "mft_wcompute_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "mft_wcompute_fwd": (_i, [_vp, _i, _i, _i, _i, _i, C.POINTER(WcomputeParams), _vp, _vp, _vp, _i, C.c_char_p,
                              _vp]),
    "mft_wcompute_bwd": (_i, [_vp, _i, _i, _i, _i, _i, C.POINTER(WcomputeParams), _vp, _vp, _vp,
                              C.POINTER(WcomputeGrads), _vp, _vp, _i, C.c_char_p, _vp]),
    "mft_gconv_saved_bytes": (_sz, [_i, _i, _i, _i]),
    "mft_gconv_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "mft_gconv_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.POINTER(GconvParams), _i, _vp, _i, _vp, _vp, _vp]),
    "mft_gconv_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, C.POINTER(GconvParams), _i, _vp, _i, _vp, _vp,
                           C.POINTER(GconvGrads), _vp, _vp, _vp]),
    "mft_gnn_saved_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "mft_gnn_saved_bytes_for": (_sz, [_i, _i, _i, _i, _i, _i]),
    "mft_gnn_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "mft_gnn_fwd": (_i, [_vp, _i, _i, _i, _i, _i, C.POINTER(GnnParams), _vp, _vp, _vp, _i, C.c_char_p, _vp]),
    "mft_gnn_bwd": (_i, [_vp, _i, _i, _i, _i, _i, C.POINTER(GnnParams), _vp, C.POINTER(GnnGrads), _vp, _vp,
                         _i, C.c_char_p, _vp]),
    "mft_head_saved_bytes": (_sz, [_i, _i, _i, _i]),
    "mft_head_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "mft_query_ce": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "mft_head_fwd": (_i, [_vp, _i, _i, _i, _i, _i, C.POINTER(GconvParams), _vp, _vp, _vp, _vp]),
    "mft_head_bwd": (_i, [_vp, _i, _i, _i, _i, _i, C.POINTER(GconvParams), _vp, _vp, C.POINTER(GconvGrads), _vp, _vp,
                          _vp]),
    "mft_debug_umma_gemm_workspace_bytes": (_sz, [_i, _i]),
    "mft_debug_umma_gemm": (_i, [_vp, _i, _vp, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "mft_debug_umma_wgrad": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "mft_debug_set_timeline": (_i, [_vp, _i]),
    "mft_debug_wcompute_saved_offsets": (_i, [_i, _i, _i, _i, _i, C.POINTER(C.c_size_t)]),
    "mft_launch_count": (C.c_ulonglong, []),
    "mft_prof_enable": (_i, [_i]),
    "mft_set_pdl": (_i, [_i]),
    "mft_prof_categories": (_i, []),
    "mft_prof_name": (C.c_char_p, [_i]),
    "mft_prof_collect": (_i, [C.POINTER(C.c_float), C.POINTER(C.c_int), _i]),
}

_lock = threading.Lock()
_lib = None


def lib_path() -> str:
    return os.path.join(_HERE, _LIB_NAME)


def load_library():
    """Load (once) and type the C ABI.  Raises LibraryMissing when the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not os.path.exists(path):
            raise LibraryMissing(
                f"{path} not found: the CUDA extension is not built and there is no other compute path "
                f"(run __graft_entry__.build() or `make -C {os.path.join(_HERE, 'csrc')}`)")
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError here = header/library drift
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().mft_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def profile_collect():
    """{category: (total_ms, launches)} of the scopes recorded since mft_prof_enable(1)."""
    lib = load_library()
    n = lib.mft_prof_categories()
    ms = (C.c_float * n)()
    cnt = (C.c_int * n)()
    check(lib.mft_prof_collect(ms, cnt, n), "mft_prof_collect")
    return {lib.mft_prof_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i] > 0}
