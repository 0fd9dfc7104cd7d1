"""The C-ABI library builds, loads on a GPU-less box, and exports every symbol that
include/mft_gnn.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import mft_b200
    if not os.path.exists(mft_b200.lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    return mft_b200.load_library()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mft_gnn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mft_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(lib):
    from mft_b200 import _lib
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mft_gnn.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == syms


def test_struct_layouts_match_header():
    from mft_b200 import _lib
    p = ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_lib.WcomputeParams) == 14 * p
    assert ctypes.sizeof(_lib.WcomputeGrads) == 18 * p
    assert ctypes.sizeof(_lib.GconvParams) == 4 * p
    assert ctypes.sizeof(_lib.GnnParams) == 3 * 14 * p + 3 * 4 * p
    assert ctypes.sizeof(_lib.GnnGrads) == 3 * 18 * p + 3 * 4 * p


def test_size_queries_and_version(lib):
    assert lib.mft_version() == 1
    # 5-way 20-shot: four pre-BN tensors over B*N(N+1)/2 unordered pairs dominate the tape
    rows = 16 * 105 * 106 // 2
    assert lib.mft_wcompute_saved_bytes(16, 105, 133, 96) >= rows * (192 + 192 + 96 + 96) * 4
    assert lib.mft_gnn_saved_bytes(16, 105, 133, 96, 5) > 3 * lib.mft_wcompute_saved_bytes(16, 105, 133, 96)
    assert lib.mft_gnn_workspace_bytes(16, 30, 133, 96, 5) > 0
    assert lib.mft_last_error() is not None


def test_sass_is_sm100a():
    from mft_b200 import lib_path
    out = subprocess.run(["cuobjdump", "-lelf", lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout
