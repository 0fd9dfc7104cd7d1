"""tcgen05 building block: one "rows x weights" GEMM with plain operands through the C ABI
(mft_debug_umma_gemm), checked against an fp32 product.  Pins the shared-memory descriptors,
the 128B swizzle, the K/N padding and pass-splitting logic, the mbarrier pipelines and the
TMEM epilogue independently of the edge-MLP fusion.  TF32 inputs (10-bit mantissa, round to
nearest), fp32 accumulate: relative L2 error ~3e-4 expected, 1e-3 asserted."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _umma_gemm(A, W, transpose_w, N, K):
    from mft_b200 import _lib
    lib = _lib.load_library()
    M = A.shape[0]
    out = torch.full((M, N), float("nan"), device="cuda")
    ws = torch.empty(max(256, lib.mft_debug_umma_gemm_workspace_bytes(N, K)), dtype=torch.uint8, device="cuda")
    rc = lib.mft_debug_umma_gemm(A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), int(transpose_w),
                                 out.data_ptr(), out.stride(0), M, N, K, ws.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mft_debug_umma_gemm")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,N,K,transpose_w", [
    (128, 96, 96, False),        # one tile, exact blocks
    (128, 192, 192, False),      # widest resident weight image
    (300, 96, 64, False),        # ragged last row tile
    (1000, 192, 133, False),     # K tail (133 -> 17 k-steps), layer-1 forward shape
    (777, 192, 229, False),      # K = 229: two N passes of 96
    (513, 96, 192, True),        # dgrad operand (W stored [K, N])
    (640, 133, 192, True),       # dgrad layer 1: N = 133 (-> 144), masked tail columns
    (90000, 192, 192, False),    # ~5w20s row count: many tiles per CTA, both accumulators, ring wrap
    (40000, 229, 192, True),     # N = 229 -> two passes of 128
])
def test_umma_gemm_matches_fp32(M, N, K, transpose_w):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(K, N, device="cuda", generator=g) if transpose_w else torch.randn(N, K, device="cuda", generator=g)
    out = _umma_gemm(A, W, transpose_w, N, K)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = (A.double() @ (W.double() if transpose_w else W.double().t())).float()
    assert torch.isfinite(out).all()
    err = float((out - ref).norm() / ref.norm())
    assert err < 1e-3, err
    # worst element relative to the row scale (catches a single wrong tile / column block)
    worst = float(((out - ref).abs().max(dim=1).values / ref.abs().max(dim=1).values).max())
    assert worst < 2e-2, worst


def test_umma_gemm_unaligned_leading_dimension():
    """lda not a multiple of 4 floats: producers must take the scalar path."""
    M, N, K = 257, 96, 133
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g)          # stride 133
    W = torch.randn(N, K, device="cuda", generator=g)
    out = _umma_gemm(A, W, False, N, K)
    ref = (A.double() @ W.double().t()).float()
    assert float((out - ref).norm() / ref.norm()) < 1e-3


@pytest.mark.parametrize("R,Cout,Cin", [
    (32, 96, 96),           # one K chunk
    (1000, 96, 96),         # layer 4
    (4000, 96, 192),        # layer 3: M = 96 (< 128: garbage accumulator rows must stay inside TMEM)
    (5000, 192, 192),       # layer 2: two overlapping M=128 blocks
    (3333, 192, 133),       # layer 1, F = 133 (N = 144, masked tail columns, ragged rows)
    (89040, 192, 229),      # 5w20s row count, F = 229 (N = 240)
    (777, 32, 16),          # tiny-fixture widths (nf = 16)
])
def test_umma_wgrad_matches_fp32(R, Cout, Cin):
    from mft_b200 import _lib
    lib = _lib.load_library()
    g = torch.Generator(device="cuda").manual_seed(R + Cout + Cin)
    P = torch.randn(R, Cout, device="cuda", generator=g)
    Q = torch.randn(R, Cin, device="cuda", generator=g)
    dW = torch.zeros(Cout, Cin, device="cuda")
    rc = lib.mft_debug_umma_wgrad(P.data_ptr(), P.stride(0), Q.data_ptr(), Q.stride(0), dW.data_ptr(),
                                  dW.stride(0), R, Cout, Cin, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "mft_debug_umma_wgrad")
    torch.cuda.synchronize()
    ref = (P.double().t() @ Q.double()).float()
    assert torch.isfinite(dW).all()
    err = float((dW - ref).norm() / ref.norm())
    assert err < 1e-3, err
