"""GPU parity: the CUDA path (through the nn.Module surface -> ctypes -> C ABI) against
the reference goldens and against the CPU oracle on seeded inputs.

Everything is measured as relative L2 error against a float64 evaluation of the reference
function.  Tolerances of the fp32 path:

  logits                       <= 1e-5   (north_star: fp32 path), every shape
  gradients, kink-free inputs  <= max(3 x the reference's own fp32 error on that tensor, 2e-5)
  gradients, large shapes      <= max(3 x the reference's own fp32 error, 3e-2)

Why two gradient bars: LeakyReLU makes the gradient a discontinuous function of the
inputs.  With ~10^7 BatchNorm outputs in a full-size head a handful always lie within float32
rounding distance of zero, and any two float32 evaluations -- including the reference's own
(see e32.* in tests/golden/gnn_5w5s.npz: 3e-3 on layer_w0.conv2d_1/2, 1e-6 elsewhere) -- may
pick different slopes for them; one flipped element moves a parameter gradient by O(1/sqrt(pairs))
~ 1e-3..1e-2.  Sharp gradient parity is therefore asserted on seeded inputs that the float64
oracle certifies kink-free (min |BN output| > 3e-5, tests/util.kink_free_problem), through the
very same kernels; at full size gradients get the loose bar plus exact structural checks
(linearity of the backward in the upstream gradient).
"""
import os

import numpy as np
import pytest
import torch

from oracle import gnn_oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu

OUT_TOL_FP32 = 1e-5
GRAD_FACTOR = 3.0
GRAD_FLOOR = 2e-5        # kink-free inputs
GRAD_FLOOR_KINK = 3e-2   # inputs with pre-activations within rounding distance of the LeakyReLU kink


def _golden(golden_dir, name):
    rec = dict(np.load(os.path.join(golden_dir, name)))
    params = {k[2:]: v for k, v in rec.items() if k.startswith("p.")}
    return rec, params


@pytest.mark.parametrize("name,fin,nf,n_way", [("gnn_tiny.npz", 13, 16, 3), ("gnn_5w5s.npz", 133, 96, 5),
                                               ("gnn_5w20s.npz", 133, 96, 5)])   # 5w20s: B=16, N=105, the benchmarked shape
@pytest.mark.parametrize("fused", [True, False])
def test_golden_fp32(golden_dir, name, fin, nf, n_way, fused):
    rec, params = _golden(golden_dir, name)
    got = U.run_cuda_gnn(rec["x"], params, rec["proj"], fin, nf, n_way, "fp32", fused)
    truth = (rec["out64"], rec["dx64"], {k: rec["g." + k].astype(np.float64) for k in params})
    out, dx, grads = got
    # gnn_tiny is kink-free at fp32 resolution (the reference's own fp32 gradients agree to 1e-6)
    floor = GRAD_FLOOR if name == "gnn_tiny.npz" else GRAD_FLOOR_KINK
    assert U.rel(out, truth[0]) < OUT_TOL_FP32
    assert U.rel(dx, truth[1]) < max(GRAD_FACTOR * U.rel(rec["dx32"], rec["dx64"]), floor)
    for k in params:
        g = grads[k].reshape(truth[2][k].shape)
        if U.is_zero_grad(k):
            assert np.abs(g).max() <= 1e-6, k
        else:
            lim = max(GRAD_FACTOR * float(rec["e32." + k]), floor)
            assert U.rel(g, truth[2][k]) < lim, (k, U.rel(g, truth[2][k]), lim)


@pytest.mark.parametrize("bsz,n,fin,nf,n_way,seed0", [
    (1, 2, 5, 4, 2, 0),          # smallest graph: one off-diagonal pair
    (3, 9, 21, 12, 4, 1),        # odd sizes everywhere (ragged tiles, K % 16 != 0)
    (2, 12, 40, 24, 5, 2),
    (1, 6, 133, 96, 5, 3),       # the real channel widths (F=133..229, 192/192/96/96), tiny graphs
    (2, 5, 133, 96, 5, 4),
])
@pytest.mark.parametrize("fused", [True, False])
def test_kink_free_sharp_gradients_fp32(bsz, n, fin, nf, n_way, seed0, fused):
    params, x, proj, seed, margin = U.kink_free_problem(bsz, n, fin, nf, n_way, seed0, tau=3e-5)
    truth = U.oracle_truth(x, params, proj, torch.float64)
    yard = U.oracle_truth(x, params, proj, torch.float32)
    got = U.run_cuda_gnn(x, params, proj, fin, nf, n_way, "fp32", fused)
    U.check_against_truth(got, truth, yard, OUT_TOL_FP32, GRAD_FACTOR, GRAD_FLOOR, f"B{bsz}N{n}s{seed}")


@pytest.mark.parametrize("bsz,n,fin,nf,n_way,seed", [
    (16, 30, 133, 96, 5, 3),     # 5-way 5-shot, train shape
    (15, 30, 133, 96, 5, 4),     # 5-way 5-shot, test shape (finetune.py: 15 queries)
    (4, 130, 133, 96, 5, 5),     # compressed 50-shot node count, fewer graphs (oracle time)
])
def test_seeded_vs_oracle_fp32(bsz, n, fin, nf, n_way, seed):
    p64 = O.random_params(fin, nf, n_way, seed, torch.float64)
    p32 = {k: v.float() for k, v in p64.items()}
    params = {k: v.numpy() for k, v in p32.items()}
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randn(bsz, n, fin, generator=g)
    proj = torch.randn(bsz, n, n_way, generator=g)
    truth = U.oracle_truth(x, params, proj, torch.float64)
    yard = U.oracle_truth(x, params, proj, torch.float32)
    got = U.run_cuda_gnn(x, params, proj, fin, nf, n_way, "fp32", True)
    U.check_against_truth(got, truth, yard, OUT_TOL_FP32, GRAD_FACTOR, GRAD_FLOOR_KINK, f"B{bsz}N{n}")


def test_backward_is_linear_in_upstream_gradient():
    """For a fixed forward the backward is a linear map of d_out: grads(a*u + v) = a*grads(u) +
    grads(v) up to float32 rounding.  Checked at the 5-way 20-shot size (B=16, N=105)."""
    import mft_b200
    fin, nf, n_way, bsz, n = 133, 96, 5, 16, 105
    mft_b200.set_precision("fp32")
    torch.manual_seed(5)
    net = mft_b200.GNN_nl(fin, nf, n_way).cuda()
    x = torch.randn(bsz, n, fin, device="cuda", requires_grad=True)
    out = net(x)
    u = torch.randn_like(out)
    v = torch.randn_like(out)
    prm = [x] + list(net.parameters())

    def grads(d):
        return torch.autograd.grad(out, prm, d, retain_graph=True)

    gu, gv, guv = grads(u), grads(v), grads(2.0 * u + v)
    for a, b, c, t in zip(gu, gv, guv, prm):
        want = 2.0 * a + b
        den = float(want.norm())
        if den < 1e-6:
            assert float(c.abs().max()) <= 1e-6
        else:
            assert float((c - want).norm()) / den < 2e-4, tuple(t.shape)


def test_full_size_5w20s_forward_and_properties():
    """5-way 20-shot (B=16, N=105): forward against the oracle, plus size-independent
    properties of the adjacency the kernels produce."""
    import mft_b200
    fin, nf, n_way, bsz, n = 133, 96, 5, 16, 105
    p64 = O.random_params(fin, nf, n_way, 11, torch.float64)
    params = {k: v.float().numpy() for k, v in p64.items()}
    g = torch.Generator().manual_seed(2024)
    x = torch.randn(bsz, n, fin, generator=g)
    with torch.no_grad():
        out_t = O.gnn_nl(x.double(), {k: torch.as_tensor(v).double() for k, v in params.items()}).numpy()
    mft_b200.set_precision("fp32")
    net = U.load_params_into(mft_b200.GNN_nl(fin, nf, n_way), params).cuda()
    with torch.no_grad():
        out = net(x.cuda())
        adj = net.layer_w0.adjacency(x.cuda())
    assert U.rel(out.double().cpu().numpy(), out_t) < OUT_TOL_FP32
    assert torch.allclose(adj.sum(2), torch.ones(bsz, n, device="cuda"), atol=1e-5)
    assert float(adj.diagonal(dim1=1, dim2=2).abs().max()) == 0.0
    # permuting the nodes of a graph permutes the adjacency (edge MLP is permutation equivariant
    # because the batch statistics are permutation invariant)
    perm = torch.randperm(n, generator=g)
    with torch.no_grad():
        adj_p = net.layer_w0.adjacency(x[:, perm].cuda())
    assert torch.allclose(adj_p, adj[:, perm][:, :, perm], atol=2e-6)


# ------------------------------------------------------------------------------------------
# tcgen05 TF32 path.  north_star: logits within rel 1e-3 of the reference.  Gradients: TF32
# rounding moves every pre-activation by ~1e-3 relative, so LeakyReLU slope flips are the rule,
# not the exception; SURVEY.md 7.4 measured 1e-2 (dx) / 3e-2 median, 6.5e-2 worst (parameters) for
# a TF32-emulated run of the reference itself.  Asserted: finite, and <= 0.15 per tensor.
# ------------------------------------------------------------------------------------------
OUT_TOL_TF32 = 1e-3
GRAD_TOL_TF32 = 0.15


def _grad_tol_tf32(rows):
    """One flipped LeakyReLU slope moves a parameter gradient by O(1/sqrt(rows)); on toy graphs
    (84 pair rows in gnn_tiny) that alone is ~0.1-0.2, on real heads (>= 7440 rows) it is below 0.15."""
    return max(GRAD_TOL_TF32, 2.0 / np.sqrt(rows))


@pytest.mark.parametrize("name,fin,nf,n_way", [("gnn_tiny.npz", 13, 16, 3), ("gnn_5w5s.npz", 133, 96, 5),
                                               ("gnn_5w20s.npz", 133, 96, 5)])
def test_golden_tf32(golden_dir, name, fin, nf, n_way):
    rec, params = _golden(golden_dir, name)
    out, dx, grads = U.run_cuda_gnn(rec["x"], params, rec["proj"], fin, nf, n_way, "tf32", True)
    bsz, n = rec["x"].shape[:2]
    tol = _grad_tol_tf32(bsz * n * (n + 1) // 2)
    assert U.rel(out, rec["out64"]) < OUT_TOL_TF32
    assert U.rel(dx, rec["dx64"]) < tol
    for k in params:
        g = grads[k].reshape(rec["g." + k].shape)
        assert np.isfinite(g).all(), k
        if U.is_zero_grad(k):
            assert np.abs(g).max() <= 1e-6, k
        else:
            assert U.rel(g, rec["g." + k]) < tol, (k, U.rel(g, rec["g." + k]))


@pytest.mark.parametrize("bsz,n,seed", [(16, 30, 3), (15, 105, 4), (6, 130, 5)])
def test_seeded_vs_oracle_tf32_logits(bsz, n, seed):
    """Forward parity of the tensor-core path at the three head sizes (5-shot, 20-shot, compressed
    50-shot node counts), F = 133/181/229 exercising the K-tail and N-pass-split code paths."""
    import mft_b200
    fin, nf, n_way = 133, 96, 5
    p64 = O.random_params(fin, nf, n_way, seed, torch.float64)
    params = {k: v.float().numpy() for k, v in p64.items()}
    g = torch.Generator().manual_seed(77 + seed)
    x = torch.randn(bsz, n, fin, generator=g)
    with torch.no_grad():
        out_t = O.gnn_nl(x.double(), {k: torch.as_tensor(v).double() for k, v in params.items()}).numpy()
    mft_b200.set_precision("tf32")
    net = U.load_params_into(mft_b200.GNN_nl(fin, nf, n_way), params).cuda()
    with torch.no_grad():
        out = net(x.cuda())
    mft_b200.set_precision("auto")
    assert U.rel(out.double().cpu().numpy(), out_t) < OUT_TOL_TF32


@pytest.mark.parametrize("bsz,n,seed", [(16, 30, 3), (8, 105, 4), (4, 130, 5)])
def test_tf32_gradients_vs_emulating_oracle(bsz, n, seed):
    """End-to-end gradients of the tensor-core path against the oracle that rounds where the kernels round
    (oracle/gnn_oracle.py emulate="tf32": TF32 operands, fp16 tape with the same power-of-two scales, TF32
    dH, bf16 dD), evaluated in float64.  Yardstick: the SAME emulation evaluated in float32 -- a second
    evaluation of the identical rounded function that differs only by fp32 arithmetic noise, as the kernels
    do.  tests/test_oracle_golden.py::test_tf32_emulation_is_close_in_value_and_chaotic_in_gradient shows why a
    flat 2e-3 cannot hold end to end (rounding-boundary flips amplified by the LeakyReLU kinks); every
    kernel is pinned to <= 2e-3 by tests/test_gpu_tape.py instead.  Bar here, per tensor:
    max(3 x yardstick, 2e-3); logits <= 5e-4 vs the emulation and <= 1e-3 vs the unrounded reference."""
    fin, nf, n_way = 133, 96, 5
    p64 = O.random_params(fin, nf, n_way, seed, torch.float64)
    params = {k: v.float().numpy() for k, v in p64.items()}
    g = torch.Generator().manual_seed(500 + seed)
    x = torch.randn(bsz, n, fin, generator=g)
    proj = torch.randn(bsz, n, n_way, generator=g)

    def emu(dtype):
        p = {k: torch.as_tensor(v).to(dtype) for k, v in params.items()}
        out, dx, gr = O.loss_and_grads(x.to(dtype), p, proj.to(dtype), emulate="tf32")
        return out.double().numpy(), dx.double().numpy(), {k: v.double().numpy() for k, v in gr.items()}

    e64, e32 = emu(torch.float64), emu(torch.float32)
    with torch.no_grad():
        out_true = O.gnn_nl(x.double(), {k: torch.as_tensor(v).double() for k, v in params.items()}).numpy()
    out, dx, grads = U.run_cuda_gnn(x, params, proj, fin, nf, n_way, "tf32", True)
    import mft_b200
    mft_b200.set_precision("auto")
    assert U.rel(out, out_true) < OUT_TOL_TF32
    assert U.rel(out, e64[0]) < 5e-4, U.rel(out, e64[0])
    lim = max(3 * U.rel(e32[1], e64[1]), 2e-3)
    assert U.rel(dx, e64[1]) < lim, ("dx", U.rel(dx, e64[1]), lim)
    for k, want in e64[2].items():
        got = grads[k].reshape(want.shape)
        if U.is_zero_grad(k):
            assert np.abs(got).max() <= 1e-6, k
            continue
        lim = max(3 * U.rel(e32[2][k], want), 2e-3)
        assert U.rel(got, want) < lim, (k, U.rel(got, want), lim)


def test_tf32_and_fp32_paths_agree_on_gradients_direction():
    """The two arithmetic paths share every kernel but the GEMMs; their gradients must point the
    same way (cosine > 0.98 per large tensor) even where slope flips blur the relative error."""
    rec, params = _golden(os.path.join(os.path.dirname(__file__), "golden"), "gnn_5w5s.npz")
    a = U.run_cuda_gnn(rec["x"], params, rec["proj"], 133, 96, 5, "fp32", True)
    b = U.run_cuda_gnn(rec["x"], params, rec["proj"], 133, 96, 5, "tf32", True)
    for k in params:
        if U.is_zero_grad(k) or a[2][k].size < 1000:
            continue
        u, v = a[2][k].ravel(), b[2][k].ravel()
        cos = float(u @ v / (np.linalg.norm(u) * np.linalg.norm(v)))
        assert cos > 0.98, (k, cos)


def test_non_cuda_input_raises():
    import mft_b200
    net = mft_b200.GNN_nl(13, 16, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net(torch.randn(2, 4, 13))
