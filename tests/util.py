"""Shared helpers of the GPU parity tests."""
import os

import numpy as np
import torch

from oracle import gnn_oracle as O


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))


ZERO_GRAD_SUFFIXES = tuple(f"conv2d_{k}.bias" for k in (1, 2, 3, 4, "last")) + (
    "layer_l0.fc.bias", "layer_l1.fc.bias")


def is_zero_grad(name):
    """Parameters whose gradient is analytically zero."""
    return name.endswith(ZERO_GRAD_SUFFIXES)


def load_params_into(module, params):
    sd = {k: torch.as_tensor(np.asarray(v)).float() for k, v in params.items()}
    module.load_state_dict(sd)
    return module


def run_cuda_gnn(x, params, proj, fin, nf, n_way, precision="fp32", fused=True):
    """Run the CUDA GNN_nl forward + backward; returns float64 numpy (out, dx, grads)."""
    import mft_b200
    mft_b200.set_precision(precision)
    net = mft_b200.GNN_nl(fin, nf, n_way)
    load_params_into(net, params)
    net = net.cuda()
    net.fused = fused
    xg = torch.as_tensor(np.asarray(x)).float().cuda().requires_grad_(True)
    out = net(xg)
    loss = (out * torch.as_tensor(np.asarray(proj)).float().cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: v.grad.detach().double().cpu().numpy() for k, v in net.named_parameters()}
    return out.detach().double().cpu().numpy(), xg.grad.double().cpu().numpy(), grads


def oracle_truth(x, params, proj, dtype=torch.float64):
    p = {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in params.items()}
    out, dx, grads = O.loss_and_grads(torch.as_tensor(np.asarray(x)).to(dtype), p,
                                      torch.as_tensor(np.asarray(proj)).to(dtype))
    return out.double().numpy(), dx.double().numpy(), {k: v.double().numpy() for k, v in grads.items()}


def check_against_truth(got, truth, yard, out_tol, grad_factor, grad_floor, label=""):
    """got/truth/yard = (out, dx, grads); yard is an fp32 evaluation of the reference
    function (its error against truth is the yardstick for gradients)."""
    out, dx, grads = got
    out_t, dx_t, grads_t = truth
    out_y, dx_y, grads_y = yard
    report = {"out": rel(out, out_t), "dx": rel(dx, dx_t)}
    assert report["out"] < out_tol, f"{label} logits rel err {report['out']:.3e} >= {out_tol}"
    lim = max(grad_factor * rel(dx_y, dx_t), grad_floor)
    assert report["dx"] < lim, f"{label} dx rel err {report['dx']:.3e} >= {lim:.3e}"
    for k, gt in grads_t.items():
        g = grads[k].reshape(gt.shape)
        if is_zero_grad(k):
            assert np.abs(g).max() <= 1e-6, f"{label} {k}: analytically-zero gradient is {np.abs(g).max():.3e}"
            continue
        e = rel(g, gt)
        lim = max(grad_factor * rel(grads_y[k], gt), grad_floor)
        report[k] = e
        assert e < lim, f"{label} {k}: rel err {e:.3e} >= {lim:.3e}"
    return report


def kink_free_problem(bsz, n, fin, nf, n_way, seed0, tau=1e-4, max_tries=200):
    """Seeded (params, x, proj) on which every BatchNorm output of the float64 forward is at
    least `tau` away from the LeakyReLU kink, so that float32 and float64 evaluations agree on
    every slope and gradients can be compared sharply.  Deterministic seed search."""
    for t in range(max_tries):
        seed = seed0 + 7919 * t
        p64 = O.random_params(fin, nf, n_way, seed, torch.float64)
        p32 = {k: v.float() for k, v in p64.items()}
        g = torch.Generator().manual_seed(1000 + seed)
        x = torch.randn(bsz, n, fin, generator=g)
        proj = torch.randn(bsz, n, n_way, generator=g)
        m = O.min_abs_preactivation(x.double(), {k: v.double() for k, v in p32.items()})
        if m > tau:
            return {k: v.numpy() for k, v in p32.items()}, x, proj, seed, m
    raise RuntimeError("no kink-free seed found")
