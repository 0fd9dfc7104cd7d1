"""Sharp parity of the tensor-core (TF32 / fp16-tape) path, kernel by kernel, at the real head sizes.

Why not simply "gradients vs the oracle <= 2e-3": the TF32 path rounds activations (TF32 operands, fp16
tape), and rounding makes the function DISCONTINUOUS on top of the LeakyReLU kinks.  The emulating
oracle (oracle/gnn_oracle.py, emulate="tf32") evaluated in float32 instead of float64 already moves its
own gradients by 2-3e-2 at 5w5s (tests/test_oracle_golden.py records the experiment): ~10^3 of the 10^7
activations sit within fp32 noise of a rounding boundary, each flip moves that activation by 1e-3
relative, and the LeakyReLU kinks downstream amplify it.  No implementation with a different fp32
summation order can match ANY reference below that level end to end.  What can be pinned sharply --
and is, here -- is every kernel given the kernels' own upstream values ("teacher forcing"):

  forward   layer k's tape H'_k, re-derived in float64 from the kernel's H'_{k-1} (BatchNorm from the
            tape's statistics, LeakyReLU, TF32 rounding, x W'^T, fp16 rounding), must equal the
            kernel's H'_k up to fp16 rounding-boundary noise (rel-L2 <= 2e-4, measured 1-3e-5; a 0.1 % systematic
            error fails); batch statistics <= 1e-5; adjacency <= 2e-5.
  backward  given the kernel's tape and adjacency the backward is a LINEAR map of the upstream gradient
            (no discontinuity left), so the float64 closed form (tests/kernel_model.py, with the
            kernels' operand roundings) must match every gradient tensor to <= 2e-3 -- measured <= 3e-4 (dx, whose
            dD is bf16), <= 1e-4 on every parameter gradient, median 2e-5 (profiles/r02/tf32_teacher_forced.txt).
            test_teacher_forced_check_is_sensitive proves that a 1 % error in the fused
            BatchNorm-backward term (DhInPlaceT / DhT in csrc/umma_layers.cu) would be caught.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import gnn_oracle as O
from tests import kernel_model as KM
from tests import util as U

pytestmark = pytest.mark.gpu

NF = 96


class Tf32Rounding:
    dh = a = w = staticmethod(O.round_tf32)
    dD = staticmethod(O.round_bf16)


def _episode_mask(n_way, n_support):
    return np.array(([True] * n_support + [False]) * n_way)


def _problem(bsz, n, fin, seed, mask=None, wscale=1.0):
    """Seeded Wcompute parameters (BN affine terms off (1, 0)), nodes and an upstream gradient."""
    g = torch.Generator().manual_seed(seed)
    full = O.random_params(fin, NF, 5, seed, torch.float64)
    p = {k[len("layer_w0."):]: v.float() for k, v in full.items() if k.startswith("layer_w0.")}
    for k in (1, 2, 3, 4):
        p[f"conv2d_{k}.weight"] = p[f"conv2d_{k}.weight"] * wscale
    x = torch.randn(bsz, n, fin, generator=g)
    if mask is not None:
        sel = torch.from_numpy(mask)
        x[:, sel] = x[0, sel]
    d_adj = torch.randn(bsz, n, n, generator=g)
    return p, x, d_adj


def _run_kernels(p, x, d_adj, mask):
    """Forward + backward of one Wcompute on the GPU (tf32); returns everything the checks need."""
    import mft_b200
    from mft_b200 import _lib
    lib = _lib.load_library()
    bsz, n, fin = x.shape
    mft_b200.set_precision("tf32")
    m = mft_b200.Wcompute(fin, NF)
    m.load_state_dict({k: v.clone() for k, v in p.items()})
    m = m.cuda()
    xg = x.cuda().requires_grad_(True)
    shared = None if mask is None else [bool(v) for v in mask]
    adj = m.adjacency(xg, shared)
    saved = adj.grad_fn.saved_tensors[2]
    off = (C.c_size_t * 8)()
    _lib.check(lib.mft_debug_wcompute_saved_offsets(bsz, n, fin, NF, _lib.PREC_TF32, off), "offsets")
    rb, ri, rj, w, allg = KM.pair_rows(bsz, n, None if mask is None else torch.from_numpy(mask))
    rows = rb.numel()
    widths = [2 * NF, 2 * NF, NF, NF]
    tape = []
    for k in range(4):
        nb = rows * widths[k] * 2
        tape.append(saved[off[k]:off[k] + nb].view(torch.float16).view(rows, widths[k]).double().cpu())
    slot, copy = int(off[6]), int(off[7])
    fs = saved[off[4]:off[4] + 4 * slot * 8].view(torch.float64).view(4, slot // copy, copy).sum(1).cpu()
    scales = saved[off[5]:off[5] + 16].view(torch.float32).double().cpu()
    adj.backward(d_adj.cuda())
    torch.cuda.synchronize()
    mft_b200.set_precision("auto")
    grads = {k: v.grad.double().cpu() for k, v in m.named_parameters()}
    return {"adj": adj.detach().double().cpu(), "tape": tape, "fsums": fs, "scales": scales,
            "dx": xg.grad.double().cpu(), "grads": grads, "rows": (rb, ri, rj, w.double(), allg)}


def _teacher_forced(p, x, d_adj, k_out, perturb=None):
    """float64 re-derivation of every stage from the kernels' own upstream values."""
    bsz, n, fin = x.shape
    rb, ri, rj, w, allg = k_out["rows"]
    pairs = float(bsz * n * n)
    x64 = x.double()
    p64 = {k: v.double() for k, v in p.items()}
    s = k_out["scales"]
    rep = {}
    # the scales themselves: exact powers of two, equal to the oracle's rule
    for k in (1, 2, 3, 4):
        want = O.tape_scale(p64[f"conv2d_{k}.weight"], p64.get(f"bn_{k - 1}.weight"), p64.get(f"bn_{k - 1}.bias")) \
            if k > 1 else O.tape_scale(p64["conv2d_1.weight"])
        assert float(s[k - 1]) == want, (k, float(s[k - 1]), want)
    d = (x64[rb, ri] - x64[rb, rj]).abs()
    a_prev = d
    saved = {"w": w, "rows": (rb, ri, rj, allg), "h": [], "mean": [], "rstd": [], "a": [d], "adj": k_out["adj"]}
    pw = dict(p64)
    for k in (1, 2, 3, 4):
        sk = float(s[k - 1])
        wk = O.round_tf32(p64[f"conv2d_{k}.weight"].flatten(1) * sk)
        pw[f"conv2d_{k}.weight"] = wk.reshape(*wk.shape, 1, 1)
        h_want = O.round_fp16_sat(O.round_tf32(a_prev) @ wk.t())
        h = k_out["tape"][k - 1]
        rep[f"H{k}"] = U.rel(h.numpy(), h_want.numpy())
        # statistics of the kernels' tape (second moment: minus the eps correction the slot carries)
        c = h.shape[1]
        s1 = (w[:, None] * h).sum(0)
        s2 = (w[:, None] * h * h).sum(0)
        corr = pairs * (sk * sk - 1.0) * float(np.float32(O.BN_EPS))     # the kernels hold eps as a float constant
        rep[f"sum{k}"] = U.rel(k_out["fsums"][k - 1, :c].numpy(), s1.numpy())
        rep[f"sumsq{k}"] = U.rel((k_out["fsums"][k - 1, c:2 * c] - corr).numpy(), s2.numpy())
        mean = s1 / pairs
        var = s2 / pairs - mean * mean
        rstd = 1.0 / torch.sqrt(var + O.BN_EPS * sk * sk)
        y = (h - mean) * rstd * p64[f"bn_{k}.weight"] + p64[f"bn_{k}.bias"]
        a_prev = torch.where(y > 0, y, y * O.LRELU_SLOPE)
        saved["h"].append(h)
        saved["mean"].append(mean)
        saved["rstd"].append(rstd)
        saved["a"].append(a_prev)
    # scores / adjacency from the kernels' H_4
    sc = a_prev @ p64["conv2d_last.weight"].flatten() + p64["conv2d_last.bias"]
    smat = torch.zeros(bsz, n, n, dtype=torch.float64)
    one = ~allg
    smat[rb[one], ri[one], rj[one]] = sc[one]
    smat[rb[one], rj[one], ri[one]] = sc[one]
    smat[:, ri[allg], rj[allg]] = sc[allg]
    smat[:, rj[allg], ri[allg]] = sc[allg]
    adj_want = torch.softmax(smat - torch.eye(n, dtype=torch.float64) * O.DIAG_MASK, dim=2)
    rep["adj"] = float((k_out["adj"] - adj_want).abs().max())
    # backward: closed form on the kernels' tape, with the kernels' operand roundings
    dx, g = KM.wcompute_bwd(x64, pw, "", saved, d_adj.double(), rounding=Tf32Rounding, perturb=perturb)
    for k in (1, 2, 3, 4):
        g[f"conv2d_{k}.weight"] = g[f"conv2d_{k}.weight"] * float(s[k - 1])      # dL/dW = s dL/dW'
    return rep, dx, g


def _fold(dx, mask):
    if mask is None:
        return [dx.numpy()]
    sel = torch.from_numpy(mask)
    return [dx[:, ~sel].numpy(), dx[:, sel].sum(0).numpy()]


def _compare(k_out, dx, g, mask):
    errs = {}
    for i, (a, b) in enumerate(zip(_fold(k_out["dx"], mask), _fold(dx, mask))):
        errs[f"dx{i}"] = U.rel(a, b)
    for name, want in g.items():
        got = k_out["grads"][name].reshape(want.shape)
        if name.endswith("bias") and "conv2d" in name:
            assert float(got.abs().max()) <= 1e-6, name
        else:
            errs[name] = U.rel(got.numpy(), want.numpy())
    return errs


CASES = [
    # bsz, n_way, n_support, fin, share
    (16, 5, 5, 133, True),       # 5w5s, layer_w0 as GnnHead runs it (shared support pairs)
    (16, 5, 5, 133, False),
    (16, 5, 20, 133, True),      # 5w20s, the benchmarked configuration
    (16, 5, 20, 181, False),     # layer_w1 width
    (16, 5, 20, 229, False),     # w_comp_last width (two N passes in dgrad layer 1)
    (6, 5, 25, 229, False),      # compressed 50-shot node count (N = 130)
]


@pytest.mark.parametrize("bsz,n_way,n_support,fin,share", CASES)
def test_tf32_kernels_teacher_forced(bsz, n_way, n_support, fin, share):
    mask = _episode_mask(n_way, n_support) if share else None
    n = n_way * (n_support + 1)
    p, x, d_adj = _problem(bsz, n, fin, 100 + fin + n, mask)
    k_out = _run_kernels(p, x, d_adj, mask)
    rep, dx, g = _teacher_forced(p, x, d_adj, k_out)
    for k in (1, 2, 3, 4):
        assert rep[f"H{k}"] < 2e-4, rep
        assert rep[f"sum{k}"] < 1e-5 and rep[f"sumsq{k}"] < 1e-5, rep
    assert rep["adj"] < 2e-5, rep
    errs = _compare(k_out, dx, g, mask)
    _report(f"B{bsz} N{n} F{fin} shared={share}", rep, errs)
    bad = {k: v for k, v in errs.items() if not v < 2e-3}
    assert not bad, (bad, errs)


def _report(label, rep, errs):
    """Append the measured errors to gpurun_out/tf32_teacher_forced.txt when MFT_PARITY_REPORT is set
    (tools copy it to profiles/)."""
    import os
    path = os.environ.get("MFT_PARITY_REPORT")
    if not path:
        return
    with open(path, "a") as f:
        f.write(f"{label}\n  forward : " + "  ".join(f"{k}={v:.1e}" for k, v in rep.items()) + "\n")
        worst = sorted(errs.items(), key=lambda kv: -kv[1])
        f.write("  backward: worst " + "  ".join(f"{k}={v:.1e}" for k, v in worst[:4]) +
                f"   median {float(np.median(list(errs.values()))):.1e}\n")


def test_teacher_forced_check_is_sensitive():
    """A 1 % error in the BatchNorm-backward term of ONE layer (what a wrong constant in DhInPlaceT / DhT
    would produce) pushes the comparison far over the 2e-3 bar; so does a 1 % error in H."""
    n_way, n_support, fin, bsz = 5, 5, 133, 16
    mask = _episode_mask(n_way, n_support)
    n = n_way * (n_support + 1)
    p, x, d_adj = _problem(bsz, n, fin, 7, mask)
    k_out = _run_kernels(p, x, d_adj, mask)
    for layer in (1, 2, 3, 4):
        _, dx, g = _teacher_forced(p, x, d_adj, k_out, perturb=lambda k, dh: dh * 1.01 if k == layer else dh)
        errs = _compare(k_out, dx, g, mask)
        assert errs[f"conv2d_{layer}.weight"] > 5e-3, (layer, errs)
    wrong = dict(k_out)
    wrong["tape"] = [t.clone() for t in k_out["tape"]]
    wrong["tape"][1] = wrong["tape"][1] * 1.01
    rep, _, _ = _teacher_forced(p, x, d_adj, wrong)
    assert rep["H2"] > 5e-3


@pytest.mark.parametrize("wscale", [1e4, 1e-3, 3e-7])
def test_fp16_tape_is_invariant_to_the_weight_scale(wscale):
    """BatchNorm makes the reference invariant to the scale of conv2d_k.weight; the fp16 tape must be as
    well (power-of-two tape scales, csrc/umma_layers.cu umma_layer_scales_kernel).  Logits of the whole
    head vs the float64 oracle <= 1e-3 with every conv weight multiplied by 1e4 / 1e-3 / 3e-7 (at 3e-7
    the BatchNorm eps is no longer negligible against the variance: the eps handling is exercised), and
    every gradient finite."""
    import mft_b200
    fin, n_way, bsz, n = 133, 5, 16, 30
    p64 = O.random_params(fin, NF, n_way, 21, torch.float64)
    for k in list(p64):
        if "conv2d_" in k and k.endswith("weight") and "last" not in k:
            p64[k] = (p64[k].float() * wscale).double()
    params = {k: v.float().numpy() for k, v in p64.items()}
    g = torch.Generator().manual_seed(5)
    x = torch.randn(bsz, n, fin, generator=g)
    proj = torch.randn(bsz, n, n_way, generator=g)
    with torch.no_grad():
        out_t = O.gnn_nl(x.double(), {k: torch.as_tensor(v).double() for k, v in params.items()}).numpy()
    out, dx, grads = U.run_cuda_gnn(x, params, proj, fin, NF, n_way, "tf32", True)
    mft_b200.set_precision("auto")
    assert U.rel(out, out_t) < 1e-3, U.rel(out, out_t)
    assert np.isfinite(dx).all() and all(np.isfinite(v).all() for v in grads.values())
    # the conv-weight gradients scale inversely with the weights: compare with the oracle's at 0.15 (the
    # end-to-end bar of test_gpu_parity.py; sharp parity is the teacher-forced test above, which also runs
    # at a non-unit scale below)
    if wscale < 1e-5:
        return      # variance << eps: every pre-activation sits at beta, the gradient is ~1e-15 of rounding noise
    _, _, g_t = U.oracle_truth(x, params, proj, torch.float64)
    for k in ("layer_w0.conv2d_2.weight", "w_comp_last.conv2d_1.weight"):
        assert U.rel(grads[k].reshape(g_t[k].shape), g_t[k]) < 0.15, k


@pytest.mark.parametrize("wscale", [1e4, 1e-3])
def test_tf32_kernels_teacher_forced_at_other_weight_scales(wscale):
    n_way, n_support, fin, bsz = 5, 5, 133, 16
    n = n_way * (n_support + 1)
    p, x, d_adj = _problem(bsz, n, fin, 9, None, wscale)
    k_out = _run_kernels(p, x, d_adj, None)
    assert float(k_out["scales"][0]) != 1.0
    rep, dx, g = _teacher_forced(p, x, d_adj, k_out)
    for k in (1, 2, 3, 4):
        assert rep[f"H{k}"] < 2e-4 and rep[f"sum{k}"] < 1e-5 and rep[f"sumsq{k}"] < 1e-5, rep
    errs = _compare(k_out, dx, g, None)
    bad = {k: v for k, v in errs.items() if not v < 2e-3}
    assert not bad, (bad, errs)
