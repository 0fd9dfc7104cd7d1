"""Executable specification of what the CUDA kernels compute (test infrastructure).

The kernels do not evaluate the reference graph literally: they work on the
*unordered* node pairs of each graph (the edge MLP input |x_i - x_j| is symmetric,
so rows (i,j) and (j,i) are identical through every layer), carry a multiplicity
w in {1 (diagonal), 2} into the batch statistics, drop the conv biases that
BatchNorm cancels, and run a hand-derived backward.  This module states that
algorithm in plain torch float64 so that tests can check, on CPU and before any
GPU time is spent, that it is exactly the reference function (test_kernel_model.py
compares it with oracle autograd).  The CUDA code in
``meta-fine-tuning_b200/csrc`` follows this file step by step.
"""
from __future__ import annotations

import torch

EPS = 1e-5
SLOPE = 0.01


def tri_rows(n: int):
    """Unordered pairs (i<=j), row-major in i: r = i*n - i*(i-1)/2 + (j-i)."""
    ii, jj = [], []
    for i in range(n):
        for j in range(i, n):
            ii.append(i)
            jj.append(j)
    return torch.tensor(ii), torch.tensor(jj)


def _lrelu(y):
    return torch.where(y > 0, y, y * SLOPE)


def _dlrelu(y):
    return torch.where(y > 0, torch.ones_like(y), torch.full_like(y, SLOPE))


def pair_rows(bsz: int, n: int, shared=None):
    """Row table of the edge MLP: (graph whose x the row reads, i, j, multiplicity, covers-all-graphs).

    Without ``shared`` every graph contributes its N(N+1)/2 unordered pairs.  With ``shared`` (bool
    [N], True where x[b, n] is the same row in every graph b -- GnnNet's support nodes) a pair of two
    shared nodes gets ONE row that stands for all graphs: it reads graph 0 and carries bsz times
    the multiplicity.  The shared rows come first, as in the CUDA pair table (common.cuh)."""
    ii, jj = tri_rows(n)
    mult = torch.where(ii == jj, 1.0, 2.0)
    if shared is None or bsz < 2:
        rg = ii.numel()
        return (torch.arange(bsz).repeat_interleave(rg), ii.repeat(bsz), jj.repeat(bsz), mult.repeat(bsz),
                torch.zeros(bsz * rg, dtype=torch.bool))
    shared = torch.as_tensor(shared, dtype=torch.bool)
    m = shared[ii] & shared[jj]
    rs, rq = int(m.sum()), int((~m).sum())
    rb = torch.cat([torch.zeros(rs, dtype=torch.long), torch.arange(bsz).repeat_interleave(rq)])
    ri = torch.cat([ii[m], ii[~m].repeat(bsz)])
    rj = torch.cat([jj[m], jj[~m].repeat(bsz)])
    w = torch.cat([mult[m] * bsz, mult[~m].repeat(bsz)])
    allg = torch.cat([torch.ones(rs, dtype=torch.bool), torch.zeros(bsz * rq, dtype=torch.bool)])
    return rb, ri, rj, w, allg


def wcompute_fwd(x, p, prefix, shared=None):
    """x [B,N,F] -> adjacency A [B,N,N] plus everything the backward needs."""
    bsz, n, f = x.shape
    rb, ri, rj, w, allg = pair_rows(bsz, n, shared)
    w = w.to(x.dtype)
    pairs = float(bsz * n * n)
    d = (x[rb, ri] - x[rb, rj]).abs()                                      # [R,F]
    saved = {"d": d, "w": w, "rows": (rb, ri, rj, allg), "h": [], "mean": [], "rstd": [], "a": [d]}
    a = d
    for k in (1, 2, 3, 4):
        wk = p[f"{prefix}conv2d_{k}.weight"].flatten(1)
        h = a @ wk.t()                                                     # bias dropped (BN cancels it)
        s1 = (w[:, None] * h).sum(0)
        s2 = (w[:, None] * h * h).sum(0)
        mean = s1 / pairs
        var = s2 / pairs - mean * mean
        rstd = 1.0 / torch.sqrt(var + EPS)
        y = (h - mean) * rstd * p[f"{prefix}bn_{k}.weight"] + p[f"{prefix}bn_{k}.bias"]
        a = _lrelu(y)
        saved["h"].append(h)
        saved["mean"].append(mean)
        saved["rstd"].append(rstd)
        saved["a"].append(a)
    wl = p[f"{prefix}conv2d_last.weight"].flatten()
    s = a @ wl + p[f"{prefix}conv2d_last.bias"]                             # [R]
    smat = torch.zeros(bsz, n, n, dtype=x.dtype)
    one = ~allg
    smat[rb[one], ri[one], rj[one]] = s[one]
    smat[rb[one], rj[one], ri[one]] = s[one]
    smat[:, ri[allg], rj[allg]] = s[allg]                                  # a shared pair scores in every graph
    smat[:, rj[allg], ri[allg]] = s[allg]
    smat = smat - torch.eye(n, dtype=x.dtype) * 1e8
    adj = torch.softmax(smat, dim=2)
    saved["adj"] = adj
    return adj, saved


class NoRounding:
    """Operand roundings of the backward GEMMs (identity here; tests/test_gpu_tape.py passes the
    tensor-core path's: dH, a and W to TF32, dD to bf16)."""
    dh = a = w = dD = staticmethod(lambda t: t)


def wcompute_bwd(x, p, prefix, saved, d_adj, rounding=NoRounding, perturb=None):
    """Closed-form backward of wcompute_fwd.  Returns dx and parameter grads.  With shared rows the
    gradient of a shared pair arrives summed over the graphs and is delivered to graph 0's nodes.
    ``perturb``: optional callable (k, dh) -> dh, used by the tests to prove that a small error in the
    BatchNorm-backward term would be caught."""
    bsz, n, f = x.shape
    w = saved["w"]
    rb, ri, rj, allg = saved["rows"]
    pairs = float(bsz * n * n)
    adj = saved["adj"]
    ds = adj * (d_adj - (adj * d_adj).sum(2, keepdim=True))                # softmax backward per row
    twin = ds + ds.transpose(1, 2)                                         # twin gradients summed
    twin = twin - torch.diag_embed(torch.diagonal(ds, dim1=1, dim2=2))     # diagonal counted once (is 0)
    g = twin[rb, ri, rj]
    g = torch.where(allg, twin.sum(0)[ri, rj], g)                          # shared rows: summed over the graphs
    grads = {}
    wl = p[f"{prefix}conv2d_last.weight"].flatten()
    grads[f"{prefix}conv2d_last.weight"] = (g[:, None] * saved["a"][4]).sum(0).reshape(1, -1, 1, 1)
    grads[f"{prefix}conv2d_last.bias"] = torch.zeros(1, dtype=x.dtype)      # sum of dS is 0 (shift invariance)
    da = g[:, None] * wl[None, :]
    for k in (4, 3, 2, 1):
        gamma = p[f"{prefix}bn_{k}.weight"]
        h, mean, rstd = saved["h"][k - 1], saved["mean"][k - 1], saved["rstd"][k - 1]
        hhat = (h - mean) * rstd
        y = hhat * gamma + p[f"{prefix}bn_{k}.bias"]
        dy = da * _dlrelu(y)
        sum_dy = dy.sum(0)
        sum_dyh = (dy * hhat).sum(0)
        grads[f"{prefix}bn_{k}.weight"] = sum_dyh
        grads[f"{prefix}bn_{k}.bias"] = sum_dy
        m1, m2 = sum_dy / pairs, sum_dyh / pairs
        dh = gamma * rstd * (dy - w[:, None] * m1 - w[:, None] * hhat * m2)
        if perturb is not None:
            dh = perturb(k, dh)
        dh = rounding.dh(dh)
        wk = p[f"{prefix}conv2d_{k}.weight"].flatten(1)
        grads[f"{prefix}conv2d_{k}.weight"] = (dh.t() @ rounding.a(saved["a"][k - 1])).reshape(*wk.shape, 1, 1)
        grads[f"{prefix}conv2d_{k}.bias"] = torch.zeros(wk.shape[0], dtype=x.dtype)   # BN removes the mean
        da = dh @ rounding.w(wk)
    da = rounding.dD(da)
    contrib = torch.sign(x[rb, ri] - x[rb, rj]) * da
    dx = torch.zeros(bsz * n, f, dtype=x.dtype)
    dx.index_add_(0, rb * n + ri, contrib)
    dx.index_add_(0, rb * n + rj, -contrib)
    return dx.reshape(bsz, n, f), grads


def gconv_fwd(adj, x, p, prefix, bn_bool=True, lrelu=False):
    """Y = x Wa^T + A (x Wb^T) + b, then BN1d over the B*N rows, optional LeakyReLU."""
    bsz, n, f = x.shape
    wfc = p[f"{prefix}fc.weight"]
    wa, wb = wfc[:, :f], wfc[:, f:]
    y = x @ wa.t() + torch.bmm(adj, x @ wb.t()) + p[f"{prefix}fc.bias"]
    saved = {"y": y}
    if bn_bool:
        rows = float(bsz * n)
        flat = y.reshape(-1, y.shape[-1])
        mean = flat.sum(0) / rows
        var = (flat * flat).sum(0) / rows - mean * mean
        rstd = 1.0 / torch.sqrt(var + EPS)
        saved.update(mean=mean, rstd=rstd)
        z = (y - mean) * rstd * p[f"{prefix}bn.weight"] + p[f"{prefix}bn.bias"]
    else:
        z = y
    saved["z"] = z
    return (_lrelu(z) if lrelu else z), saved


def gconv_bwd(adj, x, p, prefix, saved, d_out, bn_bool=True, lrelu=False):
    bsz, n, f = x.shape
    grads = {}
    dz = d_out * _dlrelu(saved["z"]) if lrelu else d_out
    if bn_bool:
        rows = float(bsz * n)
        gamma = p[f"{prefix}bn.weight"]
        hhat = (saved["y"] - saved["mean"]) * saved["rstd"]
        sum_dz = dz.sum((0, 1))
        sum_dzh = (dz * hhat).sum((0, 1))
        grads[f"{prefix}bn.weight"] = sum_dzh
        grads[f"{prefix}bn.bias"] = sum_dz
        dy = gamma * saved["rstd"] * (dz - sum_dz / rows - hhat * sum_dzh / rows)
    else:
        dy = dz
    wfc = p[f"{prefix}fc.weight"]
    wa, wb = wfc[:, :f], wfc[:, f:]
    ax = torch.bmm(adj, x)
    dyf = dy.reshape(-1, dy.shape[-1])
    grads[f"{prefix}fc.weight"] = torch.cat([dyf.t() @ x.reshape(-1, f), dyf.t() @ ax.reshape(-1, f)], dim=1)
    grads[f"{prefix}fc.bias"] = dyf.sum(0)
    du2 = dy @ wb                                                          # [B,N,F]
    dx = dy @ wa + torch.bmm(adj.transpose(1, 2), du2)
    d_adj = torch.bmm(du2, x.transpose(1, 2))
    return dx, d_adj, grads


def gnn_nl_fwd_bwd(x, p, d_out, nf2, num_layers=2, shared=None):
    """Whole GNN_nl forward + backward with the dense-concat bookkeeping the
    kernels use: one wide buffer xcat, later layers' dx accumulate into its
    leading columns."""
    bsz, n, f0 = x.shape
    ftot = f0 + nf2 * num_layers
    xcat = torch.zeros(bsz, n, ftot, dtype=x.dtype)
    xcat[:, :, :f0] = x
    tape = []
    f = f0
    for i in range(num_layers):
        xin = xcat[:, :, :f].clone()
        adj, sw = wcompute_fwd(xin, p, f"layer_w{i}.", shared if i == 0 else None)   # later layers: per-graph x
        xn, sg = gconv_fwd(adj, xin, p, f"layer_l{i}.", True, True)
        xcat[:, :, f:f + nf2] = xn
        tape.append((f, sw, sg, adj))
        f += nf2
    xin = xcat.clone()
    adj, sw = wcompute_fwd(xin, p, "w_comp_last.")
    out, sg = gconv_fwd(adj, xin, p, "layer_last.", False, False)
    # ---- backward
    grads = {}
    dxcat = torch.zeros_like(xcat)
    dx, d_adj, g = gconv_bwd(adj, xin, p, "layer_last.", sg, d_out, False, False)
    grads.update(g)
    dxcat += dx
    dx, g = wcompute_bwd(xin, p, "w_comp_last.", sw, d_adj)
    grads.update(g)
    dxcat += dx
    for i in reversed(range(num_layers)):
        f, sw, sg, adj = tape[i]
        xin = xcat[:, :, :f].clone()
        dx, d_adj, g = gconv_bwd(adj, xin, p, f"layer_l{i}.", sg, dxcat[:, :, f:f + nf2].clone(), True, True)
        grads.update(g)
        dxcat[:, :, :f] += dx
        dx, g = wcompute_bwd(xin, p, f"layer_w{i}.", sw, d_adj)
        grads.update(g)
        dxcat[:, :, :f] += dx
    return out, dxcat[:, :, :f0].clone(), grads
