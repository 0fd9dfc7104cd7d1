"""CPU oracle for the episodic GNN few-shot head -- TEST INFRASTRUCTURE ONLY.

This file is a plain PyTorch-on-CPU restatement of the reference algorithm
(johncai117/Meta-Fine-Tuning, ``methods/gnn.py`` and the ``forward_gnn`` /
``set_forward`` glue of ``methods/gnnnet.py`` / ``methods/gnnnet_copy.py``).
It is the checker, never the product:

* only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
  ``--impl reference`` legs of ``bench.py`` may import it;
* nothing under ``meta-fine-tuning_b200/`` imports it, and the product path
  raises when the CUDA library is missing instead of falling back here.

Parity pin: the reference holds no golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the
reference itself, produced in the build container by
``tests/golden/make_golden.py`` (imports ``/root/reference``) and committed as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every one.

The restatement is functional (parameters travel in a dict keyed by the
reference ``state_dict`` names), works in float32 or float64, and keeps the
pair tensor in ``[B, N, N, C]`` row layout with plain matmuls instead of the
reference's NCHW 1x1 convolutions -- same arithmetic, different summation order.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

BN_EPS = 1e-5          # torch BatchNorm default, reference gnn.py:65-74
LRELU_SLOPE = 0.01     # F.leaky_relu default, reference gnn.py:86
DIAG_MASK = 1e8        # reference gnn.py:106


# When a list is installed here, every BatchNorm output's min |y| is appended to it.  Tests use
# it to find inputs on which no pre-activation sits within rounding distance of the LeakyReLU
# kink: there the gradient is discontinuous, and ANY two float32 evaluations (including two
# of the reference itself) may legitimately pick different slopes for such an element.
KINK_PROBE = None


def _bn_batch(h: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    """Batch-statistic normalisation over every dim but the last (channel) one.

    Reference: every BatchNorm in the head is built with
    ``track_running_stats=False`` (gnn.py:41,65,70,72,74; gnnnet.py:30), so both
    train and eval mode normalise with the biased batch variance, eps 1e-5.
    """
    c = h.shape[-1]
    y = torch.nn.functional.batch_norm(h.reshape(-1, c), None, None, gamma, beta, True, 0.0, BN_EPS).reshape(h.shape)
    if KINK_PROBE is not None:
        KINK_PROBE.append(float(y.detach().abs().min()))
    return y


def min_abs_preactivation(x: torch.Tensor, p: Dict[str, torch.Tensor], emulate=None) -> float:
    """Smallest |BatchNorm output| anywhere in one GNN_nl forward (distance to the nearest
    LeakyReLU kink).  The BatchNorm1d of the last-but-one Gconvs feeds a LeakyReLU as well."""
    global KINK_PROBE
    KINK_PROBE = []
    try:
        with torch.no_grad():
            gnn_nl(x, p, emulate=emulate)
        return min(KINK_PROBE)
    finally:
        KINK_PROBE = None


def _lrelu(h: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.leaky_relu(h, LRELU_SLOPE)


# ----------------------------------------------------------------------------
# Emulation of the tensor-core path's roundings (emulate="tf32")
# ----------------------------------------------------------------------------
# The tcgen05 kernels (meta-fine-tuning_b200/csrc/umma_layers.cu) evaluate the SAME function as the
# reference but round at fixed points.  With ``emulate="tf32"`` this oracle rounds at exactly those
# points (values only -- the arithmetic between them stays in the tensor's dtype, float64 in the
# tests), so that the CUDA path can be compared sharply: LeakyReLU slopes are then chosen from the
# same rounded pre-activations the kernels see.  Rounding points of one Wcompute:
#
#   forward   A operand of every conv GEMM: |x_i - x_j| resp. LeakyReLU(BN(H_{k-1}))  -> TF32 (rna)
#             conv weights: W'_k = s_k * W_k (s_k a power of two, see tape_scale)          -> TF32 (rna)
#             conv bias: not added (BatchNorm cancels it exactly; the kernels never add it)
#             H_k = A W'_k^T, fp32 accumulate                                              -> fp16 (rn, saturating) tape
#             batch statistics from the fp16 tape values, eps' = s_k^2 * 1e-5 (BN(s h; s^2 eps) = BN(h; eps))
#             conv2d_last, softmax, Gconv: fp32 arithmetic on the unrounded LeakyReLU(BN(H_4))
#   backward  dH_k (BatchNorm backward)                                                   -> TF32, operand of dgrad and wgrad
#             dD = dL/d|x_i - x_j|                                                         -> bf16 before the dx gather
#             everything else fp32
# Forward roundings are straight-through in autograd (the kernels' hand-written backward differentiates the
# unrounded function around the rounded values); backward roundings are applied to the gradient itself.

def round_tf32(t: torch.Tensor) -> torch.Tensor:
    """Round to TF32 (10-bit mantissa), nearest, ties away from zero: what ``cvt.rna.tf32.f32`` and
    the producers' integer add of half an ulp do (csrc/umma.cuh to_tf32 / to_tf32_fast)."""
    i = t.detach().to(torch.float32).contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32).to(t.dtype)


def round_fp16_sat(t: torch.Tensor) -> torch.Tensor:
    """``cvt.rn.satfinite.f16x2.f32`` (csrc/common.cuh pack_half4): nearest-even, clamped to +-65504."""
    return t.detach().to(torch.float32).clamp(-65504.0, 65504.0).to(torch.float16).to(t.dtype)


def round_bf16(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).to(torch.bfloat16).to(t.dtype)


_ROUND = {"tf32": round_tf32, "fp16": round_fp16_sat, "bf16": round_bf16}


class _RoundValue(torch.autograd.Function):
    """y = round(x) in the forward, identity in the backward (straight-through)."""

    @staticmethod
    def forward(ctx, x, mode):
        return _ROUND[mode](x)

    @staticmethod
    def backward(ctx, g):
        return g, None


class _RoundGrad(torch.autograd.Function):
    """Identity in the forward; the gradient passing back through it is rounded."""

    @staticmethod
    def forward(ctx, x, mode):
        ctx.mode = mode
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _ROUND[ctx.mode](g), None


def tape_scale(w: torch.Tensor, gamma_prev=None, beta_prev=None) -> float:
    """Power-of-two pre-scale s_k folded into the weight image of conv layer k on the tensor-core path
    (csrc/umma_layers.cu umma_layer_scales_kernel): s_k = 2^-(e_w + e_a) with e_w the binary exponent of
    max|W_k| and e_a that of max(|gamma_{k-1}|, |beta_{k-1}|) (0 for the first layer), exponents as
    frexp gives them.  BatchNorm is invariant to the scale of its input, so only the fp16 tape sees s_k:
    it keeps the tape inside fp16's range whatever the scale of the weights."""
    def expo(v):
        v = float(v)
        if not (v > 0.0) or not math.isfinite(v):
            return 0
        return max(-60, min(60, math.frexp(v)[1]))
    e = expo(w.detach().abs().max())
    if gamma_prev is not None:
        e += expo(max(float(gamma_prev.detach().abs().max()), float(beta_prev.detach().abs().max())))
    return 2.0 ** (-max(-60, min(60, e)))


def edge_scores(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str, emulate=None) -> torch.Tensor:
    """Pairwise edge MLP up to the pre-softmax score S[b,i,j].

    Reference: Wcompute.forward, gnn.py:78-103.  ``x`` is [B,N,F].  ``emulate="tf32"``: round where the
    tensor-core kernels round (see the block comment above).
    """
    d = (x.unsqueeze(2) - x.unsqueeze(1)).abs()              # gnn.py:79-81  [B,N,N,F]
    if emulate is None:
        h = d
        for k in (1, 2, 3, 4):                               # gnn.py:84-100
            w = p[f"{prefix}conv2d_{k}.weight"].flatten(1)   # [out,in,1,1] -> [out,in]
            b = p[f"{prefix}conv2d_{k}.bias"]
            h = h @ w.t() + b
            h = _bn_batch(h, p[f"{prefix}bn_{k}.weight"], p[f"{prefix}bn_{k}.bias"])
            h = _lrelu(h)
        a_last = h
    else:
        assert emulate == "tf32", emulate
        d = _RoundGrad.apply(d, "bf16")
        a = _RoundValue.apply(d, "tf32")
        a_last = None
        for k in (1, 2, 3, 4):
            w = p[f"{prefix}conv2d_{k}.weight"].flatten(1)
            sk = tape_scale(w, p.get(f"{prefix}bn_{k - 1}.weight"), p.get(f"{prefix}bn_{k - 1}.bias")) if k > 1 \
                else tape_scale(w)
            wr = _RoundValue.apply(w * sk, "tf32")
            h = a @ wr.t()                                   # no bias: BatchNorm cancels it, the kernels never add it
            h = _RoundGrad.apply(h, "tf32")
            h = _RoundValue.apply(h, "fp16")
            c = h.shape[-1]
            y = torch.nn.functional.batch_norm(h.reshape(-1, c), None, None, p[f"{prefix}bn_{k}.weight"],
                                               p[f"{prefix}bn_{k}.bias"], True, 0.0, BN_EPS * sk * sk).reshape(h.shape)
            if KINK_PROBE is not None:
                KINK_PROBE.append(float(y.detach().abs().min()))
            a_last = _lrelu(y)
            a = _RoundValue.apply(a_last, "tf32")
    w = p[f"{prefix}conv2d_last.weight"].flatten(1)          # gnn.py:102  [1,nf]
    s = a_last @ w.t() + p[f"{prefix}conv2d_last.bias"]
    return s.squeeze(-1)                                     # [B,N,N]


def edge_adjacency(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str, emulate=None) -> torch.Tensor:
    """Row-stochastic adjacency A[b,i,:] = softmax_j(S[b,i,j] - 1e8*[i==j]).

    Reference: gnn.py:105-115 (activation == 'softmax', the only one used).
    """
    s = edge_scores(x, p, prefix, emulate)
    n = x.shape[1]
    s = s - torch.eye(n, dtype=x.dtype).unsqueeze(0) * DIAG_MASK
    return torch.softmax(s, dim=2)


def wcompute(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str, emulate=None) -> torch.Tensor:
    """Full Wcompute output [B,N,N,2] = stack(identity, A), operator 'J2'.

    Reference: gnn.py:125-132.
    """
    a = edge_adjacency(x, p, prefix, emulate)
    n = x.shape[1]
    eye = torch.eye(n, dtype=x.dtype).unsqueeze(0).expand_as(a)
    return torch.stack([eye, a], dim=3)


def gmul_j2(w: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """[W_0 x , W_1 x] concatenated on the feature axis.  Reference gnn.py:16-28."""
    outs = [torch.bmm(w[..., j], x) for j in range(w.shape[3])]
    return torch.cat(outs, dim=2)


def gconv(w: torch.Tensor, x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str,
          bn_bool: bool = True) -> torch.Tensor:
    """Graph convolution: gmul -> Linear -> (BatchNorm1d).  Reference gnn.py:43-56."""
    y = gmul_j2(w, x)
    bsz, n, _ = y.shape
    y = y.reshape(bsz * n, -1) @ p[f"{prefix}fc.weight"].t() + p[f"{prefix}fc.bias"]
    if bn_bool:
        y = _bn_batch(y, p[f"{prefix}bn.weight"], p[f"{prefix}bn.bias"])
    return y.reshape(bsz, n, -1)


def gnn_nl(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str = "",
           num_layers: int = 2, emulate=None) -> torch.Tensor:
    """GNN_nl.forward: dense-concat stack of (Wcompute, Gconv).  Reference gnn.py:154-166."""
    for i in range(num_layers):
        wi = wcompute(x, p, f"{prefix}layer_w{i}.", emulate)
        x_new = _lrelu(gconv(wi, x, p, f"{prefix}layer_l{i}."))
        x = torch.cat([x, x_new], dim=2)
    wl = wcompute(x, p, f"{prefix}w_comp_last.", emulate)
    return gconv(wl, x, p, f"{prefix}layer_last.", bn_bool=False)


# ----------------------------------------------------------------------------
# GnnNet glue (label indexing, graph assembly, score selection)
# ----------------------------------------------------------------------------

def support_label(n_way: int, n_support: int, dtype=torch.float32) -> torch.Tensor:
    """One-hot class labels for the support nodes, zeros for the query slot.

    Reference: gnnnet.py:35-38 -> [1, n_way*(n_support+1), n_way], class-major.
    """
    lab = torch.zeros(n_way, n_support + 1, n_way, dtype=dtype)
    for c in range(n_way):
        lab[c, :n_support, c] = 1
    return lab.view(1, n_way * (n_support + 1), n_way)


def build_graphs(z: torch.Tensor, n_way: int, n_support: int, n_query: int,
                 compress: bool = False) -> torch.Tensor:
    """Assemble the n_query graphs of one episode, labels attached.

    ``z`` is [n_way, n_support+n_query, D] (post ``fc``).  Graph q holds every
    support embedding plus the q-th query of each class (gnnnet.py:83), then
    the one-hot labels are concatenated (gnnnet.py:212).  ``compress`` follows
    gnnnet_copy.py:34,67-72: supports are averaged in two halves so the graph
    has n_way*(round(n_support/2)+1) nodes.
    """
    d = z.shape[2]
    if compress:
        k = round(n_support / 2)
        sup = z[:, :2 * k].reshape(n_way, 2, k, d).mean(dim=1)
        n_sup_eff, q0 = k, 2 * k
    else:
        sup, n_sup_eff, q0 = z[:, :n_support], n_support, n_support
    lab = support_label(n_way, n_sup_eff, z.dtype)
    graphs = []
    for q in range(n_query):
        g = torch.cat([sup, z[:, q0 + q:q0 + q + 1]], dim=1).reshape(1, -1, d)
        graphs.append(torch.cat([g, lab], dim=2))
    return torch.cat(graphs, dim=0)                           # [n_query, N, D+n_way]


def select_scores(out: torch.Tensor, n_way: int, n_support: int, n_query: int) -> torch.Tensor:
    """Pick the query node of every class and order rows class-major.

    Reference gnnnet.py:216: view(n_query, n_way, n_support+1, n_way)[:, :, -1]
    .permute(1,0,2) -> [n_way*n_query, n_way]; matches
    y = np.repeat(range(n_way), n_query) (gnnnet.py:220).
    """
    o = out.reshape(n_query, n_way, n_support + 1, n_way)[:, :, -1]
    return o.permute(1, 0, 2).reshape(n_way * n_query, n_way)


def query_labels(n_way: int, n_query: int) -> torch.Tensor:
    """gnnnet.py:220 / finetune.py:657."""
    return torch.from_numpy(np.repeat(np.arange(n_way), n_query)).long()


def head_scores(feat: torch.Tensor, p: Dict[str, torch.Tensor], n_way: int, n_support: int,
                n_query: int, compress: bool = False, emulate=None) -> torch.Tensor:
    """``set_forward(x, is_feature=True)``: fc (Linear 512->128 + BN1d) -> graphs -> GNN.

    Reference gnnnet.py:71-87, 210-217 (gnnnet_copy.py:51-78 when ``compress``).
    ``p`` holds ``fc.0.weight/bias``, ``fc.1.weight/bias`` and ``gnn.*``.
    """
    z = feat.reshape(-1, feat.shape[-1]) @ p["fc.0.weight"].t() + p["fc.0.bias"]
    z = _bn_batch(z, p["fc.1.weight"], p["fc.1.bias"])
    z = z.reshape(n_way, -1, z.shape[1])
    nodes = build_graphs(z, n_way, n_support, n_query, compress)
    out = gnn_nl(nodes, p, "gnn.", emulate=emulate)
    n_sup_eff = round(n_support / 2) if compress else n_support
    return select_scores(out, n_way, n_sup_eff, n_query)


def head_loss(feat, p, n_way, n_support, n_query, compress=False, emulate=None) -> torch.Tensor:
    """``set_forward_loss`` on features: cross-entropy of head_scores.  gnnnet.py:219-224."""
    s = head_scores(feat, p, n_way, n_support, n_query, compress, emulate)
    return torch.nn.functional.cross_entropy(s, query_labels(n_way, n_query))


# ----------------------------------------------------------------------------
# Parameter construction (same registration order / shapes as the reference)
# ----------------------------------------------------------------------------

def wcompute_shapes(fin: int, nf: int) -> List[Tuple[str, Tuple[int, ...]]]:
    c = [fin, 2 * nf, 2 * nf, nf, nf]
    out = []
    for k in range(1, 5):
        out += [(f"conv2d_{k}.weight", (c[k], c[k - 1], 1, 1)), (f"conv2d_{k}.bias", (c[k],)),
                (f"bn_{k}.weight", (c[k],)), (f"bn_{k}.bias", (c[k],))]
    out += [("conv2d_last.weight", (1, nf, 1, 1)), ("conv2d_last.bias", (1,))]
    return out


def gnn_nl_shapes(fin: int, nf: int, n_way: int, num_layers: int = 2):
    """(name, shape) for every GNN_nl parameter, in the reference's state_dict order."""
    out = []
    for i in range(num_layers):
        f = fin + (nf // 2) * i
        out += [(f"layer_w{i}.{n}", s) for n, s in wcompute_shapes(f, nf)]
        out += [(f"layer_l{i}.fc.weight", (nf // 2, 2 * f)), (f"layer_l{i}.fc.bias", (nf // 2,)),
                (f"layer_l{i}.bn.weight", (nf // 2,)), (f"layer_l{i}.bn.bias", (nf // 2,))]
    f = fin + (nf // 2) * num_layers
    out += [(f"w_comp_last.{n}", s) for n, s in wcompute_shapes(f, nf)]
    out += [("layer_last.fc.weight", (n_way, 2 * f)), ("layer_last.fc.bias", (n_way,))]
    return out


def random_params(fin: int, nf: int, n_way: int, seed: int, dtype=torch.float64,
                  perturb_bn: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded GNN_nl parameters for tests: PyTorch-default-like scales, with the
    BatchNorm affine terms perturbed away from (1,0) so that a wrong gamma/beta
    wiring cannot hide."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in gnn_nl_shapes(fin, nf, n_way):
        if ".bn" in name and name.endswith("weight"):
            t = 1.0 + (0.25 * torch.randn(shape, generator=g, dtype=torch.float64) if perturb_bn else 0)
            t = torch.as_tensor(t, dtype=torch.float64).expand(shape).clone()
        elif ".bn" in name and name.endswith("bias"):
            t = 0.2 * torch.randn(shape, generator=g, dtype=torch.float64) if perturb_bn \
                else torch.zeros(shape, dtype=torch.float64)
        else:
            fan_in = shape[1] if len(shape) > 1 else None
            if fan_in is None:
                # bias: fan_in of the matching weight is the previous entry's
                fan_in = prev_fan_in
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound
            if len(shape) > 1:
                prev_fan_in = shape[1]
        p[name] = t.to(dtype)
    return p


def loss_and_grads(x: torch.Tensor, p: Dict[str, torch.Tensor], proj: torch.Tensor, emulate=None):
    """Forward GNN_nl, scalar loss = sum(out * proj), gradients for x and every parameter."""
    x = x.detach().clone().requires_grad_(True)
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    out = gnn_nl(x, q, emulate=emulate)
    loss = (out * proj).sum()
    loss.backward()
    # (a parameter the function does not depend on -- the conv biases under emulation -- has gradient zero)
    return out.detach(), x.grad.detach(), {k: (v.grad.detach() if v.grad is not None else torch.zeros_like(v))
                                           for k, v in q.items()}


def flops_per_pair(fin: int, nf: int = 96) -> int:
    """Forward GEMM FLOPs (2*MAC) of one Wcompute per node pair."""
    return 2 * (fin * 2 * nf + 2 * nf * 2 * nf + 2 * nf * nf + nf * nf + nf)


def head_flops(b: int, n: int, fin: int = 133, nf: int = 96, backward: bool = True) -> float:
    """Algorithmic dense FLOPs of the three Wcomputes of one GNN_nl call."""
    per_pair = sum(flops_per_pair(fin + (nf // 2) * i, nf) for i in range(3))
    return per_pair * (3 if backward else 1) * b * n * n
