"""Drop-in module surface of the reference's ``methods/gnn.py`` on hand-written sm_100a kernels.

Same classes, constructor signatures, parameter names / shapes / registration order
(so ``state_dict`` and default-init RNG consumption are identical) and class
attributes as the reference:

* ``gmul``      -- methods/gnn.py:16-28
* ``Gconv``     -- methods/gnn.py:30-56
* ``Wcompute``  -- methods/gnn.py:58-132
* ``GNN_nl``    -- methods/gnn.py:134-166

``forward`` hands raw device pointers to ``libmft_gnn.so`` (C ABI in
``include/mft_gnn.h``) on ``torch.cuda.current_stream()``; PyTorch only owns the
memory and the autograd tape.  There is no CPU path and no library fallback: a
non-CUDA input, a missing ``.so`` or an unsupported configuration raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib

__all__ = ["gmul", "Gconv", "Wcompute", "GNN_nl", "set_precision", "get_precision"]

_PRECISIONS = {"fp32": _lib.PREC_FP32, "tf32": _lib.PREC_TF32}
_precision = os.environ.get("MFT_PRECISION", "auto").lower()


def set_precision(mode: str) -> None:
    """'fp32' (CUDA-core, exact fp32), 'tf32' (tcgen05 tensor cores) or 'auto'
    (tf32 where the tensor-core kernels support the shape, else fp32)."""
    global _precision
    mode = mode.lower()
    if mode not in ("fp32", "tf32", "auto"):
        raise ValueError(f"unknown precision {mode!r}")
    _precision = mode


def get_precision() -> str:
    return _precision


def _resolve_precision(fins: Sequence[int], nf: int) -> int:
    if _precision == "fp32":
        return _lib.PREC_FP32
    lib = _lib.load_library()
    ok = all(lib.mft_tf32_supported(int(f), int(nf)) for f in fins)
    if _precision == "tf32":
        if not ok:
            raise RuntimeError(f"tf32 tensor-core path does not support F={list(fins)}, nf={nf}")
        return _lib.PREC_TF32
    return _lib.PREC_TF32 if ok else _lib.PREC_FP32


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor, got {t.device}; this package has no CPU path")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{what}: expected float32, got {t.dtype}")


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _eff(p: torch.Tensor) -> torch.Tensor:
    """Honour the reference's fast-weight convention (backbone.py *_fw layers): a
    ``.fast`` tensor, when someone assigns one, replaces the parameter."""
    fast = getattr(p, "fast", None)
    return p if fast is None else fast


def _require_identity(W_id: torch.Tensor, what: str) -> None:
    """The kernels hard-wire what GNN_nl always passes (gnn.py:155): operator 0 / the -1e8 mask is the
    identity.  The reference API would accept any W_id, so anything else is refused loudly instead of being
    silently ignored.  One comparison + host read per call of the STANDALONE modules (GNN_nl's fused path
    never materialises W_id); skipped under CUDA-graph capture; MFT_CHECK_IDENTITY=0 switches it off."""
    if os.environ.get("MFT_CHECK_IDENTITY", "1") == "0" or not W_id.is_cuda:
        return
    if torch.cuda.is_current_stream_capturing():
        return
    n = W_id.size(1)
    eye = torch.eye(n, device=W_id.device, dtype=W_id.dtype).view(1, n, n, 1)
    if W_id.dim() != 4 or W_id.size(2) != n or not bool((W_id == eye).all()):
        raise RuntimeError(f"{what} is not the identity operator; only the identity (what GNN_nl builds, "
                           f"gnn.py:155) is implemented")


def _blob(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ---------------------------------------------------------------------------------
# parameter packing
# ---------------------------------------------------------------------------------

def _wc_tensors(m: "Wcompute") -> List[torch.Tensor]:
    """18 tensors in the reference's registration order."""
    out = []
    for k in (1, 2, 3, 4):
        conv, bn = getattr(m, f"conv2d_{k}"), getattr(m, f"bn_{k}")
        out += [_eff(conv.weight), _eff(conv.bias), _eff(bn.weight), _eff(bn.bias)]
    out += [_eff(m.conv2d_last.weight), _eff(m.conv2d_last.bias)]
    return out


def _gc_tensors(m: "Gconv") -> List[torch.Tensor]:
    out = [_eff(m.fc.weight), _eff(m.fc.bias)]
    if m.bn_bool:
        out += [_eff(m.bn.weight), _eff(m.bn.bias)]
    return out


def _fill_wc_params(dst: _lib.WcomputeParams, t: Sequence[torch.Tensor]) -> None:
    for k in range(4):
        dst.conv_w[k] = _ptr(t[4 * k])
        dst.bn_g[k] = _ptr(t[4 * k + 2])
        dst.bn_b[k] = _ptr(t[4 * k + 3])
    dst.last_w = _ptr(t[16])
    dst.last_b = _ptr(t[17])


def _fill_wc_grads(dst: _lib.WcomputeGrads, t: Sequence[torch.Tensor]) -> None:
    for k in range(4):
        dst.conv_w[k] = _ptr(t[4 * k])
        dst.conv_b[k] = _ptr(t[4 * k + 1])
        dst.bn_g[k] = _ptr(t[4 * k + 2])
        dst.bn_b[k] = _ptr(t[4 * k + 3])
    dst.last_w = _ptr(t[16])
    dst.last_b = _ptr(t[17])


def _fill_gc(dst, t: Sequence[torch.Tensor]) -> None:
    dst.fc_w = _ptr(t[0])
    dst.fc_b = _ptr(t[1])
    dst.bn_g = _ptr(t[2]) if len(t) > 2 else 0
    dst.bn_b = _ptr(t[3]) if len(t) > 2 else 0


def _contig(ts: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    return [t if t.is_contiguous() else t.contiguous() for t in ts]


def _alloc_like_flat(ts: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """One allocation, one view per tensor (16-byte aligned starts)."""
    offs, total = [], 0
    for t in ts:
        offs.append(total)
        total += (t.numel() + 3) & ~3
    flat = torch.empty(total, dtype=torch.float32, device=ts[0].device)
    return [flat[o:o + t.numel()].view(t.shape) for o, t in zip(offs, ts)]


def _shared_mask(shared_nodes, x) -> Optional[bytes]:
    """Length-N byte mask for the C ABI (host memory), or None."""
    if shared_nodes is None:
        return None
    if isinstance(shared_nodes, torch.Tensor):
        shared_nodes = shared_nodes.detach().cpu().tolist()
    mask = bytes(1 if v else 0 for v in shared_nodes)
    if len(mask) != x.size(1):
        raise ValueError("shared_nodes has %d entries for %d nodes" % (len(mask), x.size(1)))
    return mask if any(mask) else None


# ---------------------------------------------------------------------------------
# autograd functions
# ---------------------------------------------------------------------------------

class _WcomputeFn(torch.autograd.Function):
    """adj = softmax_j(edge_mlp(|x_i - x_j|) - 1e8 [i==j])  -- mft_wcompute_fwd / _bwd."""

    @staticmethod
    def forward(ctx, x, nf, prec, shared, *params):
        lib = _lib.load_library()
        _require_cuda(x, "Wcompute")
        x = x.contiguous()
        B, N, F = x.shape
        params = _contig(params)
        p = _lib.WcomputeParams()
        _fill_wc_params(p, params)
        adj = torch.empty(B, N, N, dtype=torch.float32, device=x.device)
        saved = _blob(lib.mft_wcompute_saved_bytes_for(B, N, F, nf, prec), x.device)
        ws = _blob(lib.mft_wcompute_workspace_bytes(B, N, F, nf), x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.mft_wcompute_fwd(x.data_ptr(), F, B, N, F, nf, C.byref(p), adj.data_ptr(),
                                            saved.data_ptr(), ws.data_ptr(), prec, shared, _stream()),
                       "mft_wcompute_fwd")
        ctx.save_for_backward(x, adj, saved, *params)
        ctx.meta = (nf, prec, shared)
        return adj

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_adj):
        lib = _lib.load_library()
        x, adj, saved, *params = ctx.saved_tensors
        nf, prec, shared = ctx.meta
        B, N, F = x.shape
        d_adj = d_adj.contiguous()
        p = _lib.WcomputeParams()
        _fill_wc_params(p, params)
        grads = _alloc_like_flat(params)
        g = _lib.WcomputeGrads()
        _fill_wc_grads(g, grads)
        dx = torch.zeros_like(x)
        ws = _blob(lib.mft_wcompute_workspace_bytes(B, N, F, nf), x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.mft_wcompute_bwd(x.data_ptr(), F, B, N, F, nf, C.byref(p), adj.data_ptr(),
                                            d_adj.data_ptr(), dx.data_ptr(), C.byref(g), saved.data_ptr(),
                                            ws.data_ptr(), prec, shared, _stream()), "mft_wcompute_bwd")
        return (dx, None, None, None, *grads)


class _GconvFn(torch.autograd.Function):
    """out = act(BN1d(x Wa^T + adj (x Wb^T) + b))  -- mft_gconv_fwd / _bwd."""

    @staticmethod
    def forward(ctx, adj, x, lrelu, *params):
        lib = _lib.load_library()
        _require_cuda(x, "Gconv")
        _require_cuda(adj, "Gconv")
        x = x.contiguous()
        adj = adj.contiguous()
        B, N, F = x.shape
        params = _contig(params)
        n_out = params[0].shape[0]
        p = _lib.GconvParams()
        _fill_gc(p, params)
        out = torch.empty(B, N, n_out, dtype=torch.float32, device=x.device)
        saved = _blob(lib.mft_gconv_saved_bytes(B, N, F, n_out), x.device)
        ws = _blob(lib.mft_gconv_workspace_bytes(B, N, F, n_out), x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.mft_gconv_fwd(adj.data_ptr(), x.data_ptr(), F, B, N, F, n_out, C.byref(p), int(lrelu),
                                         out.data_ptr(), n_out, saved.data_ptr(), ws.data_ptr(), _stream()),
                       "mft_gconv_fwd")
        ctx.save_for_backward(adj, x, saved, *params)
        ctx.meta = (int(lrelu), n_out)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_out):
        lib = _lib.load_library()
        adj, x, saved, *params = ctx.saved_tensors
        lrelu, n_out = ctx.meta
        B, N, F = x.shape
        d_out = d_out.contiguous()
        p = _lib.GconvParams()
        _fill_gc(p, params)
        grads = _alloc_like_flat(params)
        g = _lib.GconvGrads()
        _fill_gc(g, grads)
        dx = torch.zeros_like(x)
        d_adj = torch.empty_like(adj)
        ws = _blob(lib.mft_gconv_workspace_bytes(B, N, F, n_out), x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.mft_gconv_bwd(adj.data_ptr(), x.data_ptr(), F, B, N, F, n_out, C.byref(p), lrelu,
                                         d_out.data_ptr(), n_out, dx.data_ptr(), d_adj.data_ptr(), C.byref(g),
                                         saved.data_ptr(), ws.data_ptr(), _stream()), "mft_gconv_bwd")
        return (d_adj, dx, None, *grads)


class _GnnFn(torch.autograd.Function):
    """Whole GNN_nl stack in one library call per direction -- mft_gnn_fwd / _bwd."""

    @staticmethod
    def forward(ctx, x, nf, n_way, prec, shared, *params):
        lib = _lib.load_library()
        _require_cuda(x, "GNN_nl")
        x = x.contiguous()
        B, N, F0 = x.shape
        params = _contig(params)
        p = _lib.GnnParams()
        _pack_gnn(p, params, grads=False)
        out = torch.empty(B, N, n_way, dtype=torch.float32, device=x.device)
        saved = _blob(lib.mft_gnn_saved_bytes_for(B, N, F0, nf, n_way, prec), x.device)
        ws = _blob(lib.mft_gnn_workspace_bytes(B, N, F0, nf, n_way), x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.mft_gnn_fwd(x.data_ptr(), B, N, F0, nf, n_way, C.byref(p), out.data_ptr(),
                                       saved.data_ptr(), ws.data_ptr(), prec, shared, _stream()), "mft_gnn_fwd")
        ctx.save_for_backward(saved, *params)
        ctx.meta = (B, N, F0, nf, n_way, prec, shared)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_out):
        lib = _lib.load_library()
        saved, *params = ctx.saved_tensors
        B, N, F0, nf, n_way, prec, shared = ctx.meta
        d_out = d_out.contiguous()
        p = _lib.GnnParams()
        _pack_gnn(p, params, grads=False)
        grads = _alloc_like_flat(params)
        g = _lib.GnnGrads()
        _pack_gnn(g, grads, grads=True)
        dx = torch.empty(B, N, F0, dtype=torch.float32, device=d_out.device)
        ws = _blob(lib.mft_gnn_workspace_bytes(B, N, F0, nf, n_way), d_out.device)
        with torch.cuda.device(d_out.device):
            _lib.check(lib.mft_gnn_bwd(d_out.data_ptr(), B, N, F0, nf, n_way, C.byref(p), dx.data_ptr(),
                                       C.byref(g), saved.data_ptr(), ws.data_ptr(), prec, shared, _stream()),
                       "mft_gnn_bwd")
        return (dx, None, None, None, None, *grads)


def _pack_gnn(dst, t: Sequence[torch.Tensor], grads: bool) -> None:
    """t = [w0(18) l0(4) w1(18) l1(4) w_last(18) l_last(2)] in registration order."""
    pos = 0
    for l in range(_lib.MAX_LAYERS):
        (_fill_wc_grads if grads else _fill_wc_params)(dst.w[l], t[pos:pos + 18])
        pos += 18
        n = 4 if l < _lib.MAX_LAYERS - 1 else 2
        _fill_gc(dst.l[l], t[pos:pos + n])
        pos += n


# ---------------------------------------------------------------------------------
# modules (reference call surface)
# ---------------------------------------------------------------------------------

def _fw_layers():
    """The reference's fast-weight layer classes (backbone.py:26-87,133-213) when its
    ``backbone`` module is importable (i.e. when running inside the reference tree)."""
    try:
        import backbone  # type: ignore
        return backbone.Linear_fw, backbone.Conv2d_fw, backbone.BatchNorm2d_fw, backbone.BatchNorm1d_fw
    except Exception:
        return None


def _mark_fast(mod: nn.Module) -> nn.Module:
    for prm in mod.parameters(recurse=False):
        prm.fast = None
    return mod


def gmul(input):
    """[W_0 x, W_1 x, ...] concatenated on the feature axis (reference gnn.py:16-28).

    Kept for API parity; Gconv does not call it (the kernels never build the
    [B, N, J*F] tensor).  Runs on the device of its inputs through torch.bmm.
    """
    W, x = input
    N = W.size(-2)
    W = torch.cat(W.split(1, 3), 1).squeeze(3)
    out = torch.bmm(W, x).split(N, 1)
    return torch.cat(out, 2)


class Gconv(nn.Module):
    maml = False

    def __init__(self, nf_input, nf_output, J, bn_bool=True):
        super().__init__()
        self.J = J
        self.num_inputs = J * nf_input
        self.num_outputs = nf_output
        fw = _fw_layers() if self.maml else None
        if self.maml and fw is not None:
            self.fc = fw[0](self.num_inputs, self.num_outputs)
        else:
            self.fc = nn.Linear(self.num_inputs, self.num_outputs)
            if self.maml:
                _mark_fast(self.fc)
        self.bn_bool = bn_bool
        if self.bn_bool:
            if self.maml and fw is not None:
                self.bn = fw[3](self.num_outputs, track_running_stats=False)
            else:
                self.bn = nn.BatchNorm1d(self.num_outputs, track_running_stats=False)
                if self.maml:
                    _mark_fast(self.bn)

    def forward(self, input, _lrelu=False):
        """input = [W [B,N,N,2], x [B,N,F]] -> (W, x_new [B,N,nf_output]).

        W[..., 0] is the identity operator GNN_nl always passes (gnn.py:155,128) and is
        not multiplied out; W[..., 1] is the adjacency.
        """
        W, x = input[0], input[1]
        if self.J != 2 or W.size(3) != 2:
            raise NotImplementedError("Gconv: only J=2 (identity + adjacency) is implemented")
        _require_identity(W[..., :1], "Gconv: operator 0")
        adj = W[..., 1]
        out = _GconvFn.apply(adj, x, bool(_lrelu), *_gc_tensors(self))
        return W, out


class Wcompute(nn.Module):
    maml = False

    def __init__(self, input_features, nf, operator='J2', activation='softmax', ratio=[2, 2, 1, 1],
                 num_operators=1, drop=False):
        super().__init__()
        self.num_features = nf
        self.operator = operator
        fw = _fw_layers() if self.maml else None

        def conv(cin, cout):
            if self.maml and fw is not None:
                return fw[1](cin, cout, 1, stride=1)
            m = nn.Conv2d(cin, cout, 1, stride=1)
            return _mark_fast(m) if self.maml else m

        def bn(c):
            if self.maml and fw is not None:
                return fw[2](c, track_running_stats=False)
            m = nn.BatchNorm2d(c, track_running_stats=False)
            return _mark_fast(m) if self.maml else m

        widths = [input_features, int(nf * ratio[0]), int(nf * ratio[1]), nf * ratio[2], nf * ratio[3]]
        self.conv2d_1 = conv(widths[0], widths[1])
        self.bn_1 = bn(widths[1])
        self.drop = drop
        if self.drop:
            self.dropout = nn.Dropout(0.3)
        self.conv2d_2 = conv(widths[1], widths[2])
        self.bn_2 = bn(widths[2])
        self.conv2d_3 = conv(widths[2], widths[3])
        self.bn_3 = bn(widths[3])
        self.conv2d_4 = conv(widths[3], widths[4])
        self.bn_4 = bn(widths[4])
        self.conv2d_last = conv(nf, num_operators)
        self.activation = activation
        self._widths = widths
        self._num_operators = num_operators

    def _check_supported(self):
        nf = self.num_features
        if (self.operator != 'J2' or self.activation != 'softmax' or self.drop or self._num_operators != 1
                or self._widths[1:] != [2 * nf, 2 * nf, nf, nf]):
            raise NotImplementedError(
                "Wcompute: the CUDA path implements the configuration the reference uses everywhere "
                "(operator='J2', activation='softmax', ratio=[2,2,1,1], num_operators=1, drop=False)")

    def adjacency(self, x, shared_nodes=None):
        self._check_supported()
        nf = self.num_features
        prec = _resolve_precision([x.size(2)], nf)
        return _WcomputeFn.apply(x, nf, prec, _shared_mask(shared_nodes, x), *_wc_tensors(self))

    def forward(self, x, W_id, shared_nodes=None):
        """x [B,N,F], W_id [B,N,N,1] (identity) -> [B,N,N,2] = cat(W_id, adjacency).

        The -1e8 mask of gnn.py:106 is applied on the diagonal, i.e. W_id is taken to
        be the identity GNN_nl builds (gnn.py:155).  ``shared_nodes``: see GNN_nl.shared_nodes."""
        _require_identity(W_id, "Wcompute: W_id")
        adj = self.adjacency(x, shared_nodes)
        return torch.cat([W_id, adj.unsqueeze(3)], 3)


class GNN_nl(nn.Module):
    def __init__(self, input_features, nf, train_N_way):
        super().__init__()
        self.input_features = input_features
        self.nf = nf
        self.num_layers = 2
        self.train_N_way = train_N_way

        for i in range(self.num_layers):
            fin = self.input_features + int(nf / 2) * i
            module_w = Wcompute(fin, nf, operator='J2', activation='softmax', ratio=[2, 2, 1, 1])
            module_l = Gconv(fin, int(nf / 2), 2)
            self.add_module('layer_w{}'.format(i), module_w)
            self.add_module('layer_l{}'.format(i), module_l)

        fin = self.input_features + int(self.nf / 2) * self.num_layers
        self.w_comp_last = Wcompute(fin, nf, operator='J2', activation='softmax', ratio=[2, 2, 1, 1])
        self.layer_last = Gconv(fin, train_N_way, 2, bn_bool=False)
        self.fused = True   # one library call per direction; False = module-by-module
        # Optional promise from the caller (new; the reference recomputes everything): a length-N
        # boolean sequence, True where x[b, n, :] is the SAME row in every graph b -- GnnNet's support
        # nodes (gnnnet.py:79-80 copies them into each query's graph).  layer_w0 then evaluates each
        # support-support pair once instead of B times: same adjacency, same parameter gradients;
        # the input gradient of a shared node is delivered SUMMED over the graphs in graph 0's row
        # (zero contribution from those pairs in the other graphs' rows), which is what the backward
        # of the caller's replication adds up anyway.  None (default) = no assumption.
        self.shared_nodes = None
        self.check_shared = False   # debug: verify the promise on every call (device sync)
        # Without a promise the module looks for itself (eager calls only): node n is treated as shared when
        # x[b, n, :] is bitwise the same row in every graph b -- true for exactly the support nodes of the
        # graphs the reference's forward_gnn builds (gnnnet.py:83,212), so the unchanged scripts get the
        # shared-pair speed-up too.  One tiny comparison + a host read per call (the reference's loops
        # synchronise every step anyway, meta_template.py:88); skipped while a CUDA graph is being captured and
        # for leaf inputs that require grad (their .grad is per graph by definition).  MFT_AUTO_SHARE=0 or
        # auto_share = False switches it off.
        self.auto_share = os.environ.get("MFT_AUTO_SHARE", "1") != "0"

    def _all_tensors(self) -> List[torch.Tensor]:
        t: List[torch.Tensor] = []
        for i in range(self.num_layers):
            t += _wc_tensors(self._modules['layer_w{}'.format(i)])
            t += _gc_tensors(self._modules['layer_l{}'.format(i)])
        t += _wc_tensors(self.w_comp_last)
        t += _gc_tensors(self.layer_last)
        return t

    def _shared(self, x):
        if (self.shared_nodes is None and self.auto_share and x.is_cuda and x.size(0) >= 2
                and not (x.requires_grad and x.is_leaf) and not torch.cuda.is_current_stream_capturing()):
            with torch.no_grad():
                same = (x == x[:1]).all(dim=2).all(dim=0)
            same = same.cpu().tolist()
            return bytes(1 if v else 0 for v in same) if sum(same) >= 2 else None
        mask = _shared_mask(self.shared_nodes, x)
        if mask is not None and self.check_shared:
            sel = torch.tensor([bool(v) for v in mask], device=x.device)
            if not torch.equal(x[:, sel], x[:1, sel].expand(x.size(0), -1, -1)):
                raise ValueError("GNN_nl.shared_nodes marks nodes whose features differ between graphs")
        return mask

    def forward(self, x):
        if self.fused and self.nf % 2 == 0:
            fins = [self.input_features + (self.nf // 2) * i for i in range(self.num_layers + 1)]
            prec = _resolve_precision(fins, self.nf)
            return _GnnFn.apply(x, self.nf, self.train_N_way, prec, self._shared(x), *self._all_tensors())
        # module-by-module (same kernels, one autograd node per module)
        W_init = torch.eye(x.size(1), device=x.device).unsqueeze(0).repeat(x.size(0), 1, 1).unsqueeze(3)
        for i in range(self.num_layers):
            Wi = self._modules['layer_w{}'.format(i)](x, W_init, self._shared(x) if i == 0 else None)
            x_new = self._modules['layer_l{}'.format(i)]([Wi, x], _lrelu=True)[1]
            x = torch.cat([x, x_new], 2)
        Wl = self.w_comp_last(x, W_init)
        return self.layer_last([Wl, x])[1]
