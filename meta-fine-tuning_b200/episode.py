"""Host-side mirror of the GnnNet glue around ``GNN_nl`` (reference methods/gnnnet.py,
methods/gnnnet_copy.py): label indexing, graph assembly, score selection, and a
backbone-less head module with the reference's parameter names (``fc.*``, ``gnn.*``)
so a reference checkpoint's head loads unchanged.

Pure tensor bookkeeping on whatever device the inputs live on; the arithmetic of the
head is in ``GNN_nl`` (CUDA kernels).  The reference's own ``GnnNet`` / ``DampNet``
classes run unchanged on top of ``methods.gnn`` -- this mirror exists so that tests,
``bench.py`` and ``smoke()`` can drive the same path without the reference tree.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

import ctypes as C

from . import _lib
from .gnn import GNN_nl, _alloc_like_flat, _blob, _contig, _fill_gc, _require_cuda, _stream

__all__ = ["support_label", "build_graphs", "select_scores", "query_labels", "GnnHead"]


def support_label(n_way: int, n_support: int) -> torch.Tensor:
    """[1, n_way*(n_support+1), n_way]: one-hot class of every support node, zeros for the
    query slot that closes each class block (gnnnet.py:35-38)."""
    lab = torch.from_numpy(np.repeat(range(n_way), n_support)).unsqueeze(1)
    lab = torch.zeros(n_way * n_support, n_way).scatter(1, lab, 1).view(n_way, n_support, n_way)
    lab = torch.cat([lab, torch.zeros(n_way, 1, n_way)], dim=1)
    return lab.view(1, -1, n_way)


def query_labels(n_way: int, n_query: int) -> torch.Tensor:
    """gnnnet.py:220 / finetune.py:657: class-major query labels."""
    return torch.from_numpy(np.repeat(range(n_way), n_query))


def build_graphs(z: torch.Tensor, label: torch.Tensor, n_way: int, n_support: int, n_query: int,
                 compress: bool = False) -> torch.Tensor:
    """z [n_way, n_support+n_query, D] -> nodes [n_query, n_way*(n_s'+1), D+n_way].

    Graph q = every support embedding + the q-th query of each class (gnnnet.py:83),
    labels concatenated (gnnnet.py:212).  ``compress`` averages the supports in two halves
    first (gnnnet_copy.py:34,67-72); ``label`` must then be built for round(n_support/2)."""
    d = z.size(2)
    if compress:
        k = round(n_support / 2)
        sup = z[:, :2 * k].view(n_way, 2, k, d).mean(dim=1)
        q0 = 2 * k
    else:
        k = n_support
        sup = z[:, :n_support]
        q0 = n_support
    sup = sup.unsqueeze(0).expand(n_query, n_way, k, d)
    qry = z[:, q0:q0 + n_query].permute(1, 0, 2).unsqueeze(2)            # [n_query, n_way, 1, D]
    nodes = torch.cat([sup, qry], dim=2).reshape(n_query, n_way * (k + 1), d)
    lab = label.to(device=z.device, dtype=z.dtype).expand(n_query, -1, -1)
    return torch.cat([nodes, lab], dim=2)


def select_scores(out: torch.Tensor, n_way: int, n_support: int, n_query: int) -> torch.Tensor:
    """Query node of every class, rows class-major -> [n_way*n_query, n_way] (gnnnet.py:216)."""
    return out.view(n_query, n_way, n_support + 1, n_way)[:, :, -1].permute(1, 0, 2).contiguous().view(-1, n_way)


class _HeadFn(torch.autograd.Function):
    """feat [n_way, n_support+n_query, feat_dim] -> nodes [n_query, n_way*(n_support+1), D+n_way]:
    fc (Linear + BatchNorm1d, batch statistics) + graph assembly + label one-hots in one library call
    per direction (mft_head_fwd / _bwd; gnnnet.py:30, 71-83, 212)."""

    @staticmethod
    def forward(ctx, feat, n_way, n_support, n_query, *params):
        lib = _lib.load_library()
        _require_cuda(feat, "GnnHead")
        feat = feat.contiguous()
        feat_dim = feat.size(-1)
        params = _contig(params)                   # fc.weight [D, feat_dim], fc.bias, bn.weight, bn.bias
        D = params[0].shape[0]
        p = _lib.GconvParams()
        _fill_gc(p, params)
        nodes = torch.empty(n_query, n_way * (n_support + 1), D + n_way, dtype=torch.float32, device=feat.device)
        saved = _blob(lib.mft_head_saved_bytes(n_way, n_support, n_query, D), feat.device)
        ws = _blob(lib.mft_head_workspace_bytes(n_way, n_support, n_query, D), feat.device)
        with torch.cuda.device(feat.device):
            _lib.check(lib.mft_head_fwd(feat.data_ptr(), feat_dim, n_way, n_support, n_query, D, C.byref(p),
                                        nodes.data_ptr(), saved.data_ptr(), ws.data_ptr(), _stream()), "mft_head_fwd")
        ctx.save_for_backward(feat, saved, *params)
        ctx.meta = (n_way, n_support, n_query, D, feat_dim)
        return nodes

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_nodes):
        lib = _lib.load_library()
        feat, saved, *params = ctx.saved_tensors
        n_way, n_support, n_query, D, feat_dim = ctx.meta
        d_nodes = d_nodes.contiguous()
        p = _lib.GconvParams()
        _fill_gc(p, params)
        grads = _alloc_like_flat(params)
        g = _lib.GconvGrads()
        _fill_gc(g, grads)
        d_feat = torch.empty_like(feat) if ctx.needs_input_grad[0] else None
        ws = _blob(lib.mft_head_workspace_bytes(n_way, n_support, n_query, D), feat.device)
        with torch.cuda.device(feat.device):
            _lib.check(lib.mft_head_bwd(feat.data_ptr(), feat_dim, n_way, n_support, n_query, D, C.byref(p),
                                        d_nodes.data_ptr(), d_feat.data_ptr() if d_feat is not None else None,
                                        C.byref(g), saved.data_ptr(), ws.data_ptr(), _stream()), "mft_head_bwd")
        return (d_feat, None, None, None, *grads)


class _QueryCEFn(torch.autograd.Function):
    """GNN_nl output [n_query, n_way*(n_support+1), n_way] -> mean cross-entropy of the query nodes
    (select_scores + nn.CrossEntropyLoss against query_labels: gnnnet.py:216-224) with its gradient formed
    in the same launch (mft_query_ce)."""

    @staticmethod
    def forward(ctx, out, n_way, n_support, n_query):
        lib = _lib.load_library()
        _require_cuda(out, "GnnHead")
        out = out.contiguous()
        if tuple(out.shape) != (n_query, n_way * (n_support + 1), n_way):
            raise ValueError("query cross-entropy: output shape %s does not match the episode" % (tuple(out.shape),))
        loss = torch.empty((), dtype=torch.float32, device=out.device)
        d_out = torch.empty_like(out)
        with torch.cuda.device(out.device):
            _lib.check(lib.mft_query_ce(out.data_ptr(), n_way, n_support, n_query, loss.data_ptr(), d_out.data_ptr(),
                                        _stream()), "mft_query_ce")
        ctx.save_for_backward(d_out)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (d_out,) = ctx.saved_tensors
        return d_out * g, None, None, None


class GnnHead(nn.Module):
    """``fc`` + ``gnn`` of the reference GnnNet (gnnnet.py:30-31), driven on features.

    ``set_forward(feat)`` is ``GnnNet.set_forward(x, is_feature=True)`` (gnnnet.py:71-87);
    ``set_forward_loss(feat)`` adds the cross-entropy of gnnnet.py:219-224."""

    def __init__(self, n_way: int, n_support: int, feat_dim: int = 512, compress: bool = False,
                 share_support: bool = True):
        super().__init__()
        # The graphs build_graphs makes differ only in their query nodes: tell GNN_nl so that
        # layer_w0 evaluates the support-support pairs once (GNN_nl.shared_nodes).
        self.share_support = share_support
        # fc + BatchNorm1d + graph assembly as one library call per direction (mft_head_fwd/_bwd) instead
        # of the reference's Linear / BatchNorm1d / cat / expand kernels; the compressed variant
        # (gnnnet_copy.py) averages the supports in between and keeps the torch ops
        self.fused_pre_head = True
        # score selection + cross-entropy + its gradient as one launch (mft_query_ce) in set_forward_loss
        self.fused_loss = True
        self.n_way = n_way
        self.n_support_in = n_support
        self.compress = compress
        self.n_support = round(n_support / 2) if compress else n_support
        self.n_query = 16
        self.feat_dim = feat_dim
        self.fc = nn.Sequential(nn.Linear(feat_dim, 128), nn.BatchNorm1d(128, track_running_stats=False))
        self.gnn = GNN_nl(128 + n_way, 96, n_way)
        self.loss_fn = nn.CrossEntropyLoss()
        self.register_buffer("support_label", support_label(n_way, self.n_support), persistent=False)

    def nodes(self, feat: torch.Tensor) -> torch.Tensor:
        if self.fused_pre_head and not self.compress and feat.is_cuda and feat.dtype == torch.float32:
            lin, bn = self.fc[0], self.fc[1]
            per_class = feat.numel() // (self.n_way * feat.size(-1))
            if per_class != self.n_support + self.n_query:
                raise ValueError("GnnHead: %d rows per class, expected n_support + n_query = %d"
                                 % (per_class, self.n_support + self.n_query))
            return _HeadFn.apply(feat.reshape(self.n_way, per_class, feat.size(-1)), self.n_way, self.n_support,
                                 self.n_query, lin.weight, lin.bias, bn.weight, bn.bias)
        z = self.fc(feat.reshape(-1, feat.size(-1)))
        z = z.view(self.n_way, -1, z.size(1))
        return build_graphs(z, self.support_label, self.n_way, self.n_support_in, self.n_query, self.compress)

    def shared_mask(self):
        """True for the support nodes of a graph (node order of build_graphs: per class, the
        supports then the one query)."""
        return ([True] * self.n_support + [False]) * self.n_way

    def _set_sharing(self) -> None:
        # the head knows which nodes its graphs share: an explicit promise, no detection (GNN_nl.auto_share)
        self.gnn.shared_nodes = self.shared_mask() if self.share_support else None
        self.gnn.auto_share = False

    def forward_gnn_nodes(self, nodes: torch.Tensor) -> torch.Tensor:
        """``nodes`` must come from ``self.nodes`` / ``build_graphs`` (supports replicated)."""
        self._set_sharing()
        return select_scores(self.gnn(nodes), self.n_way, self.n_support, self.n_query)

    def set_forward(self, feat: torch.Tensor) -> torch.Tensor:
        return self.forward_gnn_nodes(self.nodes(feat))

    def _labels(self, device) -> torch.Tensor:
        key = (self.n_way, self.n_query, str(device))
        cache = self.__dict__.setdefault("_label_cache", {})
        if key not in cache:                       # built once per (shape, device): no H2D copy per step
            cache[key] = query_labels(self.n_way, self.n_query).to(device)
        return cache[key]

    def loss_from_nodes(self, nodes: torch.Tensor) -> torch.Tensor:
        """Cross-entropy of the query nodes of graphs made by ``self.nodes`` (gnnnet.py:210-224)."""
        if self.fused_loss and nodes.is_cuda and type(self.loss_fn) is nn.CrossEntropyLoss:
            self._set_sharing()
            return _QueryCEFn.apply(self.gnn(nodes), self.n_way, self.n_support, self.n_query)
        return self.loss_fn(self.forward_gnn_nodes(nodes), self._labels(nodes.device))

    def set_forward_loss(self, feat: torch.Tensor) -> torch.Tensor:
        return self.loss_from_nodes(self.nodes(feat))
