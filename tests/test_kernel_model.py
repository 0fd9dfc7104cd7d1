"""The symmetric-pair algorithm the CUDA kernels implement equals the reference function.

kernel_model.py (unordered pairs with multiplicities, biases dropped under BN,
closed-form backward) is compared in float64 with autograd through the oracle,
which test_oracle_golden.py pins to the reference.  CPU only.
"""
import numpy as np
import pytest
import torch

from oracle import gnn_oracle as O
from tests import kernel_model as K


def _rel(a, b):
    den = float(torch.linalg.norm(b))
    return float(torch.linalg.norm(a - b)) / (den if den > 0 else 1.0)


@pytest.mark.parametrize("bsz,n,fin,nf,n_way,seed", [(2, 5, 7, 8, 3, 0), (3, 9, 13, 16, 4, 1), (1, 2, 5, 4, 2, 2),
                                                      (4, 12, 21, 12, 5, 3)])
def test_symmetric_model_matches_oracle(bsz, n, fin, nf, n_way, seed):
    p = O.random_params(fin, nf, n_way, seed, torch.float64)
    g = torch.Generator().manual_seed(100 + seed)
    x = torch.randn(bsz, n, fin, generator=g, dtype=torch.float64)
    proj = torch.randn(bsz, n, n_way, generator=g, dtype=torch.float64)
    out_o, dx_o, grads_o = O.loss_and_grads(x, p, proj)
    out_k, dx_k, grads_k = K.gnn_nl_fwd_bwd(x, p, proj, nf // 2)
    assert _rel(out_k, out_o) < 1e-10
    assert _rel(dx_k, dx_o) < 1e-8
    assert set(grads_k) == set(grads_o)
    for name, go in grads_o.items():
        gk = grads_k[name].reshape(go.shape)
        if float(go.abs().max()) < 1e-9:       # analytically zero gradients (conv biases under BN, ...)
            assert float(gk.abs().max()) < 1e-9, name
        else:
            assert _rel(gk, go) < 1e-7, name


@pytest.mark.parametrize("bsz,n_way,n_support,fin,nf,seed", [(2, 2, 1, 5, 4, 0), (3, 3, 2, 9, 8, 1), (4, 5, 2, 13, 12, 2)])
def test_shared_support_rows_match_oracle(bsz, n_way, n_support, fin, nf, seed):
    """GNN_nl.shared_nodes: support-support pairs of layer_w0 as ONE row for all graphs (weight
    B x multiplicity, gradients summed over the graphs and delivered to graph 0) is still exactly
    the reference function -- logits, parameter gradients, per-graph dx of the query nodes, and the
    dx of the support nodes summed over the graphs (what the replication's backward adds up)."""
    mask = torch.tensor(([True] * n_support + [False]) * n_way)
    n = mask.numel()
    p = O.random_params(fin, nf, n_way, seed, torch.float64)
    g = torch.Generator().manual_seed(200 + seed)
    x = torch.randn(bsz, n, fin, generator=g, dtype=torch.float64)
    x[:, mask] = x[0, mask]
    proj = torch.randn(bsz, n, n_way, generator=g, dtype=torch.float64)
    out_o, dx_o, grads_o = O.loss_and_grads(x, p, proj)
    out_k, dx_k, grads_k = K.gnn_nl_fwd_bwd(x, p, proj, nf // 2, shared=mask)
    rb, ri, rj, w, allg = K.pair_rows(bsz, n, mask)
    assert int(allg.sum()) == (n_way * n_support) * (n_way * n_support + 1) // 2
    assert float(w.sum()) == bsz * n * n                    # the rows still stand for every ordered pair
    assert _rel(out_k, out_o) < 1e-10
    assert _rel(dx_k[:, ~mask], dx_o[:, ~mask]) < 1e-8
    assert _rel(dx_k[:, mask].sum(0), dx_o[:, mask].sum(0)) < 1e-8
    for name, go in grads_o.items():
        gk = grads_k[name].reshape(go.shape)
        if float(go.abs().max()) < 1e-9:
            assert float(gk.abs().max()) < 1e-9, name
        else:
            assert _rel(gk, go) < 1e-7, name


def test_wcompute_adjacency_properties():
    p = O.random_params(9, 8, 3, 7, torch.float64)
    x = torch.randn(2, 6, 9, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    adj, _ = K.wcompute_fwd(x, p, "layer_w0.")
    assert torch.allclose(adj.sum(2), torch.ones(2, 6, dtype=torch.float64))
    assert float(adj.diagonal(dim1=1, dim2=2).abs().max()) == 0.0
    ref = O.edge_adjacency(x, p, "layer_w0.")
    assert _rel(adj, ref) < 1e-12
