"""GPU tests of GNN_nl.shared_nodes (support-support pairs of layer_w0 evaluated once for all
graphs).  The contract: same logits, same parameter gradients, the same input gradient for the
per-graph nodes; for a shared node the input gradient summed over the graphs is the same (it is
delivered in graph 0's row) -- which is all the backward of GnnNet's replication (gnnnet.py:79-83,
torch.cat of the same z_support into every graph) ever uses."""
import numpy as np
import pytest
import torch

from oracle import gnn_oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu


def _episode_mask(n_way, n_support):
    return np.array(([True] * n_support + [False]) * n_way)


def _shared_problem(bsz, n_way, n_support, fin, nf, seed0, tau, max_tries=300):
    """Kink-free seeded problem whose support nodes are identical in every graph."""
    mask = _episode_mask(n_way, n_support)
    n = mask.size
    for t in range(max_tries):
        seed = seed0 + 7919 * t
        p64 = O.random_params(fin, nf, n_way, seed, torch.float64)
        p32 = {k: v.float() for k, v in p64.items()}
        g = torch.Generator().manual_seed(1000 + seed)
        x = torch.randn(bsz, n, fin, generator=g)
        x[:, torch.from_numpy(mask)] = x[0, torch.from_numpy(mask)]
        proj = torch.randn(bsz, n, n_way, generator=g)
        if tau is None or O.min_abs_preactivation(x.double(), {k: v.double() for k, v in p32.items()}) > tau:
            return {k: v.numpy() for k, v in p32.items()}, x, proj, mask
    raise RuntimeError("no kink-free seed found")


def _run(x, params, proj, fin, nf, n_way, precision, mask, fused=True, check=False):
    import mft_b200
    mft_b200.set_precision(precision)
    net = U.load_params_into(mft_b200.GNN_nl(fin, nf, n_way), params).cuda()
    net.fused = fused
    net.shared_nodes = None if mask is None else [bool(v) for v in mask]
    net.check_shared = check
    xg = x.float().cuda().requires_grad_(True)
    out = net(xg)
    (out * proj.float().cuda()).sum().backward()
    torch.cuda.synchronize()
    mft_b200.set_precision("auto")
    grads = {k: v.grad.detach().double().cpu().numpy() for k, v in net.named_parameters()}
    return out.detach().double().cpu().numpy(), xg.grad.double().cpu().numpy(), grads


def _fold(dx, mask):
    """What the replication's backward sees: per-graph rows for the queries, the sum over graphs
    for the shared nodes."""
    return dx[:, ~mask], dx[:, mask].sum(0)


@pytest.mark.parametrize("bsz,n_way,n_support,fin,nf,seed0", [
    (2, 2, 1, 5, 4, 0),
    (3, 3, 2, 21, 12, 1),
    (4, 5, 1, 40, 24, 2),
    (3, 2, 3, 133, 96, 3),
])
@pytest.mark.parametrize("fused", [True, False])
def test_shared_nodes_fp32_matches_oracle_sharply(bsz, n_way, n_support, fin, nf, seed0, fused):
    params, x, proj, mask = _shared_problem(bsz, n_way, n_support, fin, nf, seed0, tau=3e-5)
    out_t, dx_t, g_t = U.oracle_truth(x, params, proj, torch.float64)
    out_y, dx_y, g_y = U.oracle_truth(x, params, proj, torch.float32)
    out, dx, g = _run(x, params, proj, fin, nf, n_way, "fp32", mask, fused, check=True)
    assert U.rel(out, out_t) < 1e-5
    for got, want, yard in zip(_fold(dx, mask), _fold(dx_t, mask), _fold(dx_y, mask)):
        assert U.rel(got, want) < max(3 * U.rel(yard, want), 2e-5)
    for k, gt in g_t.items():
        if U.is_zero_grad(k):
            assert np.abs(g[k]).max() <= 1e-6
        else:
            assert U.rel(g[k].reshape(gt.shape), gt) < max(3 * U.rel(g_y[k], gt), 2e-5), k


def test_shared_rows_of_other_graphs_carry_no_shared_pair_gradient():
    """Documented delivery: with ONE graph-independent upstream gradient on the adjacency the
    support-support part of dx sits in graph 0."""
    import mft_b200
    mft_b200.set_precision("fp32")
    mask = _episode_mask(3, 2)
    n = mask.size
    torch.manual_seed(0)
    m = mft_b200.Wcompute(20, 8).cuda()
    x = torch.randn(4, n, 20)
    x[:, torch.from_numpy(mask)] = x[0, torch.from_numpy(mask)]
    sel = torch.from_numpy(mask).cuda()
    up = torch.zeros(4, n, n, device="cuda")
    up[:, sel.nonzero()[0], sel.nonzero()[1]] = 1.0     # upstream only on one support-support edge
    res = []
    for shared in (None, [bool(v) for v in mask]):
        xg = x.cuda().requires_grad_(True)
        m.adjacency(xg, shared).backward(up)
        res.append(xg.grad.clone())
    dense, folded = res
    assert torch.allclose(dense[:, ~sel], folded[:, ~sel], rtol=1e-4, atol=1e-7)
    assert torch.allclose(dense[:, sel].sum(0), folded[:, sel].sum(0), rtol=1e-4, atol=1e-7)
    mft_b200.set_precision("auto")


@pytest.mark.parametrize("precision,out_tol,g_tol", [("fp32", 2e-5, 3e-2), ("tf32", 2e-3, None)])
@pytest.mark.parametrize("n_support,n_query", [(5, 16), (20, 16), (5, 15)])
def test_shared_equals_dense_on_episode_shapes(precision, out_tol, g_tol, n_support, n_query):
    """Full reference shapes: shared on/off agree (inputs are not kink-free here, so gradients
    are compared at the tolerance the slope flips allow, tests/test_gpu_parity.py)."""
    n_way, fin, nf = 5, 133, 96
    params, x, proj, mask = _shared_problem(n_query, n_way, n_support, fin, nf, 11, tau=None)
    out_d, dx_d, g_d = _run(x, params, proj, fin, nf, n_way, precision, None)
    out_s, dx_s, g_s = _run(x, params, proj, fin, nf, n_way, precision, mask)
    assert U.rel(out_s, out_d) < out_tol
    n = mask.size
    tol = g_tol if g_tol is not None else max(0.15, 2.0 / np.sqrt(n_query * n * (n + 1) / 2))
    for a, b in zip(_fold(dx_s, mask), _fold(dx_d, mask)):
        assert U.rel(a, b) < tol
    for k in g_d:
        if U.is_zero_grad(k):
            assert np.abs(g_s[k]).max() <= 1e-6
        else:
            assert U.rel(g_s[k], g_d[k]) < tol, (k, U.rel(g_s[k], g_d[k]))


def test_check_shared_rejects_a_false_promise():
    import mft_b200
    net = mft_b200.GNN_nl(13, 16, 3).cuda()
    net.shared_nodes = [True, False] * 3
    net.check_shared = True
    with pytest.raises(ValueError, match="differ between graphs"):
        net(torch.randn(4, 6, 13, device="cuda"))
    with pytest.raises(ValueError, match="entries for"):
        net.shared_nodes = [True] * 5
        net(torch.randn(4, 6, 13, device="cuda"))


def test_gnn_head_shares_support_by_default_and_matches_unshared():
    import mft_b200
    mft_b200.set_precision("fp32")
    torch.manual_seed(5)
    head = mft_b200.GnnHead(5, 5).cuda()
    feat = torch.randn(5, 5 + 16, 512, device="cuda")
    res = []
    for share in (True, False):
        head.share_support = share
        head.zero_grad(set_to_none=True)
        loss = head.set_forward_loss(feat)
        loss.backward()
        res.append((float(loss.detach()), {k: v.grad.clone() for k, v in head.named_parameters()}))
        assert (head.gnn.shared_nodes is not None) == share
    assert abs(res[0][0] - res[1][0]) < 1e-5
    for k, a in res[0][1].items():
        b = res[1][1][k]
        den = float(b.norm())
        if den > 1e-6:
            assert float((a - b).norm()) / den < 3e-2, k
    mft_b200.set_precision("auto")


def test_auto_share_finds_the_replicated_supports_of_the_reference_graphs():
    """Unchanged reference callers never set shared_nodes: GNN_nl then detects (eager calls only) the nodes
    whose rows are bitwise identical in every graph -- exactly the supports of forward_gnn's graphs
    (gnnnet.py:83,212) -- and gets the same logits / parameter gradients as with sharing off."""
    import mft_b200
    mft_b200.set_precision("fp32")
    n_way, n_support, n_query, fin = 5, 5, 16, 133
    torch.manual_seed(3)
    net = mft_b200.GNN_nl(fin, 96, n_way).cuda()
    z = torch.randn(n_way, n_support + n_query, fin, device="cuda", requires_grad=True)
    res = []
    for auto in (True, False):
        net.auto_share = auto
        net.zero_grad(set_to_none=True)
        z.grad = None
        nodes = torch.cat([torch.cat([z[:, :n_support], z[:, n_support + i:n_support + i + 1]], dim=1)
                           .reshape(1, -1, fin) for i in range(n_query)], dim=0)      # gnnnet.py:83
        mask = net._shared(nodes)
        if auto:
            assert mask == bytes(([1] * n_support + [0]) * n_way)
        else:
            assert mask is None
        out = net(nodes)
        out.square().sum().backward()
        res.append((out.detach().clone(), z.grad.clone(), {k: v.grad.clone() for k, v in net.named_parameters()}))
    (o1, dz1, g1), (o0, dz0, g0) = res
    assert U.rel(o1.cpu().numpy(), o0.cpu().numpy()) < 2e-5
    assert U.rel(dz1.cpu().numpy(), dz0.cpu().numpy()) < 3e-2
    for k in g0:
        if not U.is_zero_grad(k):
            assert U.rel(g1[k].cpu().numpy(), g0[k].cpu().numpy()) < 3e-2, k
    # a leaf input that requires grad keeps per-graph gradients: no detection
    net.auto_share = True
    leaf = nodes.detach().clone().requires_grad_(True)
    assert net._shared(leaf) is None
    mft_b200.set_precision("auto")


def test_non_identity_operator_is_refused():
    import mft_b200
    mft_b200.set_precision("fp32")
    m = mft_b200.Wcompute(13, 8).cuda()
    x = torch.randn(2, 5, 13, device="cuda")
    eye = torch.eye(5, device="cuda").view(1, 5, 5, 1).repeat(2, 1, 1, 1)
    W = m(x, eye)
    assert W.shape == (2, 5, 5, 2)
    with pytest.raises(RuntimeError, match="not the identity"):
        m(x, 2 * eye)
    gc = mft_b200.Gconv(13, 4, 2).cuda()
    gc([W, x])
    bad = W.clone()
    bad[..., 0] = 0.5
    with pytest.raises(RuntimeError, match="not the identity"):
        gc([bad, x])
    mft_b200.set_precision("auto")


def test_double_backward_is_refused():
    """The hand-written backward is first order (the reference's FO-MAML needs no more): asking autograd to
    differentiate through it raises instead of silently treating the gradient as a constant."""
    import mft_b200
    mft_b200.set_precision("fp32")
    net = mft_b200.GNN_nl(13, 16, 3).cuda()
    x = torch.randn(2, 6, 13, device="cuda", requires_grad=True)
    (gx,) = torch.autograd.grad(net(x).sum(), x, create_graph=True)
    with pytest.raises(RuntimeError):
        gx.sum().backward()
    mft_b200.set_precision("auto")
