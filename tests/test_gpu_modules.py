"""GPU tests of the module surface beyond GNN_nl: Wcompute and Gconv called on their own (as the
reference API allows), inference under no_grad, eval()/train() equivalence, deepcopy, state_dict
round trip, the GnnHead mirror against the oracle (5-shot and compressed 50-shot), and CUDA-graph
replay of a training step."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import gnn_oracle as O
from tests import util as U

pytestmark = pytest.mark.gpu


def _params(fin, nf, n_way, seed):
    return {k: v.float() for k, v in O.random_params(fin, nf, n_way, seed, torch.float64).items()}


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-6), ("tf32", 2e-3)])
def test_wcompute_module_alone(prec, tol):
    import mft_b200
    mft_b200.set_precision(prec)
    fin, nf, bsz, n = 133, 96, 3, 11
    p = _params(fin, nf, 5, 21)
    m = mft_b200.Wcompute(fin, nf).cuda()
    m.load_state_dict({k[len("layer_w0."):]: v for k, v in p.items() if k.startswith("layer_w0.")})
    x = torch.randn(bsz, n, fin, generator=torch.Generator().manual_seed(1))
    eye = torch.eye(n).unsqueeze(0).repeat(bsz, 1, 1).unsqueeze(3)
    with torch.no_grad():
        out = m(x.cuda(), eye.cuda()).cpu()
        ref = O.wcompute(x.double(), {k: v.double() for k, v in p.items()}, "layer_w0.")
    assert out.shape == (bsz, n, n, 2)
    assert torch.equal(out[..., 0], eye[..., 0])
    assert U.rel(out[..., 1].numpy(), ref[..., 1].numpy()) < tol
    mft_b200.set_precision("auto")


@pytest.mark.parametrize("fin,nout,bsz,n,bn_bool", [(37, 12, 3, 9, True), (133, 48, 16, 30, True), (229, 48, 4, 105, True),
                                                    (229, 5, 4, 105, False), (70, 33, 2, 41, True), (181, 5, 3, 30, False)])
def test_gconv_module_alone_forward_backward(fin, nout, bsz, n, bn_bool):
    """Gconv alone, incl. the real widths and the BN-less last layer: output, input / adjacency gradients and every
    parameter gradient against the float64 oracle (the backward runs on n_out-wide products and needs the x Wb^T
    the forward saved; a non-multiple-of-32 everything exercises the tile edges of the dual-K product)."""
    import mft_b200
    g = torch.Generator().manual_seed(3)
    adj = torch.softmax(torch.randn(bsz, n, n, generator=g, dtype=torch.float64), dim=2)
    eye = torch.eye(n, dtype=torch.float64).expand(bsz, n, n)
    W = torch.stack([eye, adj], dim=3)
    x = torch.randn(bsz, n, fin, generator=g, dtype=torch.float64)
    m = mft_b200.Gconv(fin, nout, 2, bn_bool=bn_bool).cuda()
    with torch.no_grad():
        if bn_bool:
            m.bn.weight.uniform_(0.5, 1.5)
            m.bn.bias.uniform_(-0.5, 0.5)
    p = {"l.fc.weight": m.fc.weight.detach().double().cpu().requires_grad_(True),
         "l.fc.bias": m.fc.bias.detach().double().cpu().requires_grad_(True)}
    if bn_bool:
        p["l.bn.weight"] = m.bn.weight.detach().double().cpu().requires_grad_(True)
        p["l.bn.bias"] = m.bn.bias.detach().double().cpu().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    Wr = W.clone().requires_grad_(True)
    ref = O.gconv(Wr, xr, p, "l.", bn_bool=bn_bool)
    proj = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (ref * proj).sum().backward()
    xg = x.float().cuda().requires_grad_(True)
    Wg = W.float().cuda().requires_grad_(True)
    W_out, out = m([Wg, xg])
    assert W_out is Wg
    (out * proj.float().cuda()).sum().backward()
    assert U.rel(out.detach().cpu().numpy(), ref.detach().numpy()) < 5e-6
    assert U.rel(xg.grad.cpu().numpy(), xr.grad.numpy()) < 5e-5
    assert U.rel(Wg.grad[..., 1].cpu().numpy(), Wr.grad[..., 1].numpy()) < 5e-5
    assert U.rel(m.fc.weight.grad.cpu().numpy(), p["l.fc.weight"].grad.numpy()) < 5e-5
    if bn_bool:
        assert U.rel(m.bn.weight.grad.cpu().numpy(), p["l.bn.weight"].grad.numpy()) < 5e-5
        assert U.rel(m.bn.bias.grad.cpu().numpy(), p["l.bn.bias"].grad.numpy()) < 5e-5
        assert float(m.fc.bias.grad.abs().max()) == 0.0      # BatchNorm1d removes the mean
    else:
        assert U.rel(m.fc.bias.grad.cpu().numpy(), p["l.fc.bias"].grad.numpy()) < 5e-5


def test_inference_modes_and_copies_agree():
    import mft_b200
    mft_b200.set_precision("fp32")
    torch.manual_seed(0)
    net = mft_b200.GNN_nl(133, 96, 5).cuda()
    x = torch.randn(4, 30, 133, device="cuda")
    with torch.no_grad():
        a = net(x)
        net.eval()
        b = net(x)                    # no running statistics anywhere: eval() changes nothing
        net.train()
        c = copy.deepcopy(net)(x)
        d = mft_b200.GNN_nl(133, 96, 5).cuda()
        d.load_state_dict(net.state_dict())
        dd = d(x)
        net.fused = False
        e = net(x)
    assert torch.equal(a, b) and torch.equal(a, c) and torch.equal(a, dd)
    assert torch.allclose(a, e, rtol=0, atol=1e-5)
    mft_b200.set_precision("auto")


@pytest.mark.parametrize("n_support,compress,fixture", [(5, False, "head_5w5s.npz"), (50, True, "head_50c.npz")])
def test_gnn_head_matches_reference_scores(golden_dir, n_support, compress, fixture):
    """GnnHead (fc -> graphs -> GNN -> select) against scores the reference GnnNet / gnnnet_copy
    produced on the same features and parameters (tests/golden, is_feature path of finetune.py)."""
    import mft_b200
    rec = dict(np.load(os.path.join(golden_dir, fixture)))
    prm = dict(np.load(os.path.join(golden_dir, "head_5w5s.npz")))
    head = mft_b200.GnnHead(5, n_support, compress=compress)
    head.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in prm.items() if k.startswith("p.")})
    head = head.cuda()
    head.n_query = 15
    feat = torch.from_numpy(rec["feat15"] if "feat15" in rec else rec["feat"]).cuda()
    want = rec["scores15"] if "scores15" in rec else rec["scores"]
    for prec, tol in (("fp32", 2e-5), ("tf32", 2e-3)):
        mft_b200.set_precision(prec)
        with torch.no_grad():
            got = head.set_forward(feat).cpu().numpy()
        assert got.shape == want.shape
        assert U.rel(got, want) < tol, (prec, U.rel(got, want))
    mft_b200.set_precision("auto")


def test_gnn_head_gradients_match_reference(golden_dir):
    """set_forward_loss + backward of GnnHead (fp32 path) against the gradients of the reference's GnnNet on
    the same features and parameters (float64 run, tests/golden/head_5w5s_grads.npz): every fc.* / gnn.*
    parameter and the features, within max(3 x the reference's own fp32 error, 3e-2) -- the fixture is not
    kink-free, see tests/test_gpu_parity.py -- and the loss to 1e-5."""
    import mft_b200
    prm = dict(np.load(os.path.join(golden_dir, "head_5w5s.npz")))
    gr = dict(np.load(os.path.join(golden_dir, "head_5w5s_grads.npz")))
    mft_b200.set_precision("fp32")
    for share in (True, False):
        head = mft_b200.GnnHead(5, 5, share_support=share)
        head.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in prm.items() if k.startswith("p.")})
        head = head.cuda()
        head.n_query = 16
        feat = torch.from_numpy(prm["feat16"]).cuda().requires_grad_(True)
        loss = head.set_forward_loss(feat)
        loss.backward()
        assert abs(float(loss.detach()) - float(gr["loss64"])) < 1e-5
        assert U.rel(feat.grad.cpu().numpy(), gr["dfeat"]) < max(3 * float(gr["e32.dfeat"]), 3e-2)
        for k, v in head.named_parameters():
            want = gr["g." + k]
            if np.abs(want).max() < 1e-9:
                assert float(v.grad.abs().max()) <= 1e-6, k
            else:
                e = U.rel(v.grad.cpu().numpy(), want)
                assert e < max(3 * float(gr["e32." + k]), 3e-2), (k, e, float(gr["e32." + k]))
    mft_b200.set_precision("auto")


def test_cuda_graph_replay_equals_eager():
    import mft_b200
    mft_b200.set_precision("tf32")
    torch.manual_seed(1)
    head = mft_b200.GnnHead(5, 5).cuda()
    head.n_query = 16
    params = list(head.parameters())
    feats = [torch.randn(5, 21, 512, device="cuda") for _ in range(3)]
    step = mft_b200.GraphedStep(lambda f: head.set_forward_loss(f), [feats[0]], params, inputs_require_grad=False)
    for f in feats:
        loss_g = float(step(f))
        grads_g = [p.grad.clone() for p in params]
        for p in params:
            p.grad = None
        loss_e = head.set_forward_loss(f)
        loss_e.backward()
        assert abs(loss_g - float(loss_e.detach())) < 1e-5
        for a, p in zip(grads_g, params):
            den = float(p.grad.norm())
            if den > 1e-6:      # atomics reorder sums between runs: tiny differences only
                assert float((a - p.grad).norm()) / den < 5e-3
        for p in params:
            p.grad = None
    mft_b200.set_precision("auto")


@pytest.mark.parametrize("n_support", [5, 20])
def test_replayed_step_is_stable_over_many_replays(n_support):
    """Soak: 300 replays of the captured head step (forward + backward; side streams for the wgrads, the Gconv
    weight-gradient products, the hoisted tables / images and x W products) on the same input.  The loss, the input
    gradient and every conv / BatchNorm gradient of the tensor-core path have a fixed summation order: a replay
    that differs in one bit is a cross-stream race.  The split-K (atomic) weight gradients may move in the last
    bits only."""
    import mft_b200
    mft_b200.set_precision("tf32")
    torch.manual_seed(7)
    head = mft_b200.GnnHead(5, n_support).cuda()
    head.n_query = 16
    params = list(head.gnn.parameters())
    names = [n for n, _ in head.gnn.named_parameters()] + ["d_nodes"]
    with torch.no_grad():
        nodes = head.nodes(torch.randn(5, n_support + 16, 512, device="cuda")).contiguous()
    step = mft_b200.GraphedStep(lambda x: head.loss_from_nodes(x), [nodes], params)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def snapshot():
        loss = step(nodes).clone()
        return loss, [p.grad.clone() for p in params] + [step.static_inputs[0].grad.clone()]

    loss0, g0 = snapshot()
    for it in range(300):
        if it % 50 == 0:
            flush.fill_(it & 255)                # move the cache state around between replays
        loss, g = snapshot()
        assert torch.equal(loss, loss0), it
        for n, a, b in zip(names, g, g0):
            if ".fc." in n:                      # Gconv Linear: split-K atomics
                den = float(b.norm())
                assert den < 1e-12 or float((a - b).norm()) / den < 1e-5, (it, n)
            else:
                assert torch.equal(a, b), (it, n)
    mft_b200.set_precision("auto")


@pytest.mark.parametrize("n_support,n_query", [(5, 16), (20, 16), (1, 15), (5, 3)])
def test_fused_pre_head_equals_torch_ops(n_support, n_query):
    """mft_head_fwd/_bwd (fc Linear + BatchNorm1d + graph assembly + labels) against the reference's
    op sequence (nn.Linear, nn.BatchNorm1d, cat/expand; gnnnet.py:71-83, 212) run by torch."""
    import mft_b200
    torch.manual_seed(7)
    head = mft_b200.GnnHead(5, n_support).cuda()
    head.n_query = n_query
    feat = torch.randn(5, n_support + n_query, 512, device="cuda")
    up = torch.randn(n_query, 5 * (n_support + 1), 133, device="cuda")
    res = []
    for fused in (True, False):
        head.fused_pre_head = fused
        head.zero_grad(set_to_none=True)
        f = feat.clone().requires_grad_(True)
        nodes = head.nodes(f)
        (nodes * up).sum().backward()
        res.append((nodes.detach(), f.grad.clone(), {k: v.grad.clone() for k, v in head.fc.named_parameters()}))
    (n1, df1, g1), (n0, df0, g0) = res
    assert n1.shape == n0.shape
    assert torch.equal(n1[..., 128:], n0[..., 128:])                     # label one-hots: exact
    assert U.rel(n1.cpu().numpy(), n0.cpu().numpy()) < 2e-6
    assert U.rel(df1.cpu().numpy(), df0.cpu().numpy()) < 2e-5
    for k in g0:
        a, b = g1[k].cpu().numpy(), g0[k].cpu().numpy()
        if k == "0.bias":                       # Linear bias under BatchNorm: analytically zero
            assert np.abs(a).max() <= 1e-6 and np.abs(b).max() < 1e-3
        else:
            assert U.rel(a, b) < 5e-5, (k, U.rel(a, b))


def test_fused_pre_head_without_feature_gradient_and_bad_shape():
    import mft_b200
    head = mft_b200.GnnHead(5, 5).cuda()
    head.n_query = 16
    feat = torch.randn(5, 21, 512, device="cuda")
    loss = head.set_forward_loss(feat)          # feat carries no gradient: d_feat is not computed
    loss.backward()
    assert all(p.grad is not None for p in head.parameters())
    with pytest.raises(ValueError, match="rows per class"):
        head.nodes(torch.randn(5, 20, 512, device="cuda"))


@pytest.mark.parametrize("fin", [133, 181, 229])
def test_tf32_weight_gradients_are_bitwise_reproducible(fin):
    """The tensor-core backward has no data-dependent summation order (private partial dW copies, fixed
    reduction order): two runs on the same inputs must give IDENTICAL conv-weight gradients at the full
    5-way 20-shot row count.  (A 10 % run-to-run spread of d conv2d_1.weight once hid inside the
    slope-flip tolerance of the parity tests; this pins it.)"""
    import mft_b200
    mft_b200.set_precision("tf32")
    torch.manual_seed(fin)
    m = mft_b200.Wcompute(fin, 96).cuda()
    x = torch.randn(16, 105, fin)
    up = torch.randn(16, 105, 105).cuda()
    runs = []
    for _ in range(3):
        for prm in m.parameters():
            prm.grad = None
        xg = x.cuda().requires_grad_(True)
        m.adjacency(xg, None).backward(up)
        torch.cuda.synchronize()
        runs.append(({k: v.grad.clone() for k, v in m.named_parameters()}, xg.grad.clone()))
    mft_b200.set_precision("auto")
    for g, dx in runs[1:]:
        assert torch.equal(dx, runs[0][1])
        for k in ("conv2d_1.weight", "conv2d_2.weight", "conv2d_3.weight", "conv2d_4.weight"):
            assert torch.equal(g[k], runs[0][0][k]), k


def test_programmatic_dependent_launch_levels_give_identical_results():
    """mft_set_pdl(0/1/2) only changes how launches overlap, never what they compute."""
    import mft_b200
    lib = mft_b200.load_library()
    mft_b200.set_precision("tf32")
    torch.manual_seed(5)
    net = mft_b200.GNN_nl(133, 96, 5).cuda()
    x = torch.randn(4, 30, 133)
    res = []
    old = lib.mft_set_pdl(0)
    try:
        for level in (0, 1, 2):
            lib.mft_set_pdl(level)
            for prm in net.parameters():
                prm.grad = None
            xg = x.cuda().requires_grad_(True)
            out = net(xg)
            out.square().sum().backward()
            torch.cuda.synchronize()
            res.append((out.detach().clone(), xg.grad.clone(),
                        {k: v.grad.clone() for k, v in net.named_parameters() if "fc.weight" not in k}))
    finally:
        lib.mft_set_pdl(old)
        mft_b200.set_precision("auto")
    for out, dx, g in res[1:]:
        assert torch.equal(out, res[0][0])
        assert torch.equal(dx, res[0][1])
        for k in g:                       # (Gconv fc.weight gradients use split-K atomics: excluded)
            assert torch.equal(g[k], res[0][2][k]), k


@pytest.mark.parametrize("n_support,n_query", [(5, 16), (20, 15)])
def test_fused_query_cross_entropy_equals_torch(n_support, n_query):
    """mft_query_ce = select_scores + nn.CrossEntropyLoss(query_labels) (gnnnet.py:216-224), value and gradient."""
    import mft_b200
    from mft_b200 import episode as E
    n_way = 5
    n = n_way * (n_support + 1)
    out = torch.randn(n_query, n, n_way, generator=torch.Generator().manual_seed(n_support)).mul(3).cuda()
    a = out.clone().requires_grad_(True)
    loss_a = E._QueryCEFn.apply(a, n_way, n_support, n_query)
    (loss_a * 1.7).backward()
    b = out.clone().requires_grad_(True)
    loss_b = torch.nn.functional.cross_entropy(E.select_scores(b, n_way, n_support, n_query),
                                               E.query_labels(n_way, n_query).cuda())
    (loss_b * 1.7).backward()
    assert abs(float(loss_a.detach()) - float(loss_b.detach())) < 2e-6 * max(1.0, abs(float(loss_b.detach())))
    assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-8)
    sup = torch.ones(n, dtype=torch.bool)
    sup[n_support::n_support + 1] = False
    assert torch.count_nonzero(a.grad[:, sup]) == 0          # support nodes: exact zeros
    with pytest.raises(ValueError):
        E._QueryCEFn.apply(out[:, :-1], n_way, n_support, n_query)


def test_gnn_head_fused_loss_matches_unfused():
    import mft_b200
    from mft_b200.episode import GnnHead
    mft_b200.set_precision("fp32")      # (on the TF32 path a 1e-7 change of d_out may flip LeakyReLU slopes)
    torch.manual_seed(3)
    head = GnnHead(5, 5).cuda()
    head.n_query = 16
    feat = torch.randn(5, 21, 512).cuda()
    res = []
    for fused in (True, False):
        head.fused_loss = fused
        head.zero_grad(set_to_none=True)
        loss = head.set_forward_loss(feat)
        loss.backward()
        res.append((float(loss), {k: v.grad.double().cpu().numpy() for k, v in head.named_parameters()}))
    mft_b200.set_precision("auto")
    assert abs(res[0][0] - res[1][0]) < 1e-5
    for k in res[0][1]:
        if U.is_zero_grad("gnn." + k) or U.is_zero_grad(k):
            continue
        assert U.rel(res[0][1][k], res[1][1][k]) < 3e-2, (k, U.rel(res[0][1][k], res[1][1][k]))
