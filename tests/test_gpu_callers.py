"""The reference's OWN driver classes (GnnNet, gnnnet_copy.GnnNet, DampNet -- files taken verbatim from the
reference by oracle/make_ref.py into oracle/_ref) run on the GPU twice: once exactly as the reference is
(its methods/gnn.py on ATen / cuDNN / cuBLAS, TF32 switched off), once over this repo's ``methods/`` overlay
(same files, ``methods.gnn`` served by libmft_gnn.so -- what a user gets with PYTHONPATH set).
Same seeds, same inputs; the two runs must agree.  SURVEY.md section 8 rows a5-a8: forward_gnn / set_forward /
set_forward_loss, the first-order-MAML set_forward_finetune + MAML_update, the compressed 50-shot GnnNet and
its train_loop50, DampNet's three set_forward branches and train_loop_full.

Tolerances (fp32 path of the library against the reference's fp32 eager run): scores / loss 2e-5 relative;
gradients and SGD-updated parameters 3e-2 per tensor (two fp32 evaluations of a function with LeakyReLU
kinks, see tests/test_gpu_parity.py; sharp gradient parity is pinned there and in tests/test_gpu_tape.py).
"""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader as R

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (python oracle/make_ref.py)")]


@pytest.fixture(autouse=True)
def _exact_fp32_reference():
    import mft_b200
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32,
           torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    mft_b200.set_precision("fp32")
    yield
    mft_b200.set_precision("auto")
    (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32,
     torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark) = old


def _rel(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    den = float(b.norm())
    return float((a - b).norm()) / (den if den > 0 else 1.0)


def _build(make):
    """The same model under both variants: identical seeds -> identical initial parameters."""
    out = []
    for variant in ("reference", "overlay"):
        ns = R.load(variant)
        torch.manual_seed(1234)
        np.random.seed(10)
        out.append((ns, make(ns)))
    (_, a), (_, b) = out
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    import mft_b200
    assert isinstance(b.gnn, mft_b200.GNN_nl) and not isinstance(a.gnn, mft_b200.GNN_nl)
    return a, b


def _compare_grads(ref, ours, tol=3e-2, names=None):
    worst = ("", 0.0)
    for (k, p), (_, q) in zip(ref.named_parameters(), ours.named_parameters()):
        if names is not None and not k.startswith(names):
            continue
        if p.grad is None:
            assert q.grad is None or float(q.grad.abs().max()) <= 1e-6, k
            continue
        assert q.grad is not None, k
        if float(p.grad.norm()) < 1e-6:              # analytically zero (conv biases under BatchNorm, ...)
            assert float(q.grad.abs().max()) <= 1e-5, k
            continue
        e = _rel(q.grad, p.grad)
        if e > worst[1]:
            worst = (k, e)
        assert e < tol, (k, e)
    return worst


def test_gnnnet_set_forward_loss_on_images():
    """train.py --method gnnnet: set_forward_loss on a 5-way 5-shot 16-query episode of 224x224 images
    (gnnnet.py:68-87, 210-224), backward through head AND backbone."""
    a, b = _build(lambda ns: ns.gnnnet.GnnNet(ns.backbone.ResNet10, 5, 5).cuda())
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 21, 3, 224, 224, generator=g)
    losses = []
    for m in (a, b):
        m.n_query = 16
        loss = m.set_forward_loss(x)
        loss.backward()
        losses.append(float(loss.detach()))
    assert abs(losses[0] - losses[1]) < 2e-5 * max(1.0, abs(losses[0])), losses
    _compare_grads(a, b)


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("tf32", 1e-3), ("auto", 1e-3)])
def test_gnnnet_feature_path_of_finetune_py(precision, tol):
    """finetune.py:312-316: model.n_query = 15; model.set_forward(features, is_feature=True)."""
    import mft_b200
    a, b = _build(lambda ns: ns.gnnnet.GnnNet(ns.backbone.ResNet10, 5, 20).cuda())
    g = torch.Generator().manual_seed(4)
    feat = torch.randn(5, 35, 512, generator=g)
    mft_b200.set_precision(precision)
    with torch.no_grad():
        a.n_query = b.n_query = 15
        sa, sb = a.set_forward(feat, is_feature=True), b.set_forward(feat, is_feature=True)
    assert sa.shape == sb.shape == (75, 5)
    assert _rel(sb, sa) < tol, _rel(sb, sa)
    assert torch.equal(sa.argmax(1), sb.argmax(1)) or precision != "fp32"


def test_gnnnet_first_order_maml_loop_and_rewind():
    """train.py --fine_tune: MetaTemplate.train_loop_finetune -> set_forward_loss_finetune ->
    set_forward_finetune (15 inner epochs of Adam on a deep copy of the backbone, then the outer forward
    through forward_gnn) -> backward -> step.  After the FIRST outer step the two runs must agree (same inner
    loop bit for bit -- it does not touch the head -- so only the head's arithmetic differs); a second
    episode then exercises MAML_update's rewind with feature2 / feature3 and the final MAML_update of
    train.py:54-55 (105 more Adam steps amplify the first step's rounding differences, so only the loss and
    the bookkeeping are compared there)."""
    a, b = _build(lambda ns: ns.gnnnet.GnnNet(ns.backbone.ResNet10, 5, 5).cuda())
    init = {k: v.clone() for k, v in a.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    episodes = [(torch.randn(5, 21, 3, 224, 224, generator=g), None) for _ in range(2)]
    opts = []
    for m in (a, b):
        torch.manual_seed(77)
        np.random.seed(10)
        opts.append(torch.optim.SGD(m.parameters(), lr=0.05))
        m.train_loop_finetune(0, episodes[:1], opts[-1])
    assert not a.first and not b.first
    _compare_grads(a, b)                                   # outer gradients at the adapted point
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())              # incl. the feature2.* / feature3.* copies
    for k in init:
        if sa[k].dtype.is_floating_point and sa[k].numel() > 1 and float((sa[k] - init[k]).norm()) > 1e-7:
            assert _rel(sb[k] - init[k], sa[k] - init[k]) < 3e-2, k
    for k in sa:                                           # the inner loop's result is identical
        if k.startswith(("feature2.", "feature3.")):
            assert torch.equal(sa[k], sb[k]), k
    losses = []
    for m, opt in zip((a, b), opts):
        torch.manual_seed(78)
        np.random.seed(11)
        opt.zero_grad()
        loss = m.set_forward_loss_finetune(episodes[1][0])   # calls MAML_update (rewind) first
        loss.backward()
        opt.step()
        m.MAML_update()
        losses.append(float(loss.detach()))
    assert abs(losses[0] - losses[1]) < 2e-2 * max(1.0, abs(losses[0])), losses
    for k, v in b.state_dict().items():
        assert torch.isfinite(v).all(), k


def test_gnnnet_copy_compressed_50_shot():
    """finetune_50.py / train_50.py: gnnnet_copy.GnnNet halves the supports (n_support 50 -> 25, N = 130);
    feature path with 15 queries, then one train_loop50 step on images (5 x 66 x 3 x 224 x 224)."""
    a, b = _build(lambda ns: ns.gnnnet_copy.GnnNet(ns.backbone.ResNet10, 5, 50).cuda())
    assert a.n_support == b.n_support == 25
    g = torch.Generator().manual_seed(6)
    feat = torch.randn(5, 65, 512, generator=g)
    with torch.no_grad():
        a.n_query = b.n_query = 15
        sa, sb = a.set_forward(feat, is_feature=True), b.set_forward(feat, is_feature=True)
    assert _rel(sb, sa) < 2e-5
    x = torch.randn(5, 66, 3, 224, 224, generator=g)
    for m in (a, b):
        opt = torch.optim.SGD(m.parameters(), lr=0.05)
        m.train_loop50(0, [(x, None)], opt)
    _compare_grads(a, b)
    for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert _rel(q, p) < 3e-2, k


def test_dampnet_full_branches_and_train_loop():
    """dampnet_full.DampNet drives the same GNN_nl through three branches (dampnet_full.py:97-294): plain
    (prototypes not initialised), recovery of the clean features (call_count even, gnn.train()) and recovery
    of numpy-corrupted features (call_count odd, gnn.eval(), fc frozen); then one train_loop_full step."""
    a, b = _build(lambda ns: ns.dampnet_full.DampNet(ns.backbone.ResNet10, 5, 5))
    g = torch.Generator().manual_seed(8)
    feats = [torch.randn(5 * 21, 512, generator=g) for _ in range(3)]
    protos = torch.randn(500, 512, generator=g)
    results = []
    for m in (a, b):
        np.random.seed(3)
        m.n_query = 16
        out = []
        loss = m.set_forward_loss(feats[0].cuda())            # branch 1
        loss.backward()
        out.append(float(loss.detach()))
        m.get_all_feat(protos)
        for f in feats[1:]:                                     # call_count 151 (odd: corrupted), 152 (even)
            loss = m.set_forward_loss(f.cuda())
            loss.backward()
            out.append(float(loss.detach()))
        results.append(out)
    for la, lb in zip(*results):
        assert abs(la - lb) < 1e-4 * max(1.0, abs(la)), results
    _compare_grads(a, b)
    x = torch.randn(5, 21, 3, 224, 224, generator=g)
    for m in (a, b):
        np.random.seed(4)
        m.zero_grad()
        opt = torch.optim.SGD(m.parameters(), lr=0.05)
        m.train_loop_full(0, [(x, None)], opt, final_epoch=10)
    for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert _rel(q, p) < 3e-2, k


def test_checkpoints_cross_load_and_deepcopy():
    """train.py:46-48 / finetune.py:507-540: a state_dict saved by the reference loads into the overlay model
    (and back) with strict key matching; copy.deepcopy (finetune.py:185) keeps working."""
    a, b = _build(lambda ns: ns.gnnnet.GnnNet(ns.backbone.ResNet10, 5, 5).cuda())
    with torch.no_grad():
        for p in a.parameters():
            p.add_(0.01)
    b.load_state_dict(a.state_dict(), strict=True)
    a.load_state_dict(copy.deepcopy(b).state_dict(), strict=True)
    feat = torch.randn(5, 20, 512)
    with torch.no_grad():
        a.n_query = b.n_query = 15
        assert _rel(b.set_forward(feat, is_feature=True), a.set_forward(feat, is_feature=True)) < 2e-5
