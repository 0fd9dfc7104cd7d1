"""Generate golden fixtures from the REFERENCE itself (build container only).

Run:  python tests/golden/make_golden.py          (needs /root/reference)

The reference holds no golden vectors for the GNN head, so
the oracle is pinned against outputs of the reference's own ``methods/gnn.py``
and ``methods/gnnnet.py`` / ``gnnnet_copy.py`` imported from /root/reference and
run on CPU in float32 and float64.  /root/reference does not travel to the GPU
box, so the outputs are committed as small ``.npz`` files next to this script.

Fixtures
--------
gnn_tiny.npz       GNN_nl(13, 16, 3), B=3, N=7: full state_dict, input, output
                   and every gradient, float64 (truth) + float32 output.
gnn_5w5s.npz       GNN_nl(133, 96, 5), B=16, N=30 (config C1 head shape): input,
                   seeded state_dict (stored in full), fp64 output and every gradient.
head_5w5s.npz      GnnNet head on features (fc -> graphs -> gnn -> select):
                   scores + loss for n_query 16 and the finetune.py is_feature path (15).
head_50c.npz       gnnnet_copy (compressed 50-shot, N=130) scores on features
                   (parameters: those of head_5w5s.npz).
sampler.npz        EpisodicBatchSampler / generate_perm class draws, support_label,
                   query labels.
gnn_5w20s.npz      GNN_nl(133, 96, 5), B=16, N=105 (the benchmarked 5-way 20-shot head shape): as
                   gnn_5w5s.npz (gradients and dx stored as float32).      [python make_golden.py 5w20s]
head_5w5s_grads.npz  gradients of the n_query=16 loss of head_5w5s.npz w.r.t. every fc.* / gnn.* parameter
                   and the features (float64 run of the reference, stored float32) plus the reference's own
                   float32-vs-float64 error per tensor.                    [python make_golden.py headgrads]

loader.npz         which IMAGES form each episode: index streams recorded from the reference's own loaders
                   (CropDisease_few_shot.SetDataManager2 with num_aug=2, seed 10, num_workers=0;
                   miniImageNet_few_shot.SetDataManager with aug=False, 12 workers) run over a synthetic image
                   folder whose constant-colour images encode (class, index).   [python make_golden.py loader]

Without arguments every fixture is regenerated; with arguments only the named groups
(base, 5w20s, headgrads, loader).
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.path.insert(0, REF)
    # gnnnet.py:40,69 hard-code .cuda(); identity shim for the CPU container
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    from methods import gnn as ref_gnn
    from methods import gnnnet as ref_gnnnet
    from methods import gnnnet_copy as ref_gnnnet_copy
    import backbone as ref_backbone
    return ref_gnn, ref_gnnnet, ref_gnnnet_copy, ref_backbone


def _perturb_bn(module, gen):
    """Move BN affine terms off (1,0) so gamma/beta wiring is exercised."""
    with torch.no_grad():
        for name, prm in module.named_parameters():
            if ".bn" in "." + name and name.endswith("weight") and prm.dim() == 1:
                prm.add_(0.25 * torch.randn(prm.shape, generator=gen, dtype=prm.dtype))
            elif ".bn" in "." + name and name.endswith("bias"):
                prm.add_(0.2 * torch.randn(prm.shape, generator=gen, dtype=prm.dtype))


def _run_gnn(ref_gnn, fin, nf, n_way, bsz, n, seed):
    """fp64 (truth) and fp32 runs of the reference GNN_nl on float32-representable
    parameters / inputs, so that everything stored is self-consistent."""
    torch.manual_seed(seed)
    net32 = ref_gnn.GNN_nl(fin, nf, n_way)
    gen = torch.Generator().manual_seed(seed + 1)
    _perturb_bn(net32, gen)
    x = torch.randn(bsz, n, fin, generator=gen)
    proj = torch.randn(bsz, n, n_way, generator=gen)
    net = ref_gnn.GNN_nl(fin, nf, n_way).double()
    net.load_state_dict({k: v.double() for k, v in net32.state_dict().items()})
    x64 = x.double().requires_grad_(True)
    out64 = net(x64)
    (out64 * proj.double()).sum().backward()
    rec = {"x": x.numpy(), "proj": proj.numpy(), "out64": out64.detach().numpy(),
           "dx64": x64.grad.numpy()}
    for k, v in net32.state_dict().items():
        rec["p." + k] = v.numpy()
    for k, v in net.named_parameters():
        rec["g." + k] = v.grad.numpy()
    x32 = x.clone().requires_grad_(True)
    out32 = net32(x32)
    (out32 * proj).sum().backward()
    rec["out32"] = out32.detach().numpy()
    rec["dx32"] = x32.grad.numpy()
    # the reference's own fp32 error vs fp64, the yardstick for the fp32-path tolerance
    for k, v in net32.named_parameters():
        g64 = rec["g." + k]
        den = np.linalg.norm(g64)
        rec["e32." + k] = np.float64(np.linalg.norm(v.grad.numpy() - g64) / den if den > 0 else 0.0)
    return rec


def _head_model(ref_gnnnet, ref_backbone):
    """The GnnNet of head_5w5s.npz, rebuilt deterministically (same seeds, same draws)."""
    torch.manual_seed(5)
    np.random.seed(10)
    m = ref_gnnnet.GnnNet(ref_backbone.ResNet10, n_way=5, n_support=5)
    gen = torch.Generator().manual_seed(6)
    _perturb_bn(m.gnn, gen)
    feat15 = torch.randn(5, 5 + 15, 512, generator=gen)
    feat16 = torch.randn(5, 5 + 16, 512, generator=gen)
    return m, feat15, feat16


def _head_loss16(m, feat16):
    z = m.fc(feat16.view(-1, 512)).view(5, -1, 128)
    z_stack = [torch.cat([z[:, :5], z[:, 5 + i:5 + i + 1]], dim=1).view(1, -1, 128) for i in range(16)]
    s16 = m.forward_gnn(z_stack)
    y = torch.from_numpy(np.repeat(range(5), 16))
    return s16, m.loss_fn(s16, y)


def make_5w20s(ref_gnn):
    rec = _run_gnn(ref_gnn, 133, 96, 5, 16, 105, seed=2020)
    for k in list(rec):
        if k.startswith("g.") or k in ("dx64", "dx32"):
            rec[k] = rec[k].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "gnn_5w20s.npz"), **rec)


def make_headgrads(ref_gnnnet, ref_backbone):
    m, _, feat16 = _head_model(ref_gnnnet, ref_backbone)
    stored = dict(np.load(os.path.join(HERE, "head_5w5s.npz")))
    f32 = feat16.clone().requires_grad_(True)
    s16, loss = _head_loss16(m, f32)
    assert np.array_equal(s16.detach().numpy(), stored["scores16"]), "head model was not rebuilt identically"
    loss.backward()
    g32 = {k: v.grad.clone() for k, v in m.named_parameters() if k.startswith(("fc.", "gnn."))}
    d32 = f32.grad.clone()
    m.zero_grad()
    m.support_label = m.support_label.double()
    m.fc.double()
    m.gnn.double()
    f64 = feat16.double().requires_grad_(True)
    _, loss64 = _head_loss16(m, f64)
    loss64.backward()
    rec = {"loss64": np.float64(loss64.item()), "dfeat": f64.grad.numpy().astype(np.float32)}
    den = np.linalg.norm(f64.grad.numpy())
    rec["e32.dfeat"] = np.float64(np.linalg.norm(d32.numpy() - f64.grad.numpy()) / den)
    for k, v in m.named_parameters():
        if not k.startswith(("fc.", "gnn.")):
            continue
        g64 = v.grad.numpy()
        rec["g." + k] = g64.astype(np.float32)
        den = np.linalg.norm(g64)
        rec["e32." + k] = np.float64(np.linalg.norm(g32[k].numpy() - g64) / den if den > 0 else 0.0)
    np.savez_compressed(os.path.join(HERE, "head_5w5s_grads.npz"), **rec)


def loader_image_size(cl, idx):
    """(w, h) of synthetic image `idx` of class `cl` (tests/test_host_logic.py uses the same rule)."""
    return 40 + (7 * idx + 3 * cl) % 50, 40 + (5 * idx + cl) % 50


def _synthetic_folder(root, class_sizes):
    from PIL import Image
    for cl, n in enumerate(class_sizes):
        d = os.path.join(root, "c%03d" % cl)
        os.makedirs(d)
        for idx in range(n):
            Image.new("RGB", loader_image_size(cl, idx), (cl, idx, 128)).save(os.path.join(d, "i%03d.png" % idx))


def _decode(x):
    """[n_way, batch, 3, H, W] un-augmented views -> (class, index) of every image (colour = (cl, idx, 128))."""
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 1, 3)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 1, 3)
    px = ((x[:, :, :, 0, 0] * std + mean) * 255.0).round().long()
    assert (px[:, :, 2] == 128).all()
    return px[:, :, 0].numpy(), px[:, :, 1].numpy()


def make_loader():
    import tempfile
    import torchvision.transforms as T
    # the reference pins torchvision 0.8.2, where these are (deprecated) aliases; later releases dropped them
    if not hasattr(T, "Scale"):
        T.Scale = T.Resize
    if not hasattr(T, "RandomSizedCrop"):
        T.RandomSizedCrop = T.RandomResizedCrop
    sys.path.insert(0, REF)
    from datasets import CropDisease_few_shot as crop
    from datasets import miniImageNet_few_shot as mini
    rec = {}
    with tempfile.TemporaryDirectory() as tmp:
        crop_sizes = [24 + (cl * 5) % 17 for cl in range(38)]
        _synthetic_folder(os.path.join(tmp, "crop", "dataset", "train"), crop_sizes)
        crop.CropDisease_path = os.path.join(tmp, "crop")
        mgr = crop.SetDataManager2(64, n_eposide=6, n_query=15, n_way=5, n_support=5)
        loader = mgr.get_data_loader(num_aug=2)
        classes, indices = [], []
        for elem in loader:
            assert len(elem) == 4 and torch.equal(elem[0][0], elem[1][0])        # finetune.py:606
            cl, idx = _decode(elem[0][0])
            assert (cl == cl[:, :1]).all() and (cl[:, 0] == elem[0][1][:, 0].numpy()).all()
            classes.append(cl[:, 0])
            indices.append(idx)
        rec["crop_class_sizes"] = np.array(crop_sizes)
        rec["crop_classes"] = np.stack(classes)
        rec["crop_indices"] = np.stack(indices)
        rec["crop_rng_after"] = torch.rand(4).numpy()          # the global stream position after the 6 episodes
        mini_sizes = [22 + cl % 9 for cl in range(64)]
        _synthetic_folder(os.path.join(tmp, "mini"), mini_sizes)
        mini.miniImageNet_path = os.path.join(tmp, "mini")
        torch.manual_seed(10)
        loader = mini.SetDataManager(64, n_query=16, n_way=5, n_support=5, n_eposide=8).get_data_loader(aug=False)
        classes, indices = [], []
        for x, y in loader:
            cl, idx = _decode(x)
            assert (cl == y.numpy()).all()
            classes.append(cl[:, 0])
            indices.append(idx)
        rec["mini_class_sizes"] = np.array(mini_sizes)
        rec["mini_classes"] = np.stack(classes)
        rec["mini_indices"] = np.stack(indices)
        rec["mini_rng_after"] = torch.rand(4).numpy()
    np.savez_compressed(os.path.join(HERE, "loader.npz"), **rec)


def main():
    ref_gnn, ref_gnnnet, ref_gnnnet_copy, ref_backbone = _import_reference()
    torch.set_num_threads(8)
    groups = set(sys.argv[1:]) or {"base", "5w20s", "headgrads", "loader"}
    if "loader" in groups:
        make_loader()
    if "5w20s" in groups:
        make_5w20s(ref_gnn)
    if "headgrads" in groups:
        make_headgrads(ref_gnnnet, ref_backbone)
    if "base" in groups:
        make_base(ref_gnn, ref_gnnnet, ref_gnnnet_copy, ref_backbone)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


def make_base(ref_gnn, ref_gnnnet, ref_gnnnet_copy, ref_backbone):
    rec = _run_gnn(ref_gnn, 13, 16, 3, 3, 7, seed=1234)
    np.savez_compressed(os.path.join(HERE, "gnn_tiny.npz"), **rec)

    rec = _run_gnn(ref_gnn, 133, 96, 5, 16, 30, seed=77)
    for k in list(rec):
        if k.startswith("g."):
            rec[k] = rec[k].astype(np.float32)      # halve the file; 1e-7 rel is ample
    np.savez_compressed(os.path.join(HERE, "gnn_5w5s.npz"), **rec)

    # ---- GnnNet head on features -------------------------------------------------
    torch.manual_seed(5)
    np.random.seed(10)
    m = ref_gnnnet.GnnNet(ref_backbone.ResNet10, n_way=5, n_support=5)
    gen = torch.Generator().manual_seed(6)
    _perturb_bn(m.gnn, gen)
    head = {}
    for k, v in m.state_dict().items():
        if k.startswith("fc.") or k.startswith("gnn."):
            head["p." + k] = v.numpy()
    head["support_label"] = m.support_label.numpy()
    feat15 = torch.randn(5, 5 + 15, 512, generator=gen)
    m.n_query = 15
    s15 = m.set_forward(feat15, is_feature=True)           # finetune.py:316 path
    head["feat15"] = feat15.numpy()
    head["scores15"] = s15.detach().numpy()
    # training shape (n_query 16): drive fc + forward_gnn directly on features
    feat16 = torch.randn(5, 5 + 16, 512, generator=gen)
    m.n_query = 16
    z = m.fc(feat16.view(-1, 512)).view(5, -1, 128)
    z_stack = [torch.cat([z[:, :5], z[:, 5 + i:5 + i + 1]], dim=1).view(1, -1, 128) for i in range(16)]
    s16 = m.forward_gnn(z_stack)
    y = torch.from_numpy(np.repeat(range(5), 16))
    loss16 = m.loss_fn(s16, y)
    head["feat16"] = feat16.numpy()
    head["scores16"] = s16.detach().numpy()
    head["loss16"] = np.float64(loss16.item())
    head["y16"] = y.numpy()
    np.savez_compressed(os.path.join(HERE, "head_5w5s.npz"), **head)

    # ---- compressed 50-shot ------------------------------------------------------
    torch.manual_seed(9)
    mc = ref_gnnnet_copy.GnnNet(ref_backbone.ResNet10, n_way=5, n_support=50)
    # same fc/gnn parameters as head_5w5s.npz (identical architecture) -- not stored twice
    sd = mc.state_dict()
    for k, v in m.state_dict().items():
        if k.startswith("fc.") or k.startswith("gnn."):
            sd[k] = v.clone()
    mc.load_state_dict(sd)
    c = {}
    c["support_label"] = mc.support_label.numpy()
    c["n_support_eff"] = np.int64(mc.n_support)
    feat = torch.randn(5, 50 + 15, 512, generator=gen)
    mc.n_query = 15
    with torch.no_grad():
        sc = mc.set_forward(feat, is_feature=True)
    c["feat"] = feat.numpy()
    c["scores"] = sc.numpy()
    np.savez_compressed(os.path.join(HERE, "head_50c.npz"), **c)

    # ---- sampling / label indexing ------------------------------------------------
    s = {}
    sys.path.insert(0, REF)
    from datasets import miniImageNet_few_shot as mini
    torch.manual_seed(10)
    smp = mini.EpisodicBatchSampler(64, 5, 100)
    s["mini_train"] = torch.stack(list(iter(smp))).numpy()
    for name, n_cls, seed in (("CropDisease", 38, 10), ("EuroSAT", 10, 7), ("ISIC", 7, 10), ("Chest", 7, 11)):
        # datasets/<name>_few_shot.py: SetDataset2 re-seeds everything to `seed`, then
        # EpisodicBatchSampler2.generate_perm draws 600 class permutations
        torch.manual_seed(seed)
        np.random.seed(seed)
        from datasets import CropDisease_few_shot as crop
        smp2 = crop.EpisodicBatchSampler2(n_cls, 5, 600)
        s["perm_" + name] = torch.stack(smp2.generate_perm()).numpy()
    for n_way, n_sup in ((5, 5), (5, 20), (5, 25)):
        mm = ref_gnnnet.GnnNet(ref_backbone.ResNet10, n_way=n_way, n_support=n_sup)
        s[f"support_label_{n_way}_{n_sup}"] = mm.support_label.numpy()
    s["y_query_5_16"] = np.repeat(range(5), 16)
    s["y_query_5_15"] = np.repeat(range(5), 15)
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **s)


if __name__ == "__main__":
    main()
