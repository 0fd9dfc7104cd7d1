"""CPU tests of the host-side mirror: label indexing, graph assembly, sampling streams, the
methods/ overlay, and the rule that the product never imports the oracle."""
import importlib
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

import mft_b200
from mft_b200 import episode, sampling
from oracle import gnn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_support_label_and_query_labels_match_reference(golden_dir):
    s = np.load(os.path.join(golden_dir, "sampler.npz"))
    for n_way, n_sup in ((5, 5), (5, 20), (5, 25)):
        assert np.array_equal(episode.support_label(n_way, n_sup).numpy(), s[f"support_label_{n_way}_{n_sup}"])
    assert np.array_equal(episode.query_labels(5, 16).numpy(), s["y_query_5_16"])
    assert np.array_equal(episode.query_labels(5, 15).numpy(), s["y_query_5_15"])


def test_sampling_streams_bit_exact(golden_dir):
    s = np.load(os.path.join(golden_dir, "sampler.npz"))
    for dom, key in (("CropDisease", "perm_CropDisease"), ("EuroSAT", "perm_EuroSAT"), ("ISIC", "perm_ISIC"),
                     ("ChestX", "perm_Chest")):
        assert np.array_equal(sampling.target_domain_perms(dom).numpy(), s[key]), dom
    torch.manual_seed(10)
    assert np.array_equal(sampling.train_epoch_subsets(64, 5, 100).numpy(), s["mini_train"])


@pytest.mark.parametrize("n_support,compress", [(5, False), (20, False), (50, True)])
def test_build_graphs_and_select_scores_match_oracle(n_support, compress):
    n_way, n_query, d = 5, 4, 7
    g = torch.Generator().manual_seed(n_support)
    z = torch.randn(n_way, n_support + n_query, d, generator=g)
    k = round(n_support / 2) if compress else n_support
    lab = episode.support_label(n_way, k)
    nodes = episode.build_graphs(z, lab, n_way, n_support, n_query, compress)
    ref = O.build_graphs(z, n_way, n_support, n_query, compress)
    assert torch.equal(nodes, ref)
    out = torch.randn(n_query, n_way * (k + 1), n_way, generator=g)
    assert torch.equal(episode.select_scores(out, n_way, k, n_query), O.select_scores(out, n_way, k, n_query))


def test_gnn_head_state_dict_matches_reference_names(golden_dir):
    rec = np.load(os.path.join(golden_dir, "head_5w5s.npz"))
    names = [k[2:] for k in rec.files if k.startswith("p.")]
    head = mft_b200.GnnHead(5, 5)
    sd = head.state_dict()
    assert list(sd.keys()) == names
    for k in names:
        assert tuple(sd[k].shape) == rec["p." + k].shape
    head.load_state_dict({k: torch.from_numpy(rec["p." + k]) for k in names})   # reference checkpoint loads


def test_module_surface_and_maml_flags():
    from mft_b200.gnn import GNN_nl, Gconv, Wcompute
    assert Wcompute.maml is False and Gconv.maml is False
    net = GNN_nl(133, 96, 5)
    assert (net.input_features, net.nf, net.num_layers) == (133, 96, 2)
    assert [n for n, _ in net.named_children()] == ["layer_w0", "layer_l0", "layer_w1", "layer_l1", "w_comp_last",
                                                     "layer_last"]
    assert sum(p.numel() for p in net.parameters()) == 335994
    Wcompute.maml = Gconv.maml = True                                   # train.py:147-148
    try:
        m = GNN_nl(133, 96, 5)
        assert list(m.state_dict().keys()) == list(net.state_dict().keys())
        assert getattr(m.layer_w0.conv2d_1.weight, "fast", 0) is None
    finally:
        Wcompute.maml = Gconv.maml = False
    with pytest.raises(NotImplementedError):
        Wcompute(8, 4, activation="sigmoid").adjacency(torch.zeros(1, 2, 8))


def test_methods_overlay_resolves_gnn_here_and_everything_else_in_reference(tmp_path):
    """`from methods.gnn import GNN_nl` picks this repo's module, `methods.<other>` the reference's."""
    ref = tmp_path / "fake_reference" / "methods"
    ref.mkdir(parents=True)
    (ref / "__init__.py").write_text("")
    (ref / "gnn.py").write_text("MARK = 'reference gnn'\n")
    (ref / "gnnnet.py").write_text("from methods.gnn import GNN_nl\nMARK = 'reference gnnnet'\n")
    code = textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {str(ref.parent)!r}); sys.path.insert(0, {ROOT!r})
        import methods.gnn, methods.gnnnet, mft_b200
        assert methods.gnn.GNN_nl is mft_b200.GNN_nl
        assert methods.gnnnet.MARK == 'reference gnnnet' and methods.gnnnet.GNN_nl is mft_b200.GNN_nl
        print('ok')
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_product_never_touches_the_oracle_or_the_reference():
    pkg = os.path.join(ROOT, "meta-fine-tuning_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle|/root/reference", text, flags=re.M):
                    bad.append(f)
    assert not bad, bad


def test_missing_library_fails_loudly(monkeypatch):
    from mft_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_LIB_NAME", "libmft_gnn_missing.so")
    with pytest.raises(_lib.LibraryMissing, match="no other compute path"):
        _lib.load_library()


# ------------------------------------------------------------------------------------------
# row S, second half: WHICH images form each episode (index streams of the reference's loaders)
# ------------------------------------------------------------------------------------------

def _loader_image_size(cl, idx):       # the rule of tests/golden/make_golden.py loader_image_size
    return 40 + (7 * idx + 3 * cl) % 50, 40 + (5 * idx + cl) % 50


def test_target_domain_image_indices_match_the_reference_loader(golden_dir):
    """SetDataManager2(n_query=15, n_support=5).get_data_loader(num_aug=2) of CropDisease_few_shot.py run
    over a synthetic image folder (tests/golden/loader.npz): class subsets, the images of every class (each
    a fresh shuffled DataLoader whose seeds come from the global stream) AND the amount of randomness the
    augmentation transforms consume in between are reproduced bit for bit -- the global RNG is at the same
    position after the six episodes."""
    from mft_b200 import sampling
    rec = dict(np.load(os.path.join(golden_dir, "loader.npz")))
    eps = list(sampling.target_domain_episode_indices(rec["crop_class_sizes"].tolist(), seed=10, n_way=5,
                                                      batch_size=20, n_episodes=6, num_aug=2,
                                                      image_size_fn=_loader_image_size))
    assert np.array_equal(np.array([e[0] for e in eps]), rec["crop_classes"])
    assert np.array_equal(np.array([e[1] for e in eps]), rec["crop_indices"])
    assert np.array_equal(torch.rand(4).numpy(), rec["crop_rng_after"])
    # without modelling the transforms' consumption the very first episode already diverges
    eps0 = list(sampling.target_domain_episode_indices(rec["crop_class_sizes"].tolist(), seed=10, n_way=5,
                                                       batch_size=20, n_episodes=2, num_aug=0))
    assert np.array_equal(np.array(eps0[0][0]), rec["crop_classes"][0])
    assert not np.array_equal(np.array([e[1] for e in eps0]), rec["crop_indices"][:2])


def test_training_image_indices_match_the_reference_loader(golden_dir):
    """miniImageNet_few_shot.SetDataManager(...).get_data_loader(aug=False) (12 worker processes): class
    subsets from the main process's stream, per-class shuffles from the workers' (base_seed + worker_id)."""
    from mft_b200 import sampling
    rec = dict(np.load(os.path.join(golden_dir, "loader.npz")))
    torch.manual_seed(10)
    eps = list(sampling.train_episode_indices(rec["mini_class_sizes"].tolist(), n_way=5, batch_size=21,
                                              n_episodes=8, num_workers=12))
    assert np.array_equal(np.array([e[0] for e in eps]), rec["mini_classes"])
    assert np.array_equal(np.array([e[1] for e in eps]), rec["mini_indices"])
    assert np.array_equal(torch.rand(4).numpy(), rec["mini_rng_after"])


def test_inner_loop_batches_follow_the_numpy_stream():
    """finetune.py:271-284 / gnnnet.py:153-161: np.random.permutation per epoch, consecutive batches; a rank
    that skips an episode it does not own leaves the stream where the owner's replay leaves it."""
    from mft_b200 import sampling
    np.random.seed(10)
    want = []
    for _ in range(3):
        rand_id = np.random.permutation(25)
        want += [rand_id[j:min(j + 4, 25)] for j in range(0, 25, 4)]
    after = np.random.permutation(5)
    np.random.seed(10)
    got = list(sampling.inner_loop_batches(25, 4, 3))
    assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))
    assert np.array_equal(np.random.permutation(5), after)
    np.random.seed(10)
    sampling.skip_inner_loop(25, 3)
    assert np.array_equal(np.random.permutation(5), after)
