"""Episode sampling and label indexing, bit-exact with the reference.

The reference draws every class subset from the *global* torch RNG:

* meta-training: ``EpisodicBatchSampler.__iter__`` yields ``torch.randperm(n_classes)[:n_way]``
  per episode (datasets/miniImageNet_few_shot.py:105-107);
* target domains: the dataset constructor re-seeds torch / numpy / random
  (seed 10 CropDisease and ISIC, 7 EuroSAT, 11 ChestX -- datasets/CropDisease_few_shot.py:100-107,
  EuroSAT_few_shot.py:97, Chest_few_shot.py:183), then ``EpisodicBatchSampler2.generate_perm``
  draws all 600 subsets up front (CropDisease_few_shot.py:200-205).

Under episode sharding every rank calls these with the same seed, gets the identical
stream, and keeps the rows it owns (``parallel.owned_episodes``).
"""
from __future__ import annotations

import numpy as np
import torch

DOMAIN_SEEDS = {"CropDisease": 10, "EuroSAT": 7, "ISIC": 10, "ChestX": 11}
DOMAIN_CLASSES = {"CropDisease": 38, "EuroSAT": 10, "ISIC": 7, "ChestX": 7}


def draw_class_subsets(n_classes: int, n_way: int, n_episodes: int) -> torch.Tensor:
    """``n_episodes`` draws of ``torch.randperm(n_classes)[:n_way]`` from the global RNG."""
    return torch.stack([torch.randperm(n_classes)[:n_way] for _ in range(n_episodes)])


def target_domain_perms(domain: str, n_way: int = 5, n_episodes: int = 600) -> torch.Tensor:
    """The fixed class permutations of a target-domain evaluation run."""
    seed = DOMAIN_SEEDS[domain]
    torch.manual_seed(seed)
    np.random.seed(seed)
    return draw_class_subsets(DOMAIN_CLASSES[domain], n_way, n_episodes)


def train_epoch_subsets(n_classes: int = 64, n_way: int = 5, n_episodes: int = 100) -> torch.Tensor:
    """Class subsets of one meta-training epoch (continues the global RNG stream)."""
    return draw_class_subsets(n_classes, n_way, n_episodes)


# ------------------------------------------------------------------------------------------
# Image-index streams of the reference's episodic loaders (which images form each episode)
# ------------------------------------------------------------------------------------------
# The class subsets above are only half of row S.  WHICH images of a class enter an episode is decided by
# ``SetDataset.__getitem__`` / ``SetDataset2.__getitem__`` = ``next(iter(DataLoader(sub_dataset, shuffle=True,
# batch_size=n_support+n_query)))`` (datasets/miniImageNet_few_shot.py:73-74, CropDisease_few_shot.py:127-128):
# every call builds a fresh loader iterator, which draws a base seed and a RandomSampler seed from the GLOBAL
# torch RNG and then a ``randperm(len(class))``; on the test side the augmentation transforms of every loaded
# image (RandomResizedCrop, ImageJitter, Random{Horizontal,Vertical}Flip -- TransformLoader2,
# CropDisease_few_shot.py:244-266) consume the same global stream in between.  The replay below drives
# torch's OWN DataLoader / RandomSampler machinery over index-only datasets, so the stream consumption is
# the library's, not a transcription of one torch version's, and consumes the transforms' random numbers
# through torchvision's own ``get_params`` -- without touching an image.  tests/test_host_logic.py checks it
# against index streams recorded from the reference's loaders over a synthetic image folder
# (tests/golden/loader.npz, tests/golden/make_golden.py loader).

class _AugRng:
    """Consumes exactly the random numbers TransformLoader2's augmented pipeline draws for one image of
    size (w, h): RandomResizedCrop(image_size, scale).get_params, ImageJitter (one torch.rand(3)),
    RandomHorizontalFlip and (test side) RandomVerticalFlip (one torch.rand(1) each)."""

    def __init__(self, scale=(0.5, 0.9), vertical_flip=True, jitter_terms=3):
        self.scale = scale
        self.ratio = (3.0 / 4.0, 4.0 / 3.0)
        self.vertical_flip = vertical_flip
        self.jitter_terms = jitter_terms

    def __call__(self, w: int, h: int) -> None:
        from torchvision import transforms
        transforms.RandomResizedCrop.get_params(torch.empty(3, h, w), list(self.scale), list(self.ratio))
        torch.rand(self.jitter_terms)
        torch.rand(1)
        if self.vertical_flip:
            torch.rand(1)


class _ClassIndices(torch.utils.data.Dataset):
    """Stand-in for SubDataset / SubDataset2: item i is its own index; loading it draws what the reference's
    transforms would draw (``num_aug`` augmented views after the un-augmented ones)."""

    def __init__(self, cl, size, num_aug=0, image_size_fn=None, aug_rng=None):
        self.cl, self.size, self.num_aug = cl, size, num_aug
        self.image_size_fn, self.aug_rng = image_size_fn, aug_rng

    def __len__(self):
        return self.size

    def __getitem__(self, i):
        if self.num_aug:
            w, h = self.image_size_fn(self.cl, int(i))
            for _ in range(self.num_aug):
                self.aug_rng(w, h)
        return int(i)


class _SetIndices(torch.utils.data.Dataset):
    """Stand-in for SetDataset / SetDataset2: item cl = the first batch of a fresh shuffled loader over class cl."""

    def __init__(self, class_sizes, batch_size, num_aug=0, image_size_fn=None, aug_rng=None):
        self.loaders = [torch.utils.data.DataLoader(_ClassIndices(cl, n, num_aug, image_size_fn, aug_rng),
                                                    batch_size=batch_size, shuffle=True, num_workers=0,
                                                    pin_memory=False, collate_fn=list)
                        for cl, n in enumerate(class_sizes)]

    def __len__(self):
        return len(self.loaders)

    def __getitem__(self, cl):
        return int(cl), next(iter(self.loaders[int(cl)]))


def target_domain_episode_indices(class_sizes, seed: int, n_way: int = 5, batch_size: int = 20,
                                  n_episodes: int = 600, num_aug: int = 0, image_size_fn=None,
                                  aug_scale=(0.5, 0.9)):
    """Replay of ``SetDataManager2(...).get_data_loader(num_aug)`` + iteration (finetune.py:572-573,634):
    yields, per episode, ``(classes [n_way], indices [n_way][batch_size])`` -- ``indices[c][k]`` is the position
    inside class ``classes[c]``'s file list of the k-th image (supports first, then queries).
    ``image_size_fn(cl, idx) -> (w, h)`` is needed when ``num_aug > 0``."""
    import random
    torch.manual_seed(seed)                      # SetDataset2.__init__, CropDisease_few_shot.py:100-107
    np.random.seed(seed)
    random.seed(seed)
    data = _SetIndices(class_sizes, batch_size, num_aug, image_size_fn, _AugRng(aug_scale, True))
    perms = [torch.randperm(len(class_sizes))[:n_way] for _ in range(n_episodes)]       # generate_perm
    loader = torch.utils.data.DataLoader(data, batch_sampler=perms, num_workers=0, pin_memory=False,
                                         collate_fn=list)
    for ep in loader:
        yield [c for c, _ in ep], [idx for _, idx in ep]


def train_episode_indices(class_sizes, n_way: int = 5, batch_size: int = 21, n_episodes: int = 100,
                          num_workers: int = 12):
    """Replay of one epoch of ``miniImageNet_few_shot.SetDataManager(...).get_data_loader(aug=False)``
    (train.py:118-119; 12 worker processes, datasets/miniImageNet_few_shot.py:180): continues the global
    torch RNG exactly as the reference's epoch does (one base-seed draw, then one randperm per episode in
    the main process; the per-class shuffles happen in the workers, seeded base_seed + worker_id)."""
    data = _SetIndices(class_sizes, batch_size)

    class _Sampler:                                        # EpisodicBatchSampler, miniImageNet_few_shot.py:96-107
        def __len__(self):
            return n_episodes

        def __iter__(self):
            for _ in range(n_episodes):
                yield torch.randperm(len(class_sizes))[:n_way]

    loader = torch.utils.data.DataLoader(data, batch_sampler=_Sampler(), num_workers=num_workers,
                                         pin_memory=False, collate_fn=list)
    for ep in loader:
        yield [c for c, _ in ep], [idx for _, idx in ep]


def inner_loop_batches(n_items: int, batch_size: int, epochs: int):
    """Mini-batch index streams of the fine-tuning inner loops (finetune.py:271-284, gnnnet.py:153-161):
    one ``np.random.permutation(n_items)`` per epoch from the GLOBAL numpy RNG (seeded once,
    finetune.py:425 / train.py:70), cut into consecutive batches."""
    for _ in range(epochs):
        rand_id = np.random.permutation(n_items)
        for j in range(0, n_items, batch_size):
            yield rand_id[j:min(j + batch_size, n_items)]


def skip_inner_loop(n_items: int, epochs: int) -> None:
    """Advance the numpy stream past an episode this rank does not own (episode sharding: every rank
    replays every episode's draws)."""
    for _ in range(epochs):
        np.random.permutation(n_items)
