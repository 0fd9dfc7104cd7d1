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
