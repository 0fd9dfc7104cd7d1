"""world_size-2 gloo tests (CPU) of the episode-parallel host logic: gradient averaging equals
the single-process gradient of the mean loss over the same episodes, episode ownership, and the
final gather of per-episode results.  The per-episode compute here is the CPU oracle (tests may
use it); on the GPU box the same helpers wrap the CUDA modules under NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mft_b200 import parallel
from oracle import gnn_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _episode(seed, fin=9, n=5, bsz=2, n_way=3):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(bsz, n, fin, generator=g), torch.randn(bsz, n, n_way, generator=g)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    p = {k: v.clone().requires_grad_(True) for k, v in O.random_params(9, 8, 3, 0, torch.float64).items()}
    x, proj = _episode(100 + rank)                         # one episode per rank per step
    loss = (O.gnn_nl(x.double(), p) * proj.double()).sum()
    loss.backward()
    params = [torch.nn.Parameter(v.detach()) for v in p.values()]
    # the layout the CUDA backward produces: all but the last two gradients are views of one flat
    # allocation with 4-element alignment padding (gnn._alloc_like_flat); the rest own their storage
    vals = list(p.values())
    offs, total = [], 0
    for v in vals[:-2]:
        offs.append(total)
        total += (v.numel() + 3) & ~3
    flat = torch.full((total,), float("nan"), dtype=torch.float64)
    for prm, v, o in zip(params[:-2], vals[:-2], offs):
        view = flat[o:o + v.numel()].view(v.shape)
        view.copy_(v.grad)
        prm.grad = view
    flat[torch.isnan(flat)] = 0.0          # padding
    for prm, v in zip(params[-2:], vals[-2:]):
        prm.grad = v.grad.clone()
    n = parallel.allreduce_mean_grads(params, world)
    # first-order-MAML re-synchronisation: rank-local rewinds, then the mean over ranks
    moved = [torch.nn.Parameter(torch.full((3, 2), float(rank + 1))), torch.nn.Parameter(torch.arange(5.0) * (rank + 1))]
    n_avg = parallel.average_parameters(moved, world)
    owned = parallel.owned_episodes(7, rank, world)
    res = parallel.gather_episode_results([float(e) * 10 for e in owned], 7, rank, world)
    if rank == 0:
        np.savez(os.path.join(out_dir, "r0.npz"), n=n, res=res.numpy(), n_avg=n_avg,
                 avg0=moved[0].detach().numpy(), avg1=moved[1].detach().numpy(),
                 **{f"g{i}": prm.grad.numpy() for i, prm in enumerate(params)})
    dist.barrier()
    dist.destroy_process_group()


def test_dp_gradients_equal_single_process_mean(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "r0.npz"))
    p = {k: v.clone().requires_grad_(True) for k, v in O.random_params(9, 8, 3, 0, torch.float64).items()}
    total = 0
    for rank in range(world):
        x, proj = _episode(100 + rank)
        total = total + (O.gnn_nl(x.double(), p) * proj.double()).sum()
    (total / world).backward()
    assert int(got["n"]) == sum(v.numel() for v in p.values())
    for i, v in enumerate(p.values()):
        assert np.allclose(got[f"g{i}"], v.grad.numpy(), rtol=1e-10, atol=1e-12)
    assert np.array_equal(got["res"], np.arange(7) * 10.0)
    assert int(got["n_avg"]) == 11
    assert np.allclose(got["avg0"], 1.5) and np.allclose(got["avg1"], np.arange(5.0) * 1.5)


def test_episode_ownership_partitions_everything():
    for world in (1, 2, 4, 8):
        seen = sorted(e for r in range(world) for e in parallel.owned_episodes(600, r, world))
        assert seen == list(range(600))
        assert all(parallel.owner_of(e, world) == e % world for e in range(0, 600, 37))


def test_accuracy_summary_is_the_reference_formula():
    rng = np.random.default_rng(0)
    acc = rng.uniform(20, 100, size=600)
    mean, half = parallel.accuracy_summary(acc)
    assert abs(mean - np.mean(acc)) < 1e-12
    assert abs(half - 1.96 * np.std(acc) / np.sqrt(600)) < 1e-12      # finetune.py:672-676
