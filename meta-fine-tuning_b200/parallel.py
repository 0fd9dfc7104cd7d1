"""Episode-level data parallelism (new; the reference is single-process).

Episodes are independent units, so the path shards by episode with one process per GPU:

* evaluation / test-time fine-tuning: episode ``e`` belongs to rank ``e % world``; every
  rank replays the same seeded sampler stream and skips the episodes it does not own, so
  sampling stays bit-exact with the single-process run; the only exchange is one gather of
  the per-episode accuracies at the end;
* meta-training: one episode per rank per step; gradients are averaged with one flat
  all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests).  BatchNorm
  statistics are deliberately NOT synchronised: each episode normalises itself, as in the
  reference.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist

__all__ = ["owner_of", "owned_episodes", "allreduce_mean_grads", "gather_episode_results", "broadcast_parameters",
           "average_parameters", "accuracy_summary"]


def owner_of(episode: int, world: int) -> int:
    return episode % world


def owned_episodes(n_episodes: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_episodes, world))


def _mean_allreduce_(flat: torch.Tensor, world: int) -> None:
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)      # one kernel, no separate division
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)


def allreduce_mean_grads(params: Iterable[torch.nn.Parameter], world: int | None = None) -> int:
    """Average ``.grad`` over ranks; returns the element count.

    The gradients GNN_nl's backward returns are views of ONE flat allocation (gnn._alloc_like_flat):
    those are reduced in place with a single collective and no copy.  Gradients that live in
    storages of their own (e.g. the ``fc`` layer of GnnHead) are packed into one more flat buffer,
    reduced, and copied back with one fused launch."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    n = sum(g.numel() for g in grads)
    if world == 1:
        return n
    by_storage = {}
    for g in grads:
        by_storage.setdefault(g.untyped_storage().data_ptr(), []).append(g)
    loose: List[torch.Tensor] = []
    for group in by_storage.values():
        if len(group) < 2 or not all(g.is_contiguous() for g in group):
            loose.extend(group)
            continue
        lo = min(g.storage_offset() for g in group)
        hi = max(g.storage_offset() + g.numel() for g in group)
        g0 = group[0]
        span = torch.as_strided(g0, (hi - lo,), (1,), lo)  # covers every view (+ alignment padding)
        _mean_allreduce_(span, world)
    if loose:
        flat = torch.cat([g.reshape(-1) for g in loose])
        _mean_allreduce_(flat, world)
        outs, off = [], 0
        for g in loose:
            outs.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        torch._foreach_copy_(loose, outs)
    return n


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


def average_parameters(params: Iterable[torch.nn.Parameter], world: int | None = None) -> int:
    """Re-synchronise replicas whose parameters moved rank-locally: mean over ranks, in place.

    The first-order-MAML variants (gnnnet.py:90-103 ``MAML_update``; dampnet_full.py) rewind the last
    backbone stage by a delta each rank computes from ITS episode's inner loop, so after the rewind the
    replicas differ in exactly those tensors; averaging them is the data-parallel
    counterpart of the reference's single rewind.  One flat all-reduce; returns the element count."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    ps = [p for p in params]
    n = sum(p.numel() for p in ps)
    if world == 1 or not ps:
        return n
    flat = torch.cat([p.detach().reshape(-1) for p in ps])
    _mean_allreduce_(flat, world)
    off = 0
    with torch.no_grad():
        for p in ps:
            p.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
    return n


def gather_episode_results(local: Sequence[float], n_episodes: int, rank: int, world: int,
                           device="cpu") -> torch.Tensor:
    """Per-episode results of the episodes this rank owns -> full [n_episodes] vector on every rank."""
    out = torch.zeros(n_episodes, dtype=torch.float64, device=device)
    idx = owned_episodes(n_episodes, rank, world)
    assert len(idx) == len(local)
    if idx:
        out[torch.tensor(idx, device=device)] = torch.as_tensor(list(local), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def accuracy_summary(acc_all) -> tuple[float, float]:
    """(mean, 95 % half-width) of the per-episode accuracies exactly as the reference prints them
    (finetune.py:672-676, meta_template.py:146-149): population std, 1.96 * std / sqrt(n).  Every rank gets
    the same numbers from the vector ``gather_episode_results`` returns."""
    a = torch.as_tensor(acc_all, dtype=torch.float64)
    n = a.numel()
    mean = float(a.mean())
    std = float(a.std(unbiased=False))
    return mean, 1.96 * std / n ** 0.5
