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

__all__ = ["owner_of", "owned_episodes", "allreduce_mean_grads", "gather_episode_results", "broadcast_parameters"]


def owner_of(episode: int, world: int) -> int:
    return episode % world


def owned_episodes(n_episodes: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_episodes, world))


def allreduce_mean_grads(params: Iterable[torch.nn.Parameter], world: int | None = None) -> int:
    """Average ``.grad`` over ranks with ONE collective on a flat buffer; returns the element count."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    n = sum(g.numel() for g in grads)
    if world == 1:
        return n
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return n


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


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
