"""CUDA-graph replay of an episode step (new; the reference launches ~200 kernels per step eagerly).

One 5w20s head step is ~125 library launches plus a few dozen PyTorch ones; at that count the
gaps between dependent launches cost several hundred microseconds.  ``GraphedStep`` captures
``loss = fn(*static_inputs); loss.backward()`` once (after eager warm-up on a side stream, as
torch.cuda.graphs requires) and replays it: inputs are copied into the static tensors, the loss
and the parameter ``.grad`` tensors live at fixed addresses.

Everything the library enqueues during a step is capture-safe: kernels, cudaMemsetAsync /
cudaMemcpy2DAsync on the caller's stream, and no allocation or synchronisation of its own
(include/mft_gnn.h); tensor maps are encoded on the host from addresses that are stable across
replays because the blobs come from the graph's private pool.
"""
from __future__ import annotations

from typing import Callable, Iterable, Sequence

import torch

__all__ = ["GraphedStep"]


class GraphedStep:
    def __init__(self, fn: Callable[..., torch.Tensor], example_inputs: Sequence[torch.Tensor],
                 params: Iterable[torch.nn.Parameter], warmup: int = 3, inputs_require_grad: bool = True):
        self.fn = fn
        self.params = [p for p in params]
        self.inputs_require_grad = inputs_require_grad
        self.static_inputs = [t.detach().clone() for t in example_inputs]
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for p in self.params:
            p.grad = None
        with torch.cuda.graph(self.graph):
            self.static_loss = self._eager()
        torch.cuda.synchronize()
        # the graph writes gradients at these addresses on every replay
        self.static_grads = [p.grad for p in self.params]

    def _eager(self) -> torch.Tensor:
        for p in self.params:
            p.grad = None
        ins = []
        for t in self.static_inputs:
            if self.inputs_require_grad and t.is_floating_point():
                t.grad = None
                t.requires_grad_(True)
            ins.append(t)
        loss = self.fn(*ins)
        loss.backward()
        return loss.detach()

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        """Copy ``inputs`` into the static tensors (host tensors: asynchronously), replay, return
        the static loss tensor (valid until the next call)."""
        for dst, src in zip(self.static_inputs, inputs):
            dst.detach().copy_(src, non_blocking=True)
        self.graph.replay()
        for p, g in zip(self.params, self.static_grads):
            if p.grad is not g:         # someone cleared or replaced .grad between replays
                p.grad = g
        return self.static_loss
