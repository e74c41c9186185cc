"""CUDA-graph replay of a scoring step for a fixed input shape (latency at small batches).

At batch 1 the step is ~100 kernel launches of a few microseconds each and the host launch path (ctypes call, tensor
map lookup, cudaLaunchKernelEx) costs more than the kernels.  Every launch of the library goes to the stream it is
given and allocates nothing, so a step can be captured once — including the programmatic-dependent-launch edges —
and replayed.  torch is used for the capture plumbing only (`torch.cuda.CUDAGraph`, private memory pool).
"""
from __future__ import annotations

from typing import Callable

import torch


class GraphedStep:
    """`fn(images_device) -> tuple of device tensors`, captured for `example` 's shape/dtype.

    __call__(images) copies `images` (host or device) into the static input, replays the graph and returns the static
    output tensors (overwritten by the next call)."""

    def __init__(self, fn: Callable, example: torch.Tensor, warmup: int = 3):
        if not example.is_cuda:
            raise RuntimeError("GraphedStep needs a CUDA example tensor")
        self.static_in = example.clone()
        side = torch.cuda.Stream(example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):  # packs weights, sizes workspaces, fills the tensor-map cache
                fn(self.static_in)
        torch.cuda.current_stream(example.device).wait_stream(side)
        torch.cuda.synchronize(example.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = tuple(fn(self.static_in))

    def __call__(self, images: torch.Tensor):
        self.static_in.copy_(images, non_blocking=True)
        self.graph.replay()
        return self.static_out
