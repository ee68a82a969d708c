"""CUDA-graph capture of a whole training step (forward + loss + backward) of the hot path.

The stack launches ~115 kernels per step at cfg 4; enqueueing them from Python costs about as much host time as the
GPU needs to run them.  Every kernel on the path is capture-safe (no host synchronisation, no host-side random
numbers: dropout seeds live on the device, see ops.new_dropout_seed), so the step can be recorded once and replayed
with a single launch.  At N>1 the gradient all-reduce is captured with it: GradBucket issues the NCCL calls on a side
stream that forks from and re-joins the capturing stream (parallel.py, milestones), so they become nodes of the same graph.
"""
from __future__ import annotations

from typing import Callable

import torch


class CapturedStep:
    """graph = CapturedStep(fn); out = graph()   — `fn` must be static in shapes and use static input tensors."""

    def __init__(self, fn: Callable[[], torch.Tensor], warmup: int = 3):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                 # warm-up off the default stream, as capture requires
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count
        with torch.cuda.graph(self.graph):
            self.out = fn()
        self.kernel_launches = _lib.launch_count - n0          # library kernels recorded in (= run by) one replay

    def __call__(self):
        self.graph.replay()
        return self.out
