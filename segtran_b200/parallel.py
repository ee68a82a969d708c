"""Data-parallel gradient exchange for the hot path: one flat fp32 bucket, one NCCL all-reduce per step.

The reference wraps the whole model in DistributedDataParallel with find_unused_parameters=True
(train3d.py:671-676, train2d.py:1108-1113).  The Squeeze-and-Expansion path itself is per-sample, so the only
collective it needs is the mean of the parameter gradients over ranks.  ``GradBucket`` makes every parameter's
``.grad`` a view into ONE contiguous buffer (so no gather/scatter copies are needed), skips the parameters the
reference constructs but never uses (their .grad stays zero — same result as DDP's unused-parameter handling),
and issues a single all-reduce over NVLink/NVSwitch on a side stream.

Backend: NCCL on GPUs (one process per GPU, ``torch.distributed``); the same code runs on gloo for the CPU tests.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


_ALIGN = 64          # elements (256 bytes of fp32)


class GradBucket:
    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None, overlap_chunks: int = 0,
                 direct_accumulate: bool = False, milestones: bool = False):
        seen, self.params = set(), []
        for p in params:                      # tied parameters (query/key) appear once
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                self.params.append(p)
        if not self.params:
            raise ValueError("GradBucket: no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        # every parameter starts on a 256-byte boundary of the bucket (TMA / vector accesses on the views stay legal when
        # an optimiser lays the parameters out the same way); the padding stays zero and rides along in the all-reduce
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        self.flat = torch.zeros(self.numel, device=dev, dtype=dt)
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)      # autograd accumulates in place into the bucket
        # direct_accumulate: the weight-gradient GEMMs of the hot path add straight into these views (ops.set_grad_sink)
        # instead of producing a temporary that autograd adds in a second pass; needs zero() before every step, and the
        # per-parameter hooks of the overlap mode do not fire for those parameters
        if direct_accumulate:
            if overlap_chunks > 1:
                raise ValueError("GradBucket: direct_accumulate and overlap_chunks are mutually exclusive")
            from . import ops
            ops.set_grad_sink(True)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        # milestones: the modules announce "the gradients of these parameters are final" from tensor hooks placed at layer
        # boundaries (ops.grad_ready); each announcement starts the all-reduce of the covered bucket ranges on the side
        # stream while the rest of backward is still running.  Works with direct_accumulate (no AccumulateGrad hooks are
        # needed) and inside a captured CUDA graph (the side stream forks from and re-joins the capturing stream).
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._sent = [False] * len(self.params)
        self._milestones = bool(milestones) and self.world > 1
        if self._milestones:
            if overlap_chunks > 1:
                raise ValueError("GradBucket: milestones and overlap_chunks are mutually exclusive")
            from . import ops
            ops.set_grad_ready_callback(self.ready)
        self._comm = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        # NCCL averages inside the collective (no extra pass over the bucket); gloo has no AVG -> SUM then scale
        self._avg = self.world > 1 and dist.get_backend(process_group) == "nccl"
        self._op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._work = None
        self._works: List = []
        # optional overlap of the exchange with the rest of backward: the bucket is cut into contiguous chunks and a
        # chunk's all-reduce is issued (on the side stream) as soon as autograd has produced all of its gradients
        self._chunks = []
        if overlap_chunks > 1 and self.world > 1:
            target = (self.numel + overlap_chunks - 1) // overlap_chunks
            off, start, members = 0, 0, []
            for p in self.params:
                members.append(p)
                off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
                if off - start >= target:
                    self._chunks.append({"lo": start, "hi": off, "n": len(members), "left": len(members), "sent": False})
                    for q in members:
                        q._sx_chunk = len(self._chunks) - 1
                    start, members = off, []
            if members:
                self._chunks.append({"lo": start, "hi": off, "n": len(members), "left": len(members), "sent": False})
                for q in members:
                    q._sx_chunk = len(self._chunks) - 1
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._on_grad_ready)

    def _send_chunk(self, c):
        c["sent"] = True
        view = self.flat[c["lo"]:c["hi"]]
        if self._comm is not None:
            self._comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm):
                self._works.append(dist.all_reduce(view, op=self._op, group=self.group, async_op=True))
        else:
            self._works.append(dist.all_reduce(view, op=self._op, group=self.group, async_op=True))

    def _span(self, i):
        n = self.params[i].numel()
        return self.offsets[i], self.offsets[i] + (n + _ALIGN - 1) // _ALIGN * _ALIGN

    def _send_params(self, idx):
        """all-reduce the bucket ranges of the (sorted) parameter indices, merging neighbours into one call"""
        runs = []
        for i in idx:
            lo, hi = self._span(i)
            if runs and runs[-1][1] == lo:
                runs[-1][1] = hi
            else:
                runs.append([lo, hi])
            self._sent[i] = True
        if not runs:
            return
        if self._comm is not None:
            self._comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm):
                for lo, hi in runs:
                    self._works.append(dist.all_reduce(self.flat[lo:hi], op=self._op, group=self.group, async_op=True))
        else:
            for lo, hi in runs:
                self._works.append(dist.all_reduce(self.flat[lo:hi], op=self._op, group=self.group, async_op=True))

    def ready(self, params):
        """The gradients of `params` are final for this step: start exchanging them now (milestone mode)."""
        if not self._milestones:
            return
        self._send_params(sorted({self._index[id(p)] for p in params
                                  if id(p) in self._index and not self._sent[self._index[id(p)]]}))

    def _on_grad_ready(self, p):
        c = self._chunks[p._sx_chunk]
        c["left"] -= 1
        if c["left"] == 0 and not c["sent"]:
            self._send_chunk(c)

    def zero(self):
        """Replaces optimizer.zero_grad(): one memset instead of one per parameter; .grad views stay attached."""
        self.flat.zero_()
        for c in self._chunks:
            c["left"], c["sent"] = c["n"], False
        if self._milestones:
            self._sent = [False] * len(self.params)

    def reattach(self):
        """Call if something replaced p.grad (e.g. zero_grad(set_to_none=True))."""
        for p, off in zip(self.params, self.offsets):
            n = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat[off:off + n].data_ptr():
                p.grad = self.flat[off:off + n].view_as(p)

    def allreduce_async(self):
        """Average the bucket over ranks; returns immediately (the transfer runs on a side stream on GPUs)."""
        if self.world == 1:
            return
        if self._milestones:                  # whatever no milestone has covered (e.g. the first layer's norm / pos-code)
            self._send_params([i for i in range(len(self.params)) if not self._sent[i]])
            return
        if self._chunks:                      # overlap mode: send whatever backward has not triggered (unused params)
            for c in self._chunks:
                if not c["sent"]:
                    self._send_chunk(c)
            return
        if self._comm is not None:
            self._comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm):
                self._work = dist.all_reduce(self.flat, op=self._op, group=self.group, async_op=True)
                self._scaled = False
        else:
            self._work = dist.all_reduce(self.flat, op=self._op, group=self.group, async_op=True)
            self._scaled = False

    def wait(self):
        """Make the averaged gradients visible to the current stream (call before the optimizer step)."""
        if self.world == 1:
            return
        if self._chunks or self._milestones:
            if self._comm is not None:
                with torch.cuda.stream(self._comm):
                    for w in self._works:
                        w.wait()
                    self._scale()
                torch.cuda.current_stream().wait_stream(self._comm)
            else:
                for w in self._works:
                    w.wait()
                self._scale()
            self._works = []
            return
        if self._work is None:
            return
        if self._comm is not None:
            with torch.cuda.stream(self._comm):
                self._work.wait()
                self._scale()
            torch.cuda.current_stream().wait_stream(self._comm)
        else:
            self._work.wait()
            self._scale()
        self._work = None

    def _scale(self):
        if not self._avg:
            self.flat.mul_(1.0 / self.world)

    def bytes(self) -> int:
        """bytes moved by the all-reduce (parameters + alignment padding)"""
        return self.numel * self.flat.element_size()
