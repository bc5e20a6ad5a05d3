"""Data-parallel gradient buckets (the reference has no multi-GPU path; SURVEY.md section 8e defines the semantics).

One process per GPU, weights and Adam state replicated, the batch sharded.  Each network's gradients live in ONE flat fp32
arena (trainer._Arena); `GradBuckets` cuts that arena into contiguous buckets in the order the backward pass finishes them
(last layer first) and all-reduces (sum) each bucket as soon as its last gradient is final, on a communication stream, so the
NCCL traffic over NVLink overlaps the rest of the backward pass.  The 1/world scaling is folded into the Adam kernel
(`b200gan_adam(grad_scale=1/world)`), so the collective is a plain sum.  BatchNorm statistics stay local to each rank.

The class only touches torch tensors and `torch.distributed`, so the same code runs on CPU tensors over gloo (that is how
tests/test_dp_gloo.py covers it without GPUs) and on CUDA tensors over NCCL (also inside CUDA-graph capture: the communication
stream forks from and joins the capturing stream through events).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class GradBuckets:
    def __init__(self, grad_arena: torch.Tensor, slices: Sequence[Tuple[int, int]], group=None, bucket_numel: int = 1 << 20):
        """grad_arena: flat tensor holding every gradient; slices[i] = (start, end) of parameter i in `param_order()`
        (layer 0 first).  Buckets are built from the END of the arena (the last layer's gradients are final first)."""
        self.arena, self.group = grad_arena, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.buckets: List[Tuple[int, int, List[int]]] = []      # (start, end, parameter indices)
        cur: List[int] = []
        for i in reversed(range(len(slices))):
            cur.append(i)
            lo, hi = slices[cur[-1]][0], slices[cur[0]][1]
            if hi - lo >= bucket_numel or i == 0:
                self.buckets.append((lo, hi, cur))
                cur = []
        self._bucket_of = {}
        for b, (_, _, idx) in enumerate(self.buckets):
            for i in idx:
                self._bucket_of[i] = b
        self._pending = [set(idx) for (_, _, idx) in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._comm: Optional[torch.cuda.Stream] = torch.cuda.Stream() if grad_arena.is_cuda else None
        self.collectives = 0

    def begin(self):
        """Start of the backward pass whose gradients are final when it ends."""
        self._pending = [set(idx) for (_, _, idx) in self.buckets]
        self._launched = [False] * len(self.buckets)

    def ready(self, param_index: int):
        """The gradient of parameter `param_index` is final (called by the backward pass right after the launch that wrote it)."""
        if self.world == 1:
            return
        b = self._bucket_of[param_index]
        self._pending[b].discard(param_index)
        if not self._pending[b] and not self._launched[b]:
            self._launch(b)

    def _launch(self, b: int):
        lo, hi, _ = self.buckets[b]
        self._launched[b] = True
        self.collectives += 1
        view = self.arena[lo:hi]
        if self._comm is None:
            dist.all_reduce(view, group=self.group)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ev)
            dist.all_reduce(view, group=self.group)

    def finish(self):
        """Launch whatever has not gone out yet and make the current stream wait for every bucket (call before Adam)."""
        if self.world == 1:
            return
        for b in range(len(self.buckets)):
            if not self._launched[b]:
                self._launch(b)
        if self._comm is not None:
            torch.cuda.current_stream().wait_stream(self._comm)
