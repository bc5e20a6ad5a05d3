"""Data-parallel gradient buckets (the reference has no multi-GPU path; SURVEY.md section 8e defines the semantics).

One process per GPU, weights and Adam state replicated, the batch sharded.  Each network's gradients live in ONE flat fp32
arena (trainer._Arena); `GradBuckets` cuts that arena into contiguous buckets in the order the backward pass finishes them
(last layer first) and all-reduces (sum) each bucket as soon as its last gradient is final, on a communication stream, so the
NCCL traffic over NVLink overlaps the rest of the backward pass.  The 1/world scaling is folded into the Adam kernel
(`b200gan_adam(grad_scale=1/world)`), so the collective is a plain sum.  BatchNorm statistics stay local to each rank.

Transport: for CUDA arenas the collectives go through the library's own communicator (`DPComm` = b200gan_dp_* of
include/b200gan.h: ncclCommInitRank from a broadcast unique id, a dedicated communication stream, event fork / join against the
compute stream, capturable into the iteration's CUDA graph); `torch.distributed` only carries the 128-byte id.  For CPU arenas
(tests/test_dp_gloo.py, world_size 2 over gloo, no GPU) the same bucket plan runs over `torch.distributed.all_reduce`.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L


class DPComm:
    """The library's NCCL communicator for this process' GPU (b200gan_dp_init): rank 0 draws the unique id, `torch.distributed`
    (any backend; `group` optional) broadcasts its 128 bytes, every rank joins.  Create AFTER torch.cuda.set_device."""

    def __init__(self, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise L.B200GanError('DPComm needs an initialised torch.distributed process group to exchange the NCCL unique id')
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        buf = (C.c_ubyte * L.DP_ID_BYTES)()
        if self.rank == 0:
            L.call('b200gan_dp_unique_id', C.cast(buf, C.c_void_p))
        dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' else torch.device('cpu')
        t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_ubyte * L.DP_ID_BYTES)(*t.cpu().tolist())
        self._h = C.c_void_p()
        L.call('b200gan_dp_init', C.cast(ident, C.c_void_p), self.world, self.rank, C.byref(self._h))

    def allreduce_bucket(self, view: torch.Tensor):
        L.call('b200gan_dp_allreduce_bucket', self._h, L.ptr(view), view.numel(), L.stream_ptr())

    def allreduce_f64(self, sums: torch.Tensor):
        """Synchronised BatchNorm: sum the per-channel fp64 sums of one layer over the ranks, in stream order."""
        L.call('b200gan_dp_allreduce_f64', self._h, L.ptr(sums), sums.numel(), L.stream_ptr())

    def sync(self):
        L.call('b200gan_dp_sync', self._h, L.stream_ptr())

    @property
    def collectives(self) -> int:
        return int(L.load().b200gan_dp_collectives(self._h)) if self._h else 0

    def close(self):
        """Release the communicator (after the CUDA graphs that captured it: DCGANTrainer.close does both in order).  Not called
        from __del__: at interpreter shutdown the CUDA context may already be gone, and process exit releases everything anyway."""
        if self._h:
            h, self._h = self._h, C.c_void_p()
            L.call('b200gan_dp_destroy', h)


class GradBuckets:
    def __init__(self, grad_arena: torch.Tensor, slices: Sequence[Tuple[int, int]], group=None, bucket_numel: int = 1 << 20,
                 comm: Optional[DPComm] = None):
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
        self.comm = comm                 # CUDA arenas: the library's communicator; None: torch.distributed (CPU / gloo tests)
        if grad_arena.is_cuda and self.world > 1 and comm is None:
            raise L.B200GanError('GradBuckets over a CUDA arena needs a DPComm (the library owns the NCCL communicator)')
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
        if self.comm is None:
            dist.all_reduce(view, group=self.group)
        else:
            self.comm.allreduce_bucket(view)         # asynchronous, on the library's communication stream

    def finish(self):
        """Launch whatever has not gone out yet and make the current stream wait for every bucket (call before Adam)."""
        if self.world == 1:
            return
        for b in range(len(self.buckets)):
            if not self._launched[b]:
                self._launch(b)
        if self.comm is not None:
            self.comm.sync()
