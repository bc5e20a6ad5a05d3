"""Fused DCGAN training step: the inner loop of the reference's `train_gan.py:119-150` as one call.

`DCGANTrainer.step(real, noise)` performs, with the same arithmetic and side effects as the reference loop,

    D.zero_grad; D(real) -> BCE(.,0.9) -> backward; G(noise); D(fake.detach()) -> BCE(.,0) -> backward;
    Adam(D);  G.zero_grad; D(fake) -> BCE(.,0.9) -> backward through D into G; Adam(G)

but (a) every op is a libb200gan.so kernel launched on the current stream with no host synchronisation,
(b) parameters, gradients and Adam moments of each network live in flat fp32 arenas (the nn.Module
parameters are re-pointed into them, so `state_dict()`/checkpoints are unaffected) and the optimizer is one
fused multi-tensor launch per network, (c) Sigmoid + BCE + their backward are one kernel, (d) the dead
D-weight-gradient work of the G step (its result is discarded by `netD.zero_grad()` at train_gan.py:122) is
skipped, (e) the five per-iteration history scalars are left on the device in a (5,) tensor, (f) under
data parallelism the two gradient arenas are all-reduced over NCCL (sum, then scaled by 1/world inside the Adam
kernel) -- the reference has no multi-GPU path; semantics are "mean of per-rank gradients", BatchNorm
statistics stay local to each rank -- and (g) the whole iteration (about 150 kernel launches)
is captured once per input shape into a CUDA graph and replayed: the Adam step counts and BatchNorm's
num_batches_tracked advance on the device, so nothing about the iteration is baked into the captured launches.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib as L
from . import engine as E
from .dp import DPComm, GradBuckets

REAL_LABEL = 0.9      # train_gan.py:92
FAKE_LABEL = 0.0      # train_gan.py:93


class _Arena:
    """Flat fp32 storage for a network's parameters, gradients and Adam state."""

    def __init__(self, params):
        dev = params[0].device
        sizes = [p.numel() for p in params]
        # keep every tensor 16-byte aligned inside the arena (vectorised Adam, future TMA use)
        offs, o = [], 0
        for s in sizes:
            offs.append(o)
            o += (s + 3) // 4 * 4
        self.numel = o
        self.param = torch.zeros(o, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(o, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(o, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(o, device=dev, dtype=torch.float32)
        self.grads = []
        self.slices = [(off, off + s) for off, s in zip(offs, sizes)]    # per parameter, in param_order()
        with torch.no_grad():
            for p, off, s in zip(params, offs, sizes):
                view = self.param[off:off + s].view(p.shape)
                view.copy_(p.data)
                p.data = view                          # the module now reads/writes the arena
                self.grads.append(self.grad[off:off + s].view(p.shape))
        self.step = 0                                                    # host mirror of step_dev
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int64)   # read by the Adam kernel (graph-replay safe)


class _Captured:
    __slots__ = ('graph', 'real', 'noise', 'out', 'launches')     # graph: [(CUDAGraph, exchange 'D'/'G'/None after it)]


class DCGANTrainer:
    def __init__(self, netG, netD, lr: float = 2e-4, beta1: float = 0.5, beta2: float = 0.999, eps: float = 1e-8,
                 dtype: Optional[torch.dtype] = None, algo: Optional[int] = None, process_group=None, use_graph: Optional[bool] = None,
                 sync_bn: bool = False):
        dtype = dtype or E.default_compute_dtype()
        algo = E.default_algo() if algo is None else algo
        self.netG, self.netD = netG, netD
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.engG = E.NetEngine(netG._specs(), True, dtype, algo)
        self.engD = E.NetEngine(netD._specs(), False, dtype, algo)
        self.engG.weights_version = self.engD.weights_version = 0      # this trainer owns the weights: repack only after Adam
        dev = next(netD.parameters()).device
        if os.environ.get('B200GAN_WGRAD_STREAM', '1') != '0' and dev.type == 'cuda':
            # weight gradients off the backward pass's critical path (engine.NetEngine.wgrad_stream); B200GAN_WGRAD_STREAM=0 is the
            # single-stream order, kept for A/B timing
            self.engG.wgrad_stream = self.engD.wgrad_stream = torch.cuda.Stream(device=dev)
        self.gfwd_stream = torch.cuda.Stream(device=dev) if (os.environ.get('B200GAN_GFWD_STREAM', '1') != '0' and dev.type == 'cuda') else None
        self.arenaG = _Arena(self.engG.param_order(netG))
        self.arenaD = _Arena(self.engD.param_order(netD))
        self.dtype = dtype
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.extra_launches = 0
        # data parallel: the library's own NCCL communicator (b200gan_dp_*); torch.distributed only carried its unique id
        self.comm = DPComm(process_group) if self.world > 1 else None
        # --sync-bn: BatchNorm statistics (and the BatchNorm-backward reductions) over the GLOBAL batch instead of the rank's shard: one small
        # fp64 all-reduce per BatchNorm pass on the library's communicator (optional).  The iteration then equals a
        # single process on the concatenated batch.
        self.sync_bn = bool(sync_bn) and self.comm is not None
        if self.sync_bn:
            self.engG.sync_bn = self.engD.sync_bn = self.comm
        self.bucketsD = GradBuckets(self.arenaD.grad, self.arenaD.slices, process_group, comm=self.comm)
        self.bucketsG = GradBuckets(self.arenaG.grad, self.arenaG.slices, process_group, comm=self.comm)
        # B200GAN_DP_SEGMENTED=1: keep the collectives out of the capture (three graphs cut at the two exchanges, every bucket
        # issued between the replays) -- a diagnostic switch; the default captures the bucket all-reduces with the iteration
        self.segmented = os.environ.get('B200GAN_DP_SEGMENTED', '0') == '1'
        if use_graph is None:
            use_graph = os.environ.get('B200GAN_GRAPH', '1') != '0'
        self.use_graph = use_graph
        self._graphs = {}
        self._warm = set()
        self._static_in = {}
        self._replayed_launches = 0

    # ------------------------------------------------------------------------------------------------
    @property
    def launches(self):
        """Number of kernel-launching libb200gan calls made so far, replayed graph nodes included (bench.py's gpu_launches claim)."""
        return self.engG.launches + self.engD.launches + self.extra_launches + self._replayed_launches

    def _bce(self, logits, target, want_grad=True):
        n = logits.t.shape[0]
        dev = logits.t.device
        out2 = torch.empty(2, device=dev, dtype=torch.float32)
        dl = torch.empty((n, 1, 1, 1), device=dev, dtype=torch.float32) if want_grad else None
        L.call('b200gan_bce_sigmoid', L.ptr(logits.t), n, target, 1.0, None, L.ptr(out2), L.ptr(dl), L.stream_ptr())
        self.extra_launches += 1
        return out2, (E.Act(dl, nchw=False) if want_grad else None)

    def _adam(self, arena):
        arena.step_dev.add_(1)                   # device-side step count (a captured node under graph replay)
        L.call('b200gan_adam', L.ptr(arena.param), L.ptr(arena.grad), L.ptr(arena.exp_avg), L.ptr(arena.exp_avg_sq),
               arena.numel, self.lr, self.beta1, self.beta2, self.eps, 0, L.ptr(arena.step_dev), 1.0 / self.world, L.stream_ptr())
        self.extra_launches += 1

    def _as_input(self, t):
        """Accept the reference's NCHW tensors (fp32 or bf16) without copying."""
        if t.dim() != 4:
            raise L.B200GanError(f'expected a 4-d NCHW tensor, got shape {tuple(t.shape)}')
        if t.dtype not in (torch.float32, torch.bfloat16):
            t = t.float()
        return E.Act(t, nchw=True)

    def step(self, real: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One adversarial iteration.  `noise` (N, nz, 1, 1) defaults to torch.randn on the device (train_gan.py:132).
        Returns a (5,) float32 CUDA tensor [errD, errG, D_x, D_G_z1, D_G_z2] (the history scalars of
        train_gan.py:153-157), not synchronised.  The first call for an input shape runs kernel by kernel; the second
        captures the iteration into CUDA graph(s); later calls replay."""
        self.arenaD.step += 1
        self.arenaG.step += 1
        if not self.use_graph:
            return self._step_eager(real, noise)
        key = (tuple(real.shape), real.dtype, None if noise is None else (tuple(noise.shape), noise.dtype))
        if key not in self._warm:
            # lazy one-time initialisation (function attributes, driver entry points) must not happen under capture
            self._warm.add(key)
            return self._step_eager(real, noise)
        cap = self._graphs.get(key)
        if cap is None:
            cap = self._capture(real, noise, key)
        if real.data_ptr() != cap.real.data_ptr():
            cap.real.copy_(real)
        if noise is not None and noise.data_ptr() != cap.noise.data_ptr():
            cap.noise.copy_(noise)
        for graph, exchange in cap.graph:
            graph.replay()
            if exchange is not None:
                self._exchange(exchange, overlapped=False)
        self._replayed_launches += cap.launches
        return cap.out.clone()

    def input_buffers(self, real_shape, real_dtype=torch.float32, noise_shape=None):
        """The static input tensors of the captured iteration for this shape (created on first use): filling them directly
        (e.g. with a non-blocking H2D copy) and passing them to `step` avoids the device-to-device staging copy."""
        key = (tuple(real_shape), real_dtype, None if noise_shape is None else (tuple(noise_shape), torch.float32))
        bufs = self._static_in
        if key not in bufs:
            dev = self.arenaG.param.device
            bufs[key] = (torch.zeros(real_shape, device=dev, dtype=real_dtype),
                         None if noise_shape is None else torch.zeros(noise_shape, device=dev, dtype=torch.float32))
        return bufs[key]

    def _exchange(self, which: str, overlapped: bool = True):
        """Gradient exchange between ranks before the Adam update of network `which` ('D' or 'G'): sum over ranks (the 1/world
        factor is applied inside the Adam kernel).  The buckets were launched on the library's communication stream while the
        backward pass ran (dp.GradBuckets.ready); what remains is the last bucket and the stream join.  overlapped=False
        (segmented diagnostic mode, between graph replays): every bucket goes out now."""
        if self.comm is None:
            return
        buckets = self.bucketsD if which == 'D' else self.bucketsG
        if not overlapped:
            buckets.begin()
        buckets.finish()

    def _capture(self, real, noise, key):
        """The whole iteration is ONE graph, data parallel included: the bucket all-reduces are captured with it (the library forks
        its communication stream from the capturing stream per bucket and joins it before each Adam update), so a replayed
        iteration overlaps NCCL traffic with the backward pass exactly like the kernel-by-kernel one.  The communicator was
        warmed up by the eager first iteration (NCCL sets up its channels on first use, which must not happen under capture)."""
        cap = _Captured()
        cap.real, cap.noise = self.input_buffers(real.shape, real.dtype, None if noise is None else noise.shape)
        cap.real.copy_(real)
        if noise is not None:
            cap.noise.copy_(noise)
        l0 = self.engG.launches + self.engD.launches + self.extra_launches
        torch.cuda.synchronize()
        gen = self._segments(cap.real, cap.noise, overlap=not self.segmented)
        cap.graph = []
        pool = torch.cuda.graph_pool_handle()
        done = False
        while not done:
            g = torch.cuda.CUDAGraph()
            exchange = None
            # thread_local: NCCL's proxy / torch's watchdog threads may touch the CUDA API while this thread captures
            with torch.cuda.graph(g, pool=pool, capture_error_mode='thread_local' if self.world > 1 else 'global'):
                while True:
                    try:
                        exchange = next(gen)
                    except StopIteration as e:
                        cap.out, done, exchange = e.value, True, None
                        break
                    if self.comm is not None and self.segmented:
                        break                    # cut the graph here; the exchange runs between replays
                    self._exchange(exchange)     # captured: last bucket + join of the communication stream
                    exchange = None
            cap.graph.append((g, exchange))
        cap.launches = self.engG.launches + self.engD.launches + self.extra_launches - l0
        # capture records, it does not execute: take the recorded launches back out of the eager counters
        self.extra_launches -= cap.launches
        self._graphs[key] = cap
        return cap

    def _step_eager(self, real: torch.Tensor, noise: Optional[torch.Tensor]) -> torch.Tensor:
        gen = self._segments(real, noise, overlap=True)
        while True:
            try:
                self._exchange(next(gen))
            except StopIteration as e:
                return e.value

    def _segments(self, real: torch.Tensor, noise: Optional[torch.Tensor], overlap: bool):
        """The iteration as a generator that yields 'D' / 'G' at the two points where the gradients of that network are complete
        on this rank and must be exchanged before its Adam update; returns the history tensor."""
        netG, netD = self.netG, self.netD
        if noise is None:
            noise = torch.randn((real.shape[0], self.engG.specs[0].cin, 1, 1), device=real.device, dtype=torch.float32)
        pG = E.params_from_module(netG, self.engG.specs)
        pD = E.params_from_module(netD, self.engD.specs)
        readyD = self.bucketsD.ready if (overlap and self.comm is not None) else None
        readyG = self.bucketsG.ready if (overlap and self.comm is not None) else None
        # (1) D step ------------------------------------------------------------- train_gan.py:122-141
        self.arenaD.grad.zero_()
        if self.gfwd_stream is not None:
            # the Generator's forward does not depend on the real half of the D step: it runs beside it on its own stream
            self.gfwd_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.gfwd_stream):
                fake, ctx_g = self.engG.forward(self._as_input(noise), pG, True, True)
        logit_r, ctx_r = self.engD.forward(self._as_input(real), pD, True, True, last_act=False)
        m_real, dl = self._bce(logit_r, REAL_LABEL)
        self.engD.backward(ctx_r, pD, None, self.arenaD.grads, dlogit=dl)
        del ctx_r
        if self.gfwd_stream is not None:
            torch.cuda.current_stream().wait_stream(self.gfwd_stream)
        else:
            fake, ctx_g = self.engG.forward(self._as_input(noise), pG, True, True)
        logit_f, ctx_f = self.engD.forward(fake, pD, True, True, last_act=False)
        m_fake, dl = self._bce(logit_f, FAKE_LABEL)
        self.bucketsD.begin()                    # real + fake gradients have both accumulated once this backward has written them
        self.engD.backward(ctx_f, pD, None, self.arenaD.grads, dlogit=dl, on_ready=readyD)
        del ctx_f
        self.engD.join_wgrads()
        yield 'D'
        self._adam(self.arenaD)
        self.engD.weights_version += 1
        # (2) G step ------------------------------------------------------------- train_gan.py:144-150
        self.arenaG.grad.zero_()
        logit_g, ctx_d = self.engD.forward(fake, pD, True, True, last_act=False)
        m_g, dl = self._bce(logit_g, REAL_LABEL)
        dfake = E.Act(torch.empty_like(fake.t), nchw=False)
        self.engD.backward(ctx_d, pD, None, [None] * len(self.arenaD.grads), dinput=dfake, need_wgrad=False, dlogit=dl)
        del ctx_d
        self.bucketsG.begin()
        self.engG.backward(ctx_g, pG, dfake, self.arenaG.grads, on_ready=readyG)
        del ctx_g
        self.engG.join_wgrads()
        yield 'G'
        self._adam(self.arenaG)
        self.engG.weights_version += 1
        # errD = errD_real + errD_fake (train_gan.py:140); D_x, D_G_z1, D_G_z2 are mean probabilities
        return torch.stack([m_real[0] + m_fake[0], m_g[0], m_real[1], m_fake[1], m_g[1]])

    def close(self):
        """Data parallel: drop the captured graphs (they hold the communicator's captured collectives), drain the device, release the
        library's NCCL communicator.  Call before torch.distributed.destroy_process_group(); a no-op on one GPU."""
        self._graphs.clear()
        if self.comm is not None:
            torch.cuda.synchronize()
            self.comm.close()
            self.comm = self.bucketsD.comm = self.bucketsG.comm = None
            self.engG.sync_bn = self.engD.sync_bn = None

    def refresh_packed_weights(self):
        """Re-derive the cached bf16 weight repacks from the fp32 masters.  Call after the weights were changed from OUTSIDE the
        trainer (`load_state_dict`, manual edits): the repacks are otherwise only refreshed after the trainer's own Adam updates,
        and a captured CUDA graph reads the persistent repack buffers its previous replay wrote."""
        st = L.stream_ptr()
        for eng, net in ((self.engD, self.netD), (self.engG, self.netG)):
            for i, sp in enumerate(eng.specs):
                hit = eng._packed.get(i)
                if hit is None:
                    continue
                w = net.main[sp.conv_idx].weight
                L.call('b200gan_pack_conv_weight', L.ptr(w), w.shape[0], w.shape[1], 4, 2, L.ptr(hit[1]), st)

    @torch.no_grad()
    def sample(self, noise: torch.Tensor) -> torch.Tensor:
        """The visualisation forward of train_gan.py:166-169 (train-mode under no_grad: BatchNorm buffers move)."""
        return self.netG(noise)
