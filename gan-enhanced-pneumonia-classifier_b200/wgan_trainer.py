"""Fused WGAN-GP training iteration: the inner loop of the reference's `train_wggan.py:66-93` on the B200 kernels.

`WGANGPTrainer.critic_step(real, noise, alpha)` is one pass of train_wggan.py:71-85 (D(real), G(noise), D(fake.detach()), the
gradient penalty with its double backward, Adam with betas (beta1, 0.9)); `generator_step(noise)` is train_wggan.py:87-92;
`step(real)` runs `critic_iters` critic updates and one generator update with fresh noise / interpolation draws, like the
reference loop.  As in `trainer.DCGANTrainer`: every op is a libb200gan.so launch on the current stream without host
synchronisation, parameters / gradients / Adam moments of each network live in flat fp32 arenas (the nn.Module parameters are
views, so `state_dict()` is unchanged), one fused Adam launch per network, losses stay on the device.  Under torch.distributed the
critic's and the generator's gradient arenas are summed over ranks before their Adam updates (bucketed all-reduce on the
library's NCCL communicator, as for the DCGAN; BatchNorm statistics stay per rank).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from . import engine as E
from . import wgan_engine as W
from .dp import DPComm, GradBuckets
from .trainer import _Arena


class WGANGPTrainer:
    def __init__(self, netG, netD, lr: float = 2e-4, beta1: float = 0.5, beta2: float = 0.9, eps: float = 1e-8, lambda_gp: float = W.LAMBDA_GP,
                 critic_iters: int = 5, dtype: Optional[torch.dtype] = None, algo: Optional[int] = None, process_group=None):
        dtype = dtype or E.default_compute_dtype()
        algo = E.default_algo() if algo is None else algo
        self.netG, self.netD = netG, netD
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.lambda_gp, self.critic_iters, self.dtype = float(lambda_gp), int(critic_iters), dtype
        self.engG = E.NetEngine(netG._specs(), True, dtype, algo)
        self.engD = E.NetEngine(netD._specs(), False, dtype, algo)
        self.engG.weights_version = self.engD.weights_version = 0
        self.arenaG = _Arena(self.engG.param_order(netG))
        self.arenaD = _Arena(self.engD.param_order(netD))
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.comm = DPComm(process_group) if self.world > 1 else None
        self.bucketsD = GradBuckets(self.arenaD.grad, self.arenaD.slices, process_group, comm=self.comm)
        self.bucketsG = GradBuckets(self.arenaG.grad, self.arenaG.slices, process_group, comm=self.comm)
        self.extra_launches = 0

    @property
    def launches(self):
        return self.engG.launches + self.engD.launches + self.extra_launches

    # ------------------------------------------------------------------------------------------------
    def _adam(self, arena):
        arena.step_dev.add_(1)
        L.call('b200gan_adam', L.ptr(arena.param), L.ptr(arena.grad), L.ptr(arena.exp_avg), L.ptr(arena.exp_avg_sq), arena.numel, self.lr,
               self.beta1, self.beta2, self.eps, 0, L.ptr(arena.step_dev), 1.0 / self.world, L.stream_ptr())
        self.extra_launches += 1

    def _exchange(self, buckets):
        if self.comm is not None:
            buckets.begin()
            buckets.finish()

    def _mean(self, smap: torch.Tensor, sign: float) -> torch.Tensor:
        """sign * mean over the whole (N, 8, 8, 1) score map = sign * mean_b D(x)_b (wggan.py:69-70, train_wggan.py:74,79,90)."""
        out = torch.empty(1, device=smap.device, dtype=torch.float32)
        L.call('b200gan_mean_f32', L.ptr(smap), smap.numel(), sign / smap.numel(), L.ptr(out), L.stream_ptr())
        self.extra_launches += 1
        return out

    def _seed(self, smap: torch.Tensor, sign: float) -> E.Act:
        """d (sign * mean_b mean_hw map) / d map: a constant."""
        return E.Act(torch.full(smap.shape, sign / smap.numel(), device=smap.device, dtype=torch.float32), nchw=False)

    @staticmethod
    def _as_input(t):
        if t.dim() != 4:
            raise L.B200GanError(f'expected a 4-d NCHW tensor, got shape {tuple(t.shape)}')
        if t.dtype not in (torch.float32, torch.bfloat16):
            t = t.float()
        return E.Act(t, nchw=True)

    def critic_step(self, real: torch.Tensor, noise: Optional[torch.Tensor] = None, alpha: Optional[torch.Tensor] = None) -> torch.Tensor:
        """train_wggan.py:71-85.  Returns the (2,) float32 device tensor [d_loss, gp]."""
        n, dev = real.shape[0], real.device
        if noise is None:
            noise = torch.randn((n, self.engG.specs[0].cin, 1, 1), device=dev, dtype=torch.float32)
        if alpha is None:
            alpha = torch.rand((n, 1, 1, 1), device=dev, dtype=torch.float32)         # wggan.py:76
        pG = E.params_from_module(self.netG, self.engG.specs)
        pD = E.params_from_module(self.netD, self.engD.specs)
        self.arenaD.grad.zero_()
        real_in = self._as_input(real)
        smap_r, ctx = self.engD.forward(real_in, pD, True, True, last_act=False)
        l_real = self._mean(smap_r.t, -1.0)                                               # d_real_loss = -d_real.mean()
        self.engD.backward(ctx, pD, None, self.arenaD.grads, dlogit=self._seed(smap_r.t, -1.0))
        del ctx
        fake, _ = self.engG.forward(self._as_input(noise), pG, True, False)               # train mode: G's BatchNorm buffers move (:77)
        smap_f, ctx = self.engD.forward(fake, pD, True, True, last_act=False)
        l_fake = self._mean(smap_f.t, 1.0)                                                # d_fake_loss = d_fake.mean()
        self.engD.backward(ctx, pD, None, self.arenaD.grads, dlogit=self._seed(smap_f.t, 1.0))
        del ctx
        # x^ = alpha * real + (1 - alpha) * fake (wggan.py:77), fp32 NCHW like the reference's tensor
        xhat = torch.empty((n, real.shape[1], real.shape[2], real.shape[3]), device=dev, dtype=torch.float32)
        a = alpha.reshape(-1).contiguous().float()
        b = (1.0 - a).contiguous()
        L.call('b200gan_sample_axpby', C.byref(real_in.v), L.ptr(a), C.byref(fake.v), L.ptr(b), C.byref(L.view_nchw(xhat)), L.stream_ptr())
        self.extra_launches += 1
        gp = W.gradient_penalty(self.engD, pD, xhat, self.arenaD.grads, self.lambda_gp)
        self._exchange(self.bucketsD)
        self._adam(self.arenaD)
        self.engD.weights_version += 1
        return torch.stack([(l_real + l_fake).view(()) + gp, gp])

    def generator_step(self, noise: torch.Tensor) -> torch.Tensor:
        """train_wggan.py:87-92.  Returns the () float32 device tensor g_loss = -mean D(G(z))."""
        pG = E.params_from_module(self.netG, self.engG.specs)
        pD = E.params_from_module(self.netD, self.engD.specs)
        self.arenaG.grad.zero_()
        fake, ctx_g = self.engG.forward(self._as_input(noise), pG, True, True)
        smap, ctx_d = self.engD.forward(fake, pD, True, True, last_act=False)
        g_loss = self._mean(smap.t, -1.0)
        dfake = E.Act(torch.empty_like(fake.t), nchw=False)
        self.engD.backward(ctx_d, pD, None, [None] * len(self.arenaD.grads), dinput=dfake, need_wgrad=False, dlogit=self._seed(smap.t, -1.0))
        del ctx_d
        self.engG.backward(ctx_g, pG, dfake, self.arenaG.grads)
        del ctx_g
        self._exchange(self.bucketsG)
        self._adam(self.arenaG)
        self.engG.weights_version += 1
        return g_loss.view(())

    def step(self, real: torch.Tensor) -> torch.Tensor:
        """One iteration of the reference loop (train_wggan.py:66-93): `critic_iters` critic updates, then one generator update,
        each with fresh noise.  Returns the (critic_iters + 1,) device tensor [d_loss_1 .. d_loss_k, g_loss]."""
        n, dev, nz = real.shape[0], real.device, self.engG.specs[0].cin
        out = []
        for _ in range(self.critic_iters):
            out.append(self.critic_step(real)[0])
        out.append(self.generator_step(torch.randn((n, nz, 1, 1), device=dev, dtype=torch.float32)))
        return torch.stack(out)

    @torch.no_grad()
    def sample(self, noise: torch.Tensor) -> torch.Tensor:
        """The visualisation forward of train_wggan.py:101-103 (train mode under no_grad: BatchNorm buffers move)."""
        return self.netG(noise)

    def close(self):
        if self.comm is not None:
            torch.cuda.synchronize()
            self.comm.close()
            self.comm = self.bucketsD.comm = self.bucketsG.comm = None
