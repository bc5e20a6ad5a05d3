"""Fused training iteration of the conditional GAN (reference src/train_cgan.py:150-193) on the kernels of
libb200gan.so: no autograd, flat parameter / gradient / Adam arenas, every loss term a kernel.

What one `step` does, in the reference's order:
  D step  (:161-178)   D(real) + BCEWithLogits vs randomly smoothed 0.9 targets; G(noise, random labels); D(fake) vs smoothed 0.1 targets; the
                       Discriminator's backward + Adam -- unless the reference's rule skips it: from epoch 5 on only while D(x) < 0.8 or
                       D(G(z)) > 0.2 (the one place a value is read back to the host, as the reference's `.item()` does every iteration)
  G step  (:180-193)   D(fake) again with the updated weights (adversarial term vs the smoothed real targets), D(real) for the feature-matching
                       targets, ONE backward through the Discriminator carrying both the logit gradient and the gradients of all intermediate
                       features (the reference's second features pass over the same fake batch is bit-for-bit the first: only its BatchNorm
                       running-statistics side effect is replayed), the Generator's backward + Adam.
The VGG16 perceptual term (:57-73,186; weight 10) is `perceptual.PerceptualLoss` handed to the constructor: its gradient w.r.t. the fake batch is added to
the one coming back through the Discriminator before the Generator's backward.  The ImageNet checkpoint the reference downloads cannot be obtained
offline, so the CALLER supplies the VGG16 (checkpoint path, torchvision's cache, or random weights for benchmarks); without one the term is dropped
only on an explicit `perceptual_weight=0`, never silently.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib as L
from . import cgan_engine as CE
from .dp import DPComm, GradBuckets
from .engine import Act
from .trainer import _Arena

REAL_LABEL, FAKE_LABEL, SMOOTH = 0.9, 0.1, 0.1          # train_cgan.py:121-122,156-160
FEATURE_MULTIPLICITY = [2, 1, 2, 1, 2, 1, 2, 1, 2]      # of [a0, y1, a1, ..., y4, a4] in get_intermediate_features' aliased 14-entry list


class CGANTrainer:
    def __init__(self, netG, netD, lr: float = 2e-4, beta1: float = 0.5, beta2: float = 0.999, eps: float = 1e-8, fm_weight: float = 5.0,
                 perceptual=None, perceptual_weight: float = 10.0, dtype: Optional[torch.dtype] = None, process_group=None):
        """perceptual: a `perceptual.PerceptualLoss` (VGG16 with whatever weights the caller could obtain) or None; with None the term is only
        dropped when perceptual_weight is explicitly 0 -- asking for the reference's loss (weight 10) without a VGG16 raises."""
        if perceptual is None and perceptual_weight != 0.0:
            raise L.B200GanError('the generator loss of train_cgan.py:191 includes 10 x a VGG16 perceptual term: pass perceptual=PerceptualLoss(...) '
                                 '(perceptual.py), or perceptual_weight=0 to train with the adversarial and feature-matching terms only')
        self.perceptual, self.perceptual_weight = perceptual, (perceptual_weight if perceptual is not None else 0.0)
        self.netG, self.netD = netG, netD
        self.lr, self.beta1, self.beta2, self.eps, self.fm_weight = lr, beta1, beta2, eps, fm_weight
        if dtype is not None:
            netG.compute_dtype = netD.compute_dtype = dtype
        self.engG, self.engD = netG._engine_for(), netD._engine_for()
        self.pG, self.pD = list(netG.parameters()), list(netD.parameters())
        self.arenaG, self.arenaD = _Arena(self.pG), _Arena(self.pD)
        self.intoG = dict(zip(self.pG, self.arenaG.grads))
        self.intoD = dict(zip(self.pD, self.arenaD.grads))
        self.d_steps = 0
        self._fm_scale = None
        # data parallel (one process per GPU, the batch sharded): both gradient arenas are summed over the ranks on the library's NCCL communicator
        # before their Adam updates (1/world folded into the Adam kernel); BatchNorm statistics stay per rank; the D-step skip rule is evaluated
        # on the rank-averaged D(x) / D(G(z)) so that every rank takes the same branch (and issues the same collectives)
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.comm = DPComm(process_group) if self.world > 1 else None
        self.bucketsD = GradBuckets(self.arenaD.grad, self.arenaD.slices, process_group, comm=self.comm)
        self.bucketsG = GradBuckets(self.arenaG.grad, self.arenaG.slices, process_group, comm=self.comm)

    # -- pieces ---------------------------------------------------------------------------------------------------------------------
    def _adam(self, arena):
        arena.step_dev.add_(1)
        L.call('b200gan_adam', L.ptr(arena.param), L.ptr(arena.grad), L.ptr(arena.exp_avg), L.ptr(arena.exp_avg_sq), arena.numel, self.lr,
               self.beta1, self.beta2, self.eps, 0, L.ptr(arena.step_dev), 1.0 / self.world, L.stream_ptr())

    def _exchange(self, buckets):
        if self.comm is not None:
            buckets.begin()
            buckets.finish()

    def close(self):
        """Data parallel: release the library's NCCL communicator (call before torch.distributed.destroy_process_group(); a no-op on one GPU)."""
        if self.comm is not None:
            torch.cuda.synchronize()
            self.comm.close()
            self.comm = self.bucketsD.comm = self.bucketsG.comm = None

    @staticmethod
    def _bce(logits, target, want_grad=True):
        out2 = torch.empty(2, device=logits.device, dtype=torch.float32)            # [loss, mean sigmoid]
        dl = torch.empty_like(logits) if want_grad else None
        L.call('b200gan_bce_logits', L.ptr(logits), L.ptr(target), logits.numel(), 1.0, L.ptr(out2), L.ptr(dl), L.stream_ptr())
        return out2, dl

    def _feature_matching(self, tape_real, tape_fake):
        """Returns (loss, adders): the loss as the reference sums it over the 14 aliased entries -- a device scalar that is complete once every
        adder has run -- and, per distinct fake intermediate, a callable that ADDS that term's gradient into the tensor the Discriminator's
        backward pass hands it (`b200gan_fm_pair` computes the sum and the gradient in the same pass over the pair)."""
        real_acts, fake_acts = self.engD.feature_acts(tape_real), self.engD.feature_acts(tape_fake)
        dev = fake_acts[0][0].t.device
        sums = torch.zeros(len(fake_acts), device=dev, dtype=torch.float64)
        key = tuple(f.t.numel() for _, f in fake_acts)                            # real (unpadded) element counts
        if self._fm_scale is None or self._fm_scale[0] != key:                   # multiplicity / numel per pair, built once per batch shape
            self._fm_scale = (key, torch.tensor([mult / nel for mult, nel in zip(FEATURE_MULTIPLICITY, key)], device=dev, dtype=torch.float64))

        def adder(j, r, f):
            def add(target: Act):
                L.call('b200gan_fm_pair', C.byref(r.v), C.byref(f.v), C.byref(target.v), -2.0 * self.fm_weight * FEATURE_MULTIPLICITY[j] / key[j], 1,
                       C.c_void_p(sums.data_ptr() + 8 * j), L.stream_ptr())
            return add

        adders = [adder(j, r, f) for j, ((_, r), (_, f)) in enumerate(zip(real_acts, fake_acts))]
        return (lambda: torch.dot(sums, self._fm_scale[1]).float()), adders

    # -- one iteration ------------------------------------------------------------------------------------------------------------------
    def step(self, real: torch.Tensor, real_labels: torch.Tensor, epoch: int = 0, noise: Optional[torch.Tensor] = None,
             fake_labels: Optional[torch.Tensor] = None, smooth_real: Optional[torch.Tensor] = None, smooth_fake: Optional[torch.Tensor] = None):
        """real: (N, nc, 224, 224) CUDA tensor; real_labels: (N,) int64.  The random draws of the reference (train_cgan.py:156-160,166-167) are
        made here unless given.  Returns a (7,) CUDA tensor [errD, errG, D_x, D_G_z1, D_G_z2, perceptual, feature matching] -- no host
        synchronisation before epoch 5."""
        netG, netD = self.netG, self.netD
        if not (netG.training and netD.training):
            raise L.B200GanError('CGANTrainer.step needs both networks in training mode')
        dev, n = real.device, real.shape[0]
        if smooth_real is None:
            smooth_real = REAL_LABEL - SMOOTH * torch.rand(n, device=dev)
        if smooth_fake is None:
            smooth_fake = FAKE_LABEL + SMOOTH * torch.rand(n, device=dev)
        if noise is None:
            noise = torch.randn(n, netG.latent_dim, device=dev)
        if fake_labels is None:
            fake_labels = torch.randint(0, netG.num_classes, (n,), device=dev)
        smooth_real, smooth_fake = smooth_real.float().contiguous(), smooth_fake.float().contiguous()
        # ---- D step
        logit_r, tape_r = self.engD.forward(netD, real, real_labels, save=True)
        m_real, dl_r = self._bce(logit_r, smooth_real)
        fake, tape_g = self.engG.forward(netG, noise, fake_labels, save=True, internal=True)      # stays in the engine's layout / dtype
        logit_f, tape_f = self.engD.forward(netD, fake, fake_labels, save=True)
        m_fake, dl_f = self._bce(logit_f, smooth_fake)
        stepped = True
        if epoch >= 5:                                        # train_cgan.py:176: `if D_x < 0.8 or D_G_z1 > 0.2 or epoch < 5`
            probs = torch.stack([m_real[1], m_fake[1]]).double()
            if self.comm is not None:
                self.comm.allreduce_f64(probs)
            d_x, d_g_z1 = (probs / self.world).tolist()
            stepped = d_x < 0.8 or d_g_z1 > 0.2
        if stepped:
            self.arenaD.grad.zero_()
            self.engD.backward(netD, tape_r, dl_r, None, need_dx=False, need_dw=True, into=self.intoD)
            self.engD.backward(netD, tape_f, dl_f, None, need_dx=False, need_dw=True, into=self.intoD)
            self._exchange(self.bucketsD)
            self._adam(self.arenaD)
            self.d_steps += 1
        del tape_r, tape_f
        # ---- G step
        logit_g, tape_a = self.engD.forward(netD, fake, fake_labels, save=True)
        m_adv, dl_g = self._bce(logit_g, smooth_real)
        _, tape_fr = self.engD.forward(netD, real, real_labels, save=True, head=False)
        self.engD.replay_running_stats(netD, tape_a)          # the reference's features pass over the fake batch
        fm_loss, adders = self._feature_matching(tape_fr[0], tape_a[0])
        dfake = Act(torch.empty_like(fake.t), nchw=False)
        self.engD.backward(netD, tape_a, dl_g, adders, need_dx=True, need_dw=False, dx_out=dfake)
        l_fm = fm_loss()                                      # every pair has been visited by the backward pass
        l_p = torch.zeros((), device=dev)
        if self.perceptual is not None and self.perceptual_weight != 0.0:      # 10 x perceptual(fake, real), train_cgan.py:186,191: its gradient joins dfake
            self.perceptual.compute_dtype = self.engG.dtype
            l_p = self.perceptual._engine_for().loss_and_grad(fake, real, self.perceptual_weight, dx_into=dfake)
        self.arenaG.grad.zero_()
        self.engG.backward(netG, tape_g, dfake, need_dz=False, into=self.intoG)
        self._exchange(self.bucketsG)
        self._adam(self.arenaG)
        return torch.stack([m_real[0] + m_fake[0], m_adv[0] + self.perceptual_weight * l_p + self.fm_weight * l_fm, m_real[1], m_fake[1], m_adv[1], l_p, l_fm])
