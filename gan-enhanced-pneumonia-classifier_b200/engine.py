"""Host-side orchestration of the DCGAN Generator / Discriminator forward and backward on the B200 kernels.

Everything numeric happens in libb200gan.so (hand-written sm_100a CUDA, see csrc/); this file only sequences
the C-ABI calls per layer, owns the activation buffers (torch tensors used as plain device memory) and
exposes the result to autograd through two `torch.autograd.Function`s, so that the reference's own training
loop (`train_gan.py:119-150`: `netD(real)`, `errD.backward()`, `netD(fake.detach())`, `netD(fake)` ...) runs
unchanged on top of it.

Reference behaviour mirrored here (file:line in /root/reference/src):
  dcgan.py:25-48   Generator.main      ConvT(k7,s1,p0)-BN-ReLU, 4x[ConvT(k4,s2,p1)-BN-ReLU], ConvT(k4,s2,p1)-Tanh
  dcgan.py:64-86   Discriminator.main  Conv(k4,s2,p1)-LReLU, 4x[Conv(k4,s2,p1)-BN-LReLU], Conv(k7,s1,p0)-Sigmoid
  dcgan.py:89-90   Discriminator.forward flattens to (N,)
BatchNorm side effects (running stats with unbiased variance, num_batches_tracked) happen in every training-mode
forward, including forwards under no_grad (train_gan.py:166-169).

Internal layout: activations are NHWC in the compute dtype (float32 for the parity mode, bfloat16 for the
tensor-core mode); the reference's NCHW fp32 tensors are read and written in place through strided views at the
network edges, so no layout-conversion pass exists.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import _lib as L

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
LRELU_SLOPE = 0.2


def default_compute_dtype() -> torch.dtype:
    """bf16 (tensor-core mode) unless B200GAN_DTYPE=fp32 selects the fp32 parity mode."""
    v = os.environ.get('B200GAN_DTYPE', 'bf16').lower()
    if v in ('fp32', 'float32', 'f32'):
        return torch.float32
    if v in ('bf16', 'bfloat16'):
        return torch.bfloat16
    raise ValueError(f'B200GAN_DTYPE={v!r}: expected fp32 or bf16')


def default_algo() -> int:
    v = os.environ.get('B200GAN_ALGO', 'auto').lower()
    return {'auto': L.ALGO_AUTO, 'simt': L.ALGO_SIMT, 'tcgen05': L.ALGO_TCGEN05}[v]


@dataclass
class LayerSpec:
    conv_idx: int
    bn_idx: Optional[int]
    cin: int
    cout: int
    k: int
    stride: int
    pad: int
    act: int


def generator_specs(latent_dim: int, nc: int, ngf: int) -> List[LayerSpec]:
    ch = [latent_dim, ngf * 8, ngf * 4, ngf * 2, ngf, ngf // 2, nc]
    out = []
    for i in range(6):
        k, s, p = (7, 1, 0) if i == 0 else (4, 2, 1)
        out.append(LayerSpec(3 * i, 3 * i + 1 if i < 5 else None, ch[i], ch[i + 1], k, s, p,
                             L.ACT_RELU if i < 5 else L.ACT_TANH))
    return out


def discriminator_specs(nc: int, ndf: int) -> List[LayerSpec]:
    ch = [nc, ndf // 2, ndf, ndf * 2, ndf * 4, ndf * 8, 1]
    conv_idx = [0, 2, 5, 8, 11, 14]
    out = []
    for i in range(6):
        k, s, p = (7, 1, 0) if i == 5 else (4, 2, 1)
        out.append(LayerSpec(conv_idx[i], conv_idx[i] + 1 if 1 <= i <= 4 else None, ch[i], ch[i + 1], k, s, p,
                             L.ACT_LRELU if i < 5 else L.ACT_SIGMOID))
    return out


class Act:
    """A device tensor plus the b200gan_view describing it as logical (N,H,W,C)."""
    __slots__ = ('t', 'v', 'nchw')

    def __init__(self, t: torch.Tensor, nchw: bool):
        self.t = t
        self.nchw = nchw
        self.v = L.view_nchw(t) if nchw else L.view_nhwc(t)

    @property
    def n(self):
        return self.v.n


class LayerParams:
    __slots__ = ('w', 'gamma', 'beta', 'rm', 'rv', 'nbt')

    def __init__(self, w, gamma=None, beta=None, rm=None, rv=None, nbt=None):
        self.w, self.gamma, self.beta, self.rm, self.rv, self.nbt = w, gamma, beta, rm, rv, nbt


def params_from_module(module, specs) -> List[LayerParams]:
    out = []
    for sp in specs:
        conv = module.main[sp.conv_idx]
        if sp.bn_idx is None:
            out.append(LayerParams(conv.weight))
        else:
            bn = module.main[sp.bn_idx]
            out.append(LayerParams(conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked))
    return out


def _check_f32(t, what):
    if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
        raise L.B200GanError(f'{what} must be a contiguous float32 CUDA tensor (got {t.dtype}, contiguous={t.is_contiguous()}, {t.device})')


class LayerCtx:
    """What the forward pass keeps per layer: input x, conv output y (BatchNorm layers only), activation output a,
    BatchNorm coefficients, packed weights, and (filled during backward) the BatchNorm-backward sums."""
    __slots__ = ('x', 'y', 'a', 'scale', 'shift', 'mean', 'invstd', 'wp_down', 'wp_up', 'bsums', 'bsums2')


class NetEngine:
    """Forward / backward of one network (transposed=True: Generator, False: Discriminator)."""

    def __init__(self, specs: List[LayerSpec], transposed: bool, dtype: torch.dtype, algo: int = L.ALGO_AUTO):
        self.specs, self.transposed, self.dtype, self.algo = specs, transposed, dtype, algo
        self._conv = [L.Conv(sp.k, sp.stride, sp.pad, algo) for sp in specs]
        self.launches = 0     # C-ABI calls that launch kernels, made through this engine (bench.py's gpu_launches claim)
        self._ws = None       # zeroed fp32 scratch for the split-K reduction of the tensor-core weight gradients
        # Optional second stream for the weight-gradient launches (set by the fused trainer): nothing on the way down depends on
        # them (the next layer needs only the input gradient), so they leave the backward pass's critical path and co-run with its
        # HBM-bound BatchNorm passes.  `join_wgrads()` is the point where the compute stream waits for them.
        self.wgrad_stream = None
        # Synchronised BatchNorm (data parallel, --sync-bn): a dp.DPComm; the per-channel sums of every BatchNorm (forward statistics and
        # backward reductions) are summed over the ranks before they are used, and the sample count becomes the global one -- exactly the
        # arithmetic of one process on the concatenated batch.  None: statistics are local to the rank.
        self.sync_bn = None
        self._side_keep = []  # operands of weight gradients still in flight on wgrad_stream (kept from the caching allocator)
        self.weights_version = None   # see _pack
        self._packed = {}

    # -- thin wrappers over the C ABI --------------------------------------------------------------
    def _tc_layer(self, i):
        """Layers whose convolutions qualify for the tcgen05 implicit GEMM (bf16, k4 s2 p1, channels % 32 == 0)."""
        sp = self.specs[i]
        return (self.dtype == torch.bfloat16 and self.algo != L.ALGO_SIMT and sp.k == 4 and sp.stride == 2 and sp.pad == 1
                and sp.cin % 32 == 0 and sp.cout % 32 == 0)

    def _act_fused(self, i):
        """Layers without BatchNorm whose activation rides on the convolution (b200gan_fuse.out_act / dy_act): the k4 s2 p1
        image-side layers D0 (LeakyReLU) and G5 (Tanh)."""
        sp = self.specs[i]
        return sp.bn_idx is None and sp.k == 4 and sp.act in (L.ACT_LRELU, L.ACT_RELU, L.ACT_TANH)

    def _pack(self, i, w, st):
        """bf16 GEMM-operand repacks of the fp32 master weight: ('down' form, 'up' form), see b200gan_pack_conv_weight.
        One launch produces both.  While `weights_version` is not None the result is reused until the owner of the weights
        bumps the version (DCGANTrainer does after each Adam update: the two Discriminator forwards of the D step share one
        repack); with version None (module-level autograd path, weights may change behind our back) every forward repacks."""
        if not self._tc_layer(i):
            return None, None
        n = w.numel()
        if self.weights_version is None:
            both = torch.empty(2 * n, device=w.device, dtype=torch.bfloat16)
        else:
            # ONE persistent buffer per layer (fixed address): a captured CUDA graph may read, in its first Discriminator forward,
            # the repack its own previous replay wrote after the Adam update
            hit = self._packed.get(i)
            if hit is None or hit[1].device != w.device or hit[1].numel() != 2 * n:
                hit = self._packed[i] = [None, torch.empty(2 * n, device=w.device, dtype=torch.bfloat16)]
            both = hit[1]
            if hit[0] == self.weights_version:
                return both[:n], both[n:]
            hit[0] = self.weights_version
        L.call('b200gan_pack_conv_weight', L.ptr(w), w.shape[0], w.shape[1], 4, 2, L.ptr(both), st)
        self.launches += 1
        return both[:n], both[n:]

    # -- 64-channel image-side layers (the WGAN-GP critic's Conv2d(nc -> 64) and generator's ConvTranspose2d(64 -> nc)) ---------------
    # The warp-MMA kernels of the image-side layers are built for a 32-channel feature side.  A 64-channel one is served as two 32-channel
    # SLICES of the same tensor (the TMA-staged kernels take the slice's strides in their tensor map): the "up" direction sums two fp32 partial
    # images before the activation, the weight gradient writes the two halves of dw directly.  Measured at batch 512: 3.2 ms -> 0.4 ms ("up"),
    # 1.15 ms -> 0.3 ms (weight gradient) against the fp32-FMA kernels of conv_edge.cu, which remain the path for fused gradient masks.
    def _split64(self, coarse: Act, fine: Act, fuse_has_coarse_ref: bool) -> bool:
        sp_ok = self.dtype == torch.bfloat16 and self.algo != L.ALGO_SIMT
        return (sp_ok and coarse.v.c == 64 and fine.v.c in (1, 3) and not fuse_has_coarse_ref and not coarse.nchw and coarse.t.dtype == torch.bfloat16
                and coarse.t.is_contiguous() and coarse.v.w % 16 == 0 and coarse.v.w <= 112 and fine.v.h == 2 * coarse.v.h)

    def _up64(self, name, i, coarse: Act, w, fine: Act, st, out_act):
        """fine = out_act(up(coarse[..., :32], w[:32]) + up(coarse[..., 32:], w[32:]))."""
        parts = []
        for h in range(2):
            t = Act(torch.empty((fine.v.n, fine.v.h, fine.v.w, fine.v.c), device=coarse.t.device, dtype=torch.float32), nchw=False)
            cs, ws = Act(coarse.t[..., 32 * h:32 * h + 32], nchw=False), w[32 * h:32 * h + 32]
            L.call(name, C.byref(self._conv[i]), C.byref(cs.v), L.ptr(ws), None, C.byref(t.v), None, st)
            parts.append(t)
        L.call('b200gan_sample_axpby', C.byref(parts[0].v), None, C.byref(parts[1].v), None, C.byref(parts[0].v), st)
        L.call('b200gan_bn_act_fwd', C.byref(parts[0].v), None, None, out_act, LRELU_SLOPE, C.byref(fine.v), st)
        self.launches += 4

    def _fprop(self, i, x: Act, w, y: Act, st, wp_down=None, wp_up=None, fuse=None):
        # ConvTranspose2d forward is the 'up' geometry, Conv2d forward the 'down' geometry
        name, wp = ('b200gan_convT2d_fprop', wp_up) if self.transposed else ('b200gan_conv2d_fprop', wp_down)
        if self.transposed and (fuse is None or not fuse.bn_sums) and self._conv[i].k == 4 and self._split64(x, y, False):
            return self._up64(name, i, x, w, y, st, fuse.out_act if fuse is not None else L.ACT_NONE)
        if not self.transposed and (fuse is None or not fuse.bn_sums) and self._conv[i].k == 4 and self._split64(y, x, False):
            for h in range(2):                       # "down" into the two 32-channel halves of the result (activation epilogue included)
                ys = Act(y.t[..., 32 * h:32 * h + 32], nchw=False)
                L.call(name, C.byref(self._conv[i]), C.byref(x.v), L.ptr(w[32 * h:32 * h + 32]), None, C.byref(ys.v), C.byref(fuse) if fuse is not None else None, st)
            self.launches += 2
            return
        L.call(name, C.byref(self._conv[i]), C.byref(x.v), L.ptr(w), L.ptr(wp), C.byref(y.v), C.byref(fuse) if fuse is not None else None, st)
        self.launches += 1

    def _dgrad(self, i, dy: Act, w, dx: Act, st, wp_down=None, wp_up=None, fuse=None):
        name, wp = ('b200gan_convT2d_dgrad', wp_down) if self.transposed else ('b200gan_conv2d_dgrad', wp_up)
        if not self.transposed and fuse is None and self._conv[i].k == 4 and self._split64(dy, dx, False):
            return self._up64(name, i, dy, w, dx, st, L.ACT_NONE)
        L.call(name, C.byref(self._conv[i]), C.byref(dy.v), L.ptr(w), L.ptr(wp), C.byref(dx.v), C.byref(fuse) if fuse is not None else None, st)
        self.launches += 1

    def _wgrad(self, i, x: Act, dy: Act, dw, st, fuse=None):
        name = 'b200gan_convT2d_wgrad' if self.transposed else 'b200gan_conv2d_wgrad'
        coarse, fine = (x, dy) if self.transposed else (dy, x)
        # a fused activation backward on the COARSE gradient (the critic's first layer) keeps the one-call path; on the fine side (tanh of the
        # generator's last layer) it rides along with the slices
        coarse_ref = fuse is not None and bool(fuse.dy_ref) and not self.transposed
        if self._conv[i].k == 4 and self._split64(coarse, fine, coarse_ref):
            for h in range(2):
                cs = Act(coarse.t[..., 32 * h:32 * h + 32], nchw=False)
                xa, da = (cs, dy) if self.transposed else (x, cs)
                L.call(name, C.byref(self._conv[i]), C.byref(xa.v), C.byref(da.v), L.ptr(dw[32 * h:32 * h + 32]), None,
                       C.byref(fuse) if fuse is not None else None, st)
            self.launches += 2
            return
        ws = None
        need = int(L.load().b200gan_conv_wgrad_workspace_floats(C.byref(self._conv[i]), C.byref(x.v), C.byref(dy.v), 1 if self.transposed else 0))
        if need > 0:
            # all zero on entry, handed back all zero by the library: one buffer serves every layer
            if self._ws is None or self._ws.numel() < need or self._ws.device != dw.device:
                self._ws = torch.zeros(max(need, max(sp.cin * sp.cout * sp.k * sp.k for sp in self.specs)), device=dw.device, dtype=torch.float32)
            ws = self._ws
        L.call(name, C.byref(self._conv[i]), C.byref(x.v), C.byref(dy.v), L.ptr(dw), L.ptr(ws), C.byref(fuse) if fuse is not None else None, st)
        self.launches += 1

    def out_hw(self, i, h, w):
        sp = self.specs[i]
        if self.transposed:
            return (h - 1) * sp.stride - 2 * sp.pad + sp.k, (w - 1) * sp.stride - 2 * sp.pad + sp.k
        return (h + 2 * sp.pad - sp.k) // sp.stride + 1, (w + 2 * sp.pad - sp.k) // sp.stride + 1

    # -- forward -----------------------------------------------------------------------------------
    def forward(self, x: Act, params: List[LayerParams], training: bool, save: bool, out: Optional[Act] = None,
                last_act: bool = True):
        """Runs all six layers.  `out` optionally receives the final activation (e.g. the fp32 NCHW tensor
        handed back to torch); with last_act=False the last layer's pre-activation is returned instead
        (the Discriminator logits, so that Sigmoid can be fused with the BCE loss).
        Returns (final Act, [LayerCtx] or None)."""
        st = L.stream_ptr()
        dev = x.t.device
        ctxs = [] if save else None
        cur = x
        nl = len(self.specs)
        for i, sp in enumerate(self.specs):
            p = params[i]
            _check_f32(p.w, 'conv weight')
            oh, ow = self.out_hw(i, cur.v.h, cur.v.w)
            last = i == nl - 1
            ydt = torch.float32 if (last and not self.transposed) else self.dtype      # D logits stay fp32
            shape = (cur.v.n, oh, ow, sp.cout)
            wp_down, wp_up = self._pack(i, p.w, st)
            lc = LayerCtx() if save else None
            if save:
                lc.x, lc.y, lc.wp_down, lc.wp_up = cur, None, wp_down, wp_up
                lc.scale = lc.shift = lc.mean = lc.invstd = lc.bsums = None
            if sp.bn_idx is not None:
                # conv (+ BatchNorm statistics in its epilogue) -> finalize -> normalise + activation
                cch = sp.cout
                y = Act(torch.empty(shape, device=dev, dtype=self.dtype), nchw=False)
                scale = torch.empty(cch, device=dev, dtype=torch.float32)
                shift = torch.empty(cch, device=dev, dtype=torch.float32)
                if training:
                    sums = torch.empty(2 * cch, device=dev, dtype=torch.float64)
                    mean = torch.empty(cch, device=dev, dtype=torch.float32)
                    invstd = torch.empty(cch, device=dev, dtype=torch.float32)
                    self._fprop(i, cur, p.w, y, st, wp_down, wp_up, fuse=L.fuse(bn_sums=sums))
                    ranks = 1
                    fused_apply = self.sync_bn is None and os.environ.get('B200GAN_BN_ONE_LAUNCH', '1') != '0'
                    if self.sync_bn is not None:
                        self.sync_bn.allreduce_f64(sums)
                        ranks = self.sync_bn.world
                    if fused_apply:
                        # finalize + normalise + activation in one launch (local statistics): one launch less per BatchNorm layer
                        a = Act(torch.empty(shape, device=dev, dtype=self.dtype), nchw=False)
                        L.call('b200gan_bn_finalize_act_fwd', L.ptr(sums), cch, cur.v.n * oh * ow, L.ptr(p.gamma), L.ptr(p.beta),
                               L.ptr(p.rm), L.ptr(p.rv), L.ptr(p.nbt), BN_MOMENTUM, BN_EPS, L.ptr(scale), L.ptr(shift),
                               L.ptr(mean), L.ptr(invstd), C.byref(y.v), sp.act, LRELU_SLOPE, C.byref(a.v), st)
                        self.launches += 1
                    else:
                        L.call('b200gan_bn_finalize', L.ptr(sums), cch, cur.v.n * oh * ow * ranks, L.ptr(p.gamma), L.ptr(p.beta),
                               L.ptr(p.rm), L.ptr(p.rv), L.ptr(p.nbt), BN_MOMENTUM, BN_EPS, L.ptr(scale), L.ptr(shift),
                               L.ptr(mean), L.ptr(invstd), st)
                    if save:
                        lc.mean, lc.invstd = mean, invstd
                else:
                    self._fprop(i, cur, p.w, y, st, wp_down, wp_up)
                    fused_apply = False
                    L.call('b200gan_bn_eval_coeffs', cch, L.ptr(p.gamma), L.ptr(p.beta), L.ptr(p.rm), L.ptr(p.rv), BN_EPS,
                           L.ptr(scale), L.ptr(shift), st)
                if not fused_apply:
                    a = Act(torch.empty(shape, device=dev, dtype=self.dtype), nchw=False)
                    L.call('b200gan_bn_act_fwd', C.byref(y.v), L.ptr(scale), L.ptr(shift), sp.act, LRELU_SLOPE, C.byref(a.v), st)
                    self.launches += 2
                if save:
                    lc.y, lc.scale, lc.shift = y, scale, shift
            elif last and not last_act:
                a = Act(torch.empty(shape, device=dev, dtype=ydt), nchw=False)          # logits
                self._fprop(i, cur, p.w, a, st, wp_down, wp_up)
            elif self._act_fused(i):
                a = out if (last and out is not None) else Act(torch.empty(shape, device=dev, dtype=ydt), nchw=False)
                self._fprop(i, cur, p.w, a, st, wp_down, wp_up, fuse=L.fuse(out_act=sp.act, out_slope=LRELU_SLOPE))
            else:
                y = Act(torch.empty(shape, device=dev, dtype=ydt), nchw=False)
                self._fprop(i, cur, p.w, y, st, wp_down, wp_up)
                a = out if (last and out is not None) else Act(torch.empty(shape, device=dev, dtype=ydt), nchw=False)
                L.call('b200gan_bn_act_fwd', C.byref(y.v), None, None, sp.act, LRELU_SLOPE, C.byref(a.v), st)
                self.launches += 1
            if save:
                lc.a = a
                ctxs.append(lc)
            cur = a
        return cur, ctxs

    # -- backward ----------------------------------------------------------------------------------
    def backward(self, ctxs: List[LayerCtx], params: List[LayerParams], dout: Optional[Act], grads: List[Optional[torch.Tensor]],
                 dinput: Optional[Act] = None, need_wgrad: bool = True, dlogit: Optional[Act] = None, on_ready=None):
        """`dout`: gradient w.r.t. the final activation (or `dlogit`: w.r.t. the last conv output, when the
        loss kernel already applied Sigmoid').  `grads` is the flat list over `param_order()` of fp32 tensors
        that are ACCUMULATED into (autograd semantics); entries may be None to skip.  `dinput`, when given,
        receives the gradient w.r.t. the network input.  `on_ready(j)` is called as soon as the launches that finish the
        gradient of parameter j (index into `param_order()`) have been issued (data-parallel bucket all-reduce, dp.py).

        Per layer, going down: the incoming gradient d is w.r.t. the layer's activation output.
          * BatchNorm layers: the input-gradient convolution of the layer ABOVE already applied the activation backward and
            accumulated the BatchNorm-backward sums in its epilogue (b200gan_fuse.prev_*), so d is dz; one pass finishes
            native_batch_norm_backward in place (dy, dgamma, dbeta);
          * D0 / G5 (no BatchNorm): the activation backward rides on the operand read of wgrad / dgrad (dy_act, dy_ref)."""
        st = L.stream_ptr()
        nl = len(self.specs)
        gi = len(grads)
        d = dout
        masked = False        # d already carries this layer's activation backward (applied by the dgrad epilogue of the layer above)
        for i in reversed(range(nl)):
            sp, p, lc = self.specs[i], params[i], ctxs[i]
            dev = lc.a.t.device
            nparam = 3 if sp.bn_idx is not None else 1
            gi -= nparam
            fuse_kw = {}
            if i == nl - 1 and dlogit is not None:
                dy = dlogit
            elif sp.bn_idx is not None:
                cnt = lc.y.v.n * lc.y.v.h * lc.y.v.w
                dg = grads[gi + 1] if need_wgrad else None
                db = grads[gi + 2] if need_wgrad else None
                if self.sync_bn is not None:
                    # the input gradient needs the GLOBAL sums; dgamma / dbeta stay this rank's share (the gradient exchange sums them)
                    local = lc.bsums.clone()
                    self.sync_bn.allreduce_f64(lc.bsums)
                    cnt *= self.sync_bn.world
                    cch = sp.cout
                    if dg is not None:
                        L.call('b200gan_accumulate_2d', L.ptr(dg), L.ptr(local[cch:]), 1, 1, cch, 0, 1, st)
                        L.call('b200gan_accumulate_2d', L.ptr(db), L.ptr(local), 1, 1, cch, 0, 1, st)
                        self.launches += 2
                    dg = db = None
                L.call('b200gan_bn_act_bwd_apply', C.byref(d.v), C.byref(lc.y.v), None, L.ptr(lc.scale), L.ptr(lc.shift),
                       L.ptr(lc.mean), L.ptr(lc.invstd), L.ptr(p.gamma), L.ptr(lc.bsums), cnt, L.ACT_NONE, LRELU_SLOPE,
                       C.byref(d.v), L.ptr(dg), L.ptr(db), st)
                self.launches += 1
                dy = d
            elif self._act_fused(i):
                dy = d
                if not masked:
                    fuse_kw = dict(dy_act=sp.act, dy_slope=LRELU_SLOPE, dy_ref=lc.a.v)
            else:
                # activation without BatchNorm and without a fused kernel (Sigmoid of the module-level Discriminator forward)
                dy = Act(torch.empty_like(lc.a.t), nchw=lc.a.nchw)          # same layout as the saved output (NCHW at the network edge)
                L.call('b200gan_bn_act_bwd_apply', C.byref(d.v), C.byref(lc.a.v), C.byref(lc.a.v), None, None, None, None, None, None, 0,
                       sp.act, LRELU_SLOPE, C.byref(dy.v), None, None, st)
                self.launches += 1
            side = self.wgrad_stream if (need_wgrad and grads[gi] is not None) else None
            if side is not None:
                side.wait_stream(torch.cuda.current_stream())      # dy (and dgamma / dbeta of this layer) are final
                with torch.cuda.stream(side):
                    self._wgrad(i, lc.x, dy, grads[gi], L.stream_ptr(), fuse=L.fuse(**fuse_kw) if fuse_kw else None)
                    if on_ready is not None:                       # the bucket's all-reduce forks from the stream the gradient is on
                        for j in range(gi + nparam - 1, gi - 1, -1):
                            on_ready(j)
                self._side_keep.append((lc, dy))
            else:
                if need_wgrad and grads[gi] is not None:
                    self._wgrad(i, lc.x, dy, grads[gi], st, fuse=L.fuse(**fuse_kw) if fuse_kw else None)
                if need_wgrad and on_ready is not None:
                    for j in range(gi + nparam - 1, gi - 1, -1):
                        on_ready(j)
            masked = False
            if i > 0:
                below, lb = self.specs[i - 1], ctxs[i - 1]
                if below.bn_idx is not None:
                    lb.bsums = torch.empty(2 * below.cout, device=dev, dtype=torch.float64)
                    fuse_kw.update(prev_act=below.act, prev_slope=LRELU_SLOPE, prev_y=lb.y.v, prev_scale=lb.scale, prev_shift=lb.shift,
                                   prev_mean=lb.mean, prev_invstd=lb.invstd, prev_sums=lb.bsums)
                elif self._act_fused(i - 1) and below.act in (L.ACT_LRELU, L.ACT_RELU) and (need_wgrad or dinput is not None):
                    # D0 (LeakyReLU, no BatchNorm): its activation backward rides on THIS layer's dgrad epilogue, so D0's own
                    # wgrad / dgrad read a plain gradient tensor (measured at B=512: the mask on D0's wgrad operand instead sends
                    # that kernel down its register-staged variant, 299 us against 120 us, more than the epilogue costs here)
                    fuse_kw.update(prev_act=below.act, prev_slope=LRELU_SLOPE, prev_y=lb.a.v)
                    masked = True
                d = Act(torch.empty(lc.x.t.shape, device=dev, dtype=self.dtype), nchw=False)
                self._dgrad(i, dy, p.w, d, st, lc.wp_down, lc.wp_up, fuse=L.fuse(**fuse_kw) if fuse_kw else None)
            elif dinput is not None:
                self._dgrad(i, dy, p.w, dinput, st, lc.wp_down, lc.wp_up, fuse=L.fuse(**fuse_kw) if fuse_kw else None)
        return dinput

    def join_wgrads(self):
        """Make the current stream wait for the weight gradients launched on `wgrad_stream` (call before the gradients are read)."""
        if self.wgrad_stream is not None:
            torch.cuda.current_stream().wait_stream(self.wgrad_stream)
        self._side_keep.clear()

    def param_order(self, module):
        """The parameters in `module.parameters()` order: conv weight, then BN weight, bias per layer."""
        out = []
        for sp in self.specs:
            out.append(module.main[sp.conv_idx].weight)
            if sp.bn_idx is not None:
                out += [module.main[sp.bn_idx].weight, module.main[sp.bn_idx].bias]
        return out


def _forward_raw(module, engine: NetEngine, x: torch.Tensor, save: bool):
    """Shared by the autograd and the no-grad paths: fp32 NCHW in, fp32 NCHW out (the reference's interface)."""
    params = params_from_module(module, engine.specs)
    training = module.training
    if save and not training:
        raise L.B200GanError('backward through an eval-mode network is not on the DCGAN training path and is not implemented')
    xin = x.detach()
    if xin.dtype != torch.float32:
        xin = xin.float()
    if xin.dim() != 4 or xin.shape[1] != engine.specs[0].cin:
        raise L.B200GanError(f'expected input of shape (N,{engine.specs[0].cin},H,W), got {tuple(xin.shape)}')
    n, h, w = xin.shape[0], xin.shape[2], xin.shape[3]
    for i in range(len(engine.specs)):
        h, w = engine.out_hw(i, h, w)
        if h < 1 or w < 1:
            raise L.B200GanError(f'input spatial size {tuple(xin.shape[2:])} is too small for this network (the reference is hard-wired to 224x224)')
    out_t = torch.empty((n, engine.specs[-1].cout, h, w), device=x.device, dtype=torch.float32)
    _, ctxs = engine.forward(Act(xin, nchw=True), params, training, save, out=Act(out_t, nchw=True))
    return out_t, ctxs


class _NetFunction(torch.autograd.Function):
    """autograd bridge: forward/backward of a whole network through the C ABI."""

    @staticmethod
    def forward(ctx, module, engine: NetEngine, x, *flat_params):
        out_t, ctxs = _forward_raw(module, engine, x, save=True)
        ctx.engine, ctx.module, ctx.ctxs, ctx.xshape = engine, module, ctxs, tuple(x.shape)
        return out_t

    @staticmethod
    def backward(ctx, dout):
        engine, module, ctxs = ctx.engine, ctx.module, ctx.ctxs
        if ctxs is None:
            raise L.B200GanError('backward called twice on the same forward (activations were released)')
        params = params_from_module(module, engine.specs)
        plist = engine.param_order(module)
        needs = ctx.needs_input_grad        # (module, engine, x, *params)
        grads = [torch.zeros_like(p, dtype=torch.float32) if needs[3 + j] else None for j, p in enumerate(plist)]
        dout = dout.contiguous()
        if dout.dtype != torch.float32:
            dout = dout.float()
        dx = dinput = None
        if needs[2]:
            dx = torch.empty(ctx.xshape, device=dout.device, dtype=torch.float32)
            dinput = Act(dx, nchw=True)
        engine.backward(ctxs, params, Act(dout, nchw=True), grads, dinput=dinput,
                        need_wgrad=any(g is not None for g in grads))
        ctx.ctxs = None
        return (None, None, dx, *grads)


def run_network(module, engine: NetEngine, x: torch.Tensor) -> torch.Tensor:
    """What `Generator.forward` / `Discriminator.forward` call for CUDA inputs."""
    plist = engine.param_order(module)
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in plist)):
        return _NetFunction.apply(module, engine, x, *plist)
    return _forward_raw(module, engine, x, save=False)[0]
