"""Host-side orchestration of the conditional GAN's Generator / Discriminator (reference src/cgan.py) on the
kernels of libb200gan.so.  Like engine.py this file only sequences C-ABI calls and owns the activation buffers; nothing numeric happens in
torch (the few torch calls are allocation and weight-layout plumbing: a transpose of the `fc` weight, slices of gradient buffers).

How the reference's layers map onto the library (file:line in /root/reference/src/cgan.py):
  :22,55-56  z + label_emb(labels)            b200gan_embed_add (writes [x, 1]: the constant feature carries fc's bias through the GEMM)
  :24,57-58  fc: Linear(nz -> 8nf*7*7) + view  b200gan_convT2d_* with k=7,s=1,p=0 on the 1x1 "image": a Linear whose output is viewed as
                                              (8nf,7,7) IS that transposed convolution with weight fc.weight.T viewed (nz, 8nf, 7, 7)
  :26-27     BatchNorm2d + ReLU on fc's output b200gan_bn_stats / bn_finalize / bn_act_fwd
  :28-49     5 x [Upsample(2), Conv2d(3,1,1)+bias]   ONE stride-2 transposed convolution each (b200gan_upconv3_fold folds the 3x3 weight to
                                              the ConvTranspose2d(4,2,1) form; csrc/cgan_ops.cu has the algebra): the wide layers run on the
                                              tcgen05 kernels of the DCGAN Generator, no upsampled tensor is ever written
  :70-89     Conv2d(4,2,1)+bias [+BN] + LeakyReLU    the DCGAN Discriminator's convolutions; the bias is a per-channel shift pass
                                              (b200gan_bn_act_fwd with unit scale), its gradient a channel sum (b200gan_bn_stats)
  :91        Conv2d(8nf -> 1, 7,1,0)+bias      b200gan_conv2d_* k=7
  :103-106   projection: out + <label_emb(labels), features>   b200gan_class_proj_fwd / _bwd
  :108-113   get_intermediate_features         the same forward, every intermediate handed out (and taking gradients back in)
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib as L
from .engine import Act, BN_EPS, BN_MOMENTUM, LRELU_SLOPE, default_algo, default_compute_dtype

INIT_SIZE = 7            # cgan.py:18


def _st():
    return L.stream_ptr()


class _ConvOp:
    """One convolution geometry + the bf16 operand repack the tcgen05 kernels want when the layer qualifies for them."""

    def __init__(self, k, stride, pad, transposed, cin, cout, dtype, algo):
        self.transposed, self.cin, self.cout = transposed, cin, cout
        self.desc = L.Conv(k, stride, pad, algo)
        self.tc = (dtype == torch.bfloat16 and algo != L.ALGO_SIMT and k == 4 and stride == 2 and pad == 1 and cin % 32 == 0 and cout % 32 == 0)
        self._ws = None

    def pack(self, w):
        """('down' form, 'up' form) of the fp32 master `w`, or (None, None).  Repacked per call: at module level the weights may change
        between any two calls (optimizer.step(), load_state_dict)."""
        if not self.tc:
            return None, None
        n = w.numel()
        both = torch.empty(2 * n, device=w.device, dtype=torch.bfloat16)
        L.call('b200gan_pack_conv_weight', L.ptr(w), w.shape[0], w.shape[1], 4, 2, L.ptr(both), _st())
        return both[:n], both[n:]

    def fprop(self, x: Act, w, y: Act, packs):
        name, wp = ('b200gan_convT2d_fprop', packs[1]) if self.transposed else ('b200gan_conv2d_fprop', packs[0])
        L.call(name, C.byref(self.desc), C.byref(x.v), L.ptr(w), L.ptr(wp), C.byref(y.v), None, _st())

    def dgrad(self, dy: Act, w, dx: Act, packs):
        name, wp = ('b200gan_convT2d_dgrad', packs[0]) if self.transposed else ('b200gan_conv2d_dgrad', packs[1])
        L.call(name, C.byref(self.desc), C.byref(dy.v), L.ptr(w), L.ptr(wp), C.byref(dx.v), None, _st())

    def wgrad(self, x: Act, dy: Act, dw):
        """dw += (accumulating, like autograd)."""
        coarse = x if self.transposed else dy
        if self.tc and coarse.v.c == 32:
            # The tensor-core weight-gradient kernel tiles the coarse side's channels by 64.  A 32-channel coarse operand is copied into the
            # lower half of a zeroed 64-channel tensor (one elementwise pass) instead of dropping to the generic SIMT kernel (measured at batch
            # 1024: 13.7 ms per call on that kernel); the coarse channels are dimension 0 of dw in both geometries, so the real block of the
            # padded gradient is its leading half.
            wide = _nhwc(coarse.v.n, coarse.v.h, coarse.v.w, 64, coarse.t.device, coarse.t.dtype, zero=True)
            L.call('b200gan_copy_view', C.byref(coarse.v), C.byref(Act(wide.t[..., :32], nchw=False).v), _st())
            dwp = torch.zeros((64,) + tuple(dw.shape[1:]), device=dw.device, dtype=torch.float32)
            self._wgrad_call(wide if self.transposed else x, dy if self.transposed else wide, dwp)
            L.call('b200gan_accumulate_2d', L.ptr(dw), L.ptr(dwp), 0, 1, dw.numel(), 0, 1, _st())
            return
        self._wgrad_call(x, dy, dw)

    def _wgrad_call(self, x: Act, dy: Act, dw):
        name = 'b200gan_convT2d_wgrad' if self.transposed else 'b200gan_conv2d_wgrad'
        need = int(L.load().b200gan_conv_wgrad_workspace_floats(C.byref(self.desc), C.byref(x.v), C.byref(dy.v), 1 if self.transposed else 0))
        ws = None
        if need > 0:
            if self._ws is None or self._ws.numel() < need or self._ws.device != dw.device:
                self._ws = torch.zeros(need, device=dw.device, dtype=torch.float32)       # zero on entry, handed back zero
            ws = self._ws
        L.call(name, C.byref(self.desc), C.byref(x.v), C.byref(dy.v), L.ptr(dw), L.ptr(ws), None, _st())


class _BNState:
    __slots__ = ('scale', 'shift', 'mean', 'invstd', 'sums')


def _bn_forward(y: Act, bn, act, out: Act, training: bool) -> _BNState:
    """BatchNorm2d + activation (training: batch statistics, running-stat side effects; eval: running statistics).  y / out may be stored
    wider than the BatchNorm (zero-padded channels, see _padded): the statistics pass runs over the stored width, the finalize step over the
    real channels, and the padded channels get scale = shift = 0, so they stay exactly zero through the dense apply pass."""
    dev, p, c = y.t.device, y.v.c, bn.weight.shape[0]
    s = _BNState()
    coef = torch.zeros(4 * p, device=dev, dtype=torch.float32) if p != c else torch.empty(4 * c, device=dev, dtype=torch.float32)
    s.scale, s.shift, s.mean, s.invstd = coef[:p], coef[p:2 * p], coef[2 * p:3 * p], coef[3 * p:]
    s.sums = None
    if training:
        sums = torch.empty(2 * p, device=dev, dtype=torch.float64)
        L.call('b200gan_bn_stats', C.byref(y.v), L.ptr(sums), _st())
        s.sums = sums if p == c else torch.cat([sums[:c], sums[p:p + c]])
        L.call('b200gan_bn_finalize', L.ptr(s.sums), c, y.v.n * y.v.h * y.v.w, L.ptr(bn.weight), L.ptr(bn.bias), L.ptr(bn.running_mean),
               L.ptr(bn.running_var), L.ptr(bn.num_batches_tracked), BN_MOMENTUM, BN_EPS, L.ptr(s.scale), L.ptr(s.shift), L.ptr(s.mean),
               L.ptr(s.invstd), _st())
    else:
        s.mean = s.invstd = None
        L.call('b200gan_bn_eval_coeffs', c, L.ptr(bn.weight), L.ptr(bn.bias), L.ptr(bn.running_mean), L.ptr(bn.running_var), BN_EPS,
               L.ptr(s.scale), L.ptr(s.shift), _st())
    L.call('b200gan_bn_act_fwd', C.byref(y.v), L.ptr(s.scale), L.ptr(s.shift), act, LRELU_SLOPE, C.byref(out.v), _st())
    return s


def _bn_backward(da: Act, y: Act, s: _BNState, bn, act, dy: Act, dgamma, dbeta):
    """native_batch_norm_backward behind the activation backward: dy (may alias da), dgamma += , dbeta +=.  Padded channels (see _bn_forward):
    gamma = invstd = 0 there, so dy stays exactly zero; their dgamma / dbeta land in scratch."""
    p, c = y.v.c, bn.weight.shape[0]
    dev = y.t.device
    sums = torch.empty(2 * p, device=dev, dtype=torch.float64)
    gamma, dg, db = bn.weight, dgamma, dbeta
    if p != c:
        gamma = _pad_vec(bn.weight.detach(), p)
        scratch = torch.zeros(2 * p, device=dev, dtype=torch.float32)
        dg, db = (scratch[:p], scratch[p:]) if dgamma is not None else (None, None)
    L.call('b200gan_bn_act_bwd_reduce', C.byref(da.v), C.byref(y.v), None, L.ptr(s.scale), L.ptr(s.shift), L.ptr(s.mean), L.ptr(s.invstd), act,
           LRELU_SLOPE, L.ptr(sums), _st())
    L.call('b200gan_bn_act_bwd_apply', C.byref(da.v), C.byref(y.v), None, L.ptr(s.scale), L.ptr(s.shift), L.ptr(s.mean), L.ptr(s.invstd),
           L.ptr(gamma), L.ptr(sums), y.v.n * y.v.h * y.v.w, act, LRELU_SLOPE, C.byref(dy.v), L.ptr(dg), L.ptr(db), _st())
    if p != c and dgamma is not None:
        L.call('b200gan_accumulate_2d', L.ptr(dgamma), L.ptr(dg), 0, 1, c, 0, 1, _st())
        L.call('b200gan_accumulate_2d', L.ptr(dbeta), L.ptr(db), 0, 1, c, 0, 1, _st())


def _bn_again(s: _BNState, y: Act, bn):
    """The running-statistics side effect of one more training-mode forward over the same batch with the same weights (its outputs would be
    identical): the finalize step once more on the saved sums."""
    c = bn.weight.shape[0]
    scratch = torch.empty(4 * c, device=y.t.device, dtype=torch.float32)
    L.call('b200gan_bn_finalize', L.ptr(s.sums), c, y.v.n * y.v.h * y.v.w, L.ptr(bn.weight), L.ptr(bn.bias), L.ptr(bn.running_mean),
           L.ptr(bn.running_var), L.ptr(bn.num_batches_tracked), BN_MOMENTUM, BN_EPS, L.ptr(scratch[:c]), L.ptr(scratch[c:2 * c]),
           L.ptr(scratch[2 * c:3 * c]), L.ptr(scratch[3 * c:]), _st())


_ones_cache = {}


def _ones(c, dev):
    key = (c, dev)
    if key not in _ones_cache:
        _ones_cache[key] = torch.ones(c, device=dev, dtype=torch.float32)
    return _ones_cache[key]


def _bias_act(y: Act, bias, act, out: Act):
    """out = act(y + bias[c])  (out may alias y)."""
    L.call('b200gan_bn_act_fwd', C.byref(y.v), L.ptr(_ones(y.v.c, y.t.device)), L.ptr(bias), act, LRELU_SLOPE, C.byref(out.v), _st())


def _act_backward(da: Act, a: Act, act, dz: Act):
    """dz = da * act'(.) from the saved activation output (no BatchNorm in between)."""
    L.call('b200gan_bn_act_bwd_apply', C.byref(da.v), C.byref(a.v), C.byref(a.v), None, None, None, None, None, None, 0, act, LRELU_SLOPE,
           C.byref(dz.v), None, None, _st())


def _add_channel_sum(dst: torch.Tensor, dy: Act, c: Optional[int] = None):
    """Gradient of a per-channel bias: dst += sum of dy over (N,H,W) (fp64 accumulation); c: only the first c channels (padded tensors)."""
    sums = torch.empty(2 * dy.v.c, device=dy.t.device, dtype=torch.float64)
    L.call('b200gan_bn_stats', C.byref(dy.v), L.ptr(sums), _st())
    L.call('b200gan_accumulate_2d', L.ptr(dst), L.ptr(sums), 1, 1, dy.v.c if c is None else c, 0, 1, _st())


def _buf(grads: dict, into: Optional[dict], p) -> torch.Tensor:
    """The fp32 buffer the gradient of parameter `p` is accumulated into: the caller's (`into`, e.g. a trainer's arena) or a fresh zero one."""
    if p not in grads:
        grads[p] = into[p] if into is not None else torch.zeros_like(p, dtype=torch.float32)
    return grads[p]


def _accumulate(dst: Act, src: Act):
    """dst += src (views of equal extents, any layout / dtype)."""
    L.call('b200gan_sample_axpby', C.byref(dst.v), None, C.byref(src.v), None, C.byref(dst.v), _st())


def _nhwc(n, h, w, c, dev, dtype, zero=False) -> Act:
    return Act((torch.zeros if zero else torch.empty)((n, h, w, c), device=dev, dtype=dtype), nchw=False)


# Channel padding.  The tcgen05 / warp-MMA convolution kernels want channel counts that are multiples of 32; the CLI-default widths have one
# 16-channel tensor per network (feature_maps // 2).  In the tensor-core mode such a tensor is STORED 32 channels wide, the upper half exactly
# zero (zero-padded weight blocks produce it, zero-padded weight blocks consume it), which doubles the multiply-adds of the two layers that touch
# it and moves them from the generic SIMT kernels onto the tensor cores (measured at batch 1024: 161 -> see DESIGN.md ms per iteration).
# Bias and BatchNorm passes run densely over the stored width with zero-padded coefficients (the padding stays exactly zero); what must see the
# REAL channels only (feature matching, the features handed to autograd) goes through a strided view.
def _padded(c: int, enable: bool) -> int:
    return (c + 31) // 32 * 32 if (enable and c >= 16 and c % 32) else c


def _real(act: Act, c: int) -> Act:
    return act if act.v.c == c else Act(act.t[..., :c], nchw=False)


def _pad_block(w: torch.Tensor, d0: int, d1: int) -> torch.Tensor:
    """w (a, b, k, k) -> (d0, d1, k, k), zero outside the original block (a copy only when something is padded)."""
    if w.shape[0] == d0 and w.shape[1] == d1:
        return w
    out = torch.zeros((d0, d1) + tuple(w.shape[2:]), device=w.device, dtype=w.dtype)
    out[:w.shape[0], :w.shape[1]] = w
    return out


def _pad_vec(v: torch.Tensor, d: int) -> torch.Tensor:
    if v.shape[0] == d:
        return v
    out = torch.zeros(d, device=v.device, dtype=v.dtype)
    out[:v.shape[0]] = v
    return out


def _f32(t, what):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise L.B200GanError(f'{what} must be a float32 CUDA tensor (got {t.dtype} on {t.device})')
    return t.contiguous()


def _labels(labels, n, num_classes, dev):
    if labels.dim() != 1 or labels.shape[0] != n:
        raise L.B200GanError(f'labels must have shape ({n},), got {tuple(labels.shape)}')
    return labels.to(device=dev, dtype=torch.int64).contiguous()


# ------------------------------------------------------------------------------------------------ Generator
class _GLayer:
    __slots__ = ('x', 'y', 'a', 'bn', 'w4', 'packs')


class GeneratorEngine:
    """cgan.Generator on the library.  Layer 0 is the conditioned latent GEMM (+BN+ReLU); layers 1..5 the folded upsample-convolutions."""

    def __init__(self, latent_dim, num_classes, nc, nf, dtype=None, algo=None):
        self.dtype = dtype or default_compute_dtype()
        self.algo = default_algo() if algo is None else algo
        self.nz, self.classes, self.nc, self.nf = latent_dim, num_classes, nc, nf
        self.ch = [nf * 8, nf * 4, nf * 2, nf, nf // 2, nc]
        pad = self.dtype == torch.bfloat16 and self.algo != L.ALGO_SIMT
        self.pch = [_padded(c, pad) for c in self.ch[:5]] + [nc]          # stored widths (see _padded)
        self.fc = _ConvOp(INIT_SIZE, 1, 0, True, latent_dim + 1, self.ch[0], self.dtype, self.algo)
        self.up = [_ConvOp(4, 2, 1, True, self.pch[i], self.pch[i + 1], self.dtype, self.algo) for i in range(5)]

    @staticmethod
    def conv_of(mod, i):          # i = 1..5: main[3], [7], [11], [15], [19]
        return mod.main[4 * i - 1]

    @staticmethod
    def bn_of(mod, i):            # i = 0..4: main[0], [4], [8], [12], [16]
        return mod.main[4 * i]

    def fc_weight(self, mod):
        """[fc.weight^T ; fc.bias] viewed (nz+1, 8nf, 7, 7): the ConvTranspose2d weight of the latent GEMM, bias as the last input feature."""
        w = torch.cat([mod.fc.weight.detach().t(), mod.fc.bias.detach()[None]], 0).contiguous()
        return w.view(self.nz + 1, self.ch[0], INIT_SIZE, INIT_SIZE)

    def forward(self, mod, z, labels, save: bool, internal: bool = False):
        """Returns (image, tape): the image as the reference's (N, nc, 224, 224) fp32 tensor, or with internal=True as an Act in the engine's own
        layout and compute dtype (what the fused trainer hands straight to the Discriminator)."""
        training = mod.training
        if save and not training:
            raise L.B200GanError('backward through an eval-mode network is not on the training path and is not implemented')
        dev = z.device
        n = z.shape[0]
        if z.dim() != 2 or z.shape[1] != self.nz:
            raise L.B200GanError(f'expected noise of shape (N,{self.nz}), got {tuple(z.shape)}')
        z = _f32(z.detach(), 'noise')
        labels = _labels(labels, n, self.classes, dev)
        table = _f32(mod.label_emb.weight.detach(), 'label_emb.weight')
        xa = torch.empty((n, 1, 1, self.nz + 1), device=dev, dtype=torch.float32)
        L.call('b200gan_embed_add', L.ptr(table), L.ptr(labels), L.ptr(z), n, self.nz, 1, L.ptr(xa), _st())
        tape: List[_GLayer] = []
        # layer 0: fc (+bias) as the k7 transposed convolution of the 1x1 input, then main[0] BatchNorm + main[1] ReLU
        l0 = _GLayer()
        l0.x, l0.w4, l0.packs = Act(xa, nchw=False), self.fc_weight(mod), (None, None)
        if self.pch[0] != self.ch[0]:
            raise L.B200GanError(f'feature_maps_g = {self.nf}: the first generator width must be a multiple of 32 or below 16 in the bf16 mode')
        l0.y = _nhwc(n, INIT_SIZE, INIT_SIZE, self.ch[0], dev, self.dtype)
        self.fc.fprop(l0.x, l0.w4, l0.y, l0.packs)
        l0.a = _nhwc(n, INIT_SIZE, INIT_SIZE, self.ch[0], dev, self.dtype)
        l0.bn = _bn_forward(l0.y, self.bn_of(mod, 0), L.ACT_RELU, l0.a, training)
        tape.append(l0)
        cur = l0.a
        out_t = None
        for i in range(1, 6):
            conv, op = self.conv_of(mod, i), self.up[i - 1]
            lay = _GLayer()
            lay.x = cur
            cin, cout = self.ch[i - 1], self.ch[i]                    # real widths; op.cin / op.cout are the stored ones
            folded = torch.empty((cin, cout, 4, 4), device=dev, dtype=torch.float32)
            w3 = _f32(conv.weight.detach(), 'conv weight')
            L.call('b200gan_upconv3_fold', L.ptr(w3), cout, cin, L.ptr(folded), _st())
            lay.w4 = _pad_block(folded, op.cin, op.cout)
            lay.packs = op.pack(lay.w4)
            h = cur.v.h * 2
            lay.y = _nhwc(n, h, h, op.cout, dev, self.dtype)
            op.fprop(cur, lay.w4, lay.y, lay.packs)
            if i < 5:
                _bias_act(lay.y, _pad_vec(conv.bias.detach(), op.cout), L.ACT_NONE, lay.y)
                lay.a = _nhwc(n, h, h, op.cout, dev, self.dtype)
                lay.bn = _bn_forward(lay.y, self.bn_of(mod, i), L.ACT_RELU, lay.a, training)
            elif internal:
                lay.a, lay.bn = lay.y, None
                _bias_act(lay.y, conv.bias.detach(), L.ACT_TANH, lay.a)                          # in place
                lay.y = None
            else:
                out_t = torch.empty((n, self.nc, h, h), device=dev, dtype=torch.float32)        # the reference's NCHW fp32 image
                lay.a, lay.bn = Act(out_t, nchw=True), None
                _bias_act(lay.y, conv.bias.detach(), L.ACT_TANH, lay.a)
                lay.y = None
            tape.append(lay)
            cur = lay.a
        return (cur if internal else out_t), (tape, labels)

    def backward(self, mod, saved, dout, need_dz: bool, into: Optional[dict] = None):
        """Returns (dz or None, {parameter: gradient}) for one upstream gradient `dout`: an (N, nc, 224, 224) fp32 tensor or an Act.  With `into`
        ({parameter: fp32 buffer}) the gradients are accumulated there instead of into fresh tensors."""
        tape, labels = saved
        grads = {}
        d = dout if isinstance(dout, Act) else Act(_f32(dout, 'grad_output'), nchw=True)
        dev = d.t.device
        n = d.v.n
        for i in range(5, 0, -1):
            lay, conv, op = tape[i], self.conv_of(mod, i), self.up[i - 1]
            cin, cout = self.ch[i - 1], self.ch[i]
            dy = _nhwc(n, lay.a.v.h, lay.a.v.w, op.cout, dev, self.dtype)
            if lay.bn is None:
                _act_backward(d, lay.a, L.ACT_TANH, dy)
            else:
                bn = self.bn_of(mod, i)
                _bn_backward(d, lay.y, lay.bn, bn, L.ACT_RELU, dy, _buf(grads, into, bn.weight), _buf(grads, into, bn.bias))
            _add_channel_sum(_buf(grads, into, conv.bias), dy, cout)
            dw4 = torch.zeros_like(lay.w4)
            op.wgrad(lay.x, dy, dw4)
            if dw4.shape[0] != cin or dw4.shape[1] != cout:
                dw4 = dw4[:cin, :cout].contiguous()                 # the real block of the padded weight gradient
            L.call('b200gan_upconv3_unfold', L.ptr(dw4), cout, cin, L.ptr(_buf(grads, into, conv.weight)), _st())
            d = _nhwc(n, lay.x.v.h, lay.x.v.w, op.cin, dev, self.dtype)
            op.dgrad(dy, lay.w4, d, lay.packs)
        l0, bn = tape[0], self.bn_of(mod, 0)
        _bn_backward(d, l0.y, l0.bn, bn, L.ACT_RELU, d, _buf(grads, into, bn.weight), _buf(grads, into, bn.bias))
        dwa = torch.zeros_like(l0.w4)
        self.fc.wgrad(l0.x, d, dwa)
        m = dwa.numel() // (self.nz + 1)                      # back in the Linear's layout: weight (out, in) from the GEMM's (in, out), bias = last row
        L.call('b200gan_accumulate_2d', L.ptr(_buf(grads, into, mod.fc.weight)), L.ptr(dwa), 0, m, self.nz, 1, m, _st())
        L.call('b200gan_accumulate_2d', L.ptr(_buf(grads, into, mod.fc.bias)), L.ptr(dwa.view(-1)[self.nz * m:]), 0, 1, m, 0, 1, _st())
        dxa = torch.empty((n, 1, 1, self.nz + 1), device=dev, dtype=torch.float32)
        self.fc.dgrad(d, l0.w4, Act(dxa, nchw=False), l0.packs)
        L.call('b200gan_embed_bwd', L.ptr(dxa), L.ptr(labels), n, self.nz, self.nz + 1, self.classes, L.ptr(_buf(grads, into, mod.label_emb.weight)), _st())
        dz = dxa.view(n, self.nz + 1)[:, :self.nz].contiguous() if need_dz else None
        return dz, grads


# ------------------------------------------------------------------------------------------------ Discriminator
class _DLayer:
    __slots__ = ('x', 'y', 'a', 'bn', 'packs', 'w')


class DiscriminatorEngine:
    """cgan.Discriminator on the library: five k4 s2 p1 convolutions with bias (BatchNorm on 1..4), the 7x7 head, the projection term."""

    CONV_IDX = [0, 2, 5, 8, 11, 14]

    def __init__(self, num_classes, nc, nf, dtype=None, algo=None):
        self.dtype = dtype or default_compute_dtype()
        self.algo = default_algo() if algo is None else algo
        self.classes, self.nc, self.nf = num_classes, nc, nf
        self.ch = [nc, nf // 2, nf, nf * 2, nf * 4, nf * 8]
        self.pch = list(self.ch)                                          # stored widths: the first layer's output may be padded (see _padded)
        self.pch[1] = _padded(self.ch[1], self.dtype == torch.bfloat16 and self.algo != L.ALGO_SIMT)
        self.down = [_ConvOp(4, 2, 1, False, self.pch[i], self.pch[i + 1], self.dtype, self.algo) for i in range(5)]
        self.head = _ConvOp(INIT_SIZE, 1, 0, False, self.ch[5], 1, self.dtype, self.algo)

    def conv_of(self, mod, i):
        return mod.main[self.CONV_IDX[i]]

    def bn_of(self, mod, i):      # i = 1..4
        return mod.main[self.CONV_IDX[i] + 1]

    def forward(self, mod, x, labels, save: bool, head: bool = True):
        """Returns (logits (N,) or None, tape).  head=False stops after the last LeakyReLU (get_intermediate_features)."""
        training = mod.training
        if save and not training:
            raise L.B200GanError('backward through an eval-mode network is not on the training path and is not implemented')
        if isinstance(x, Act):                      # already in the engine's layout (the fused trainer's fake batch)
            cur = x
        else:
            if x.dim() != 4:
                raise L.B200GanError(f'expected images of shape (N,{self.nc},{32 * INIT_SIZE},{32 * INIT_SIZE}), got {tuple(x.shape)}')
            xin = x.detach()
            if xin.dtype not in (torch.float32, torch.bfloat16):
                xin = xin.float()
            cur = Act(xin, nchw=True)
        dev, n = cur.t.device, cur.v.n
        if cur.v.c != self.nc or cur.v.h != 32 * INIT_SIZE or cur.v.w != 32 * INIT_SIZE:
            raise L.B200GanError(f'expected images of shape (N,{self.nc},{32 * INIT_SIZE},{32 * INIT_SIZE}), got (N,{cur.v.c},{cur.v.h},{cur.v.w})')
        labels = _labels(labels, n, self.classes, dev)
        tape: List[_DLayer] = []
        for i in range(5):
            conv, op = self.conv_of(mod, i), self.down[i]
            lay = _DLayer()
            lay.w = _pad_block(_f32(conv.weight.detach(), 'conv weight'), op.cout, op.cin)
            lay.x, lay.packs = cur, op.pack(lay.w)
            h = cur.v.h // 2
            lay.y = _nhwc(n, h, h, op.cout, dev, self.dtype)
            op.fprop(cur, lay.w, lay.y, lay.packs)
            if i == 0:
                lay.a, lay.bn = lay.y, None
                _bias_act(lay.y, _pad_vec(conv.bias.detach(), op.cout), L.ACT_LRELU, lay.a)      # in place, as the reference's inplace LeakyReLU
            else:
                _bias_act(lay.y, conv.bias.detach(), L.ACT_NONE, lay.y)
                lay.a = _nhwc(n, h, h, op.cout, dev, self.dtype)
                lay.bn = _bn_forward(lay.y, self.bn_of(mod, i), L.ACT_LRELU, lay.a, training)
            tape.append(lay)
            cur = lay.a
        logits = None
        if head:
            conv = self.conv_of(mod, 5)
            logits = torch.empty((n, 1, 1, 1), device=dev, dtype=torch.float32)
            lg = Act(logits, nchw=False)
            self.head.fprop(cur, conv.weight.detach(), lg, (None, None))
            _bias_act(lg, conv.bias.detach(), L.ACT_NONE, lg)
            table = _f32(mod.label_emb.weight.detach(), 'label_emb.weight')
            L.call('b200gan_class_proj_fwd', C.byref(cur.v), L.ptr(table), L.ptr(labels), L.ptr(logits), _st())
            logits = logits.view(n)
        return logits, (tape, labels)

    def feature_tensors(self, tape) -> List[torch.Tensor]:
        """The distinct intermediates of main[:-1] as NCHW fp32 tensors, in order [a0, y1, a1, y2, a2, y3, a3, y4, a4]."""
        outs = []
        for _, src in self.feature_acts(tape):
            t = torch.empty((src.v.n, src.v.c, src.v.h, src.v.w), device=src.t.device, dtype=torch.float32)
            L.call('b200gan_copy_view', C.byref(src.v), C.byref(Act(t, nchw=True).v), _st())
            outs.append(t)
        return outs

    def feature_acts(self, tape):
        """[(stored Act, the same restricted to its real channels)] for [a0, y1, a1, y2, a2, y3, a3, y4, a4]."""
        out = []
        for i, lay in enumerate(tape):
            for src in ([lay.a] if i == 0 else [lay.y, lay.a]):
                out.append((src, _real(src, self.ch[i + 1])))
        return out

    def replay_running_stats(self, mod, saved):
        """What one more training-mode forward over the same batch would do to the BatchNorm buffers (train_cgan.py:181 and :188 run the
        Discriminator twice on the same fake batch with the same weights; the second pass's outputs are the first's)."""
        for i, lay in enumerate(saved[0]):
            if lay.bn is not None:
                _bn_again(lay.bn, lay.y, self.bn_of(mod, i))

    def backward(self, mod, saved, dlogits: Optional[torch.Tensor], dfeats: Optional[list], need_dx: bool, need_dw: bool, into: Optional[dict] = None,
                 dx_out: Optional[Act] = None):
        """Gradients for the upstream gradient of the logits (N,) and / or of the nine feature tensors [a0, y1, a1, ..., y4, a4] (NCHW fp32
        tensors, Acts in any layout, or callables `g(target_act)` that add their gradient into the target themselves; entries may be None).  Returns (dx or None, {param: grad}); `into`: see GeneratorEngine.backward;
        `dx_out`: where the input gradient goes (default: a fresh NCHW fp32 tensor)."""
        tape, labels = saved
        dev = tape[0].x.t.device
        n = tape[0].x.v.n
        grads = {}
        top = tape[4]
        d = _nhwc(n, INIT_SIZE, INIT_SIZE, self.ch[5], dev, self.dtype)        # gradient w.r.t. a4
        have_d = False
        if dlogits is not None:
            conv = self.conv_of(mod, 5)
            dl = _f32(dlogits, 'grad of logits').view(n, 1, 1, 1)
            dla = Act(dl, nchw=False)
            self.head.dgrad(dla, conv.weight.detach(), d, (None, None))
            have_d = True
            table = _f32(mod.label_emb.weight.detach(), 'label_emb.weight')
            dtable = None
            if need_dw:
                self.head.wgrad(top.a, dla, _buf(grads, into, conv.weight))
                _add_channel_sum(_buf(grads, into, conv.bias), dla)
                dtable = _buf(grads, into, mod.label_emb.weight)
            L.call('b200gan_class_proj_bwd', C.byref(top.a.v), L.ptr(table), L.ptr(labels), L.ptr(dl), C.byref(d.v), self.classes, L.ptr(dtable), _st())

        def add_feat(j, i, d):
            """d += gradient of feature j (an intermediate of layer i): a tensor / Act to accumulate, or a callable that adds it itself
            (the fused trainer's feature-matching kernel writes straight into d: no gradient tensor, no accumulate pass)."""
            g = None if dfeats is None else dfeats[j]
            if g is None:
                return
            tgt = _real(d, self.ch[i + 1])
            if callable(g):
                g(tgt)
            else:
                _accumulate(tgt, g if isinstance(g, Act) else Act(_f32(g, 'grad of features'), nchw=True))

        for i in range(4, -1, -1):
            lay, conv, op = tape[i], self.conv_of(mod, i), self.down[i]
            if not have_d:
                d.t.zero_()                   # no gradient reaches this depth from above (memset)
                have_d = True
            add_feat(2 * i, i, d)             # a_i sits at position 2i of [a0, y1, a1, ..., y4, a4]
            if lay.bn is None:
                _act_backward(d, lay.a, L.ACT_LRELU, d)
            else:
                bn = self.bn_of(mod, i)
                dg = _buf(grads, into, bn.weight) if need_dw else None
                db = _buf(grads, into, bn.bias) if need_dw else None
                _bn_backward(d, lay.y, lay.bn, bn, L.ACT_LRELU, d, dg, db)
                add_feat(2 * i - 1, i, d)
            if need_dw:
                cin, cout = self.ch[i], self.ch[i + 1]
                _add_channel_sum(_buf(grads, into, conv.bias), d, cout)
                if op.cin == cin and op.cout == cout:
                    op.wgrad(lay.x, d, _buf(grads, into, conv.weight))
                else:                         # padded layer: the real block of the padded weight gradient
                    dwp = torch.zeros_like(lay.w)
                    op.wgrad(lay.x, d, dwp)
                    L.call('b200gan_accumulate_2d', L.ptr(_buf(grads, into, conv.weight)), L.ptr(dwp), 0, cout, cin * 16, op.cin * 16, 1, _st())
            if i > 0:
                nd = _nhwc(n, lay.x.v.h, lay.x.v.w, op.cin, dev, self.dtype)
                op.dgrad(d, lay.w, nd, lay.packs)
                d = nd
            elif need_dx:
                if dx_out is None:
                    dx_out = Act(torch.empty((n, self.nc, lay.x.v.h, lay.x.v.w), device=dev, dtype=torch.float32), nchw=True)
                op.dgrad(d, lay.w, dx_out, lay.packs)
                return dx_out.t, grads
        return None, grads
