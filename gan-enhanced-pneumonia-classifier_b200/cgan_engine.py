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
        name = 'b200gan_convT2d_wgrad' if self.transposed else 'b200gan_conv2d_wgrad'
        need = int(L.load().b200gan_conv_wgrad_workspace_floats(C.byref(self.desc), C.byref(x.v), C.byref(dy.v), 1 if self.transposed else 0))
        ws = None
        if need > 0:
            if self._ws is None or self._ws.numel() < need or self._ws.device != dw.device:
                self._ws = torch.zeros(need, device=dw.device, dtype=torch.float32)       # zero on entry, handed back zero
            ws = self._ws
        L.call(name, C.byref(self.desc), C.byref(x.v), C.byref(dy.v), L.ptr(dw), L.ptr(ws), None, _st())


class _BNState:
    __slots__ = ('scale', 'shift', 'mean', 'invstd')


def _bn_forward(y: Act, bn, act, out: Act, training: bool) -> _BNState:
    """BatchNorm2d + activation (training: batch statistics, running-stat side effects; eval: running statistics)."""
    dev, c = y.t.device, y.v.c
    s = _BNState()
    s.scale = torch.empty(c, device=dev, dtype=torch.float32)
    s.shift = torch.empty(c, device=dev, dtype=torch.float32)
    s.mean = s.invstd = None
    if training:
        sums = torch.empty(2 * c, device=dev, dtype=torch.float64)
        s.mean = torch.empty(c, device=dev, dtype=torch.float32)
        s.invstd = torch.empty(c, device=dev, dtype=torch.float32)
        L.call('b200gan_bn_stats', C.byref(y.v), L.ptr(sums), _st())
        L.call('b200gan_bn_finalize', L.ptr(sums), c, y.v.n * y.v.h * y.v.w, L.ptr(bn.weight), L.ptr(bn.bias), L.ptr(bn.running_mean),
               L.ptr(bn.running_var), L.ptr(bn.num_batches_tracked), BN_MOMENTUM, BN_EPS, L.ptr(s.scale), L.ptr(s.shift), L.ptr(s.mean),
               L.ptr(s.invstd), _st())
    else:
        L.call('b200gan_bn_eval_coeffs', c, L.ptr(bn.weight), L.ptr(bn.bias), L.ptr(bn.running_mean), L.ptr(bn.running_var), BN_EPS,
               L.ptr(s.scale), L.ptr(s.shift), _st())
    L.call('b200gan_bn_act_fwd', C.byref(y.v), L.ptr(s.scale), L.ptr(s.shift), act, LRELU_SLOPE, C.byref(out.v), _st())
    return s


def _bn_backward(da: Act, y: Act, s: _BNState, bn, act, dy: Act, dgamma, dbeta):
    """native_batch_norm_backward behind the activation backward: dy (may alias da), dgamma += , dbeta +=."""
    c = y.v.c
    sums = torch.empty(2 * c, device=y.t.device, dtype=torch.float64)
    L.call('b200gan_bn_act_bwd_reduce', C.byref(da.v), C.byref(y.v), None, L.ptr(s.scale), L.ptr(s.shift), L.ptr(s.mean), L.ptr(s.invstd), act,
           LRELU_SLOPE, L.ptr(sums), _st())
    L.call('b200gan_bn_act_bwd_apply', C.byref(da.v), C.byref(y.v), None, L.ptr(s.scale), L.ptr(s.shift), L.ptr(s.mean), L.ptr(s.invstd),
           L.ptr(bn.weight), L.ptr(sums), y.v.n * y.v.h * y.v.w, act, LRELU_SLOPE, C.byref(dy.v), L.ptr(dgamma), L.ptr(dbeta), _st())


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


def _channel_sum(dy: Act) -> torch.Tensor:
    """Gradient of a per-channel bias: sum of dy over (N,H,W), fp64 accumulation, returned as fp32."""
    sums = torch.empty(2 * dy.v.c, device=dy.t.device, dtype=torch.float64)
    L.call('b200gan_bn_stats', C.byref(dy.v), L.ptr(sums), _st())
    return sums[:dy.v.c].float()


def _accumulate(dst: Act, src: Act):
    """dst += src (views of equal extents, any layout / dtype)."""
    L.call('b200gan_sample_axpby', C.byref(dst.v), None, C.byref(src.v), None, C.byref(dst.v), _st())


def _nhwc(n, h, w, c, dev, dtype) -> Act:
    return Act(torch.empty((n, h, w, c), device=dev, dtype=dtype), nchw=False)


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
        self.fc = _ConvOp(INIT_SIZE, 1, 0, True, latent_dim + 1, self.ch[0], self.dtype, self.algo)
        self.up = [_ConvOp(4, 2, 1, True, self.ch[i], self.ch[i + 1], self.dtype, self.algo) for i in range(5)]

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

    def forward(self, mod, z, labels, save: bool):
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
            lay.w4 = torch.empty((op.cin, op.cout, 4, 4), device=dev, dtype=torch.float32)
            w3 = _f32(conv.weight.detach(), 'conv weight')
            L.call('b200gan_upconv3_fold', L.ptr(w3), op.cout, op.cin, L.ptr(lay.w4), _st())
            lay.packs = op.pack(lay.w4)
            h = cur.v.h * 2
            lay.y = _nhwc(n, h, h, op.cout, dev, self.dtype)
            op.fprop(cur, lay.w4, lay.y, lay.packs)
            if i < 5:
                _bias_act(lay.y, conv.bias.detach(), L.ACT_NONE, lay.y)
                lay.a = _nhwc(n, h, h, op.cout, dev, self.dtype)
                lay.bn = _bn_forward(lay.y, self.bn_of(mod, i), L.ACT_RELU, lay.a, training)
            else:
                out_t = torch.empty((n, self.nc, h, h), device=dev, dtype=torch.float32)        # the reference's NCHW fp32 image
                lay.a, lay.bn = Act(out_t, nchw=True), None
                _bias_act(lay.y, conv.bias.detach(), L.ACT_TANH, lay.a)
                lay.y = None
            tape.append(lay)
            cur = lay.a
        return out_t, (tape, labels)

    def backward(self, mod, saved, dout, need_dz: bool):
        """Returns (dz or None, {parameter: gradient}) for one upstream gradient `dout` (N, nc, 224, 224)."""
        tape, labels = saved
        dev = dout.device
        n = dout.shape[0]
        grads = {}
        d = Act(_f32(dout, 'grad_output'), nchw=True)
        for i in range(5, 0, -1):
            lay, conv, op = tape[i], self.conv_of(mod, i), self.up[i - 1]
            dy = _nhwc(n, lay.a.v.h, lay.a.v.w, op.cout, dev, self.dtype)
            if lay.bn is None:
                _act_backward(d, lay.a, L.ACT_TANH, dy)
            else:
                bn = self.bn_of(mod, i)
                grads[bn.weight] = torch.zeros_like(bn.weight)
                grads[bn.bias] = torch.zeros_like(bn.bias)
                _bn_backward(d, lay.y, lay.bn, bn, L.ACT_RELU, dy, grads[bn.weight], grads[bn.bias])
            grads[conv.bias] = _channel_sum(dy)
            dw4 = torch.zeros_like(lay.w4)
            op.wgrad(lay.x, dy, dw4)
            grads[conv.weight] = torch.zeros_like(conv.weight)
            L.call('b200gan_upconv3_unfold', L.ptr(dw4), op.cout, op.cin, L.ptr(grads[conv.weight]), _st())
            d = _nhwc(n, lay.x.v.h, lay.x.v.w, op.cin, dev, self.dtype)
            op.dgrad(dy, lay.w4, d, lay.packs)
        l0, bn = tape[0], self.bn_of(mod, 0)
        grads[bn.weight] = torch.zeros_like(bn.weight)
        grads[bn.bias] = torch.zeros_like(bn.bias)
        _bn_backward(d, l0.y, l0.bn, bn, L.ACT_RELU, d, grads[bn.weight], grads[bn.bias])
        dwa = torch.zeros_like(l0.w4)
        self.fc.wgrad(l0.x, d, dwa)
        flat = dwa.view(self.nz + 1, -1)
        grads[mod.fc.weight] = flat[:self.nz].t()             # back in the Linear's (out, in) layout
        grads[mod.fc.bias] = flat[self.nz]
        dxa = torch.empty((n, 1, 1, self.nz + 1), device=dev, dtype=torch.float32)
        self.fc.dgrad(d, l0.w4, Act(dxa, nchw=False), l0.packs)
        grads[mod.label_emb.weight] = torch.zeros_like(mod.label_emb.weight)
        L.call('b200gan_embed_bwd', L.ptr(dxa), L.ptr(labels), n, self.nz, self.nz + 1, self.classes, L.ptr(grads[mod.label_emb.weight]), _st())
        dz = dxa.view(n, self.nz + 1)[:, :self.nz].contiguous() if need_dz else None
        return dz, grads


# ------------------------------------------------------------------------------------------------ Discriminator
class _DLayer:
    __slots__ = ('x', 'y', 'a', 'bn', 'packs')


class DiscriminatorEngine:
    """cgan.Discriminator on the library: five k4 s2 p1 convolutions with bias (BatchNorm on 1..4), the 7x7 head, the projection term."""

    CONV_IDX = [0, 2, 5, 8, 11, 14]

    def __init__(self, num_classes, nc, nf, dtype=None, algo=None):
        self.dtype = dtype or default_compute_dtype()
        self.algo = default_algo() if algo is None else algo
        self.classes, self.nc, self.nf = num_classes, nc, nf
        self.ch = [nc, nf // 2, nf, nf * 2, nf * 4, nf * 8]
        self.down = [_ConvOp(4, 2, 1, False, self.ch[i], self.ch[i + 1], self.dtype, self.algo) for i in range(5)]
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
        dev = x.device
        n = x.shape[0]
        if x.dim() != 4 or x.shape[1] != self.nc or x.shape[2] != 32 * INIT_SIZE or x.shape[3] != 32 * INIT_SIZE:
            raise L.B200GanError(f'expected images of shape (N,{self.nc},{32 * INIT_SIZE},{32 * INIT_SIZE}), got {tuple(x.shape)}')
        xin = x.detach()
        if xin.dtype != torch.float32:
            xin = xin.float()
        labels = _labels(labels, n, self.classes, dev)
        tape: List[_DLayer] = []
        cur = Act(xin, nchw=True)
        for i in range(5):
            conv, op = self.conv_of(mod, i), self.down[i]
            lay = _DLayer()
            lay.x, lay.packs = cur, op.pack(_f32(conv.weight.detach(), 'conv weight'))
            h = cur.v.h // 2
            lay.y = _nhwc(n, h, h, op.cout, dev, self.dtype)
            op.fprop(cur, conv.weight.detach(), lay.y, lay.packs)
            if i == 0:
                lay.a, lay.bn = lay.y, None
                _bias_act(lay.y, conv.bias.detach(), L.ACT_LRELU, lay.a)           # in place, as the reference's inplace LeakyReLU
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
        for i, lay in enumerate(tape):
            for src in ([lay.a] if i == 0 else [lay.y, lay.a]):
                t = torch.empty((src.v.n, src.v.c, src.v.h, src.v.w), device=src.t.device, dtype=torch.float32)
                L.call('b200gan_copy_view', C.byref(src.v), C.byref(Act(t, nchw=True).v), _st())
                outs.append(t)
        return outs

    def backward(self, mod, saved, dlogits: Optional[torch.Tensor], dfeats: Optional[List[Optional[torch.Tensor]]], need_dx: bool, need_dw: bool):
        """Gradients for the upstream gradient of the logits (N,) and / or of the nine feature tensors.  Returns (dx or None, {param: grad})."""
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
                grads[conv.weight] = torch.zeros_like(conv.weight)
                self.head.wgrad(top.a, dla, grads[conv.weight])
                grads[conv.bias] = _channel_sum(dla)
                dtable = grads[mod.label_emb.weight] = torch.zeros_like(mod.label_emb.weight)
            L.call('b200gan_class_proj_bwd', C.byref(top.a.v), L.ptr(table), L.ptr(labels), L.ptr(dl), C.byref(d.v), self.classes, L.ptr(dtable), _st())

        def feat(j):
            return None if dfeats is None or dfeats[j] is None else Act(_f32(dfeats[j], 'grad of features'), nchw=True)

        for i in range(4, -1, -1):
            lay, conv, op = tape[i], self.conv_of(mod, i), self.down[i]
            ga = feat(2 * i)                  # a_i sits at position 2i of [a0, y1, a1, ..., y4, a4]
            if ga is not None:
                if have_d:
                    _accumulate(d, ga)
                else:
                    L.call('b200gan_copy_view', C.byref(ga.v), C.byref(d.v), _st())
                    have_d = True
            if not have_d:
                d.t.zero_()                   # no gradient reaches this depth from above (memset)
                have_d = True
            if lay.bn is None:
                _act_backward(d, lay.a, L.ACT_LRELU, d)
            else:
                bn = self.bn_of(mod, i)
                dg = db = None
                if need_dw:
                    dg = grads[bn.weight] = torch.zeros_like(bn.weight)
                    db = grads[bn.bias] = torch.zeros_like(bn.bias)
                _bn_backward(d, lay.y, lay.bn, bn, L.ACT_LRELU, d, dg, db)
                gy = feat(2 * i - 1)
                if gy is not None:
                    _accumulate(d, gy)
            if need_dw:
                grads[conv.bias] = _channel_sum(d)
                grads[conv.weight] = torch.zeros_like(conv.weight)
                op.wgrad(lay.x, d, grads[conv.weight])
            if i > 0:
                nd = _nhwc(n, lay.x.v.h, lay.x.v.w, op.cin, dev, self.dtype)
                op.dgrad(d, conv.weight.detach(), nd, lay.packs)
                d = nd
            elif need_dx:
                dx = torch.empty((n, self.nc, lay.x.v.h, lay.x.v.w), device=dev, dtype=torch.float32)
                op.dgrad(d, conv.weight.detach(), Act(dx, nchw=True), lay.packs)
                return dx, grads
        return None, grads
