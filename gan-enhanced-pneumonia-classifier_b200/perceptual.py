"""The VGG16 perceptual loss of the conditional GAN on the kernels of libb200gan.so (reference src/train_cgan.py:57-73 `PerceptualLoss`, used at
:186 as `10.0 * perceptual_loss(fake_images, real_images)`).

`PerceptualLoss` is the drop-in: the same constructor-less call contract (`loss = perceptual(x, y)`), the same `self.blocks` ModuleList of torchvision's
`vgg16.features[:4], [4:9], [9:16]` in eval mode with frozen parameters (so a checkpoint loads the same way), stock torch on CPU tensors.  On CUDA
tensors the arithmetic runs through `PerceptualEngine`:

  * every Conv2d(3,1,1) + ReLU is ONE stride-2 4x4 convolution producing all four output parities as 4*Co channels (b200gan_conv3x3_fold builds the
    weight once -- the network is frozen), followed by one pass that adds the bias, applies ReLU and scatters depth-to-space (b200gan_bias_relu_d2s);
    the 64..256-channel layers therefore run on the tcgen05 kernels of the Discriminator, input gradient included (b200gan_conv2d_dgrad with the same
    folded weight behind b200gan_relu_bwd_s2d); the 3-channel image is stored 32 channels wide in the tensor-core mode;
  * MaxPool2d(2,2): b200gan_maxpool2_fwd / _bwd (ATen's first-maximum tie rule);
  * the three MSE terms and their gradients: b200gan_fm_pair, written / added straight into the backward pass's tensors;
  * only the gradient w.r.t. the first argument exists (the reference's `real_images` carry no gradient and VGG is frozen).

Where the weights come from is the caller's business: the reference downloads torchvision's ImageNet checkpoint at start-up (train_cgan.py:60), which
an offline machine cannot; `PerceptualLoss(weights=...)` takes 'imagenet' (torchvision's loader: download or cache), a path to a vgg16 state_dict, or
'random' (architecture only: benchmarks and parity tests).  Parity is pinned for the operator on arbitrary weights (oracle/vgg_oracle.py against
torchvision's own vgg16), not for the checkpoint's values.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib as L
from .engine import Act, default_algo, default_compute_dtype

CONV_POS = [(0, 0), (0, 2), (1, 1), (1, 3), (2, 1), (2, 3), (2, 5)]       # (block, index inside the block) of the seven Conv2d modules
POOL_BEFORE = {2, 4}
BLOCK_END = [1, 3, 6]


def _st():
    return L.stream_ptr()


def _nhwc(n, h, w, c, dev, dtype, zero=False) -> Act:
    return Act((torch.zeros if zero else torch.empty)((n, h, w, c), device=dev, dtype=dtype), nchw=False)


class _Layer:
    __slots__ = ('x', 'a', 'pool_src')


class PerceptualEngine:
    def __init__(self, blocks: nn.ModuleList, dtype: Optional[torch.dtype] = None, algo: Optional[int] = None):
        self.dtype = dtype or default_compute_dtype()
        self.algo = default_algo() if algo is None else algo
        self.tc = self.dtype == torch.bfloat16 and self.algo != L.ALGO_SIMT
        self.desc = L.Conv(4, 2, 1, self.algo)
        self.convs = [blocks[b][i] for b, i in CONV_POS]
        self.cin_pad = 32 if self.tc else self.convs[0].in_channels
        self.w4: List[torch.Tensor] = []
        self.packs = []
        self._scale = None
        for li, conv in enumerate(self.convs):       # frozen network: fold (and repack for the tensor cores) once
            co, ci = conv.out_channels, conv.in_channels
            cip = self.cin_pad if li == 0 else ci
            w3 = conv.weight.detach().float().contiguous()
            w4 = torch.empty((4 * co, cip, 4, 4), device=w3.device, dtype=torch.float32)
            L.call('b200gan_conv3x3_fold', L.ptr(w3), co, ci, cip, L.ptr(w4), _st())
            self.w4.append(w4)
            if self.tc and cip % 32 == 0:
                both = torch.empty(2 * w4.numel(), device=w4.device, dtype=torch.bfloat16)
                L.call('b200gan_pack_conv_weight', L.ptr(w4), w4.shape[0], w4.shape[1], 4, 2, L.ptr(both), _st())
                self.packs.append((both[:w4.numel()], both[w4.numel():]))
            else:
                self.packs.append((None, None))

    # -- plumbing ---------------------------------------------------------------------------------------------------------------------
    def _image(self, x) -> Act:
        """The image as an NHWC tensor of the compute dtype, `cin_pad` channels wide (zero beyond the real ones)."""
        src = x if isinstance(x, Act) else Act(x.detach() if x.dtype in (torch.float32, torch.bfloat16) else x.detach().float(), nchw=True)
        c = self.convs[0].in_channels
        if src.v.c != c:
            raise L.B200GanError(f'the perceptual loss expects {c}-channel images, got {src.v.c}')
        if src.v.h % 8 or src.v.w % 8:
            raise L.B200GanError(f'image extents must be multiples of 8 (two 2x2 poolings under stride-2 kernels), got {src.v.h}x{src.v.w}')
        out = _nhwc(src.v.n, src.v.h, src.v.w, self.cin_pad, src.t.device, self.dtype, zero=self.cin_pad != c)
        L.call('b200gan_copy_view', C.byref(src.v), C.byref(Act(out.t[..., :c], nchw=False).v), _st())
        return out

    def _conv_relu(self, li: int, x: Act) -> Act:
        co = self.convs[li].out_channels
        t = _nhwc(x.v.n, x.v.h // 2, x.v.w // 2, 4 * co, x.t.device, self.dtype)
        L.call('b200gan_conv2d_fprop', C.byref(self.desc), C.byref(x.v), L.ptr(self.w4[li]), L.ptr(self.packs[li][0]), C.byref(t.v), None, _st())
        a = _nhwc(x.v.n, x.v.h, x.v.w, co, x.t.device, self.dtype)
        L.call('b200gan_bias_relu_d2s', C.byref(t.v), L.ptr(self.convs[li].bias.detach()), C.byref(a.v), _st())
        return a

    def features(self, x, save: bool):
        """Returns ([relu1_2, relu2_2, relu3_3] as Acts, tape or None)."""
        cur = self._image(x)
        feats, tape = [], ([] if save else None)
        for li in range(7):
            lay = _Layer()
            lay.pool_src = None
            if li in POOL_BEFORE:
                lay.pool_src = cur
                p = _nhwc(cur.v.n, cur.v.h // 2, cur.v.w // 2, cur.v.c, cur.t.device, self.dtype)
                L.call('b200gan_maxpool2_fwd', C.byref(cur.v), C.byref(p.v), _st())
                cur = p
            lay.x = cur
            lay.a = cur = self._conv_relu(li, cur)
            if save:
                tape.append(lay)
            if li in BLOCK_END:
                feats.append(cur)
        return feats, tape

    # -- the loss -------------------------------------------------------------------------------------------------------------------------
    def loss_and_grad(self, x, y, weight: float = 1.0, dx_into: Optional[Act] = None):
        """loss = sum_b mean((f_b(x) - f_b(y))^2) (a device scalar, unweighted) and, when `dx_into` is given, dx_into += weight * d loss / d x
        (an Act over the real image channels, any layout)."""
        fy, _ = self.features(y, save=False)
        fx, tape = self.features(x, save=dx_into is not None)
        dev = fx[0].t.device
        sums = torch.zeros(3, device=dev, dtype=torch.float64)
        numel = [f.t.numel() for f in fx]
        if dx_into is None:
            for b in range(3):
                L.call('b200gan_fm_pair', C.byref(fy[b].v), C.byref(fx[b].v), None, 0.0, 0, C.c_void_p(sums.data_ptr() + 8 * b), _st())
        else:
            d = None
            for li in range(6, -1, -1):
                lay = tape[li]
                if d is None:
                    d = Act(torch.empty_like(lay.a.t), nchw=False)
                if li in BLOCK_END:
                    b = BLOCK_END.index(li)
                    L.call('b200gan_fm_pair', C.byref(fy[b].v), C.byref(lay.a.v), C.byref(d.v), -2.0 * weight / numel[b], 0 if li == 6 else 1,
                           C.c_void_p(sums.data_ptr() + 8 * b), _st())
                co = self.convs[li].out_channels
                dt = _nhwc(lay.a.v.n, lay.a.v.h // 2, lay.a.v.w // 2, 4 * co, dev, self.dtype)
                L.call('b200gan_relu_bwd_s2d', C.byref(d.v), C.byref(lay.a.v), C.byref(dt.v), _st())
                dx = Act(torch.empty_like(lay.x.t), nchw=False)
                L.call('b200gan_conv2d_dgrad', C.byref(self.desc), C.byref(dt.v), L.ptr(self.w4[li]), L.ptr(self.packs[li][1]), C.byref(dx.v), None, _st())
                if lay.pool_src is not None:
                    d = Act(torch.empty_like(lay.pool_src.t), nchw=False)
                    L.call('b200gan_maxpool2_bwd', C.byref(lay.pool_src.v), C.byref(dx.v), C.byref(d.v), 0, _st())
                else:
                    d = dx
            c = self.convs[0].in_channels
            real = Act(d.t[..., :c], nchw=False) if d.v.c != c else d
            L.call('b200gan_sample_axpby', C.byref(dx_into.v), None, C.byref(real.v), None, C.byref(dx_into.v), _st())
        if self._scale is None or self._scale[0] != tuple(numel):            # 1 / numel per block, built once per batch shape
            self._scale = (tuple(numel), torch.tensor([1.0 / v for v in numel], device=dev, dtype=torch.float64))
        return torch.dot(sums, self._scale[1]).float()


class _PerceptualFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, x, y):
        dx = torch.zeros_like(x, dtype=torch.float32) if x.requires_grad else None
        loss = eng.loss_and_grad(x, y, 1.0, Act(dx, nchw=True) if dx is not None else None)
        ctx.dx = dx
        return loss

    @staticmethod
    def backward(ctx, go):
        dx, ctx.dx = ctx.dx, None
        return None, (dx * go if dx is not None else None), None


def _vgg_features(weights):
    import torchvision.models as models
    if weights == 'imagenet':
        return models.vgg16(weights=models.VGG16_Weights.IMAGENET1K_V1).features      # train_cgan.py:60 (downloads unless cached)
    vgg = models.vgg16(weights=None)
    if weights != 'random':
        vgg.load_state_dict(torch.load(weights, map_location='cpu'))
    return vgg.features


class PerceptualLoss(nn.Module):
    """Drop-in for train_cgan.py:57-73.  weights: 'imagenet' (the reference's behaviour), a path to a torchvision vgg16 state_dict, or 'random'."""

    def __init__(self, weights='imagenet'):
        super().__init__()
        vgg = _vgg_features(weights)
        self.blocks = nn.ModuleList([vgg[:4], vgg[4:9], vgg[9:16]]).eval()
        for p in self.parameters():
            p.requires_grad = False

    def _engine_for(self):
        dtype = getattr(self, 'compute_dtype', None) or default_compute_dtype()
        cache = self.__dict__.setdefault('_b200_engines', {})
        key = (dtype, default_algo(), next(self.parameters()).device)
        if key not in cache:
            cache[key] = PerceptualEngine(self.blocks, dtype=dtype)
        return cache[key]

    def forward(self, x, y):
        if not x.is_cuda:
            total = 0.0
            for block in self.blocks:
                x, y = block(x), block(y)
                total = total + torch.mean((x - y) ** 2)
            return total
        eng = self._engine_for()
        if torch.is_grad_enabled() and x.requires_grad:
            return _PerceptualFn.apply(eng, x, y)
        return eng.loss_and_grad(x, y)
