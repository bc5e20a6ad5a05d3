"""CPU oracle for the VGG16 perceptual loss of the conditional GAN (reference src/train_cgan.py:57-73: `PerceptualLoss`, used at :186 with weight
10) -- TEST INFRASTRUCTURE ONLY.  numpy restatement (NCHW, float32 storage, float64 accumulation through dcgan_oracle's convolution primitives).

The arithmetic lives in third-party dependencies: the architecture is torchvision's `models.vgg16(...).features[:16]` (torchvision 0.26 in this
image; requirements.txt leaves it unpinned) -- Conv2d(3,1,1)+ReLU x2, MaxPool2d(2,2), Conv+ReLU x2, MaxPool2d(2,2), Conv+ReLU x3, cut into the three
blocks [:4], [4:9], [9:16] -- and the loss is sum over the blocks of mean((block(x) - block(y))^2), the blocks chained (train_cgan.py:66-73).
Restated here: conv2d with bias, relu, 2x2 max pooling with ATen's tie rule (the FIRST maximum of a window in row-major order takes the gradient),
their backward w.r.t. the input; the network is frozen (train_cgan.py:64-65), so no weight gradients exist.

Pinned against torchvision itself (the same architecture instantiated with random weights: the ImageNet checkpoint the reference downloads at
train_cgan.py:60 cannot be obtained offline, so parity is pinned for the OPERATOR, on arbitrary weights, not for the checkpoint's values):
tests/test_oracle_golden.py::test_vgg_oracle_matches_torchvision.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
from __future__ import annotations

import numpy as np

import dcgan_oracle as orc

CONV_IDX = [0, 2, 5, 7, 10, 12, 14]          # positions of the Conv2d modules in vgg16.features[:16]
CHANNELS = [(3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 256)]
POOL_BEFORE = {2, 4}                         # a MaxPool2d(2,2) sits in front of these convolutions (features[4], features[9])
BLOCK_END = {1, 3, 6}                        # the feature maps compared: after relu1_2, relu2_2, relu3_3


def init_weights(rng, scale=1.0):
    """Random weights in torchvision's key layout (`{idx}.weight`, `{idx}.bias` of the features Sequential); He-style so that activations keep O(1) scale."""
    sd = {}
    for idx, (ci, co) in zip(CONV_IDX, CHANNELS):
        sd[f'{idx}.weight'] = (rng.randn(co, ci, 3, 3) * scale * np.sqrt(2.0 / (ci * 9))).astype(np.float32)
        sd[f'{idx}.bias'] = (rng.randn(co) * 0.05).astype(np.float32)
    return sd


def maxpool2(a):
    n, c, h, w = a.shape
    return a.reshape(n, c, h // 2, 2, w // 2, 2).transpose(0, 1, 2, 4, 3, 5).reshape(n, c, h // 2, w // 2, 4).max(axis=4)


def maxpool2_bwd(a, dp):
    n, c, h, w = a.shape
    win = a.reshape(n, c, h // 2, 2, w // 2, 2).transpose(0, 1, 2, 4, 3, 5).reshape(n, c, h // 2, w // 2, 4)
    first = win.argmax(axis=4)                                   # numpy's argmax returns the first maximum, like ATen's max_pool2d
    dwin = np.zeros_like(win)
    np.put_along_axis(dwin, first[..., None], dp[..., None], axis=4)
    return dwin.reshape(n, c, h // 2, w // 2, 2, 2).transpose(0, 1, 2, 4, 3, 5).reshape(n, c, h, w)


def features(x, sd):
    """Returns ([f1, f2, f3], cache)."""
    feats, cache = [], []
    cur = x
    for li, idx in enumerate(CONV_IDX):
        pooled_from = None
        if li in POOL_BEFORE:
            pooled_from = cur
            cur = maxpool2(cur)
        a = np.maximum(orc.conv2d_fprop(cur, sd[f'{idx}.weight'], 1, 1) + sd[f'{idx}.bias'][None, :, None, None], 0)
        cache.append((cur, a, pooled_from))
        cur = a
        if li in BLOCK_END:
            feats.append(a)
    return feats, cache


def perceptual(x, y, sd):
    """Returns (loss, d loss / d x): loss = sum_b mean((f_b(x) - f_b(y))^2)."""
    fx, cache = features(x, sd)
    fy, _ = features(y, sd)
    loss = 0.0
    top = {}
    for b, (a, t) in enumerate(zip(fx, fy)):
        diff = a.astype(np.float64) - t
        loss += (diff ** 2).mean()
        top[sorted(BLOCK_END)[b]] = (2.0 * diff / diff.size).astype(x.dtype)
    d = None
    for li in reversed(range(len(CONV_IDX))):
        inp, a, pooled_from = cache[li]
        if li in top:
            d = top[li] if d is None else d + top[li]
        dz = d * (a > 0)
        d = orc.conv2d_dgrad(dz, sd[f'{CONV_IDX[li]}.weight'], 1, 1, inp.shape[2:])
        if pooled_from is not None:
            d = maxpool2_bwd(pooled_from, d)
    return np.float32(loss), d
