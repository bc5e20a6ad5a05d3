#!/usr/bin/env python
"""Development aid: the D3 forward convolution (Conv2d 128->256, 28x28 -> 14x14, B=512) alone on the CTA-pair kernel at several cluster
counts (B200GAN_PAIR_CLUSTERS), CUDA events, L2 flushed.  Needs B200GAN_PAIR=1 (the pair kernel is opt-in); without it the one-CTA kernel is timed."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gan_enhanced_pneumonia_classifier_b200 as pkg
L = pkg._lib
n, ci, h, co = 512, 128, 28, 256
x = torch.randn((n, h, h, ci), device='cuda').bfloat16()
w = torch.randn((co, ci, 4, 4), device='cuda') * 0.02
y = torch.empty((n, h // 2, h // 2, co), device='cuda', dtype=torch.bfloat16)
wp = torch.empty(w.numel(), device='cuda', dtype=torch.bfloat16)
cv = L.Conv(4, 2, 1, L.ALGO_TCGEN05)
st = L.stream_ptr
L.call('b200gan_pack_conv_weight', L.ptr(w), co, ci, 4, 0, L.ptr(wp), st())
sums = torch.zeros(2 * co, device='cuda', dtype=torch.float64)
flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)
def run(fz):
    L.call('b200gan_conv2d_fprop', C.byref(cv), C.byref(L.view_nhwc(x)), L.ptr(w), L.ptr(wp), C.byref(L.view_nhwc(y)), C.byref(fz) if fz else None, st())
def timeit(fz, iters=10):
    for _ in range(3): run(fz)
    torch.cuda.synchronize(); tot = 0
    for _ in range(iters):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(fz); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / iters * 1e3
flops = 2.0 * n * 196 * co * 16 * ci
for k in (74, 73, 72, 70, 64, 56, 37, 20):
    os.environ['B200GAN_PAIR_CLUSTERS'] = str(k)
    for name, fz in (('plain', None), ('stats', L.fuse(bn_sums=sums))):
        us = timeit(fz)
        print(f'clusters={k:3d} {name}: {us:7.1f} us  {flops / us / 1e6:7.1f} TFLOP/s', flush=True)
