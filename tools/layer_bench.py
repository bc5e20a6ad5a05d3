#!/usr/bin/env python
"""Per-kernel timing of every tensor-core convolution of the DCGAN step (CUDA events, L2 flushed between launches).
Development aid: prints TFLOP/s per layer/primitive so that the slowest kernels can be attacked first."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gan_enhanced_pneumonia_classifier_b200 as pkg  # noqa: E402

L = pkg._lib


def timeit(fn, flush, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=512)
    ap.add_argument('--algo', default='auto')
    a = ap.parse_args()
    algo = {'auto': L.ALGO_AUTO, 'simt': L.ALGO_SIMT, 'tcgen05': L.ALGO_TCGEN05}[a.algo]
    B = a.batch
    flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)
    st = L.stream_ptr
    rows = []
    # conv geometry (Ci, H(fine), Co): D1..D4 == G4..G1
    for name, ci, h, co in [('D1/G4', 32, 112, 64), ('D2/G3', 64, 56, 128), ('D3/G2', 128, 28, 256), ('D4/G1', 256, 14, 512)]:
        x = torch.randn((B, h, h, ci), device='cuda').to(torch.bfloat16)
        dy = torch.randn((B, h // 2, h // 2, co), device='cuda').to(torch.bfloat16)
        w = torch.randn((co, ci, 4, 4), device='cuda') * 0.02
        wd = torch.empty(w.numel(), device='cuda', dtype=torch.bfloat16)
        wu = torch.empty(w.numel(), device='cuda', dtype=torch.bfloat16)
        L.call('b200gan_pack_conv_weight', L.ptr(w), co, ci, 4, 0, L.ptr(wd), st())
        L.call('b200gan_pack_conv_weight', L.ptr(w), co, ci, 4, 1, L.ptr(wu), st())
        y = torch.empty_like(dy)
        dx = torch.empty_like(x)
        dw = torch.zeros_like(w)
        ws = torch.zeros_like(w)
        cv = L.Conv(4, 2, 1, algo)
        flops = 2.0 * B * (h // 2) ** 2 * co * 16 * ci
        for prim, fn in [
            ('down(fprop)', lambda: L.call('b200gan_conv2d_fprop', C.byref(cv), C.byref(L.view_nhwc(x)), L.ptr(w), L.ptr(wd), C.byref(L.view_nhwc(y)), None, st())),
            ('up(dgrad)', lambda: L.call('b200gan_conv2d_dgrad', C.byref(cv), C.byref(L.view_nhwc(dy)), L.ptr(w), L.ptr(wu), C.byref(L.view_nhwc(dx)), None, st())),
            ('wgrad', lambda: L.call('b200gan_conv2d_wgrad', C.byref(cv), C.byref(L.view_nhwc(x)), C.byref(L.view_nhwc(dy)), L.ptr(dw), L.ptr(ws), None, st())),
        ]:
            ms = timeit(fn, flush)
            byt = (x.numel() + dy.numel()) * 2
            rows.append(dict(layer=name, prim=prim, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1), min_gbs=round(byt / ms / 1e6, 0)))
            print(rows[-1], flush=True)
    print(json.dumps(rows))


if __name__ == '__main__':
    main()
