#!/usr/bin/env python
"""Launch ONE convolution of the step a few times (for a single-kernel ncu capture). usage: one_kernel.py {d1_up|d1_down|d3_down} [B]"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gan_enhanced_pneumonia_classifier_b200 as pkg
L = pkg._lib
which = sys.argv[1] if len(sys.argv) > 1 else 'd1_up'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
st = L.stream_ptr
bf = torch.bfloat16
cfg = {'d1_up': (32, 112, 64), 'd1_down': (32, 112, 64), 'd3_down': (128, 28, 256), 'd2_up': (64, 56, 128)}[which]
ci, h, co = cfg
x = torch.randn((B, h, h, ci), device='cuda').to(bf); dy = torch.randn((B, h // 2, h // 2, co), device='cuda').to(bf)
w = torch.randn((co, ci, 4, 4), device='cuda') * 0.02
wd = torch.empty(w.numel(), device='cuda', dtype=bf); wu = torch.empty(w.numel(), device='cuda', dtype=bf)
L.call('b200gan_pack_conv_weight', L.ptr(w), co, ci, 4, 0, L.ptr(wd), st()); L.call('b200gan_pack_conv_weight', L.ptr(w), co, ci, 4, 1, L.ptr(wu), st())
y = torch.empty_like(dy); dx = torch.empty_like(x)
cv = L.Conv(4, 2, 1, L.ALGO_TCGEN05)
epi = sys.argv[3] if len(sys.argv) > 3 else 'none'
aprev = torch.randn((B, h, h, ci), device='cuda').to(bf)
sums = torch.zeros(2 * ci, device='cuda', dtype=torch.float64)
apv = L.view_nhwc(aprev)
fz = None
if epi == 'mask':
    fz = L.fuse(prev_act=L.ACT_LRELU, prev_slope=0.2, prev_y=apv)
coef_u = torch.rand((4, ci), device='cuda') + 0.5
flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)
def up():
    if epi == 'bnbwd':
        f1 = L.fuse(prev_act=L.ACT_LRELU, prev_slope=0.2, prev_y=apv, prev_scale=coef_u[0], prev_shift=coef_u[1], prev_mean=coef_u[2], prev_invstd=coef_u[3], prev_sums=sums)
        L.call('b200gan_conv2d_dgrad', C.byref(cv), C.byref(L.view_nhwc(dy)), L.ptr(w), L.ptr(wu), C.byref(L.view_nhwc(dx)), C.byref(f1), st())
    elif epi == 'stats':
        f1 = L.fuse(bn_sums=sums)
        L.call('b200gan_convT2d_fprop', C.byref(cv), C.byref(L.view_nhwc(dy)), L.ptr(w), L.ptr(wu), C.byref(L.view_nhwc(dx)), C.byref(f1), st())
    else:
        L.call('b200gan_conv2d_dgrad', C.byref(cv), C.byref(L.view_nhwc(dy)), L.ptr(w), L.ptr(wu), C.byref(L.view_nhwc(dx)), C.byref(fz) if fz is not None else None, st())
def down():
    if epi == 'stats':
        f1 = L.fuse(bn_sums=sums2)
    elif epi == 'bnbwd':
        f1 = L.fuse(prev_act=L.ACT_RELU, prev_y=L.view_nhwc(yprev), prev_scale=coef[0], prev_shift=coef[1], prev_mean=coef[2], prev_invstd=coef[3], prev_sums=sums2)
    else:
        f1 = None
    L.call('b200gan_convT2d_dgrad' if epi == 'bnbwd' else 'b200gan_conv2d_fprop', C.byref(cv), C.byref(L.view_nhwc(x)), L.ptr(w), L.ptr(wd), C.byref(L.view_nhwc(y)), C.byref(f1) if f1 is not None else None, st())
sums2 = torch.zeros(2 * co, device='cuda', dtype=torch.float64)
yprev = torch.randn((B, h // 2, h // 2, co), device='cuda').to(bf)
coef = torch.rand((4, co), device='cuda') + 0.5
if which.endswith('down'):
    up = down
if len(sys.argv) > 4:
    for _ in range(3): up()
    torch.cuda.synchronize(); tot = 0
    for _ in range(10):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); up(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    print(which, epi, 'ms', tot / 10); sys.exit(0)
for _ in range(3):
    if which.endswith('up'):
        up()
    else:
        L.call('b200gan_conv2d_fprop', C.byref(cv), C.byref(L.view_nhwc(x)), L.ptr(w), L.ptr(wd), C.byref(L.view_nhwc(y)), None, st())
torch.cuda.synchronize()
print('done')
