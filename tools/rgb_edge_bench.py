#!/usr/bin/env python
"""Development aid: the five image-side calls of the nc=3 DCGAN iteration alone (batch 512 by default), L2 flushed, with the bytes each must move and
the fraction of the measured copy bandwidth it reaches.  usage: rgb_edge_bench.py [B] [once]   (`once`: one launch each, for ncu)"""
import ctypes as C, json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import gan_enhanced_pneumonia_classifier_b200 as pkg
L = pkg._lib; st = L.stream_ptr; bf = torch.bfloat16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
once = len(sys.argv) > 2
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbps', 6542.1) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6542.1
flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)
real = torch.rand((B, 3, 224, 224), device='cuda') * 2 - 1                       # fp32 NCHW, as the loader delivers it
fake = torch.tanh(torch.randn((B, 224, 224, 3), device='cuda')).to(bf)           # NHWC bf16, as the Generator leaves it
dfake = torch.randn((B, 224, 224, 3), device='cuda').to(bf)
a0 = torch.randn((B, 112, 112, 32), device='cuda').to(bf); d0 = torch.randn_like(a0); a4 = torch.randn_like(a0); yprev = torch.randn_like(a0)
wD = torch.randn((32, 3, 4, 4), device='cuda') * 0.02; wG = torch.randn((32, 3, 4, 4), device='cuda') * 0.02
dwD = torch.zeros_like(wD); dwG = torch.zeros_like(wG)
coef = [torch.rand(32, device='cuda') + 0.5 for _ in range(4)]; sums = torch.zeros(64, device='cuda', dtype=torch.float64)
cv = L.Conv(4, 2, 1, L.ALGO_AUTO)
V, VC = (lambda t: L.view_nhwc(t)), (lambda t: L.view_nchw(t))
def call(name, *args): L.call(name, *args)
MB = 1e6
cases = []
f = L.fuse(out_act=L.ACT_LRELU, out_slope=0.2)
cases.append(('D0 forward, real (fp32 NCHW)', real.numel() * 4 + a0.numel() * 2, lambda: call('b200gan_conv2d_fprop', C.byref(cv), C.byref(VC(real)), L.ptr(wD), None, C.byref(V(a0)), C.byref(f), st())))
cases.append(('D0 forward, fake (bf16 NHWC)', fake.numel() * 2 + a0.numel() * 2, lambda: call('b200gan_conv2d_fprop', C.byref(cv), C.byref(V(fake)), L.ptr(wD), None, C.byref(V(a0)), C.byref(f), st())))
vf, vy = V(fake), V(yprev)
f2 = L.fuse(dy_act=L.ACT_TANH, dy_ref=vf, prev_act=L.ACT_RELU, prev_y=vy, prev_scale=coef[0], prev_shift=coef[1], prev_mean=coef[2], prev_invstd=coef[3], prev_sums=sums)
cases.append(('G5 input gradient (+tanh backward, BatchNorm-backward sums)', dfake.numel() * 2 * 2 + a4.numel() * 2 * 2,
              lambda: call('b200gan_convT2d_dgrad', C.byref(cv), C.byref(V(dfake)), L.ptr(wG), None, C.byref(V(d0)), C.byref(f2), st())))
f3 = L.fuse(out_act=L.ACT_TANH)
cases.append(('G5 forward (+tanh)', a4.numel() * 2 + fake.numel() * 2, lambda: call('b200gan_convT2d_fprop', C.byref(cv), C.byref(V(a4)), L.ptr(wG), None, C.byref(V(fake)), C.byref(f3), st())))
cases.append(('D0 input gradient', d0.numel() * 2 + dfake.numel() * 2, lambda: call('b200gan_conv2d_dgrad', C.byref(cv), C.byref(V(d0)), L.ptr(wD), None, C.byref(V(dfake)), None, st())))
cases.append(('D0 weight gradient, real', real.numel() * 4 + d0.numel() * 2, lambda: call('b200gan_conv2d_wgrad', C.byref(cv), C.byref(VC(real)), C.byref(V(d0)), L.ptr(dwD), None, None, st())))
cases.append(('D0 weight gradient, fake', fake.numel() * 2 + d0.numel() * 2, lambda: call('b200gan_conv2d_wgrad', C.byref(cv), C.byref(V(fake)), C.byref(V(d0)), L.ptr(dwD), None, None, st())))
f4 = L.fuse(dy_act=L.ACT_TANH, dy_ref=vf)
cases.append(('G5 weight gradient (+tanh backward)', a4.numel() * 2 + dfake.numel() * 2 * 2, lambda: call('b200gan_convT2d_wgrad', C.byref(cv), C.byref(V(a4)), C.byref(V(dfake)), L.ptr(dwG), None, C.byref(f4), st())))
tot_us = tot_floor = 0.0
for name, nbytes, fn in cases:
    if once:
        fn(); torch.cuda.synchronize(); continue
    for _ in range(3): fn()
    torch.cuda.synchronize(); t = 0.0
    for _ in range(10):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); t += e0.elapsed_time(e1)
    us = t / 10 * 1e3; floor = nbytes / (peak * 1e9) * 1e6
    tot_us += us; tot_floor += floor
    print(f'{us:8.1f} us  {nbytes / MB:7.0f} MB  floor {floor:6.1f} us  {floor / us:5.2f} of the copy bandwidth   {name}')
if not once:
    print(f'sum {tot_us:.0f} us against {tot_floor:.0f} us of traffic')
