#!/usr/bin/env python
"""Launch each hot kernel of the step once or twice at the benchmark shapes (B=512 by default) so that one short
`ncu --set full` pass can capture them (summaries under profiles/).  usage: ncu_target.py [B] [reps] [which,...]"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gan_enhanced_pneumonia_classifier_b200 as pkg
L = pkg._lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
which = set(sys.argv[3].split(',')) if len(sys.argv) > 3 else {'thin', 'd1', 'd3', 'ew'}
st = L.stream_ptr
bf = torch.bfloat16
def rnd(shape, dt=bf): return torch.randn(shape, device='cuda').to(dt)
def V(t): return C.byref(L.view_nhwc(t))
auto = L.Conv(4, 2, 1, L.ALGO_AUTO)
SL = 0.2

def mid(ci, h, co):
    x = rnd((B, h, h, ci)); dy = rnd((B, h // 2, h // 2, co)); w = rnd((co, ci, 4, 4), torch.float32) * 0.02
    wd = torch.empty(w.numel(), device='cuda', dtype=bf); wu = torch.empty(w.numel(), device='cuda', dtype=bf)
    L.call('b200gan_pack_conv_weight', L.ptr(w), co, ci, 4, 0, L.ptr(wd), st()); L.call('b200gan_pack_conv_weight', L.ptr(w), co, ci, 4, 1, L.ptr(wu), st())
    d = dict(x=x, dy=dy, w=w, wd=wd, wu=wu, y=torch.empty_like(dy), dx=torch.empty_like(x), dw=torch.zeros_like(w))
    d['sums_y'] = torch.zeros(2 * co, device='cuda', dtype=torch.float64)
    d['sums_x'] = torch.zeros(2 * ci, device='cuda', dtype=torch.float64)
    d['coef_x'] = [torch.rand(ci, device='cuda') + 0.5 for _ in range(4)]
    d['yprev'] = rnd((B, h, h, ci))
    return d

def run_mid(d):
    f1 = L.fuse(bn_sums=d['sums_y'])
    L.call('b200gan_conv2d_fprop', C.byref(auto), V(d['x']), L.ptr(d['w']), L.ptr(d['wd']), V(d['y']), C.byref(f1), st())
    yv = L.view_nhwc(d['yprev'])
    f2 = L.fuse(prev_act=L.ACT_LRELU, prev_slope=SL, prev_y=yv, prev_scale=d['coef_x'][0], prev_shift=d['coef_x'][1], prev_mean=d['coef_x'][2],
                prev_invstd=d['coef_x'][3], prev_sums=d['sums_x'])
    L.call('b200gan_conv2d_dgrad', C.byref(auto), V(d['dy']), L.ptr(d['w']), L.ptr(d['wu']), V(d['dx']), C.byref(f2), st())
    L.call('b200gan_conv2d_dgrad', C.byref(auto), V(d['dy']), L.ptr(d['w']), L.ptr(d['wu']), V(d['dx']), None, st())
    L.call('b200gan_conv2d_wgrad', C.byref(auto), V(d['x']), V(d['dy']), L.ptr(d['dw']), None, None, st())

if 'thin' in which:
    img = rnd((B, 224, 224, 1)); img32 = torch.randn((B, 1, 224, 224), device='cuda')
    c32 = rnd((B, 112, 112, 32)); a0 = rnd((B, 112, 112, 32)); w0 = rnd((32, 1, 4, 4), torch.float32) * 0.02
    out32 = torch.empty_like(c32); outimg = torch.empty_like(img); dw0 = torch.zeros_like(w0)
if 'd1' in which: d1 = mid(32, 112, 64)
if 'd3' in which: d3 = mid(128, 28, 256)
if 'lat' in which:
    zl = torch.randn((B, 100, 1, 1), device='cuda'); wl = torch.randn((100, 512, 7, 7), device='cuda') * 0.02
    yl = torch.empty((B, 7, 7, 512), device='cuda', dtype=bf); dyl = rnd((B, 7, 7, 512)); dwl = torch.zeros_like(wl)
    cvl = L.Conv(7, 1, 0, L.ALGO_AUTO)
if 'mask' in which:
    dm = mid(32, 112, 64); a0m = rnd((B, 112, 112, 32))
if 'ew' in which:
    yb = rnd((B, 56, 56, 64)); dab = rnd((B, 56, 56, 64)); ab = torch.empty_like(yb)
    sc, sh, mu, isd, gam = (torch.rand(64, device='cuda') + 0.5 for _ in range(5)); sums = torch.zeros(128, device='cuda', dtype=torch.float64)
for _ in range(reps):
    if 'thin' in which:
        a0v = L.view_nhwc(a0)
        L.call('b200gan_conv2d_fprop', C.byref(auto), C.byref(L.view_nchw(img32)), L.ptr(w0), None, V(out32), C.byref(L.fuse(out_act=L.ACT_LRELU, out_slope=SL)), st())
        L.call('b200gan_convT2d_fprop', C.byref(auto), V(c32), L.ptr(w0), None, V(outimg), C.byref(L.fuse(out_act=L.ACT_TANH)), st())
        fz = L.fuse(dy_act=L.ACT_LRELU, dy_slope=SL, dy_ref=a0v)
        L.call('b200gan_conv2d_dgrad', C.byref(auto), V(c32), L.ptr(w0), None, V(outimg), C.byref(fz), st())
        L.call('b200gan_conv2d_wgrad', C.byref(auto), C.byref(L.view_nchw(img32)), V(c32), L.ptr(dw0), None, C.byref(fz), st())
    if 'd1' in which: run_mid(d1)
    if 'd3' in which: run_mid(d3)
    if 'lat' in which:
        L.call('b200gan_convT2d_fprop', C.byref(cvl), C.byref(L.view_nchw(zl)), L.ptr(wl), None, V(yl), None, st())
        L.call('b200gan_convT2d_wgrad', C.byref(cvl), C.byref(L.view_nchw(zl)), V(dyl), L.ptr(dwl), None, None, st())
    if 'mask' in which:
        fm = L.fuse(prev_act=L.ACT_LRELU, prev_slope=SL, prev_y=L.view_nhwc(a0m))
        L.call('b200gan_conv2d_dgrad', C.byref(auto), V(dm['dy']), L.ptr(dm['w']), L.ptr(dm['wu']), V(dm['dx']), C.byref(fm), st())
        yv2 = L.view_nhwc(dm['y'])
        f2 = L.fuse(prev_act=L.ACT_RELU, prev_y=yv2, prev_scale=dm['coef_x'][0].repeat(2), prev_shift=dm['coef_x'][1].repeat(2), prev_mean=dm['coef_x'][2].repeat(2),
                    prev_invstd=dm['coef_x'][3].repeat(2), prev_sums=dm['sums_y'])
        L.call('b200gan_convT2d_dgrad', C.byref(auto), V(dm['x']), L.ptr(dm['w']), L.ptr(dm['wd']), V(dm['y'].clone()), C.byref(f2), st())
    if 'ew' in which:
        L.call('b200gan_bn_act_fwd', V(yb), L.ptr(sc), L.ptr(sh), L.ACT_LRELU, SL, V(ab), st())
        L.call('b200gan_bn_act_bwd_apply', V(dab), V(yb), None, L.ptr(sc), L.ptr(sh), L.ptr(mu), L.ptr(isd), L.ptr(gam), L.ptr(sums), B * 56 * 56, L.ACT_NONE, SL, V(dab), None, None, st())
    torch.cuda.synchronize()
print('done')
