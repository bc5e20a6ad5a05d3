"""G5 input gradient (ConvTranspose2d(32->1) dgrad = 'down' 1->32 channels) alone: with / without the tanh backward on the gradient
operand and with / without the BatchNorm-backward epilogue."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import gan_enhanced_pneumonia_classifier_b200 as pkg
L = pkg._lib; st = L.stream_ptr; bf = torch.bfloat16
n = 512
flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)
dimg = torch.randn((n, 224, 224, 1), device='cuda').to(bf); img = torch.tanh(torch.randn((n, 224, 224, 1), device='cuda')).to(bf)
w = torch.randn((32, 1, 4, 4), device='cuda') * 0.02
dx = torch.empty((n, 112, 112, 32), device='cuda', dtype=bf); yprev = torch.randn((n, 112, 112, 32), device='cuda').to(bf)
coef = [torch.rand(32, device='cuda') + 0.5 for _ in range(4)]; sums = torch.zeros(64, device='cuda', dtype=torch.float64)
cv = L.Conv(4, 2, 1, L.ALGO_AUTO)
for ref in (False, True):
    for bn in (False, True):
        kw = {}
        if ref: kw.update(dy_act=L.ACT_TANH, dy_ref=L.view_nhwc(img))
        if bn: kw.update(prev_act=L.ACT_RELU, prev_y=L.view_nhwc(yprev), prev_scale=coef[0], prev_shift=coef[1], prev_mean=coef[2], prev_invstd=coef[3], prev_sums=sums)
        fz = L.fuse(**kw) if kw else None
        f = lambda: L.call('b200gan_convT2d_dgrad', C.byref(cv), C.byref(L.view_nhwc(dimg)), L.ptr(w), None, C.byref(L.view_nhwc(dx)), C.byref(fz) if fz is not None else None, st())
        for _ in range(3): f()
        torch.cuda.synchronize(); tot = 0
        for _ in range(10):
            flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
        print(f'tanh-backward on operand {ref}, BatchNorm-backward epilogue {bn}: {tot / 10 * 1e3:.1f} us')
