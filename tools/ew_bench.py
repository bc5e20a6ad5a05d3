"""Time bn_act_bwd_apply (in place, act=NONE) on the D1 / D2 / D4 tensors alone, L2 flushed."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import gan_enhanced_pneumonia_classifier_b200 as pkg
L = pkg._lib; st = L.stream_ptr; bf = torch.bfloat16
flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)
for (h, c) in ((56, 64), (28, 128), (7, 512)):
    n = 512
    yb = torch.randn((n, h, h, c), device='cuda').to(bf); dz = torch.randn((n, h, h, c), device='cuda').to(bf)
    vec = [torch.rand(c, device='cuda') + 0.5 for _ in range(5)]; bs = torch.zeros(2 * c, device='cuda', dtype=torch.float64)
    f = lambda: L.call('b200gan_bn_act_bwd_apply', C.byref(L.view_nhwc(dz)), C.byref(L.view_nhwc(yb)), None, L.ptr(vec[0]), L.ptr(vec[1]), L.ptr(vec[2]),
                       L.ptr(vec[3]), L.ptr(vec[4]), L.ptr(bs), n * h * h, L.ACT_NONE, 0.2, C.byref(L.view_nhwc(dz)), None, None, st())
    for _ in range(3): f()
    torch.cuda.synchronize(); tot = 0
    for _ in range(10):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    ms = tot / 10; gb = 3 * yb.numel() * 2 / 1e9
    ab = torch.empty_like(yb)
    f2 = lambda: L.call('b200gan_bn_act_fwd', C.byref(L.view_nhwc(yb)), L.ptr(vec[0]), L.ptr(vec[1]), L.ACT_LRELU, 0.2, C.byref(L.view_nhwc(ab)), st())
    for _ in range(3): f2()
    torch.cuda.synchronize(); tot = 0
    for _ in range(10):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f2(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    ms2 = tot / 10; gb2 = 2 * yb.numel() * 2 / 1e9
    print(f'{h}x{h}x{c}: bwd {ms * 1e3:.1f} us  {gb / ms * 1e3:.0f} GB/s   fwd {ms2 * 1e3:.1f} us {gb2 / ms2 * 1e3:.0f} GB/s')
