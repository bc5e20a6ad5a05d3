"""Time the tensor-core weight gradient of the middle layers alone (workspace path, L2 flushed)."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import gan_enhanced_pneumonia_classifier_b200 as pkg
L = pkg._lib; st = L.stream_ptr; bf = torch.bfloat16
flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)
cv = L.Conv(4, 2, 1, L.ALGO_TCGEN05)
for (ci, h, co) in ((32, 112, 64), (64, 56, 128), (128, 28, 256), (256, 14, 512)):
    n = 512
    x = torch.randn((n, h, h, ci), device='cuda').to(bf); dy = torch.randn((n, h // 2, h // 2, co), device='cuda').to(bf)
    dw = torch.zeros((co, ci, 4, 4), device='cuda'); ws = torch.zeros_like(dw)
    f = lambda: L.call('b200gan_conv2d_wgrad', C.byref(cv), C.byref(L.view_nhwc(x)), C.byref(L.view_nhwc(dy)), L.ptr(dw), L.ptr(ws), None, st())
    for _ in range(3): f()
    torch.cuda.synchronize(); tot = 0
    for _ in range(10):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    ms = tot / 10; fl = 2.0 * n * (h // 2) ** 2 * co * 16 * ci
    print(f'wgrad {ci}->{co} @{h}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.0f} TFLOP/s')
