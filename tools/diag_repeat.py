"""Repeat the module-level forward+backward of the small fp32 networks with fixed inputs and weights; report any repetition whose
outputs / gradients differ from the first beyond atomics noise (a sporadic race shows up as a rare large deviation)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import gan_enhanced_pneumonia_classifier_b200 as pkg
from parity_utils import synthetic_noise, synthetic_real
from test_gpu_step import build
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dtype = torch.bfloat16 if 'bf16' in sys.argv else torch.float32
m = dict(seed=3, nz=16, nc=1, fm=8)
G, D = build(m, dtype)
real = torch.from_numpy(synthetic_real(5, 4, 1)).cuda()
z = torch.from_numpy(synthetic_noise(10, 4, 16)).cuda()
g = torch.Generator(device='cuda').manual_seed(1)
rG = torch.randn((4, 1, 224, 224), device='cuda', generator=g)
rD = torch.randn((4,), device='cuda', generator=g)
def once(net, x, r, need_in):
    for p in net.parameters(): p.grad = None
    x = x.clone().requires_grad_(need_in)
    out = net(x)
    (out * r).sum().backward()
    res = {'out': out.detach().clone()}
    for k, p in net.named_parameters(): res['grad.' + k] = p.grad.detach().clone()
    if need_in: res['grad.input'] = x.grad.detach().clone()
    for k, b in net.named_buffers():
        if 'running' in k: res['buf.' + k] = b.detach().clone()
    return res
for name, net, x, r, need_in in (('G', G, z, rG, False), ('D', D, real, rD, True)):
    base = once(net, x, r, need_in)
    bad = 0
    for it in range(N):
        cur = once(net, x, r, need_in)
        for k in base:
            if k.startswith('buf.'): continue
            d = (cur[k].double() - base[k].double()).abs().max().item()
            scale = base[k].double().abs().max().item() + 1e-30
            if d / scale > (1e-5 if dtype == torch.float32 else 3e-2):
                bad += 1
                print(f'{name} rep {it}: {k} deviates {d:.3e} (scale {scale:.3e}, rel {d / scale:.2e})')
    print(name, 'done', N, 'repetitions, deviations:', bad)
