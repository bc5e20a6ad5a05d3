#!/usr/bin/env python
"""Development aid: per-kernel time of the fused training step via torch.profiler (CUPTI), cheaper than an ncu pass."""
import argparse, os, sys, collections, re
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gan_enhanced_pneumonia_classifier_b200 as pkg
from gan_enhanced_pneumonia_classifier_b200.trainer import DCGANTrainer
from torch.profiler import profile, ProfilerActivity

ap = argparse.ArgumentParser(); ap.add_argument('--batch', type=int, default=512); ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--nc', type=int, default=1); ap.add_argument('--dtype', default='bf16')
a = ap.parse_args()
torch.manual_seed(0)
G, D = pkg.Generator(100, a.nc, 64).cuda(), pkg.Discriminator(a.nc, 64).cuda()
tr = DCGANTrainer(G, D, dtype=torch.bfloat16 if a.dtype == 'bf16' else torch.float32)
real = torch.rand((a.batch, a.nc, 224, 224), device='cuda') * 2 - 1
for _ in range(3):
    tr.step(real, torch.randn((a.batch, 100, 1, 1), device='cuda'))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(a.steps):
        tr.step(real, torch.randn((a.batch, 100, 1, 1), device='cuda'))
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r'\(.*', '', e.name.replace('(anonymous namespace)::', '')).replace('void b200gan::', '').replace('b200gan::', '')
        agg[name[:80]][0] += 1; agg[name[:80]][1] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
tot = sum(v[1] for v in agg.values())
print(f'total GPU kernel time {tot/1e3/a.steps:.3f} ms/step over {a.steps} steps')
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print(f'{t/1e3/a.steps:9.3f} ms/step {100*t/tot:5.1f}%  n/step={n/a.steps:5.1f} each={t/n:9.1f} us  {k}')
