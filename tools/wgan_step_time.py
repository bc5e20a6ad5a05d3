#!/usr/bin/env python
"""Development aid: time of one WGAN-GP iteration (critic_iters critic updates + one generator update, BASELINE.json configs[4] widths) and
its per-kernel breakdown."""
import argparse, collections, os, re, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gan_enhanced_pneumonia_classifier_b200 import wggan
from gan_enhanced_pneumonia_classifier_b200.wgan_trainer import WGANGPTrainer
from torch.profiler import ProfilerActivity, profile
ap = argparse.ArgumentParser(); ap.add_argument('--batch', type=int, default=512); ap.add_argument('--nc', type=int, default=1)
ap.add_argument('--critic-iters', type=int, default=5); ap.add_argument('--steps', type=int, default=3)
a = ap.parse_args()
torch.manual_seed(0)
G, D = wggan.Generator(100, a.nc, 64).cuda(), wggan.Discriminator(a.nc, 64).cuda()
tr = WGANGPTrainer(G, D, critic_iters=a.critic_iters, dtype=torch.bfloat16)
real = torch.rand((a.batch, a.nc, 224, 224), device='cuda') * 2 - 1
for _ in range(2):
    out = tr.step(real)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    out = tr.step(real)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(f'WGAN-GP iteration (B={a.batch}, nc={a.nc}, critic_iters={a.critic_iters}): {ms:.2f} ms -> {a.batch / ms * 1e3:.0f} images/s; last losses {out.tolist()}')
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(real); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r'\(.*', '', e.name.replace('(anonymous namespace)::', '')).replace('void b200gan::', '').replace('b200gan::', '')
        agg[name[:80]][0] += 1; agg[name[:80]][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f'total GPU kernel time {tot / 1e3:.3f} ms per iteration')
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
    print(f'{t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  n={n:4d} each={t / n:9.1f} us  {k}')
