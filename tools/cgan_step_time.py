#!/usr/bin/env python
"""Development aid: time of one conditional-GAN iteration (CGANTrainer.step, CLI-default widths: feature_maps 32, nc 3) and its per-kernel breakdown."""
import argparse, collections, os, re, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gan_enhanced_pneumonia_classifier_b200 import cgan
from gan_enhanced_pneumonia_classifier_b200.cgan_trainer import CGANTrainer
from torch.profiler import ProfilerActivity, profile
ap = argparse.ArgumentParser(); ap.add_argument('--batch', type=int, default=1024); ap.add_argument('--nc', type=int, default=3)
ap.add_argument('--nf', type=int, default=32); ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--perceptual', action='store_true', help='include the VGG16 perceptual term (random weights)')
a = ap.parse_args()
torch.manual_seed(0)
G, D = cgan.Generator(100, 2, a.nc, a.nf).cuda(), cgan.Discriminator(2, a.nc, a.nf).cuda()
vgg = None
if a.perceptual:
    from gan_enhanced_pneumonia_classifier_b200.perceptual import PerceptualLoss
    vgg = PerceptualLoss('random').cuda()
tr = CGANTrainer(G, D, perceptual=vgg, perceptual_weight=10.0 if vgg is not None else 0.0, dtype=torch.bfloat16)
real = torch.rand((a.batch, a.nc, 224, 224), device='cuda') * 2 - 1
labels = torch.randint(0, 2, (a.batch,), device='cuda')
for _ in range(2):
    out = tr.step(real, labels)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    out = tr.step(real, labels)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(f'CGAN iteration (B={a.batch}, nc={a.nc}, nf={a.nf}, perceptual={a.perceptual}): {ms:.2f} ms -> {a.batch / ms * 1e3:.0f} images/s; last history {out.tolist()}')
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(real, labels); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r'\(.*', '', e.name.replace('(anonymous namespace)::', '')).replace('void b200gan::', '').replace('b200gan::', '')
        agg[name[:90]][0] += 1; agg[name[:90]][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f'total GPU kernel time {tot / 1e3:.3f} ms per iteration')
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print(f'{t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  n={n:4d} each={t / n:9.1f} us  {k}')
