"""Many eager fp32 runs of the small trainer: find runs whose history departs from the majority and where it starts."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import gan_enhanced_pneumonia_classifier_b200 as pkg
from gan_enhanced_pneumonia_classifier_b200.trainer import DCGANTrainer
from parity_utils import synthetic_noise, synthetic_real
from test_gpu_step import build
m = dict(seed=3, nz=16, nc=1, fm=8)
real = torch.from_numpy(synthetic_real(5, 4, 1)).cuda()
noises = [torch.from_numpy(synthetic_noise(10 + i, 4, 16)).cuda() for i in range(5)]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
poison = 'poison' in sys.argv
graph = 'graph' in sys.argv
def dirty():
    big = torch.full((32 << 20,), float('nan'), device='cuda')
    small = [torch.full((n,), float('nan'), device='cuda') for n in (64, 128, 256, 1024, 4096, 16384, 65536, 200000) for _ in range(40)]
    dbl = [torch.full((n,), float('nan'), device='cuda', dtype=torch.float64) for n in (16, 32, 64, 128, 256) for _ in range(40)]
    torch.cuda.synchronize(); del big, small, dbl
hs = []
for r in range(N):
    if poison: dirty()
    G, D = build(m, torch.float32)
    tr = DCGANTrainer(G, D, dtype=torch.float32, use_graph=graph)
    hs.append(torch.stack([tr.step(real, z) for z in noises]).cpu().numpy())
hs = np.stack(hs)
med = np.median(hs, axis=0)
dev = np.abs(hs - med).reshape(N, 5, 5)
print('nan runs', int(np.isnan(hs).any(axis=(1, 2)).sum()), 'max dev per run:', ' '.join(f'{d.max():.1e}' for d in dev))
for r in range(N):
    if dev[r].max() > 1e-5 or np.isnan(hs[r]).any():
        print('run', r, 'per-step max dev', [f'{x:.1e}' for x in dev[r].max(axis=1)], 'first bad step row', hs[r][np.argmax(dev[r].max(axis=1) > 1e-5)], 'median row', med[np.argmax(dev[r].max(axis=1) > 1e-5)])
