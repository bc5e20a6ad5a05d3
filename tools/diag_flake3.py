"""Eager fp32 small-trainer runs with per-step snapshots of the parameter / gradient arenas: characterise the rare divergent run."""
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
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
runs = []
for r in range(N):
    G, D = build(m, torch.float32)
    tr = DCGANTrainer(G, D, dtype=torch.float32, use_graph=False)
    snaps = []
    for z in noises:
        h = tr.step(real, z)
        snaps.append(dict(h=h.cpu().numpy(), pG=tr.arenaG.param.cpu().numpy().copy(), pD=tr.arenaD.param.cpu().numpy().copy(),
                          gG=tr.arenaG.grad.cpu().numpy().copy(), gD=tr.arenaD.grad.cpu().numpy().copy()))
    runs.append(snaps)
    names = [k for k, _ in G.named_parameters()]
    slicesG = tr.arenaG.slices
ref = runs[0]
# pick as reference a run that agrees with the majority at the last step
hl = np.stack([r[-1]['h'] for r in runs]); med = np.median(hl, axis=0)
ok = [i for i in range(N) if np.abs(hl[i] - med).max() < 1e-5]
ref = runs[ok[0]]
print('runs', N, 'divergent', N - len(ok))
for i in range(N):
    if i in ok: continue
    for s in range(5):
        for key in ('gD', 'pD', 'gG', 'pG'):
            d = np.abs(runs[i][s][key] - ref[s][key])
            scale = np.abs(ref[s][key]).max()
            big = d > (1e-6 if key[0] == 'p' else 1e-5 * scale)
            if big.any():
                idx = np.nonzero(big)[0]
                owner = ''
                if key == 'pG' or key == 'gG':
                    owner = sorted({names[j] for j, (a, b) in enumerate(slicesG) for t in idx[:2000] if a <= t < b})[:6]
                print(f'run {i} step {s + 1} {key}: {big.sum()} of {d.size} entries differ, max {d.max():.3e} (scale {scale:.2e}), median of differing {np.median(d[big]):.2e} {owner}')
    print('   history dev per step', [f"{np.abs(runs[i][s]['h'] - ref[s]['h']).max():.1e}" for s in range(5)])
