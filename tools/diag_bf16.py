"""Development aid: per-tensor error table of the bf16 mode against the golden fixture (and fp32 mode)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import dcgan_oracle as orc
import gan_enhanced_pneumonia_classifier_b200 as pkg
from parity_utils import synthetic_noise, synthetic_real
from test_gpu_step import build, reference_loop_step

name = sys.argv[1] if len(sys.argv) > 1 else 'step_small_nc1.npz'
g = np.load(os.path.join(ROOT, 'tests', 'golden', name)); m = json.loads(str(g['meta']))
res = {}
for tag, dtype, algo in (('fp32', torch.float32, 'auto'), ('bf16-simt', torch.bfloat16, 'simt'), ('bf16-auto', torch.bfloat16, 'auto')):
    os.environ['B200GAN_ALGO'] = algo
    G, D = build(m, dtype)
    optD = torch.optim.Adam(D.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999)); optG = torch.optim.Adam(G.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    real = torch.from_numpy(synthetic_real(m['real_seed'], m['batch'], m['nc'])).cuda()
    noises = synthetic_noise(m['noise_seed'], m['batch'] * m['iters'], m['nz']).reshape(m['iters'], m['batch'], m['nz'], 1, 1)
    res[tag] = reference_loop_step(G, D, optG, optD, real, torch.from_numpy(noises[0]).cuda())
for k in ('errG', 'errD', 'D_x', 'D_G_z1', 'D_G_z2'):
    print(k, float(g[f'it0.{k}']), {t: round(res[t][k], 6) for t in res})
for net in ('grads_D', 'grads_G'):
    for k in res['fp32'][net]:
        ref = g[f'it0.{net}.{k}'].astype(np.float64)
        row = []
        for t in res:
            a = res[t][net][k].astype(np.float64)
            row.append(f"{t}: relL2 {np.sqrt(((a-ref)**2).sum())/np.sqrt((ref**2).sum()):.2e} max {np.abs(a-ref).max()/np.abs(ref).max():.2e}")
        print(f'{net}.{k:16s}', ' | '.join(row))
