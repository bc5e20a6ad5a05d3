"""Diagnose run-to-run differences of the fp32 trainer: eager vs eager vs graph, 5 steps, small nets (the test's configuration)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import dcgan_oracle as orc
import gan_enhanced_pneumonia_classifier_b200 as pkg
from gan_enhanced_pneumonia_classifier_b200.trainer import DCGANTrainer
from parity_utils import synthetic_noise, synthetic_real
from test_gpu_step import build
m = dict(seed=3, nz=16, nc=1, fm=8)
real = torch.from_numpy(synthetic_real(5, 4, 1)).cuda()
noises = [torch.from_numpy(synthetic_noise(10 + i, 4, 16)).cuda() for i in range(5)]
def run(mode, steps=5):
    G, D = build(m, torch.float32)
    tr = DCGANTrainer(G, D, dtype=torch.float32, use_graph=mode)
    h = torch.stack([tr.step(real, z) for z in noises[:steps]]).cpu().numpy()
    sd = {('G.' + k): v.detach().cpu().numpy().copy() for k, v in G.state_dict().items()}
    sd.update({('D.' + k): v.detach().cpu().numpy().copy() for k, v in D.state_dict().items()})
    return h, sd
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    for steps in (1, 5):
        h0, s0 = run(False, steps); h1, s1 = run(False, steps); h2, s2 = run(True, steps)
        def cmp(a, b):
            out = []
            for k in a:
                if a[k].dtype.kind != 'f': continue
                d = np.abs(a[k].astype(np.float64) - b[k]); 
                if d.max() > 0: out.append((k, float(d.max()), float((d > 2e-6 + 1e-4 * np.abs(b[k])).mean())))
            return out
        ee, eg = cmp(s0, s1), cmp(s0, s2)
        worst = lambda l: sorted(l, key=lambda t: -t[2])[:3]
        print(f'rep {rep} steps {steps}: eager-eager differing tensors {len(ee)} hist diff {np.abs(h0 - h1).max():.2e} worst {worst(ee)}')
        print(f'            eager-graph differing tensors {len(eg)} hist diff {np.abs(h0 - h2).max():.2e} worst {worst(eg)}')
