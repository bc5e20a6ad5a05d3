"""One eager fp32 trainer run (small nets, 3 steps) for compute-sanitizer --tool initcheck; optional NaN poisoning of the allocator."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import gan_enhanced_pneumonia_classifier_b200 as pkg
from gan_enhanced_pneumonia_classifier_b200.trainer import DCGANTrainer
from parity_utils import synthetic_noise, synthetic_real
from test_gpu_step import build
dtype = torch.bfloat16 if 'bf16' in sys.argv else torch.float32
m = dict(seed=3, nz=16, nc=1, fm=8)
real = torch.from_numpy(synthetic_real(5, 4, 1)).cuda()
noises = [torch.from_numpy(synthetic_noise(10 + i, 4, 16)).cuda() for i in range(3)]
def run(poison):
    if poison:
        x = torch.full((64 << 20,), float('nan'), device='cuda'); torch.cuda.synchronize(); del x
    G, D = build(m, dtype)
    tr = DCGANTrainer(G, D, dtype=dtype, use_graph=False)
    return torch.stack([tr.step(real, z) for z in noises]).cpu().numpy()
if 'poison' in sys.argv:
    a = run(False); b = run(True); c = run(True)
    print('clean vs poisoned', np.abs(a - b).max(), np.abs(a - c).max(), 'nan' if np.isnan(b).any() or np.isnan(c).any() else 'finite')
else:
    print(run(False))
