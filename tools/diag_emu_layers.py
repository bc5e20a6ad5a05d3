#!/usr/bin/env python
"""Development aid: Discriminator forward (bf16 kernels) layer by layer against the bf16-storage oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import dcgan_oracle as orc
import gan_enhanced_pneumonia_classifier_b200 as pkg
from gan_enhanced_pneumonia_classifier_b200 import engine as E
from parity_utils import synthetic_real
nc, B = 1, 8
rng = np.random.RandomState(701)
sdG = orc.init_state(orc.generator_plan(100, nc, 64), True, rng)
sdD = orc.init_state(orc.discriminator_plan(nc, 64), False, rng)
D = pkg.Discriminator(nc, 64)
D.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sdD.items()})
D = D.cuda()
real = synthetic_real(702, B, nc)
eng = E.NetEngine(D._specs(), False, torch.bfloat16)
logit, ctxs = eng.forward(E.Act(torch.from_numpy(real).cuda(), nchw=True), E.params_from_module(D, eng.specs), True, True, last_act=False)
torch.cuda.synchronize()
for name, storage in (('emu', orc.bf16_round), ('fp32', None)):
    oD = orc.DiscriminatorOracle(nc, 64, {k: v.copy() for k, v in sdD.items()}, storage=storage)
    p, cache = oD.probs(real, train=True)
    print(f'--- kernels vs oracle[{name}]')
    for i, (lc, (a_in, z, out, xhat, invstd)) in enumerate(zip(ctxs, cache)):
        got = lc.a.t.float().cpu().numpy().transpose(0, 3, 1, 2)
        ref = z if i == 5 else out
        d = np.abs(got - ref)
        rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        print(f'layer {i}: relL2 {rel:.3e}  frac differing {np.mean(d > 0):.4f}  max abs {d.max():.3e}  ref rms {np.sqrt((ref**2).mean()):.3e}')
        if lc.mean is not None:
            mean_ref = None
    print('probs kernels', torch.sigmoid(logit.t.float()).view(-1).cpu().numpy())
    print('probs oracle ', p)
