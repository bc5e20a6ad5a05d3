"""Development aid: where the drop-in cgan modules and the numpy oracle part ways, layer by layer.
    python tools/diag_cgan.py cgan_step_nc1.npz        # fp32 mode on a fixture's networks
    python tools/diag_cgan.py full                     # bf16 mode, feature_maps 32 (the CLI default width)"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import cgan_oracle as co
from gan_enhanced_pneumonia_classifier_b200 import cgan
from test_oracle_golden import cgan_state

name = sys.argv[1] if len(sys.argv) > 1 else 'cgan_step_nc1.npz'
if name == 'full':
    m = dict(nz=100, nc=3, nf=32)
    torch.manual_seed(7)
    G, D = cgan.Generator(m['nz'], 2, m['nc'], m['nf']), cgan.Discriminator(2, m['nc'], m['nf'])
    with torch.no_grad():
        D.label_emb.weight.mul_(0.02)
    sdG = {k: v.numpy().copy() for k, v in G.state_dict().items()}; sdD = {k: v.numpy().copy() for k, v in D.state_dict().items()}
    rng = np.random.RandomState(3); z, fl = rng.randn(4, 100).astype(np.float32), rng.randint(0, 2, 4).astype(np.int64)
    dtype = torch.bfloat16
else:
    g = np.load(os.path.join(ROOT, 'tests', 'golden', name)); m = json.loads(str(g['meta']))
    G, D = cgan.Generator(m['nz'], 2, m['nc'], m['nf']), cgan.Discriminator(2, m['nc'], m['nf'])
    sdG, sdD = cgan_state(g, 'G'), cgan_state(g, 'D')
    G.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sdG.items()}); D.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sdD.items()})
    z, fl = g['it0.noise'], g['it0.fake_labels']
    dtype = torch.float32
G, D = G.cuda(), D.cuda(); G.compute_dtype = D.compute_dtype = dtype
oG, oD = co.GeneratorOracle(m['nz'], 2, m['nc'], m['nf'], sdG), co.DiscriminatorOracle(2, m['nc'], m['nf'], sdD)
fake_ref, cg = oG.forward(z, fl)
eng = G._engine_for()
out, (tape, _) = eng.forward(G, torch.from_numpy(z).cuda(), torch.from_numpy(fl).cuda(), save=True)
def rel(a, b): return float(np.linalg.norm((a - b).astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))
def nchw(act): return (act.t if act.nchw else act.t.permute(0, 3, 1, 2)).float().cpu().numpy()
print('G layer0 act', rel(nchw(tape[0].a), cg['bn0'][2]))
for i in range(1, 6):
    print('G layer', i, 'act relL2', rel(nchw(tape[i].a), cg['layers'][i - 1][3]))
x = torch.from_numpy(fake_ref).cuda()
logit_ref, cd = oD.forward(fake_ref, fl)
de = D._engine_for()
lg, (dt, _) = de.forward(D, x, torch.from_numpy(fl).cuda(), save=True)
for i in range(5):
    y = cd['y'][i]
    cond = np.abs(y.mean(axis=(0, 2, 3))) / y.std(axis=(0, 2, 3))
    print('D layer', i, 'y relL2', rel(nchw(dt[i].y), y if i else cd['a'][0]), 'a relL2', rel(nchw(dt[i].a), cd['a'][i]), ' |mean|/std of y per channel: max %.1f median %.1f' % (cond.max(), np.median(cond)))
print('logits', lg.cpu().numpy(), logit_ref)
