"""Parity of the BENCHMARKED configuration: DCGANTrainer.step in bf16 on the full-width networks (nz=100, ngf=ndf=64: every
tcgen05 kernel, the halo-tile kernels, the image-side and latent kernels and the fused epilogues are active), kernel by kernel
AND replayed from its CUDA graph, against fixtures produced by the reference itself on CPU fp32
(tests/golden/step_full_b32_nc{1,3}.npz, written by oracle/make_golden.py from /root/reference/src/dcgan.py and the op
sequence of train_gan.py:121-150; batch 32, two iterations, nc = 1 (the benchmark workload) and nc = 3 (the CLI default)).

north_star tolerance: bf16 rtol 2e-2 against the fp32 reference for generator outputs, discriminator probabilities, losses;
post-step weights inside the 2*lr-per-step envelope a sign flip of a ~0 gradient can cause.

Gradients (not named by north_star).  With hard ReLU / LeakyReLU branches the distance of ANY bf16-storage implementation from
the fp32 reference is not 2^-9: every pre-activation within rounding distance of zero takes the other branch, which perturbs the
gradients by ~sqrt(fraction of flipped elements) per layer.  The numpy oracle run with bf16 rounding at the points where the CUDA
path stores bf16 (dcgan_oracle.Net(storage=bf16_round), wide accumulation everywhere: the IDEAL bf16 implementation) sits at
relative L2 0.01 (last D layer) .. 0.12 (first D layer) .. 0.2 (G) from the reference at batch 32; even two fp32 implementations
(stock torch vs the float64-accumulating oracle) differ by 1e-3 (D) .. 2e-2 (G) because single branch flips suffice.  Those
per-tensor distances are stored in the fixture (`bf16_model.*`, oracle/make_golden.py) and the CUDA path is held to
GRAD_MODEL_FACTOR x that ideal distance + GRAD_MODEL_SLACK -- i.e. "not measurably worse than perfect bf16 storage".  Kernel-level
parity (one bf16 ulp against the float64 oracle on every layer shape) is pinned in tests/test_gpu_tc.py; a bf16 trajectory
cannot be matched more tightly end to end, because one-ulp rounding differences spread to every element within three layers
(tools/diag_emu_layers.py: 0.06 % of D1's outputs differ, 1 % of D2's, 11 % of D3's, 36 % of D4's).
"""
import json
import os

import numpy as np
import pytest
import torch

import dcgan_oracle as orc
import gan_enhanced_pneumonia_classifier_b200 as pkg
from conftest import GOLDEN
from parity_utils import close, synthetic_noise, synthetic_real

pytestmark = pytest.mark.gpu

GRAD_MODEL_FACTOR, GRAD_MODEL_SLACK = 1.5, 0.02      # bf16: relL2(kernel, reference) <= FACTOR * relL2(ideal bf16 storage, reference) + SLACK
GRAD_NORM_RTOL = 5e-2                                 # every gradient tensor's L2 norm within 5 %
WEIGHT_TIGHT_FRAC = 0.7                               # share of post-step weights within 0.25 lr per step (measured: >= 0.75 G, >= 0.93 D)


def _sample(v, n):
    v = np.asarray(v).reshape(-1)
    return v[::max(1, v.size // n)][:n]


def _build(m, dtype):
    rng = np.random.RandomState(m['seed'])
    sdG = orc.init_state(orc.generator_plan(m['nz'], m['nc'], m['fm']), True, rng)
    sdD = orc.init_state(orc.discriminator_plan(m['nc'], m['fm']), False, rng)
    G, D = pkg.Generator(m['nz'], m['nc'], m['fm']), pkg.Discriminator(m['nc'], m['fm'])
    G.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sdG.items()})
    D.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sdD.items()})
    G, D = G.cuda(), D.cuda()
    G.compute_dtype = D.compute_dtype = dtype
    return G, D


class _Snapshot:
    """Everything DCGANTrainer.step mutates, so that a trainer whose graph is already captured can be put back to the
    fixture's initial state (same device addresses: the captured graph stays valid)."""

    def __init__(self, tr):
        self.tr = tr
        self.p = [a.param.clone() for a in (tr.arenaG, tr.arenaD)]
        self.buf = [(b, b.clone()) for net in (tr.netG, tr.netD) for b in net.buffers()]

    def restore(self, buffers_only=False):
        tr = self.tr
        with torch.no_grad():
            for b, v in self.buf:
                b.copy_(v)
            if buffers_only:
                return
            for a, p in zip((tr.arenaG, tr.arenaD), self.p):
                a.param.copy_(p)
                a.exp_avg.zero_()
                a.exp_avg_sq.zero_()
                a.grad.zero_()
                a.step_dev.zero_()
                a.step = 0
        tr.refresh_packed_weights()


@pytest.mark.parametrize('mode', ['eager', 'graph'])
@pytest.mark.parametrize('nc', [1, 3])
def test_bf16_full_width_trainer_matches_reference_fixture(nc, mode, capsys):
    _run(nc, mode, torch.bfloat16, capsys, rtol=2e-2, atol=2e-3, grad_rel=None, grad_norm=GRAD_NORM_RTOL, tight=WEIGHT_TIGHT_FRAC)


def test_fp32_full_width_trainer_matches_reference_fixture(capsys):
    """The same comparison in the fp32 parity mode (SIMT kernels): separates wiring from bf16 precision.  north_star: fp32 rtol 1e-4
    on outputs / losses.  Gradient bounds: stock torch itself sits 1e-3 (D) / 2.5e-2 (G) from the float64-accumulated truth on this
    fixture (single LeakyReLU / ReLU branch flips, then the +-lr Adam step of D before the G step), so that is the floor."""
    _run(1, 'eager', torch.float32, capsys, rtol=1e-4, atol=1e-6, grad_rel=(5e-3, 6e-2), grad_norm=5e-3, tight=0.9, it1_rtol=2e-2)


def _run(nc, mode, dtype, capsys, rtol, atol, grad_rel, grad_norm, tight, it1_rtol=None):
    from gan_enhanced_pneumonia_classifier_b200.trainer import DCGANTrainer
    g = np.load(os.path.join(GOLDEN, f'step_full_b32_nc{nc}.npz'))
    m = json.loads(str(g['meta']))
    G, D = _build(m, dtype)
    tr = DCGANTrainer(G, D, lr=m['lr'], beta1=m['beta1'], dtype=dtype, use_graph=(mode == 'graph'))
    snap = _Snapshot(tr)
    real = torch.from_numpy(synthetic_real(m['real_seed'], m['batch'], nc)).cuda()
    noises = synthetic_noise(m['noise_seed'], m['batch'] * m['iters'], m['nz']).reshape(m['iters'], m['batch'], m['nz'], 1, 1)
    noises = [torch.from_numpy(z).cuda() for z in noises]
    if mode == 'graph':
        # first call: kernel by kernel, second: capture + replay; then rewind to the initial state so that the compared
        # iterations are pure graph replays
        tr.step(real, noises[0])
        tr.step(real, noises[1])
        assert len(tr._graphs) == 1
        snap.restore()
    # ---- generator output of the first iteration (module-level forward, train mode; its BatchNorm side effects are undone)
    with torch.no_grad():
        fake = G(noises[0]).cpu().numpy()
    snap.restore(buffers_only=True)
    fs = m['fake_stride']
    close(fake[:, :, ::fs, ::fs], g['it0.fake_sample'], rtol=rtol, atol=max(atol, 1e-5) * 10, what='generator output')
    report, fails = [], []
    keysD, keysG = orc.param_keys(orc.discriminator_plan(nc, m['fm'])), orc.param_keys(orc.generator_plan(m['nz'], nc, m['fm']))
    for it in range(m['iters']):
        got = tr.step(real, noises[it]).cpu().numpy()
        want = np.array([g[f'it{it}.{k}'] for k in ('errD', 'errG', 'D_x', 'D_G_z1', 'D_G_z2')])
        report.append(f'it{it} history got {got} want {want}')
        # north star: bf16 rtol 2e-2 against the fp32 reference (losses, mean probabilities)
        rt = rtol if it == 0 or it1_rtol is None else it1_rtol          # later iterations carry the first Adam step's sign-flip noise
        if not (np.abs(got - want) <= atol + rt * np.abs(want)).all():
            fails.append(f'history scalars it{it}: got {got} want {want}')
        if mode == 'graph':
            assert len(tr._graphs) == 1, 'the compared iterations must be graph replays'
        # ---- every gradient tensor of this iteration, relative L2 on the fixture's strided sample + the tensor's norm
        worst = 0.0
        for arena, keys, tag in ((tr.arenaD, keysD, 'grads_D'), (tr.arenaG, keysG, 'grads_G')):
            for k, (lo, hi) in zip(keys, arena.slices):
                v = arena.grad[lo:hi].float().cpu().numpy().astype(np.float64)
                ref = g[f'it{it}.{tag}.{k}.sample'].astype(np.float64)
                a = _sample(v, m['sample'])
                rel = float(np.linalg.norm(a - ref) / max(np.linalg.norm(ref), 1e-30))
                nrm = float(np.sqrt((v ** 2).sum()) / g[f'it{it}.{tag}.{k}.l2'])
                report.append(f'it{it} {tag}.{k:16s} relL2 {rel:.3e}  norm ratio {nrm:.4f}')
                worst = max(worst, rel)
                if it > 0:
                    continue                 # later iterations: trajectories of a GAN step are chaotic at the gradient level; history is compared
                if grad_rel is None:
                    ideal = float(g[f'bf16_model.it0.{tag}.{k}.rel_l2'])
                    bound = GRAD_MODEL_FACTOR * ideal + GRAD_MODEL_SLACK
                    report[-1] += f'  (ideal bf16 storage {ideal:.3e}, bound {bound:.3e})'
                else:
                    bound = grad_rel[0] if tag == 'grads_D' else grad_rel[1]
                if rel >= bound:
                    fails.append(f'it{it} {tag}.{k}: relative L2 {rel:.3e} (bound {bound:.3e})')
                if abs(nrm - 1) >= max(grad_norm, bound):
                    fails.append(f'it{it} {tag}.{k}: norm ratio {nrm:.4f}')
        report.append(f'it{it} worst gradient relL2 {worst:.3e}')
    # ---- post-step state: weights inside the sign-flip envelope and mostly tight; BatchNorm buffers; counters exact
    for tag, net in (('G', G), ('D', D)):
        for k, v in net.state_dict().items():
            v = v.cpu().numpy()
            if k.endswith('num_batches_tracked'):
                assert int(v) == int(g[f'final.{tag}.{k}']), k
                continue
            ref = g[f'final.{tag}.{k}'] if f'final.{tag}.{k}' in g.files else g[f'final.{tag}.{k}.sample']
            a = v if ref.shape == v.shape else _sample(v, m['wsample'])
            d = np.abs(a.astype(np.float64) - ref)
            if 'running' in k:
                # two iterations of trajectory noise on top of the storage precision: relative to the tensor's scale
                if d.max() > (2e-2 if dtype == torch.bfloat16 else 5e-3) * np.abs(ref).max():
                    fails.append(f'{tag}.{k}: max diff {d.max():.3e} (ref max {np.abs(ref).max():.3e})')
            else:
                frac = float((d <= 0.25 * m['lr'] * m['iters']).mean())
                report.append(f'final {tag}.{k:16s} max diff {d.max():.2e}  within 0.25*lr*iters: {frac:.4f}')
                if d.max() > 2.1 * m['lr'] * m['iters']:      # 2 x (lr + 1.054 lr): Adam's second bias-corrected step can exceed lr by 5 %
                    fails.append(f'{tag}.{k}: max diff {d.max():.3e} exceeds the sign-flip envelope')
                if frac < tight:
                    fails.append(f'{tag}.{k}: only {frac:.3f} of the entries within 0.25 lr per step')
    with capsys.disabled():
        print(f'\n[full-width {dtype} parity nc={nc} {mode}]\n  ' + '\n  '.join(report))
    assert not fails, '\n'.join(fails)
