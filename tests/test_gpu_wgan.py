"""WGAN-GP on the GPU: the drop-in `wggan` modules, `gradient_penalty` with its explicitly launched
double backward, and the fused `WGANGPTrainer`, against the numpy oracle (oracle/wgan_oracle.py) and fixtures produced by the
reference itself (tests/golden/wgan_small_nc{1,3}.npz: /root/reference/src/wggan.py + torch.autograd's double backward)."""
import json
import os

import numpy as np
import pytest
import torch

import dcgan_oracle as orc
import wgan_oracle as wo
import gan_enhanced_pneumonia_classifier_b200 as pkg
from conftest import GOLDEN
from gan_enhanced_pneumonia_classifier_b200 import wggan
from gan_enhanced_pneumonia_classifier_b200.wgan_trainer import WGANGPTrainer
from parity_utils import close, grad_close, synthetic_noise, synthetic_real, weights_close

pytestmark = pytest.mark.gpu


def build(m, dtype):
    rng = np.random.RandomState(m['seed'])
    sdG = orc.init_state(wo.wgan_generator_plan(m['nz'], m['nc'], m['fm']), True, rng)
    sdD = orc.init_state(wo.critic_plan(m['nc'], m['fm']), False, rng)
    G, D = wggan.Generator(m['nz'], m['nc'], m['fm']), wggan.Discriminator(m['nc'], m['fm'])
    G.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sdG.items()})
    D.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sdD.items()})
    G, D = G.cuda(), D.cuda()
    G.compute_dtype = D.compute_dtype = dtype
    return G, D, sdG, sdD


@pytest.mark.parametrize('nc,batch,seed', [(1, 3, 32), (3, 3, 34), (3, 4, 35), (1, 5, 36)])
def test_gradient_penalty_and_its_double_backward_match_the_oracle(nc, batch, seed):
    """wggan.gradient_penalty (drop-in signature) on CUDA: value and d gp / d (every critic parameter) through `.backward()`."""
    m = dict(seed=seed, nz=16, nc=nc, fm=8)
    G, D, sdG, sdD = build(m, torch.float32)
    real, fake = synthetic_real(seed + 10, batch, nc), synthetic_real(seed + 11, batch, nc) * 0.5
    alpha = np.random.RandomState(seed + 12).rand(batch, 1, 1, 1).astype(np.float32)
    orig = torch.rand
    torch.rand = lambda *a, **k: torch.from_numpy(alpha).to(k.get('device', 'cpu'))      # the draw of wggan.py:76
    try:
        gp = wggan.gradient_penalty(D, torch.from_numpy(real).cuda(), torch.from_numpy(fake).cuda(), torch.device('cuda'), lambda_gp=10.)
    finally:
        torch.rand = orig
    gp.backward()
    oD = wo.Critic(nc, m['fm'], {k: v.copy() for k, v in sdD.items()})
    xhat = alpha * real + (1 - alpha) * fake
    gp_ref, grads = oD.gradient_penalty(xhat.astype(np.float32), 10.)
    close(float(gp.detach()), gp_ref, rtol=1e-4, atol=1e-6, what='gradient penalty')
    fails = []
    for k, p in D.named_parameters():
        try:
            grad_close(p.grad.cpu().numpy(), grads[k], f'd gp / d {k}', bulk=2e-4)
        except AssertionError as e:
            fails.append(str(e))
    assert not fails, '\n'.join(fails)
    for k in oD.sd:                                        # one train-mode forward of the critic: its BatchNorm buffers moved once
        if 'running' in k:
            close(D.state_dict()[k].cpu().numpy(), oD.sd[k], rtol=1e-4, atol=1e-6, what=k)
        elif k.endswith('num_batches_tracked'):
            assert int(D.state_dict()[k]) == 1


@pytest.mark.parametrize('name', ['wgan_small_nc1.npz', 'wgan_small_nc3.npz'])
def test_fused_trainer_matches_reference_fixture(name):
    g = np.load(os.path.join(GOLDEN, name))
    m = json.loads(str(g['meta']))
    G, D, _, _ = build(m, torch.float32)
    tr = WGANGPTrainer(G, D, lr=m['lr'], beta1=m['beta1'], beta2=m['beta2'], lambda_gp=m['lambda_gp'], critic_iters=m['critic_iters'], dtype=torch.float32)
    real = torch.from_numpy(synthetic_real(m['real_seed'], m['batch'], m['nc'])).cuda()
    noises = synthetic_noise(m['noise_seed'], m['batch'] * (m['critic_iters'] + 1), m['nz']).reshape(m['critic_iters'] + 1, m['batch'], m['nz'], 1, 1)
    keysD = orc.param_keys(wo.critic_plan(m['nc'], m['fm']))
    for it in range(m['critic_iters']):
        out = tr.critic_step(real, torch.from_numpy(noises[it]).cuda(), torch.from_numpy(g[f'c{it}.alpha']).cuda()).cpu().numpy()
        tol = dict(rtol=1e-4, atol=1e-6) if it == 0 else dict(rtol=2e-3, atol=1e-5)
        close(out[0], g[f'c{it}.d_loss'], what=f'c{it}.d_loss', **tol)
        close(out[1], g[f'c{it}.gp'], what=f'c{it}.gp', **tol)
        if it == 0:
            for k, (lo, hi) in zip(keysD, tr.arenaD.slices):
                grad_close(tr.arenaD.grad[lo:hi].cpu().numpy().reshape(g[f'c0.grads_D.{k}'].shape), g[f'c0.grads_D.{k}'], f'critic gradient {k}', bulk=2e-4)
    g_loss = tr.generator_step(torch.from_numpy(noises[m['critic_iters']]).cuda())
    close(float(g_loss), g['g_loss'], rtol=2e-3, atol=1e-5, what='g_loss')
    for tag, net in (('G', G), ('D', D)):
        for k, v in net.state_dict().items():
            ref, v = g[f'final.{tag}.{k}'], v.cpu().numpy()
            if k.endswith('num_batches_tracked'):
                assert int(v) == int(ref), k
            elif 'running' in k:
                close(v, ref, rtol=1e-3, atol=1e-5, what=k)
            else:
                weights_close(v, ref, what=f'final.{tag}.{k}', steps=m['critic_iters'] if tag == 'D' else 1, rtol=1e-3, atol=1e-5, frac=0.97)


def test_reference_loop_runs_unchanged_over_the_drop_in_modules():
    """The reference's own op sequence (train_wggan.py:70-92: module calls, `gradient_penalty`, `d_loss.backward()`, torch.optim.Adam)
    over OUR modules on CUDA, against the fixture."""
    g = np.load(os.path.join(GOLDEN, 'wgan_small_nc1.npz'))
    m = json.loads(str(g['meta']))
    G, D, _, _ = build(m, torch.float32)
    optG = torch.optim.Adam(G.parameters(), lr=m['lr'], betas=(m['beta1'], m['beta2']))
    optD = torch.optim.Adam(D.parameters(), lr=m['lr'], betas=(m['beta1'], m['beta2']))
    real = torch.from_numpy(synthetic_real(m['real_seed'], m['batch'], m['nc'])).cuda()
    noises = synthetic_noise(m['noise_seed'], m['batch'] * (m['critic_iters'] + 1), m['nz']).reshape(m['critic_iters'] + 1, m['batch'], m['nz'], 1, 1)
    orig = torch.rand
    for it in range(m['critic_iters']):
        D.zero_grad()
        d_real_loss = -D(real).mean()
        fake = G(torch.from_numpy(noises[it]).cuda())
        d_fake_loss = D(fake.detach()).mean()
        torch.rand = lambda *a, **k: torch.from_numpy(g[f'c{it}.alpha']).to(k.get('device', 'cpu'))
        try:
            gp = wggan.gradient_penalty(D, real.data, fake.data, torch.device('cuda'), lambda_gp=m['lambda_gp'])
        finally:
            torch.rand = orig
        d_loss = d_real_loss + d_fake_loss + gp
        d_loss.backward()
        tol = dict(rtol=1e-4, atol=1e-6) if it == 0 else dict(rtol=2e-3, atol=1e-5)
        close(d_loss.item(), g[f'c{it}.d_loss'], what=f'c{it}.d_loss', **tol)
        if it == 0:
            for k, p in D.named_parameters():
                grad_close(p.grad.cpu().numpy(), g[f'c0.grads_D.{k}'], f'critic gradient {k}', bulk=2e-4)
        optD.step()
    G.zero_grad()
    g_loss = -D(G(torch.from_numpy(noises[m['critic_iters']]).cuda())).mean()
    g_loss.backward()
    close(g_loss.item(), g['g_loss'], rtol=2e-3, atol=1e-5, what='g_loss')
    for k, p in G.named_parameters():
        grad_close(p.grad.cpu().numpy(), g[f'grads_G.{k}'], f'generator gradient {k}', bulk=5e-4, l2=1e-2, worst=5e-2)


def test_full_width_bf16_critic_step_tracks_the_fp32_path():
    """BASELINE configs[4] widths (nz=100, ngf=ndf=64: the k4 s2 p1 layers run on the tcgen05 kernels in bf16) at batch 4: the bf16
    critic step against the library's own fp32 parity mode -- losses within bf16 tolerance (north_star: rtol 2e-2), gradient penalty
    within 5e-2 (it is a function of a GRADIENT, which bf16 storage perturbs more than a forward value; see test_gpu_fullsize.py)."""
    m = dict(seed=77, nz=100, nc=1, fm=64)
    real = torch.from_numpy(synthetic_real(78, 4, 1)).cuda()
    noise = torch.from_numpy(synthetic_noise(79, 4, 100)).cuda()
    alpha = torch.from_numpy(np.random.RandomState(80).rand(4, 1, 1, 1).astype(np.float32)).cuda()
    out = {}
    for dt in (torch.float32, torch.bfloat16):
        G, D, _, _ = build(m, dt)
        tr = WGANGPTrainer(G, D, dtype=dt)
        out[dt] = tr.critic_step(real, noise, alpha).cpu().numpy()
        assert np.isfinite(out[dt]).all()
    close(out[torch.bfloat16][0] - out[torch.bfloat16][1], out[torch.float32][0] - out[torch.float32][1], rtol=2e-2, atol=2e-3, what='-D(real) + D(fake)')
    close(out[torch.bfloat16][1], out[torch.float32][1], rtol=5e-2, atol=1e-3, what='gradient penalty')


@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
def test_wgan_cli_trains_on_the_gpu(tmp_path, dtype):
    """train_wggan.py CLI on CUDA (fused WGANGPTrainer): 6 images at batch 4 = 2 iterations per epoch (the second one ragged),
    critic_iters 2 -> 4 critic losses + 2 generator losses per epoch, reference artefacts written."""
    from gan_enhanced_pneumonia_classifier_b200 import train_wggan as tw
    d = str(tmp_path)
    argv = ['--synthetic', '6', '--batch-size', '4', '--epochs', '2', '--latent-dim', '16', '--feature-maps-g', '8', '--feature-maps-d', '8',
            '--num-channels', '3', '--vis-batch-size', '4', '--critic-iters', '2', '--model-dir', d + '/models', '--output-dir', d + '/results',
            '--results-dir', d + '/results/metrics', '--figures-dir', d + '/results/figures', '--seed', '0', '--dtype', dtype]
    hist = tw.main(tw.build_parser().parse_args(argv))
    assert len(hist['D_losses']) == 8 and len(hist['G_losses']) == 4
    assert all(np.isfinite(hist['D_losses'])) and all(np.isfinite(hist['G_losses']))
    sd = torch.load(d + '/models/wgan/generator_final.pth')
    assert len(sd) == 31 and sd['main.0.weight'].shape == (16, 128, 7, 7) and int(sd['main.1.num_batches_tracked']) > 0


@pytest.mark.parametrize('n,c', [(3, 64), (5, 512), (2, 32)])
def test_score_map_kernels_match_the_oracle(n, c):
    """The critic's last layer Conv2d(C -> 1, k7 s1 p0) on the 14 x 14 map (wggan.py:63) and its gradients through the dedicated
    one-CTA-per-image kernels (bf16 activations, fp32 weights / scores, as the training step uses them) against the numpy oracle."""
    import ctypes as C
    L = pkg._lib
    rng = np.random.RandomState(c)
    bf = lambda a: torch.from_numpy(a).to(torch.bfloat16).float().numpy()
    x = bf(rng.randn(n, c, 14, 14).astype(np.float32))
    w = (rng.randn(1, c, 7, 7) * 0.05).astype(np.float32)
    dy = rng.randn(n, 1, 8, 8).astype(np.float32)
    xt = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 3, 1))).cuda().to(torch.bfloat16)
    wt, dyt = torch.from_numpy(w).cuda(), torch.from_numpy(np.ascontiguousarray(dy.transpose(0, 2, 3, 1))).cuda()
    cv, st = L.Conv(7, 1, 0, L.ALGO_AUTO), L.stream_ptr()
    y = torch.full((n, 8, 8, 1), float('nan'), device='cuda')
    L.call('b200gan_conv2d_fprop', C.byref(cv), C.byref(L.view_nhwc(xt)), L.ptr(wt), None, C.byref(L.view_nhwc(y)), None, st)
    close(y.cpu().numpy().transpose(0, 3, 1, 2), orc.conv2d_fprop(x, w, 1, 0), rtol=1e-4, atol=1e-4, what='score map forward')
    dx = torch.full((n, 14, 14, c), float('nan'), device='cuda', dtype=torch.bfloat16)
    L.call('b200gan_conv2d_dgrad', C.byref(cv), C.byref(L.view_nhwc(dyt)), L.ptr(wt), None, C.byref(L.view_nhwc(dx)), None, st)
    ref = orc.conv2d_dgrad(dy, w, 1, 0, (14, 14))
    close(dx.float().cpu().numpy().transpose(0, 3, 1, 2), ref, rtol=2.0 ** -8 * 1.05, atol=1e-5 * max(1.0, np.abs(ref).max()), what='score map input gradient')
    base = torch.randn((1, c, 7, 7), device='cuda')
    dw = base.clone()
    L.call('b200gan_conv2d_wgrad', C.byref(cv), C.byref(L.view_nhwc(xt)), C.byref(L.view_nhwc(dyt)), L.ptr(dw), None, None, st)
    ref = orc.conv2d_wgrad(x, dy, 7, 1, 0)
    close((dw - base).cpu().numpy(), ref, rtol=1e-4, atol=3e-5 * max(1.0, np.abs(ref).max()), what='score map weight gradient')
