"""Drop-in parity on the GPU: the reference's own training-loop op sequence (train_gan.py:121-150) driven over
OUR Generator / Discriminator (CUDA kernels behind the C ABI) against fixtures produced by the reference itself."""
import json
import os

import numpy as np
import pytest
import torch

import dcgan_oracle as orc
import gan_enhanced_pneumonia_classifier_b200 as pkg
from conftest import GOLDEN
from parity_utils import close, grad_close, synthetic_noise, synthetic_real, weights_close

pytestmark = pytest.mark.gpu


def build(m, dtype):
    rng = np.random.RandomState(m['seed'])
    sdG = orc.init_state(orc.generator_plan(m['nz'], m['nc'], m['fm']), True, rng)
    sdD = orc.init_state(orc.discriminator_plan(m['nc'], m['fm']), False, rng)
    G, D = pkg.Generator(m['nz'], m['nc'], m['fm']), pkg.Discriminator(m['nc'], m['fm'])
    G.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sdG.items()})
    D.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sdD.items()})
    G, D = G.cuda(), D.cuda()
    G.compute_dtype = D.compute_dtype = dtype
    return G, D


def reference_loop_step(G, D, optG, optD, real, noise):
    """train_gan.py:121-150 verbatim in structure (module calls, BCELoss, backward, optimizer steps)."""
    crit = torch.nn.BCELoss()
    b = real.size(0)
    D.zero_grad()
    label = torch.full((b,), 0.9, dtype=torch.float, device=real.device)
    out_real = D(real).view(-1)
    errD_real = crit(out_real, label)
    errD_real.backward()
    fake = G(noise)
    label.fill_(0.0)
    out_fake = D(fake.detach()).view(-1)
    errD_fake = crit(out_fake, label)
    errD_fake.backward()
    errD = errD_real + errD_fake
    gD = {k: p.grad.detach().cpu().numpy().copy() for k, p in D.named_parameters()}
    optD.step()
    G.zero_grad()
    label.fill_(0.9)
    out2 = D(fake).view(-1)
    errG = crit(out2, label)
    errG.backward()
    gG = {k: p.grad.detach().cpu().numpy().copy() for k, p in G.named_parameters()}
    optG.step()
    return dict(errG=errG.item(), errD=errD.item(), D_x=out_real.mean().item(), D_G_z1=out_fake.mean().item(),
                D_G_z2=out2.mean().item(), fake=fake.detach().cpu().numpy(), p_real=out_real.detach().cpu().numpy(),
                p_fake=out_fake.detach().cpu().numpy(), p_fake_for_G=out2.detach().cpu().numpy(), grads_D=gD, grads_G=gG)


@pytest.mark.parametrize('name', ['step_small_nc1.npz', 'step_small_nc3.npz'])
def test_fp32_mode_matches_reference_fixture(name):
    g = np.load(os.path.join(GOLDEN, name))
    m = json.loads(str(g['meta']))
    G, D = build(m, torch.float32)
    optD = torch.optim.Adam(D.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    optG = torch.optim.Adam(G.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    real = torch.from_numpy(synthetic_real(m['real_seed'], m['batch'], m['nc'])).cuda()
    noises = synthetic_noise(m['noise_seed'], m['batch'] * m['iters'], m['nz']).reshape(m['iters'], m['batch'], m['nz'], 1, 1)
    for it in range(m['iters']):
        r = reference_loop_step(G, D, optG, optD, real, torch.from_numpy(noises[it]).cuda())
        tol = dict(rtol=1e-4, atol=1e-6) if it == 0 else dict(rtol=2e-3, atol=1e-5)
        for k in ('errG', 'errD', 'D_x', 'D_G_z1', 'D_G_z2', 'p_real', 'p_fake', 'p_fake_for_G'):
            close(r[k], g[f'it{it}.{k}'], what=f'it{it}.{k}', **tol)
        close(r['fake'][:, :, ::3, ::3], g[f'it{it}.fake'], rtol=tol['rtol'], atol=1e-5 if it == 0 else 1e-3, what='fake')
        if it == 0:
            for net in ('grads_D', 'grads_G'):
                for k, v in r[net].items():
                    grad_close(v, g[f'it0.{net}.{k}'], f'{net}.{k}')
    for tag, net in (('G', G), ('D', D)):
        for k, v in net.state_dict().items():
            ref = g[f'final.{tag}.{k}']
            v = v.cpu().numpy()
            if k.endswith('num_batches_tracked'):
                assert int(v) == int(ref), k
            elif 'running' in k:
                close(v, ref, rtol=1e-3, atol=1e-5, what=k)
            else:
                weights_close(v, ref, what=f'final.{tag}.{k}', steps=m['iters'], rtol=1e-3, atol=1e-5, frac=0.97)


def test_bf16_mode_within_tolerance_of_fp32_reference():
    g = np.load(os.path.join(GOLDEN, 'step_small_nc1.npz'))
    m = json.loads(str(g['meta']))
    G, D = build(m, torch.bfloat16)
    optD = torch.optim.Adam(D.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    optG = torch.optim.Adam(G.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    real = torch.from_numpy(synthetic_real(m['real_seed'], m['batch'], m['nc'])).cuda()
    noises = synthetic_noise(m['noise_seed'], m['batch'] * m['iters'], m['nz']).reshape(m['iters'], m['batch'], m['nz'], 1, 1)
    r = reference_loop_step(G, D, optG, optD, real, torch.from_numpy(noises[0]).cuda())
    # north star: bf16 rtol 2e-2 against the fp32 reference (probabilities, losses, generator output)
    for k in ('errG', 'errD', 'D_x', 'D_G_z1', 'D_G_z2', 'p_real', 'p_fake', 'p_fake_for_G'):
        close(r[k], g[f'it0.{k}'], rtol=2e-2, atol=2e-3, what=k)
    close(r['fake'][:, :, ::3, ::3], g['it0.fake'], rtol=2e-2, atol=2e-2, what='fake')
    # Gradients of the bf16 path are compared where BatchNorm is well conditioned: full width, batch 32, against the reference fixture with
    # bounds calibrated by the bf16-storage oracle (tests/test_gpu_fullsize.py).  At this fixture's batch of 3 a BatchNorm layer normalises over
    # as few as 147 samples and bf16 storage noise alone moves the gradients by 0.2 relative L2 (SIMT and tcgen05 kernels alike), so only their
    # wiring is checked here: direction and magnitude of every gradient tensor.
    for net in ('grads_D', 'grads_G'):
        for k, v in r[net].items():
            ref = g[f'it0.{net}.{k}'].astype(np.float64).reshape(-1)
            a = v.astype(np.float64).reshape(-1)
            cos = float(a @ ref / (np.linalg.norm(a) * np.linalg.norm(ref)))
            rel = float(np.linalg.norm(a - ref) / np.linalg.norm(ref))
            assert cos > 0.9 and rel < 0.45, f'{net}.{k}: cosine {cos:.3f} relL2 {rel:.3f}'
    for tag, net in (('G', G), ('D', D)):
        for k, v in net.state_dict().items():
            if k.endswith('weight') or k.endswith('bias'):
                # one Adam step moves every weight by ~lr; bf16 gradient noise may flip near-zero ones (2*lr)
                assert np.abs(v.cpu().numpy() - g[f'final.{tag}.{k}']).max() <= 2.05 * m['lr'] * m['iters'] + 1e-6, k


def test_full_size_fp32_checksums():
    g = np.load(os.path.join(GOLDEN, 'step_full_nc1.npz'))
    m = json.loads(str(g['meta']))
    G, D = build(m, torch.float32)
    optD = torch.optim.Adam(D.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    optG = torch.optim.Adam(G.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    real = torch.from_numpy(synthetic_real(m['real_seed'], m['batch'], m['nc'])).cuda()
    noise = torch.from_numpy(synthetic_noise(m['noise_seed'], m['batch'], m['nz'])).cuda()
    r = reference_loop_step(G, D, optG, optD, real, noise)
    for k in ('errG', 'errD', 'D_x', 'D_G_z1', 'D_G_z2', 'p_real', 'p_fake', 'p_fake_for_G'):
        close(r[k], g[f'it0.{k}'], rtol=1e-4, atol=1e-6, what=k)
    close(r['fake'][:, :, ::7, ::7], g['it0.fake_sample'], atol=1e-5, what='fake_sample')
    for net in ('grads_D', 'grads_G'):
        for k, v in r[net].items():
            l2 = np.sqrt((v.astype(np.float64) ** 2).sum())
            close(l2, g[f'it0.{net}.{k}.l2'], rtol=1e-3, what=f'{net}.{k}.l2')
            grad_close(v.reshape(-1)[::max(1, v.size // 64)][:64], g[f'it0.{net}.{k}.sample'], f'{net}.{k}', bulk=1e-4)


def test_state_dict_roundtrip_eval_forward_and_vis_side_effects():
    """generate_synthetic.py's path: load_state_dict -> eval() -> forward under no_grad; plus SURVEY fact X5:
    a train-mode forward under no_grad still updates the BatchNorm buffers."""
    m = dict(seed=7, nz=16, nc=3, fm=8)
    G, _ = build(m, torch.float32)
    z = torch.from_numpy(synthetic_noise(3, 4, 16)).cuda()
    sd_np = {k: v.cpu().numpy().copy() for k, v in G.state_dict().items()}
    oracle = orc.GeneratorOracle(16, 3, 8, {k: v.copy() for k, v in sd_np.items()})
    with torch.no_grad():
        out_train = G(z)                                   # train mode: buffers move
    ref_train, _ = oracle.forward(z.cpu().numpy(), train=True)
    close(out_train.cpu().numpy(), ref_train, rtol=1e-4, atol=1e-5, what='train-mode no_grad forward')
    assert int(G.main[1].num_batches_tracked) == 1
    close(G.main[1].running_mean.cpu().numpy(), oracle.sd['main.1.running_mean'], rtol=1e-4, atol=1e-6, what='running_mean')
    close(G.main[4].running_var.cpu().numpy(), oracle.sd['main.4.running_var'], rtol=1e-4, atol=1e-6, what='running_var')
    G.eval()
    with torch.no_grad():
        out_eval = G(z)
    ref_eval, _ = oracle.forward(z.cpu().numpy(), train=False)
    close(out_eval.cpu().numpy(), ref_eval, rtol=1e-4, atol=1e-5, what='eval forward')
    assert int(G.main[1].num_batches_tracked) == 1        # eval does not touch the buffers


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_graph_replay_matches_kernel_by_kernel(dtype):
    """DCGANTrainer.step replayed from its CUDA graph (second call onwards) against the same trainer launching kernel by kernel.

    Exact, every run: launch counts, the device-side Adam step counts and every BatchNorm `num_batches_tracked` (they live on the
    device precisely so that replay is exact).  Tight: the first two iterations (history scalars, and the complete state after
    them) -- the second call is already a replay.  Afterwards only the sign-flip envelope is asserted: the step is not bitwise
    reproducible from run to run (fp32 shared-memory / split-K accumulation order), and a GAN iteration amplifies that noise
    discretely -- one ReLU derivative within 1e-7 of zero taking the other branch puts a run on an alternative, equally valid
    trajectory ~1e-4 away (tools/diag_flake3.py measured 4 such runs in 150, kernel by kernel against itself).  A retry loop
    used to paper over that; the assertions below hold on either trajectory instead."""
    from gan_enhanced_pneumonia_classifier_b200.trainer import DCGANTrainer
    m = dict(seed=3, nz=16, nc=1, fm=8)
    real = torch.from_numpy(synthetic_real(5, 4, 1)).cuda()
    noises = [torch.from_numpy(synthetic_noise(10 + i, 4, 16)).cuda() for i in range(5)]
    tol = dict(rtol=1e-4, atol=1e-5) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-3)
    hist, nets, early = {}, {}, {}
    for mode in (False, True):
        G, D = build(m, dtype)
        tr = DCGANTrainer(G, D, dtype=dtype, use_graph=mode)
        rows = []
        for i, z in enumerate(noises):
            rows.append(tr.step(real, z))
            if i == 1:
                early[mode] = {f'{tag}.{k}': v.detach().clone() for tag, net in (('G', G), ('D', D)) for k, v in net.state_dict().items()}
        hist[mode] = torch.stack(rows).cpu().numpy()
        nets[mode] = (G, D, tr)
    assert len(nets[True][2]._graphs) == 1 and nets[True][2].launches == nets[False][2].launches
    assert int(nets[True][2].arenaD.step_dev) == 5 and int(nets[True][2].arenaG.step_dev) == 5
    for a, b in zip(nets[True][:2], nets[False][:2]):
        sa, sb = a.state_dict(), b.state_dict()
        for k in sa:
            if k.endswith('num_batches_tracked'):
                assert int(sa[k]) == int(sb[k]), k
    # iterations 1-2 (the second one replayed): history and the whole state, tight
    close(hist[True][:2], hist[False][:2], what='history scalars, graph vs eager (iterations 1-2)', **tol)
    for k, v in early[True].items():
        if k.endswith('num_batches_tracked'):
            assert int(v) == int(early[False][k]), k
        elif 'running' in k:
            close(v.cpu().numpy(), early[False][k].cpu().numpy(), what=f'after 2 iterations: {k}', **tol)
        else:
            weights_close(v.cpu().numpy(), early[False][k].cpu().numpy(), what=f'after 2 iterations: {k}', steps=2, rtol=tol['rtol'], atol=max(tol['atol'], 2e-6),
                          frac=0.98 if dtype == torch.float32 else 0.9)
    # iterations 3-5: finite, and every weight inside the envelope two valid trajectories can be apart
    assert np.isfinite(hist[True]).all() and np.isfinite(hist[False]).all()
    close(hist[True], hist[False], what='history scalars, graph vs eager (all iterations)', rtol=5e-2, atol=5e-3)
    for a, b in zip(nets[True][:2], nets[False][:2]):
        sa, sb = a.state_dict(), b.state_dict()
        for k in sa:
            if not k.endswith('num_batches_tracked') and 'running' not in k:
                assert float((sa[k] - sb[k]).abs().max()) <= 2.1 * 2e-4 * 5, k


def test_fused_trainer_matches_oracle_over_three_iterations():
    """DCGANTrainer.step (fused BCE, skipped dead D-wgrad, flat arenas, fused Adam) against the oracle's train_iteration:
    history scalars and every post-step tensor of both state_dicts, fp32 mode, three iterations."""
    from gan_enhanced_pneumonia_classifier_b200.trainer import DCGANTrainer
    m = dict(seed=21, nz=16, nc=1, fm=8)
    G, D = build(m, torch.float32)
    tr = DCGANTrainer(G, D, dtype=torch.float32, use_graph=False)
    rng = np.random.RandomState(m['seed'])
    sdG = orc.init_state(orc.generator_plan(16, 1, 8), True, rng)
    sdD = orc.init_state(orc.discriminator_plan(1, 8), False, rng)
    oG, oD = orc.GeneratorOracle(16, 1, 8, sdG), orc.DiscriminatorOracle(1, 8, sdD)
    aG, aD = orc.AdamOracle(orc.param_keys(oG.plan), 2e-4, 0.5), orc.AdamOracle(orc.param_keys(oD.plan), 2e-4, 0.5)
    for it in range(3):
        real, noise = synthetic_real(300 + it, 4, 1), synthetic_noise(400 + it, 4, 16)
        got = tr.step(torch.from_numpy(real).cuda(), torch.from_numpy(noise).cuda()).cpu().numpy()
        r = orc.train_iteration(oG, oD, aG, aD, real, noise)
        want = np.array([r[k] for k in ('errD', 'errG', 'D_x', 'D_G_z1', 'D_G_z2')])
        close(got, want, what=f'history it{it}', **(dict(rtol=1e-4, atol=1e-6) if it == 0 else dict(rtol=2e-3, atol=1e-5)))
    for tag, net, o in (('G', G, oG), ('D', D, oD)):
        for k, v in net.state_dict().items():
            if k.endswith('num_batches_tracked'):
                assert int(v) == int(o.sd[k]), k
            elif 'running' in k:
                # three iterations of trajectory noise (fp32 atomics reorder sums from run to run): measured drift up to 2e-5
                close(v.cpu().numpy(), o.sd[k], rtol=2e-3, atol=1e-4, what=k)
            else:
                if k.endswith('.bias'):
                    # BatchNorm biases (8-64 entries, each the sum of three ~lr-sized normalised Adam updates of a near-cancelling
                    # gradient) are the noisiest numbers of the step: fp32 atomics alone move single entries by ~2e-5 from run to
                    # run, so a fraction-of-entries criterion is meaningless on them; every entry must stay within 0.25 lr
                    assert np.abs(v.cpu().numpy() - o.sd[k]).max() < 5e-5, f'{tag}.{k}'
                else:
                    weights_close(v.cpu().numpy(), o.sd[k], what=f'{tag}.{k}', steps=3, rtol=1e-3, atol=1e-5, frac=0.97)


def test_cli_train_then_sample_on_gpu(tmp_path):
    """train_gan.py CLI (fused trainer, bf16, graph replay) -> generator_final.pth -> generate_synthetic.py on the GPU; the PNG
    pixels must equal the oracle's eval-mode forward of the saved checkpoint on the same noise (reference generate_synthetic.py:34-54)."""
    from PIL import Image
    from gan_enhanced_pneumonia_classifier_b200 import generate_synthetic as gs
    from gan_enhanced_pneumonia_classifier_b200 import train_gan as tg
    d = str(tmp_path)
    argv = ['--synthetic', '8', '--batch-size', '4', '--epochs', '2', '--latent-dim', '16', '--feature-maps-g', '8', '--feature-maps-d', '8',
            '--num-channels', '3', '--vis-batch-size', '4', '--model-dir', d + '/models', '--output-dir', d + '/results',
            '--results-dir', d + '/results/metrics', '--figures-dir', d + '/results/figures', '--seed', '0', '--dtype', 'fp32']
    hist = tg.main(tg.build_parser().parse_args(argv))
    assert len(hist['G_losses_iter']) == 4 and all(np.isfinite(hist['D_losses_iter']))
    ckpt = d + '/models/gan/generator_final.pth'
    torch.manual_seed(5)
    gs.generate_images(ckpt, d + '/synthetic', 3, 16, 8, 2, torch.device('cuda'))
    torch.manual_seed(5)
    z = torch.cat([torch.randn(2, 16, 1, 1, device='cuda'), torch.randn(1, 16, 1, 1, device='cuda')]).cpu().numpy()
    sd = {k: v.cpu().numpy() for k, v in torch.load(ckpt).items()}
    ref, _ = orc.GeneratorOracle(16, 3, 8, sd).forward(z, train=False)
    want = np.clip((ref * 0.5 + 0.5) * 255 + 0.5, 0, 255).astype(np.uint8).transpose(0, 2, 3, 1)
    for i in range(3):
        got = np.asarray(Image.open(f'{d}/synthetic/synthetic_{i + 1:05d}.png'))
        assert got.shape == (224, 224, 3)
        assert np.abs(got.astype(int) - want[i].astype(int)).max() <= 1, f'image {i}'       # fp32 path: at most one grey level


@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
def test_cli_trains_from_the_device_resident_cache(tmp_path, dtype):
    """train_gan.py --cache-dataset: batches come from DeviceImageCache (uint8 in HBM, gather / flip / normalise on the GPU)
    instead of a host DataLoader; 10 images at batch 4 = 3 iterations per epoch, the last one ragged."""
    from gan_enhanced_pneumonia_classifier_b200 import train_gan as tg
    d = str(tmp_path)
    argv = ['--synthetic', '10', '--cache-dataset', '--batch-size', '4', '--epochs', '2', '--latent-dim', '16', '--feature-maps-g', '8',
            '--feature-maps-d', '8', '--num-channels', '3', '--vis-batch-size', '4', '--model-dir', d + '/models', '--output-dir', d + '/results',
            '--results-dir', d + '/results/metrics', '--figures-dir', d + '/results/figures', '--seed', '0', '--dtype', dtype]
    hist = tg.main(tg.build_parser().parse_args(argv))
    assert len(hist['G_losses_iter']) == 6 and all(np.isfinite(hist['D_losses_iter'])) and all(np.isfinite(hist['G_losses_iter']))
    assert os.path.exists(d + '/models/gan/generator_final.pth')


def test_packed_weight_cache_follows_the_weights_under_graph_replay():
    """Full-size networks (ngf = ndf = 64: the tcgen05 layers and their bf16 weight repacks are active), batch 2, graph replay.
    The repacks are cached per Adam update in persistent buffers that the captured graph reads across replays; after every step
    each cached repack must equal a fresh repack of the current fp32 master weights bit for bit, and the history must stay finite
    and close to the kernel-by-kernel run."""
    import ctypes as C
    from gan_enhanced_pneumonia_classifier_b200.trainer import DCGANTrainer
    L = pkg._lib
    m = dict(seed=5, nz=100, nc=1, fm=64)
    real = torch.from_numpy(synthetic_real(7, 2, 1)).cuda()
    noises = [torch.from_numpy(synthetic_noise(20 + i, 2, 100)).cuda() for i in range(4)]
    hist = {}
    for mode in (False, True):
        G, D = build(m, torch.bfloat16)
        tr = DCGANTrainer(G, D, dtype=torch.bfloat16, use_graph=mode)
        rows = []
        for z in noises:
            rows.append(tr.step(real, z).cpu().numpy())
            for eng, net in ((tr.engD, D), (tr.engG, G)):
                for i, sp in enumerate(eng.specs):
                    if not eng._tc_layer(i):
                        continue
                    w = net.main[sp.conv_idx].weight
                    fresh = torch.empty(2 * w.numel(), device='cuda', dtype=torch.bfloat16)
                    L.call('b200gan_pack_conv_weight', L.ptr(w), w.shape[0], w.shape[1], 4, 2, L.ptr(fresh), L.stream_ptr())
                    ver, cached = eng._packed[i]
                    if ver == eng.weights_version:          # a cache entry of the current version must be current
                        assert torch.equal(cached, fresh), f'stale repack: graph={mode} layer {i}'
        hist[mode] = np.stack(rows)
        assert np.isfinite(hist[mode]).all()
    close(hist[True][:2], hist[False][:2], rtol=5e-2, atol=5e-2, what='history, graph vs eager, full-size bf16')
