"""Generate the golden fixtures under tests/golden/ from the REFERENCE ITSELF -- test infrastructure only.

Run in the build container (the only place /root/reference exists):

    python oracle/make_golden.py [--ref /root/reference]

It imports the reference's own modules (`src/dcgan.py`, `src/train_gan.py`) and executes them with the
installed torch on CPU fp32, then stores inputs that cannot be regenerated and all outputs as compressed
npz files.  Nothing at test/bench run time reads /root/reference; the fixtures travel instead.

Fixtures
  step_small_nc1.npz / step_small_nc3.npz
      reference `dcgan.Generator/Discriminator` (nz=16, ngf=ndf=8) driven by the op sequence of
      `train_gan.py:121-150` with explicit noise, 2 iterations, batch 3.  Initial weights come from
      `dcgan_oracle.init_state(RandomState(seed))`, so tests regenerate them.
  main_small_nc3.npz
      the UNMODIFIED `train_gan.main(args)` (matplotlib stubbed, `get_dataloaders` patched to a seeded
      in-memory dataset of 10 images -> batches 4,4,2; save_interval=2 so the train-mode visualisation
      forward of train_gan.py:166-169 runs 3 times).  Initial weights (torch RNG), fixed noise and
      per-iteration noise are recorded from the run; history JSON and final state dicts are the outputs.
  wgan_small_nc1.npz / wgan_small_nc3.npz
      the reference's WGAN-GP iteration (src/wggan.py modules and its unmodified `gradient_penalty`, op sequence of
      src/train_wggan.py:70-92): nz=16, ngf=ndf=8, two critic updates + one generator update; recorded noise / alpha, losses,
      gradient penalties, critic scores, every gradient and the final state dicts.
  cgan_step_nc1.npz / cgan_step_nc3.npz
      the reference's conditional-GAN modules (src/cgan.py) through the op sequence of src/train_cgan.py:150-193 minus the VGG16 perceptual term
      (ImageNet weights cannot be downloaded offline): nz=16, nf=8, two iterations; recorded labels / smoothing / noise, logits, sampled fake
      image, every gradient of both networks, history rows, initial and final state dicts.
  cgan_main_nc3.npz
      the UNMODIFIED `train_cgan.main(args)` (matplotlib stubbed, dataset class -> seeded in-memory dataset, PerceptualLoss -> 0), 6 epochs x 2
      batches of 4: every torch.rand / randn / randint draw in order, the optimizers' initial parameters, history JSON, final state dicts.
  step_full_nc1.npz
      full-size nets (nz=100, ngf=ndf=64, nc=1), batch 2, 1 iteration; scalars, D probabilities, a
      strided sample of the fake image and per-tensor checksums of grads / post-step weights.
  step_full_b32_nc1.npz / step_full_b32_nc3.npz
      the configuration bench.py times, at a batch the CPU finishes in seconds: full-size nets, batch 32 (BatchNorm is well
      conditioned: >= 1568 samples per channel), 2 iterations; as above, but every gradient / post-step tensor is kept as its
      L2 norm plus a strided sample of up to 4096 entries (relative-L2 parity of the bf16 tensor-core path is measured on
      those samples), small tensors (<= 4096 entries, incl. every BatchNorm vector) in full.  `bf16_model.*`: the same first
      iteration by the numpy oracle with bf16 STORAGE rounding (dcgan_oracle.Net(storage=bf16_round)): its history scalars and
      the relative L2 distance of each of its gradient tensors from the reference's -- what bf16 storage costs by itself.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import dcgan_oracle as orc  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def load_state(module, sd_np):
    module.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()})


def to_np(sd):
    return {k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def synthetic_real(seed, n, nc, size=224):
    """Uniform[-1,1) images from numpy's RandomState -- regenerated identically by the tests."""
    return (np.random.RandomState(seed).rand(n, nc, size, size).astype(np.float32) * 2 - 1)


def synthetic_noise(seed, n, nz):
    return np.random.RandomState(seed).randn(n, nz, 1, 1).astype(np.float32)


def reference_step(dcgan, netG, netD, optG, optD, real, noise):
    """The op sequence of train_gan.py:121-150, executed with the reference's modules and stock torch."""
    crit = torch.nn.BCELoss()
    b = real.size(0)
    netD.zero_grad()
    label = torch.full((b,), 0.9, dtype=torch.float)
    out_real = netD(real).view(-1)
    errD_real = crit(out_real, label)
    errD_real.backward()
    fake = netG(noise)
    label.fill_(0.0)
    out_fake = netD(fake.detach()).view(-1)
    errD_fake = crit(out_fake, label)
    errD_fake.backward()
    errD = errD_real + errD_fake
    gradsD = {k: p.grad.detach().numpy().copy() for k, p in netD.named_parameters()}
    optD.step()
    netG.zero_grad()
    label.fill_(0.9)
    out2 = netD(fake).view(-1)
    errG = crit(out2, label)
    errG.backward()
    gradsG = {k: p.grad.detach().numpy().copy() for k, p in netG.named_parameters()}
    optG.step()
    return dict(errG=errG.item(), errD=errD.item(), D_x=out_real.mean().item(), D_G_z1=out_fake.mean().item(),
                D_G_z2=out2.mean().item(), fake=fake.detach().numpy().copy(), p_real=out_real.detach().numpy().copy(),
                p_fake=out_fake.detach().numpy().copy(), p_fake_for_G=out2.detach().numpy().copy(),
                grads_D=gradsD, grads_G=gradsG)


def make_step_fixture(dcgan, name, nz, fm, nc, batch, iters, seed, full=False, sample=64, wsample=256, fake_stride=7, bf16_model=False):
    rng = np.random.RandomState(seed)
    sdG = orc.init_state(orc.generator_plan(nz, nc, fm), True, rng)
    sdD = orc.init_state(orc.discriminator_plan(nc, fm), False, rng)
    netG, netD = dcgan.Generator(nz, nc, fm), dcgan.Discriminator(nc, fm)
    load_state(netG, sdG)
    load_state(netD, sdD)
    optD = torch.optim.Adam(netD.parameters(), lr=2e-4, betas=(0.5, 0.999))
    optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.5, 0.999))
    out = dict(meta=json.dumps(dict(nz=nz, fm=fm, nc=nc, batch=batch, iters=iters, seed=seed, lr=2e-4, beta1=0.5,
                                    real_seed=seed + 1, noise_seed=seed + 2, torch=torch.__version__,
                                    sample=sample, wsample=wsample, fake_stride=fake_stride)))
    real = synthetic_real(seed + 1, batch, nc)
    noises = synthetic_noise(seed + 2, batch * iters, nz).reshape(iters, batch, nz, 1, 1)
    for it in range(iters):
        r = reference_step(dcgan, netG, netD, optG, optD, torch.from_numpy(real), torch.from_numpy(noises[it]))
        for k in ('errG', 'errD', 'D_x', 'D_G_z1', 'D_G_z2'):
            out[f'it{it}.{k}'] = np.float64(r[k])
        for k in ('p_real', 'p_fake', 'p_fake_for_G'):
            out[f'it{it}.{k}'] = r[k]
        if full:
            out[f'it{it}.fake_sample'] = r['fake'][:, :, ::fake_stride, ::fake_stride].copy()
            out[f'it{it}.fake_sum'] = np.float64(r['fake'].astype(np.float64).sum())
            out[f'it{it}.fake_sqsum'] = np.float64((r['fake'].astype(np.float64) ** 2).sum())
            for net in ('grads_D', 'grads_G'):
                for k, v in r[net].items():
                    out[f'it{it}.{net}.{k}.l2'] = np.float64(np.sqrt((v.astype(np.float64) ** 2).sum()))
                    out[f'it{it}.{net}.{k}.sample'] = v.reshape(-1)[::max(1, v.size // sample)][:sample].copy()
        else:
            out[f'it{it}.fake'] = r['fake'][:, :, ::3, ::3].copy()     # strided sample keeps the fixture small
            if it == 0:
                for net in ('grads_D', 'grads_G'):
                    for k, v in r[net].items():
                        out[f'it{it}.{net}.{k}'] = v
    if bf16_model:
        # Calibration of the bf16 comparison: how far does IDEAL bf16 storage (the numpy oracle rounding to bf16 wherever the CUDA path
        # stores bf16, wide accumulation everywhere) sit from the reference's fp32 gradients of the first iteration?  With hard ReLU /
        # LeakyReLU branches that distance is not 2^-9: pre-activations within rounding distance of zero change branch.  The CUDA bf16
        # path is held to a small multiple of these per-tensor numbers (tests/test_gpu_fullsize.py).
        rng2 = np.random.RandomState(seed)
        oG = orc.GeneratorOracle(nz, nc, fm, orc.init_state(orc.generator_plan(nz, nc, fm), True, rng2), storage=orc.bf16_round)
        oD = orc.DiscriminatorOracle(nc, fm, orc.init_state(orc.discriminator_plan(nc, fm), False, rng2), storage=orc.bf16_round)
        r = orc.train_iteration(oG, oD, orc.AdamOracle(orc.param_keys(oG.plan), 2e-4, 0.5), orc.AdamOracle(orc.param_keys(oD.plan), 2e-4, 0.5),
                                real, noises[0])
        for k in ('errG', 'errD', 'D_x', 'D_G_z1', 'D_G_z2'):
            out[f'bf16_model.it0.{k}'] = np.float64(r[k])
        for net in ('grads_D', 'grads_G'):
            for k, v in r[net].items():
                a = v.reshape(-1)[::max(1, v.size // sample)][:sample].astype(np.float64)
                ref = out[f'it0.{net}.{k}.sample'].astype(np.float64)
                out[f'bf16_model.it0.{net}.{k}.rel_l2'] = np.float64(np.linalg.norm(a - ref) / np.linalg.norm(ref))
    for tag, net in (('G', netG), ('D', netD)):
        for k, v in to_np(net.state_dict()).items():
            if full and v.size > 4096:
                out[f'final.{tag}.{k}.l2'] = np.float64(np.sqrt((v.astype(np.float64) ** 2).sum()))
                out[f'final.{tag}.{k}.sample'] = v.reshape(-1)[::max(1, v.size // wsample)][:wsample].copy()
            else:
                out[f'final.{tag}.{k}'] = v
    path = os.path.join(GOLDEN_DIR, name)
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def make_main_fixture(ref_src, name):
    """Run the reference's unmodified train_gan.main on CPU and record what the oracle needs to replay it."""
    # matplotlib is not installed in the image: stub it before importing train_gan
    mpl = types.ModuleType('matplotlib')
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType('matplotlib.pyplot')
    for fn in ('figure', 'plot', 'title', 'xlabel', 'ylabel', 'legend', 'grid', 'tight_layout', 'savefig', 'close'):
        setattr(plt, fn, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules.setdefault('matplotlib', mpl)
    sys.modules.setdefault('matplotlib.pyplot', plt)
    import train_gan  # the reference's file, unmodified

    nz, fm, nc, bs, n_img, data_seed = 16, 8, 3, 4, 10, 777
    real = synthetic_real(data_seed, n_img, nc)

    def fake_get_dataloaders(data_dir, batch_size, num_workers):
        ds = torch.utils.data.TensorDataset(torch.from_numpy(real), torch.zeros(n_img, dtype=torch.long))
        dl = torch.utils.data.DataLoader(ds, batch_size=batch_size, shuffle=False)
        return dl, dl

    rec = dict(randn=[], init=[])
    real_randn = torch.randn

    def rec_randn(*a, **k):
        t = real_randn(*a, **k)
        rec['randn'].append(t.detach().cpu().numpy().copy())
        return t

    real_adam = torch.optim.Adam

    def rec_adam(params, *a, **k):
        params = list(params)
        rec['init'].append([p.detach().cpu().numpy().copy() for p in params])
        return real_adam(params, *a, **k)

    train_gan.get_dataloaders = fake_get_dataloaders
    train_gan.optim.Adam = rec_adam
    torch.randn = rec_randn
    tmp = tempfile.mkdtemp(prefix='golden_main_')
    args = argparse.Namespace(
        data_dir=tmp, model_dir=os.path.join(tmp, 'models'), output_dir=os.path.join(tmp, 'results'),
        results_dir=os.path.join(tmp, 'results', 'metrics'), figures_dir=os.path.join(tmp, 'results', 'figures'),
        num_channels=nc, latent_dim=nz, feature_maps_g=fm, feature_maps_d=fm, epochs=1, batch_size=bs, lr=2e-4,
        beta1=0.5, workers=0, vis_batch_size=5, save_interval=2, checkpoint_interval=10, cpu=True)
    torch.manual_seed(1234)
    try:
        train_gan.main(args)
    finally:
        torch.randn = real_randn
        train_gan.optim.Adam = real_adam
    hist = json.load(open(os.path.join(args.results_dir, 'gan_training_history.json')))
    sdG = to_np(torch.load(os.path.join(args.model_dir, 'gan', 'generator_final.pth')))
    sdD = to_np(torch.load(os.path.join(args.model_dir, 'gan', 'discriminator_final.pth')))
    out = dict(meta=json.dumps(dict(nz=nz, fm=fm, nc=nc, batch=bs, n_img=n_img, data_seed=data_seed, lr=2e-4, beta1=0.5,
                                    save_interval=2, vis_batch=5, torch=torch.__version__,
                                    files=sorted(os.listdir(os.path.join(args.output_dir, 'gan_images'))))),
               history=json.dumps(hist), fixed_noise=rec['randn'][0])
    for i, z in enumerate(rec['randn'][1:]):
        out[f'noise{i}'] = z
    keysD = orc.param_keys(orc.discriminator_plan(nc, fm))
    keysG = orc.param_keys(orc.generator_plan(nz, nc, fm))
    for k, v in zip(keysD, rec['init'][0]):      # optimizerD is created first (train_gan.py:94)
        out[f'init.D.{k}'] = v
    for k, v in zip(keysG, rec['init'][1]):
        out[f'init.G.{k}'] = v
    for k, v in sdG.items():
        out[f'final.G.{k}'] = v
    for k, v in sdD.items():
        out[f'final.D.{k}'] = v
    path = os.path.join(GOLDEN_DIR, name)
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB', 'history:', {k: v[:3] for k, v in hist.items()})


def make_wgan_fixture(wggan, name, nz, fm, nc, batch, critic_iters, seed, lambda_gp=10.0):
    """The reference's WGAN-GP iteration (train_wggan.py:70-92) with its own modules and its own, unmodified `gradient_penalty`
    (torch.autograd does the double backward); noise and the interpolation draw `alpha` (wggan.py:76) are recorded from the run."""
    import wgan_oracle as wo
    rng = np.random.RandomState(seed)
    sdG = orc.init_state(wo.wgan_generator_plan(nz, nc, fm), True, rng)
    sdD = orc.init_state(wo.critic_plan(nc, fm), False, rng)
    netG, netD = wggan.Generator(nz, nc, fm), wggan.Discriminator(nc, fm)
    load_state(netG, sdG)
    load_state(netD, sdD)
    optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.5, 0.9))
    optD = torch.optim.Adam(netD.parameters(), lr=2e-4, betas=(0.5, 0.9))
    real = torch.from_numpy(synthetic_real(seed + 1, batch, nc))
    noises = synthetic_noise(seed + 2, batch * (critic_iters + 1), nz).reshape(critic_iters + 1, batch, nz, 1, 1)
    out = dict(meta=json.dumps(dict(nz=nz, fm=fm, nc=nc, batch=batch, critic_iters=critic_iters, seed=seed, lr=2e-4, beta1=0.5, beta2=0.9,
                                    lambda_gp=lambda_gp, real_seed=seed + 1, noise_seed=seed + 2, torch=torch.__version__)))
    real_rand = torch.rand
    alphas = []

    def rec_rand(*a, **k):
        t = real_rand(*a, **k)
        alphas.append(t.detach().numpy().copy())
        return t

    for it in range(critic_iters):                       # train_wggan.py:70-85
        netD.zero_grad()
        d_real = netD(real)
        d_real_loss = -d_real.mean()
        fake = netG(torch.from_numpy(noises[it]))
        d_fake = netD(fake.detach())
        d_fake_loss = d_fake.mean()
        torch.rand = rec_rand
        try:
            gp = wggan.gradient_penalty(netD, real.data, fake.data, torch.device('cpu'), lambda_gp=lambda_gp)
        finally:
            torch.rand = real_rand
        d_loss = d_real_loss + d_fake_loss + gp
        d_loss.backward()
        out[f'c{it}.d_loss'], out[f'c{it}.gp'] = np.float64(d_loss.item()), np.float64(gp.item())
        out[f'c{it}.alpha'] = alphas[-1]
        out[f'c{it}.d_real'], out[f'c{it}.d_fake'] = d_real.detach().numpy().copy(), d_fake.detach().numpy().copy()
        if it == 0:
            for k, p in netD.named_parameters():
                out[f'c{it}.grads_D.{k}'] = p.grad.detach().numpy().copy()
        optD.step()
    netG.zero_grad()                                     # train_wggan.py:87-92
    fake = netG(torch.from_numpy(noises[critic_iters]))
    g_loss = -netD(fake).mean()
    g_loss.backward()
    out['g_loss'] = np.float64(g_loss.item())
    out['fake'] = fake.detach().numpy()[:, :, ::3, ::3].copy()
    for k, p in netG.named_parameters():
        out[f'grads_G.{k}'] = p.grad.detach().numpy().copy()
    optG.step()
    for tag, net in (('G', netG), ('D', netD)):
        for k, v in to_np(net.state_dict()).items():
            out[f'final.{tag}.{k}'] = v
    path = os.path.join(GOLDEN_DIR, name)
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB', 'd_loss', [float(out[f'c{i}.d_loss']) for i in range(critic_iters)], 'g_loss', float(out['g_loss']))


def _stub_matplotlib():
    """matplotlib is not installed in the image: stub it before importing the reference's training scripts."""
    mpl = types.ModuleType('matplotlib')
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType('matplotlib.pyplot')
    for fn in ('figure', 'plot', 'title', 'xlabel', 'ylabel', 'legend', 'grid', 'tight_layout', 'savefig', 'close', 'subplot'):
        setattr(plt, fn, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules.setdefault('matplotlib', mpl)
    sys.modules.setdefault('matplotlib.pyplot', plt)


def make_cgan_step_fixture(cgan, name, nz, nf, nc, batch, iters, seed):
    """The reference's conditional-GAN modules (src/cgan.py, unmodified) driven by the op sequence of train_cgan.py:150-193 WITHOUT the VGG16
    perceptual term (its ImageNet weights cannot be downloaded here): adversarial BCEWithLogits + 5 x feature matching.  Inputs that torch's RNG
    would draw (noise, labels, label smoothing) come from numpy and are stored; initial weights are the modules' own (torch RNG), stored."""
    torch.manual_seed(seed)
    netG, netD = cgan.Generator(nz, 2, nc, nf), cgan.Discriminator(2, nc, nf)
    rng = np.random.RandomState(seed)
    # the reference zero-initialises nothing but BatchNorm biases; give every bias / embedding a visible value so that their paths are exercised
    with torch.no_grad():
        for net in (netG, netD):
            for k, p in net.named_parameters():
                if k.endswith('bias') and p.abs().max() == 0:
                    p.copy_(torch.from_numpy((rng.randn(*p.shape) * 0.05).astype(np.float32)))
    out = dict(meta=json.dumps(dict(nz=nz, nf=nf, nc=nc, batch=batch, iters=iters, seed=seed, lr=2e-4, beta1=0.5, fm_weight=5.0, torch=torch.__version__)))
    for tag, net in (('G', netG), ('D', netD)):
        for k, v in to_np(net.state_dict()).items():
            out[f'init.{tag}.{k}'] = v
    optD = torch.optim.Adam(netD.parameters(), lr=2e-4, betas=(0.5, 0.999))
    optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.5, 0.999))
    crit = torch.nn.BCEWithLogitsLoss()
    hist = []
    for it in range(iters):
        real = torch.from_numpy(synthetic_real(seed + 10 + it, batch, nc))
        real_labels = torch.from_numpy(rng.randint(0, 2, batch).astype(np.int64))
        fake_labels = torch.from_numpy(rng.randint(0, 2, batch).astype(np.int64))
        smooth_real = torch.from_numpy((0.9 - 0.1 * rng.rand(batch)).astype(np.float32))
        smooth_fake = torch.from_numpy((0.1 + 0.1 * rng.rand(batch)).astype(np.float32))
        noise = torch.from_numpy(rng.randn(batch, nz).astype(np.float32))
        for k, v in (('real_labels', real_labels), ('fake_labels', fake_labels), ('smooth_real', smooth_real), ('smooth_fake', smooth_fake), ('noise', noise)):
            out[f'it{it}.{k}'] = v.numpy().copy()
        netD.zero_grad()
        output_real = netD(real, real_labels, 1.0)
        D_x = torch.sigmoid(output_real).mean().item()
        errD_real = crit(output_real, smooth_real)
        fake_images = netG(noise, fake_labels, 1.0)
        output_fake = netD(fake_images.detach(), fake_labels, 1.0)
        D_G_z1 = torch.sigmoid(output_fake).mean().item()
        errD = errD_real + crit(output_fake, smooth_fake)
        errD.backward()
        out[f'it{it}.out_real'], out[f'it{it}.out_fake'] = output_real.detach().numpy().copy(), output_fake.detach().numpy().copy()
        out[f'it{it}.fake'] = fake_images.detach().numpy()[:, :, ::5, ::5].copy()
        for k, p in netD.named_parameters():
            out[f'it{it}.grads_D.{k}'] = p.grad.detach().numpy().copy()
        optD.step()
        netG.zero_grad()
        output_fake = netD(fake_images, fake_labels, 1.0)
        D_G_z2 = torch.sigmoid(output_fake).mean().item()
        errG_adv = crit(output_fake, smooth_real)
        feats_real = netD.get_intermediate_features(real, real_labels, 1.0)
        feats_fake = netD.get_intermediate_features(fake_images, fake_labels, 1.0)
        errG_fm = sum(torch.mean((a - b) ** 2) for a, b in zip(feats_real, feats_fake))
        errG = errG_adv + 5.0 * errG_fm
        errG.backward()
        out[f'it{it}.feat_fake_l2'] = np.array([np.sqrt((f.detach().numpy().astype(np.float64) ** 2).sum()) for f in feats_fake])
        for k, p in netG.named_parameters():
            out[f'it{it}.grads_G.{k}'] = p.grad.detach().numpy().copy()
        optG.step()
        hist.append([errD.item(), errG.item(), D_x, D_G_z1, D_G_z2, errG_fm.item()])
    out['history'] = np.array(hist, dtype=np.float64)
    for tag, net in (('G', netG), ('D', netD)):
        for k, v in to_np(net.state_dict()).items():
            out[f'final.{tag}.{k}'] = v
    path = os.path.join(GOLDEN_DIR, name)
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB', 'history', hist)


def make_cgan_main_fixture(name):
    """The reference's UNMODIFIED `train_cgan.main(args)` on CPU.  Stubs, none of which touch the reference's files: matplotlib (absent); the dataset
    class -> a seeded in-memory dataset; `PerceptualLoss` -> a module returning 0 (VGG16's ImageNet weights cannot be downloaded offline; the term
    is the one part of train_cgan.py this fixture does not pin).  Every torch.rand / randn / randint draw of the run is recorded in order."""
    _stub_matplotlib()
    import train_cgan  # the reference's file, unmodified

    nz, nf, nc, bs, n_img, data_seed, epochs = 16, 8, 3, 4, 8, 4242, 6
    real = synthetic_real(data_seed, n_img, nc)
    labels = np.random.RandomState(data_seed + 1).randint(0, 2, n_img).astype(np.int64)

    class MemoryDataset(torch.utils.data.Dataset):
        def __init__(self, *a, **k):
            pass

        def __len__(self):
            return n_img

        def __getitem__(self, i):
            return torch.from_numpy(real[i]), torch.tensor(labels[i])

    class NoPerceptual(torch.nn.Module):
        def forward(self, x, y):
            return torch.zeros(())

    class OrderedLoader(torch.utils.data.DataLoader):          # the reference asks for shuffle=True: keep the order replayable
        def __init__(self, ds, batch_size, shuffle, num_workers):
            super().__init__(ds, batch_size=batch_size, shuffle=False, num_workers=0)

    rec = dict(draws=[], init=[])
    originals = dict(rand=torch.rand, randn=torch.randn, randint=torch.randint)

    def recorder(kind):
        def fn(*a, **k):
            t = originals[kind](*a, **k)
            rec['draws'].append((kind, t.detach().cpu().numpy().copy()))
            return t
        return fn

    real_adam = torch.optim.Adam

    def rec_adam(params, *a, **k):
        params = list(params)
        rec['init'].append([p.detach().cpu().numpy().copy() for p in params])
        return real_adam(params, *a, **k)

    train_cgan.RSNAPneumoniaDataset = MemoryDataset
    train_cgan.DataLoader = OrderedLoader
    train_cgan.PerceptualLoss = NoPerceptual
    train_cgan.optim.Adam = rec_adam
    tmp = tempfile.mkdtemp(prefix='golden_cgan_')
    args = argparse.Namespace(
        data_dir=tmp, model_dir=os.path.join(tmp, 'models'), output_dir=os.path.join(tmp, 'results'),
        results_dir=os.path.join(tmp, 'results', 'metrics'), figures_dir=os.path.join(tmp, 'results', 'figures'),
        num_channels=nc, latent_dim=nz, feature_maps_g=nf, feature_maps_d=nf, epochs=epochs, batch_size=bs, lr=2e-4, beta1=0.5, workers=0,
        vis_batch_size=5, save_interval=5, checkpoint_interval=100, cpu=True)
    torch.manual_seed(4321)
    for kind in originals:
        setattr(torch, kind, recorder(kind))
    try:
        train_cgan.main(args)
    finally:
        for kind, fn in originals.items():
            setattr(torch, kind, fn)
        train_cgan.optim.Adam = real_adam
    hist = json.load(open(os.path.join(args.results_dir, 'gan_training_history.json')))
    sdG = to_np(torch.load(os.path.join(args.model_dir, 'gan', 'generator_final.pth')))
    sdD = to_np(torch.load(os.path.join(args.model_dir, 'gan', 'discriminator_final.pth')))
    out = dict(meta=json.dumps(dict(nz=nz, nf=nf, nc=nc, batch=bs, n_img=n_img, data_seed=data_seed, epochs=epochs, lr=2e-4, beta1=0.5, save_interval=5,
                                    vis_batch=5, torch=torch.__version__, draw_kinds=[k for k, _ in rec['draws']],
                                    files=sorted(os.listdir(os.path.join(args.output_dir, 'gan_images'))))),
               history=json.dumps(hist))
    for i, (_, v) in enumerate(rec['draws']):
        out[f'draw{i}'] = v
    netG = train_cgan.Generator(nz, 2, nc, nf)
    netD = train_cgan.Discriminator(2, nc, nf)
    for (k, _), v in zip(netD.named_parameters(), rec['init'][0]):      # optimizerD is created first (train_cgan.py:124)
        out[f'init.D.{k}'] = v
    for (k, _), v in zip(netG.named_parameters(), rec['init'][1]):
        out[f'init.G.{k}'] = v
    for k, v in sdG.items():
        out[f'final.G.{k}'] = v
    for k, v in sdD.items():
        out[f'final.D.{k}'] = v
    path = os.path.join(GOLDEN_DIR, name)
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB', 'draws', len(rec['draws']), 'history:', {k: v[:3] for k, v in hist.items()})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference')
    ap.add_argument('--only', default=None, help='substring filter on the fixture names (default: regenerate all)')
    a = ap.parse_args()
    ref_src = os.path.join(a.ref, 'src')
    sys.path.insert(0, ref_src)
    import dcgan  # the reference's file, unmodified
    assert os.path.abspath(dcgan.__file__).startswith(os.path.abspath(a.ref)), dcgan.__file__
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.manual_seed(0)
    want = lambda n: a.only is None or a.only in n
    if want('step_small_nc1.npz'):
        make_step_fixture(dcgan, 'step_small_nc1.npz', nz=16, fm=8, nc=1, batch=3, iters=2, seed=100)
    if want('step_small_nc3.npz'):
        make_step_fixture(dcgan, 'step_small_nc3.npz', nz=16, fm=8, nc=3, batch=3, iters=2, seed=200)
    if want('step_full_nc1.npz'):
        make_step_fixture(dcgan, 'step_full_nc1.npz', nz=100, fm=64, nc=1, batch=2, iters=1, seed=300, full=True)
    if want('step_full_b32_nc1.npz'):
        make_step_fixture(dcgan, 'step_full_b32_nc1.npz', nz=100, fm=64, nc=1, batch=32, iters=2, seed=500, full=True, sample=4096, wsample=4096,
                          fake_stride=13, bf16_model=True)
    if want('step_full_b32_nc3.npz'):
        make_step_fixture(dcgan, 'step_full_b32_nc3.npz', nz=100, fm=64, nc=3, batch=32, iters=2, seed=600, full=True, sample=4096, wsample=4096,
                          fake_stride=13, bf16_model=True)
    if want('main_small_nc3.npz'):
        make_main_fixture(ref_src, 'main_small_nc3.npz')
    import wggan  # the reference's file, unmodified
    assert os.path.abspath(wggan.__file__).startswith(os.path.abspath(a.ref)), wggan.__file__
    if want('wgan_small_nc1.npz'):
        make_wgan_fixture(wggan, 'wgan_small_nc1.npz', nz=16, fm=8, nc=1, batch=3, critic_iters=2, seed=800)
    if want('wgan_small_nc3.npz'):
        make_wgan_fixture(wggan, 'wgan_small_nc3.npz', nz=16, fm=8, nc=3, batch=4, critic_iters=2, seed=900)
    import cgan  # the reference's file, unmodified
    assert os.path.abspath(cgan.__file__).startswith(os.path.abspath(a.ref)), cgan.__file__
    if want('cgan_step_nc1.npz'):
        make_cgan_step_fixture(cgan, 'cgan_step_nc1.npz', nz=16, nf=8, nc=1, batch=3, iters=2, seed=1100)
    if want('cgan_step_nc3.npz'):
        make_cgan_step_fixture(cgan, 'cgan_step_nc3.npz', nz=16, nf=8, nc=3, batch=4, iters=2, seed=1200)
    if want('cgan_main_nc3.npz'):
        make_cgan_main_fixture('cgan_main_nc3.npz')


if __name__ == '__main__':
    main()
