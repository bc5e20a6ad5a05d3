"""Conditional GAN on the GPU: the drop-in `cgan` modules (label conditioning, Linear as latent GEMM, nearest
Upsample + Conv2d(3) folded into one stride-2 transposed convolution, biased convolutions, projection head, get_intermediate_features) through
the C ABI, against the numpy oracle (oracle/cgan_oracle.py, which states upsample-then-convolve literally) and fixtures produced by the reference
itself (tests/golden/cgan_*.npz: /root/reference/src/cgan.py modules and the unmodified train_cgan.main with the VGG16 perceptual term stubbed
to 0 -- its ImageNet weights cannot be downloaded offline, and that term is the part of train_cgan.py this path does not cover)."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

import cgan_oracle as co
import dcgan_oracle as orc
from conftest import GOLDEN
from gan_enhanced_pneumonia_classifier_b200 import _lib as L
from gan_enhanced_pneumonia_classifier_b200 import cgan
from gan_enhanced_pneumonia_classifier_b200.engine import Act
from parity_utils import close, grad_close, synthetic_real
from test_oracle_golden import PRE_BN_BIASES_D, PRE_BN_BIASES_G, cgan_state

pytestmark = pytest.mark.gpu


def st():
    return L.stream_ptr()


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def adam_scale_close(v, ref, lr, iters, what, bulk=True):
    """Post-Adam weights on Adam's own scale: every entry inside the sign-flip envelope 2 lr iters; for tensors large enough for a fraction to
    mean something, four in five within an eighth of it."""
    dd = np.abs(np.asarray(v, np.float64) - ref)
    assert dd.max() <= 2.05 * lr * iters, f'{what}: max diff {dd.max():.3e} outside the Adam envelope'
    if bulk and dd.size >= 64:
        frac = (dd <= 0.25 * lr * iters).mean()
        assert frac >= 0.8, f'{what}: only {frac:.3f} of the entries within 0.25 lr iters'


# ---------------------------------------------------------------------------------------------------------------- single entry points
def test_embed_add_and_backward_match_numpy():
    rng = np.random.RandomState(1)
    table, z = rng.randn(3, 20).astype(np.float32), rng.randn(7, 20).astype(np.float32)
    labels = rng.randint(0, 3, 7).astype(np.int64)
    out = torch.empty((7, 21), device='cuda')
    table_d, labels_d, z_d = dev(table), dev(labels), dev(z)          # (named: the raw pointers below do not keep temporaries alive)
    L.call('b200gan_embed_add', L.ptr(table_d), L.ptr(labels_d), L.ptr(z_d), 7, 20, 1, L.ptr(out), st())
    np.testing.assert_array_equal(out.cpu().numpy(), np.concatenate([table[labels] + z, np.ones((7, 1), np.float32)], 1))
    L.call('b200gan_embed_add', L.ptr(table_d), L.ptr(labels_d), None, 7, 20, 0, L.ptr(out), st())
    np.testing.assert_array_equal(out.view(-1)[:140].cpu().numpy().reshape(7, 20), table[labels])
    dx = rng.randn(7, 21).astype(np.float32)
    dx_d = dev(dx)
    dtable = torch.full((3, 20), 0.5, device='cuda')
    L.call('b200gan_embed_bwd', L.ptr(dx_d), L.ptr(labels_d), 7, 20, 21, 3, L.ptr(dtable), st())
    close(dtable.cpu().numpy(), 0.5 + co.embedding_bwd(dx[:, :20], labels, 3), rtol=1e-6, atol=1e-6, what='embedding backward (accumulating)')


@pytest.mark.parametrize('ci,co_,h', [(8, 4, 7), (64, 32, 14), (4, 3, 28), (16, 1, 5)])
def test_folded_upsample_conv_equals_upsample_then_conv(ci, co_, h):
    """b200gan_upconv3_fold + ConvTranspose2d(4,2,1) kernels == nearest Upsample(2) followed by Conv2d(3,1,1), forward, input gradient and
    weight gradient (through b200gan_upconv3_unfold), against the oracle's literal upsample-then-convolve."""
    rng = np.random.RandomState(ci * 100 + co_)
    x = rng.randn(3, ci, h, h).astype(np.float32)
    w3 = (rng.randn(co_, ci, 3, 3) * 0.1).astype(np.float32)
    dy = rng.randn(3, co_, 2 * h, 2 * h).astype(np.float32)
    up = co.upsample2(x)
    y_ref = orc.conv2d_fprop(up, w3, 1, 1)
    dx_ref = co.upsample2_bwd(orc.conv2d_dgrad(dy, w3, 1, 1, up.shape[2:]))
    dw_ref = orc.conv2d_wgrad(up, dy, 3, 1, 1)
    w4 = torch.empty((ci, co_, 4, 4), device='cuda')
    w3_d = dev(w3)
    L.call('b200gan_upconv3_fold', L.ptr(w3_d), co_, ci, L.ptr(w4), st())
    conv = L.Conv(4, 2, 1, L.ALGO_AUTO)
    xa, dya = Act(dev(x), nchw=True), Act(dev(dy), nchw=True)
    y = torch.empty((3, co_, 2 * h, 2 * h), device='cuda')
    L.call('b200gan_convT2d_fprop', C.byref(conv), C.byref(xa.v), L.ptr(w4), None, C.byref(Act(y, nchw=True).v), None, st())
    close(y.cpu().numpy(), y_ref, rtol=1e-4, atol=1e-5, what='folded forward')
    dx = torch.empty_like(xa.t)
    L.call('b200gan_convT2d_dgrad', C.byref(conv), C.byref(dya.v), L.ptr(w4), None, C.byref(Act(dx, nchw=True).v), None, st())
    close(dx.cpu().numpy(), dx_ref, rtol=1e-4, atol=1e-4, what='folded input gradient')
    dw4 = torch.zeros_like(w4)
    L.call('b200gan_convT2d_wgrad', C.byref(conv), C.byref(xa.v), C.byref(dya.v), L.ptr(dw4), None, None, st())
    dw3 = torch.full((co_, ci, 3, 3), 0.25, device='cuda')
    L.call('b200gan_upconv3_unfold', L.ptr(dw4), co_, ci, L.ptr(dw3), st())
    grad_close(dw3.cpu().numpy() - 0.25, dw_ref, 'unfolded weight gradient', bulk=1e-5, l2=1e-4, worst=1e-3)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_class_projection_matches_numpy(dtype):
    rng = np.random.RandomState(5)
    n, c, hw, classes = 5, 16, 7, 3
    x = rng.randn(n, c, hw, hw).astype(np.float32)
    x_t = dev(x).permute(0, 2, 3, 1).contiguous().to(dtype)            # NHWC storage, as the engine keeps it
    x_seen = x_t.float().permute(0, 3, 1, 2).cpu().numpy()
    table = rng.randn(classes, c * hw * hw).astype(np.float32)
    labels = rng.randint(0, classes, n).astype(np.int64)
    out = torch.full((n,), 2.0, device='cuda')
    xa = Act(x_t, nchw=False)
    table_d, labels_d = dev(table), dev(labels)
    L.call('b200gan_class_proj_fwd', C.byref(xa.v), L.ptr(table_d), L.ptr(labels_d), L.ptr(out), st())
    ref = 2.0 + (table[labels].astype(np.float64) * x_seen.reshape(n, -1)).sum(1)
    close(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-4, what='projection')
    dout = rng.randn(n).astype(np.float32)
    dx = torch.zeros((n, hw, hw, c), device='cuda', dtype=torch.float32)
    dtable = torch.zeros((classes, c * hw * hw), device='cuda')
    dout_d = dev(dout)
    L.call('b200gan_class_proj_bwd', C.byref(xa.v), L.ptr(table_d), L.ptr(labels_d), L.ptr(dout_d), C.byref(Act(dx, nchw=False).v), classes,
           L.ptr(dtable), st())
    close(dx.permute(0, 3, 1, 2).cpu().numpy(), (dout[:, None] * table[labels]).reshape(n, c, hw, hw), rtol=1e-6, atol=1e-6, what='projection dx')
    close(dtable.cpu().numpy(), co.embedding_bwd(dout[:, None] * x_seen.reshape(n, -1), labels, classes), rtol=1e-5, atol=1e-5, what='projection dtable')


# ---------------------------------------------------------------------------------------------------------------- the modules
def build(g, m, dtype):
    G, D = cgan.Generator(m['nz'], 2, m['nc'], m['nf']), cgan.Discriminator(2, m['nc'], m['nf'])
    G.load_state_dict({k: torch.from_numpy(v) for k, v in cgan_state(g, 'G').items()}, strict=False)
    D.load_state_dict({k: torch.from_numpy(v) for k, v in cgan_state(g, 'D').items()}, strict=False)
    G, D = G.cuda(), D.cuda()
    G.compute_dtype = D.compute_dtype = dtype
    return G, D


def reference_iteration(G, D, optG, optD, real, real_labels, smooth_real, smooth_fake, noise, fake_labels, epoch=0, perceptual=None):
    """train_cgan.py:150-193 restated over whatever modules it is given (the perceptual term pluggable; None = dropped, as in the fixtures)."""
    crit = torch.nn.BCEWithLogitsLoss()
    D.zero_grad()
    out_real = D(real, real_labels, 1.0)
    d_x = torch.sigmoid(out_real).mean().item()
    err_real = crit(out_real, smooth_real)
    fake = G(noise, fake_labels, 1.0)
    out_fake = D(fake.detach(), fake_labels, 1.0)
    d_g_z1 = torch.sigmoid(out_fake).mean().item()
    err_d = err_real + crit(out_fake, smooth_fake)
    stepped = d_x < 0.8 or d_g_z1 > 0.2 or epoch < 5
    grads_d = None
    if stepped:
        err_d.backward()
        grads_d = {k: p.grad.detach().cpu().numpy().copy() for k, p in D.named_parameters()}
        optD.step()
    G.zero_grad()
    out_g = D(fake, fake_labels, 1.0)
    d_g_z2 = torch.sigmoid(out_g).mean().item()
    err_adv = crit(out_g, smooth_real)
    fr = D.get_intermediate_features(real, real_labels, 1.0)
    ff = D.get_intermediate_features(fake, fake_labels, 1.0)
    err_fm = sum(torch.mean((a - b) ** 2) for a, b in zip(fr, ff))
    err_g = err_adv + 5.0 * err_fm + (10.0 * perceptual(fake, real) if perceptual is not None else 0.0)
    err_g.backward()
    grads_g = {k: p.grad.detach().cpu().numpy().copy() for k, p in G.named_parameters()}
    optG.step()
    return dict(row=np.array([err_d.item(), err_g.item(), d_x, d_g_z1, d_g_z2, err_fm.item()]), out_real=out_real.detach().cpu().numpy(),
                out_fake=out_fake.detach().cpu().numpy(), fake=fake.detach().cpu().numpy(), grads_D=grads_d, grads_G=grads_g,
                feat_l2=np.array([f.detach().double().norm().item() for f in ff]), stepped=stepped)


@pytest.mark.parametrize('name', ['cgan_step_nc1.npz', 'cgan_step_nc3.npz'])
def test_reference_loop_over_dropin_modules_matches_reference_fixture(name):
    """fp32 parity mode: logits, image, feature norms, every gradient of both networks (first iteration, entry by entry), history rows and
    the weights after two Adam steps, against what the reference's own modules produced on CPU."""
    g = np.load(os.path.join(GOLDEN, name))
    m = json.loads(str(g['meta']))
    G, D = build(g, m, torch.float32)
    optD = torch.optim.Adam(D.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    optG = torch.optim.Adam(G.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    for it in range(m['iters']):
        real = dev(synthetic_real(m['seed'] + 10 + it, m['batch'], m['nc']))
        r = reference_iteration(G, D, optG, optD, real, dev(g[f'it{it}.real_labels']), dev(g[f'it{it}.smooth_real']), dev(g[f'it{it}.smooth_fake']),
                                dev(g[f'it{it}.noise']), dev(g[f'it{it}.fake_labels']))
        ref = g['history'][it]
        # Tolerances: fp32 rtol 1e-4 where the arithmetic is well conditioned (logits of real images, the image, errD).  The fake images of an
        # untrained Generator are nearly constant, the Discriminator's BatchNorm inputs then have |mean|/std in the tens to hundreds, and every
        # fp32 BatchNorm (torch's included) carries eps * |mean|/std of rounding there; the projection head sums 3136 N(0,1)-weighted features on
        # top.  Quantities downstream of D(fake) are therefore held to 5e-4 of the logit scale (tools/diag_cgan.py shows the layer-by-layer growth).
        close(r['row'][[0, 5]], ref[[0, 5]], what=f'errD / feature matching of iteration {it}', **(dict(rtol=1e-4, atol=1e-5) if it == 0 else dict(rtol=5e-3, atol=1e-4)))
        close(r['row'][1], ref[1], what=f'errG of iteration {it}', rtol=5e-4 if it == 0 else 2e-2)
        close(r['row'][2:5], ref[2:5], what=f'sigmoid means of iteration {it}', rtol=0, atol=3e-3)
        if it == 0:
            close(r['out_real'], g['it0.out_real'], rtol=1e-4, atol=2e-4, what='logits(real)')
            close(r['out_fake'], g['it0.out_fake'], rtol=0, atol=5e-4 * float(np.abs(g['it0.out_fake']).max()), what='logits(fake)')
            close(r['fake'][:, :, ::5, ::5], g['it0.fake'], rtol=1e-4, atol=1e-5, what='fake image')
            act_idx = [0, 1, 3, 4, 6, 7, 9, 10, 12, 13]       # LeakyReLU outputs; the other four are conv outputs INCLUDING the pre-BatchNorm bias,
            close(r['feat_l2'][act_idx], g['it0.feat_fake_l2'][act_idx], rtol=1e-3, what='norms of the activation features')
            close(r['feat_l2'], g['it0.feat_fake_l2'], rtol=1e-2, what='norms of the 14 intermediate features')      # which took a noise-driven Adam step
            fails = []
            for net, pre in (('grads_D', PRE_BN_BIASES_D), ('grads_G', PRE_BN_BIASES_G)):
                for k, v in r[net].items():
                    if k in pre:
                        continue
                    try:
                        # conditioning-limited (see the tolerance note above); the kernel-level fp32 check is test_full_width_modules_against_the_oracle
                        grad_close(v, g[f'it0.{net}.{k}'], f'{net}.{k}', **(dict(bulk=2e-3, l2=5e-3, worst=2e-2) if net == 'grads_D' else dict(bulk=3e-2, l2=6e-2, worst=0.1)))
                    except AssertionError as e:
                        fails.append(str(e))
            assert not fails, '\n'.join(fails)
    for tag, net, pre in (('G', G, PRE_BN_BIASES_G), ('D', D, PRE_BN_BIASES_D)):
        for k, v in net.state_dict().items():
            v, ref = v.cpu().numpy(), g[f'final.{tag}.{k}']
            if k.endswith('num_batches_tracked'):
                assert int(v) == int(ref), k
            elif 'running' in k:
                close(v, ref, rtol=1e-3, atol=1e-4 * max(1.0, float(np.abs(ref).max())) + (2.05 * m['lr'] * m['iters'] if k.endswith('running_mean') else 0), what=k)
            elif k in pre:
                assert np.abs(v - ref).max() <= 2.05 * m['lr'] * m['iters'], k
            else:
                # on Adam's scale (the second step runs on conditioning-limited gradients, see above): all inside the sign-flip envelope, nine in
                # ten within an eighth of it
                adam_scale_close(v, ref, m['lr'], m['iters'], f'final.{tag}.{k}')


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


GRAD_MODEL_FACTOR, GRAD_MODEL_SLACK = 1.5, 0.02          # as tests/test_gpu_fullsize.py: bound = factor * (what bf16 storage costs by itself) + slack


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_full_width_modules_against_the_oracle(dtype):
    """fp32: outputs and gradients against the oracle on well-conditioned inputs (the tight kernel-level check that the reference-fixture test
    above cannot give: its fake batches sit where BatchNorm is ill conditioned).  bf16: the CLI-default width (feature_maps 32: Generator 256-128-64-32-16-nc, its first three folded upsample-convolutions and the
    Discriminator's 32-64-128-256 layers on the tcgen05 kernels) with bf16 storage against the fp32 numpy oracle.  The gradient bounds are
    calibrated per tensor by the same oracle run with bf16 STORAGE rounding (cgan_oracle's `storage` hook): at batch 8 the sign flips of ReLU /
    LeakyReLU units and the cancellation inside BatchNorm's backward make bf16 storage alone cost 2-15 % relative L2 on these gradients, in any
    implementation.  Each network gets inputs of its own (the Discriminator noise images, the Generator a given upstream gradient): an untrained
    Generator's near-constant images put |mean|/std of the Discriminator's first BatchNorm inputs at 10-25, which multiplies storage rounding by
    that factor (tools/diag_cgan.py full)."""
    nz, nf, nc, n = 100, 32, 3, 8
    torch.manual_seed(7)
    G, D = cgan.Generator(nz, 2, nc, nf), cgan.Discriminator(2, nc, nf)
    with torch.no_grad():
        D.label_emb.weight.mul_(0.02)
    sdG = {k: v.numpy().copy() for k, v in G.state_dict().items()}
    sdD = {k: v.numpy().copy() for k, v in D.state_dict().items()}
    G, D = G.cuda(), D.cuda()
    G.compute_dtype = D.compute_dtype = dtype
    rng = np.random.RandomState(3)
    z, labels = rng.randn(n, nz).astype(np.float32), rng.randint(0, 2, n).astype(np.int64)
    target = (0.9 - 0.1 * rng.rand(n)).astype(np.float32)
    x = synthetic_real(99, n, nc)
    r = (rng.randn(n, nc, 224, 224) * 1e-3).astype(np.float32)

    def oracle(storage):
        oD = co.DiscriminatorOracle(2, nc, nf, {k: v.copy() for k, v in sdD.items()}, storage=storage)
        logit, cd = oD.forward(x, labels)
        dx, gd = oD.backward(cd, co.bce_logits(logit, target)[1], None, need_input_grad=True)
        oG = co.GeneratorOracle(nz, 2, nc, nf, {k: v.copy() for k, v in sdG.items()}, storage=storage)
        fake, cg = oG.forward(z, labels)
        _, gg = oG.backward(cg, r)
        out = {'D.logits': logit, 'D.input': dx, 'G.image': fake}
        out.update({f'D.{k}': v for k, v in gd.items() if k not in PRE_BN_BIASES_D})
        out.update({f'G.{k}': v for k, v in gg.items() if k not in PRE_BN_BIASES_G})
        return out

    ref = oracle(None)
    model = oracle(orc.bf16_round) if dtype == torch.bfloat16 else None
    # ---- the CUDA path
    x_t = dev(x).requires_grad_(True)
    logits = D(x_t, dev(labels))
    torch.nn.BCEWithLogitsLoss()(logits, dev(target)).backward()
    fake = G(dev(z), dev(labels))
    (fake * dev(r)).sum().backward()
    got = {'D.logits': logits.detach().cpu().numpy(), 'D.input': x_t.grad.cpu().numpy(), 'G.image': fake.detach().cpu().numpy()}
    got.update({f'D.{k}': p.grad.cpu().numpy() for k, p in D.named_parameters() if k not in PRE_BN_BIASES_D})
    got.update({f'G.{k}': p.grad.cpu().numpy() for k, p in G.named_parameters() if k not in PRE_BN_BIASES_G})
    assert set(got) == set(ref)
    report, bad = {}, {}
    for k in ref:
        err = _rel(got[k], ref[k])
        cost = _rel(model[k], ref[k]) if model is not None else 0.0
        report[k] = (float('%.3g' % err), float('%.3g' % cost))
        if err > (GRAD_MODEL_FACTOR * cost + GRAD_MODEL_SLACK if model is not None else 5e-3):
            bad[k] = report[k]
    print(f'relative L2 per tensor (CUDA {dtype} path, bf16-storage oracle):', report)
    assert not bad, bad
    if dtype == torch.bfloat16:
        assert report['G.image'][0] < 2e-2 and report['D.logits'][0] < 5e-2
    else:
        # fp32: 1e-6 wherever no unit changed branch; ONE ReLU among the 8e5 of a layer landing on the other side of 0 (its pre-activation is
        # within fp32 rounding of 0 in one of the two summation orders) moves every gradient below it by 1/sqrt(8e5) ~ 1e-3 relative L2
        # (tests/parity_utils.py), hence 5e-3 per tensor above and a tight median here
        assert np.median([v[0] for v in report.values()]) < 2e-5
        assert max(report[k][0] for k in ('D.logits', 'G.image', 'D.main.14.weight', 'G.main.19.weight')) < 2e-5


def test_reference_main_replayed_over_dropin_modules():
    """The reference's unmodified train_cgan.main (fixture cgan_main_nc3.npz), replayed on the GPU from its recorded random draws: twelve
    iterations, train-mode visualisation forwards, final state dicts."""
    g = np.load(os.path.join(GOLDEN, 'cgan_main_nc3.npz'))
    m = json.loads(str(g['meta']))
    hist = json.loads(str(g['history']))
    G, D = build(g, m, torch.float32)
    optD = torch.optim.Adam(D.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    optG = torch.optim.Adam(G.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    real = synthetic_real(m['data_seed'], m['n_img'], m['nc'])
    labels = np.random.RandomState(m['data_seed'] + 1).randint(0, 2, m['n_img']).astype(np.int64)
    fixed_noise = dev(g['draw0'])
    fixed_labels = dev(np.tile(np.arange(2), m['vis_batch'] // 2 + 1)[:m['vis_batch']].astype(np.int64))
    per_epoch = m['n_img'] // m['batch']
    it, rows = 0, []
    for epoch in range(m['epochs']):
        for b in range(per_epoch):
            d = 1 + 4 * it
            sl = slice(b * m['batch'], (b + 1) * m['batch'])
            r = reference_iteration(G, D, optG, optD, dev(real[sl]), dev(labels[sl]), dev((0.9 - 0.1 * g[f'draw{d}']).astype(np.float32)),
                                    dev((0.1 + 0.1 * g[f'draw{d + 1}']).astype(np.float32)), dev(g[f'draw{d + 2}']), dev(g[f'draw{d + 3}']), epoch=epoch)
            rows.append(r['row'])
            if it % m['save_interval'] == 0 or (epoch == m['epochs'] - 1 and b == per_epoch - 1):
                with torch.no_grad():
                    vis = G(fixed_noise, fixed_labels, 1.0)
                assert tuple(vis.shape) == (m['vis_batch'], m['nc'], 224, 224)
            it += 1
    rows = np.array(rows).reshape(m['epochs'], per_epoch, -1).mean(axis=1)
    for col, key in ((0, 'D_losses_epoch'), (1, 'G_losses_epoch'), (5, 'feature_matching_losses')):
        close(rows[:2, col], hist[key][:2], rtol=5e-3, atol=1e-3, what=f'{key}, first two epochs')        # chaotic afterwards (see the oracle's replay)
        close(rows[:, col], hist[key], rtol=0.15, atol=1e-2, what=key)
    n_it = m['epochs'] * per_epoch
    for tag, net, pre in (('G', G, PRE_BN_BIASES_G), ('D', D, PRE_BN_BIASES_D)):
        sd = net.state_dict()
        for k in sd:
            v, ref = sd[k].cpu().numpy(), g[f'final.{tag}.{k}']
            if k.endswith('num_batches_tracked'):
                assert int(v) == int(ref), k
            elif 'running' not in k:
                adam_scale_close(v, ref, m['lr'], n_it, f'final.{tag}.{k}', bulk=k not in pre)


def test_state_dicts_interoperate_and_eval_forward_matches_oracle():
    """A checkpoint written by the reference's modules loads into the drop-in ones (same keys / shapes / dtypes) and the eval-mode sampler
    forward (running statistics) reproduces the oracle's."""
    g = np.load(os.path.join(GOLDEN, 'cgan_main_nc3.npz'))
    m = json.loads(str(g['meta']))
    sd = {k[len('final.G.'):]: np.array(g[k]) for k in g.files if k.startswith('final.G.')}
    G = cgan.Generator(m['nz'], 2, m['nc'], m['nf'])
    assert list(G.state_dict().keys()) == list(sd.keys())
    G.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    G = G.cuda().eval()
    G.compute_dtype = torch.float32
    rng = np.random.RandomState(11)
    z, labels = rng.randn(3, m['nz']).astype(np.float32), np.array([0, 1, 1], dtype=np.int64)
    with torch.no_grad():
        img = G(dev(z), dev(labels))
    ref, _ = co.GeneratorOracle(m['nz'], 2, m['nc'], m['nf'], sd).forward(z, labels, train=False)
    close(img.cpu().numpy(), ref, rtol=1e-4, atol=1e-5, what='eval-mode sample')
    assert int(G.main[0].num_batches_tracked) == int(sd['main.0.num_batches_tracked'])       # eval forward leaves the buffers alone


# ---------------------------------------------------------------------------------------------------------------- the fused trainer
def test_loss_kernels_match_numpy():
    rng = np.random.RandomState(17)
    x, t = (rng.randn(37) * 8).astype(np.float32), rng.rand(37).astype(np.float32)
    x_d, t_d = dev(x), dev(t)
    out2, dl = torch.empty(2, device='cuda'), torch.empty(37, device='cuda')
    L.call('b200gan_bce_logits', L.ptr(x_d), L.ptr(t_d), 37, 1.0, L.ptr(out2), L.ptr(dl), st())
    loss_ref, dl_ref = co.bce_logits(x, t)
    close(out2.cpu().numpy(), [loss_ref, orc.sigmoid(x).mean(dtype=np.float64)], rtol=1e-5, atol=1e-6, what='BCEWithLogits / mean sigmoid')
    close(dl.cpu().numpy(), dl_ref, rtol=1e-5, atol=1e-7, what='BCEWithLogits gradient')
    r, f = rng.randn(3, 5, 6, 6).astype(np.float32), rng.randn(3, 5, 6, 6).astype(np.float32)
    ra, fa = Act(dev(r), nchw=True), Act(dev(f).permute(0, 2, 3, 1).contiguous(), nchw=False)      # mixed layouts
    d = Act(torch.full((3, 6, 6, 5), 1.0, device='cuda'), nchw=False)
    s = torch.zeros(1, device='cuda', dtype=torch.float64)
    L.call('b200gan_fm_pair', C.byref(ra.v), C.byref(fa.v), C.byref(d.v), -0.5, 1, L.ptr(s), st())
    close(s.item(), ((r.astype(np.float64) - f) ** 2).sum(), rtol=1e-6, what='feature-matching sum')
    close(d.t.permute(0, 3, 1, 2).cpu().numpy(), 1.0 - 0.5 * (r - f), rtol=1e-6, atol=1e-6, what='feature-matching gradient (accumulating)')
    # the dense bf16 fast path (three NHWC tensors of one layout), accumulating and overwriting
    rb, fb = dev(rng.randn(4, 6, 6, 32).astype(np.float32), torch.bfloat16), dev(rng.randn(4, 6, 6, 32).astype(np.float32), torch.bfloat16)
    for add in (1, 0):
        db = torch.full((4, 6, 6, 32), 0.5, device='cuda', dtype=torch.bfloat16)
        s.zero_()
        L.call('b200gan_fm_pair', C.byref(Act(rb, nchw=False).v), C.byref(Act(fb, nchw=False).v), C.byref(Act(db, nchw=False).v), 0.25, add, L.ptr(s), st())
        diff = rb.double() - fb.double()
        close(s.item(), (diff ** 2).sum().item(), rtol=1e-6, what='feature-matching sum, dense bf16')
        close(db.float().cpu().numpy(), ((0.5 if add else 0.0) + 0.25 * diff).float().cpu().numpy(), rtol=8e-3, atol=4e-3, what='feature-matching gradient, dense bf16')
    src = rng.randn(4, 6)
    dst = torch.full((6, 4), 2.0, device='cuda')
    src_d = dev(src)
    L.call('b200gan_accumulate_2d', L.ptr(dst), L.ptr(src_d), 1, 6, 4, 1, 6, st())             # dst += src^T, src in fp64
    close(dst.cpu().numpy(), 2.0 + src.T.astype(np.float32), rtol=1e-6, atol=1e-6, what='accumulate_2d')


@pytest.mark.parametrize('name', ['cgan_step_nc1.npz', 'cgan_step_nc3.npz'])
def test_fused_trainer_matches_reference_fixture_and_the_module_loop(name):
    """CGANTrainer.step (fp32) against the reference-generated fixture, with the tolerances of the module-level test above, and against the
    reference loop run over the drop-in modules on the same device (same kernels, different bookkeeping: arenas, one combined backward
    through the Discriminator, replayed running statistics): tightly."""
    from gan_enhanced_pneumonia_classifier_b200.cgan_trainer import CGANTrainer
    g = np.load(os.path.join(GOLDEN, name))
    m = json.loads(str(g['meta']))
    G, D = build(g, m, torch.float32)
    G2, D2 = build(g, m, torch.float32)
    tr = CGANTrainer(G, D, lr=m['lr'], beta1=m['beta1'], perceptual_weight=0.0, dtype=torch.float32)
    optD = torch.optim.Adam(D2.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    optG = torch.optim.Adam(G2.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    for it in range(m['iters']):
        real = dev(synthetic_real(m['seed'] + 10 + it, m['batch'], m['nc']))
        draws = [dev(g[f'it{it}.{k}']) for k in ('real_labels', 'smooth_real', 'smooth_fake', 'noise', 'fake_labels')]
        row = tr.step(real, draws[0], epoch=0, noise=draws[3], fake_labels=draws[4], smooth_real=draws[1], smooth_fake=draws[2]).cpu().numpy().astype(np.float64)
        r = reference_iteration(G2, D2, optG, optD, real, *draws)
        ref = g['history'][it]
        mine = row[[0, 1, 2, 3, 4, 6]]
        close(mine[[0, 5]], ref[[0, 5]], what=f'errD / feature matching of iteration {it}', **(dict(rtol=1e-4, atol=1e-5) if it == 0 else dict(rtol=5e-3, atol=1e-4)))
        close(mine[1], ref[1], what=f'errG of iteration {it}', rtol=5e-4 if it == 0 else 2e-2)
        close(mine[2:5], ref[2:5], what=f'sigmoid means of iteration {it}', rtol=0, atol=3e-3)
        assert row[5] == 0.0
        close(mine, r['row'], rtol=2e-5 if it == 0 else 2e-3, atol=1e-5 if it == 0 else 1e-3, what=f'trainer vs module loop, iteration {it}')
    assert tr.d_steps == m['iters']
    for (k, a), b in zip(G.state_dict().items(), G2.state_dict().values()):
        if k.endswith('num_batches_tracked'):
            assert int(a) == int(b) == int(g[f'final.G.{k}']), k
    for (k, a), b in zip(D.state_dict().items(), D2.state_dict().values()):
        if k.endswith('num_batches_tracked'):
            assert int(a) == int(b) == int(g[f'final.D.{k}']), k            # five train-mode passes per iteration, one of them replayed
        elif 'running' in k:
            close(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-3, atol=1e-4 * max(1.0, float(b.abs().max())) + 2.05 * m['lr'] * m['iters'], what=f'D.{k}')
    for tag, net, pre in (('G', G, PRE_BN_BIASES_G), ('D', D, PRE_BN_BIASES_D)):
        for k, v in net.state_dict().items():
            if 'running' in k or k.endswith('num_batches_tracked'):
                continue
            adam_scale_close(v.cpu().numpy(), g[f'final.{tag}.{k}'], m['lr'], m['iters'], f'final.{tag}.{k}', bulk=k not in pre)


def test_fused_trainer_skips_the_d_step_by_the_reference_rule():
    """train_cgan.py:176-178: from epoch 5 on the Discriminator is only updated while D(x) < 0.8 or D(G(z)) > 0.2.  The projection table is set so
    that the Discriminator is confidently right (real batch labelled 0 scores high, fake batch labelled 1 scores low): the step must be skipped at
    epoch 5 and taken at epoch 4; the Generator steps either way."""
    from gan_enhanced_pneumonia_classifier_b200.cgan_trainer import CGANTrainer
    g = np.load(os.path.join(GOLDEN, 'cgan_step_nc3.npz'))
    m = json.loads(str(g['meta']))
    n = m['batch']
    real = dev(synthetic_real(m['seed'] + 10, n, m['nc']))
    real_labels, fake_labels = torch.zeros(n, dtype=torch.int64, device='cuda'), torch.ones(n, dtype=torch.int64, device='cuda')
    draws = [dev(g[f'it0.{k}']) for k in ('smooth_real', 'smooth_fake', 'noise')]
    for epoch, expect in ((5, False), (4, True)):
        G, D = build(g, m, torch.float32)
        with torch.no_grad():
            a_real = D.get_intermediate_features(real, real_labels)[-1].flatten(1).mean(0)
            a_fake = D.get_intermediate_features(G(draws[2], fake_labels), fake_labels)[-1].flatten(1).mean(0)
            D.main[14].weight.zero_()
            D.main[14].bias.zero_()
            D.label_emb.weight[0] = 8.0 * a_real / a_real.dot(a_real)
            D.label_emb.weight[1] = -8.0 * a_fake / a_fake.dot(a_fake)
        tr = CGANTrainer(G, D, lr=m['lr'], beta1=m['beta1'], perceptual_weight=0.0, dtype=torch.float32)
        before = tr.arenaD.param.clone()
        row = tr.step(real, real_labels, epoch=epoch, noise=draws[2], fake_labels=fake_labels, smooth_real=draws[0], smooth_fake=draws[1]).cpu().numpy()
        assert row[2] >= 0.8 and row[3] <= 0.2, row
        assert (tr.d_steps == 1) == expect and bool((tr.arenaD.param != before).any()) == expect
        assert bool((tr.arenaG.exp_avg != 0).any())


def test_fused_trainer_bf16_full_width_runs_and_learns_finite():
    from gan_enhanced_pneumonia_classifier_b200.cgan_trainer import CGANTrainer
    torch.manual_seed(5)
    G, D = cgan.Generator(100, 2, 3, 32).cuda(), cgan.Discriminator(2, 3, 32).cuda()
    tr = CGANTrainer(G, D, perceptual_weight=0.0, dtype=torch.bfloat16)
    with pytest.raises(L.B200GanError):
        CGANTrainer(G, D)                                  # the reference's loss (perceptual weight 10) without a VGG16: refused, not dropped
    real = dev(synthetic_real(1, 16, 3))
    labels = dev(np.random.RandomState(2).randint(0, 2, 16).astype(np.int64))
    rows = torch.stack([tr.step(real, labels, epoch=e) for e in range(3)]).cpu().numpy()
    assert np.isfinite(rows).all() and (rows[:, 6] > 0).all()
    for p in list(G.parameters()) + list(D.parameters()):
        assert torch.isfinite(p).all()
    assert int(D.main[3].num_batches_tracked) == 15 and int(G.main[0].num_batches_tracked) == 3


def test_cgan_cli_on_gpu(tmp_path):
    from gan_enhanced_pneumonia_classifier_b200 import train_cgan as tc
    d = str(tmp_path)
    base = ['--synthetic', '8', '--batch-size', '4', '--epochs', '1', '--vis-batch-size', '4', '--model-dir', d + '/models', '--output-dir', d + '/results',
            '--results-dir', d + '/results/metrics', '--figures-dir', d + '/results/figures', '--seed', '0', '--checkpoint-interval', '1']
    # the reference's default needs torchvision's ImageNet checkpoint: offline (and uncached) the CLI stops instead of training something else
    assert tc.main(tc.build_parser().parse_args(base + ['--vgg-weights', os.path.join(d, 'no_such_checkpoint.pth')])) is None
    hist = tc.main(tc.build_parser().parse_args(base + ['--no-perceptual']))
    assert len(hist['G_losses_epoch']) == 1 and np.isfinite(hist['G_losses_epoch'][0]) and hist['perceptual_losses'] == [0.0]
    hist = tc.main(tc.build_parser().parse_args(base + ['--vgg-weights', 'random']))                 # the full generator loss of train_cgan.py:191
    assert np.isfinite(hist['G_losses_epoch'][0]) and hist['perceptual_losses'][0] > 0
    sd = torch.load(d + '/models/gan/generator_final.pth')
    assert sd['fc.weight'].shape == (256 * 49, 100) and sd['fc.weight'].device.type == 'cpu'
