"""The VGG16 perceptual loss of the conditional GAN on the GPU (reference src/train_cgan.py:57-73,186): the 3x3
convolutions folded onto the stride-2 4x4 kernels, the fused bias / ReLU / depth-to-space passes, max pooling, the drop-in `PerceptualLoss` and its
place in `CGANTrainer` -- against the numpy oracle (oracle/vgg_oracle.py, pinned to torchvision's vgg16 in tests/test_oracle_golden.py) and against
torchvision's own module on the CPU.  Weights are random: the ImageNet checkpoint the reference downloads cannot be obtained offline, so what is pinned
is the operator on arbitrary weights."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

import cgan_oracle as co
import dcgan_oracle as orc
import vgg_oracle as vo
from conftest import GOLDEN
from gan_enhanced_pneumonia_classifier_b200 import _lib as L
from gan_enhanced_pneumonia_classifier_b200.engine import Act
from gan_enhanced_pneumonia_classifier_b200.perceptual import PerceptualLoss
from parity_utils import close, grad_close, synthetic_real

pytestmark = pytest.mark.gpu


def st():
    return L.stream_ptr()


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def nhwc(a, dtype=torch.float32):
    return dev(a).permute(0, 2, 3, 1).contiguous().to(dtype)


def nchw(t):
    return t.float().permute(0, 3, 1, 2).cpu().numpy()


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def random_vgg(seed):
    sd = vo.init_weights(np.random.RandomState(seed))
    mod = PerceptualLoss('random')
    feats = torch.nn.Sequential(*[m for blk in mod.blocks for m in blk])
    feats.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return mod, sd


@pytest.mark.parametrize('ci,co_,h,pad', [(3, 8, 8, 3), (3, 64, 16, 32), (64, 64, 16, 64), (128, 256, 8, 128)])
def test_folded_3x3_convolution_equals_the_oracle(ci, co_, h, pad):
    """conv3x3_fold + Conv2d(4,2,1) + bias_relu_d2s == relu(conv3x3(x) + bias); relu_bwd_s2d + the k4 input gradient == the 3x3 input gradient."""
    rng = np.random.RandomState(ci + co_)
    x = rng.randn(2, ci, h, h).astype(np.float32)
    w3 = (rng.randn(co_, ci, 3, 3) * np.sqrt(2.0 / (9 * ci))).astype(np.float32)
    bias = (rng.randn(co_) * 0.1).astype(np.float32)
    da = rng.randn(2, co_, h, h).astype(np.float32)
    a_ref = np.maximum(orc.conv2d_fprop(x, w3, 1, 1) + bias[None, :, None, None], 0)
    dx_ref = orc.conv2d_dgrad(da * (a_ref > 0), w3, 1, 1, (h, h))
    w3_d, bias_d = dev(w3), dev(bias)
    w4 = torch.empty((4 * co_, pad, 4, 4), device='cuda')
    L.call('b200gan_conv3x3_fold', L.ptr(w3_d), co_, ci, pad, L.ptr(w4), st())
    xp = torch.zeros((2, h, h, pad), device='cuda')
    xp[..., :ci] = nhwc(x)
    conv = L.Conv(4, 2, 1, L.ALGO_AUTO)
    t = torch.empty((2, h // 2, h // 2, 4 * co_), device='cuda')
    L.call('b200gan_conv2d_fprop', C.byref(conv), C.byref(Act(xp, nchw=False).v), L.ptr(w4), None, C.byref(Act(t, nchw=False).v), None, st())
    a = torch.empty((2, h, h, co_), device='cuda')
    L.call('b200gan_bias_relu_d2s', C.byref(Act(t, nchw=False).v), L.ptr(bias_d), C.byref(Act(a, nchw=False).v), st())
    close(nchw(a), a_ref, rtol=1e-4, atol=1e-5, what='relu(conv3x3 + bias) through the folded stride-2 convolution')
    dt = torch.empty_like(t)
    da_d = nhwc(da)
    L.call('b200gan_relu_bwd_s2d', C.byref(Act(da_d, nchw=False).v), C.byref(Act(a, nchw=False).v), C.byref(Act(dt, nchw=False).v), st())
    dx = torch.empty_like(xp)
    L.call('b200gan_conv2d_dgrad', C.byref(conv), C.byref(Act(dt, nchw=False).v), L.ptr(w4), None, C.byref(Act(dx, nchw=False).v), None, st())
    close(nchw(dx[..., :ci]), dx_ref, rtol=1e-4, atol=1e-4, what='input gradient through the folded convolution')
    if pad > ci:
        assert float(dx[..., ci:].abs().max()) == 0.0


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_max_pooling_forward_and_first_maximum_backward(dtype):
    rng = np.random.RandomState(9)
    a = np.round(rng.randn(2, 16, 8, 8) * 2).astype(np.float32) / 2          # coarse values: plenty of ties inside the windows
    a[0, :, :2, :2] = 0.0                                                   # an all-equal window
    dp = rng.randn(2, 16, 4, 4).astype(np.float32)
    a_d, dp_d = nhwc(a, dtype), nhwc(dp, dtype)
    p = torch.empty((2, 4, 4, 16), device='cuda', dtype=dtype)
    L.call('b200gan_maxpool2_fwd', C.byref(Act(a_d, nchw=False).v), C.byref(Act(p, nchw=False).v), st())
    np.testing.assert_array_equal(nchw(p), vo.maxpool2(a))
    for add in (0, 1):
        da = torch.full((2, 8, 8, 16), 0.25, device='cuda', dtype=dtype)
        L.call('b200gan_maxpool2_bwd', C.byref(Act(a_d, nchw=False).v), C.byref(Act(dp_d, nchw=False).v), C.byref(Act(da, nchw=False).v), add, st())
        want = vo.maxpool2_bwd(a, nchw(dp_d)) + (0.25 if add else 0.0)
        close(nchw(da), want, rtol=8e-3 if dtype == torch.bfloat16 else 1e-6, atol=1e-6, what=f'max-pool backward (add={add})')


def test_perceptual_loss_fp32_matches_the_oracle():
    mod, sd = random_vgg(21)
    mod = mod.cuda()
    mod.compute_dtype = torch.float32
    rng = np.random.RandomState(22)
    x, y = synthetic_real(31, 2, 3, size=32), synthetic_real(32, 2, 3, size=32)
    xt = dev(x).requires_grad_(True)
    loss = mod(xt, dev(y))
    (loss * 3.0).backward()
    loss_ref, dx_ref = vo.perceptual(x, y, sd)
    close(loss.item(), loss_ref, rtol=1e-4, what='perceptual loss')
    grad_close(xt.grad.cpu().numpy() / 3.0, dx_ref, 'd perceptual / d x', bulk=2e-5, l2=2e-3, worst=2e-2)
    with torch.no_grad():
        close(mod(dev(x), dev(y)).item(), loss_ref, rtol=1e-4, what='perceptual loss without autograd')


@pytest.mark.parametrize('dtype,loss_tol,grad_tol', [(torch.float32, 1e-4, 5e-3), (torch.bfloat16, 3e-2, 0.15)])
def test_perceptual_loss_full_size_matches_torchvision_on_the_cpu(dtype, loss_tol, grad_tol):
    """224x224, the size train_cgan.py runs it at, against torchvision's own module on the CPU (same weights).  bf16: the 64..256-channel layers on
    the tcgen05 kernels, the image stored 32 channels wide."""
    mod, _ = random_vgg(23)
    x, y = synthetic_real(41, 2, 3), synthetic_real(42, 2, 3) * 0.5
    xc = torch.from_numpy(x).requires_grad_(True)
    ref = mod(xc, torch.from_numpy(y))                                       # CPU tensors: the stock torch path (= the reference's PerceptualLoss.forward)
    ref.backward()
    gpu = PerceptualLoss('random')
    gpu.load_state_dict(mod.state_dict())
    gpu = gpu.cuda()
    gpu.compute_dtype = dtype
    xt = dev(x).requires_grad_(True)
    loss = gpu(xt, dev(y))
    loss.backward()
    close(loss.item(), ref.item(), rtol=loss_tol, what=f'perceptual loss ({dtype})')
    err = rel(xt.grad.cpu().numpy(), xc.grad.numpy())
    print(f'{dtype}: loss {loss.item():.6f} vs {ref.item():.6f}, relative L2 of the image gradient {err:.3e}')
    assert err < grad_tol, err


def test_trainer_with_the_perceptual_term_equals_the_reference_loop_over_the_modules():
    """CGANTrainer(perceptual=...) against train_cgan.py:150-193 restated with stock torch glue over the drop-in modules and the drop-in PerceptualLoss
    (fp32): the same kernels, different bookkeeping (the perceptual gradient is added to the Discriminator's inside the trainer)."""
    from gan_enhanced_pneumonia_classifier_b200.cgan_trainer import CGANTrainer
    from test_gpu_cgan import build, reference_iteration
    g = np.load(os.path.join(GOLDEN, 'cgan_step_nc3.npz'))
    m = json.loads(str(g['meta']))
    vgg, _ = random_vgg(25)
    vgg = vgg.cuda()
    vgg.compute_dtype = torch.float32
    G, D = build(g, m, torch.float32)
    G2, D2 = build(g, m, torch.float32)
    tr = CGANTrainer(G, D, lr=m['lr'], beta1=m['beta1'], perceptual=vgg, perceptual_weight=10.0, dtype=torch.float32)
    optD = torch.optim.Adam(D2.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    optG = torch.optim.Adam(G2.parameters(), lr=m['lr'], betas=(m['beta1'], 0.999))
    for it in range(m['iters']):
        real = dev(synthetic_real(m['seed'] + 10 + it, m['batch'], m['nc']))
        draws = [dev(g[f'it{it}.{k}']) for k in ('real_labels', 'smooth_real', 'smooth_fake', 'noise', 'fake_labels')]
        row = tr.step(real, draws[0], epoch=0, noise=draws[3], fake_labels=draws[4], smooth_real=draws[1], smooth_fake=draws[2]).cpu().numpy().astype(np.float64)
        r = reference_iteration(G2, D2, optG, optD, real, *draws, perceptual=vgg)
        assert row[5] > 0
        close(row[[0, 1, 2, 3, 4, 6]], r['row'], rtol=2e-5 if it == 0 else 2e-3, atol=1e-5 if it == 0 else 1e-3, what=f'trainer vs module loop, iteration {it}')
        if it == 0:
            grads = {k: p.grad.detach().cpu().numpy() for k, p in G2.named_parameters()}
            for (k, _), view in zip(G.named_parameters(), tr.arenaG.grads):
                if 'main.3.bias' in k or 'main.7.bias' in k or 'main.11.bias' in k or 'main.15.bias' in k:
                    continue                                               # conv biases in front of a BatchNorm: rounding noise (see test_gpu_cgan.py)
                grad_close(view.cpu().numpy(), grads[k], f'generator gradient {k} (adversarial + 10 perceptual + 5 feature matching)', bulk=2e-4, l2=2e-3, worst=2e-2)
