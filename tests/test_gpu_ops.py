"""Per-operator parity of the C-ABI kernels (called through ctypes exactly as the product does) against the
numpy oracle, on the GPU.  fp32 storage: rtol 1e-4 (north star); bf16 storage: rtol 2e-2 against the fp32 oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

import dcgan_oracle as orc
import gan_enhanced_pneumonia_classifier_b200 as pkg
from parity_utils import close, grad_close

L = pkg._lib
pytestmark = pytest.mark.gpu

DT = {'fp32': torch.float32, 'bf16': torch.bfloat16}
TOL = {'fp32': dict(rtol=1e-4, atol=1e-5), 'bf16': dict(rtol=2e-2, atol=2e-2)}


def dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda().to(dtype)


def nhwc(a, dtype):
    """numpy NCHW -> device NHWC tensor of the given storage dtype."""
    return dev(a.transpose(0, 2, 3, 1), dtype)


def to_nchw(t):
    return t.float().cpu().numpy().transpose(0, 3, 1, 2)


def st():
    return L.stream_ptr()


CONV_CASES = [
    # (N, Ci, H, W, Co, k, s, p)
    (2, 3, 16, 16, 8, 4, 2, 1),       # D0-like (nc=3)
    (3, 8, 12, 20, 16, 4, 2, 1),      # middle layer, non-square
    (2, 1, 18, 18, 5, 4, 2, 1),       # nc=1, odd channel count
    (4, 16, 7, 7, 1, 7, 1, 0),        # D5-like: 7x7 valid conv to one logit
    (2, 100, 7, 7, 24, 7, 1, 0),      # G0 geometry seen from the conv side (Co = latent, Ci = ngf*8)
    (1, 70, 10, 10, 67, 4, 2, 1),     # ragged tiles (not multiples of 64/16)
]


@pytest.mark.parametrize('case', CONV_CASES)
@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_conv2d_fprop_dgrad_wgrad(case, mode):
    n, ci, h, w, co, k, s, p = case
    rng = np.random.RandomState(hash(case) % 2 ** 31)
    x = rng.randn(n, ci, h, w).astype(np.float32)
    wt = (rng.randn(co, ci, k, k) * 0.1).astype(np.float32)
    oh, ow = orc.conv_out_size(h, k, s, p), orc.conv_out_size(w, k, s, p)
    dy = rng.randn(n, co, oh, ow).astype(np.float32)
    dt = DT[mode]
    if mode == 'bf16':      # the oracle sees the same bf16-rounded inputs
        x = dev(x, dt).float().cpu().numpy()
        dy = dev(dy, dt).float().cpu().numpy()
    cv = L.Conv(k, s, p, L.ALGO_SIMT)
    xd, dyd, wd = nhwc(x, dt), nhwc(dy, dt), dev(wt)
    # fprop
    y = torch.empty((n, oh, ow, co), device='cuda', dtype=dt)
    L.call('b200gan_conv2d_fprop', C.byref(cv), C.byref(L.view_nhwc(xd)), L.ptr(wd), None, C.byref(L.view_nhwc(y)), None, st())
    close(to_nchw(y), orc.conv2d_fprop(x, wt, s, p), what='fprop', **TOL[mode])
    # dgrad
    dx = torch.empty((n, h, w, ci), device='cuda', dtype=dt)
    L.call('b200gan_conv2d_dgrad', C.byref(cv), C.byref(L.view_nhwc(dyd)), L.ptr(wd), None, C.byref(L.view_nhwc(dx)), None, st())
    close(to_nchw(dx), orc.conv2d_dgrad(dy, wt, s, p, (h, w)), what='dgrad', **TOL[mode])
    # wgrad accumulates: start from a non-zero buffer
    base = rng.randn(co, ci, k, k).astype(np.float32)
    dw = dev(base)
    L.call('b200gan_conv2d_wgrad', C.byref(cv), C.byref(L.view_nhwc(xd)), C.byref(L.view_nhwc(dyd)), L.ptr(dw), None, None, st())
    ref = orc.conv2d_wgrad(x, dy, k, s, p)
    close(dw.cpu().numpy() - base, ref, rtol=1e-4, atol=1e-4 * max(1.0, np.abs(ref).max()), what='wgrad')


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_convT2d_on_reference_nchw_views(mode):
    """ConvTranspose2d (weight (Cin,Cout,k,k)) reading / writing the reference's NCHW fp32 tensors in place."""
    rng = np.random.RandomState(5)
    n, cin, h, cout, k, s, p = 2, 12, 7, 6, 4, 2, 1
    x = rng.randn(n, cin, h, h).astype(np.float32)
    wt = (rng.randn(cin, cout, k, k) * 0.1).astype(np.float32)
    oh = orc.convT2d_out_size(h, k, s, p)
    dy = rng.randn(n, cout, oh, oh).astype(np.float32)
    cv = L.Conv(k, s, p, L.ALGO_SIMT)
    xd, dyd, wd = dev(x), dev(dy), dev(wt)                     # NCHW fp32, as the reference holds them
    y = torch.empty((n, cout, oh, oh), device='cuda')
    L.call('b200gan_convT2d_fprop', C.byref(cv), C.byref(L.view_nchw(xd)), L.ptr(wd), None, C.byref(L.view_nchw(y)), None, st())
    close(y.cpu().numpy(), orc.convT2d_fprop(x, wt, s, p), what='convT fprop')
    dx = torch.empty_like(xd)
    L.call('b200gan_convT2d_dgrad', C.byref(cv), C.byref(L.view_nchw(dyd)), L.ptr(wd), None, C.byref(L.view_nchw(dx)), None, st())
    close(dx.cpu().numpy(), orc.convT2d_dgrad(dy, wt, s, p), what='convT dgrad')
    dw = torch.zeros_like(wd)
    L.call('b200gan_convT2d_wgrad', C.byref(cv), C.byref(L.view_nchw(xd)), C.byref(L.view_nchw(dyd)), L.ptr(dw), None, None, st())
    close(dw.cpu().numpy(), orc.convT2d_wgrad(x, dy, k, s, p), rtol=1e-4, atol=1e-4, what='convT wgrad')
    # G0: latent (N,nz,1,1) -> (N,C,7,7)
    z = rng.randn(3, 10, 1, 1).astype(np.float32)
    w0 = (rng.randn(10, 9, 7, 7) * 0.1).astype(np.float32)
    cv0 = L.Conv(7, 1, 0, L.ALGO_SIMT)
    y0 = torch.empty((3, 9, 7, 7), device='cuda')
    zd, w0d = dev(z), dev(w0)          # keep the device tensors alive across the asynchronous launch
    L.call('b200gan_convT2d_fprop', C.byref(cv0), C.byref(L.view_nchw(zd)), L.ptr(w0d), None, C.byref(L.view_nchw(y0)), None, st())
    close(y0.cpu().numpy(), orc.convT2d_fprop(z, w0, 1, 0), what='G0 fprop')


def test_conv_shape_errors_are_loud():
    cv = L.Conv(4, 2, 1, L.ALGO_SIMT)
    x = torch.zeros((1, 8, 8, 4), device='cuda')
    y = torch.zeros((1, 5, 4, 4), device='cuda')            # wrong OH
    w = torch.zeros((4, 4, 4, 4), device='cuda')
    with pytest.raises(L.B200GanError, match='shapes do not match'):
        L.call('b200gan_conv2d_fprop', C.byref(cv), C.byref(L.view_nhwc(x)), L.ptr(w), None, C.byref(L.view_nhwc(y)), None, st())
    cvt = L.Conv(4, 2, 1, L.ALGO_TCGEN05)                   # forced tensor-core path on a shape it cannot take
    y2 = torch.zeros((1, 4, 4, 4), device='cuda')
    with pytest.raises(L.B200GanError):
        L.call('b200gan_conv2d_fprop', C.byref(cvt), C.byref(L.view_nhwc(x)), L.ptr(w), None, C.byref(L.view_nhwc(y2)), None, st())


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('act', [L.ACT_RELU, L.ACT_LRELU])
@pytest.mark.parametrize('shape', [(3, 8, 7, 7), (2, 32, 14, 10), (2, 300, 3, 5)])
def test_batchnorm_act_forward_backward(mode, act, shape):
    rng = np.random.RandomState(11)
    n, c, h, w = shape
    dt = DT[mode]
    y = (rng.randn(*shape) * 1.7 + 0.3).astype(np.float32)
    da = rng.randn(*shape).astype(np.float32)
    if mode == 'bf16':
        y, da = dev(y, dt).float().cpu().numpy(), dev(da, dt).float().cpu().numpy()
    gamma = rng.normal(1, 0.2, c).astype(np.float32)
    beta = rng.normal(0, 0.2, c).astype(np.float32)
    rm, rv, nbt = rng.randn(c).astype(np.float32), (rng.rand(c) + 0.5).astype(np.float32), np.array(3, np.int64)
    rm0, rv0 = rm.copy(), rv.copy()
    # oracle
    out_o, xhat, invstd = orc.bn_train_fwd(y, gamma, beta, rm, rv, nbt)
    a_o = np.maximum(out_o, 0) if act == L.ACT_RELU else np.where(out_o > 0, out_o, 0.2 * out_o)
    dz = da * ((out_o > 0) if act == L.ACT_RELU else np.where(out_o > 0, 1.0, 0.2)).astype(np.float32)
    dy_o, dg_o, db_o = orc.bn_train_bwd(dz, xhat, gamma, invstd)
    # device
    yd, dad = nhwc(y, dt), nhwc(da, dt)
    g, b, rmd, rvd = dev(gamma), dev(beta), dev(rm0), dev(rv0)
    nbtd = torch.tensor(3, device='cuda', dtype=torch.int64)
    sums = torch.empty(2 * c, device='cuda', dtype=torch.float64)
    scale, shift, mean, istd = (torch.empty(c, device='cuda') for _ in range(4))
    L.call('b200gan_bn_stats', C.byref(L.view_nhwc(yd)), L.ptr(sums), st())
    L.call('b200gan_bn_finalize', L.ptr(sums), c, n * h * w, L.ptr(g), L.ptr(b), L.ptr(rmd), L.ptr(rvd), L.ptr(nbtd), 0.1, 1e-5,
           L.ptr(scale), L.ptr(shift), L.ptr(mean), L.ptr(istd), st())
    a = torch.empty_like(yd)
    L.call('b200gan_bn_act_fwd', C.byref(L.view_nhwc(yd)), L.ptr(scale), L.ptr(shift), act, 0.2, C.byref(L.view_nhwc(a)), st())
    assert int(nbtd) == 4
    close(rmd.cpu().numpy(), rm, what='running_mean')
    close(rvd.cpu().numpy(), rv, what='running_var')
    close(istd.cpu().numpy(), invstd, what='invstd')
    close(to_nchw(a), a_o, what='bn+act', **TOL[mode])
    dgam, dbet = torch.full((c,), 0.5, device='cuda'), torch.full((c,), -0.25, device='cuda')     # accumulate semantics
    dyd = torch.empty_like(yd)
    L.call('b200gan_bn_act_bwd_reduce', C.byref(L.view_nhwc(dad)), C.byref(L.view_nhwc(yd)), None, L.ptr(scale), L.ptr(shift),
           L.ptr(mean), L.ptr(istd), act, 0.2, L.ptr(sums), st())
    L.call('b200gan_bn_act_bwd_apply', C.byref(L.view_nhwc(dad)), C.byref(L.view_nhwc(yd)), None, L.ptr(scale), L.ptr(shift),
           L.ptr(mean), L.ptr(istd), L.ptr(g), L.ptr(sums), n * h * w, act, 0.2, C.byref(L.view_nhwc(dyd)), L.ptr(dgam), L.ptr(dbet), st())
    grad_close(dgam.cpu().numpy() - 0.5, dg_o, 'dgamma', bulk=1e-4, l2=2e-3)
    grad_close(dbet.cpu().numpy() + 0.25, db_o, 'dbeta', bulk=1e-4, l2=2e-3)
    if mode == 'fp32':
        grad_close(to_nchw(dyd), dy_o, 'bn dgrad', bulk=1e-5, l2=1e-3)
    else:
        grad_close(to_nchw(dyd), dy_o, 'bn dgrad', bulk=5e-3, l2=2e-2, worst=5e-2)


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('shape', [(3, 8, 7, 7), (2, 64, 14, 10), (2, 1024, 3, 5), (2, 300, 3, 5)])
def test_bn_finalize_act_fwd_one_launch_is_bit_identical_to_the_two_passes(mode, shape):
    """b200gan_bn_finalize_act_fwd (one launch: coefficients + running statistics + normalise + activation) against b200gan_bn_finalize
    followed by b200gan_bn_act_fwd: the same arithmetic, so every output is BIT-identical; 300 channels (not a multiple of the vector
    width in bf16) take the library's two-pass route inside the call."""
    rng = np.random.RandomState(5)
    n, c, h, w = shape
    dt = DT[mode]
    yd = nhwc((rng.randn(*shape) * 1.3 + 0.2).astype(np.float32), dt)
    gamma, beta = dev(rng.normal(1, 0.2, c).astype(np.float32)), dev(rng.normal(0, 0.2, c).astype(np.float32))
    rm0, rv0 = rng.randn(c).astype(np.float32), (rng.rand(c) + 0.5).astype(np.float32)
    sums = torch.empty(2 * c, device='cuda', dtype=torch.float64)
    L.call('b200gan_bn_stats', C.byref(L.view_nhwc(yd)), L.ptr(sums), st())
    outs = []
    for fused in (False, True):
        rmd, rvd, nbtd = dev(rm0), dev(rv0), torch.tensor(7, device='cuda', dtype=torch.int64)
        scale, shift, mean, istd = (torch.full((c,), float('nan'), device='cuda') for _ in range(4))
        a = torch.full_like(yd, float('nan'))
        if fused:
            L.call('b200gan_bn_finalize_act_fwd', L.ptr(sums), c, n * h * w, L.ptr(gamma), L.ptr(beta), L.ptr(rmd), L.ptr(rvd), L.ptr(nbtd), 0.1, 1e-5,
                   L.ptr(scale), L.ptr(shift), L.ptr(mean), L.ptr(istd), C.byref(L.view_nhwc(yd)), L.ACT_LRELU, 0.2, C.byref(L.view_nhwc(a)), st())
        else:
            L.call('b200gan_bn_finalize', L.ptr(sums), c, n * h * w, L.ptr(gamma), L.ptr(beta), L.ptr(rmd), L.ptr(rvd), L.ptr(nbtd), 0.1, 1e-5,
                   L.ptr(scale), L.ptr(shift), L.ptr(mean), L.ptr(istd), st())
            L.call('b200gan_bn_act_fwd', C.byref(L.view_nhwc(yd)), L.ptr(scale), L.ptr(shift), L.ACT_LRELU, 0.2, C.byref(L.view_nhwc(a)), st())
        torch.cuda.synchronize()
        assert int(nbtd) == 8
        outs.append([t.clone() for t in (a, scale, shift, mean, istd, rmd, rvd)])
    for name, t0, t1 in zip(('a', 'scale', 'shift', 'mean', 'invstd', 'running_mean', 'running_var'), *outs):
        assert not torch.isnan(t1.float()).any(), name
        assert torch.equal(t0, t1), f'{name}: one-launch result differs from the two passes'


def test_bn_eval_tanh_sigmoid_and_copy():
    rng = np.random.RandomState(2)
    c = 6
    y = rng.randn(2, c, 5, 5).astype(np.float32)
    gamma, beta = rng.normal(1, .1, c).astype(np.float32), rng.normal(0, .1, c).astype(np.float32)
    rm, rv = rng.randn(c).astype(np.float32), (rng.rand(c) + .5).astype(np.float32)
    scale, shift = torch.empty(c, device='cuda'), torch.empty(c, device='cuda')
    gd, bd, rmd, rvd = dev(gamma), dev(beta), dev(rm), dev(rv)
    L.call('b200gan_bn_eval_coeffs', c, L.ptr(gd), L.ptr(bd), L.ptr(rmd), L.ptr(rvd), 1e-5, L.ptr(scale), L.ptr(shift), st())
    yd = dev(y)                                                 # NCHW view in, NHWC bf16 out (mixed dtypes + layouts)
    a = torch.empty((2, 5, 5, c), device='cuda', dtype=torch.bfloat16)
    L.call('b200gan_bn_act_fwd', C.byref(L.view_nchw(yd)), L.ptr(scale), L.ptr(shift), L.ACT_TANH, 0.0, C.byref(L.view_nhwc(a)), st())
    close(to_nchw(a), np.tanh(orc.bn_eval_fwd(y, gamma, beta, rm, rv)), rtol=1e-2, atol=1e-2, what='eval+tanh')
    out = torch.empty_like(yd)
    L.call('b200gan_bn_act_fwd', C.byref(L.view_nchw(yd)), None, None, L.ACT_SIGMOID, 0.0, C.byref(L.view_nchw(out)), st())
    close(out.cpu().numpy(), orc.sigmoid(y), what='sigmoid')
    cp = torch.empty((2, 5, 5, c), device='cuda')
    L.call('b200gan_copy_view', C.byref(L.view_nchw(yd)), C.byref(L.view_nhwc(cp)), st())
    assert np.array_equal(to_nchw(cp), y)


@pytest.mark.parametrize('target', [0.9, 0.0])
def test_bce_sigmoid_including_saturation(target):
    # logits that saturate fp32 sigmoid to exactly 1.0 / 0.0 exercise the -100 clamp and the 1e-12 divisor
    logit = np.array([-120., -90., -30., -3., -0.1, 0., 0.2, 4., 16., 17., 40., 1e-3, -1e-3], np.float32)
    p = orc.sigmoid(logit)
    assert p[0] == 0.0 and p[-3] == 1.0
    ld = dev(logit)
    b = logit.size
    prob, out2, dl = torch.empty(b, device='cuda'), torch.empty(2, device='cuda'), torch.empty(b, device='cuda')
    L.call('b200gan_bce_sigmoid', L.ptr(ld), b, target, 1.0, L.ptr(prob), L.ptr(out2), L.ptr(dl), st())
    close(prob.cpu().numpy(), p, rtol=1e-5, atol=1e-30, what='prob')
    close(out2.cpu().numpy()[0], orc.bce_fwd(p, target), rtol=1e-5, what='loss')
    close(out2.cpu().numpy()[1], p.mean(), rtol=1e-5, what='mean prob')
    close(dl.cpu().numpy(), orc.bce_bwd(p, target) * ((1 - p) * p), rtol=1e-4, atol=1e-9, what='dlogit')


@pytest.mark.parametrize('numel', [1, 7, 4096 + 3])
def test_adam_matches_torch_formula(numel):
    rng = np.random.RandomState(numel)
    p0, m, v = rng.randn(numel).astype(np.float32), np.zeros(numel, np.float32), np.zeros(numel, np.float32)
    pd, md, vd = dev(p0), dev(m), dev(v)
    po = p0.copy()
    for step in (1, 2, 3, 10):
        g = (rng.randn(numel) * 10.0 ** rng.randint(-6, 1, numel)).astype(np.float32)
        orc.adam_step(po, g, m, v, step, 2e-4, 0.5)
        gdev = dev(g)
        L.call('b200gan_adam', L.ptr(pd), L.ptr(gdev), L.ptr(md), L.ptr(vd), numel, 2e-4, 0.5, 0.999, 1e-8, step, None, 1.0, st())
        close(pd.cpu().numpy(), po, rtol=1e-6, atol=1e-7, what=f'param step {step}')
        close(md.cpu().numpy(), m, rtol=1e-5, atol=1e-12, what='exp_avg')
        close(vd.cpu().numpy(), v, rtol=1e-5, atol=1e-20, what='exp_avg_sq')
