"""b200gan_fuse: every fusion a convolution call can carry (activation on the way out, activation backward on the gradient
operand, BatchNorm statistics in the forward epilogue, activation backward + BatchNorm-backward sums in the input-gradient
epilogue) must give the same numbers as the unfused chain of the numpy oracle -- whichever kernel serves the call
(warp-MMA image-side kernels, tcgen05 implicit GEMM, SIMT)."""
import ctypes as C

import numpy as np
import pytest
import torch

import dcgan_oracle as orc
import gan_enhanced_pneumonia_classifier_b200 as pkg
from parity_utils import close

L = pkg._lib
pytestmark = pytest.mark.gpu
SLOPE = 0.2


def st():
    return L.stream_ptr()


def rnd(shape, seed, scale=1.0):
    return (np.random.RandomState(seed).randn(*shape) * scale).astype(np.float32)


def bf16_round(a):
    return torch.from_numpy(a).to(torch.bfloat16).float().numpy()


def dev_nhwc(a, dtype):
    """numpy NCHW -> device tensor stored NHWC in `dtype`, with its view."""
    t = torch.from_numpy(np.ascontiguousarray(a.transpose(0, 2, 3, 1))).cuda().to(dtype)
    return t, L.view_nhwc(t)


def dev_nchw(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t, L.view_nchw(t)


def back_nchw(t, nhwc):
    a = t.float().cpu().numpy()
    return a.transpose(0, 3, 1, 2) if nhwc else a


def lrelu(x):
    return np.where(x > 0, x, SLOPE * x).astype(np.float32)


def conv(algo=L.ALGO_AUTO):
    return L.Conv(4, 2, 1, algo)


# ---------------------------------------------------------------------------------------------------
# image-side layers: D0 = Conv2d(nc->32)+LeakyReLU, G5 = ConvTranspose2d(32->nc)+Tanh, forward and backward
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('nc', [1, 3])
@pytest.mark.parametrize('fine_kind', ['nchw_f32', 'nhwc_bf16'])
@pytest.mark.parametrize('algo', ['auto', 'simt'])
def test_image_side_layers_fused(nc, fine_kind, algo):
    n, hc, wc = 3, 12, 32                         # coarse 12x32 (W % 16 == 0: the warp-MMA kernels), fine 24x64
    cv = conv(L.ALGO_AUTO if algo == 'auto' else L.ALGO_SIMT)
    w = rnd((32, nc, 4, 4), 1, 0.1)               # conv geometry (Co=32, Ci=nc); as a ConvTranspose2d weight it is (Cin_T=32, Cout_T=nc)
    wd = torch.from_numpy(w).cuda()
    fine = rnd((n, nc, 2 * hc, 2 * wc), 2)
    coarse = bf16_round(rnd((n, 32, hc, wc), 3))
    nhwc_fine = fine_kind == 'nhwc_bf16'
    if nhwc_fine:
        fine = bf16_round(fine)
        mk_fine = lambda a: dev_nhwc(a, torch.bfloat16)
    else:
        mk_fine = dev_nchw
    tol = dict(rtol=2e-2, atol=2e-2)

    # --- D0 forward: a0 = LeakyReLU(conv(x))  (dcgan.py:65-66)
    x_t, x_v = mk_fine(fine)
    a0_t = torch.empty((n, hc, wc, 32), device='cuda', dtype=torch.bfloat16)
    f = L.fuse(out_act=L.ACT_LRELU, out_slope=SLOPE)
    L.call('b200gan_conv2d_fprop', C.byref(cv), C.byref(x_v), L.ptr(wd), None, C.byref(L.view_nhwc(a0_t)), C.byref(f), st())
    a0_ref = lrelu(orc.conv2d_fprop(fine, w, 2, 1))
    close(back_nchw(a0_t, True), a0_ref, what='D0 fprop + LeakyReLU', **tol)

    # --- D0 backward: dy0 = da0 * LeakyReLU'(a0) formed on the fly in wgrad and dgrad
    a0 = back_nchw(a0_t, True)                    # the saved (bf16) activation, as the kernels see it
    da0 = coarse
    da0_t, da0_v = dev_nhwc(da0, torch.bfloat16)
    a0_v = L.view_nhwc(a0_t)
    dy0 = bf16_round(da0 * np.where(a0 > 0, 1.0, SLOPE).astype(np.float32))
    f = L.fuse(dy_act=L.ACT_LRELU, dy_slope=SLOPE, dy_ref=a0_v)
    base = rnd((32, nc, 4, 4), 4)
    dw_t = torch.from_numpy(base.copy()).cuda()
    L.call('b200gan_conv2d_wgrad', C.byref(cv), C.byref(x_v), C.byref(da0_v), L.ptr(dw_t), None, C.byref(f), st())
    dw_ref = orc.conv2d_wgrad(fine, dy0, 4, 2, 1)
    close(dw_t.cpu().numpy() - base, dw_ref, rtol=1e-2, atol=1e-2 * np.abs(dw_ref).max(), what='D0 wgrad with fused LeakyReLU backward')
    dx_t, dx_v = mk_fine(np.zeros_like(fine))
    L.call('b200gan_conv2d_dgrad', C.byref(cv), C.byref(da0_v), L.ptr(wd), None, C.byref(dx_v), C.byref(f), st())
    dx_ref = orc.conv2d_dgrad(dy0, w, 2, 1, (2 * hc, 2 * wc))
    close(back_nchw(dx_t, nhwc_fine), dx_ref, what='D0 dgrad with fused LeakyReLU backward', **tol)

    # --- G5 forward: fake = Tanh(convT(a4))  (dcgan.py:46-47); a4 plays the coarse side
    a4 = coarse
    a4_t, a4_v = dev_nhwc(a4, torch.bfloat16)
    fake_t, fake_v = mk_fine(np.zeros_like(fine))
    f = L.fuse(out_act=L.ACT_TANH)
    L.call('b200gan_convT2d_fprop', C.byref(cv), C.byref(a4_v), L.ptr(wd), None, C.byref(fake_v), C.byref(f), st())
    fake_ref = np.tanh(orc.convT2d_fprop(a4, w, 2, 1))
    close(back_nchw(fake_t, nhwc_fine), fake_ref, what='G5 fprop + Tanh', **tol)

    # --- G5 backward: dy5 = dfake * (1 - fake^2) formed on the fly
    fake = back_nchw(fake_t, nhwc_fine)
    dfake = fine                                     # any gradient tensor of the image shape
    dfake_t, dfake_v = mk_fine(dfake)
    dy5 = dfake * (1 - fake * fake)
    f = L.fuse(dy_act=L.ACT_TANH, dy_ref=fake_v)
    dw_t = torch.from_numpy(base.copy()).cuda()
    L.call('b200gan_convT2d_wgrad', C.byref(cv), C.byref(a4_v), C.byref(dfake_v), L.ptr(dw_t), None, C.byref(f), st())
    dw_ref = orc.convT2d_wgrad(a4, dy5, 4, 2, 1)
    close(dw_t.cpu().numpy() - base, dw_ref, rtol=1e-2, atol=1e-2 * np.abs(dw_ref).max(), what='G5 wgrad with fused Tanh backward')
    da4_t = torch.empty((n, hc, wc, 32), device='cuda', dtype=torch.bfloat16)
    L.call('b200gan_convT2d_dgrad', C.byref(cv), C.byref(dfake_v), L.ptr(wd), None, C.byref(L.view_nhwc(da4_t)), C.byref(f), st())
    da4_ref = orc.convT2d_dgrad(dy5, w, 2, 1)
    close(back_nchw(da4_t, True), da4_ref, what='G5 dgrad with fused Tanh backward', **tol)


def test_image_side_full_size_row_tiles():
    """The benchmark geometry (224x224 image, 112x112x32 coarse, partial last tile impossible: 112 = 14 x 8) on a small batch,
    plus a coarse height that is NOT a multiple of the 8-row tile (ragged last tile)."""
    for hc, wc in ((112, 112), (13, 16)):
        n, nc = 2, 1
        w = rnd((32, nc, 4, 4), 5, 0.1)
        wd = torch.from_numpy(w).cuda()
        fine = rnd((n, nc, 2 * hc, 2 * wc), 6)
        x_t, x_v = dev_nchw(fine)
        y_t = torch.empty((n, hc, wc, 32), device='cuda', dtype=torch.bfloat16)
        cv = conv()
        L.call('b200gan_conv2d_fprop', C.byref(cv), C.byref(x_v), L.ptr(wd), None, C.byref(L.view_nhwc(y_t)), None, st())
        ref = orc.conv2d_fprop(bf16_round(fine), bf16_round(w), 2, 1)
        close(back_nchw(y_t, True), ref, rtol=1e-2, atol=1e-2, what=f'thin down {hc}x{wc}')
        c = bf16_round(rnd((n, 32, hc, wc), 7))
        c_t, c_v = dev_nhwc(c, torch.bfloat16)
        dx_t, dx_v = dev_nchw(np.zeros_like(fine))
        L.call('b200gan_conv2d_dgrad', C.byref(cv), C.byref(c_v), L.ptr(wd), None, C.byref(dx_v), None, st())
        close(dx_t.cpu().numpy(), orc.conv2d_dgrad(c, bf16_round(w), 2, 1, (2 * hc, 2 * wc)), rtol=1e-2, atol=2e-2, what=f'thin up {hc}x{wc}')
        dw_t = torch.zeros((32, nc, 4, 4), device='cuda')
        L.call('b200gan_conv2d_wgrad', C.byref(cv), C.byref(x_v), C.byref(c_v), L.ptr(dw_t), None, None, st())
        dw_ref = orc.conv2d_wgrad(bf16_round(fine), c, 4, 2, 1)
        close(dw_t.cpu().numpy(), dw_ref, rtol=5e-3, atol=5e-3 * np.abs(dw_ref).max(), what=f'thin wgrad {hc}x{wc}')


# ---------------------------------------------------------------------------------------------------
# BatchNorm fusions of the middle layers: statistics in the forward epilogue, activation backward + BN-backward sums in the
# input-gradient epilogue.  mode fp32 -> SIMT kernels, bf16 -> tcgen05 kernels (channels % 32 == 0).
# ---------------------------------------------------------------------------------------------------
# geometry (n, ci, co, hf, wf): fine (n,ci,hf,wf), coarse (n,co,hf/2,wf/2).  The second case is the thin 32<->64-channel layer at a size
# the halo-tile tcgen05 kernels take (conv_down4_tc_kernel<1>, <2>, conv_up4_tc_kernel<1>), with ragged tiles in both directions
# (24 = 16 + 8 output rows, 20 = 8 + 8 + 4 output columns); the last two are the 64<->128-channel layer: its "up" direction runs on
# conv_up4w_tc_kernel<1> (ConvTranspose2d forward + statistics) and <2> (Conv2d input gradient + BatchNorm backward), ragged and full tiles
@pytest.mark.parametrize('geom', [(4, 32, 64, 16, 16), (3, 32, 64, 48, 40), (3, 64, 128, 48, 40), (5, 64, 128, 56, 56)])
@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('transposed', [False, True])
def test_batchnorm_epilogue_fusions(mode, transposed, geom):
    dt = torch.float32 if mode == 'fp32' else torch.bfloat16
    rt = (lambda a: a) if mode == 'fp32' else bf16_round
    tol = dict(rtol=1e-4, atol=1e-5) if mode == 'fp32' else dict(rtol=2e-2, atol=2e-2)
    n, ci, co, hf, wf = geom
    cv = conv()
    w = rnd((co, ci, 4, 4), 1, 0.05)
    wd = torch.from_numpy(w).cuda()
    wp = [torch.empty(w.size, device='cuda', dtype=torch.bfloat16) for _ in range(2)]
    for form in (0, 1):
        L.call('b200gan_pack_conv_weight', L.ptr(wd), co, ci, 4, form, L.ptr(wp[form]), st())
    wq = rt(w) if mode == 'bf16' else w
    fine = rt(rnd((n, ci, hf, wf), 2))
    coarse = rt(rnd((n, co, hf // 2, wf // 2), 3))
    fine_t, fine_v = dev_nhwc(fine, dt)
    coarse_t, coarse_v = dev_nhwc(coarse, dt)

    # ---- forward with statistics: Conv2d: fine -> coarse; ConvTranspose2d: coarse -> fine
    if not transposed:
        src_v, res_shape, cres, name, wpk = fine_v, (n, hf // 2, wf // 2, co), co, 'b200gan_conv2d_fprop', wp[0]
        ref = orc.conv2d_fprop(fine, wq, 2, 1)
    else:
        src_v, res_shape, cres, name, wpk = coarse_v, (n, hf, wf, ci), ci, 'b200gan_convT2d_fprop', wp[1]
        ref = orc.convT2d_fprop(coarse, wq, 2, 1)
    y_t = torch.empty(res_shape, device='cuda', dtype=dt)
    sums = torch.full((2 * cres,), 123.0, device='cuda', dtype=torch.float64)       # must be OVERWRITTEN
    f = L.fuse(bn_sums=sums)
    L.call(name, C.byref(cv), C.byref(src_v), L.ptr(wd), L.ptr(wpk), C.byref(L.view_nhwc(y_t)), C.byref(f), st())
    y = back_nchw(y_t, True)
    close(y, ref, what='conv with bn_sums: result', **tol)
    s = sums.cpu().numpy()
    cnt = y.shape[0] * y.shape[2] * y.shape[3]
    # statistics are those of the STORED tensor (what the normalisation pass will read)
    close(s[:cres] / cnt, y.astype(np.float64).mean(axis=(0, 2, 3)), rtol=1e-4, atol=1e-5, what='bn_sums: mean')
    close(s[cres:] / cnt, (y.astype(np.float64) ** 2).mean(axis=(0, 2, 3)), rtol=1e-4, atol=1e-6, what='bn_sums: mean of squares')

    # ---- input gradient with the previous layer's activation + BatchNorm backward in the epilogue
    if not transposed:      # Conv2d dgrad: coarse -> fine; previous layer produced a_prev of the fine shape
        g_v, res_np, cprev, name, wpk = coarse_v, fine, ci, 'b200gan_conv2d_dgrad', wp[1]
        dx_ref = orc.conv2d_dgrad(coarse, wq, 2, 1, (hf, wf))
    else:                   # ConvTranspose2d dgrad: fine -> coarse
        g_v, res_np, cprev, name, wpk = fine_v, coarse, co, 'b200gan_convT2d_dgrad', wp[0]
        dx_ref = orc.convT2d_dgrad(fine, wq, 2, 1)
    y_prev = rt(rnd(res_np.shape, 7))
    yp_t, yp_v = dev_nhwc(y_prev, dt)
    scale, shift = rnd((cprev,), 8, 0.5) + 1.0, rnd((cprev,), 9, 0.3)
    mean, invstd = rnd((cprev,), 10, 0.2), np.abs(rnd((cprev,), 11)) + 0.5
    dvec = [torch.from_numpy(v).cuda() for v in (scale, shift, mean, invstd)]
    for act, act_np in ((L.ACT_LRELU, lambda z: np.where(z > 0, 1.0, SLOPE)), (L.ACT_RELU, lambda z: np.where(z > 0, 1.0, 0.0))):
        dz_t = torch.empty(tuple(np.array(res_np.shape)[[0, 2, 3, 1]]), device='cuda', dtype=dt)
        psums = torch.full((2 * cprev,), -7.0, device='cuda', dtype=torch.float64)
        f = L.fuse(prev_act=act, prev_slope=SLOPE, prev_y=yp_v, prev_scale=dvec[0], prev_shift=dvec[1], prev_mean=dvec[2], prev_invstd=dvec[3],
                   prev_sums=psums)
        L.call(name, C.byref(cv), C.byref(g_v), L.ptr(wd), L.ptr(wpk), C.byref(L.view_nhwc(dz_t)), C.byref(f), st())
        z = y_prev * scale[None, :, None, None] + shift[None, :, None, None]
        dz_ref = dx_ref * act_np(z)
        dz = back_nchw(dz_t, True)
        # elements whose pre-activation is within rounding of the kink may take the other branch: exclude |z| < 1e-3
        safe = np.abs(z) > 1e-3
        close(np.where(safe, dz, 0), np.where(safe, dz_ref, 0), what='dgrad with prev_*: dz', **tol)
        xhat = (y_prev - mean[None, :, None, None]) * invstd[None, :, None, None]
        s = psums.cpu().numpy()
        s0, s1 = dz_ref.astype(np.float64).sum(axis=(0, 2, 3)), (dz_ref.astype(np.float64) * xhat).sum(axis=(0, 2, 3))
        mag = np.abs(dz_ref).sum(axis=(0, 2, 3))
        btol = 1e-4 if mode == 'fp32' else 1e-2
        assert np.all(np.abs(s[:cprev] - s0) <= btol * mag + 1e-4), 'prev_sums: sum dz'
        assert np.all(np.abs(s[cprev:] - s1) <= btol * (np.abs(dz_ref * xhat).sum(axis=(0, 2, 3))) + 1e-4), 'prev_sums: sum dz*xhat'


def test_fuse_argument_checking():
    cv = conv(L.ALGO_SIMT)
    x = torch.zeros((1, 8, 8, 4), device='cuda')
    y = torch.zeros((1, 4, 4, 4), device='cuda')
    w = torch.zeros((4, 4, 4, 4), device='cuda')
    with pytest.raises(L.B200GanError, match='do not apply to a forward'):
        f = L.fuse(dy_act=L.ACT_TANH, dy_ref=L.view_nhwc(y))
        L.call('b200gan_conv2d_fprop', C.byref(cv), C.byref(L.view_nhwc(x)), L.ptr(w), None, C.byref(L.view_nhwc(y)), C.byref(f), st())
    with pytest.raises(L.B200GanError, match='only apply to a forward'):
        f = L.fuse(out_act=L.ACT_TANH)
        L.call('b200gan_conv2d_dgrad', C.byref(cv), C.byref(L.view_nhwc(y)), L.ptr(w), None, C.byref(L.view_nhwc(x)), C.byref(f), st())
    with pytest.raises(L.B200GanError, match='extents differ'):
        f = L.fuse(dy_act=L.ACT_LRELU, dy_slope=0.2, dy_ref=L.view_nhwc(x))
        L.call('b200gan_conv2d_wgrad', C.byref(cv), C.byref(L.view_nhwc(x)), C.byref(L.view_nhwc(y)), L.ptr(w), None, C.byref(f), st())


@pytest.mark.parametrize('case', [('bf16', 64, 32, 28), ('bf16', 64, 64, 12), ('bf16', 128, 96, 8), ('fp32', 24, 10, 6)])
def test_dgrad_activation_only_epilogue(case):
    """prev_* with prev_scale == NULL: the layer below has no BatchNorm (D0: Conv2d + LeakyReLU, dcgan.py:65-66); prev_y is its saved
    activation output and the input-gradient convolution returns dx * act'(a_prev).  Cases: the fused 4-class "up" kernel
    (64 -> 32 channels), the generic tcgen05 kernel, and the fp32 SIMT kernels."""
    mode, co, ci, hc = case
    dt = torch.float32 if mode == 'fp32' else torch.bfloat16
    rt = (lambda a: a) if mode == 'fp32' else bf16_round
    tol = dict(rtol=1e-4, atol=1e-5) if mode == 'fp32' else dict(rtol=2e-2, atol=2e-2)
    n = 3
    w = rt(rnd((co, ci, 4, 4), 1, 0.05))
    wd = torch.from_numpy(w).cuda()
    wp = torch.empty(w.size, device='cuda', dtype=torch.bfloat16)
    L.call('b200gan_pack_conv_weight', L.ptr(wd), co, ci, 4, 1, L.ptr(wp), st())
    dy = rt(rnd((n, co, hc, hc), 2))
    a_prev = rt(rnd((n, ci, 2 * hc, 2 * hc), 3))
    dy_t, dy_v = dev_nhwc(dy, dt)
    ap_t, ap_v = dev_nhwc(a_prev, dt)
    dz_t = torch.full((n, 2 * hc, 2 * hc, ci), float('nan'), device='cuda', dtype=dt)
    cv = conv()
    f = L.fuse(prev_act=L.ACT_LRELU, prev_slope=SLOPE, prev_y=ap_v)
    L.call('b200gan_conv2d_dgrad', C.byref(cv), C.byref(dy_v), L.ptr(wd), L.ptr(wp), C.byref(L.view_nhwc(dz_t)), C.byref(f), st())
    ref = orc.conv2d_dgrad(dy, w, 2, 1, (2 * hc, 2 * hc)) * np.where(a_prev > 0, 1.0, SLOPE).astype(np.float32)
    close(back_nchw(dz_t, True), ref, what=f'dgrad + activation-only epilogue {case}', **tol)


@pytest.mark.parametrize('nc', [1, 3])
def test_image_side_down_kernel_batchnorm_epilogues(nc):
    """G5's input gradient (ConvTranspose2d(32->nc) dgrad, the image-side "down" kernel) with BOTH fusions the generator step uses
    at once: Tanh backward on the gradient operand (dy_act) and, in the epilogue, ReLU backward + BatchNorm-backward sums of the
    layer below (prev_*); plus the forward-statistics epilogue (bn_sums) of the same kernel."""
    n, hc, wc = 3, 12, 32
    cv = conv()
    w = rnd((32, nc, 4, 4), 1, 0.1)
    wd = torch.from_numpy(w).cuda()
    fake = bf16_round(np.tanh(rnd((n, nc, 2 * hc, 2 * wc), 2)))
    dfake = bf16_round(rnd((n, nc, 2 * hc, 2 * wc), 3))
    y4 = bf16_round(rnd((n, 32, hc, wc), 4))
    scale, shift = rnd((32,), 5, 0.5) + 1.0, rnd((32,), 6, 0.3)
    mean, invstd = rnd((32,), 7, 0.2), np.abs(rnd((32,), 8)) + 0.5
    dvec = [torch.from_numpy(v).cuda() for v in (scale, shift, mean, invstd)]
    fake_t, fake_v = dev_nhwc(fake, torch.bfloat16)
    dfake_t, dfake_v = dev_nhwc(dfake, torch.bfloat16)
    y4_t, y4_v = dev_nhwc(y4, torch.bfloat16)
    dz_t = torch.full((n, hc, wc, 32), float('nan'), device='cuda', dtype=torch.bfloat16)
    psums = torch.full((64,), -3.0, device='cuda', dtype=torch.float64)
    f = L.fuse(dy_act=L.ACT_TANH, dy_ref=fake_v, prev_act=L.ACT_RELU, prev_y=y4_v, prev_scale=dvec[0], prev_shift=dvec[1], prev_mean=dvec[2],
               prev_invstd=dvec[3], prev_sums=psums)
    L.call('b200gan_convT2d_dgrad', C.byref(cv), C.byref(dfake_v), L.ptr(wd), None, C.byref(L.view_nhwc(dz_t)), C.byref(f), st())
    dy5 = bf16_round(dfake * (1 - fake * fake))
    da4 = orc.convT2d_dgrad(dy5, bf16_round(w), 2, 1)
    z = y4 * scale[None, :, None, None] + shift[None, :, None, None]
    dz_ref = da4 * (z > 0)
    dz = back_nchw(dz_t, True)
    safe = np.abs(z) > 1e-3
    close(np.where(safe, dz, 0), np.where(safe, dz_ref, 0), rtol=2e-2, atol=2e-2, what='thin down + prev_*: dz')
    xhat = (y4 - mean[None, :, None, None]) * invstd[None, :, None, None]
    s = psums.cpu().numpy()
    # the kernel's sums are those of the stored dz
    close(s[:32], dz.astype(np.float64).sum(axis=(0, 2, 3)), rtol=1e-3, atol=1e-2, what='prev_sums: sum dz')
    close(s[32:], (dz.astype(np.float64) * xhat).sum(axis=(0, 2, 3)), rtol=1e-3, atol=2e-2, what='prev_sums: sum dz*xhat')
    # forward statistics epilogue of the same kernel (Conv2d(nc->32) forward with bn_sums)
    x_t, x_v = dev_nhwc(dfake, torch.bfloat16)
    y_t = torch.empty((n, hc, wc, 32), device='cuda', dtype=torch.bfloat16)
    sums = torch.full((64,), 9.0, device='cuda', dtype=torch.float64)
    f = L.fuse(bn_sums=sums)
    L.call('b200gan_conv2d_fprop', C.byref(cv), C.byref(x_v), L.ptr(wd), None, C.byref(L.view_nhwc(y_t)), C.byref(f), st())
    y = back_nchw(y_t, True).astype(np.float64)
    close(y, orc.conv2d_fprop(dfake, bf16_round(w), 2, 1), rtol=2e-2, atol=2e-2, what='thin down + bn_sums: result')
    close(sums.cpu().numpy()[:32], y.sum(axis=(0, 2, 3)), rtol=1e-4, atol=1e-2, what='bn_sums: sum')
    close(sums.cpu().numpy()[32:], (y ** 2).sum(axis=(0, 2, 3)), rtol=1e-4, atol=1e-2, what='bn_sums: sum of squares')


@pytest.mark.parametrize('n,c', [(37, 64), (512, 512), (5, 8)])
def test_full_window_dgrad_batchnorm_epilogue(n, c):
    """D5 = Conv2d(C->1, k7) on a 7x7 map (dcgan.py:84): its input gradient (an outer product) carries the LeakyReLU backward and the
    BatchNorm-backward sums of D4 in the same kernel.  Checked against the unfused chain in float64 from the same bf16 inputs."""
    k = 7
    cv = L.Conv(k, 1, 0, L.ALGO_AUTO)
    w = rnd((1, c, k, k), 1, 0.05)
    dl = rnd((n, 1, 1, 1), 2)
    y_prev = bf16_round(rnd((n, c, k, k), 3))
    scale, shift = rnd((c,), 8, 0.5) + 1.0, rnd((c,), 9, 0.3)
    mean, invstd = rnd((c,), 10, 0.2), np.abs(rnd((c,), 11)) + 0.5
    wd, dld = torch.from_numpy(w).cuda(), torch.from_numpy(dl).cuda()
    yp_t, yp_v = dev_nhwc(y_prev, torch.bfloat16)
    dvec = [torch.from_numpy(v).cuda() for v in (scale, shift, mean, invstd)]
    dz_t = torch.empty((n, k, k, c), device='cuda', dtype=torch.bfloat16)
    psums = torch.full((2 * c,), -7.0, device='cuda', dtype=torch.float64)          # must be OVERWRITTEN
    f = L.fuse(prev_act=L.ACT_LRELU, prev_slope=SLOPE, prev_y=yp_v, prev_scale=dvec[0], prev_shift=dvec[1], prev_mean=dvec[2], prev_invstd=dvec[3],
               prev_sums=psums)
    L.call('b200gan_conv2d_dgrad', C.byref(cv), C.byref(L.view_nchw(dld)), L.ptr(wd), None, C.byref(L.view_nhwc(dz_t)), C.byref(f), st())
    dx = dl.astype(np.float64) * w.astype(np.float64)                                 # (n,1,1,1) * (1,c,7,7)
    z = y_prev * scale[None, :, None, None] + shift[None, :, None, None]
    dz_ref = dx * np.where(z > 0, 1.0, SLOPE)
    dz = back_nchw(dz_t, True)
    safe = np.abs(z) > 1e-3
    close(np.where(safe, dz, 0), np.where(safe, dz_ref, 0), rtol=1e-2, atol=1e-6, what='k7 dgrad with prev_*: dz')
    # the sums are those of the STORED (bf16) dz
    xhat = (y_prev.astype(np.float64) - mean[None, :, None, None]) * invstd[None, :, None, None]
    s = psums.cpu().numpy()
    s0, s1 = dz.astype(np.float64).sum(axis=(0, 2, 3)), (dz.astype(np.float64) * xhat).sum(axis=(0, 2, 3))
    assert np.all(np.abs(s[:c] - s0) <= 1e-4 * np.abs(dz).sum(axis=(0, 2, 3)) + 1e-6), 'prev_sums: sum dz'
    assert np.all(np.abs(s[c:] - s1) <= 1e-4 * np.abs(dz * xhat).sum(axis=(0, 2, 3)) + 1e-6), 'prev_sums: sum dz*xhat'

