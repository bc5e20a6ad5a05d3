"""Shape-specialised CUDA-core kernels (thin image-side convs, latent GEMM, 7x7 GEMV) selected by ALGO_AUTO,
checked against the generic SIMT kernels (ALGO_SIMT) and the numpy oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

import dcgan_oracle as orc
import gan_enhanced_pneumonia_classifier_b200 as pkg
from parity_utils import close

L = pkg._lib
pytestmark = pytest.mark.gpu
AUTO, SIMT = L.Conv(4, 2, 1, L.ALGO_AUTO), L.Conv(4, 2, 1, L.ALGO_SIMT)


def st():
    return L.stream_ptr()


def rnd(shape, seed, dtype=torch.float32, scale=1.0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return (torch.randn(shape, device='cuda', generator=g) * scale).to(dtype)


@pytest.mark.parametrize('nc', [1, 3])
@pytest.mark.parametrize('fine_kind', ['nchw_f32', 'nhwc_bf16'])
def test_thin_layers_match_generic(nc, fine_kind):
    n, h = 3, 16                                   # fine 32x32, coarse 16x16 (W % 16 == 0: the warp-MMA kernels)
    w = rnd((32, nc, 4, 4), 1, scale=0.1)
    if fine_kind == 'nchw_f32':
        fine = rnd((n, nc, 2 * h, 2 * h), 2)
        fview = L.view_nchw
        fine_out = lambda: torch.empty_like(fine)
    else:
        fine = rnd((n, 2 * h, 2 * h, nc), 2, torch.bfloat16)
        fview = L.view_nhwc
        fine_out = lambda: torch.empty_like(fine)
    coarse = rnd((n, h, h, 32), 3, torch.bfloat16)
    # down: conv2d fprop (D0) -- fine -> coarse
    ya, yb = torch.empty_like(coarse), torch.empty_like(coarse)
    L.call('b200gan_conv2d_fprop', C.byref(AUTO), C.byref(fview(fine)), L.ptr(w), None, C.byref(L.view_nhwc(ya)), None, st())
    L.call('b200gan_conv2d_fprop', C.byref(SIMT), C.byref(fview(fine)), L.ptr(w), None, C.byref(L.view_nhwc(yb)), None, st())
    close(ya.float().cpu().numpy(), yb.float().cpu().numpy(), rtol=1e-2, atol=1e-2, what='thin down')
    # up: conv2d dgrad (D0 dgrad / G5 fprop) -- coarse -> fine
    da, db = fine_out(), fine_out()
    L.call('b200gan_conv2d_dgrad', C.byref(AUTO), C.byref(L.view_nhwc(coarse)), L.ptr(w), None, C.byref(fview(da)), None, st())
    L.call('b200gan_conv2d_dgrad', C.byref(SIMT), C.byref(L.view_nhwc(coarse)), L.ptr(w), None, C.byref(fview(db)), None, st())
    close(da.float().cpu().numpy(), db.float().cpu().numpy(), rtol=1e-2, atol=2e-2, what='thin up')
    # wgrad
    base = rnd((32, nc, 4, 4), 4)
    wa, wb = base.clone(), base.clone()
    L.call('b200gan_conv2d_wgrad', C.byref(AUTO), C.byref(fview(fine)), C.byref(L.view_nhwc(coarse)), L.ptr(wa), None, None, st())
    L.call('b200gan_conv2d_wgrad', C.byref(SIMT), C.byref(fview(fine)), C.byref(L.view_nhwc(coarse)), L.ptr(wb), None, None, st())
    ref = (wb - base).cpu().numpy()
    # the warp-MMA kernel stages the image as bf16 (the SIMT reference keeps fp32 image values): bf16 tolerance
    close((wa - base).cpu().numpy(), ref, rtol=5e-3, atol=5e-3 * np.abs(ref).max(), what='thin wgrad')


def test_thin_down_against_oracle():
    n, h, nc = 2, 16, 3
    w = rnd((32, nc, 4, 4), 1, scale=0.1)
    fine = rnd((n, nc, 2 * h, 2 * h), 2)
    y = torch.empty((n, h, h, 32), device='cuda', dtype=torch.bfloat16)
    L.call('b200gan_conv2d_fprop', C.byref(AUTO), C.byref(L.view_nchw(fine)), L.ptr(w), None, C.byref(L.view_nhwc(y)), None, st())
    ref = orc.conv2d_fprop(fine.cpu().numpy(), w.cpu().numpy(), 2, 1)
    close(y.float().cpu().numpy().transpose(0, 3, 1, 2), ref, rtol=1e-2, atol=1e-2, what='thin down vs oracle')


def test_window_gemv_matches_generic():
    n, c, k = 5, 64, 7
    cva, cvs = L.Conv(k, 1, 0, L.ALGO_AUTO), L.Conv(k, 1, 0, L.ALGO_SIMT)
    x = rnd((n, k, k, c), 1, torch.bfloat16)
    w = rnd((1, c, k, k), 2, scale=0.05)
    ya, yb = torch.empty((n, 1, 1, 1), device='cuda'), torch.empty((n, 1, 1, 1), device='cuda')
    L.call('b200gan_conv2d_fprop', C.byref(cva), C.byref(L.view_nhwc(x)), L.ptr(w), None, C.byref(L.view_nhwc(ya)), None, st())
    L.call('b200gan_conv2d_fprop', C.byref(cvs), C.byref(L.view_nhwc(x)), L.ptr(w), None, C.byref(L.view_nhwc(yb)), None, st())
    close(ya.cpu().numpy(), yb.cpu().numpy(), rtol=1e-4, atol=1e-4, what='window fprop')
    dl = rnd((n, 1, 1, 1), 3)
    da, db = torch.empty_like(x), torch.empty_like(x)
    L.call('b200gan_conv2d_dgrad', C.byref(cva), C.byref(L.view_nhwc(dl)), L.ptr(w), None, C.byref(L.view_nhwc(da)), None, st())
    L.call('b200gan_conv2d_dgrad', C.byref(cvs), C.byref(L.view_nhwc(dl)), L.ptr(w), None, C.byref(L.view_nhwc(db)), None, st())
    close(da.float().cpu().numpy(), db.float().cpu().numpy(), rtol=1e-2, atol=1e-3, what='window dgrad')
    base = rnd((1, c, k, k), 4)
    wa, wb = base.clone(), base.clone()
    L.call('b200gan_conv2d_wgrad', C.byref(cva), C.byref(L.view_nhwc(x)), C.byref(L.view_nhwc(dl)), L.ptr(wa), None, None, st())
    L.call('b200gan_conv2d_wgrad', C.byref(cvs), C.byref(L.view_nhwc(x)), C.byref(L.view_nhwc(dl)), L.ptr(wb), None, None, st())
    close((wa - base).cpu().numpy(), (wb - base).cpu().numpy(), rtol=1e-3, atol=1e-3, what='window wgrad')


def test_latent_gemm_matches_generic():
    n, nz, c, k = 70, 100, 24, 7
    cva, cvs = L.Conv(k, 1, 0, L.ALGO_AUTO), L.Conv(k, 1, 0, L.ALGO_SIMT)
    z = rnd((n, nz, 1, 1), 1)
    w = rnd((nz, c, k, k), 2, scale=0.05)
    ya = torch.empty((n, k, k, c), device='cuda', dtype=torch.bfloat16)
    yb = torch.empty_like(ya)
    L.call('b200gan_convT2d_fprop', C.byref(cva), C.byref(L.view_nchw(z)), L.ptr(w), None, C.byref(L.view_nhwc(ya)), None, st())
    L.call('b200gan_convT2d_fprop', C.byref(cvs), C.byref(L.view_nchw(z)), L.ptr(w), None, C.byref(L.view_nhwc(yb)), None, st())
    close(ya.float().cpu().numpy(), yb.float().cpu().numpy(), rtol=1e-2, atol=1e-2, what='latent fprop')
    dy = rnd((n, k, k, c), 3, torch.bfloat16)
    base = rnd((nz, c, k, k), 4)
    wa, wb = base.clone(), base.clone()
    L.call('b200gan_convT2d_wgrad', C.byref(cva), C.byref(L.view_nchw(z)), C.byref(L.view_nhwc(dy)), L.ptr(wa), None, None, st())
    L.call('b200gan_convT2d_wgrad', C.byref(cvs), C.byref(L.view_nchw(z)), C.byref(L.view_nhwc(dy)), L.ptr(wb), None, None, st())
    ref = (wb - base).cpu().numpy()
    # the tensor-core kernel rounds z to bf16 (2^-9 relative per product, like every tensor-core operand of the step) ...
    close((wa - base).cpu().numpy(), ref, rtol=1e-2, atol=1e-2 * np.abs(ref).max(), what='latent wgrad')
    # ... and is exact (fp32 accumulation order aside) against the same GEMMs on the rounded operands
    zb, wr = z.reshape(n, nz).to(torch.bfloat16).float(), w.to(torch.bfloat16).float().reshape(nz, c, k * k)
    y_ref = torch.einsum('nk,kcs->nsc', zb, wr).reshape(n, k, k, c)
    close(ya.float().cpu().numpy(), y_ref.cpu().numpy(), rtol=8e-3, atol=1e-5, what='latent fprop vs rounded-operand GEMM')
    dw_ref = torch.einsum('nk,nsc->kcs', zb, dy.float().reshape(n, k * k, c)).reshape(nz, c, k, k).cpu().numpy()
    close((wa - base).cpu().numpy(), dw_ref, rtol=1e-4, atol=1e-4 * np.abs(dw_ref).max(), what='latent wgrad vs rounded-operand GEMM')


@pytest.mark.parametrize('n,nz,c', [(512, 100, 512), (3, 16, 8), (65, 128, 16)])
def test_latent_gemm_shapes(n, nz, c):
    """Full-size G0 (B=512, nz=100, C=512) and the edges: batch not a multiple of 64, nz at 16 and at the 128 limit."""
    k = 7
    cv = L.Conv(k, 1, 0, L.ALGO_AUTO)
    z = rnd((n, nz, 1, 1), 5)
    w = rnd((nz, c, k, k), 6, scale=0.05)
    y = torch.empty((n, k, k, c), device='cuda', dtype=torch.bfloat16)
    L.call('b200gan_convT2d_fprop', C.byref(cv), C.byref(L.view_nchw(z)), L.ptr(w), None, C.byref(L.view_nhwc(y)), None, st())
    zb, wr = z.reshape(n, nz).to(torch.bfloat16).float(), w.to(torch.bfloat16).float().reshape(nz, c, k * k)
    y_ref = torch.einsum('nk,kcs->nsc', zb, wr).reshape(n, k, k, c)
    close(y.float().cpu().numpy(), y_ref.cpu().numpy(), rtol=8e-3, atol=1e-5, what='latent fprop')
    dy = rnd((n, k, k, c), 7, torch.bfloat16)
    base = rnd((nz, c, k, k), 8)
    dw = base.clone()
    L.call('b200gan_convT2d_wgrad', C.byref(cv), C.byref(L.view_nchw(z)), C.byref(L.view_nhwc(dy)), L.ptr(dw), None, None, st())
    dw_ref = torch.einsum('nk,nsc->kcs', zb, dy.float().reshape(n, k * k, c)).reshape(nz, c, k, k).cpu().numpy()
    close((dw - base).cpu().numpy(), dw_ref, rtol=2e-4, atol=2e-4 * np.abs(dw_ref).max(), what='latent wgrad')


@pytest.mark.parametrize('nc,c,fine_kind', [(1, 64, 'nchw_f32'), (3, 64, 'nchw_f32'), (1, 64, 'nhwc_bf16'), (3, 16, 'nhwc_bf16'), (2, 128, 'nchw_f32')])
def test_edge_layers_match_oracle_and_generic(nc, c, fine_kind):
    """The image-side layers whose feature side is not 32 channels wide (conv_edge.cu: the WGAN-GP critic's Conv2d(nc -> 64) and
    generator's ConvTranspose2d(64 -> nc)), forward / input gradient / weight gradient with the fused activations the training step
    uses, against the numpy oracle and the generic SIMT kernels."""
    n, h = 3, 12
    w = rnd((c, nc, 4, 4), 1, scale=0.1)
    if fine_kind == 'nchw_f32':
        fine = rnd((n, nc, 2 * h, 2 * h), 2)
        fview, to_nchw = L.view_nchw, lambda t: t.float().cpu().numpy()
    else:
        fine = rnd((n, 2 * h, 2 * h, nc), 2, torch.bfloat16)
        fview, to_nchw = L.view_nhwc, lambda t: t.float().cpu().numpy().transpose(0, 3, 1, 2)
    coarse = rnd((n, h, h, c), 3, torch.bfloat16)
    cn = coarse.float().cpu().numpy().transpose(0, 3, 1, 2)
    fnp, wnp = to_nchw(fine), w.cpu().numpy()
    # down + LeakyReLU on the way out (the critic's first layer, wggan.py:52-53)
    ya, yb = torch.empty_like(coarse), torch.empty_like(coarse)
    fz = L.fuse(out_act=L.ACT_LRELU, out_slope=0.2)
    L.call('b200gan_conv2d_fprop', C.byref(AUTO), C.byref(fview(fine)), L.ptr(w), None, C.byref(L.view_nhwc(ya)), C.byref(fz), st())
    L.call('b200gan_conv2d_fprop', C.byref(SIMT), C.byref(fview(fine)), L.ptr(w), None, C.byref(L.view_nhwc(yb)), C.byref(fz), st())
    ref = orc.conv2d_fprop(fnp, wnp, 2, 1)
    ref = np.where(ref > 0, ref, 0.2 * ref)
    close(ya.float().cpu().numpy().transpose(0, 3, 1, 2), ref, rtol=2.0 ** -8 * 1.05, atol=1e-5, what='edge down + LeakyReLU vs oracle')
    close(ya.float().cpu().numpy(), yb.float().cpu().numpy(), rtol=2.0 ** -7 * 1.05, atol=1e-5, what='edge down vs generic (one bf16 ulp)')
    # up + Tanh on the way out (the generator's last layer, wggan.py:41-42)
    da, db = torch.empty_like(fine), torch.empty_like(fine)
    ft = L.fuse(out_act=L.ACT_TANH)
    L.call('b200gan_convT2d_fprop', C.byref(AUTO), C.byref(L.view_nhwc(coarse)), L.ptr(w), None, C.byref(fview(da)), C.byref(ft), st())
    L.call('b200gan_convT2d_fprop', C.byref(SIMT), C.byref(L.view_nhwc(coarse)), L.ptr(w), None, C.byref(fview(db)), C.byref(ft), st())
    ref = np.tanh(orc.conv2d_dgrad(cn, wnp, 2, 1, (2 * h, 2 * h)))
    tol = dict(rtol=1e-5, atol=1e-5) if fine_kind == 'nchw_f32' else dict(rtol=2.0 ** -8 * 1.05, atol=1e-5)
    close(to_nchw(da), ref, what='edge up + Tanh vs oracle', **tol)
    close(to_nchw(da), to_nchw(db), what='edge up vs generic', **tol)
    # up with the LeakyReLU backward on the gradient operand (the critic's first-layer input gradient: dy * act'(a0))
    a0 = rnd((n, h, h, c), 5, torch.bfloat16)
    fm = L.fuse(dy_act=L.ACT_LRELU, dy_slope=0.2, dy_ref=L.view_nhwc(a0))
    L.call('b200gan_conv2d_dgrad', C.byref(AUTO), C.byref(L.view_nhwc(coarse)), L.ptr(w), None, C.byref(fview(da)), C.byref(fm), st())
    mask = np.where(a0.float().cpu().numpy().transpose(0, 3, 1, 2) > 0, 1.0, 0.2).astype(np.float32)
    ref = orc.conv2d_dgrad(cn * mask, wnp, 2, 1, (2 * h, 2 * h))
    close(to_nchw(da), ref, what='edge up with activation backward vs oracle', **(dict(rtol=1e-5, atol=1e-5) if fine_kind == 'nchw_f32' else dict(rtol=2.0 ** -8 * 1.05, atol=1e-4)))
    # weight gradient, plain and with the mask on the gradient operand
    for fuse, g in ((None, cn), (fm, cn * mask)):
        base = rnd((c, nc, 4, 4), 4)
        dw = base.clone()
        L.call('b200gan_conv2d_wgrad', C.byref(AUTO), C.byref(fview(fine)), C.byref(L.view_nhwc(coarse)), L.ptr(dw), None, C.byref(fuse) if fuse is not None else None, st())
        ref = orc.conv2d_wgrad(fnp, g, 4, 2, 1)
        close((dw - base).cpu().numpy(), ref, rtol=1e-4, atol=2e-5 * max(1.0, np.abs(ref).max()), what='edge wgrad vs oracle')
