"""tcgen05 implicit-GEMM convolutions (forced with ALGO_TCGEN05, so a silent SIMT fallback cannot pass) against
the numpy oracle AND the fp32-accumulating SIMT kernels, on every layer shape of the DCGAN (small / ragged batches).

Oracle tolerance: operands are bf16-exact, the oracle accumulates in float64; the kernel accumulates in fp32 and rounds the
result to bf16 once, so it must sit within ONE bf16 ulp (2^-8 relative) of the oracle plus the fp32 accumulation error."""
import ctypes as C

import numpy as np
import pytest
import torch

import dcgan_oracle as orc
import gan_enhanced_pneumonia_classifier_b200 as pkg
from parity_utils import close

L = pkg._lib
pytestmark = pytest.mark.gpu


def st():
    return L.stream_ptr()


def bf16_exact(shape, scale, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return (torch.randn(shape, device='cuda', generator=g) * scale).to(torch.bfloat16)


def pack(w, form):
    co, ci, k, _ = w.shape
    out = torch.empty(co * ci * k * k, device='cuda', dtype=torch.bfloat16)
    L.call('b200gan_pack_conv_weight', L.ptr(w), co, ci, k, form, L.ptr(out), st())
    return out


def run_pair(n, ci, h, w_, co, direction):
    """direction 'down': y = conv2d(x) with x (n,h,w,ci) -> (n,h/2,w/2,co); 'up': dx = conv2d_dgrad(dy)."""
    cv_tc, cv_simt = L.Conv(4, 2, 1, L.ALGO_TCGEN05), L.Conv(4, 2, 1, L.ALGO_SIMT)
    wt = bf16_exact((co, ci, 4, 4), 0.05, 1).float().contiguous()          # bf16-representable fp32 master
    if direction == 'down':
        x = bf16_exact((n, h, w_, ci), 1.0, 2)
        y_tc = torch.full((n, h // 2, w_ // 2, co), float('nan'), device='cuda', dtype=torch.bfloat16)
        y_ref = torch.empty_like(y_tc)
        wp = pack(wt, 0)
        L.call('b200gan_conv2d_fprop', C.byref(cv_tc), C.byref(L.view_nhwc(x)), L.ptr(wt), L.ptr(wp), C.byref(L.view_nhwc(y_tc)), None, st())
        L.call('b200gan_conv2d_fprop', C.byref(cv_simt), C.byref(L.view_nhwc(x)), L.ptr(wt), None, C.byref(L.view_nhwc(y_ref)), None, st())
        torch.cuda.synchronize()
        return x, wt, y_tc, y_ref
    dy = bf16_exact((n, h, w_, co), 1.0, 3)                                  # coarse side (n,h,w,co) -> fine (n,2h,2w,ci)
    dx_tc = torch.full((n, 2 * h, 2 * w_, ci), float('nan'), device='cuda', dtype=torch.bfloat16)
    dx_ref = torch.empty_like(dx_tc)
    wp = pack(wt, 1)
    L.call('b200gan_conv2d_dgrad', C.byref(cv_tc), C.byref(L.view_nhwc(dy)), L.ptr(wt), L.ptr(wp), C.byref(L.view_nhwc(dx_tc)), None, st())
    L.call('b200gan_conv2d_dgrad', C.byref(cv_simt), C.byref(L.view_nhwc(dy)), L.ptr(wt), None, C.byref(L.view_nhwc(dx_ref)), None, st())
    torch.cuda.synchronize()
    return dy, wt, dx_tc, dx_ref


BF16_ULP = 2.0 ** -8


def oracle_close(got_nhwc, ref_nchw, what):
    """bf16 result against the float64-accumulated oracle: one bf16 ulp + fp32 accumulation noise."""
    ref = np.ascontiguousarray(ref_nchw.transpose(0, 2, 3, 1))
    close(got_nhwc.float().cpu().numpy(), ref, rtol=1.05 * BF16_ULP, atol=2e-5 * max(1.0, float(np.abs(ref).max())), what=what + ' vs numpy oracle')


def nchw(t):
    return t.float().cpu().numpy().transpose(0, 3, 1, 2)


def report(tag, got, ref):
    g, r = got.float(), ref.float()
    err = (g - r).abs().max().item()
    print(f'{tag}: max|diff|={err:.4e} ref max={r.abs().max().item():.3e} nan={torch.isnan(g).sum().item()}')
    return err


# (n, cin, h, w, cout): every k4s2p1 layer shape of the DCGAN at a small batch + ragged batches
DOWN = [(2, 64, 16, 16, 128), (3, 32, 112, 112, 64), (5, 64, 56, 56, 128), (9, 128, 28, 28, 256), (130, 256, 14, 14, 512),
        (2, 32, 8, 24, 32), (3, 64, 12, 20, 64)]
UP = [(2, 128, 8, 8, 64), (130, 512, 7, 7, 256), (9, 256, 14, 14, 128), (5, 128, 28, 28, 64), (3, 64, 56, 56, 32),
      (2, 32, 4, 12, 32), (3, 64, 6, 10, 64), (3, 128, 24, 20, 64), (150, 128, 12, 8, 64)]


@pytest.mark.parametrize('case', DOWN)
def test_tc_down_conv_matches_simt(case):
    n, ci, h, w_, co = case
    x, wt, y_tc, y_ref = run_pair(n, ci, h, w_, co, 'down')
    report(f'down {case}', y_tc, y_ref)
    assert not torch.isnan(y_tc.float()).any()
    close(y_tc.float().cpu().numpy(), y_ref.float().cpu().numpy(), rtol=1.6e-2, atol=2e-2, what=f'down {case}')
    oracle_close(y_tc, orc.conv2d_fprop(nchw(x), wt.cpu().numpy(), 2, 1), f'down {case}')


@pytest.mark.parametrize('case', UP)
def test_tc_up_conv_matches_simt(case):
    n, co, h, w_, ci = case          # dy has `co` channels, result has `ci`
    dy, wt, dx_tc, dx_ref = run_pair(n, ci, h, w_, co, 'up')
    report(f'up {case}', dx_tc, dx_ref)
    assert not torch.isnan(dx_tc.float()).any()
    close(dx_tc.float().cpu().numpy(), dx_ref.float().cpu().numpy(), rtol=1.6e-2, atol=2e-2, what=f'up {case}')
    oracle_close(dx_tc, orc.conv2d_dgrad(nchw(dy), wt.cpu().numpy(), 2, 1, (2 * h, 2 * w_)), f'up {case}')


@pytest.mark.parametrize('case', [(5, 128, 28, 28, 64), (3, 128, 24, 20, 64)])
def test_halo_wide_kernel_agrees_with_generic_kernel(case, monkeypatch):
    """conv_up4w_tc_kernel (128 -> 64 channels "up", conv_tc_halo_wide.cu) against the generic implicit-GEMM kernel on the same operands
    (B200GAN_NO_UP4W=1 is read per call): the same bf16 products accumulated in fp32 in a different order, so the two bf16 results are at
    most one ulp apart -- 2^-7 relative at the bottom of a binade; both are held to the numpy oracle (2^-8) by test_tc_up_conv_matches_simt."""
    n, co, h, w_, ci = case
    monkeypatch.setenv('B200GAN_NO_UP4W', '0')
    dy, wt, dx_new, _ = run_pair(n, ci, h, w_, co, 'up')
    monkeypatch.setenv('B200GAN_NO_UP4W', '1')
    _, _, dx_gen, _ = run_pair(n, ci, h, w_, co, 'up')
    assert not torch.isnan(dx_new.float()).any()
    close(dx_new.float().cpu().numpy(), dx_gen.float().cpu().numpy(), rtol=2.1 * BF16_ULP, atol=1e-4, what='halo-wide vs generic kernel')
    assert (dx_new != dx_gen).float().mean().item() < 0.05, 'more than 5 % of the elements differ by an ulp'
    oracle_close(dx_new, orc.conv2d_dgrad(nchw(dy), wt.cpu().numpy(), 2, 1, (2 * h, 2 * w_)), f'halo-wide up {case}')


def test_tc_against_numpy_oracle():
    x, wt, y_tc, _ = run_pair(2, 64, 16, 16, 128, 'down')
    ref = orc.conv2d_fprop(x.float().cpu().numpy().transpose(0, 3, 1, 2), wt.cpu().numpy(), 2, 1)
    close(y_tc.float().cpu().numpy().transpose(0, 3, 1, 2), ref, rtol=1e-2, atol=1e-2, what='down vs oracle')
    dy, wt, dx_tc, _ = run_pair(2, 64, 8, 8, 128, 'up')
    ref = orc.conv2d_dgrad(dy.float().cpu().numpy().transpose(0, 3, 1, 2), wt.cpu().numpy(), 2, 1, (16, 16))
    close(dx_tc.float().cpu().numpy().transpose(0, 3, 1, 2), ref, rtol=1e-2, atol=1e-2, what='up vs oracle')


@pytest.mark.parametrize('case,direction', [((9, 128, 28, 28, 256), 'down'), ((130, 256, 14, 14, 512), 'down'), ((5, 64, 28, 28, 256), 'down'),
                                            ((130, 512, 7, 7, 256), 'up')])
def test_cta_pair_kernel_matches_oracle(case, direction, monkeypatch):
    """The opt-in CTA-pair kernel (tcgen05.mma.cta_group::2, 256 x 256 tiles, conv_tc_pair.cu; B200GAN_PAIR=1) on the wide layers, odd
    M-tile counts included (the last pair's second tile is empty), against the numpy oracle and the one-CTA kernel."""
    monkeypatch.setenv('B200GAN_PAIR', '1')
    if direction == 'down':
        n, ci, h, w_, co = case
        x, wt, y_pair, _ = run_pair(n, ci, h, w_, co, 'down')
        oracle_close(y_pair, orc.conv2d_fprop(nchw(x), wt.cpu().numpy(), 2, 1), f'pair down {case}')
        monkeypatch.setenv('B200GAN_PAIR', '0')
        _, _, y_one, _ = run_pair(n, ci, h, w_, co, 'down')
    else:
        n, co, h, w_, ci = case
        dy, wt, y_pair, _ = run_pair(n, ci, h, w_, co, 'up')
        oracle_close(y_pair, orc.conv2d_dgrad(nchw(dy), wt.cpu().numpy(), 2, 1, (2 * h, 2 * w_)), f'pair up {case}')
        monkeypatch.setenv('B200GAN_PAIR', '0')
        _, _, y_one, _ = run_pair(n, ci, h, w_, co, 'up')
    # same operands, same fp32 accumulation order per output element up to the MMA's internal order: one bf16 ulp at most
    close(y_pair.float().cpu().numpy(), y_one.float().cpu().numpy(), rtol=1.05 * BF16_ULP, atol=1e-4, what='pair vs one-CTA kernel')


# (n, Ci, H, W, Co): x is (n,H,W,Ci), dy is (n,H/2,W/2,Co) -- the four tensor-core layer geometries + odd sizes
WGRAD = [(3, 32, 112, 112, 64), (5, 64, 56, 56, 128), (9, 128, 28, 28, 256), (70, 256, 14, 14, 512), (2, 64, 8, 24, 64),
         (3, 128, 12, 20, 192), (2, 512, 4, 4, 128)]


@pytest.mark.parametrize('use_ws', [False, True])
@pytest.mark.parametrize('case', WGRAD)
def test_tc_wgrad_matches_simt(case, use_ws):
    """use_ws: the split-K partial sums go through the caller's zeroed workspace (coalesced atomics + transpose), which must
    come back all zero; without it the kernel adds straight into dw."""
    n, ci, h, w_, co = case
    x = bf16_exact((n, h, w_, ci), 1.0, 4)
    dy = bf16_exact((n, h // 2, w_ // 2, co), 1.0, 5)
    base = torch.randn((co, ci, 4, 4), device='cuda')
    dw_tc, dw_ref = base.clone(), base.clone()
    cv_tc, cv_simt = L.Conv(4, 2, 1, L.ALGO_TCGEN05), L.Conv(4, 2, 1, L.ALGO_SIMT)
    ws = torch.zeros(co * ci * 16, device='cuda') if use_ws else None
    L.call('b200gan_conv2d_wgrad', C.byref(cv_tc), C.byref(L.view_nhwc(x)), C.byref(L.view_nhwc(dy)), L.ptr(dw_tc), L.ptr(ws), None, st())
    L.call('b200gan_conv2d_wgrad', C.byref(cv_simt), C.byref(L.view_nhwc(x)), C.byref(L.view_nhwc(dy)), L.ptr(dw_ref), None, None, st())
    torch.cuda.synchronize()
    if use_ws:
        assert float(ws.abs().max()) == 0.0, 'workspace must be handed back zeroed'
    a, b = (dw_tc - base).cpu().numpy(), (dw_ref - base).cpu().numpy()
    print(f'wgrad {case}: max|diff|={np.abs(a - b).max():.4e} ref max={np.abs(b).max():.3e}')
    close(a, b, rtol=2e-3, atol=2e-3 * max(1.0, np.abs(b).max()), what=f'wgrad {case}')
    # fp32 result (fp32 accumulation over up to n*OH*OW = 9408 terms, then `base + .` in fp32) against the float64 oracle
    ref = orc.conv2d_wgrad(nchw(x), nchw(dy), 4, 2, 1)
    close(a, ref, rtol=1e-4, atol=3e-5 * max(1.0, np.abs(ref).max()), what=f'wgrad {case} vs numpy oracle')


def test_tc_wgrad_against_numpy_oracle():
    n, ci, h, co = 2, 64, 16, 128
    x = bf16_exact((n, h, h, ci), 1.0, 6)
    dy = bf16_exact((n, h // 2, h // 2, co), 1.0, 7)
    dw = torch.zeros((co, ci, 4, 4), device='cuda')
    cv = L.Conv(4, 2, 1, L.ALGO_TCGEN05)
    L.call('b200gan_conv2d_wgrad', C.byref(cv), C.byref(L.view_nhwc(x)), C.byref(L.view_nhwc(dy)), L.ptr(dw), None, None, st())
    ref = orc.conv2d_wgrad(x.float().cpu().numpy().transpose(0, 3, 1, 2), dy.float().cpu().numpy().transpose(0, 3, 1, 2), 4, 2, 1)
    close(dw.cpu().numpy(), ref, rtol=1e-3, atol=1e-3, what='wgrad vs oracle')
