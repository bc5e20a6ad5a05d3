"""The WGAN-GP critic on the B200 kernels: layer plans for `wggan.Generator` / `wggan.Discriminator` and the gradient penalty with its
explicit double backward (reference /root/reference/src/wggan.py:15-89).

The Generator (wggan.py:18-42) and the critic (wggan.py:51-64) are the same kinds of layers as the DCGAN's, with other widths, so
`engine.NetEngine` runs their forward and first-order backward unchanged.  What is new is `gradient_penalty` (wggan.py:72-89): the
reference differentiates the critic's input gradient w.r.t. the critic's parameters with `torch.autograd.grad(create_graph=True)`;
here the three sweeps are launched explicitly (math and numpy restatement: oracle/wgan_oracle.py):

  1. forward of the critic on x^ (train mode: BatchNorm batch statistics, running buffers move), activations saved;
  2. first backward down to the input, g = d sum_b D(x^)_b / d x^, KEEPING per layer dz (after the LeakyReLU derivative), dy (after
     BatchNorm backward) and the BatchNorm-backward sums;  gp = lambda mean_b (||g_b|| - 1)^2;
  3. reverse sweep through (2), bottom to top, with u = d gp / d g = coeff_b g_b:  per layer  dW += wgrad(x = u, dy),
     r = conv(u, W) (a FORWARD convolution), then the second-order BatchNorm terms (`b200gan_bn_bwd_bwd`): the next u, the adjoint
     `inj` of the forward conv output, and the gamma adjoint;
  4. an ordinary backward of the forward graph that starts without a loss gradient and picks up `inj` at every BatchNorm layer.

Every convolution goes through the same C-ABI entry points as the DCGAN step (tcgen05 kernels for the k4 s2 p1 layers with channels
% 32 == 0; the 64-channel image-side layers and the 7x7 valid convolution on the 14x14 map run on the SIMT kernels for now).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib as L
from . import engine as E
from .engine import Act, LayerSpec

LAMBDA_GP = 10.0


def wgan_generator_specs(latent_dim: int, nc: int, ngf: int) -> List[LayerSpec]:
    """wggan.py:18-42: channels nz -> 16 ngf -> 8 -> 4 -> 2 -> ngf -> nc; Sequential indices as in the DCGAN generator."""
    ch = [latent_dim, ngf * 16, ngf * 8, ngf * 4, ngf * 2, ngf, nc]
    out = []
    for i in range(6):
        k, s, p = (7, 1, 0) if i == 0 else (4, 2, 1)
        out.append(LayerSpec(3 * i, 3 * i + 1 if i < 5 else None, ch[i], ch[i + 1], k, s, p, L.ACT_RELU if i < 5 else L.ACT_TANH))
    return out


def critic_specs(nc: int, ndf: int) -> List[LayerSpec]:
    """wggan.py:51-64: Conv(nc->ndf)-LReLU, 3 x [Conv-BN-LReLU] (ndf -> 2 -> 4 -> 8 ndf), Conv(8 ndf -> 1, k7 s1 p0) without activation."""
    ch = [nc, ndf, ndf * 2, ndf * 4, ndf * 8, 1]
    conv_idx = [0, 2, 5, 8, 11]
    out = []
    for i in range(5):
        k, s, p = (7, 1, 0) if i == 4 else (4, 2, 1)
        out.append(LayerSpec(conv_idx[i], conv_idx[i] + 1 if 1 <= i <= 3 else None, ch[i], ch[i + 1], k, s, p, L.ACT_LRELU if i < 4 else L.ACT_NONE))
    return out


def _empty_like(a: Act) -> Act:
    return Act(torch.empty_like(a.t), nchw=False)


def gradient_penalty(eng: E.NetEngine, params, xhat: torch.Tensor, grads, lambda_gp: float = LAMBDA_GP) -> torch.Tensor:
    """gp = lambda * mean_b (||d sum D(x^) / d x^_b|| - 1)^2 through the train-mode critic `eng`, and its gradient w.r.t. every critic
    parameter ACCUMULATED into `grads` (list over `eng.param_order()`: conv weight, then BatchNorm weight, bias per layer).
    `xhat`: (N, nc, H, W) float32 NCHW.  Returns the () float32 tensor gp (device, not synchronised)."""
    st = L.stream_ptr()
    dev = xhat.device
    specs = eng.specs
    nl = len(specs)
    n = xhat.shape[0]
    # ---- (1) forward, activations saved ---------------------------------------------------------------------------------------
    smap, ctxs = eng.forward(Act(xhat, nchw=True), params, True, True, last_act=False)
    hw = smap.t.shape[1] * smap.t.shape[2]
    # parameter index of each layer's conv weight in `grads`
    gidx, gi = [], 0
    for sp in specs:
        gidx.append(gi)
        gi += 3 if sp.bn_idx is not None else 1

    # ---- (2) first backward, keeping dz / dy / sums ----------------------------------------------------------------------------
    seed = Act(torch.full(smap.t.shape, 1.0 / hw, device=dev, dtype=torch.float32), nchw=False)     # d sum_b mean_hw(map) / d map
    dys, dzs = [None] * nl, [None] * nl
    dys[nl - 1] = seed
    g = torch.empty_like(xhat)
    for i in reversed(range(nl)):
        sp, p, lc = specs[i], params[i], ctxs[i]
        dy = dys[i]
        if i == 0:
            eng._dgrad(0, dy, p.w, Act(g, nchw=True), st, lc.wp_down, lc.wp_up)
            break
        below, lb = specs[i - 1], ctxs[i - 1]
        d = _empty_like(lb.a)
        if below.bn_idx is not None:
            lb.bsums = torch.empty(2 * below.cout, device=dev, dtype=torch.float64)
            fz = L.fuse(prev_act=below.act, prev_slope=E.LRELU_SLOPE, prev_y=lb.y.v, prev_scale=lb.scale, prev_shift=lb.shift, prev_mean=lb.mean,
                        prev_invstd=lb.invstd, prev_sums=lb.bsums)
            eng._dgrad(i, dy, p.w, d, st, lc.wp_down, lc.wp_up, fuse=fz)                 # d = dz of the layer below, sums in lb.bsums
            dzs[i - 1] = d
            dyb = _empty_like(d)
            cnt = lb.y.v.n * lb.y.v.h * lb.y.v.w
            L.call('b200gan_bn_act_bwd_apply', C.byref(d.v), C.byref(lb.y.v), None, L.ptr(lb.scale), L.ptr(lb.shift), L.ptr(lb.mean), L.ptr(lb.invstd),
                   L.ptr(params[i - 1].gamma), L.ptr(lb.bsums), cnt, L.ACT_NONE, E.LRELU_SLOPE, C.byref(dyb.v), None, None, st)
            eng.launches += 1
            dys[i - 1] = dyb
        else:
            fz = L.fuse(prev_act=below.act, prev_slope=E.LRELU_SLOPE, prev_y=lb.a.v)
            eng._dgrad(i, dy, p.w, d, st, lc.wp_down, lc.wp_up, fuse=fz)                 # LeakyReLU derivative from the saved output's sign
            dzs[i - 1] = dys[i - 1] = d
    # ---- penalty and d gp / d g ---------------------------------------------------------------------------------------------------
    sumsq = torch.empty(n, device=dev, dtype=torch.float64)
    gp = torch.empty(1, device=dev, dtype=torch.float32)
    coeff = torch.empty(n, device=dev, dtype=torch.float32)
    gv = L.view_nchw(g)
    L.call('b200gan_sample_sumsq', C.byref(gv), L.ptr(sumsq), st)
    L.call('b200gan_gp_from_norms', L.ptr(sumsq), n, lambda_gp, L.ptr(gp), L.ptr(coeff), st)
    u = Act(torch.empty_like(xhat), nchw=True)
    L.call('b200gan_sample_axpby', C.byref(gv), L.ptr(coeff), None, None, C.byref(u.v), st)
    eng.launches += 3

    # ---- (3) reverse sweep through the first backward ---------------------------------------------------------------------------
    injs = [None] * nl
    for i in range(nl):
        sp, p, lc = specs[i], params[i], ctxs[i]
        eng._wgrad(i, u, dys[i], grads[gidx[i]], st)                                     # dW_i += wgrad(x = u, dy = dy_i)
        if i == nl - 1:
            break
        r = Act(torch.empty(lc.a.t.shape, device=dev, dtype=lc.a.t.dtype), nchw=False)
        eng._fprop(i, u, p.w, r, st, lc.wp_down, lc.wp_up)                               # adjoint of dy_i: a forward convolution of u
        nxt = _empty_like(r)
        if sp.bn_idx is not None:
            inj = _empty_like(r)
            ws = torch.empty(3 * sp.cout, device=dev, dtype=torch.float64)
            cnt = lc.y.v.n * lc.y.v.h * lc.y.v.w
            L.call('b200gan_bn_bwd_bwd', C.byref(r.v), C.byref(lc.y.v), C.byref(dzs[i].v), L.ptr(lc.scale), L.ptr(lc.shift), L.ptr(lc.mean),
                   L.ptr(lc.invstd), L.ptr(p.gamma), L.ptr(lc.bsums), cnt, sp.act, E.LRELU_SLOPE, C.byref(nxt.v), C.byref(inj.v),
                   L.ptr(grads[gidx[i] + 1]), L.ptr(ws), st)
            eng.launches += 2
            injs[i] = inj
        else:
            L.call('b200gan_bn_act_bwd_apply', C.byref(r.v), C.byref(lc.a.v), C.byref(lc.a.v), None, None, None, None, None, None, 0, sp.act,
                   E.LRELU_SLOPE, C.byref(nxt.v), None, None, st)                        # u = r * lrelu'(.) from the saved output's sign
            eng.launches += 1
        u = nxt

    # ---- (4) ordinary backward of the forward graph, fed by the injected adjoints ------------------------------------------------
    d = None
    for i in reversed(range(nl - 1)):
        sp, p, lc = specs[i], params[i], ctxs[i]
        dy = None
        if d is not None:
            if sp.bn_idx is not None:
                cnt = lc.y.v.n * lc.y.v.h * lc.y.v.w
                L.call('b200gan_bn_act_bwd_apply', C.byref(d.v), C.byref(lc.y.v), None, L.ptr(lc.scale), L.ptr(lc.shift), L.ptr(lc.mean), L.ptr(lc.invstd),
                       L.ptr(p.gamma), L.ptr(lc.bsums2), cnt, L.ACT_NONE, E.LRELU_SLOPE, C.byref(d.v), L.ptr(grads[gidx[i] + 1]), L.ptr(grads[gidx[i] + 2]), st)
                eng.launches += 1
            dy = d                                                                       # (no BatchNorm: the dgrad above already applied the mask)
        if injs[i] is not None:
            if dy is None:
                dy = injs[i]
            else:
                L.call('b200gan_sample_axpby', C.byref(dy.v), None, C.byref(injs[i].v), None, C.byref(dy.v), st)     # dy += inj
                eng.launches += 1
        if dy is None:
            continue
        eng._wgrad(i, lc.x, dy, grads[gidx[i]], st)
        if i == 0:
            break
        below, lb = specs[i - 1], ctxs[i - 1]
        d = _empty_like(lb.a)
        if below.bn_idx is not None:
            lb.bsums2 = torch.empty(2 * below.cout, device=dev, dtype=torch.float64)
            fz = L.fuse(prev_act=below.act, prev_slope=E.LRELU_SLOPE, prev_y=lb.y.v, prev_scale=lb.scale, prev_shift=lb.shift, prev_mean=lb.mean,
                        prev_invstd=lb.invstd, prev_sums=lb.bsums2)
        else:
            fz = L.fuse(prev_act=below.act, prev_slope=E.LRELU_SLOPE, prev_y=lb.a.v)
        eng._dgrad(i, dy, p.w, d, st, lc.wp_down, lc.wp_up, fuse=fz)
    return gp.view(())
