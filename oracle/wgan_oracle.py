"""CPU oracle for the WGAN-GP training step -- TEST INFRASTRUCTURE ONLY.

numpy restatement of what the reference executes in `src/wggan.py` + `src/train_wggan.py:70-93` (harlanljones/gan-enhanced-pneumonia-classifier):

* `wggan.Generator` (wggan.py:15-46): ConvT(nz -> 16 ngf, k7) - BN - ReLU, 4 x [ConvT(k4 s2 p1) - BN - ReLU] (16 -> 8 -> 4 -> 2 -> 1 ngf), ConvT(ngf -> nc) - Tanh;
* `wggan.Discriminator` (the critic, wggan.py:48-70): Conv(nc -> ndf) - LeakyReLU(0.2), 3 x [Conv(k4 s2 p1) - BN - LeakyReLU], Conv(8 ndf -> 1, k7 s1 p0) on the
  14 x 14 map (an 8 x 8 score map), then the spatial mean (wggan.py:69-70): one score per image, no sigmoid;
* `gradient_penalty` (wggan.py:72-89): x^ = a x_real + (1-a) x_fake, g = d sum_b D(x^)_b / d x^ through the TRAIN-MODE critic (BatchNorm batch statistics couple the
  samples), gp = lambda * mean_b (||g_b||_2 - 1)^2, differentiated w.r.t. the critic's parameters through BOTH the forward pass and the first backward pass
  (`create_graph=True`).  The reference leaves that double backward to torch.autograd; here it is written out (`Critic.gradient_penalty`), because the CUDA path
  has to launch it kernel by kernel:
      first backward, layer l (top to bottom):  dz_l = da_l * lrelu'(z_l);  dy_l = gamma_l invstd_l P_l(dz_l);  da_{l-1} = dgrad_l(dy_l, W_l)
      with P_l(v) = v - mean(v) - xhat_l mean(v xhat_l) (the projection batch-norm backward applies per channel; symmetric).
      Reverse mode through that chain (bottom to top), u_0 = d gp / d g:
          r_l      = conv_l(u_{l-1}, W_l)                          adjoint of dy_l      (a FORWARD convolution of the adjoint)
          dW_l    += wgrad(x = u_{l-1}, dy = dy_l)                 the first backward's own use of W_l
          dgamma_l+= sum r_l invstd_l P_l(dz_l)
          u_l      = gamma_l invstd_l P_l(r_l) * lrelu'(z_l)       adjoint of da_l      (LeakyReLU is piecewise linear: no second-derivative term)
          inj_l    = -gamma_l invstd_l^2 [ xhat_l mean(r_l P_l(dz_l)) + mean(dz_l xhat_l) P_l(r_l) + mean(r_l xhat_l) P_l(dz_l) ]
                                                                   adjoint of the FORWARD conv output y_l (dy_l depends on y_l through xhat_l and invstd_l)
      followed by an ordinary backward pass of the forward graph that starts with no loss gradient at the top and picks up inj_l at every BatchNorm layer.
* the loop of train_wggan.py:70-93: `critic_iters` critic updates (each: D(real), G(noise), D(fake.detach()), gradient_penalty, Adam(beta1, 0.9)), then one
  generator update through the critic.

Pinned against fixtures generated from the reference itself (oracle/make_golden.py imports /root/reference/src/wggan.py and lets torch.autograd do the double
backward): tests/test_oracle_golden.py::test_wgan_*.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
from __future__ import annotations

import numpy as np

import dcgan_oracle as orc

LAMBDA_GP = 10.0        # train_wggan.py:141
BETA2 = 0.9             # train_wggan.py:53-54


def wgan_generator_plan(latent_dim, nc, ngf):
    """[(conv_index, bn_index or None, Cin, Cout, k, s, p)] for wggan.Generator.main (wggan.py:18-42)."""
    ch = [latent_dim, ngf * 16, ngf * 8, ngf * 4, ngf * 2, ngf, nc]
    return [(3 * i, 3 * i + 1 if i < 5 else None, ch[i], ch[i + 1], *((7, 1, 0) if i == 0 else (4, 2, 1))) for i in range(6)]


def critic_plan(nc, ndf):
    """wggan.Discriminator.main (wggan.py:51-64): Sequential indices 0 conv, 1 lrelu, 2 conv, 3 bn, 4 lrelu, 5 conv, 6 bn, 7 lrelu, 8 conv, 9 bn, 10 lrelu, 11 conv."""
    ch = [nc, ndf, ndf * 2, ndf * 4, ndf * 8, 1]
    conv_idx = [0, 2, 5, 8, 11]
    return [(conv_idx[i], conv_idx[i] + 1 if 1 <= i <= 3 else None, ch[i], ch[i + 1], *((7, 1, 0) if i == 4 else (4, 2, 1))) for i in range(5)]


class WGANGenerator(orc.Net):
    def __init__(self, latent_dim, nc, ngf, state, storage=None):
        super().__init__(wgan_generator_plan(latent_dim, nc, ngf), True, state, storage)

    def _act(self, i, z):
        return np.tanh(z) if i == 5 else np.maximum(z, 0)

    def _act_bwd(self, i, dout, z, out):
        return dout * (1 - out * out) if i == 5 else dout * (z > 0)


class Critic(orc.Net):
    """wggan.Discriminator: forward returns the (N, 1, 8, 8) score map; `scores` applies the spatial mean of wggan.py:69-70."""
    fp32_weight_layers = ()
    fp32_output_layers = (4,)

    def __init__(self, nc, ndf, state, storage=None):
        super().__init__(critic_plan(nc, ndf), False, state, storage)

    def _act(self, i, z):
        return z if i == 4 else np.where(z > 0, z, z * z.dtype.type(orc.LRELU_SLOPE))

    def _act_bwd(self, i, dout, z, out):
        return dout if i == 4 else dout * np.where(z > 0, z.dtype.type(1), z.dtype.type(orc.LRELU_SLOPE))

    def scores(self, x, train=True):
        smap, cache = self.forward(x, train)
        return smap.mean(axis=(2, 3), dtype=np.float64).astype(x.dtype).reshape(-1), cache

    def backward_from_scores(self, cache, dscore, need_input_grad):
        """dscore: (N,) gradient w.r.t. the per-image score; the spatial mean spreads it evenly over the score map."""
        smap = cache[-1][2]
        hw = smap.shape[2] * smap.shape[3]
        d = np.broadcast_to((dscore / dscore.dtype.type(hw)).reshape(-1, 1, 1, 1), smap.shape).astype(smap.dtype)
        return self.backward(cache, d, need_input_grad)

    # -- gradient penalty with its double backward (wggan.py:72-89), see the module docstring -----------------------------------
    def gradient_penalty(self, xhat_in, lambda_gp=LAMBDA_GP):
        """Returns (gp, grads over parameter keys).  One train-mode forward of the critic (BatchNorm buffers move, as in the reference)."""
        dt = xhat_in.dtype
        smap, cache = self.forward(xhat_in, train=True)
        nl = len(self.plan)
        slope = dt.type(orc.LRELU_SLOPE)
        n_img = xhat_in.shape[0]
        hw = smap.shape[2] * smap.shape[3]

        def proj(v, xh):
            m1 = v.mean(axis=(0, 2, 3), dtype=np.float64).astype(dt)[None, :, None, None]
            m2 = (v.astype(np.float64) * xh).mean(axis=(0, 2, 3)).astype(dt)[None, :, None, None]
            return v - m1 - xh * m2

        def cmean(v):
            return v.mean(axis=(0, 2, 3), dtype=np.float64).astype(dt)[None, :, None, None]

        # ---- first backward: g = d sum_b score_b / d xhat -------------------------------------------------------------------------
        d = np.full(smap.shape, 1.0 / hw, dtype=dt)
        dys, dzs, masks = [None] * nl, [None] * nl, [None] * nl
        for l in reversed(range(nl)):
            conv_i, bn_i, cin, cout, k, s, p = self.plan[l]
            a_in, z, out, xh, invstd = cache[l]
            if l == nl - 1:
                dz = d
            else:
                masks[l] = np.where(z > 0, dt.type(1), slope)
                dz = d * masks[l]
            dzs[l] = dz
            if bn_i is not None:
                gam = self.sd[f'main.{bn_i}.weight']
                dy = (gam * invstd)[None, :, None, None] * proj(dz, xh)
            else:
                dy = dz
            dys[l] = dy.astype(dt)
            d = orc.conv2d_dgrad(dys[l], self.sd[f'main.{conv_i}.weight'], s, p, a_in.shape[2:])
        g = d
        norms = np.sqrt((g.astype(np.float64) ** 2).reshape(n_img, -1).sum(axis=1))
        gp = dt.type(lambda_gp * np.mean((norms - 1.0) ** 2))
        # d gp / d g = lambda * (2/N) (||g_b|| - 1) g_b / ||g_b||
        u = (lambda_gp * 2.0 / n_img * (norms - 1.0) / norms).astype(dt).reshape(-1, 1, 1, 1) * g

        # ---- reverse mode through the first backward (bottom to top) ---------------------------------------------------------------
        grads = {}
        inj = [None] * nl
        for l in range(nl):
            conv_i, bn_i, cin, cout, k, s, p = self.plan[l]
            a_in, z, out, xh, invstd = cache[l]
            w = self.sd[f'main.{conv_i}.weight']
            grads[f'main.{conv_i}.weight'] = orc.conv2d_wgrad(u, dys[l], k, s, p)
            if l == nl - 1:
                break
            r = orc.conv2d_fprop(u, w, s, p)
            if bn_i is not None:
                gam = self.sd[f'main.{bn_i}.weight']
                gs = (gam * invstd)[None, :, None, None]
                pdz, pr = proj(dzs[l], xh), proj(r, xh)
                grads[f'main.{bn_i}.weight'] = (r.astype(np.float64) * (invstd[None, :, None, None] * pdz)).sum(axis=(0, 2, 3)).astype(dt)
                grads[f'main.{bn_i}.bias'] = np.zeros_like(gam)
                m2 = (dzs[l].astype(np.float64) * xh).mean(axis=(0, 2, 3)).astype(dt)[None, :, None, None]
                inj[l] = -(gs * invstd[None, :, None, None]) * (xh * cmean(r * pdz) + m2 * pr + cmean(r * xh) * pdz)
                adj_dz = gs * pr
            else:
                adj_dz = r
            u = (adj_dz * masks[l]).astype(dt)

        # ---- ordinary backward of the forward graph, fed only by the injected adjoints of the conv outputs ---------------------------
        d = None
        for l in reversed(range(nl - 1)):
            conv_i, bn_i, cin, cout, k, s, p = self.plan[l]
            a_in, z, out, xh, invstd = cache[l]
            dy = None
            if d is not None:
                dz = d * masks[l]
                if bn_i is not None:
                    dy, dg, db = orc.bn_train_bwd(dz, xh, self.sd[f'main.{bn_i}.weight'], invstd)
                    grads[f'main.{bn_i}.weight'] = grads[f'main.{bn_i}.weight'] + dg
                    grads[f'main.{bn_i}.bias'] = grads[f'main.{bn_i}.bias'] + db
                else:
                    dy = dz
            if inj[l] is not None:
                dy = inj[l] if dy is None else dy + inj[l]
            if dy is None:
                continue
            dy = dy.astype(dt)
            grads[f'main.{conv_i}.weight'] = grads[f'main.{conv_i}.weight'] + orc.conv2d_wgrad(a_in, dy, k, s, p)
            d = orc.conv2d_dgrad(dy, self.sd[f'main.{conv_i}.weight'], s, p, a_in.shape[2:]) if l > 0 else None
        return gp, grads


def critic_iteration(G, D, optD, real, noise, alpha, lambda_gp=LAMBDA_GP):
    """One pass of train_wggan.py:71-85.  `alpha` (N,1,1,1) is the interpolation draw of wggan.py:76."""
    dt = real.dtype
    n = real.shape[0]
    s_real, c_real = D.scores(real, train=True)
    gD, _ = D.backward_from_scores(c_real, np.full(n, -1.0 / n, dtype=dt), need_input_grad=False)
    fake, _ = G.forward(noise, train=True)
    s_fake, c_fake = D.scores(fake, train=True)
    g2, _ = D.backward_from_scores(c_fake, np.full(n, 1.0 / n, dtype=dt), need_input_grad=False)
    gD = orc.accumulate(gD, g2)
    xhat = (alpha * real + (1 - alpha) * fake).astype(dt)
    gp, g3 = D.gradient_penalty(xhat, lambda_gp)
    gD = orc.accumulate(gD, g3)
    d_loss = dt.type(-s_real.mean(dtype=np.float64) + s_fake.mean(dtype=np.float64) + gp)
    optD.step(D.sd, gD)
    return dict(d_loss=float(d_loss), gp=float(gp), grads_D=gD, fake=fake)


def generator_iteration(G, D, optG, noise):
    """train_wggan.py:87-92: g_loss = -mean D(G(z)) through the (train-mode) critic."""
    dt = noise.dtype
    n = noise.shape[0]
    fake, c_g = G.forward(noise, train=True)
    s, c_d = D.scores(fake, train=True)
    _, dfake = D.backward_from_scores(c_d, np.full(n, -1.0 / n, dtype=dt), need_input_grad=True)
    gG, _ = G.backward(c_g, dfake, need_input_grad=False)
    optG.step(G.sd, gG)
    return dict(g_loss=float(dt.type(-s.mean(dtype=np.float64))), grads_G=gG, fake=fake)
