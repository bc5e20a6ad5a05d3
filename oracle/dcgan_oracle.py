"""CPU oracle for the DCGAN adversarial training step -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the arithmetic the reference executes on its hot path
(`src/dcgan.py` + `src/train_gan.py:119-169` of harlanljones/gan-enhanced-pneumonia-classifier).
It is the checker for the CUDA path: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.  Nothing under
`gan-enhanced-pneumonia-classifier_b200/` imports it, and the product fails loudly when the CUDA library is absent.

Where the arithmetic lives: the reference tree holds no kernels; every op is a call into the
third-party dependency **PyTorch** (`requirements.txt:5`, `torch>=1.9.0`, unpinned; the container
has torch 2.11.0).  The published ATen semantics restated here are

* `conv2d` / `conv_transpose2d` (cross-correlation; transposed conv == conv input-gradient),
* `batch_norm` in training mode (biased batch variance for normalisation, unbiased variance in the
  running-stat update, momentum 0.1, eps 1e-5, `num_batches_tracked += 1`) and in eval mode,
* `relu`, `leaky_relu(0.2)`, `tanh`, `sigmoid`,
* `binary_cross_entropy` (log terms clamped at -100, `log1p(-p)` for the negative branch, backward
  divisor clamped at 1e-12),
* `optim.Adam` (lr, betas, eps 1e-8, no weight decay, no amsgrad).

Parity pin: the reference has no tests or golden vectors for this path, so the
oracle is pinned against outputs of the reference itself, generated in the build container by
`oracle/make_golden.py` (which imports `/root/reference/src/dcgan.py` and runs the unmodified
`train_gan.main`) and committed under `tests/golden/`.  `tests/test_oracle_golden.py` checks this
file against those fixtures.

All tensors are NCHW numpy arrays.  `dtype` defaults to float32, the storage/arithmetic type of the
reference; the long sums (convolutions, BatchNorm statistics, means) are accumulated in float64 and rounded
once, so the oracle sits at least as close to exact arithmetic as any fp32 summation order.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

BN_EPS = 1e-5          # nn.BatchNorm2d default (dcgan.py:27 etc.)
BN_MOMENTUM = 0.1
LRELU_SLOPE = 0.2      # dcgan.py:66
REAL_LABEL = 0.9       # train_gan.py:92
FAKE_LABEL = 0.0       # train_gan.py:93


# ----------------------------------------------------------------------------------------------
# convolution primitives on the "conv geometry": weight (Co, Ci, k, k), x (N, Ci, H, W),
# y (N, Co, OH, OW), OH = (H + 2p - k)//s + 1
# ----------------------------------------------------------------------------------------------
def conv_out_size(h, k, s, p):
    return (h + 2 * p - k) // s + 1


def conv2d_fprop(x, w, stride, pad):
    """y[n,co,oh,ow] = sum_{ci,kh,kw} x[n,ci,oh*s-p+kh,ow*s-p+kw] * w[co,ci,kh,kw]  (nn.Conv2d, dcgan.py:65-84)."""
    n, ci, h, wd = x.shape
    co, ci2, k, _ = w.shape
    assert ci == ci2
    oh, ow = conv_out_size(h, k, stride, pad), conv_out_size(wd, k, stride, pad)
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad))) if pad else x
    y = np.zeros((n, co, oh, ow), dtype=np.float64)           # accumulate in float64, round once
    for kh in range(k):
        for kw in range(k):
            xs = xp[:, :, kh:kh + stride * (oh - 1) + 1:stride, kw:kw + stride * (ow - 1) + 1:stride]
            y += np.einsum('nchw,oc->nohw', xs.astype(np.float64), w[:, :, kh, kw].astype(np.float64), optimize=True)
    return y.astype(x.dtype)


def conv2d_dgrad(dy, w, stride, pad, in_hw):
    """dx[n,ci,ih,iw] = sum_{co,kh,kw} dy[n,co,oh,ow] * w[co,ci,kh,kw] with ih = oh*s-p+kh (autograd of conv2d)."""
    n, co, oh, ow = dy.shape
    co2, ci, k, _ = w.shape
    assert co == co2
    h, wd = in_hw
    dxp = np.zeros((n, ci, h + 2 * pad, wd + 2 * pad), dtype=np.float64)
    for kh in range(k):
        for kw in range(k):
            dxp[:, :, kh:kh + stride * (oh - 1) + 1:stride, kw:kw + stride * (ow - 1) + 1:stride] += \
                np.einsum('nohw,oc->nchw', dy.astype(np.float64), w[:, :, kh, kw].astype(np.float64), optimize=True)
    return np.ascontiguousarray(dxp[:, :, pad:pad + h, pad:pad + wd]).astype(dy.dtype)


def conv2d_wgrad(x, dy, k, stride, pad):
    """dw[co,ci,kh,kw] = sum_{n,oh,ow} dy[n,co,oh,ow] * x[n,ci,oh*s-p+kh,ow*s-p+kw]."""
    n, ci, h, wd = x.shape
    _, co, oh, ow = dy.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad))) if pad else x
    dw = np.zeros((co, ci, k, k), dtype=x.dtype)
    for kh in range(k):
        for kw in range(k):
            xs = xp[:, :, kh:kh + stride * (oh - 1) + 1:stride, kw:kw + stride * (ow - 1) + 1:stride]
            # the reduction runs over N*OH*OW (up to millions of) terms: accumulate in float64
            dw[:, :, kh, kw] = np.tensordot(dy.astype(np.float64), xs.astype(np.float64), axes=([0, 2, 3], [0, 2, 3]))
    return dw


# nn.ConvTranspose2d (dcgan.py:26-46): weight (Cin_T, Cout_T, k, k).  With the conv geometry
# Co := Cin_T, Ci := Cout_T the transposed conv IS the conv input-gradient, and its own gradients are
# the conv forward / conv weight-gradient with the roles of x and dy exchanged.
def convT2d_out_size(h, k, s, p):
    return (h - 1) * s - 2 * p + k


def convT2d_fprop(x, w, stride, pad):
    k = w.shape[2]
    oh, ow = convT2d_out_size(x.shape[2], k, stride, pad), convT2d_out_size(x.shape[3], k, stride, pad)
    return conv2d_dgrad(x, w, stride, pad, (oh, ow))


def convT2d_dgrad(dy, w, stride, pad):
    return conv2d_fprop(dy, w, stride, pad)


def convT2d_wgrad(x, dy, k, stride, pad):
    return conv2d_wgrad(dy, x, k, stride, pad)


# ----------------------------------------------------------------------------------------------
# BatchNorm2d, activations, loss, optimiser
# ----------------------------------------------------------------------------------------------
def bn_train_fwd(y, gamma, beta, running_mean, running_var, nbt):
    """Training-mode batch_norm over (N,H,W) per channel.  Mutates the running buffers in place and
    returns (out, xhat, invstd).  `nbt` is a 1-element int64 array (num_batches_tracked)."""
    dt = y.dtype
    n = y.shape[0] * y.shape[2] * y.shape[3]
    mean = y.mean(axis=(0, 2, 3), dtype=np.float64)
    var = ((y.astype(np.float64) - mean[None, :, None, None]) ** 2).mean(axis=(0, 2, 3))   # biased
    invstd = (1.0 / np.sqrt(var + BN_EPS)).astype(dt)
    xhat = (y - mean.astype(dt)[None, :, None, None]) * invstd[None, :, None, None]
    out = xhat * gamma[None, :, None, None] + beta[None, :, None, None]
    unbiased = var * (n / max(n - 1, 1))
    running_mean[...] = ((1 - BN_MOMENTUM) * running_mean + BN_MOMENTUM * mean).astype(dt)
    running_var[...] = ((1 - BN_MOMENTUM) * running_var + BN_MOMENTUM * unbiased).astype(dt)
    nbt += 1
    return out.astype(dt), xhat.astype(dt), invstd


def bn_eval_fwd(y, gamma, beta, running_mean, running_var):
    scale = gamma / np.sqrt(running_var + np.asarray(BN_EPS, dtype=y.dtype))
    shift = beta - running_mean * scale
    return (y * scale[None, :, None, None] + shift[None, :, None, None]).astype(y.dtype)


def bn_train_bwd(dz, xhat, gamma, invstd):
    """native_batch_norm_backward in training mode: returns (dy, dgamma, dbeta)."""
    n = dz.shape[0] * dz.shape[2] * dz.shape[3]
    dbeta = dz.sum(axis=(0, 2, 3), dtype=np.float64)
    dgamma = (dz.astype(np.float64) * xhat).sum(axis=(0, 2, 3))
    m1 = (dbeta / n).astype(dz.dtype)[None, :, None, None]
    m2 = (dgamma / n).astype(dz.dtype)[None, :, None, None]
    dy = (gamma * invstd)[None, :, None, None] * (dz - m1 - xhat * m2)
    return dy.astype(dz.dtype), dgamma.astype(dz.dtype), dbeta.astype(dz.dtype)


def sigmoid(x):
    one = np.asarray(1, dtype=x.dtype)
    with np.errstate(over='ignore'):
        return (one / (one + np.exp(-x))).astype(x.dtype)


def bce_fwd(p, target):
    """nn.BCELoss(reduction='mean') (train_gan.py:90): log terms clamped at -100, log1p(-p) form."""
    dt = p.dtype
    t = np.asarray(target, dtype=dt)
    with np.errstate(divide='ignore'):
        lp = np.maximum(np.log(p), dt.type(-100))
        l1p = np.maximum(np.log1p(-p), dt.type(-100))
    loss = (t - dt.type(1)) * l1p - t * lp
    return dt.type(loss.mean(dtype=np.float64))


def bce_bwd(p, target):
    """d mean-BCE / dp = (p - t) / max(p (1-p), 1e-12) / N."""
    dt = p.dtype
    t = dt.type(target)
    return ((p - t) / np.maximum((dt.type(1) - p) * p, dt.type(1e-12)) / dt.type(p.size)).astype(dt)


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (train_gan.py:94-95,141,150); `step` is the 1-based count
    AFTER the increment.  Updates param/exp_avg/exp_avg_sq in place."""
    dt = param.dtype
    exp_avg += (grad - exp_avg) * dt.type(1 - beta1)                       # lerp_
    exp_avg_sq *= dt.type(beta2)
    exp_avg_sq += dt.type(1 - beta2) * grad * grad                         # addcmul_
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = np.sqrt(exp_avg_sq) / dt.type(math.sqrt(bc2)) + dt.type(eps)
    param -= dt.type(step_size) * (exp_avg / denom)


# ----------------------------------------------------------------------------------------------
# network plans (dcgan.py:25-48 and :64-86): the Sequential indices define the state_dict keys
# ----------------------------------------------------------------------------------------------
def generator_plan(latent_dim, nc, ngf):
    """[(conv_index, bn_index or None, Cin, Cout, k, s, p)] for Generator.main."""
    ch = [latent_dim, ngf * 8, ngf * 4, ngf * 2, ngf, ngf // 2, nc]
    plan = []
    for i in range(6):
        k, s, p = (7, 1, 0) if i == 0 else (4, 2, 1)
        plan.append((3 * i, 3 * i + 1 if i < 5 else None, ch[i], ch[i + 1], k, s, p))
    return plan


def discriminator_plan(nc, ndf):
    """[(conv_index, bn_index or None, Cin, Cout, k, s, p)] for Discriminator.main."""
    ch = [nc, ndf // 2, ndf, ndf * 2, ndf * 4, ndf * 8, 1]
    conv_idx = [0, 2, 5, 8, 11, 14]
    plan = []
    for i in range(6):
        k, s, p = (7, 1, 0) if i == 5 else (4, 2, 1)
        bn = conv_idx[i] + 1 if 1 <= i <= 4 else None
        plan.append((conv_idx[i], bn, ch[i], ch[i + 1], k, s, p))
    return plan


def init_state(plan, transposed, rng, dtype=np.float32):
    """A weights_init-like state dict (dcgan.py:6-12) drawn from a numpy RandomState (NOT torch's RNG)."""
    sd = OrderedDict()
    for conv_i, bn_i, cin, cout, k, s, p in plan:
        shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        sd[f'main.{conv_i}.weight'] = rng.normal(0.0, 0.02, size=shape).astype(dtype)
        if bn_i is not None:
            sd[f'main.{bn_i}.weight'] = rng.normal(1.0, 0.02, size=(cout,)).astype(dtype)
            sd[f'main.{bn_i}.bias'] = np.zeros((cout,), dtype)
            sd[f'main.{bn_i}.running_mean'] = np.zeros((cout,), dtype)
            sd[f'main.{bn_i}.running_var'] = np.ones((cout,), dtype)
            sd[f'main.{bn_i}.num_batches_tracked'] = np.zeros((), np.int64)
    return sd


def param_keys(plan):
    """Parameter order of `net.parameters()` (what optim.Adam sees)."""
    keys = []
    for conv_i, bn_i, *_ in plan:
        keys.append(f'main.{conv_i}.weight')
        if bn_i is not None:
            keys += [f'main.{bn_i}.weight', f'main.{bn_i}.bias']
    return keys


def bf16_round(x):
    """Round-to-nearest-even to bfloat16 precision, returned as float32 (what storing a tensor in bf16 and reading it back does)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


class Net:
    """Common forward/backward driver over a plan and a state dict (numpy arrays, updated in place).

    `storage`: None (the reference's fp32 arithmetic), or a rounding function (e.g. `bf16_round`) applied wherever the B200
    build's bf16 mode STORES a tensor: network input as read by the first convolution, the multiplicand copy of every conv
    weight (the fp32 masters stay exact; layers in `fp32_weight_layers` multiply by the fp32 weights), every convolution
    output that feeds a BatchNorm, every activation output, every activation-backward result dz, every BatchNorm-backward result dy, and the
    gradient handed out of the network.  Accumulation stays wide, statistics are those of the STORED tensors.  This is the
    model the bf16 tensor-core path is held to tightly (tests/test_gpu_fullsize.py); how far that model sits from the fp32
    reference is a property of bf16 storage, not of the kernels."""
    fp32_weight_layers = ()
    fp32_output_layers = ()

    def __init__(self, plan, transposed, state, storage=None):
        self.plan, self.transposed, self.sd = plan, transposed, state
        self.q = storage if storage is not None else (lambda t: t)

    # activation after layer i
    def _act(self, i, z):
        raise NotImplementedError

    def _act_bwd(self, i, dout, z, out):
        raise NotImplementedError

    def forward(self, x, train=True):
        cache = []
        q = self.q
        a = q(x)
        for i, (conv_i, bn_i, cin, cout, k, s, p) in enumerate(self.plan):
            w = self.sd[f'main.{conv_i}.weight']
            w = w if i in self.fp32_weight_layers else q(w)
            y = convT2d_fprop(a, w, s, p) if self.transposed else conv2d_fprop(a, w, s, p)
            # a convolution output is stored (rounded) only where BatchNorm needs its statistics; without BatchNorm the activation
            # rides on the convolution's epilogue and only the activation output is stored
            y = q(y) if bn_i is not None else y
            xhat = invstd = None
            if bn_i is not None:
                g, b = self.sd[f'main.{bn_i}.weight'], self.sd[f'main.{bn_i}.bias']
                rm, rv = self.sd[f'main.{bn_i}.running_mean'], self.sd[f'main.{bn_i}.running_var']
                if train:
                    z, xhat, invstd = bn_train_fwd(y, g, b, rm, rv, self.sd[f'main.{bn_i}.num_batches_tracked'])
                else:
                    z = bn_eval_fwd(y, g, b, rm, rv)
            else:
                z = y
            out = self._act(i, z)
            out = out if i in self.fp32_output_layers else q(out)
            cache.append((a, z, out, xhat, invstd))
            a = out
        return a, cache

    def backward(self, cache, dout, need_input_grad=True):
        """Returns (grads dict over parameter keys, dinput or None)."""
        grads = {}
        q = self.q
        d = dout
        for i in reversed(range(len(self.plan))):
            conv_i, bn_i, cin, cout, k, s, p = self.plan[i]
            a_in, z, out, xhat, invstd = cache[i]
            dz = self._act_bwd(i, d, z, out)
            dz = dz if i in self.fp32_output_layers else q(dz)
            if bn_i is not None:
                dy, dg, db = bn_train_bwd(dz, xhat, self.sd[f'main.{bn_i}.weight'], invstd)
                dy = q(dy)
                grads[f'main.{bn_i}.weight'], grads[f'main.{bn_i}.bias'] = dg, db
            else:
                dy = dz
            w = self.sd[f'main.{conv_i}.weight']
            w = w if i in self.fp32_weight_layers else q(w)
            if self.transposed:
                grads[f'main.{conv_i}.weight'] = convT2d_wgrad(a_in, dy, k, s, p)
                d = convT2d_dgrad(dy, w, s, p) if (i > 0 or need_input_grad) else None
            else:
                grads[f'main.{conv_i}.weight'] = conv2d_wgrad(a_in, dy, k, s, p)
                d = conv2d_dgrad(dy, w, s, p, a_in.shape[2:]) if (i > 0 or need_input_grad) else None
        return grads, (q(d) if d is not None else None)


class GeneratorOracle(Net):
    """dcgan.py:14-52: ConvT -> BN -> ReLU (x5), ConvT -> Tanh."""

    def __init__(self, latent_dim, nc, ngf, state, storage=None):
        super().__init__(generator_plan(latent_dim, nc, ngf), True, state, storage)

    def _act(self, i, z):
        return np.tanh(z) if i == 5 else np.maximum(z, 0)

    def _act_bwd(self, i, dout, z, out):
        if i == 5:
            return dout * (1 - out * out)
        return dout * (z > 0)


class DiscriminatorOracle(Net):
    """dcgan.py:54-90: Conv -> LeakyReLU, (Conv -> BN -> LeakyReLU) x4, Conv -> Sigmoid, flattened to (N,)."""

    fp32_weight_layers = (5,)      # the 7x7 GEMV multiplies bf16 activations by the fp32 master weights
    fp32_output_layers = (5,)      # logits, probabilities and the loss gradient stay fp32

    def __init__(self, nc, ndf, state, storage=None):
        super().__init__(discriminator_plan(nc, ndf), False, state, storage)

    def _act(self, i, z):
        if i == 5:
            return sigmoid(z)
        return np.where(z > 0, z, z * z.dtype.type(LRELU_SLOPE))

    def _act_bwd(self, i, dout, z, out):
        if i == 5:
            return dout * ((1 - out) * out)
        return dout * np.where(z > 0, z.dtype.type(1), z.dtype.type(LRELU_SLOPE))

    def probs(self, x, train=True):
        out, cache = self.forward(x, train)
        return out.reshape(-1), cache

    def backward_from_probs(self, cache, dp, need_input_grad):
        return self.backward(cache, dp.reshape(-1, 1, 1, 1), need_input_grad)


class AdamOracle:
    def __init__(self, keys, lr, beta1, beta2=0.999, eps=1e-8):
        self.keys, self.lr, self.beta1, self.beta2, self.eps = keys, lr, beta1, beta2, eps
        self.t = 0
        self.m, self.v = {}, {}

    def step(self, sd, grads):
        self.t += 1
        for key in self.keys:
            if key not in self.m:
                self.m[key] = np.zeros_like(sd[key])
                self.v[key] = np.zeros_like(sd[key])
            adam_step(sd[key], grads[key], self.m[key], self.v[key], self.t, self.lr, self.beta1, self.beta2, self.eps)


def accumulate(dst, src):
    for k, v in src.items():
        dst[k] = v.copy() if k not in dst else dst[k] + v
    return dst


def train_iteration(G, D, optG, optD, real, noise):
    """One pass of train_gan.py:121-150.  Returns the five per-iteration history scalars plus internals."""
    dt = real.dtype
    # (1) update D: real batch, then detached fake batch; gradients accumulate (train_gan.py:122-141)
    p_real, c_real = D.probs(real, train=True)
    errD_real = bce_fwd(p_real, REAL_LABEL)
    gD, _ = D.backward_from_probs(c_real, bce_bwd(p_real, REAL_LABEL), need_input_grad=False)
    D_x = float(p_real.mean(dtype=np.float64))
    fake, c_g = G.forward(noise, train=True)
    p_fake, c_fake = D.probs(fake, train=True)
    errD_fake = bce_fwd(p_fake, FAKE_LABEL)
    g2, _ = D.backward_from_probs(c_fake, bce_bwd(p_fake, FAKE_LABEL), need_input_grad=False)
    gD = accumulate(gD, g2)
    D_G_z1 = float(p_fake.mean(dtype=np.float64))
    errD = dt.type(errD_real + errD_fake)
    optD.step(D.sd, gD)
    # (2) update G through the already-updated D (train_gan.py:144-150)
    p2, c2 = D.probs(fake, train=True)
    errG = bce_fwd(p2, REAL_LABEL)
    _, dfake = D.backward_from_probs(c2, bce_bwd(p2, REAL_LABEL), need_input_grad=True)
    gG, _ = G.backward(c_g, dfake, need_input_grad=False)
    D_G_z2 = float(p2.mean(dtype=np.float64))
    optG.step(G.sd, gG)
    return {
        'errG': float(errG), 'errD': float(errD), 'D_x': D_x, 'D_G_z1': D_G_z1, 'D_G_z2': D_G_z2,
        'fake': fake, 'p_real': p_real, 'p_fake': p_fake, 'p_fake_for_G': p2, 'grads_D': gD, 'grads_G': gG,
    }


def train_iteration_dp(Gs, Ds, optGs, optDs, reals, noises):
    """Data-parallel semantics of the B200 build (the reference itself is single-device): R emulated
    ranks, each with its own replica (identical weights, LOCAL BatchNorm statistics and buffers) and its own shard
    (reals[r], noises[r]).  Each optimizer update uses the MEAN over ranks of the per-rank gradients -- what an all-reduce
    (sum) followed by a 1/R scaling delivers.  Returns the per-rank results of `train_iteration`'s bookkeeping plus the
    averaged gradients."""
    R = len(Gs)
    dt = reals[0].dtype
    out = [dict() for _ in range(R)]
    gDs, caches = [], []
    for r in range(R):
        G, D = Gs[r], Ds[r]
        p_real, c_real = D.probs(reals[r], train=True)
        gD, _ = D.backward_from_probs(c_real, bce_bwd(p_real, REAL_LABEL), need_input_grad=False)
        fake, c_g = G.forward(noises[r], train=True)
        p_fake, c_fake = D.probs(fake, train=True)
        g2, _ = D.backward_from_probs(c_fake, bce_bwd(p_fake, FAKE_LABEL), need_input_grad=False)
        gDs.append(accumulate(gD, g2))
        caches.append((fake, c_g))
        out[r].update(errD=float(dt.type(bce_fwd(p_real, REAL_LABEL) + bce_fwd(p_fake, FAKE_LABEL))),
                      D_x=float(p_real.mean(dtype=np.float64)), D_G_z1=float(p_fake.mean(dtype=np.float64)))
    gD_mean = {k: (sum(g[k].astype(np.float64) for g in gDs) / R).astype(dt) for k in gDs[0]}
    for r in range(R):
        optDs[r].step(Ds[r].sd, gD_mean)
    gGs = []
    for r in range(R):
        fake, c_g = caches[r]
        p2, c2 = Ds[r].probs(fake, train=True)
        _, dfake = Ds[r].backward_from_probs(c2, bce_bwd(p2, REAL_LABEL), need_input_grad=True)
        gG, _ = Gs[r].backward(c_g, dfake, need_input_grad=False)
        gGs.append(gG)
        out[r].update(errG=float(bce_fwd(p2, REAL_LABEL)), D_G_z2=float(p2.mean(dtype=np.float64)))
    gG_mean = {k: (sum(g[k].astype(np.float64) for g in gGs) / R).astype(dt) for k in gGs[0]}
    for r in range(R):
        optGs[r].step(Gs[r].sd, gG_mean)
    return out, gD_mean, gG_mean


def run_training(G, D, optG, optD, real_batches, noises, fixed_noise=None, save_interval=500):
    """train_gan.py:112-171 for one epoch over `real_batches`, including the train-mode visualisation
    forward (train_gan.py:166-169) that mutates G's BatchNorm buffers."""
    hist = {k: [] for k in ('G_losses_iter', 'D_losses_iter', 'D_x_iter', 'D_G_z1_iter', 'D_G_z2_iter')}
    vis = []
    last = len(real_batches) - 1
    for it, (real, z) in enumerate(zip(real_batches, noises)):
        r = train_iteration(G, D, optG, optD, real, z)
        hist['G_losses_iter'].append(r['errG'])
        hist['D_losses_iter'].append(r['errD'])
        hist['D_x_iter'].append(r['D_x'])
        hist['D_G_z1_iter'].append(r['D_G_z1'])
        hist['D_G_z2_iter'].append(r['D_G_z2'])
        if fixed_noise is not None and (it % save_interval == 0 or it == last):
            vis.append(G.forward(fixed_noise, train=True)[0])
    return hist, vis
