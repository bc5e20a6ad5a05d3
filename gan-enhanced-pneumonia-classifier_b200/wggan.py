"""Drop-in for the reference's `src/wggan.py` (WGAN-GP): same import names (`weights_init`, `Generator`,
`Discriminator`, `gradient_penalty`), constructor signatures and `self.main` Sequential indices, hence identical `state_dict` keys,
shapes and dtypes (reference wggan.py:5-13 weights_init, :15-46 Generator, :48-70 Discriminator, :72-89 gradient_penalty).

For CUDA tensors the arithmetic runs on the hand-written sm_100a kernels of libb200gan.so: the networks through `engine.NetEngine`
(first-order autograd is bridged by one `torch.autograd.Function` per network, as for the DCGAN), and `gradient_penalty` through
`wgan_engine.gradient_penalty`, which launches the penalty's DOUBLE backward explicitly (the reference gets it from
`torch.autograd.grad(create_graph=True)`) and hands the resulting parameter gradients to autograd, so that the reference's own loop
`d_loss = d_real_loss + d_fake_loss + gp; d_loss.backward()` (train_wggan.py:74-82) works unchanged.  CPU tensors run the stock torch
modules and the reference's autograd formulation (that is the oracle / `--cpu` path, not a product path).
"""
import torch
import torch.nn as nn

from . import _lib as L
from . import engine as _engine
from . import wgan_engine as _wgan
from .dcgan import _B200Net


def weights_init(m):
    """reference wggan.py:5-13: Conv / Linear weights ~ N(0, 0.02); BatchNorm / InstanceNorm weights ~ N(1, 0.02); every bias 0."""
    classname = m.__class__.__name__
    if hasattr(m, 'weight') and m.weight is not None:
        if 'Conv' in classname or 'Linear' in classname:
            nn.init.normal_(m.weight.data, 0.0, 0.02)
        elif 'BatchNorm' in classname or 'InstanceNorm' in classname:
            nn.init.normal_(m.weight.data, 1.0, 0.02)
    if hasattr(m, 'bias') and m.bias is not None:
        nn.init.constant_(m.bias.data, 0.)


class Generator(_B200Net):
    """Generator(latent_dim, num_channels, feature_maps_g): z (N, latent_dim, 1, 1) -> image (N, nc, 224, 224), base width 16 x feature_maps_g."""
    _transposed = True

    def __init__(self, latent_dim, num_channels, feature_maps_g):
        super().__init__()
        self._cfg = (latent_dim, num_channels, feature_maps_g)
        layers = []
        for sp in self._specs():
            layers.append(nn.ConvTranspose2d(sp.cin, sp.cout, sp.k, sp.stride, sp.pad, bias=False))
            layers += [nn.BatchNorm2d(sp.cout), nn.ReLU(True)] if sp.bn_idx is not None else [nn.Tanh()]
        self.main = nn.Sequential(*layers)
        self.apply(weights_init)

    def _specs(self):
        return _wgan.wgan_generator_specs(*self._cfg)

    def forward(self, z):
        return self._run(z)


class Discriminator(_B200Net):
    """The critic: Discriminator(num_channels, feature_maps_d): image (N, nc, 224, 224) -> one unbounded score per image (N,)."""

    def __init__(self, num_channels, feature_maps_d):
        super().__init__()
        self._cfg = (num_channels, feature_maps_d)
        layers = []
        for sp in self._specs():
            layers.append(nn.Conv2d(sp.cin, sp.cout, sp.k, sp.stride, sp.pad, bias=False))
            if sp.bn_idx is not None:
                layers.append(nn.BatchNorm2d(sp.cout))
            if sp.act == L.ACT_LRELU:
                layers.append(nn.LeakyReLU(0.2, inplace=True))
        self.main = nn.Sequential(*layers)
        self.apply(weights_init)

    def _specs(self):
        return _wgan.critic_specs(*self._cfg)

    def forward(self, x):
        out = self._run(x)                 # (N, 1, 8, 8) score map
        out = out.mean([2, 3])             # reference wggan.py:69
        return out.view(-1)


class _GradientPenalty(torch.autograd.Function):
    """gp and d gp / d (critic parameters) in one go: forward launches all four sweeps (wgan_engine.gradient_penalty), backward
    hands the stored gradients to autograd scaled by the incoming gradient."""

    @staticmethod
    def forward(ctx, D, eng, xhat, lambda_gp, *plist):
        params = _engine.params_from_module(D, eng.specs)
        grads = [torch.zeros_like(p, dtype=torch.float32) for p in plist]
        gp = _wgan.gradient_penalty(eng, params, xhat, grads, lambda_gp)
        ctx.grads = grads
        return gp

    @staticmethod
    def backward(ctx, go):
        grads, ctx.grads = ctx.grads, None
        return (None, None, None, None, *[g * go for g in grads])


def gradient_penalty(D, real_samples, fake_samples, device, lambda_gp=10.):
    """reference wggan.py:72-89, same signature.  The returned scalar carries d gp / d (parameters of D) into `.backward()`."""
    batch_size = real_samples.size(0)
    alpha = torch.rand(batch_size, 1, 1, 1, device=device)
    if not real_samples.is_cuda:
        interpolates = (alpha * real_samples + (1 - alpha) * fake_samples).requires_grad_(True)
        d_interpolates = D(interpolates)
        gradients = torch.autograd.grad(outputs=d_interpolates, inputs=interpolates, grad_outputs=torch.ones_like(d_interpolates),
                                        create_graph=True, retain_graph=True, only_inputs=True)[0]
        gradients = gradients.view(batch_size, -1)
        return ((gradients.norm(2, dim=1) - 1) ** 2).mean() * lambda_gp
    if not D.training:
        raise L.B200GanError('gradient_penalty through an eval-mode critic is not on the WGAN-GP training path and is not implemented')
    interpolates = (alpha * real_samples.detach().float() + (1 - alpha) * fake_samples.detach().float()).contiguous()
    eng = D._engine_for()
    return _GradientPenalty.apply(D, eng, interpolates, float(lambda_gp), *eng.param_order(D))
