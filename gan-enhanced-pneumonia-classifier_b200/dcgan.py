"""Drop-in for the reference's `src/dcgan.py`: same import names, constructor signatures, `self.main`
Sequential indices (hence identical `state_dict` keys, shapes and dtypes -- `generator_final.pth` stays
loadable by the reference's `generate_synthetic.py:23-25`) and `forward(x) -> Tensor` contract.

What changes is who does the arithmetic.  The `nn.ConvTranspose2d` / `nn.Conv2d` / `nn.BatchNorm2d` objects in
`self.main` are kept purely as parameter and buffer containers.  For a CUDA input, `forward` hands the whole
network to the hand-written sm_100a kernels of libb200gan.so through a `torch.autograd.Function`
(`engine.run_network`), so `loss.backward()`, `fake.detach()`, `optim.Adam(net.parameters())`,
`.train()/.eval()`, `.state_dict()` all keep working for foreign callers.  For a CPU input the stock torch
modules run (`--cpu` in train_gan.py must stay the reference's own path: it is the oracle / CPU baseline, not
a product path).  There is no silent fallback on CUDA: a missing library raises.

Reference: /root/reference/src/dcgan.py:6-12 (weights_init), :14-52 (Generator), :54-90 (Discriminator).
"""
import os
import sys

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
if __package__ in (None, ''):
    # imported the way the reference imports it (`from dcgan import Generator` with this directory on
    # sys.path, e.g. by the reference's generate_synthetic.py): load the package through its import shim
    sys.path.insert(0, os.path.dirname(_HERE))
    import gan_enhanced_pneumonia_classifier_b200 as _pkg  # noqa: F401
    from gan_enhanced_pneumonia_classifier_b200 import engine as _engine
else:
    from . import engine as _engine


def weights_init(m):
    """DCGAN initialisation applied with `net.apply(weights_init)` (reference dcgan.py:6-12): conv and
    transposed-conv weights ~ N(0, 0.02); BatchNorm weight ~ N(1, 0.02), bias = 0."""
    kind = type(m).__name__
    if 'Conv' in kind:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif 'BatchNorm' in kind:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


class _B200Net(nn.Module):
    """Shared plumbing: lazily builds the per-dtype kernel engine for CUDA inputs."""
    _transposed = False

    def _specs(self):
        raise NotImplementedError

    def _engine_for(self, dtype=None, algo=None):
        dtype = dtype or getattr(self, 'compute_dtype', None) or _engine.default_compute_dtype()
        algo = _engine.default_algo() if algo is None else algo
        cache = self.__dict__.setdefault('_b200_engines', {})
        key = (dtype, algo)
        if key not in cache:
            cache[key] = _engine.NetEngine(self._specs(), self._transposed, dtype, algo)
        return cache[key]

    def _run(self, x):
        if x.is_cuda:
            return _engine.run_network(self, self._engine_for(), x)
        return self.main(x)          # the reference's own torch path (CPU oracle / --cpu)


class Generator(_B200Net):
    """Generator(latent_dim, num_channels, feature_maps_g): z (N, latent_dim, 1, 1) -> image (N, nc, 224, 224)."""
    _transposed = True

    def __init__(self, latent_dim, num_channels, feature_maps_g):
        super().__init__()
        self._cfg = (latent_dim, num_channels, feature_maps_g)
        layers = []
        for sp in _engine.generator_specs(latent_dim, num_channels, feature_maps_g):
            layers.append(nn.ConvTranspose2d(sp.cin, sp.cout, sp.k, sp.stride, sp.pad, bias=False))
            if sp.bn_idx is not None:
                layers += [nn.BatchNorm2d(sp.cout), nn.ReLU(True)]
            else:
                layers.append(nn.Tanh())
        self.main = nn.Sequential(*layers)
        self.apply(weights_init)

    def _specs(self):
        return _engine.generator_specs(*self._cfg)

    def forward(self, input):
        return self._run(input)


class Discriminator(_B200Net):
    """Discriminator(num_channels, feature_maps_d): image (N, nc, 224, 224) -> probability of "real" (N,)."""

    def __init__(self, num_channels, feature_maps_d):
        super().__init__()
        self._cfg = (num_channels, feature_maps_d)
        layers = []
        for sp in _engine.discriminator_specs(num_channels, feature_maps_d):
            layers.append(nn.Conv2d(sp.cin, sp.cout, sp.k, sp.stride, sp.pad, bias=False))
            if sp.bn_idx is not None:
                layers.append(nn.BatchNorm2d(sp.cout))
            layers.append(nn.LeakyReLU(0.2, inplace=True) if sp.cout != 1 or sp.k != 7 else nn.Sigmoid())
        self.main = nn.Sequential(*layers)
        self.apply(weights_init)

    def _specs(self):
        return _engine.discriminator_specs(*self._cfg)

    def forward(self, input):
        return self._run(input).view(-1, 1).squeeze(1)
