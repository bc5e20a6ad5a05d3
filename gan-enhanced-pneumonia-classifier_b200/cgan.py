"""Drop-in for the reference's `src/cgan.py` (conditional GAN with a projection discriminator): same import
names (`weights_init`, `Generator`, `Discriminator`, `ProgressiveGenerator`, `ProgressiveDiscriminator`), constructor signatures, attribute
names (`label_emb`, `fc`, `main` with the reference's Sequential indices), hence identical `state_dict` keys, shapes and dtypes, and the
call contracts `netG(z, labels, alpha)`, `netD(x, labels, alpha)`, `netD.get_intermediate_features(x, labels, alpha)` (reference
cgan.py:6-12 weights_init, :14-61 Generator, :63-113 Discriminator; `alpha` is accepted and unused there too, :54,94,108).

For CUDA tensors the arithmetic runs on the hand-written sm_100a kernels of libb200gan.so (`cgan_engine`): the nn.Embedding / nn.Linear /
nn.Conv2d / nn.BatchNorm2d objects are parameter containers only; one `torch.autograd.Function` per network bridges autograd, so the
reference's loop (`train_cgan.py:150-193`: BCEWithLogits on `netD(...)`, feature matching on `get_intermediate_features`, `errG.backward()`,
`optim.Adam(net.parameters())`) runs unchanged on top.  CPU tensors run the stock torch modules (oracle / `--cpu` path, not a product path).
"""
import torch
import torch.nn as nn

from . import _lib as L
from . import cgan_engine as _ce


def weights_init(m):
    """reference cgan.py:6-12: Conv weights ~ N(0, 0.02); BatchNorm weight ~ N(1, 0.02), bias 0 (Linear / Embedding keep torch's defaults)."""
    kind = type(m).__name__
    if 'Conv' in kind:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif 'BatchNorm' in kind:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


class _CGANNet(nn.Module):
    def _engine_for(self):
        dtype = getattr(self, 'compute_dtype', None) or _ce.default_compute_dtype()
        algo = _ce.default_algo()
        cache = self.__dict__.setdefault('_b200_engines', {})
        if (dtype, algo) not in cache:
            cache[(dtype, algo)] = self._make_engine(dtype, algo)
        return cache[(dtype, algo)]

    def _wants_grad(self, *tensors):
        return torch.is_grad_enabled() and (any(t.requires_grad for t in tensors) or any(p.requires_grad for p in self.parameters()))


def _grad_list(plist, grads, needs, offset):
    return [grads.get(p) if needs[offset + j] else None for j, p in enumerate(plist)]


class _GenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, eng, z, labels, *plist):
        out, saved = eng.forward(mod, z, labels, save=True)
        ctx.mod, ctx.eng, ctx.saved, ctx.plist = mod, eng, saved, plist
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.saved is None:
            raise L.B200GanError('backward called twice on the same forward (activations were released)')
        dz, grads = ctx.eng.backward(ctx.mod, ctx.saved, dout.contiguous().float(), need_dz=ctx.needs_input_grad[2])
        ctx.saved = None
        return (None, None, dz, None, *_grad_list(ctx.plist, grads, ctx.needs_input_grad, 4))


class Generator(_CGANNet):
    """Generator(latent_dim, num_classes, num_channels, feature_maps_g): (z (N, latent_dim), labels (N,)) -> image (N, nc, 224, 224)."""

    def __init__(self, latent_dim, num_classes, num_channels, feature_maps_g):
        super().__init__()
        self.latent_dim, self.num_classes, self.init_size = latent_dim, num_classes, _ce.INIT_SIZE
        self._cfg = (latent_dim, num_classes, num_channels, feature_maps_g)
        widths = [feature_maps_g * 8, feature_maps_g * 4, feature_maps_g * 2, feature_maps_g, feature_maps_g // 2]
        self.label_emb = nn.Embedding(num_classes, latent_dim)
        self.fc = nn.Linear(latent_dim, widths[0] * self.init_size ** 2)
        layers = [nn.BatchNorm2d(widths[0]), nn.ReLU(True)]
        for cin, cout in zip(widths, widths[1:]):
            layers += [nn.Upsample(scale_factor=2), nn.Conv2d(cin, cout, 3, 1, 1), nn.BatchNorm2d(cout), nn.ReLU(True)]
        layers += [nn.Upsample(scale_factor=2), nn.Conv2d(widths[-1], num_channels, 3, 1, 1), nn.Tanh()]
        self.main = nn.Sequential(*layers)
        self.apply(weights_init)

    def _make_engine(self, dtype, algo):
        return _ce.GeneratorEngine(*self._cfg, dtype=dtype, algo=algo)

    def forward(self, z, labels, alpha=1.0):
        if not z.is_cuda:
            seed = self.fc(z + self.label_emb(labels))
            return self.main(seed.view(seed.size(0), -1, self.init_size, self.init_size))
        eng = self._engine_for()
        if self._wants_grad(z):
            return _GenFn.apply(self, eng, z, labels, *self.parameters())
        return eng.forward(self, z, labels, save=False)[0]


class _DiscFn(torch.autograd.Function):
    """Outputs: logits (N,) when `head`, followed by the nine distinct intermediates when `feats`."""

    @staticmethod
    def forward(ctx, mod, eng, head, feats, x, labels, *plist):
        logits, saved = eng.forward(mod, x, labels, save=True, head=head)
        ctx.mod, ctx.eng, ctx.saved, ctx.plist, ctx.head, ctx.feats = mod, eng, saved, plist, head, feats
        outs = ([logits] if head else []) + (eng.feature_tensors(saved[0]) if feats else [])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        if ctx.saved is None:
            raise L.B200GanError('backward called twice on the same forward (activations were released)')
        douts = list(douts)
        dlogits = douts.pop(0) if ctx.head else None
        needs = ctx.needs_input_grad
        dx, grads = ctx.eng.backward(ctx.mod, ctx.saved, dlogits, douts if ctx.feats else None, need_dx=needs[4], need_dw=any(needs[6:]))
        ctx.saved = None
        return (None, None, None, None, dx, None, *_grad_list(ctx.plist, grads, needs, 6))


class Discriminator(_CGANNet):
    """Discriminator(num_classes, num_channels, feature_maps_d): (image (N, nc, 224, 224), labels (N,)) -> one logit per image (N,)."""

    def __init__(self, num_classes, num_channels, feature_maps_d):
        super().__init__()
        nf = feature_maps_d
        self.num_classes = num_classes
        self._cfg = (num_classes, num_channels, feature_maps_d)
        self.label_emb = nn.Embedding(num_classes, nf * 8 * _ce.INIT_SIZE ** 2)
        widths = [num_channels, nf // 2, nf, nf * 2, nf * 4, nf * 8]
        layers = []
        for i, (cin, cout) in enumerate(zip(widths, widths[1:])):
            layers.append(nn.Conv2d(cin, cout, 4, 2, 1))
            if i > 0:
                layers.append(nn.BatchNorm2d(cout))
            layers.append(nn.LeakyReLU(0.2, inplace=True))
        layers.append(nn.Conv2d(widths[-1], 1, _ce.INIT_SIZE, 1, 0))
        self.main = nn.Sequential(*layers)
        self.apply(weights_init)

    def _make_engine(self, dtype, algo):
        return _ce.DiscriminatorEngine(*self._cfg, dtype=dtype, algo=algo)

    def _run(self, x, labels):
        eng = self._engine_for()
        if self._wants_grad(x):
            return _DiscFn.apply(self, eng, True, False, x, labels, *self.parameters())[0]
        return eng.forward(self, x, labels, save=False)[0]

    def forward(self, x, labels, alpha=1.0):
        if not x.is_cuda:
            h = self.main[:-1](x)
            proj = (self.label_emb(labels) * h.flatten(1)).sum(dim=1)
            return self.main[-1](h).view(-1) + proj
        return self._run(x, labels)

    def get_intermediate_features(self, x, labels, alpha=1.0):
        """The outputs of main[0..13] (cgan.py:108-113).  The reference's LeakyReLU layers are in place, so its list aliases: the entry of a
        Conv2d that feeds LeakyReLU directly (layer 0) and of every BatchNorm2d hold the LeakyReLU OUTPUT; the same tensors are returned here."""
        if not x.is_cuda:
            out = []
            for layer in self.main[:-1]:
                x = layer(x)
                out.append(x)
            return out
        a0, y1, a1, y2, a2, y3, a3, y4, a4 = self._features_cuda(x, labels)
        return [a0, a0, y1, a1, a1, y2, a2, a2, y3, a3, a3, y4, a4, a4]

    def _features_cuda(self, x, labels):
        eng = self._engine_for()
        if self._wants_grad(x):
            return _DiscFn.apply(self, eng, False, True, x, labels, *self.parameters())
        return eng.feature_tensors(eng.forward(self, x, labels, save=False, head=False)[1][0])


ProgressiveGenerator = Generator
ProgressiveDiscriminator = Discriminator
