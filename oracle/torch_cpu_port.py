"""CPU timing baseline: the reference's DCGAN step executed by stock torch.nn on the host cores -- TEST /
BENCH INFRASTRUCTURE ONLY (imported by bench.py's `cpu_baseline` and `--impl reference` legs and by tests).

/root/reference is pure Python and does not exist on the GPU box, so it cannot be timed there directly.  This
port restates what the reference executes on its `--cpu` path with the very same third-party engine (stock
`torch.nn` layers on CPU, oneDNN/ATen kernels, all host threads): the layer stacks of `dcgan.py:25-48,64-86`
and the per-iteration op sequence of `train_gan.py:121-150` (BCELoss, label smoothing 0.9, two Adam
optimizers lr 2e-4 betas (0.5, 0.999)).  `tests/test_oracle_golden.py::test_torch_port_matches_fixture` pins it
against the fixtures generated from the reference itself.  kind = "port" in bench.py's JSON.
"""
from __future__ import annotations

import time

import torch
import torch.nn as nn


def _stack(kind, chans, first_last_k7):
    layers = []
    n = len(chans) - 1
    for i in range(n):
        k7 = (i == 0) if kind == 'G' else (i == n - 1)
        k, s, p = (7, 1, 0) if k7 else (4, 2, 1)
        conv = nn.ConvTranspose2d if kind == 'G' else nn.Conv2d
        layers.append(conv(chans[i], chans[i + 1], k, s, p, bias=False))
        has_bn = (i < n - 1) if kind == 'G' else (0 < i < n - 1)
        if has_bn:
            layers.append(nn.BatchNorm2d(chans[i + 1]))
        if i == n - 1:
            layers.append(nn.Tanh() if kind == 'G' else nn.Sigmoid())
        else:
            layers.append(nn.ReLU(True) if kind == 'G' else nn.LeakyReLU(0.2, inplace=True))
    return nn.Sequential(*layers)


class PortNet(nn.Module):
    def __init__(self, kind, chans):
        super().__init__()
        self.kind = kind
        self.main = _stack(kind, chans, True)
        for m in self.modules():
            name = type(m).__name__
            if 'Conv' in name:
                nn.init.normal_(m.weight.data, 0.0, 0.02)
            elif 'BatchNorm' in name:
                nn.init.normal_(m.weight.data, 1.0, 0.02)
                nn.init.constant_(m.bias.data, 0)

    def forward(self, x):
        y = self.main(x)
        return y if self.kind == 'G' else y.view(-1, 1).squeeze(1)


def make_nets(nz=100, nc=1, ngf=64, ndf=64):
    G = PortNet('G', [nz, ngf * 8, ngf * 4, ngf * 2, ngf, ngf // 2, nc])
    D = PortNet('D', [nc, ndf // 2, ndf, ndf * 2, ndf * 4, ndf * 8, 1])
    return G, D


class CpuStepper:
    """One object = the state the reference's main() holds between iterations."""

    def __init__(self, nz=100, nc=1, ngf=64, ndf=64, lr=2e-4, beta1=0.5, threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.nz = nz
        self.G, self.D = make_nets(nz, nc, ngf, ndf)
        self.crit = nn.BCELoss()
        self.optD = torch.optim.Adam(self.D.parameters(), lr=lr, betas=(beta1, 0.999))
        self.optG = torch.optim.Adam(self.G.parameters(), lr=lr, betas=(beta1, 0.999))

    def step(self, real, noise=None):
        G, D, crit = self.G, self.D, self.crit
        b = real.size(0)
        D.zero_grad()
        label = torch.full((b,), 0.9, dtype=torch.float)
        out_real = D(real).view(-1)
        errD_real = crit(out_real, label)
        errD_real.backward()
        D_x = out_real.mean().item()
        if noise is None:
            noise = torch.randn(b, self.nz, 1, 1)
        fake = G(noise)
        label.fill_(0.0)
        out_fake = D(fake.detach()).view(-1)
        errD_fake = crit(out_fake, label)
        errD_fake.backward()
        D_G_z1 = out_fake.mean().item()
        errD = errD_real + errD_fake
        self.optD.step()
        G.zero_grad()
        label.fill_(0.9)
        out2 = D(fake).view(-1)
        errG = crit(out2, label)
        errG.backward()
        D_G_z2 = out2.mean().item()
        self.optG.step()
        return errD.item(), errG.item(), D_x, D_G_z1, D_G_z2


def host_threads():
    """The host threads this process may use (its CPU affinity mask; os.cpu_count() where affinity is not available)."""
    import os
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def time_cpu_steps(batch=64, steps=3, warmup=1, nc=1, seed=1, threads=None):
    """images/s of the CPU path on all host threads: `steps` timed iterations at `batch` images.  The thread count is set
    explicitly (torchrun exports OMP_NUM_THREADS=1, which would otherwise leave the CPU arm on one thread)."""
    torch.set_num_threads(threads or host_threads())
    torch.manual_seed(0)
    st = CpuStepper(nc=nc)
    g = torch.Generator().manual_seed(seed)
    real = torch.rand(batch, nc, 224, 224, generator=g) * 2 - 1
    for _ in range(warmup):
        st.step(real)
    t0 = time.perf_counter()
    for _ in range(steps):
        st.step(real)
    dt = time.perf_counter() - t0
    return dict(images_per_s=batch * steps / dt, seconds=dt, ms_per_step=dt / steps * 1e3, threads=torch.get_num_threads(),
                batch=batch, steps=steps)
