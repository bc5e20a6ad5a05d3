"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/b200gan.h
declares, the ctypes prototypes cover them one to one, and the module surface equals the reference's."""
import ctypes
import os
import re

import pytest
import torch

import gan_enhanced_pneumonia_classifier_b200 as pkg
from conftest import ROOT

L = pkg._lib


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'b200gan.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(b200gan_[A-Za-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(L.LIB_PATH), 'run __graft_entry__.build() first'
    lib = ctypes.CDLL(L.LIB_PATH)
    names = header_functions()
    assert len(names) >= 27
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/b200gan.h but not exported'
    assert sorted(list(L.PROTOTYPES) + L.OTHER_SYMBOLS) == names, 'ctypes prototypes and header differ'
    assert lib.b200gan_version() == 440        # 0.4.4: + b200gan_bn_finalize_act_fwd (0.4.3: conditional-GAN entry points, b200gan_dp_allreduce_f64)


def test_bad_arguments_fail_loudly_without_a_gpu():
    # argument validation happens before any CUDA call, so it is testable on the CPU box
    with pytest.raises(L.B200GanError, match='bad argument'):
        L.call('b200gan_adam', None, None, None, None, 0, 1e-3, 0.5, 0.999, 1e-8, 1, None, 1.0, None)
    v = L.View(0, 0, 1, 1, 1, 1, 1, 1, 1, 1)
    with pytest.raises(L.B200GanError, match='null view'):
        L.call('b200gan_bn_stats', ctypes.byref(v), None, None)


def test_dp_entry_points_validate_arguments_without_a_gpu():
    ident = (ctypes.c_ubyte * L.DP_ID_BYTES)()
    h = ctypes.c_void_p()
    with pytest.raises(L.B200GanError, match='bad argument'):
        L.call('b200gan_dp_init', ctypes.cast(ident, ctypes.c_void_p), 2, 5, ctypes.byref(h))         # rank outside the world
    with pytest.raises(L.B200GanError, match='bad argument'):
        L.call('b200gan_dp_allreduce_bucket', None, None, 0, None)
    with pytest.raises(L.B200GanError, match='null handle'):
        L.call('b200gan_dp_sync', None, None)
    L.call('b200gan_dp_destroy', None)                                                               # destroying nothing is fine
    assert L.load().b200gan_dp_collectives(None) == 0


def test_cuda_tensors_never_fall_back(monkeypatch):
    monkeypatch.setattr(L, '_lib', None)
    monkeypatch.setattr(L, 'LIB_PATH', '/nonexistent/libb200gan.so')
    with pytest.raises(L.B200GanError, match='no CPU/cuDNN fallback'):
        L.load()


def test_module_surface_matches_reference_contract():
    G, D = pkg.Generator(100, 3, 64), pkg.Discriminator(3, 64)
    gk, dk = list(G.state_dict()), list(D.state_dict())
    assert len(gk) == 31 and len(dk) == 26                              # SURVEY.md section 8a (a3, a5)
    assert G.state_dict()['main.0.weight'].shape == (100, 512, 7, 7)
    assert G.state_dict()['main.15.weight'].shape == (32, 3, 4, 4)
    assert D.state_dict()['main.14.weight'].shape == (1, 512, 7, 7)
    assert D.state_dict()['main.3.num_batches_tracked'].dtype == torch.int64
    assert sum(p.numel() for p in G.parameters()) == 5296576 + 32 * 2 * 16   # nc=3 adds 2*32*16 to the nc=1 count
    assert sum(p.numel() for p in D.parameters()) == 2812800 + 32 * 2 * 16
    assert [type(m).__name__ for m in D.main][:3] == ['Conv2d', 'LeakyReLU', 'Conv2d']
    # CPU tensors run the stock torch modules (the reference's --cpu path), shapes as dcgan.py:108-118 asserts
    G8, D8 = pkg.Generator(8, 3, 4), pkg.Discriminator(3, 4)
    img = G8(torch.randn(2, 8, 1, 1))
    assert img.shape == (2, 3, 224, 224) and D8(img).shape == (2,)
    with pytest.raises(RuntimeError):                                    # fact X1: 64x64 is not runnable
        D8(torch.randn(2, 3, 64, 64))
