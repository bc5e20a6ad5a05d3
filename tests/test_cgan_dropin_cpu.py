"""Conditional-GAN drop-in modules, CPU side: constructor / state_dict contract against the reference-generated
fixtures, and the CPU (`--cpu`) path of the modules against the numpy oracle.  The CUDA path is covered by tests/test_gpu_cgan.py."""
import json
import os
import sys

import numpy as np
import pytest
import torch

import cgan_oracle as co
from conftest import GOLDEN
from gan_enhanced_pneumonia_classifier_b200 import cgan
from parity_utils import close, synthetic_real
from test_oracle_golden import cgan_state

REF_SRC = '/root/reference/src'


def test_state_dict_layout_matches_the_reference_checkpoints():
    g = np.load(os.path.join(GOLDEN, 'cgan_main_nc3.npz'))
    m = json.loads(str(g['meta']))
    G, D = cgan.Generator(m['nz'], 2, m['nc'], m['nf']), cgan.Discriminator(2, m['nc'], m['nf'])
    for tag, net in (('G', G), ('D', D)):
        want = [k[len(f'final.{tag}.'):] for k in g.files if k.startswith(f'final.{tag}.')]
        sd = net.state_dict()
        assert list(sd.keys()) == want
        for k in want:
            assert tuple(sd[k].shape) == g[f'final.{tag}.{k}'].shape and str(sd[k].dtype).replace('torch.', '') == str(g[f'final.{tag}.{k}'].dtype), k
    assert cgan.ProgressiveGenerator is cgan.Generator and cgan.ProgressiveDiscriminator is cgan.Discriminator
    assert G.latent_dim == m['nz'] and G.init_size == 7 and G.num_classes == 2 and D.num_classes == 2


@pytest.mark.parametrize('name', ['cgan_step_nc1.npz', 'cgan_step_nc3.npz'])
def test_cpu_path_of_the_modules_matches_the_oracle(name):
    g = np.load(os.path.join(GOLDEN, name))
    m = json.loads(str(g['meta']))
    G, D = cgan.Generator(m['nz'], 2, m['nc'], m['nf']), cgan.Discriminator(2, m['nc'], m['nf'])
    sdG, sdD = cgan_state(g, 'G'), cgan_state(g, 'D')
    G.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sdG.items()})
    D.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sdD.items()})
    z, fl = g['it0.noise'], g['it0.fake_labels']
    fake = G(torch.from_numpy(z), torch.from_numpy(fl), 1.0)
    ref_fake, _ = co.GeneratorOracle(m['nz'], 2, m['nc'], m['nf'], sdG).forward(z, fl)
    close(fake.detach().numpy(), ref_fake, rtol=1e-4, atol=1e-5, what='Generator forward')
    real = synthetic_real(m['seed'] + 10, m['batch'], m['nc'])
    oD = co.DiscriminatorOracle(2, m['nc'], m['nf'], sdD)
    logits = D(torch.from_numpy(real), torch.from_numpy(g['it0.real_labels']))
    ref_logits, cache = oD.forward(real, g['it0.real_labels'])
    close(logits.detach().numpy(), ref_logits, rtol=1e-4, atol=2e-4, what='Discriminator logits')
    close(logits.detach().numpy(), g['it0.out_real'], rtol=1e-4, atol=2e-4, what='Discriminator logits vs the reference run')
    feats = D.get_intermediate_features(torch.from_numpy(real), torch.from_numpy(g['it0.real_labels']))
    assert len(feats) == 14 and feats[0] is feats[1] and feats[3] is feats[4] and feats[2] is not feats[3]       # the in-place LeakyReLU aliasing
    for a, b in zip(feats, oD.features(oD.forward(real, g['it0.real_labels'])[1])):
        close(a.detach().numpy(), b, rtol=1e-3, atol=1e-4, what='intermediate feature')


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason='the reference tree only exists in the build container')
def test_checkpoints_round_trip_with_the_reference_modules():
    sys.path.insert(0, REF_SRC)
    try:
        import cgan as ref_cgan
    finally:
        sys.path.remove(REF_SRC)
    assert os.path.abspath(ref_cgan.__file__).startswith('/root/reference')
    torch.manual_seed(3)
    mine_g, mine_d = cgan.Generator(16, 2, 3, 8), cgan.Discriminator(2, 3, 8)
    ref_g, ref_d = ref_cgan.Generator(16, 2, 3, 8), ref_cgan.Discriminator(2, 3, 8)
    ref_g.load_state_dict(mine_g.state_dict())            # strict: same keys and shapes both ways
    mine_d.load_state_dict(ref_d.state_dict())
    z, labels = torch.randn(3, 16), torch.tensor([0, 1, 1])
    x = torch.randn(3, 3, 224, 224)
    assert torch.equal(mine_g(z, labels), ref_g(z, labels))
    assert torch.allclose(mine_d(x, labels), ref_d(x, labels), rtol=1e-5, atol=1e-5)


def test_perceptual_loss_dropin_cpu_path_is_the_reference_formula():
    """perceptual.PerceptualLoss on CPU tensors: torchvision's vgg16.features[:4], [4:9], [9:16], chained, MSE summed (train_cgan.py:57-73), frozen
    parameters -- checked against the numpy oracle on the same random weights."""
    import vgg_oracle as vo
    from gan_enhanced_pneumonia_classifier_b200.perceptual import PerceptualLoss
    sd = vo.init_weights(np.random.RandomState(2))
    mod = PerceptualLoss('random')
    assert len(mod.blocks) == 3 and [len(b) for b in mod.blocks] == [4, 5, 7] and not mod.blocks.training
    assert all(not p.requires_grad for p in mod.parameters())
    torch.nn.Sequential(*[m for blk in mod.blocks for m in blk]).load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    rng = np.random.RandomState(3)
    x, y = rng.rand(1, 3, 16, 16).astype(np.float32), rng.rand(1, 3, 16, 16).astype(np.float32)
    xt = torch.from_numpy(x).requires_grad_(True)
    loss = mod(xt, torch.from_numpy(y))
    loss.backward()
    ref, dx = vo.perceptual(x, y, sd)
    close(loss.item(), ref, rtol=1e-5, what='perceptual loss')
    close(xt.grad.numpy(), dx, rtol=1e-3, atol=1e-7, what='d perceptual / d x')
