"""Checkpoint interoperability with the UNMODIFIED reference, on CPU:

  * the reference's own `generate_synthetic.generate_images` (src/generate_synthetic.py:18-61: `Generator(latent_dim, 3, fm)`,
    `load_state_dict(torch.load(path))`, eval-mode forward, PNG output) accepts a `generator_final.pth` written by OUR train_gan.py;
  * OUR `Generator` / `Discriminator` load the `generator_final.pth` / `discriminator_final.pth` the reference's unmodified
    `train_gan.main` wrote (held in tests/golden/main_small_nc3.npz as `final.G.*` / `final.D.*`) with strict key / shape / dtype
    matching, and reproduce the reference module's eval-mode forward on it.

The first half needs /root/reference (present in the build container only, where the driver runs the `-m "not gpu"` suite); it is
skipped elsewhere.  The second half runs everywhere from the committed fixture."""
import importlib.util
import json
import os
import sys

import numpy as np
import pytest
import torch

import gan_enhanced_pneumonia_classifier_b200 as pkg
from conftest import GOLDEN
from gan_enhanced_pneumonia_classifier_b200 import train_gan as tg

REF_SRC = '/root/reference/src'


def _load_reference(name):
    """Import /root/reference/src/<name>.py under a private module name, with the reference directory first on sys.path so that its
    own `from dcgan import Generator` / `from utils import check_create_dir` resolve to the reference's files."""
    sys.path.insert(0, REF_SRC)
    saved = {k: sys.modules.pop(k) for k in ('dcgan', 'cgan', 'wggan', 'utils') if k in sys.modules}
    try:
        spec = importlib.util.spec_from_file_location(f'_reference_{name}', os.path.join(REF_SRC, f'{name}.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        sys.path.remove(REF_SRC)
        for k in ('dcgan', 'cgan', 'wggan', 'utils'):
            sys.modules.pop(k, None)
        sys.modules.update(saved)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_SRC, 'generate_synthetic.py')), reason='/root/reference is only present in the build container')
def test_unmodified_reference_sampler_loads_our_checkpoint(tmp_path):
    d = str(tmp_path)
    argv = ['--cpu', '--synthetic', '4', '--batch-size', '2', '--epochs', '1', '--latent-dim', '8', '--feature-maps-g', '4', '--feature-maps-d', '4',
            '--num-channels', '3', '--vis-batch-size', '2', '--model-dir', d + '/models', '--output-dir', d + '/results',
            '--results-dir', d + '/results/metrics', '--figures-dir', d + '/results/figures', '--seed', '0']
    tg.main(tg.build_parser().parse_args(argv))
    ckpt = d + '/models/gan/generator_final.pth'
    ref_gs = _load_reference('generate_synthetic')
    assert ref_gs.__file__.startswith(REF_SRC) and ref_gs.Generator.__init__.__code__.co_filename.startswith(REF_SRC)   # the reference's own classes
    # the call the reference's __main__ makes (generate_synthetic.py:82-90); any load error there ends in sys.exit(1)
    torch.manual_seed(3)
    ref_gs.generate_images(ckpt, d + '/ref_synthetic', 3, 8, 4, 2, torch.device('cpu'))
    assert sorted(os.listdir(d + '/ref_synthetic')) == ['synthetic_00001.png', 'synthetic_00002.png', 'synthetic_00003.png']
    # and the reference Generator (its own class, dcgan.py:14-52) computes what our module computes from the same file
    ref_dcgan = _load_reference('dcgan')
    refG = ref_dcgan.Generator(8, 3, 4)
    refG.load_state_dict(torch.load(ckpt))          # strict
    ours = pkg.Generator(8, 3, 4)
    ours.load_state_dict(torch.load(ckpt))
    z = torch.randn(2, 8, 1, 1)
    refG.eval(), ours.eval()
    with torch.no_grad():
        assert torch.equal(refG(z), ours(z))


def test_our_modules_load_the_reference_produced_checkpoints():
    g = np.load(os.path.join(GOLDEN, 'main_small_nc3.npz'))
    m = json.loads(str(g['meta']))
    sdG = {k[len('final.G.'):]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith('final.G.')}
    sdD = {k[len('final.D.'):]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith('final.D.')}
    assert len(sdG) == 31 and len(sdD) == 26
    G, D = pkg.Generator(m['nz'], m['nc'], m['fm']), pkg.Discriminator(m['nc'], m['fm'])
    res = G.load_state_dict(sdG, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    res = D.load_state_dict(sdD, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for net, sd in ((G, sdG), (D, sdD)):
        for k, v in net.state_dict().items():
            assert v.dtype == sd[k].dtype and v.shape == sd[k].shape and torch.equal(v, sd[k]), k
    # the numpy oracle's eval-mode forward of that checkpoint equals our module's CPU forward (what generate_synthetic.py runs)
    import dcgan_oracle as orc
    z = g['fixed_noise']
    ref, _ = orc.GeneratorOracle(m['nz'], m['nc'], m['fm'], {k: v.numpy() for k, v in sdG.items()}).forward(z, train=False)
    G.eval()
    with torch.no_grad():
        out = G(torch.from_numpy(z)).numpy()
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-5)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_SRC, 'generate_synthetic_cgan.py')), reason='/root/reference is only present in the build container')
def test_unmodified_reference_cgan_and_wgan_samplers_load_our_checkpoints(tmp_path):
    """generate_synthetic_cgan.py / generate_synthetic_wgan.py of the reference (their own Generator classes, strict load_state_dict, eval-mode forward,
    PNG output) accept the `generator_final.pth` files written by OUR train_cgan.py / train_wggan.py."""
    from gan_enhanced_pneumonia_classifier_b200 import train_cgan as tc
    from gan_enhanced_pneumonia_classifier_b200 import train_wggan as tw
    d = str(tmp_path)
    common = ['--cpu', '--synthetic', '4', '--batch-size', '2', '--epochs', '1', '--latent-dim', '8', '--feature-maps-g', '2', '--feature-maps-d', '2',
              '--num-channels', '3', '--vis-batch-size', '2', '--seed', '0']

    def dirs(tag):
        return ['--model-dir', f'{d}/{tag}/models', '--output-dir', f'{d}/{tag}/results', '--results-dir', f'{d}/{tag}/results/metrics',
                '--figures-dir', f'{d}/{tag}/results/figures']

    tc.main(tc.build_parser().parse_args(common + dirs('cgan') + ['--no-perceptual']))
    tw.main(tw.build_parser().parse_args(common + dirs('wgan') + ['--critic-iters', '1']))
    for name, ckpt, out in (('generate_synthetic_cgan', f'{d}/cgan/models/gan/generator_final.pth', f'{d}/ref_cgan'),
                            ('generate_synthetic_wgan', f'{d}/wgan/models/wgan/generator_final.pth', f'{d}/ref_wgan')):
        ref = _load_reference(name)
        assert ref.__file__.startswith(REF_SRC) and ref.Generator.__init__.__code__.co_filename.startswith(REF_SRC)      # the reference's own classes
        torch.manual_seed(4)
        ref.generate_images(ckpt, out, 3, 8, 2, 2, torch.device('cpu'))          # any load error there ends in sys.exit(1)
        assert sorted(os.listdir(out)) == ['synthetic_00001.png', 'synthetic_00002.png', 'synthetic_00003.png']
