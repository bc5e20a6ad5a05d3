"""Drop-in CLIs (CPU part): our train_gan.py / generate_synthetic.py expose exactly the reference's flags and defaults
(src/train_gan.py:217-240, src/generate_synthetic.py:63-70; the lists below were read off the reference) plus additive flags only;
a CPU round trip train -> generator_final.pth -> generate_synthetic writes the reference's artefacts."""
import json
import os

import numpy as np
import pytest
import torch

from gan_enhanced_pneumonia_classifier_b200 import generate_synthetic as gs
from gan_enhanced_pneumonia_classifier_b200 import train_gan as tg

REF_TRAIN_FLAGS = {
    'data_dir': './data/processed', 'model_dir': './models', 'output_dir': './results', 'results_dir': './results/metrics',
    'figures_dir': './results/figures', 'num_channels': 3, 'latent_dim': 100, 'feature_maps_g': 64, 'feature_maps_d': 64, 'epochs': 50,
    'batch_size': 128, 'lr': 0.0002, 'beta1': 0.5, 'workers': 4, 'vis_batch_size': 64, 'save_interval': 500, 'checkpoint_interval': 10,
    'cpu': False,
}
REF_SAMPLE_FLAGS = {'output_dir': './data/synthetic', 'num_images': 5000, 'latent_dim': 100, 'feature_maps_g': 64, 'batch_size': 64, 'cpu': False}


def test_train_cli_keeps_reference_flags_and_defaults():
    args = vars(tg.build_parser().parse_args([]))
    for k, v in REF_TRAIN_FLAGS.items():
        assert args[k] == v, k
    assert set(args) - set(REF_TRAIN_FLAGS) == {'dtype', 'synthetic', 'max_iters', 'log_interval', 'seed', 'cache_dataset', 'sync_bn'}        # additive only


def test_sampler_cli_keeps_reference_flags_and_defaults():
    args = vars(gs.build_parser().parse_args(['--model-path', 'x.pth']))
    for k, v in REF_SAMPLE_FLAGS.items():
        assert args[k] == v, k
    assert set(args) - set(REF_SAMPLE_FLAGS) == {'model_path', 'num_channels', 'seed'}


def test_cpu_round_trip_writes_reference_artefacts(tmp_path):
    d = str(tmp_path)
    argv = ['--cpu', '--synthetic', '4', '--batch-size', '2', '--epochs', '1', '--latent-dim', '8', '--feature-maps-g', '4', '--feature-maps-d', '4',
            '--num-channels', '3', '--vis-batch-size', '2', '--model-dir', d + '/models', '--output-dir', d + '/results',
            '--results-dir', d + '/results/metrics', '--figures-dir', d + '/results/figures', '--seed', '0', '--checkpoint-interval', '1']
    hist = tg.main(tg.build_parser().parse_args(argv))
    assert len(hist['G_losses_iter']) == 2 and len(hist['D_losses_epoch']) == 1
    for f in ('models/gan/generator_final.pth', 'models/gan/discriminator_final.pth', 'models/gan/generator_epoch_001.pth',
              'results/metrics/gan_training_history.json'):
        assert os.path.exists(os.path.join(d, f)), f
    assert any(n.startswith('fake_samples_epoch_001_iter_') for n in os.listdir(d + '/results/gan_images'))
    saved = json.load(open(d + '/results/metrics/gan_training_history.json'))
    assert set(saved) == set(tg.HISTORY_KEYS)
    sd = torch.load(d + '/models/gan/generator_final.pth')
    assert len(sd) == 31 and sd['main.0.weight'].shape == (8, 32, 7, 7) and sd['main.1.num_batches_tracked'].dtype == torch.int64
    n = gs.generate_images(d + '/models/gan/generator_final.pth', d + '/synthetic', 3, 8, 4, 2, torch.device('cpu'))
    assert n == 3 and sorted(os.listdir(d + '/synthetic')) == ['synthetic_00001.png', 'synthetic_00002.png', 'synthetic_00003.png']
    from PIL import Image
    img = np.asarray(Image.open(d + '/synthetic/synthetic_00001.png'))
    assert img.shape == (224, 224, 3) and img.dtype == np.uint8


def test_device_cache_has_no_host_path(tmp_path, capsys):
    """The device-resident image cache is CUDA only: a host tensor is rejected loudly, and `--cache-dataset --cpu` refuses to run
    instead of quietly falling back to a host loader."""
    from gan_enhanced_pneumonia_classifier_b200.data_cache import DeviceImageCache
    with pytest.raises(RuntimeError, match='GPU memory'):
        DeviceImageCache(torch.zeros((2, 3, 8, 8), dtype=torch.uint8))
    d = str(tmp_path)
    argv = ['--cpu', '--cache-dataset', '--synthetic', '4', '--batch-size', '2', '--epochs', '1', '--model-dir', d + '/models', '--output-dir', d + '/results',
            '--results-dir', d + '/results/metrics', '--figures-dir', d + '/results/figures']
    assert tg.main(tg.build_parser().parse_args(argv)) is None
    assert '--cache-dataset keeps the training images in GPU memory' in capsys.readouterr().out
    assert not os.path.exists(d + '/models/gan/generator_final.pth')


REF_WGAN_FLAGS = {
    'data_dir': './data/processed', 'model_dir': './models', 'output_dir': './results', 'results_dir': './results/metrics',
    'figures_dir': './results/figures', 'num_channels': 3, 'latent_dim': 100, 'feature_maps_g': 64, 'feature_maps_d': 64, 'epochs': 30,
    'batch_size': 64, 'lr': 0.0002, 'beta1': 0.5, 'workers': 4, 'vis_batch_size': 64, 'save_interval': 500, 'checkpoint_interval': 10,
    'critic_iters': 5, 'lambda_gp': 10.0, 'cpu': False,
}


def test_wgan_cli_keeps_reference_flags_and_defaults():
    """src/train_wggan.py:127-148 (the list above was read off the reference)."""
    from gan_enhanced_pneumonia_classifier_b200 import train_wggan as tw
    args = vars(tw.build_parser().parse_args([]))
    for k, v in REF_WGAN_FLAGS.items():
        assert args[k] == v, k
    assert set(args) - set(REF_WGAN_FLAGS) == {'dtype', 'synthetic', 'max_iters', 'log_interval', 'seed'}


def test_wgan_cpu_round_trip_writes_reference_artefacts(tmp_path):
    """--cpu runs the reference's stock-torch WGAN-GP loop over the drop-in modules (autograd double backward) and writes the reference's files."""
    from gan_enhanced_pneumonia_classifier_b200 import train_wggan as tw
    d = str(tmp_path)
    argv = ['--cpu', '--synthetic', '4', '--batch-size', '2', '--epochs', '1', '--latent-dim', '8', '--feature-maps-g', '2', '--feature-maps-d', '2',
            '--num-channels', '1', '--vis-batch-size', '2', '--critic-iters', '2', '--model-dir', d + '/models', '--output-dir', d + '/results',
            '--results-dir', d + '/results/metrics', '--figures-dir', d + '/results/figures', '--seed', '0', '--checkpoint-interval', '1']
    hist = tw.main(tw.build_parser().parse_args(argv))
    assert len(hist['D_losses']) == 4 and len(hist['G_losses']) == 2 and len(hist['D_losses_epoch']) == 1
    assert all(np.isfinite(hist['D_losses'])) and all(np.isfinite(hist['G_losses']))
    for f in ('models/wgan/generator_final.pth', 'models/wgan/discriminator_final.pth', 'models/wgan/generator_epoch_001.pth',
              'results/metrics/wgan_training_history.json'):
        assert os.path.exists(os.path.join(d, f)), f
    assert any(n.startswith('fake_samples_epoch_001_iter_') for n in os.listdir(d + '/results/wgan_images'))
    sd = torch.load(d + '/models/wgan/discriminator_final.pth')
    assert len(sd) == 20 and sd['main.11.weight'].shape == (1, 16, 7, 7)


# ---- conditional GAN CLI ---------------------------------------------------------------------------------
REF_CGAN_FLAGS = dict(data_dir='./data/processed', model_dir='./models', output_dir='./results', results_dir='./results/metrics',
                      figures_dir='./results/figures', num_channels=3, latent_dim=100, feature_maps_g=32, feature_maps_d=32, epochs=50, batch_size=32,
                      lr=0.0002, beta1=0.5, workers=4, vis_batch_size=32, save_interval=1000, checkpoint_interval=5, cpu=False)


def test_cgan_cli_keeps_reference_flags_and_defaults():
    """src/train_cgan.py:249-267 (the list above was read off the reference)."""
    from gan_enhanced_pneumonia_classifier_b200 import train_cgan as tc
    args = vars(tc.build_parser().parse_args([]))
    for k, v in REF_CGAN_FLAGS.items():
        assert args[k] == v, k
    assert set(args) - set(REF_CGAN_FLAGS) == {'no_perceptual', 'vgg_weights', 'dtype', 'synthetic', 'max_iters', 'log_interval', 'seed'}


def test_cgan_cpu_round_trip_writes_reference_artefacts(tmp_path):
    """--cpu --no-perceptual runs the reference's stock-torch loop (minus the VGG16 term) over the drop-in modules and writes the reference's files."""
    from gan_enhanced_pneumonia_classifier_b200 import train_cgan as tc
    d = str(tmp_path)
    argv = ['--cpu', '--no-perceptual', '--synthetic', '4', '--batch-size', '2', '--epochs', '1', '--latent-dim', '8', '--feature-maps-g', '2',
            '--feature-maps-d', '2', '--num-channels', '1', '--vis-batch-size', '3', '--model-dir', d + '/models', '--output-dir', d + '/results',
            '--results-dir', d + '/results/metrics', '--figures-dir', d + '/results/figures', '--seed', '0', '--checkpoint-interval', '1']
    hist = tc.main(tc.build_parser().parse_args(argv))
    assert set(hist) == set(tc.HISTORY_KEYS)
    assert len(hist['G_losses_epoch']) == 1 and np.isfinite(hist['G_losses_epoch'][0]) and np.isfinite(hist['feature_matching_losses'][0])
    assert hist['perceptual_losses'] == [0.0] and hist['G_losses_iter'] == []        # the reference never fills the per-iteration lists either
    for f in ('models/gan/generator_final.pth', 'models/gan/discriminator_final.pth', 'models/gan/generator_epoch_001.pth',
              'results/metrics/gan_training_history.json'):
        assert os.path.exists(os.path.join(d, f)), f
    assert any(n.startswith('fake_samples_epoch_001_iter_') for n in os.listdir(d + '/results/gan_images'))
    sd = torch.load(d + '/models/gan/generator_final.pth')
    assert sd['label_emb.weight'].shape == (2, 8) and sd['fc.weight'].shape == (16 * 49, 8) and sd['main.19.weight'].shape == (1, 1, 3, 3)
