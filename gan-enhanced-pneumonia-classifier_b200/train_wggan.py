"""Drop-in for the reference's `src/train_wggan.py` (WGAN-GP): same CLI flags and defaults (reference :127-148),
same artefacts (`<model-dir>/wgan/generator_epoch_%03d.pth`, `discriminator_epoch_%03d.pth`, `generator_final.pth`,
`discriminator_final.pth`, `<output-dir>/wgan_images/fake_samples_epoch_%03d_iter_%06d.png`, `<results-dir>/wgan_training_history.json` with
the four history lists of reference :57, `<figures-dir>/wgan_loss_curve.png`), same training semantics (reference :59-117: `critic_iters`
critic updates with gradient penalty per generator update, Adam betas (beta1, 0.9)).

On a CUDA device the inner loop (reference :66-93) is `WGANGPTrainer.critic_step` / `generator_step`: hand-written B200 kernels, no host
synchronisation per iteration (losses are flushed every `--log-interval` iterations instead of the reference's four `.item()` calls per
critic update).  `--cpu` keeps the reference's stock-torch loop (the oracle / CPU baseline, not a product path).
Additive flags only: --dtype {bf16,fp32}, --synthetic N, --max-iters, --log-interval, --seed.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.optim as optim

_HERE = os.path.dirname(os.path.abspath(__file__))
if __package__ in (None, ''):
    sys.path.insert(0, os.path.dirname(_HERE))
    sys.path.insert(0, _HERE)
    from gan_enhanced_pneumonia_classifier_b200.wggan import Discriminator, Generator, gradient_penalty, weights_init
    from gan_enhanced_pneumonia_classifier_b200.wgan_trainer import WGANGPTrainer
    from gan_enhanced_pneumonia_classifier_b200.train_gan import _DATA_LOADER_HINT, _save_image_grid, _save_state, _synthetic_loader
else:
    from .wggan import Discriminator, Generator, gradient_penalty, weights_init
    from .wgan_trainer import WGANGPTrainer
    from .train_gan import _DATA_LOADER_HINT, _save_image_grid, _save_state, _synthetic_loader

HISTORY_KEYS = ('D_losses', 'G_losses', 'D_losses_epoch', 'G_losses_epoch')


def plot_gan_losses(history, out_path):
    """reference :16-27; matplotlib is optional here."""
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
    except ImportError:
        print(f'matplotlib is not installed: skipping {out_path}')
        return
    plt.figure(figsize=(12, 6))
    plt.plot(history['D_losses'], label='Critic (D) Loss')
    plt.plot(history['G_losses'], label='Generator Loss')
    plt.legend()
    plt.xlabel('Iterations')
    plt.ylabel('Loss')
    plt.grid(True)
    plt.tight_layout()
    plt.savefig(out_path)
    plt.close()


def main(args):
    use_cuda = torch.cuda.is_available() and not args.cpu
    device = torch.device('cuda' if use_cuda else 'cpu')
    print(f'Device: {device}')
    if getattr(args, 'seed', None) is not None:
        torch.manual_seed(args.seed)
    model_dir = os.path.join(args.model_dir, 'wgan')
    image_dir = os.path.join(args.output_dir, 'wgan_images')
    for d in (model_dir, image_dir, args.results_dir, args.figures_dir):
        os.makedirs(d, exist_ok=True)

    n_syn = getattr(args, 'synthetic', 0)
    if n_syn:
        train_loader = _synthetic_loader(n_syn, args.num_channels, args.batch_size, 1)
    else:
        try:
            from data_loader import get_dataloaders          # the reference's src/data_loader.py:158
        except ImportError as e:
            print(f'Error: {e}')
            print(_DATA_LOADER_HINT)
            return None
        train_loader, _ = get_dataloaders(data_dir=args.data_dir, batch_size=args.batch_size, num_workers=args.workers)

    netG = Generator(args.latent_dim, args.num_channels, args.feature_maps_g).to(device)
    netD = Discriminator(args.num_channels, args.feature_maps_d).to(device)
    netG.apply(weights_init)
    netD.apply(weights_init)
    fixed_noise = torch.randn(args.vis_batch_size, args.latent_dim, device=device).unsqueeze(2).unsqueeze(3)
    trainer = None
    if use_cuda:
        dtype = {'bf16': torch.bfloat16, 'fp32': torch.float32}[getattr(args, 'dtype', 'bf16')]
        netG.compute_dtype = netD.compute_dtype = dtype
        trainer = WGANGPTrainer(netG, netD, lr=args.lr, beta1=args.beta1, beta2=0.9, lambda_gp=args.lambda_gp, critic_iters=args.critic_iters, dtype=dtype)
    else:
        optimizerG = optim.Adam(netG.parameters(), lr=args.lr, betas=(args.beta1, 0.9))
        optimizerD = optim.Adam(netD.parameters(), lr=args.lr, betas=(args.beta1, 0.9))

    history = {k: [] for k in HISTORY_KEYS}
    log_interval = max(1, getattr(args, 'log_interval', 50))
    max_iters = getattr(args, 'max_iters', 0) or 0
    iters, stop = 0, False
    start = time.time()
    for epoch in range(args.epochs):
        pending, d_epoch, g_epoch = [], [], []

        def flush():
            if not pending:
                return
            rows = torch.stack(pending).float().cpu().tolist()       # ONE sync for log_interval iterations
            pending.clear()
            for row in rows:
                history['D_losses'].extend(row[:-1])
                d_epoch.extend(row[:-1])
                history['G_losses'].append(row[-1])
                g_epoch.append(row[-1])

        for i, data in enumerate(train_loader):
            real_images = data[0].to(device, non_blocking=True)
            b_size = real_images.size(0)
            if trainer is not None:
                pending.append(trainer.step(real_images))
            else:
                row = []
                for _ in range(args.critic_iters):                   # reference :70-85
                    netD.zero_grad()
                    d_real_loss = -netD(real_images).mean()
                    noise = torch.randn(b_size, args.latent_dim, device=device).unsqueeze(2).unsqueeze(3)
                    fake_images = netG(noise)
                    d_fake_loss = netD(fake_images.detach()).mean()
                    gp = gradient_penalty(netD, real_images.data, fake_images.data, device, lambda_gp=args.lambda_gp)
                    d_loss = d_real_loss + d_fake_loss + gp
                    d_loss.backward()
                    optimizerD.step()
                    row.append(d_loss.detach())
                netG.zero_grad()                                     # reference :87-92
                noise = torch.randn(b_size, args.latent_dim, device=device).unsqueeze(2).unsqueeze(3)
                g_loss = -netD(netG(noise)).mean()
                g_loss.backward()
                optimizerG.step()
                row.append(g_loss.detach())
                pending.append(torch.stack(row))
            last = (epoch == args.epochs - 1 and i == len(train_loader) - 1) or (max_iters and iters + 1 >= max_iters)
            if (iters % args.save_interval == 0) or last:
                with torch.no_grad():
                    fake_vis = netG(fixed_noise).detach().float().cpu()
                _save_image_grid(fake_vis, f'{image_dir}/fake_samples_epoch_{epoch + 1:03d}_iter_{iters:06d}.png')
            iters += 1
            if len(pending) >= log_interval:
                flush()
            if max_iters and iters >= max_iters:
                stop = True
                break
        flush()
        history['D_losses_epoch'].append(float(np.mean(d_epoch)) if d_epoch else float('nan'))
        history['G_losses_epoch'].append(float(np.mean(g_epoch)) if g_epoch else float('nan'))
        print(f"Epoch {epoch + 1}/{args.epochs} Summary -  Avg Loss_D: {history['D_losses_epoch'][-1]:.4f}, Avg Loss_G: {history['G_losses_epoch'][-1]:.4f}")
        if (epoch + 1) % args.checkpoint_interval == 0 or (epoch + 1) == args.epochs:
            _save_state(netG, os.path.join(model_dir, f'generator_epoch_{epoch + 1:03d}.pth'))
            _save_state(netD, os.path.join(model_dir, f'discriminator_epoch_{epoch + 1:03d}.pth'))
        if stop:
            break
    _save_state(netG, os.path.join(model_dir, 'generator_final.pth'))
    _save_state(netD, os.path.join(model_dir, 'discriminator_final.pth'))
    print(f'Saved final models. ({time.time() - start:.1f} s)')
    with open(os.path.join(args.results_dir, 'wgan_training_history.json'), 'w') as f:
        json.dump(history, f, indent=4)
    plot_gan_losses(history, os.path.join(args.figures_dir, 'wgan_loss_curve.png'))
    return history


def build_parser():
    parser = argparse.ArgumentParser(description='Train Wasserstein GAN-GP on RSNA Pneumonia images')
    parser.add_argument('--data-dir', type=str, default='./data/processed')
    parser.add_argument('--model-dir', type=str, default='./models')
    parser.add_argument('--output-dir', type=str, default='./results')
    parser.add_argument('--results-dir', type=str, default='./results/metrics')
    parser.add_argument('--figures-dir', type=str, default='./results/figures')
    parser.add_argument('--num-channels', type=int, default=3)
    parser.add_argument('--latent-dim', type=int, default=100)
    parser.add_argument('--feature-maps-g', type=int, default=64)
    parser.add_argument('--feature-maps-d', type=int, default=64)
    parser.add_argument('--epochs', type=int, default=30)
    parser.add_argument('--batch-size', type=int, default=64)
    parser.add_argument('--lr', type=float, default=0.0002)
    parser.add_argument('--beta1', type=float, default=0.5)
    parser.add_argument('--workers', type=int, default=4)
    parser.add_argument('--vis-batch-size', type=int, default=64)
    parser.add_argument('--save-interval', type=int, default=500)
    parser.add_argument('--checkpoint-interval', type=int, default=10)
    parser.add_argument('--critic-iters', type=int, default=5, help='Number of D updates per G update')
    parser.add_argument('--lambda-gp', type=float, default=10., help='Gradient penalty coefficient')
    parser.add_argument('--cpu', action='store_true')
    # --- additive (not in the reference) --- #
    parser.add_argument('--dtype', choices=['bf16', 'fp32'], default='bf16', help='compute dtype of the B200 kernels')
    parser.add_argument('--synthetic', type=int, default=0, help='train on N synthetic uniform[-1,1] images instead of the RSNA loader')
    parser.add_argument('--max-iters', type=int, default=0, help='stop after this many generator iterations (0 = all epochs)')
    parser.add_argument('--log-interval', type=int, default=50, help='iterations between host synchronisations of the loss history')
    parser.add_argument('--seed', type=int, default=None)
    return parser


if __name__ == '__main__':
    a = build_parser().parse_args()
    print('--- Args ---')
    for k, v in vars(a).items():
        print(f'  {k}: {v}')
    main(a)
