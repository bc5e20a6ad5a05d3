"""Drop-in for the reference's `src/train_cgan.py` (conditional GAN): same CLI flags and defaults (reference
:249-267), same artefacts (`<model-dir>/gan/generator_epoch_%03d.pth`, `discriminator_epoch_%03d.pth`, `generator_final.pth`,
`discriminator_final.pth`, `<output-dir>/gan_images/fake_samples_epoch_%03d_iter_%06d.png`, `<results-dir>/gan_training_history.json` with the
nine history lists of reference :127-128 -- the five per-iteration ones stay empty there too --, `<figures-dir>/gan_loss_curve.png`), same training
semantics (reference :150-193: randomly smoothed BCEWithLogits targets, the D-step skip rule from epoch 5 on, adversarial + feature-matching
Generator loss, Adam betas (beta1, 0.999)).

The VGG16 perceptual loss (:57-73, 10 x in the Generator loss :186,191) is `perceptual.PerceptualLoss` (same `blocks`, frozen, B200 kernels on CUDA).
Like the reference it asks torchvision for the ImageNet checkpoint at start-up (`--vgg-weights imagenet`, the default): where that cannot be had
(offline, no cache) the CLI stops with a message instead of training something else -- give it a vgg16 state_dict file (`--vgg-weights FILE`), or say
explicitly that the term may be dropped (`--no-perceptual`) or run on random weights (`--vgg-weights random`, benchmarks only).
`--cpu` keeps the reference's stock-torch loop.
Additive flags only: --no-perceptual, --vgg-weights, --dtype {bf16,fp32}, --synthetic N, --max-iters, --log-interval, --seed.
Under `torchrun` (CUDA) the run is data parallel like train_gan.py's: one process per GPU, every rank trains on its own shard (DistributedSampler /
its own synthetic images and random draws), weights start from rank 0's initialisation, `CGANTrainer` sums the gradient arenas over the ranks on the
library's NCCL communicator and evaluates the D-step skip rule on rank-averaged probabilities; rank 0 writes the artefacts.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

_HERE = os.path.dirname(os.path.abspath(__file__))
if __package__ in (None, ''):
    sys.path.insert(0, os.path.dirname(_HERE))
    sys.path.insert(0, _HERE)
    from gan_enhanced_pneumonia_classifier_b200.cgan import Discriminator, Generator, weights_init
    from gan_enhanced_pneumonia_classifier_b200.cgan_trainer import CGANTrainer
    from gan_enhanced_pneumonia_classifier_b200.perceptual import PerceptualLoss
    from gan_enhanced_pneumonia_classifier_b200.train_gan import _DATA_LOADER_HINT, _save_image_grid, _save_state
else:
    from .cgan import Discriminator, Generator, weights_init
    from .cgan_trainer import CGANTrainer
    from .perceptual import PerceptualLoss
    from .train_gan import _DATA_LOADER_HINT, _save_image_grid, _save_state

HISTORY_KEYS = ('G_losses_iter', 'D_losses_iter', 'D_x_iter', 'D_G_z1_iter', 'D_G_z2_iter', 'G_losses_epoch', 'D_losses_epoch', 'perceptual_losses',
                'feature_matching_losses')
NUM_CLASSES = 2          # reference :106


def plot_gan_losses(history, out_path):
    """reference :19-55 (two panels: adversarial losses, additional loss components); matplotlib is optional here."""
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
    except ImportError:
        print(f'matplotlib is not installed: skipping {out_path}')
        return
    epochs = range(1, len(history['G_losses_epoch']) + 1)
    plt.figure(figsize=(12, 6))
    for panel, keys in ((1, (('G_losses_epoch', 'Generator Loss'), ('D_losses_epoch', 'Discriminator Loss'))),
                        (2, (('perceptual_losses', 'Perceptual Loss'), ('feature_matching_losses', 'Feature Matching Loss')))):
        plt.subplot(2, 1, panel)
        for key, label in keys:
            plt.plot(epochs, history[key], label=label, alpha=0.8)
        plt.xlabel('Epochs')
        plt.ylabel('Loss')
        plt.legend()
        plt.grid(True, linestyle='--', alpha=0.6)
    plt.tight_layout()
    plt.savefig(out_path)
    plt.close()


def _synthetic_labelled_loader(n_images, nc, batch_size, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n_images, nc, 224, 224, generator=g) * 2 - 1
    y = torch.randint(0, NUM_CLASSES, (n_images,), generator=g)
    return torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=batch_size, shuffle=True, generator=g,
                                       pin_memory=torch.cuda.is_available())


def _reference_cpu_iteration(netG, netD, optG, optD, criterion, perceptual, real, real_labels, epoch, latent_dim):
    """The reference's stock-torch iteration (:150-193), used for --cpu.  Returns the seven history values."""
    b, device = real.size(0), real.device
    smooth_real = torch.full((b,), 0.9, device=device) - 0.1 * torch.rand(b, device=device)
    smooth_fake = torch.full((b,), 0.1, device=device) + 0.1 * torch.rand(b, device=device)
    netD.zero_grad()
    out_real = netD(real, real_labels, 1.0)
    d_x = torch.sigmoid(out_real).mean().item()
    err_real = criterion(out_real, smooth_real)
    noise = torch.randn(b, latent_dim, device=device)
    fake_labels = torch.randint(0, NUM_CLASSES, (b,), device=device)
    fake = netG(noise, fake_labels, 1.0)
    out_fake = netD(fake.detach(), fake_labels, 1.0)
    d_g_z1 = torch.sigmoid(out_fake).mean().item()
    err_d = err_real + criterion(out_fake, smooth_fake)
    if d_x < 0.8 or d_g_z1 > 0.2 or epoch < 5:
        err_d.backward()
        optD.step()
    netG.zero_grad()
    out_g = netD(fake, fake_labels, 1.0)
    d_g_z2 = torch.sigmoid(out_g).mean().item()
    err_p = perceptual(fake, real) if perceptual is not None else torch.zeros((), device=device)
    err_fm = sum(torch.mean((r - f) ** 2) for r, f in zip(netD.get_intermediate_features(real, real_labels, 1.0),
                                                         netD.get_intermediate_features(fake, fake_labels, 1.0)))
    err_g = criterion(out_g, smooth_real) + 10.0 * err_p + 5.0 * err_fm
    err_g.backward()
    optG.step()
    return torch.tensor([err_d.item(), err_g.item(), d_x, d_g_z1, d_g_z2, float(err_p), err_fm.item()])


def main(args):
    use_cuda = torch.cuda.is_available() and not args.cpu
    world, rank, local_rank = (int(os.environ.get(k, d)) for k, d in (('WORLD_SIZE', '1'), ('RANK', '0'), ('LOCAL_RANK', '0')))
    if world > 1 and not use_cuda:
        print('Error: data-parallel training (torchrun) needs CUDA devices.')
        return None
    device = torch.device('cuda', local_rank) if use_cuda else torch.device('cpu')
    is_main = rank == 0
    if is_main:
        print(f'Using device: {device}' + (f' (data parallel over {world} ranks)' if world > 1 else ''))
    no_perceptual = getattr(args, 'no_perceptual', False)
    perceptual = None
    if not no_perceptual:
        source = getattr(args, 'vgg_weights', 'imagenet')
        try:
            perceptual = PerceptualLoss(source)               # reference :60,112: downloads torchvision's ImageNet checkpoint unless cached
        except Exception as e:   # noqa: BLE001
            if is_main:
                print(f'Error: cannot obtain the VGG16 weights of the perceptual loss ({source}): {e}')
                print('Pass --vgg-weights FILE (a torchvision vgg16 state_dict), --vgg-weights random (architecture only), or --no-perceptual to train '
                      'with the adversarial and feature-matching terms only.')
            return None
    if use_cuda:
        torch.cuda.set_device(device)
    if world > 1:
        torch.distributed.init_process_group('nccl')
    seed = getattr(args, 'seed', None)
    if seed is not None:
        torch.manual_seed(seed)          # common stream: weight initialisation and the fixed visualisation noise are the same on every rank
    model_dir = os.path.join(args.model_dir, 'gan')
    image_dir = os.path.join(args.output_dir, 'gan_images')
    for d in (model_dir, image_dir, args.results_dir, args.figures_dir):
        os.makedirs(d, exist_ok=True)

    n_syn = getattr(args, 'synthetic', 0)
    if n_syn:
        dataloader = _synthetic_labelled_loader(n_syn, args.num_channels, args.batch_size, 1 + rank)
    else:
        try:
            from data_loader import RSNAPneumoniaDataset, data_transforms          # the reference's src/data_loader.py
        except ImportError as e:
            print(f'Error loading data: {e}')
            print(_DATA_LOADER_HINT)
            return None
        try:
            dataset = RSNAPneumoniaDataset(data_dir=os.path.join(args.data_dir, 'Training', 'Images'),
                                           metadata_file=os.path.join(args.data_dir, 'stage2_train_metadata.csv'),
                                           transform=data_transforms['train'], is_test=False)
            if world > 1:
                sampler = torch.utils.data.distributed.DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=True, seed=seed or 0, drop_last=True)
                dataloader = torch.utils.data.DataLoader(dataset, batch_size=args.batch_size, sampler=sampler, num_workers=args.workers, drop_last=True)
            else:
                dataloader = torch.utils.data.DataLoader(dataset, batch_size=args.batch_size, shuffle=True, num_workers=args.workers)
        except Exception as e:                                                      # reference :101-103
            print(f'Error loading data: {e}')
            return None
    if is_main:
        print(f'Loaded training data with {len(dataloader.dataset)} samples.')
    if world > 1:                        # every rank must issue the same number of gradient exchanges per epoch
        nb = torch.tensor([len(dataloader), -len(dataloader)], device=device, dtype=torch.int64)
        torch.distributed.all_reduce(nb, op=torch.distributed.ReduceOp.MAX)
        if int(nb[0]) != -int(nb[1]):
            raise RuntimeError(f'data-parallel ranks disagree on the number of batches per epoch ({-int(nb[1])}..{int(nb[0])})')

    netG = Generator(args.latent_dim, NUM_CLASSES, args.num_channels, args.feature_maps_g).to(device)
    netD = Discriminator(NUM_CLASSES, args.num_channels, args.feature_maps_d).to(device)
    netG.apply(weights_init)
    netD.apply(weights_init)
    if world > 1:                        # all replicas start from rank 0's initialisation
        for t in list(netG.state_dict().values()) + list(netD.state_dict().values()):
            torch.distributed.broadcast(t, 0)
    fixed_noise = torch.randn(args.vis_batch_size, args.latent_dim, device=device)
    fixed_labels = torch.tensor(np.tile(np.arange(NUM_CLASSES), args.vis_batch_size // NUM_CLASSES + 1)[:args.vis_batch_size], dtype=torch.long,
                                device=device)
    if seed is not None and world > 1:
        torch.manual_seed(seed + rank)   # from here on (noise, fake labels, label smoothing) every rank draws its own stream
    trainer = None
    if use_cuda:
        dtype = {'bf16': torch.bfloat16, 'fp32': torch.float32}[getattr(args, 'dtype', 'bf16')]
        if perceptual is not None:
            perceptual = perceptual.to(device)
        trainer = CGANTrainer(netG, netD, lr=args.lr, beta1=args.beta1, perceptual=perceptual, perceptual_weight=10.0 if perceptual is not None else 0.0,
                              dtype=dtype)
    else:
        criterion = nn.BCEWithLogitsLoss()
        optimizerD = optim.Adam(netD.parameters(), lr=args.lr, betas=(args.beta1, 0.999))
        optimizerG = optim.Adam(netG.parameters(), lr=args.lr, betas=(args.beta1, 0.999))

    history = {k: [] for k in HISTORY_KEYS}
    log_interval = max(1, getattr(args, 'log_interval', 50))
    max_iters = getattr(args, 'max_iters', 0) or 0
    iters, stop = 0, False
    start = time.time()
    for epoch in range(args.epochs):
        epoch_start = time.time()
        pending, rows = [], []

        def flush():
            if pending:
                rows.extend(torch.stack(pending).float().cpu().tolist())          # ONE sync for log_interval iterations
                pending.clear()

        n_batches = len(dataloader)
        if world > 1 and hasattr(getattr(dataloader, 'sampler', None), 'set_epoch'):
            dataloader.sampler.set_epoch(epoch)
        for i, (real_images, real_labels) in enumerate(dataloader):
            real_images = real_images.to(device, non_blocking=True)
            real_labels = real_labels.to(device, non_blocking=True)
            if trainer is not None:
                pending.append(trainer.step(real_images, real_labels, epoch=epoch))
            else:
                pending.append(_reference_cpu_iteration(netG, netD, optimizerG, optimizerD, criterion, perceptual, real_images, real_labels, epoch,
                                                        args.latent_dim))
            last = (epoch == args.epochs - 1 and i == n_batches - 1) or (max_iters and iters + 1 >= max_iters)
            if (iters % args.save_interval == 0) or last:
                with torch.no_grad():                      # the networks stay in training mode, as in the reference (:208-210); every rank
                    fake_vis = netG(fixed_noise, fixed_labels, 1.0).detach().float().cpu()        # runs it: the BatchNorm buffers move
                if is_main:
                    _save_image_grid(fake_vis, f'{image_dir}/fake_samples_epoch_{epoch + 1:03d}_iter_{iters:06d}.png')
            iters += 1
            if len(pending) >= log_interval:
                flush()
            if max_iters and iters >= max_iters:
                stop = True
                break
        flush()
        r = np.array(rows) if rows else np.full((1, 7), np.nan)
        history['D_losses_epoch'].append(float(r[:, 0].mean()))
        history['G_losses_epoch'].append(float(r[:, 1].mean()))
        history['perceptual_losses'].append(float(r[:, 5].mean()))
        history['feature_matching_losses'].append(float(r[:, 6].mean()))
        if is_main:
            print(f"Epoch {epoch + 1}/{args.epochs} Summary - Time: {time.time() - epoch_start:.2f}s, Avg Loss_D: {history['D_losses_epoch'][-1]:.4f}, "
                  f"Avg Loss_G: {history['G_losses_epoch'][-1]:.4f}, D(x): {r[:, 2].mean():.3f}, D(G(z)): {r[:, 4].mean():.3f}")
        if is_main and ((epoch + 1) % args.checkpoint_interval == 0 or (epoch + 1) == args.epochs):
            _save_state(netG, os.path.join(model_dir, f'generator_epoch_{epoch + 1:03d}.pth'))
            _save_state(netD, os.path.join(model_dir, f'discriminator_epoch_{epoch + 1:03d}.pth'))
            print(f'Saved checkpoints for epoch {epoch + 1} to {model_dir}')
        if stop:
            break
    if is_main:
        print(f'Training finished in {time.time() - start:.2f} seconds.')
        _save_state(netG, os.path.join(model_dir, 'generator_final.pth'))
        _save_state(netD, os.path.join(model_dir, 'discriminator_final.pth'))
        print(f'Saved final models to {model_dir}')
        with open(os.path.join(args.results_dir, 'gan_training_history.json'), 'w') as f:
            json.dump(history, f, indent=4)
        plot_gan_losses(history, os.path.join(args.figures_dir, 'gan_loss_curve.png'))
    if world > 1:
        if trainer is not None:
            trainer.close()
        torch.distributed.destroy_process_group()
    return history


def build_parser():
    parser = argparse.ArgumentParser(description='Train cDCGAN on RSNA Pneumonia Dataset with Enhanced Logging')
    parser.add_argument('--data-dir', type=str, default='./data/processed')
    parser.add_argument('--model-dir', type=str, default='./models')
    parser.add_argument('--output-dir', type=str, default='./results')
    parser.add_argument('--results-dir', type=str, default='./results/metrics')
    parser.add_argument('--figures-dir', type=str, default='./results/figures')
    parser.add_argument('--num-channels', type=int, default=3)
    parser.add_argument('--latent-dim', type=int, default=100)
    parser.add_argument('--feature-maps-g', type=int, default=32)
    parser.add_argument('--feature-maps-d', type=int, default=32)
    parser.add_argument('--epochs', type=int, default=50)
    parser.add_argument('--batch-size', type=int, default=32)
    parser.add_argument('--lr', type=float, default=0.0002)
    parser.add_argument('--beta1', type=float, default=0.5)
    parser.add_argument('--workers', type=int, default=4)
    parser.add_argument('--vis-batch-size', type=int, default=32)
    parser.add_argument('--save-interval', type=int, default=1000)
    parser.add_argument('--checkpoint-interval', type=int, default=5)
    parser.add_argument('--cpu', action='store_true')
    # --- additive (not in the reference) --- #
    parser.add_argument('--no-perceptual', action='store_true', help='drop the VGG16 perceptual term of the Generator loss')
    parser.add_argument('--vgg-weights', type=str, default='imagenet', help="VGG16 weights of the perceptual loss: 'imagenet' (torchvision's checkpoint, "
                        "as the reference), a path to a vgg16 state_dict, or 'random' (architecture only)")
    parser.add_argument('--dtype', choices=['bf16', 'fp32'], default='bf16', help='compute dtype of the B200 kernels')
    parser.add_argument('--synthetic', type=int, default=0, help='train on N synthetic uniform[-1,1] images with random labels instead of the RSNA set')
    parser.add_argument('--max-iters', type=int, default=0, help='stop after this many iterations (0 = all epochs)')
    parser.add_argument('--log-interval', type=int, default=50, help='iterations between host synchronisations of the loss history')
    parser.add_argument('--seed', type=int, default=None)
    return parser


if __name__ == '__main__':
    a = build_parser().parse_args()
    print('--- Training Arguments ---')
    for k, v in vars(a).items():
        print(f'  {k}: {v}')
    print('-------------------------')
    main(a)
