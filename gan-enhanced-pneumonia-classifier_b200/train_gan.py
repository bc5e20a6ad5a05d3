"""Drop-in for the reference's `src/train_gan.py`: same CLI flags and defaults (reference :217-240), same artefacts
(`<model-dir>/gan/generator_epoch_%03d.pth`, `discriminator_epoch_%03d.pth`, `generator_final.pth`,
`discriminator_final.pth`, `<output-dir>/gan_images/fake_samples_epoch_%03d_iter_%06d.png`,
`<results-dir>/gan_training_history.json` with the seven history lists of reference :100-108,
`<figures-dir>/gan_loss_curve.png`), same training semantics (reference :112-196).

What differs: on a CUDA device the inner loop (reference :119-150) is ONE call to `DCGANTrainer.step`, which runs
the hand-written B200 kernels with no host synchronisation; the nine per-iteration `.item()` syncs of the reference
(:130,:138,:149,:153-163) are replaced by a device-side list that is flushed every `--log-interval` iterations.
`--cpu` keeps the reference's own stock-torch loop (it is the oracle / CPU baseline, not a product path).

Additive flags only: --dtype {bf16,fp32}, --synthetic N (train on N synthetic uniform[-1,1] images instead of the
RSNA loader), --max-iters, --log-interval, --seed.  Launched under torchrun it trains data-parallel (one process per
GPU, NCCL all-reduce of the two gradient arenas, per-rank BatchNorm statistics; rank 0 writes the artefacts).
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn as nn
import torch.optim as optim

_HERE = os.path.dirname(os.path.abspath(__file__))
if __package__ in (None, ''):
    sys.path.insert(0, os.path.dirname(_HERE))
    sys.path.insert(0, _HERE)
    from gan_enhanced_pneumonia_classifier_b200.dcgan import Discriminator, Generator, weights_init
    from gan_enhanced_pneumonia_classifier_b200.trainer import DCGANTrainer
    from gan_enhanced_pneumonia_classifier_b200.data_cache import DeviceImageCache
else:
    from .dcgan import Discriminator, Generator, weights_init
    from .trainer import DCGANTrainer
    from .data_cache import DeviceImageCache

HISTORY_KEYS = ('G_losses_iter', 'D_losses_iter', 'D_x_iter', 'D_G_z1_iter', 'D_G_z2_iter', 'G_losses_epoch', 'D_losses_epoch')


def plot_gan_losses(history, output_path):
    """Loss curves (reference :18-45).  matplotlib is optional here: without it the plot is skipped with a message."""
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
    except ImportError:
        print(f'matplotlib is not installed: skipping {output_path}')
        return
    g, d = history.get('G_losses_iter', []), history.get('D_losses_iter', [])
    if not g or not d:
        print('Warning: Loss data missing or empty in history. Skipping plot generation.')
        return
    plt.figure(figsize=(12, 6))
    plt.plot(range(len(g)), g, label='Generator Loss', alpha=0.8)
    plt.plot(range(len(d)), d, label='Discriminator Loss', alpha=0.8)
    plt.title('Generator and Discriminator Loss During Training (Per Iteration)')
    plt.xlabel('Iterations')
    plt.ylabel('Loss (BCELoss)')
    plt.legend()
    plt.grid(True, linestyle='--', alpha=0.6)
    plt.tight_layout()
    try:
        plt.savefig(output_path)
        print(f'Saved GAN loss plot to {output_path}')
    except Exception as e:       # noqa: BLE001  (reference behaviour: report and continue)
        print(f'Error saving plot to {output_path}: {e}')
    plt.close()


def _synthetic_loader(n_images, nc, batch_size, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n_images, nc, 224, 224, generator=g) * 2 - 1
    ds = torch.utils.data.TensorDataset(x, torch.zeros(n_images, dtype=torch.long))
    return torch.utils.data.DataLoader(ds, batch_size=batch_size, shuffle=False, pin_memory=torch.cuda.is_available())


_DATA_LOADER_HINT = ("The RSNA loaders are the reference's own src/data_loader.py (not shipped here): put the reference's src/ directory on "
                     'PYTHONPATH (see INTEGRATION.md), or train on synthetic images with --synthetic N.')


def _shard_loader(loader, world, rank, batch_size, workers, seed):
    """Data parallel over the reference's DataLoader (data_loader.py:189-192): the same dataset behind a DistributedSampler
    (shuffle, equal shard sizes), so that each rank walks 1/world of the images per epoch."""
    sampler = torch.utils.data.distributed.DistributedSampler(loader.dataset, num_replicas=world, rank=rank, shuffle=True, seed=seed,
                                                              drop_last=True)
    return torch.utils.data.DataLoader(loader.dataset, batch_size=batch_size, sampler=sampler, num_workers=workers, pin_memory=True)


def _save_image_grid(t, path):
    import torchvision.utils as vutils
    vutils.save_image(t, path, normalize=True, nrow=8)


def _reference_cpu_iteration(netG, netD, optG, optD, criterion, real, noise, real_label=0.9, fake_label=0.0):
    """The reference's stock-torch iteration (:121-150), used for --cpu."""
    b = real.size(0)
    netD.zero_grad()
    label = torch.full((b,), real_label, dtype=torch.float, device=real.device)
    out_real = netD(real).view(-1)
    errD_real = criterion(out_real, label)
    errD_real.backward()
    fake = netG(noise)
    label.fill_(fake_label)
    out_fake = netD(fake.detach()).view(-1)
    errD_fake = criterion(out_fake, label)
    errD_fake.backward()
    errD = errD_real + errD_fake
    optD.step()
    netG.zero_grad()
    label.fill_(real_label)
    out2 = netD(fake).view(-1)
    errG = criterion(out2, label)
    errG.backward()
    optG.step()
    return torch.stack([errD.detach(), errG.detach(), out_real.mean().detach(), out_fake.mean().detach(), out2.mean().detach()])


def main(args):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    use_cuda = torch.cuda.is_available() and not args.cpu
    device = torch.device('cuda', local_rank) if use_cuda else torch.device('cpu')
    if use_cuda:
        torch.cuda.set_device(device)
    if world > 1:
        torch.distributed.init_process_group('nccl' if use_cuda else 'gloo')
    is_main = rank == 0
    if is_main:
        print(f'Using device: {device}' + (f' (data parallel over {world} ranks)' if world > 1 else ''))
    seed = getattr(args, 'seed', None)
    if seed is not None:
        torch.manual_seed(seed)          # common stream: weight initialisation and fixed_noise are the same on every rank

    gan_model_dir = os.path.join(args.model_dir, 'gan')
    gan_output_dir = os.path.join(args.output_dir, 'gan_images')
    if is_main:
        for d in (gan_model_dir, gan_output_dir, args.results_dir, args.figures_dir):
            os.makedirs(d, exist_ok=True)

    # --- data ------------------------------------------------------------------------------------
    n_syn = getattr(args, 'synthetic', 0)
    cached = getattr(args, 'cache_dataset', False)
    if cached and not use_cuda:
        print('Error: --cache-dataset keeps the training images in GPU memory; it needs CUDA (drop the flag for --cpu runs).')
        return
    cache_dtype = {'bf16': torch.bfloat16, 'fp32': torch.float32}[getattr(args, 'dtype', 'bf16')]
    if n_syn and cached:
        g = torch.Generator().manual_seed(1 + rank)
        u8 = torch.randint(0, 256, (n_syn, args.num_channels, 224, 224), dtype=torch.uint8, generator=g)
        half = (0.5,) * args.num_channels
        train_loader = DeviceImageCache(u8.to(device), batch_size=args.batch_size, mean=half, std=half, dtype=cache_dtype,
                                        seed=None if seed is None else seed + rank)
        print(f'Loaded training data with {len(train_loader.dataset)} samples.' if is_main else '', end='\n' if is_main else '')
    elif cached:
        try:
            import torchvision.transforms as T
            from data_loader import RSNAPneumoniaDataset, check_dataset_availability     # the reference's src/data_loader.py:75,119
            if not check_dataset_availability(args.data_dir):
                raise FileNotFoundError(f'Dataset not available in {args.data_dir}. Please download using the provided script.')
            ds = RSNAPneumoniaDataset(os.path.join(args.data_dir, 'Training', 'Images'), os.path.join(args.data_dir, 'stage2_train_metadata.csv'),
                                      transform=T.Resize((224, 224)), is_test=False)
            if world > 1:        # each rank caches and draws from its own slice of the images; equal slices, so that every rank
                ds = torch.utils.data.Subset(ds, range(rank, len(ds) // world * world, world))   # issues the same collectives
            train_loader = DeviceImageCache.from_dataset(ds, device, num_workers=args.workers, batch_size=args.batch_size, dtype=cache_dtype,
                                                         seed=None if seed is None else seed + rank)
            if train_loader.images.shape[1] != args.num_channels:
                print(f'Error: the dataset has {train_loader.images.shape[1]} channels, --num-channels is {args.num_channels}')
                return
            print(f'Loaded training data with {len(train_loader.dataset)} samples (device-resident cache, '
                  f'{train_loader.images.numel() / 2**30:.2f} GiB).')
        except ImportError as e:
            print(f'Error: {e}')
            print(_DATA_LOADER_HINT)
            return
        except FileNotFoundError as e:
            print(f'Error: {e}')
            print(f"Please ensure the dataset exists at '{args.data_dir}' and is structured correctly.")
            print('Run `python src/download_dataset.py` first if needed.')
            return
    elif n_syn:
        train_loader = _synthetic_loader(n_syn, args.num_channels, args.batch_size, 1 + rank)
        print(f'Loaded training data with {len(train_loader.dataset)} samples.' if is_main else '', end='\n' if is_main else '')
    else:
        try:
            from data_loader import get_dataloaders          # the reference's src/data_loader.py:158
            train_loader, _ = get_dataloaders(data_dir=args.data_dir, batch_size=args.batch_size, num_workers=args.workers)
            if world > 1:
                train_loader = _shard_loader(train_loader, world, rank, args.batch_size, args.workers, seed or 0)
            print(f'Loaded training data with {len(train_loader.dataset)} samples.')
        except ImportError as e:
            print(f'Error: {e}')
            print(_DATA_LOADER_HINT)
            return
        except FileNotFoundError as e:
            print(f'Error: {e}')
            print(f"Please ensure the dataset exists at '{args.data_dir}' and is structured correctly.")
            print('Run `python src/download_dataset.py` first if needed.')
            return
        except Exception as e:   # noqa: BLE001  (reference :75-77)
            print(f'Error loading data: {e}')
            return

    if world > 1:
        # every rank must issue the same number of gradient exchanges per epoch, on the same batch shapes
        nb = torch.tensor([len(train_loader), -len(train_loader)], device=device, dtype=torch.int64)
        torch.distributed.all_reduce(nb, op=torch.distributed.ReduceOp.MAX)
        if int(nb[0]) != -int(nb[1]):
            raise RuntimeError(f'data-parallel ranks disagree on the number of batches per epoch ({-int(nb[1])}..{int(nb[0])}): '
                               'shard the dataset into equal parts')

    # --- networks (reference :80-84) ----------------------------------------------------------------
    netG = Generator(args.latent_dim, args.num_channels, args.feature_maps_g).to(device)
    netD = Discriminator(args.num_channels, args.feature_maps_d).to(device)
    netG.apply(weights_init)
    netD.apply(weights_init)
    if world > 1:            # all replicas start from rank 0's initialisation
        for t in list(netG.state_dict().values()) + list(netD.state_dict().values()):
            torch.distributed.broadcast(t, 0)
    if is_main:
        print('Generator Architecture Initialized.')
        print('Discriminator Architecture Initialized.')

    fixed_noise = torch.randn(args.vis_batch_size, args.latent_dim, 1, 1, device=device)
    if seed is not None and world > 1:
        torch.manual_seed(seed + rank)   # from here on (per-step noise, train_gan.py:132) every rank draws its own stream
    trainer = None
    if use_cuda:
        dtype = {'bf16': torch.bfloat16, 'fp32': torch.float32}[getattr(args, 'dtype', 'bf16')]
        netG.compute_dtype = netD.compute_dtype = dtype      # the module-level forwards (visualisation, train_gan.py:166-169) too
        trainer = DCGANTrainer(netG, netD, lr=args.lr, beta1=args.beta1, dtype=dtype, sync_bn=getattr(args, 'sync_bn', False))
    else:
        criterion = nn.BCELoss()
        optimizerD = optim.Adam(netD.parameters(), lr=args.lr, betas=(args.beta1, 0.999))
        optimizerG = optim.Adam(netG.parameters(), lr=args.lr, betas=(args.beta1, 0.999))

    if is_main:
        print('Starting Training Loop...')
    history = {k: [] for k in HISTORY_KEYS}
    log_interval = max(1, getattr(args, 'log_interval', 50))
    max_iters = getattr(args, 'max_iters', 0) or 0
    iters = 0
    start_time = time.time()
    try:
        from tqdm import tqdm
    except ImportError:
        tqdm = None
    stop = False

    for epoch in range(args.epochs):
        epoch_start = time.time()
        pending = []                      # device-side (5,) tensors, not yet synchronised
        epoch_rows = []
        num_batches = len(train_loader)
        if hasattr(getattr(train_loader, 'sampler', None), 'set_epoch'):
            train_loader.sampler.set_epoch(epoch)
        it = enumerate(train_loader)
        bar = tqdm(it, total=num_batches, desc=f'Epoch {epoch + 1}/{args.epochs}', leave=True) if (tqdm and is_main) else it

        def flush():
            if not pending:
                return
            rows = torch.stack(pending).float().cpu().tolist()     # ONE sync for log_interval iterations
            pending.clear()
            for errD, errG, d_x, z1, z2 in rows:
                history['G_losses_iter'].append(errG)
                history['D_losses_iter'].append(errD)
                history['D_x_iter'].append(d_x)
                history['D_G_z1_iter'].append(z1)
                history['D_G_z2_iter'].append(z2)
            epoch_rows.extend(rows)
            if tqdm and is_main:
                errD, errG, d_x, z1, z2 = rows[-1]
                bar.set_postfix({'Loss_D': f'{errD:.4f}', 'Loss_G': f'{errG:.4f}', 'D(x)': f'{d_x:.4f}', 'D(G(z))': f'{z1:.4f}/{z2:.4f}'})

        for i, data in bar:
            real = data[0].to(device, non_blocking=True)
            b = real.size(0)
            noise = torch.randn(b, args.latent_dim, 1, 1, device=device)
            if trainer is not None:
                pending.append(trainer.step(real, noise))
            else:
                pending.append(_reference_cpu_iteration(netG, netD, optimizerG, optimizerD, criterion, real, noise))
            last = (epoch == args.epochs - 1 and i == num_batches - 1) or (max_iters and iters + 1 >= max_iters)
            if (iters % args.save_interval == 0) or last:
                with torch.no_grad():                              # train mode on purpose (reference :166-169)
                    fake_vis = netG(fixed_noise).detach().float().cpu()
                if is_main:
                    _save_image_grid(fake_vis, f'{gan_output_dir}/fake_samples_epoch_{epoch + 1:03d}_iter_{iters:06d}.png')
            iters += 1
            if len(pending) >= log_interval:
                flush()
            if max_iters and iters >= max_iters:
                stop = True
                break
        flush()
        n = max(1, len(epoch_rows))
        history['G_losses_epoch'].append(sum(r[1] for r in epoch_rows) / n)
        history['D_losses_epoch'].append(sum(r[0] for r in epoch_rows) / n)
        if is_main:
            dt = time.time() - epoch_start
            print(f"Epoch {epoch + 1}/{args.epochs} Summary - Time: {dt:.2f}s, Avg Loss_D: {history['D_losses_epoch'][-1]:.4f}, "
                  f"Avg Loss_G: {history['G_losses_epoch'][-1]:.4f}, {len(epoch_rows) * args.batch_size * world / max(dt, 1e-9):.1f} images/s")
            if (epoch + 1) % args.checkpoint_interval == 0 or (epoch + 1) == args.epochs:
                _save_state(netG, os.path.join(gan_model_dir, f'generator_epoch_{epoch + 1:03d}.pth'))
                _save_state(netD, os.path.join(gan_model_dir, f'discriminator_epoch_{epoch + 1:03d}.pth'))
                print(f'Saved checkpoints for epoch {epoch + 1} to {gan_model_dir}')
        if stop:
            break

    if is_main:
        print(f'Training finished in {time.time() - start_time:.2f} seconds.')
        _save_state(netG, os.path.join(gan_model_dir, 'generator_final.pth'))
        _save_state(netD, os.path.join(gan_model_dir, 'discriminator_final.pth'))
        print(f'Saved final models to {gan_model_dir}')
        history_filename = os.path.join(args.results_dir, 'gan_training_history.json')
        try:
            with open(history_filename, 'w') as f:
                json.dump(history, f, indent=4)
            print(f'Saved training history to {history_filename}')
        except Exception as e:   # noqa: BLE001
            print(f'Error saving training history to {history_filename}: {e}')
        plot_gan_losses(history, os.path.join(args.figures_dir, 'gan_loss_curve.png'))
    if world > 1:
        if trainer is not None:
            trainer.close()
        torch.distributed.destroy_process_group()
    return history


def _save_state(net, path):
    # parameters live in the trainer's flat arena: clone so that each file holds plain per-tensor fp32 storage
    torch.save({k: v.detach().clone().cpu() for k, v in net.state_dict().items()}, path)


def build_parser():
    parser = argparse.ArgumentParser(description='Train DCGAN on RSNA Pneumonia Dataset with Enhanced Logging')
    # --- Paths (reference :217-221) --- #
    parser.add_argument('--data-dir', type=str, default='./data/processed', help='Path to the processed dataset directory')
    parser.add_argument('--model-dir', type=str, default='./models', help='Base directory to save model checkpoints (GAN models saved to ./models/gan/)')
    parser.add_argument('--output-dir', type=str, default='./results', help='Base directory for outputs (generated images saved to ./results/gan_images/)')
    parser.add_argument('--results-dir', type=str, default='./results/metrics', help='Directory to save training history JSON (gan_training_history.json)')
    parser.add_argument('--figures-dir', type=str, default='./results/figures', help='Directory to save generated plot images (gan_loss_curve.png)')
    # --- Model Hyperparameters (reference :224-227) --- #
    parser.add_argument('--num-channels', type=int, default=3, help='Number of image channels (3 for RGB)')
    parser.add_argument('--latent-dim', type=int, default=100, help='Size of the latent z vector')
    parser.add_argument('--feature-maps-g', type=int, default=64, help='Base feature maps for Generator')
    parser.add_argument('--feature-maps-d', type=int, default=64, help='Base feature maps for Discriminator')
    # --- Training Hyperparameters (reference :230-234) --- #
    parser.add_argument('--epochs', type=int, default=50, help='Number of training epochs')
    parser.add_argument('--batch-size', type=int, default=128, help='Batch size for training')
    parser.add_argument('--lr', type=float, default=0.0002, help='Learning rate for Adam optimizer')
    parser.add_argument('--beta1', type=float, default=0.5, help='Beta1 hyperparameter for Adam optimizers')
    parser.add_argument('--workers', type=int, default=4, help='Number of data loading workers')
    # --- Logging and Saving (reference :237-240) --- #
    parser.add_argument('--vis-batch-size', type=int, default=64, help='Batch size for generating visualization images')
    parser.add_argument('--save-interval', type=int, default=500, help='Save generated image samples every N iterations')
    parser.add_argument('--checkpoint-interval', type=int, default=10, help='Save model checkpoints every N epochs')
    parser.add_argument('--cpu', action='store_true', help='Force use CPU even if CUDA is available')
    # --- additive flags of the B200 build --- #
    parser.add_argument('--dtype', choices=['bf16', 'fp32'], default='bf16', help='B200 compute mode: bf16 tensor-core path or fp32 parity path')
    parser.add_argument('--synthetic', type=int, default=0, metavar='N', help='train on N synthetic uniform[-1,1] images instead of the RSNA loader')
    parser.add_argument('--cache-dataset', action='store_true', help='decode the training images once into a device-resident uint8 cache; '
                        'shuffle, horizontal flip and normalisation then run on the GPU (CUDA only)')
    parser.add_argument('--max-iters', type=int, default=0, help='stop after this many iterations (0 = run all epochs)')
    parser.add_argument('--log-interval', type=int, default=50, help='flush the device-side history scalars every N iterations')
    parser.add_argument('--seed', type=int, default=None, help='torch.manual_seed (the reference is unseeded)')
    parser.add_argument('--sync-bn', action='store_true', help='data parallel only: BatchNorm statistics over the global batch (one small all-reduce per '
                        'BatchNorm pass) instead of per rank; the iteration then equals one process on the concatenated batch')
    return parser


if __name__ == '__main__':
    args = build_parser().parse_args()
    print('--- Training Arguments ---')
    for k, v in vars(args).items():
        print(f'  {k}: {v}')
    print('-------------------------')
    main(args)
