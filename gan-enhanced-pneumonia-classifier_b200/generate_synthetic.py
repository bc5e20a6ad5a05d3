"""Drop-in for the reference's `src/generate_synthetic.py`: same CLI flags and defaults (reference :63-70),
same `generate_images(generator_path, output_dir, num_images, latent_dim, feature_maps_g, batch_size, device)` signature, same
output files (`<output-dir>/synthetic_%05d.png`, values `(G(z) * 0.5) + 0.5`, reference :52-54).

On a CUDA device the eval-mode Generator forward (BatchNorm folded into per-channel scale/shift from the running statistics,
reference :34) runs on the B200 kernels of libb200gan.so; the `x*0.5+0.5 -> uint8` quantisation of a whole batch happens on the
device with torchvision's own rounding (`mul(255).add(0.5).clamp(0,255)`), ONE device-to-host copy per batch follows, and only the
PNG encoding stays on the host.  `--cpu` keeps the reference's stock-torch path.
Additive flags: --num-channels (the reference hard-codes 3, :23), --seed.
"""
import argparse
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if __package__ in (None, ''):
    sys.path.insert(0, os.path.dirname(_HERE))
    from gan_enhanced_pneumonia_classifier_b200.dcgan import Generator
else:
    from .dcgan import Generator


def _save_png(u8_chw, path):
    from PIL import Image
    arr = u8_chw.permute(1, 2, 0).numpy()
    Image.fromarray(arr[:, :, 0] if arr.shape[2] == 1 else arr).save(path)


def generate_images(generator_path, output_dir, num_images, latent_dim, feature_maps_g, batch_size, device, num_channels=3):
    """Generates synthetic images using a trained generator (reference :18-61)."""
    os.makedirs(output_dir, exist_ok=True)
    netG = Generator(latent_dim, num_channels, feature_maps_g).to(device)
    try:
        netG.load_state_dict(torch.load(generator_path, map_location=device))
    except FileNotFoundError:
        print(f'Error: Generator model not found at {generator_path}')
        sys.exit(1)
    except Exception as e:    # noqa: BLE001  (reference behaviour: report and exit)
        print(f'Error loading generator state dict: {e}')
        print('Ensure the Generator class definition matches the saved model.')
        sys.exit(1)
    netG.eval()
    print(f'Generating {num_images} synthetic images...')
    generated = 0
    with torch.no_grad():
        while generated < num_images:
            b = min(batch_size, num_images - generated)
            noise = torch.randn(b, latent_dim, 1, 1, device=device)
            fake = netG(noise)
            # (img * 0.5) + 0.5, then torchvision.utils.save_image's quantisation, on the device for the whole batch
            u8 = fake.mul(0.5).add_(0.5).mul_(255).add_(0.5).clamp_(0, 255).to(torch.uint8).cpu()
            for i in range(b):
                _save_png(u8[i], os.path.join(output_dir, f'synthetic_{generated + 1:05d}.png'))
                generated += 1
            print(f'Generated {generated}/{num_images} images...')
    print(f'Finished generating {generated} images in {output_dir}')
    return generated


def build_parser():
    parser = argparse.ArgumentParser(description='Generate synthetic images using a trained DCGAN generator.')
    parser.add_argument('--model-path', type=str, required=True, help='Path to the trained generator checkpoint (e.g., models/gan/generator_final.pth)')
    parser.add_argument('--output-dir', type=str, default='./data/synthetic', help='Directory to save generated images.')
    parser.add_argument('--num-images', type=int, default=5000, help='Number of synthetic images to generate.')
    parser.add_argument('--latent-dim', type=int, default=100, help='Size of the latent z vector (must match training).')
    parser.add_argument('--feature-maps-g', type=int, default=64, help='Generator base feature maps (must match training).')
    parser.add_argument('--batch-size', type=int, default=64, help='Batch size for generation.')
    parser.add_argument('--cpu', action='store_true', help='Force CPU usage even if CUDA is available.')
    parser.add_argument('--num-channels', type=int, default=3, help='image channels of the checkpoint (the reference hard-codes 3)')
    parser.add_argument('--seed', type=int, default=None, help='torch.manual_seed (the reference is unseeded)')
    return parser


if __name__ == '__main__':
    args = build_parser().parse_args()
    device = torch.device('cpu') if args.cpu else torch.device('cuda:0' if torch.cuda.is_available() else 'cpu')
    print(f'Using device: {device}')
    if args.seed is not None:
        torch.manual_seed(args.seed)
    generate_images(generator_path=args.model_path, output_dir=args.output_dir, num_images=args.num_images, latent_dim=args.latent_dim,
                    feature_maps_g=args.feature_maps_g, batch_size=args.batch_size, device=device, num_channels=args.num_channels)
