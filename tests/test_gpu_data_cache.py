"""Device-resident image cache: `b200gan_gather_augment` against torchvision's arithmetic restated
with torch on the CPU -- ToTensor (v / 255), RandomHorizontalFlip, Normalize ((x - mean) / std) of src/data_loader.py:17-23.
Integer gather and fp32 arithmetic: the f32 output must be bit-exact."""
import numpy as np
import pytest
import torch

import gan_enhanced_pneumonia_classifier_b200 as pkg
from gan_enhanced_pneumonia_classifier_b200.data_cache import IMAGENET_MEAN, IMAGENET_STD, DeviceImageCache

pytestmark = pytest.mark.gpu


def reference_batch(images_u8, index, flip, mean, std):
    x = images_u8[index].to(torch.float32).div(255)                 # ToTensor
    f = flip.bool()
    x[f] = x[f].flip(-1)                                            # RandomHorizontalFlip where drawn
    m = torch.tensor(mean, dtype=torch.float32).view(1, -1, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, -1, 1, 1)
    return x.sub(m).div(s)                                          # Normalize


@pytest.mark.parametrize('n,c,h,w,b', [(37, 3, 224, 224, 16), (5, 1, 224, 224, 8), (9, 3, 10, 7, 4), (3, 4, 6, 12, 3)])
def test_gather_augment_bit_exact(n, c, h, w, b):
    g = torch.Generator().manual_seed(n * 31 + w)
    imgs = torch.randint(0, 256, (n, c, h, w), dtype=torch.uint8, generator=g)
    mean, std = (IMAGENET_MEAN + (0.5,))[:c], (IMAGENET_STD + (0.25,))[:c]
    cache = DeviceImageCache(imgs.cuda(), batch_size=b, mean=mean, std=std)
    index = torch.randint(0, n, (b,), generator=g)
    flip = torch.randint(0, 2, (b,), generator=g).to(torch.uint8)
    out = cache.batch(index.cuda(), flip.cuda())
    ref = reference_batch(imgs, index, flip, mean, std)
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert torch.equal(out.cpu(), ref), float((out.cpu() - ref).abs().max())
    # no flip vector = never mirrored; bf16 output = the fp32 value rounded once
    cache16 = DeviceImageCache(imgs.cuda(), batch_size=b, mean=mean, std=std, dtype=torch.bfloat16)
    out16 = cache16.batch(index.cuda())
    ref16 = reference_batch(imgs, index, torch.zeros(b, dtype=torch.uint8), mean, std).to(torch.bfloat16)
    assert torch.equal(out16.cpu(), ref16)


def test_epoch_is_a_permutation_with_flips():
    n, b = 50, 16
    g = torch.Generator().manual_seed(3)
    imgs = torch.randint(0, 256, (n, 3, 8, 8), dtype=torch.uint8, generator=g)
    labels = torch.arange(n)
    cache = DeviceImageCache(imgs.cuda(), labels, batch_size=b, seed=11)
    assert len(cache) == 4 and len(cache.dataset) == n
    seen, flips = [], 0
    for i, data in enumerate(cache):                 # the loop shape of src/train_gan.py:121-123
        x, y = data[0], data[1]
        assert x.shape == (min(b, n - i * b), 3, 8, 8)
        assert torch.equal(y, cache.last_index)      # labels travel with their images
        ref = reference_batch(imgs, cache.last_index.cpu(), cache.last_flip.cpu(), IMAGENET_MEAN, IMAGENET_STD)
        assert torch.equal(x.cpu(), ref)
        seen += y.cpu().tolist()
        flips += int(cache.last_flip.sum())
    assert sorted(seen) == list(range(n))            # every image exactly once per epoch
    assert 0 < flips < n                             # p = 0.5 flips
    second = [y.cpu().tolist() for _, y in cache]
    assert sum(second, []) != seen                   # reshuffled each epoch
    dl = DeviceImageCache(imgs.cuda(), batch_size=b, drop_last=True, shuffle=False, flip=False)
    assert len(dl) == 3
    assert torch.equal(next(iter(dl))[0].cpu(), reference_batch(imgs, torch.arange(b), torch.zeros(b, dtype=torch.uint8), IMAGENET_MEAN, IMAGENET_STD))


def test_from_dataset_and_argument_checks():
    class DS(torch.utils.data.Dataset):
        def __init__(self):
            self.a = np.random.RandomState(0).randint(0, 256, (7, 12, 12, 3), dtype=np.uint8)     # HWC like a PIL RGB image

        def __len__(self):
            return 7

        def __getitem__(self, i):
            return self.a[i], i % 2

    ds = DS()
    cache = DeviceImageCache.from_dataset(ds, 'cuda', batch_size=4, shuffle=False, flip=False)
    assert cache.images.shape == (7, 3, 12, 12) and cache.images.is_cuda
    assert torch.equal(cache.images.cpu(), torch.from_numpy(ds.a).permute(0, 3, 1, 2))
    assert cache.labels.cpu().tolist() == [0, 1, 0, 1, 0, 1, 0]
    with pytest.raises(RuntimeError):
        DeviceImageCache(torch.zeros((2, 3, 4, 4), dtype=torch.uint8))         # host tensor: no CPU path
    with pytest.raises(ValueError):
        DeviceImageCache(torch.zeros((2, 3, 4, 4), device='cuda'))             # not uint8
    with pytest.raises(pkg._lib.B200GanError):                                  # the C ABI checks its arguments
        pkg._lib.call('b200gan_gather_augment', None, 0, None, None, None, None, None, None)
