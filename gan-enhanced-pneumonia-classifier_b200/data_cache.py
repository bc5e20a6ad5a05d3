"""Device-resident training-image cache: the input side of the training loop.

The reference decodes, resizes, flips and normalises every PNG on the host in every epoch (src/data_loader.py:17-23,102-116,
DataLoader at :189-192, consumed at src/train_gan.py:121-123); at ~5 x 10^4 images/s per GPU that loader is three orders of
magnitude too slow.  `DeviceImageCache` holds the RESIZED uint8 images in HBM (26 684 x 3 x 224 x 224 = 4.0 GB) and produces
each batch with one kernel (`b200gan_gather_augment`): shuffled gather, RandomHorizontalFlip(p=0.5), ToTensor and Normalize with
torchvision's arithmetic.  It iterates like the reference's DataLoader: `for i, data in enumerate(cache): real = data[0]`,
`len(cache)` batches, `cache.dataset` sized.  There is no host fallback: the cache is CUDA only.
"""
import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # reference src/data_loader.py:22
IMAGENET_STD = (0.229, 0.224, 0.225)


class _Sized:
    def __init__(self, n):
        self._n = n

    def __len__(self):
        return self._n


class DeviceImageCache:
    def __init__(self, images_u8: torch.Tensor, labels: Optional[torch.Tensor] = None, batch_size: int = 128,
                 mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD, flip: bool = True, shuffle: bool = True,
                 drop_last: bool = False, dtype: torch.dtype = torch.float32, seed: Optional[int] = None):
        if not images_u8.is_cuda:
            raise RuntimeError('DeviceImageCache holds its images in GPU memory: pass a CUDA uint8 tensor (N, C, H, W)')
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or not 1 <= images_u8.shape[1] <= 4:
            raise ValueError('images_u8 must be a uint8 tensor (N, C, H, W) with 1..4 channels')
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError('dtype must be torch.float32 or torch.bfloat16')
        self.images = images_u8.contiguous()
        n, c = self.images.shape[:2]
        self.labels = labels.to(self.images.device) if labels is not None else torch.zeros(n, dtype=torch.long, device=self.images.device)
        self.batch_size, self.flip, self.shuffle, self.drop_last, self.dtype = int(batch_size), flip, shuffle, drop_last, dtype
        mean, std = list(mean)[:c], list(std)[:c]
        if len(mean) != c or len(std) != c:
            raise ValueError(f'mean / std need {c} entries')
        self._mean = (C.c_float * c)(*mean)
        self._std = (C.c_float * c)(*std)
        self._gen = torch.Generator(device=self.images.device)
        if seed is not None:
            self._gen.manual_seed(seed)
        self.dataset = _Sized(n)
        self.last_index = self.last_flip = None      # of the most recent batch (tests, debugging)

    @classmethod
    def from_dataset(cls, dataset, device, num_workers: int = 0, load_batch: int = 256, **kw):
        """Decode a map-style dataset ONCE.  Items are (image, label) with image a PIL image, a uint8 HWC numpy array or a uint8
        CHW tensor, all of one size (for the reference: RSNAPneumoniaDataset(..., transform=transforms.Resize((224, 224))))."""
        import numpy as np

        def to_u8(img):
            if isinstance(img, torch.Tensor):
                if img.dtype != torch.uint8:
                    raise ValueError('tensor items must be uint8 (C, H, W)')
                return img
            a = np.asarray(img)
            if a.dtype != np.uint8:
                raise ValueError('image items must be 8-bit')
            if a.ndim == 2:
                a = a[:, :, None]
            return torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1)))

        def collate(items):
            return torch.stack([to_u8(i[0]) for i in items]), torch.tensor([int(i[1]) for i in items], dtype=torch.long)

        loader = torch.utils.data.DataLoader(dataset, batch_size=load_batch, shuffle=False, num_workers=num_workers, collate_fn=collate)
        images, labels, at = None, None, 0
        for x, y in loader:
            if images is None:
                images = torch.empty((len(dataset),) + tuple(x.shape[1:]), dtype=torch.uint8, device=device)
                labels = torch.empty(len(dataset), dtype=torch.long, device=device)
            images[at:at + x.shape[0]].copy_(x, non_blocking=False)
            labels[at:at + x.shape[0]].copy_(y)
            at += x.shape[0]
        if images is None:
            raise ValueError('empty dataset')
        return cls(images[:at], labels[:at], **kw)

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def batch(self, index: torch.Tensor, flip: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
        """out[b] = Normalize(ToTensor(hflip_if(flip[b])(images[index[b]]))) as an NCHW tensor of self.dtype."""
        n, c, h, w = self.images.shape
        index = index.to(device=self.images.device, dtype=torch.int64).contiguous()
        b = index.numel()
        if out is None:
            out = torch.empty((b, c, h, w), device=self.images.device, dtype=self.dtype)
        if flip is not None:
            flip = flip.to(device=self.images.device, dtype=torch.uint8).contiguous()
        L.call('b200gan_gather_augment', L.ptr(self.images), n, L.ptr(index), L.ptr(flip), self._mean, self._std,
               C.byref(L.view_nchw(out)), L.stream_ptr())
        return out

    def __iter__(self):
        n, dev = len(self.dataset), self.images.device
        order = torch.randperm(n, device=dev, generator=self._gen) if self.shuffle else torch.arange(n, device=dev)
        for i in range(len(self)):
            index = order[i * self.batch_size:(i + 1) * self.batch_size]
            flip = (torch.rand(index.numel(), device=dev, generator=self._gen) < 0.5).to(torch.uint8) if self.flip else None
            self.last_index, self.last_flip = index, flip
            yield self.batch(index, flip), self.labels[index]
