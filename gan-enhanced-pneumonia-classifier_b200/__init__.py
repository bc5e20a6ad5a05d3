"""B200-native DCGAN training step (drop-in for the hot path of harlanljones/gan-enhanced-pneumonia-classifier).

Layout of this directory (it doubles as the reference's `src/` script directory: put it on sys.path and
`from dcgan import Generator` / `python train_gan.py ...` work as in the reference):

  dcgan.py       drop-in Generator / Discriminator / weights_init      (reference src/dcgan.py)
  train_gan.py   drop-in training CLI                                   (reference src/train_gan.py)
  engine.py      layer sequencing over the C ABI + autograd bridge
  trainer.py     fused training step (flat arenas, fused Adam, data-parallel buckets, CUDA graph)
  data_cache.py  device-resident uint8 image cache + fused gather / flip / normalise (input side of train_gan.py)
  generate_synthetic.py  drop-in sampler CLI                             (reference src/generate_synthetic.py)
  _lib.py        ctypes binding of libb200gan.so
  csrc/          hand-written sm_100a CUDA kernels and the C ABI (include/b200gan.h)
"""
from . import _lib, engine          # noqa: F401
from .dcgan import Discriminator, Generator, weights_init   # noqa: F401

__all__ = ['Generator', 'Discriminator', 'weights_init', 'engine']
