"""ic_gan_b200 — B200-native (sm_100a) kernels behind IC-GAN's generator/discriminator hot path.

Only the data-parallel G/D step and the kNN conditioning build of facebookresearch/ic_gan are covered;
the host side mirrors the reference's module surface, the math runs in libicgan_b200.so (C ABI, include/icgan_b200.h).
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
