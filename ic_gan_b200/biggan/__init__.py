"""Drop-in BigGAN backbone of IC-GAN (mirrors BigGAN_PyTorch.layers / BigGAN_PyTorch.BigGAN)."""
from . import layers  # noqa: F401
from .model import D_arch, Discriminator, G_arch, G_D, Generator  # noqa: F401
from . import deep  # noqa: F401  (BigGANdeep.py GBlock / DBlock)
