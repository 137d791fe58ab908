"""StyleGAN2-ADA backbone of IC-GAN: B200 replacements of torch_utils.ops (bias_act, upfirdn2d)."""
