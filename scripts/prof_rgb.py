"""A few launches of the RGB-input conv (first layer of D) for ncu: python scripts/prof_rgb.py [B]"""
import sys

import torch

sys.path.insert(0, ".")
from ic_gan_b200 import _lib as L  # noqa: E402
from ic_gan_b200._lib import call, dt, ptr  # noqa: E402

L.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda")
x = torch.randn(B, 256, 256, 3, device=dev).to(torch.bfloat16)
w = torch.zeros(96, 32, device=dev, dtype=torch.bfloat16)
w[:, :27] = torch.randn(96, 27, device=dev) * 0.2
bias = torch.randn(96, device=dev)
y = torch.empty(B, 256, 256, 96, device=dev, dtype=torch.bfloat16)
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call("icgan_conv2d_rgb_tc", ptr(x), ptr(w), None, ptr(bias), ptr(y), B, 256, 256, 3, 96, 3, dt(y), L.ACT_NONE,
         torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rgb conv B={B}: {ms:.3f} ms, output {y.numel() * 2 / ms * 1e-6:.0f} GB/s")
