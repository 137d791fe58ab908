"""Achieved HBM bandwidth of the memory-bound kernels of the G/D step on representative NHWC bf16 shapes (GPU only)."""
import sys

import torch

sys.path.insert(0, ".")
from ic_gan_b200 import _lib as L  # noqa: E402
from ic_gan_b200._lib import call, dt, ptr  # noqa: E402

L.load()
dev = torch.device("cuda")
sp = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(B, H, W, C):
    P = B * H * W
    x = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
    dy = torch.randn_like(x)
    y = torch.empty_like(x)
    yup = torch.empty(B, 2 * H, 2 * W, C, device=dev, dtype=torch.bfloat16)
    dyup = torch.randn(B, 2 * H, 2 * W, C, device=dev).to(torch.bfloat16)
    half = torch.empty(B, H // 2, W // 2, C, device=dev, dtype=torch.bfloat16)
    mean = torch.zeros(C, device=dev)
    invstd = torch.ones(C, device=dev)
    gain = torch.ones(B, C, device=dev)
    bias = torch.zeros(B, C, device=dev)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    ws = torch.zeros(2 * C, device=dev)
    s1, s2 = torch.zeros(B, C, device=dev), torch.zeros(B, C, device=dev)
    m1, m2 = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    nb = x.numel() * 2
    rows = []
    rows.append(("bn_train_stats", nb, timeit(lambda: call("icgan_bn_train_stats", ptr(x), P, C, dt(x), ptr(ws), None, ptr(rm),
                                                             ptr(rv), ptr(mean), ptr(invstd), 1e-5, 0.1, sp()))))
    rows.append(("bn_apply relu", 2 * nb, timeit(lambda: call("icgan_bn_apply", ptr(x), ptr(y), ptr(mean), ptr(invstd), ptr(gain),
                                                                ptr(bias), C, B, H, W, C, 1, 0, dt(x), dt(y), sp()))))
    rows.append(("bn_apply relu+up2", 5 * nb, timeit(lambda: call("icgan_bn_apply", ptr(x), ptr(yup), ptr(mean), ptr(invstd),
                                                                    ptr(gain), ptr(bias), C, B, H, W, C, 1, 1, dt(x), dt(yup), sp()))))
    rows.append(("bn_bwd_reduce", 2 * nb, timeit(lambda: call("icgan_bn_bwd_reduce", ptr(x), ptr(dy), ptr(mean), ptr(invstd),
                                                                ptr(gain), ptr(bias), C, ptr(s1), ptr(s2), B, H, W, C, 1, 0, dt(x),
                                                                dt(dy), sp()))))
    rows.append(("bn_bwd_reduce up2", 5 * nb, timeit(lambda: call("icgan_bn_bwd_reduce", ptr(x), ptr(dyup), ptr(mean), ptr(invstd),
                                                                    ptr(gain), ptr(bias), C, ptr(s1), ptr(s2), B, H, W, C, 1, 1,
                                                                    dt(x), dt(dyup), sp()))))
    rows.append(("bn_bwd_apply", 3 * nb, timeit(lambda: call("icgan_bn_bwd_apply", ptr(x), ptr(dy), ptr(y), ptr(mean), ptr(invstd),
                                                               ptr(gain), ptr(bias), C, ptr(m1), ptr(m2), B, H, W, C, 1, 0, dt(x),
                                                               dt(dy), sp()))))
    rows.append(("bn_bwd_apply up2", 6 * nb, timeit(lambda: call("icgan_bn_bwd_apply", ptr(x), ptr(dyup), ptr(y), ptr(mean),
                                                                   ptr(invstd), ptr(gain), ptr(bias), C, ptr(m1), ptr(m2), B, H, W,
                                                                   C, 1, 1, dt(x), dt(dyup), sp()))))
    rows.append(("relu", 2 * nb, timeit(lambda: call("icgan_relu", ptr(x), ptr(y), x.numel(), dt(x), sp()))))
    rows.append(("relu_bwd", 3 * nb, timeit(lambda: call("icgan_relu_bwd", ptr(dy), ptr(x), ptr(y), x.numel(), dt(x), dt(dy), sp()))))
    rows.append(("pool2 avg", 1.25 * nb, timeit(lambda: call("icgan_pool2", ptr(x), None, ptr(half), B, H // 2, W // 2, C, 0.25, 0,
                                                               dt(x), sp()))))
    rows.append(("unpool2 avg", 1.25 * nb, timeit(lambda: call("icgan_unpool2", ptr(half), None, ptr(y), B, H // 2, W // 2, C, 0.25,
                                                                 0, dt(half), dt(half), sp()))))
    rows.append(("torch add", 3 * nb, timeit(lambda: torch.add(x, dy, out=y))))
    rows.append(("torch copy", 2 * nb, timeit(lambda: y.copy_(x))))
    print(f"== B={B} {H}x{W} C={C}  ({nb / 1e6:.0f} MB per tensor)")
    for name, b, ms in rows:
        print(f"  {name:20s} {ms:8.3f} ms  {b / ms * 1e-6:8.0f} GB/s")


if __name__ == "__main__":
    run(64, 256, 256, 96)
    run(128, 128, 128, 192)
    run(128, 64, 64, 384)
    run(128, 32, 32, 768)
    run(128, 8, 8, 1536)
