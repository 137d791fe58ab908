"""Device time of the non-local block core (ops.AttentionCoreFn forward + backward) at the cc-256 shapes, fused
(csrc/tc_attn.cu) against the three-GEMM path with fp32 logits in HBM.  Run on a B200:  python scripts/bench_attention.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ic_gan_b200 import ops  # noqa: E402


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    dev = torch.device("cuda:0")
    only = sys.argv[1] if len(sys.argv) > 1 else ""          # substring of a shape name, e.g. "G cc-256"
    modes = (True,) if "--fused-only" in sys.argv else (False, True)
    for name, (B, Q, Kk, d, dv) in {"G cc-256 (C=384)": (128, 4096, 1024, 48, 192),
                                    "D cc-256 (C=192), fake+real": (256, 4096, 1024, 24, 96),
                                    "G ic-128 (C=128)": (256, 4096, 1024, 16, 64)}.items():
        if only and not only.startswith("--") and only not in name:
            continue
        mk = lambda *s: torch.randn(*s, device=dev).bfloat16()
        theta, phi, g, do = mk(B, Q, d), mk(B, Kk, d), mk(B, Kk, dv), mk(B, Q, dv)
        flops_f = 2.0 * B * Q * Kk * (d + dv)
        for fused in modes:
            ops.FUSED_ATTENTION = fused
            with torch.no_grad():
                t_f = timed(lambda: ops.AttentionCoreFn.apply(theta, phi, g))

            def both():
                t, p, v = (x.detach().requires_grad_(True) for x in (theta, phi, g))
                ops.AttentionCoreFn.apply(t, p, v).backward(do)
            t_fb = timed(both)
            print(f"{name:30s} {'fused  ' if fused else 'unfused'}  forward (no grad) {t_f:7.3f} ms "
                  f"({flops_f / t_f / 1e9:6.1f} TFLOP/s)   forward+backward {t_fb:7.3f} ms", flush=True)
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
