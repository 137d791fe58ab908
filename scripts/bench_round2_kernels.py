"""Round-2 kernels on representative shapes (GPU only): achieved HBM GB/s of the StyleGAN2 NHWC kernels against the
algorithmic bytes (each operand once), achieved TFLOP/s of the tap-table tensor-core convolutions (stride 2, transposed
parity classes, sub-pixel up, pooled down).  CUDA events, 3 warm-up + 10 timed launches, tensors far larger than L2.
    python scripts/bench_round2_kernels.py            # table on stdout
    python scripts/bench_round2_kernels.py one NAME   # a single kernel, 3 launches (for `ncu --set full -k regex:...`)"""
import sys

import torch

sys.path.insert(0, ".")
from ic_gan_b200 import _lib as L  # noqa: E402
from ic_gan_b200._lib import call, dt, float_array, int_array, ptr  # noqa: E402

L.load()
dev = torch.device("cuda")
sp = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
PEAK_GBS, PEAK_TF = 6572.5, 1436.0


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def cases():
    out = {}
    N, C, H = 64, 64, 256
    bf = torch.bfloat16
    x = torch.randn(N, H, H, C, device=dev).to(bf)
    x257 = torch.randn(N, H + 1, H + 1, C, device=dev).to(bf)
    y = torch.empty_like(x)
    yh = torch.empty(N, H // 2, H // 2, C, device=dev, dtype=bf)
    y2 = torch.empty_like(x)
    dxb = torch.empty_like(x)
    f = (torch.tensor([1., 3., 3., 1.])[:, None] * torch.tensor([1., 3., 3., 1.])[None, :] / 64).to(dev).contiguous()
    taps = float_array([0.125, 0.375, 0.375, 0.125])
    s = torch.rand(N, C, device=dev) + 0.5
    noise = torch.randn(N, H, H, device=dev)
    bias = torch.randn(C, device=dev)
    dpre, dbn, dnz = torch.empty(N, C, device=dev), torch.empty(N, C, device=dev), torch.empty(N, H, H, device=dev)
    b = x.numel() * 2

    def fir(xin, yout, inH, up, down, pad, sep, post=False):
        return lambda: call("icgan_upfirdn2d_nhwc", ptr(xin), ptr(f), ptr(yout), N, C, inH, inH, up, down, pad, pad, pad, pad, 0,
                            float(up * up), ptr(s) if post else None, ptr(noise) if post else None, None, 1,
                            ptr(bias) if post else None, 3 if post else 0, 0.2, 2 ** 0.5, 256.0, ptr(s) if post else None,
                            ptr(y2) if post else None, taps if sep else None, taps if sep else None, dt(xin), sp())
    out["upfirdn2d blur 257->256 (separable)"] = (fir(x257, y, H + 1, 1, 1, 1, True), x257.numel() * 2 + b, None)
    out["upfirdn2d blur 257->256 (2-D taps)"] = (fir(x257, y, H + 1, 1, 1, 1, False), x257.numel() * 2 + b, None)
    out["upfirdn2d blur + demod/noise/bias/lrelu/clamp + 2nd output"] = (fir(x257, y, H + 1, 1, 1, 1, True, True),
                                                                      x257.numel() * 2 + 2 * b, None)
    out["upfirdn2d down2 256->128"] = (fir(x, yh, H, 1, 2, 1, True), b + yh.numel() * 2, None)
    xh = torch.randn(N, H // 2, H // 2, C, device=dev).to(bf)
    out["upfirdn2d up2 128->256"] = (lambda: call("icgan_upfirdn2d_nhwc", ptr(xh), ptr(f), ptr(y), N, C, H // 2, H // 2, 2, 1, 2, 1,
                                                  2, 1, 0, 4.0, None, None, None, 0, None, 0, 0.0, 1.0, -1.0, None, None, taps,
                                                  taps, dt(xh), sp()), xh.numel() * 2 + b, None)
    out["bias_act_nhwc fwd (demod+noise+bias+lrelu+clamp)"] = (lambda: call(
        "icgan_bias_act_nhwc", ptr(x), None, ptr(y), ptr(bias), ptr(s), ptr(noise), None, 1, N, H * H, C, 0, 3, 0.2, 2 ** 0.5,
        256.0, dt(x), sp()), 2 * b, None)
    out["mod_bias_act_bwd (dx, dpre, dnoise, dbias in one pass)"] = (lambda: call(
        "icgan_mod_bias_act_bwd", ptr(x), ptr(y), ptr(y2), ptr(s), ptr(dxb), ptr(dpre),
        ptr(dbn), ptr(dnz), N, H * H, C, 3, 0.2, 2 ** 0.5, 256.0, dt(x), sp()), 4 * b, None)
    out["modulate x*s[n,c]"] = (lambda: call("icgan_modulate", ptr(x), ptr(s), ptr(y), N, H * H, C, dt(x), dt(y), sp()), 2 * b, None)
    out["chan_dot sum_p a*b"] = (lambda: call("icgan_chan_dot", ptr(x), ptr(y), ptr(dpre), N, H * H, C, dt(x), dt(y), sp()), 2 * b, None)
    rgb = torch.empty(N, H, H, 3, device=dev, dtype=bf)
    w3 = torch.randn(3, 1, 1, C, device=dev)
    out["toRGB 1x1 64->3 (128-bit loads)"] = (lambda: call("icgan_conv2d_small", ptr(x), ptr(w3), None, None, ptr(rgb), N, H, H, C, 3, 1,
                                                           dt(x), dt(rgb), 0, sp()), b + rgb.numel() * 2, None)
    g3 = torch.zeros(3, 1, 1, C, device=dev)
    out["toRGB wgrad 1x1 (128-bit loads)"] = (lambda: call("icgan_conv2d_wgrad_small", ptr(x), ptr(rgb), ptr(g3), N, H, H, C, 3, 1, dt(x),
                                                           dt(rgb), sp()), b + rgb.numel() * 2, None)

    # tensor-core tap-table kernels
    def conv_ex(B, Hin, ci, co, taps_, in_stride, dom, omap, OH):
        xx = torch.randn(B, Hin, Hin, ci, device=dev).to(bf)
        wk = torch.randn(co, 16, ci, device=dev).to(bf)
        yy = torch.empty(B, OH, OH, co, device=dev, dtype=bf)
        fn = lambda: call("icgan_conv2d_tc_ex", ptr(xx), ptr(wk), None, None, None, ptr(yy), B, Hin, Hin, ci, co, 16, len(taps_),
                          int_array([t[0] for t in taps_]), int_array([t[1] for t in taps_]), int_array([t[2] for t in taps_]),
                          in_stride, dom, dom, OH, OH, omap[0], omap[1], omap[2], omap[3], dt(yy), L.F32, 0, sp())
        return fn, None, 2.0 * B * dom * dom * ci * co * len(taps_)
    t16 = [(r - 1, s_ - 1, r * 4 + s_) for r in range(4) for s_ in range(4)]
    out["pooled down-conv 16-tap stride 2, B=128 256^2 96->96 (BigGAN D)"] = conv_ex(128, 256, 96, 96, t16, 2, 128, (1, 0, 1, 0), 128)
    out["pooled down-conv 16-tap stride 2, B=128 64^2 384->384"] = conv_ex(128, 64, 384, 384, t16, 2, 32, (1, 0, 1, 0), 32)
    t4 = [(-1, -1, 0), (-1, 0, 1), (0, -1, 2), (0, 0, 3)]
    out["sub-pixel up-conv phase (4 taps), B=128 128^2 192->96 (BigGAN G)"] = conv_ex(128, 128, 192, 96, t4, 1, 128, (2, 0, 2, 0), 256)
    t9 = [(kh, kw, kh * 3 + kw) for kh in range(3) for kw in range(3)]
    out["stride-2 3x3 conv, B=64 257^2 64->128 (StyleGAN2 D)"] = conv_ex(64, 257, 64, 128, t9, 2, 128, (1, 0, 1, 0), 128)
    t22 = [(0, 0, 0), (0, -1, 2), (-1, 0, 6), (-1, -1, 8)]
    out["transposed stride-2 parity class (4 taps), B=64 128^2 128->64 (StyleGAN2 G)"] = conv_ex(64, 128, 128, 64, t22, 1, 129, (2, 0, 2, 0), 257)
    return out


def main():
    cs = cases()
    if len(sys.argv) > 2 and sys.argv[1] == "one":
        fn = cs[sys.argv[2]][0]
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        return
    print(f"# {torch.cuda.get_device_name()}; peaks: {PEAK_GBS} GB/s copy, {PEAK_TF} TFLOP/s bf16 sustained (MEASURED_PEAKS.json)")
    for name, (fn, nbytes, flops) in cs.items():
        ms = timeit(fn)
        if nbytes is not None:
            gbs = nbytes / ms * 1e-6
            print(f"{name:78s} {ms * 1e3:9.1f} us  {nbytes / 1e6:9.1f} MB  {gbs:8.0f} GB/s  {gbs / PEAK_GBS:5.2f} of HBM peak")
        else:
            tf = flops / ms * 1e-9
            print(f"{name:78s} {ms * 1e3:9.1f} us  {flops / 1e9:9.1f} GF  {tf:8.0f} TF/s  {tf / PEAK_TF:5.2f} of bf16 peak")


if __name__ == "__main__":
    main()
