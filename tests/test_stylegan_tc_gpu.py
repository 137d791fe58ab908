"""GPU numerics of the StyleGAN2 kernels added in round 2, each against a plain PyTorch float32 evaluation of the same
op on the same (rounded) inputs: strided / transposed / arbitrary-padding tensor-core convolutions (bf16 and the
split-bf16 float32 mode; forward, dgrad, wgrad), the tiled NHWC upfirdn2d incl. its fused epilogue, and the
modulation / activation elementwise ops with first- and second-order gradients."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20))


def _rel_l2(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))


CONV_CASES = [
    # name, transpose, k, stride, pad, Ci, Co, H
    ("same3x3", False, 3, 1, 1, 32, 64, 16),
    ("pointwise", False, 1, 1, 0, 64, 32, 16),
    ("valid3x3", False, 3, 1, 0, 48, 24, 18),
    ("stride2_odd_input", False, 3, 2, 0, 32, 64, 33),
    ("stride2_big", False, 3, 2, 0, 64, 128, 65),
    ("transposed_stride2", True, 3, 2, 0, 64, 32, 16),
    ("transposed_stride2_small", True, 3, 2, 0, 32, 32, 4),
    ("transposed_stride1", True, 3, 1, 1, 32, 48, 12),
    # RGB-side 1x1 layers: the 128-bit streaming kernels (bf16) / generic small-channel kernels (float32)
    ("torgb_64", False, 1, 1, 0, 64, 3, 32),
    ("torgb_512", False, 1, 1, 0, 512, 3, 8),
    ("torgb_96", False, 1, 1, 0, 96, 3, 16),
    ("fromrgb_64", False, 1, 1, 0, 3, 64, 32),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "f32split"])
def test_tensor_core_conv_family(cuda_device, case, dtype):
    from ic_gan_b200.stylegan2.ops import conv2d_gradfix as CG
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    name, transpose, k, stride, pad, ci, co, H = case
    g = torch.Generator(device=cuda_device).manual_seed(11)
    B = 3
    x = torch.randn(B, ci, H, H, device=cuda_device, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    wshape = (ci, co, k, k) if transpose else (co, ci, k, k)
    w = (torch.randn(wshape, device=cuda_device, generator=g) / (ci * k * k) ** 0.5).to(dtype)
    x.requires_grad_(True); w.requires_grad_(True)
    op = CG.conv_transpose2d if transpose else CG.conv2d
    y = op(x, w, stride=stride, padding=pad)
    xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    ref = (F.conv_transpose2d if transpose else F.conv2d)(xr, wr, stride=stride, padding=pad)
    assert y.shape == ref.shape and y.dtype == dtype
    gy = torch.randn(ref.shape, device=cuda_device, generator=g).to(dtype)
    dx, dw = torch.autograd.grad(y, [x, w], gy)
    dxr, dwr = torch.autograd.grad(ref, [xr, wr], gy.float())
    tol = 2e-2 if dtype == torch.bfloat16 else 2e-4
    errs = {"y": _rel(y, ref), "dx": _rel(dx, dxr), "dw": _rel(dw, dwr)}
    print(f"{name} {dtype}: {errs}")
    assert all(e <= tol for e in errs.values()), errs


FIR_CASES = [
    # up, down, pad(x0,x1,y0,y1), H, C
    (1, 1, (1, 1, 1, 1), 17, 64), (1, 1, (2, 2, 2, 2), 16, 72), (2, 1, (2, 1, 2, 1), 8, 64), (2, 1, (2, 1, 2, 1), 33, 16),
    (1, 2, (1, 1, 1, 1), 32, 64), (1, 2, (0, 0, 0, 0), 18, 8), (1, 1, (-1, 2, 0, 1), 20, 40), (1, 1, (1, 1, 1, 1), 257, 64),
]


def _fir_ref(x, f, up, down, pad, flip, gain):
    N, C, H, W = x.shape
    z = torch.zeros(N, C, H * up, W * up, device=x.device)
    z[:, :, ::up, ::up] = x
    z = F.pad(z, [max(pad[0], 0), max(pad[1], 0), max(pad[2], 0), max(pad[3], 0)])
    z = z[:, :, max(-pad[2], 0): z.shape[2] - max(-pad[3], 0), max(-pad[0], 0): z.shape[3] - max(-pad[1], 0)]
    ff = f * gain
    if not flip:
        ff = ff.flip([0, 1])
    z = F.conv2d(z, ff[None, None].repeat(C, 1, 1, 1), groups=C)
    return z[:, :, ::down, ::down]


@pytest.mark.parametrize("case", FIR_CASES, ids=[f"up{c[0]}down{c[1]}pad{c[2]}_{c[3]}x{c[4]}" for c in FIR_CASES])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_tiled_upfirdn2d(cuda_device, case, dtype):
    from ic_gan_b200.stylegan2.ops import upfirdn2d as U
    up, down, pad, H, C = case
    g = torch.Generator(device=cuda_device).manual_seed(5)
    x = torch.randn(2, C, H, H + 3, device=cuda_device, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    f0 = U.setup_filter([1, 3, 3, 1], device=cuda_device)
    general = f0 * torch.arange(1, 17, device=cuda_device).reshape(4, 4) / 8               # rank 4: 2-D tap loop
    rank1 = torch.outer(torch.tensor([1., 2., 4., 3.]), torch.tensor([2., 1., 5., 3.])).to(cuda_device) / 40  # separable path
    for flip, f in ((False, general), (True, general), (False, rank1), (True, rank1), (False, f0)):
        xg = x.clone().requires_grad_(True)
        y = U.upfirdn2d(xg, f, up=up, down=down, padding=list(pad), flip_filter=flip, gain=up * up)
        ref = _fir_ref(x.float(), f, up, down, pad, flip, up * up)
        assert y.shape == ref.shape and y.is_contiguous(memory_format=torch.channels_last)
        tol = 1e-5 if dtype == torch.float32 else 1e-2
        assert _rel(y, ref) <= tol, (flip, _rel(y, ref))
        gy = torch.randn(ref.shape, device=cuda_device, generator=g)
        (dx,) = torch.autograd.grad(y, xg, gy.to(dtype))
        xr = x.float().requires_grad_(True)
        (dxr,) = torch.autograd.grad(_fir_ref(xr, f, up, down, pad, flip, up * up), xr, gy)
        assert _rel(dx, dxr) <= tol * 2, (flip, "adjoint", _rel(dx, dxr))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_upfirdn2d_fused_epilogue(cuda_device, dtype):
    """icgan_upfirdn2d_nhwc with demodulation scale, noise, bias, lrelu, gain, clamp and the second (re-modulated) output."""
    from ic_gan_b200._lib import call, dt, ptr, stream_ptr
    from ic_gan_b200.stylegan2.ops import upfirdn2d as U
    N, C, H = 3, 64, 17
    g = torch.Generator(device=cuda_device).manual_seed(8)
    x = torch.randn(N, H, H, C, device=cuda_device, generator=g).to(dtype)
    f = U.setup_filter([1, 3, 3, 1], device=cuda_device).contiguous()
    pre = torch.rand(N, C, device=cuda_device, generator=g) + 0.5
    s2 = torch.rand(N, C, device=cuda_device, generator=g) + 0.5
    noise = torch.randn(N, 16, 16, device=cuda_device, generator=g)
    ns = torch.tensor([0.3], device=cuda_device)
    bias = torch.randn(C, device=cuda_device, generator=g)
    y = torch.empty(N, 16, 16, C, device=cuda_device, dtype=dtype)
    y2 = torch.empty_like(y)
    from ic_gan_b200._lib import float_array
    taps = [0.125, 0.375, 0.375, 0.125]
    call("icgan_upfirdn2d_nhwc", ptr(x), ptr(f), ptr(y), N, C, H, H, 1, 1, 1, 1, 1, 1, 0, 4.0, ptr(pre), ptr(noise), ptr(ns), 1,
         ptr(bias), 3, 0.2, 2 ** 0.5, 1.5, ptr(s2), ptr(y2), float_array(taps), float_array(taps), dt(x), stream_ptr())
    base = _fir_ref(x.float().permute(0, 3, 1, 2), f, 1, 1, (1, 1, 1, 1), False, 4.0)
    t = base * pre[:, :, None, None] + 0.3 * noise[:, None] + bias[None, :, None, None]
    ref = (F.leaky_relu(t, 0.2) * 2 ** 0.5).clamp(-1.5, 1.5).permute(0, 2, 3, 1)
    tol = 1e-5 if dtype == torch.float32 else 1.5e-2
    assert _rel(y, ref) <= tol
    assert _rel(y2, ref.to(dtype).float() * s2[:, None, None, :]) <= tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_modulation_ops_first_and_second_order(cuda_device, dtype):
    from ic_gan_b200.stylegan2.ops import elementwise as E
    N, C, H = 3, 64, 9
    g = torch.Generator(device=cuda_device).manual_seed(2)
    x = torch.randn(N, C, H, H, device=cuda_device, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    s = torch.rand(N, C, device=cuda_device, generator=g) + 0.5
    pre = torch.rand(N, C, device=cuda_device, generator=g) + 0.5
    noise = torch.randn(N, 1, H, H, device=cuda_device, generator=g)
    bias = torch.randn(C, device=cuda_device, generator=g)

    def mine(x, s, pre, noise, bias):
        return E.mod_bias_act(E.modulate(x, s), pre=pre, noise=noise, bias=bias, act="lrelu", gain=2 ** 0.5, clamp=2.0)

    def ref(x, s, pre, noise, bias):
        t = (x * s[:, :, None, None]).to(dtype).float() * pre[:, :, None, None] + noise + bias[None, :, None, None]
        return (F.leaky_relu(t, 0.2) * 2 ** 0.5).clamp(-2.0, 2.0)

    args = [x.clone().requires_grad_(True), s.clone().requires_grad_(True), pre.clone().requires_grad_(True),
            noise.clone().requires_grad_(True), bias.clone().requires_grad_(True)]
    rargs = [x.float().clone().requires_grad_(True)] + [a.detach().clone().requires_grad_(True) for a in args[1:]]
    y, yr = mine(*args), ref(*rargs)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert _rel(y, yr) <= tol
    gy = torch.randn(yr.shape, device=cuda_device, generator=g)
    grads = torch.autograd.grad(y, args, gy.to(dtype), create_graph=True)
    rgrads = torch.autograd.grad(yr, rargs, gy, create_graph=True)
    # bf16: the activation / clamp masks are taken from the ROUNDED output (as bias_act.cu does from y), so a handful of
    # elements that round onto a boundary flip their mask: compare in relative L2, not max-abs
    metric = _rel if dtype == torch.float32 else _rel_l2
    for name, a, b in zip("x s pre noise bias".split(), grads, rgrads):
        assert metric(a, b) <= tol * 3, (name, metric(a, b))
    # the fused one-pass backward (taken when no graph is being recorded) against the composed one
    if dtype == torch.bfloat16:
        y2 = mine(*args)
        fused = torch.autograd.grad(y2, args, gy.to(dtype))
        for name, a, b in zip("x s pre noise bias".split(), fused, rgrads):
            assert _rel_l2(a, b) <= tol * 3, ("fused backward", name, _rel_l2(a, b))
    # second order: d/ds and d/dx of |d y / d s|^2 (the path-length pattern)
    sec = torch.autograd.grad(grads[1].square().sum(), [args[0], args[2]])
    rsec = torch.autograd.grad(rgrads[1].square().sum(), [rargs[0], rargs[2]])
    for name, a, b in zip(("x", "pre"), sec, rsec):
        assert metric(a, b) <= tol * 10, ("second order", name, metric(a, b))
