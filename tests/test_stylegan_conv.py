"""StyleGAN2 conv path (conv2d_gradfix / conv2d_resample / modulated_conv2d): oracle vs golden on CPU; the B200 ops vs the
golden vectors of the live reference on GPU, with first- and second-order gradients (R1, path length)."""
import os

import numpy as np
import pytest
import torch

from oracle import stylegan_ops_oracle as S
from tests.helpers import GOLD

MODS = [("mod_plain", 1, 1, 3), ("mod_up", 2, 1, 3), ("mod_rgb", 1, 0, 1)]


@pytest.fixture(scope="module")
def gold():
    d = np.load(os.path.join(GOLD, "stylegan_conv.npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files}


def test_oracle_conv2d_resample_matches_golden(gold):
    for name, ci, co, k, up, down, pad, hw in S.CONV_SITES:
        y = S.conv2d_resample(gold[f"cr_{name}_x"], gold[f"cr_{name}_w"], gold["f"], up=up, down=down, padding=pad)
        assert y.shape == gold[f"cr_{name}_y"].shape and (y - gold[f"cr_{name}_y"]).abs().max() <= 2e-5, name


def test_oracle_modulated_conv_matches_golden(gold):
    for name, up, pad, k in MODS:
        dem = name != "mod_rgb"
        y = S.modulated_conv2d(gold[name + "_x"], gold[name + "_w"], gold[name + "_s"],
                               gold[name + "_noise"] if dem else None, up=up, padding=pad, resample_filter=gold["f"],
                               demodulate=dem)
        assert (y - gold[name + "_y"]).abs().max() <= 5e-5, name


def _ok(a, b, tol, what):
    err = (a.detach().float().cpu() - b).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), f"{what}: {err:.3e}"


@pytest.mark.gpu
def test_gpu_conv2d_resample(cuda_device, gold):
    from ic_gan_b200.stylegan2.ops import conv2d_resample as CR
    dev = cuda_device
    f = gold["f"].to(dev)
    for name, ci, co, k, up, down, pad, hw in S.CONV_SITES:
        p = f"cr_{name}_"
        x = gold[p + "x"].to(dev).requires_grad_(True)
        w = gold[p + "w"].to(dev).requires_grad_(True)
        y = CR.conv2d_resample(x, w, f=f, up=up, down=down, padding=pad)
        _ok(y, gold[p + "y"], 2e-5, name + " y")
        dx, dw = torch.autograd.grad(y, [x, w], gold[p + "gy"].to(dev), create_graph=True)
        _ok(dx, gold[p + "dx"], 5e-5, name + " dx")
        _ok(dw, gold[p + "dw"], 5e-5, name + " dw")
        (ddw,) = torch.autograd.grad(dx, w, gold[p + "v"].to(dev))
        _ok(ddw, gold[p + "ddw"], 1e-4, name + " second-order dw")


@pytest.mark.gpu
def test_gpu_modulated_conv2d(cuda_device, gold):
    from ic_gan_b200.stylegan2.modconv import modulated_conv2d
    dev = cuda_device
    f = gold["f"].to(dev)
    for name, up, pad, k in MODS:
        dem = name != "mod_rgb"
        x = gold[name + "_x"].to(dev).requires_grad_(True)
        w = gold[name + "_w"].to(dev).requires_grad_(True)
        s = gold[name + "_s"].to(dev).requires_grad_(True)
        noise = gold[name + "_noise"].to(dev) if dem else None
        y = modulated_conv2d(x, w, s, noise=noise, up=up, padding=pad, resample_filter=f, demodulate=dem,
                             fused_modconv=False)
        _ok(y, gold[name + "_y"], 5e-5, name + " y (training form)")
        with torch.no_grad():
            yf = modulated_conv2d(x, w, s, noise=noise, up=up, padding=pad, resample_filter=f, demodulate=dem,
                                  fused_modconv=True)
        _ok(yf, gold[name + "_y"], 5e-5, name + " y (fused inference form)")
        dx, dw, ds = torch.autograd.grad(y, [x, w, s], gold[name + "_gy"].to(dev), create_graph=True)
        _ok(dx, gold[name + "_dx"], 1e-4, name + " dx")
        _ok(dw, gold[name + "_dw"], 1e-4, name + " dw")
        _ok(ds, gold[name + "_ds"], 1e-4, name + " dstyles")
        (dds,) = torch.autograd.grad(dx.square().sum(), s)
        _ok(dds, gold[name + "_dds"], 5e-4, name + " path-length style second order")


@pytest.mark.gpu
def test_gpu_conv_bf16_tensor_core_path(cuda_device):
    """bf16, 3x3, stride 1: conv2d_gradfix dispatches to the tcgen05 kernels (forward, dgrad, wgrad)."""
    from ic_gan_b200.stylegan2.ops import conv2d_gradfix as CG
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device=cuda_device).manual_seed(3)
    x = torch.randn(4, 64, 32, 32, device=cuda_device, generator=g).bfloat16().requires_grad_(True)
    w = (torch.randn(128, 64, 3, 3, device=cuda_device, generator=g) / 24).bfloat16().requires_grad_(True)
    y = CG.conv2d(x, w, padding=1)
    gy = torch.randn_like(y)
    dx, dw = torch.autograd.grad(y, [x, w], gy)
    xr, wr = x.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    yr = F.conv2d(xr, wr, padding=1)
    dxr, dwr = torch.autograd.grad(yr, [xr, wr], gy.float())
    assert (y.float() - yr).abs().max() <= 3e-2 * yr.abs().max()
    assert (dx.float() - dxr).abs().max() <= 3e-2 * dxr.abs().max()
    assert (dw.float() - dwr).abs().max() <= 3e-2 * dwr.abs().max()
