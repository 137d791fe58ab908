"""StyleGAN2-ADA bias_act / upfirdn2d: oracle vs golden (CPU) and the B200 kernels vs golden incl. 1st/2nd-order
gradients (GPU).  Golden vectors are the reference's own impl='ref' outputs (oracle/make_golden_extra.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import stylegan_ops_oracle as S
from tests.helpers import GOLD

ACTS = list(S.ACTS)
VARIANTS = [("def", {}), ("clamp", dict(gain=1.7, clamp=1.1, alpha=0.3))]


@pytest.fixture(scope="module")
def gold():
    d = np.load(os.path.join(GOLD, "stylegan_ops.npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files}


@pytest.mark.parametrize("act", ACTS)
def test_oracle_bias_act_matches_golden(gold, act):
    for tag, kw in VARIANTS:
        x = gold["ba_x"].clone().requires_grad_(True)
        b = gold["ba_b"].clone().requires_grad_(True)
        y = S.bias_act(x, b, act=act, **kw)
        dx, db = torch.autograd.grad(y, [x, b], gold["ba_gy"])
        k = f"ba_{act}_{tag}"
        assert (y - gold[k + "_y"]).abs().max() <= 1e-6
        assert (dx - gold[k + "_dx"]).abs().max() <= 1e-5
        assert (db - gold[k + "_db"]).abs().max() <= 1e-4


def test_oracle_upfirdn2d_matches_golden(gold):
    f = gold["uf_f"]
    for site in S.UPFIRDN_SITES:
        k = "uf_" + site["name"]
        y = S.upfirdn2d(gold[k + "_x"], f, up=site["up"], down=site["down"], padding=site["padding"], gain=site["gain"])
        assert y.shape == gold[k + "_y"].shape and (y - gold[k + "_y"]).abs().max() <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("act", ACTS)
def test_gpu_bias_act(cuda_device, gold, act):
    from ic_gan_b200.stylegan2.ops import bias_act as B
    for tag, kw in VARIANTS:
        k = f"ba_{act}_{tag}"
        for layout in ("nchw", "channels_last"):
            x = gold["ba_x"].to(cuda_device)
            if layout == "channels_last":
                x = x.contiguous(memory_format=torch.channels_last)
            x = x.clone().requires_grad_(True)
            b = gold["ba_b"].to(cuda_device).clone().requires_grad_(True)
            y = B.bias_act(x, b, act=act, **kw)
            dx, db = torch.autograd.grad(y, [x, b], gold["ba_gy"].to(cuda_device), create_graph=True)
            assert (y.cpu() - gold[k + "_y"]).abs().max() <= 2e-5, (act, tag, layout)
            assert (dx.cpu() - gold[k + "_dx"]).abs().max() <= 5e-5, (act, tag, layout)
            assert (db.cpu() - gold[k + "_db"]).abs().max() <= 5e-4, (act, tag, layout)
            if dx.requires_grad:  # second order: d/dx of <dx, gg>  (R1 / path-length regularisers need it)
                ddx = torch.autograd.grad(dx, x, gold["ba_gg"].to(cuda_device), allow_unused=True)[0]
                ddx = torch.zeros_like(x) if ddx is None else ddx
                assert (ddx.cpu() - gold[k + "_ddx"]).abs().max() <= 2e-4, (act, tag, layout)
    # reduced precision storage (fp16 as in the reference's num_fp16_res layers, bf16 on B200)
    for dtp, tol in ((torch.float16, 2e-2), (torch.bfloat16, 8e-2)):
        y = B.bias_act(gold["ba_x"].to(cuda_device, dtp), gold["ba_b"].to(cuda_device, dtp), act=act)
        assert y.dtype == dtp and (y.float().cpu() - gold[f"ba_{act}_def_y"]).abs().max() <= tol * 4


@pytest.mark.gpu
def test_gpu_upfirdn2d(cuda_device, gold):
    from ic_gan_b200.stylegan2.ops import upfirdn2d as U
    f = U.setup_filter([1, 3, 3, 1], device=cuda_device)
    assert (f.cpu() - gold["uf_f"]).abs().max() == 0
    for site in S.UPFIRDN_SITES:
        k = "uf_" + site["name"]
        for layout in ("nchw", "channels_last"):
            x = gold[k + "_x"].to(cuda_device)
            if layout == "channels_last":
                x = x.contiguous(memory_format=torch.channels_last)
            x = x.clone().requires_grad_(True)
            y = U.upfirdn2d(x, f, up=site["up"], down=site["down"], padding=site["padding"], gain=site["gain"])
            assert y.shape == gold[k + "_y"].shape
            assert (y.cpu() - gold[k + "_y"]).abs().max() <= 1e-5, (site["name"], layout)
            gy = gold[k + "_gy"].to(cuda_device).requires_grad_(True)
            (dx,) = torch.autograd.grad(y, x, gy, create_graph=True)
            assert (dx.cpu() - gold[k + "_dx"]).abs().max() <= 1e-5, (site["name"], layout)
            # double backward (R1 / path length): d<dx, v>/d(gy) is the forward op applied to v
            v = torch.randn_like(dx)
            (ddy,) = torch.autograd.grad(dx, gy, v)
            ref = U.upfirdn2d(v, f, up=site["up"], down=site["down"], padding=site["padding"], gain=site["gain"])
            assert (ddy - ref).abs().max() <= 1e-5
    x = gold["ufh_x"].to(cuda_device)
    assert (U.upsample2d(x, f).cpu() - gold["ufh_up"]).abs().max() <= 1e-5
    assert (U.downsample2d(x, f).cpu() - gold["ufh_down"]).abs().max() <= 1e-5
    assert (U.filter2d(x, f).cpu() - gold["ufh_filt"]).abs().max() <= 1e-5
    y = U.upfirdn2d(x, U.setup_filter([1, 2, 4], device=cuda_device), up=[2, 1], down=[1, 2], padding=[1, 0, 2, 1],
                    flip_filter=True, gain=1.5)
    assert y.shape == gold["ufh_flip"].shape and (y.cpu() - gold["ufh_flip"]).abs().max() <= 1e-5
    # full-size property: upsample then downsample with the same FIR preserves a constant image (DC gain 1)
    big = torch.full((2, 64, 256, 256), 0.75, device=cuda_device, dtype=torch.bfloat16)
    rt = U.downsample2d(U.upsample2d(big, f), f)
    assert rt.shape == big.shape and (rt[:, :, 4:-4, 4:-4].float() - 0.75).abs().max() <= 2e-2
