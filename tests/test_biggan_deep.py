"""BigGAN-deep blocks: the oracle and the B200 blocks against golden vectors frozen from the
live reference's BigGANdeep.GBlock / DBlock (oracle/make_golden_deep.py)."""
import functools
import os

import numpy as np
import pytest
import torch

from oracle import biggan_oracle as O
from oracle.make_golden_deep import D_SPEC, G_SPEC, inputs, synth
from tests.helpers import GOLD, rel_l2


BN_FED_BIASES = ("conv1.bias", "conv2.bias", "conv3.bias")  # of the G block: each is followed by a batch norm


@pytest.fixture(scope="module")
def gold():
    d = np.load(os.path.join(GOLD, "biggan_deep_blocks.npz"))
    return {k: torch.from_numpy(d[k]) for k in d.files}


def _g_shapes():
    ci, co, hid, cond = G_SPEC["in_channels"], G_SPEC["out_channels"], G_SPEC["in_channels"] // 4, G_SPEC["cond"]
    sd = {}
    for name, (o, i, k) in {"conv1": (hid, ci, 1), "conv2": (hid, hid, 3), "conv3": (hid, hid, 3), "conv4": (co, hid, 1)}.items():
        sd[f"{name}.weight"], sd[f"{name}.bias"] = torch.zeros(o, i, k, k), torch.zeros(o)
        sd[f"{name}.u0"], sd[f"{name}.sv0"] = torch.zeros(1, o), torch.zeros(1)
    for name, c in {"bn1": ci, "bn2": hid, "bn3": hid, "bn4": hid}.items():
        for part in ("gain", "bias"):
            sd[f"{name}.{part}.weight"] = torch.zeros(c, cond)
            sd[f"{name}.{part}.u0"], sd[f"{name}.{part}.sv0"] = torch.zeros(1, c), torch.zeros(1)
        sd[f"{name}.stored_mean"], sd[f"{name}.stored_var"] = torch.zeros(c), torch.zeros(c)
    return sd


def _d_shapes():
    ci, co, hid = D_SPEC["in_channels"], D_SPEC["out_channels"], D_SPEC["out_channels"] // 4
    sd = {}
    for name, (o, i, k) in {"conv1": (hid, ci, 1), "conv2": (hid, hid, 3), "conv3": (hid, hid, 3), "conv4": (co, hid, 1),
                            "conv_sc": (co - ci, ci, 1)}.items():
        sd[f"{name}.weight"], sd[f"{name}.bias"] = torch.zeros(o, i, k, k), torch.zeros(o)
        sd[f"{name}.u0"], sd[f"{name}.sv0"] = torch.zeros(1, o), torch.zeros(1)
    return sd


def test_oracle_deep_blocks_match_golden(gold):
    cfg = O.BigGANConfig()
    sd = {k: v.clone().requires_grad_(O.is_param(k, v)) for k, v in synth(_g_shapes(), 51).items()}
    x, y = inputs(G_SPEC, 52)
    x.requires_grad_(True); y.requires_grad_(True)
    out = O.g_block_deep(sd, "", x, y, True, cfg, G_SPEC["out_channels"], True)
    out.backward(gold["g_gy"])
    assert (out - gold["g_out"]).abs().max() <= 2e-5 and (x.grad - gold["g_dx"]).abs().max() <= 2e-5
    for k in [k for k in gold if k.startswith("g_grad/")]:
        if not k.endswith(BN_FED_BIASES):  # conv biases feeding a batch norm: analytically zero, rounding noise only
            assert rel_l2(sd[k[7:]].grad, gold[k]) <= 2e-4, k
    sd = {k: v.clone().requires_grad_(O.is_param(k, v)) for k, v in synth(_d_shapes(), 61).items()}
    x, _ = inputs(D_SPEC, 62)
    x.requires_grad_(True)
    out = O.d_block_deep(sd, "", x, True, cfg, True)
    out.backward(gold["d_gy"])
    assert (out - gold["d_out"]).abs().max() <= 2e-5 and (x.grad - gold["d_dx"]).abs().max() <= 2e-5


def _build_g(dev, cdt):
    from ic_gan_b200.biggan import deep, layers
    conv = functools.partial(layers.SNConv2d, kernel_size=3, padding=1, num_svs=1, num_itrs=1, eps=1e-8)
    lin = functools.partial(layers.SNLinear, num_svs=1, num_itrs=1, eps=1e-8, bias=False)
    bn = functools.partial(layers.ccbn, which_linear=lin, input_size=G_SPEC["cond"], norm_style="bn", eps=1e-5)
    blk = deep.GBlock(G_SPEC["in_channels"], G_SPEC["out_channels"], which_conv=conv, which_bn=bn,
                      activation=torch.nn.ReLU(inplace=False), upsample=True)
    blk.load_state_dict(synth(blk.state_dict(), 51), strict=True)
    blk = blk.to(dev)
    for m in blk.modules():
        if isinstance(m, layers.SN):
            m.compute_dtype = cdt
    return blk


def test_deep_block_state_layout_matches_reference():
    from ic_gan_b200.biggan import deep, layers
    conv = functools.partial(layers.SNConv2d, kernel_size=3, padding=1)
    lin = functools.partial(layers.SNLinear, bias=False)
    bn = functools.partial(layers.ccbn, which_linear=lin, input_size=G_SPEC["cond"])
    g = deep.GBlock(G_SPEC["in_channels"], G_SPEC["out_channels"], which_conv=conv, which_bn=bn, upsample=True)
    d = deep.DBlock(D_SPEC["in_channels"], D_SPEC["out_channels"], which_conv=conv, downsample=True)
    assert {k: tuple(v.shape) for k, v in g.state_dict().items()} == {k: tuple(v.shape) for k, v in _g_shapes().items()}
    assert {k: tuple(v.shape) for k, v in d.state_dict().items()} == {k: tuple(v.shape) for k, v in _d_shapes().items()}


@pytest.mark.gpu
@pytest.mark.parametrize("cdt", [torch.float32, torch.bfloat16])
def test_gpu_deep_gblock_matches_reference(cuda_device, gold, cdt):
    blk = _build_g(cuda_device, cdt)
    blk.train()
    x, y = inputs(G_SPEC, 52)
    x = x.to(cuda_device).to(cdt).requires_grad_(True)
    y = y.to(cuda_device).requires_grad_(True)
    out = blk(x, y)
    out.float().backward(gold["g_gy"].to(cuda_device).to(out.dtype).float())
    tol, gtol = (1e-4, 5e-3) if cdt == torch.float32 else (5e-2, 0.15)
    assert (out.float().cpu() - gold["g_out"]).abs().max() <= tol * max(1.0, gold["g_out"].abs().max().item())
    assert rel_l2(x.grad.float(), gold["g_dx"]) <= gtol
    params = dict(blk.named_parameters())
    worst = max(rel_l2(params[k[7:]].grad, gold[k]) for k in gold if k.startswith("g_grad/") and not k.endswith(BN_FED_BIASES))
    print(f"deep GBlock {cdt}: worst parameter-gradient rel-L2 {worst:.3e}")
    assert worst <= gtol * 2


@pytest.mark.gpu
@pytest.mark.parametrize("cdt", [torch.float32, torch.bfloat16])
def test_gpu_deep_dblock_matches_reference(cuda_device, gold, cdt):
    from ic_gan_b200.biggan import deep, layers
    conv = functools.partial(layers.SNConv2d, kernel_size=3, padding=1, num_svs=1, num_itrs=1, eps=1e-8)
    blk = deep.DBlock(D_SPEC["in_channels"], D_SPEC["out_channels"], which_conv=conv, preactivation=True,
                      activation=torch.nn.ReLU(inplace=False), downsample=True)
    blk.load_state_dict(synth(blk.state_dict(), 61), strict=True)
    blk = blk.to(cuda_device)
    for m in blk.modules():
        if isinstance(m, layers.SN):
            m.compute_dtype = cdt
    blk.train()
    x, _ = inputs(D_SPEC, 62)
    x = x.to(cuda_device).requires_grad_(True)
    out = blk(x)
    out.float().backward(gold["d_gy"].to(cuda_device))
    tol, gtol = (1e-4, 5e-3) if cdt == torch.float32 else (5e-2, 0.15)
    assert (out.float().cpu() - gold["d_out"]).abs().max() <= tol * max(1.0, gold["d_out"].abs().max().item())
    assert rel_l2(x.grad.float(), gold["d_dx"]) <= gtol
    params = dict(blk.named_parameters())
    worst = max(rel_l2(params[k[7:]].grad, gold[k]) for k in gold if k.startswith("d_grad/") and gold[k].abs().max() > 1e-6)
    print(f"deep DBlock {cdt}: worst parameter-gradient rel-L2 {worst:.3e}")
    assert worst <= gtol * 2
