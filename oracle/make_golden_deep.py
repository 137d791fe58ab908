"""Golden vectors for the BigGAN-deep blocks from the LIVE reference (build container only; needs /root/reference):
   python oracle/make_golden_deep.py  ->  tests/golden/biggan_deep_blocks.npz
BigGANdeep.GBlock (ccbn, x2 upsampling, channel-dropping shortcut) and BigGANdeep.DBlock (down-sampling, concatenating
shortcut), training mode, forward + gradients w.r.t. the input, the conditioning vector and every parameter.  The
oracle's g_block_deep / d_block_deep are asserted equal to the reference first."""
from __future__ import annotations

import contextlib
import functools
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ICGAN_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from oracle import biggan_oracle as O  # noqa: E402

G_SPEC = dict(in_channels=64, out_channels=32, cond=24, H=8, B=3, upsample=True)
D_SPEC = dict(in_channels=32, out_channels=64, H=16, B=3, downsample=True)


def synth(sd, seed):
    """Deterministic weights for a block state_dict (same generator on the test side)."""
    out = {}
    for i, (k, v) in enumerate(sorted(sd.items())):
        rng = np.random.default_rng([seed, i])
        if k.endswith("weight"):
            fan = int(np.prod(v.shape[1:])) if v.dim() > 1 else v.shape[0]
            a = rng.standard_normal(v.shape) / np.sqrt(fan)
        elif k.endswith("bias"):
            a = rng.standard_normal(v.shape) * 0.1
        elif k.endswith("stored_var"):
            a = rng.uniform(0.5, 1.5, v.shape)
        elif k.endswith("sv0"):
            a = np.ones(v.shape)
        else:
            a = rng.standard_normal(v.shape) * (0.1 if k.endswith("stored_mean") else 1.0)
        out[k] = torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(v.shape))
    return out


def inputs(spec, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(spec["B"], spec["in_channels"], spec["H"], spec["H"], generator=g)
    y = torch.randn(spec["B"], spec.get("cond", 1), generator=g)
    return x, y


def main():
    sys.path[:0] = [REF, os.path.join(REF, "BigGAN_PyTorch")]  # BigGANdeep.py does a bare `import layers`
    with contextlib.redirect_stdout(io.StringIO()):
        import BigGANdeep as RD
        import layers as RL
    torch.manual_seed(0)
    cfg = O.BigGANConfig()
    conv = functools.partial(RL.SNConv2d, kernel_size=3, padding=1, num_svs=1, num_itrs=1, eps=cfg.SN_eps)
    lin = functools.partial(RL.SNLinear, num_svs=1, num_itrs=1, eps=cfg.SN_eps, bias=False)
    bn = functools.partial(RL.ccbn, which_linear=lin, input_size=G_SPEC["cond"], norm_style="bn", eps=cfg.BN_eps)
    fx = {}

    gb = RD.GBlock(G_SPEC["in_channels"], G_SPEC["out_channels"], which_conv=conv, which_bn=bn,
                   activation=torch.nn.ReLU(inplace=False),
                   upsample=functools.partial(torch.nn.functional.interpolate, scale_factor=2))
    sd = synth(gb.state_dict(), 51)
    gb.load_state_dict(sd); gb.train()
    x, y = inputs(G_SPEC, 52)
    x.requires_grad_(True); y.requires_grad_(True)
    out = gb(x, y)
    gy = torch.randn(out.shape, generator=torch.Generator().manual_seed(53))
    out.backward(gy)
    osd = {k: v.clone().requires_grad_(O.is_param(k, v)) for k, v in sd.items()}
    xo, yo = x.detach().clone().requires_grad_(True), y.detach().clone().requires_grad_(True)
    oo = O.g_block_deep(osd, "", xo, yo, True, cfg, G_SPEC["out_channels"], True)
    oo.backward(gy)
    assert (oo - out).abs().max() <= 2e-5 and (xo.grad - x.grad).abs().max() <= 2e-5
    for k, p in gb.named_parameters():
        assert (osd[k].grad - p.grad).abs().max() <= 2e-4 * max(1.0, p.grad.abs().max().item()), k
    fx.update({"g_out": out.detach(), "g_gy": gy, "g_dx": x.grad, "g_dy": y.grad})
    fx.update({"g_grad/" + k: p.grad for k, p in gb.named_parameters()})
    fx.update({"g_buf/" + k: v.clone() for k, v in gb.state_dict().items() if not O.is_param(k, v)})

    db = RD.DBlock(D_SPEC["in_channels"], D_SPEC["out_channels"], which_conv=conv, preactivation=True,
                   activation=torch.nn.ReLU(inplace=False), downsample=torch.nn.AvgPool2d(2))
    sd = synth(db.state_dict(), 61)
    db.load_state_dict(sd); db.train()
    x, _ = inputs(D_SPEC, 62)
    x.requires_grad_(True)
    out = db(x)
    gy = torch.randn(out.shape, generator=torch.Generator().manual_seed(63))
    out.backward(gy)
    osd = {k: v.clone().requires_grad_(O.is_param(k, v)) for k, v in sd.items()}
    xo = x.detach().clone().requires_grad_(True)
    oo = O.d_block_deep(osd, "", xo, True, cfg, True)
    oo.backward(gy)
    assert (oo - out).abs().max() <= 2e-5 and (xo.grad - x.grad).abs().max() <= 2e-5
    for k, p in db.named_parameters():
        assert (osd[k].grad - p.grad).abs().max() <= 2e-4 * max(1.0, p.grad.abs().max().item()), k
    fx.update({"d_out": out.detach(), "d_gy": gy, "d_dx": x.grad})
    fx.update({"d_grad/" + k: p.grad for k, p in db.named_parameters()})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "biggan_deep_blocks.npz"),
                        **{k: v.detach().numpy() for k, v in fx.items()})
    print("[golden] biggan_deep_blocks: oracle == reference (GBlock with ccbn + upsample, DBlock with downsample + concat shortcut)")


if __name__ == "__main__":
    main()
