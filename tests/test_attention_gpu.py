"""The fused non-local block kernels (csrc/tc_attn.cu: logits in tensor memory) against float32 PyTorch on the same
bf16-rounded operands, at the shapes the BigGAN configs use (BigGAN_PyTorch/layers.py:227-244 with ch = 96 / 64 at
64x64) and at shapes that exercise several tiles per CTA, one key chunk, and the smallest channel counts."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# (B, Q, Kk, d, dv)
SHAPES = [
    (2, 4096, 1024, 48, 192),   # cc-256 generator: C = 384
    (2, 4096, 1024, 24, 96),    # cc-256 discriminator: C = 192
    (3, 1024, 256, 16, 64),     # ic-128 generator (d = 12 padded to 16 ... here 16)
    (1, 256, 128, 8, 32),       # one key chunk, smallest channels
    (5, 4096, 384, 32, 128),    # 160 tiles: some CTAs take two; three key chunks
]


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _reference(theta, phi, g, do):
    t, p, v = (x.float().requires_grad_(True) for x in (theta, phi, g))
    S = t @ p.transpose(1, 2)
    P = torch.softmax(S, -1)
    o = P @ v
    o.backward(do.float())
    return o.detach(), P.detach(), torch.logsumexp(S.detach(), -1) * 1.4426950408889634, t.grad, p.grad, v.grad


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_fused_attention_matches_float32(cuda_device, shape):
    from ic_gan_b200 import ops
    from ic_gan_b200._lib import call, ptr, stream_ptr
    B, Q, Kk, d, dv = shape
    gen = torch.Generator(device=cuda_device).manual_seed(Q + d)
    mk = lambda *s, scale=1.0: (torch.randn(*s, device=cuda_device, generator=gen) * scale).bfloat16()
    theta, phi = mk(B, Q, d, scale=1.5 / d ** 0.25), mk(B, Kk, d, scale=1.5 / d ** 0.25)   # logits of a few units
    g, do = mk(B, Kk, dv), mk(B, Q, dv)
    o_ref, P_ref, lse_ref, dt_ref, dp_ref, dg_ref = _reference(theta, phi, g, do)

    # raw C-ABI entry points first: forward output, log-sum-exp, ds, every gradient
    o = torch.empty(B, Q, dv, device=cuda_device, dtype=torch.bfloat16)
    lse = torch.empty(B, Q, device=cuda_device)
    call("icgan_attn_fwd", ptr(theta), ptr(phi), ptr(g), ptr(o), ptr(lse), B, Q, Kk, d, dv, stream_ptr())
    torch.cuda.synchronize()
    report = {"lse max abs": float((lse - lse_ref).abs().max()), "o": _rel(o, o_ref)}
    o2 = torch.empty_like(o)
    call("icgan_attn_fwd", ptr(theta), ptr(phi), ptr(g), ptr(o2), None, B, Q, Kk, d, dv, stream_ptr())
    report["o without lse == o"] = float((o2.float() - o.float()).abs().max())
    dtheta, dS = torch.empty_like(theta), torch.empty(B, Q, Kk, device=cuda_device, dtype=torch.bfloat16)
    dsum = torch.empty_like(lse)
    call("icgan_attn_bwd_q", ptr(theta), ptr(phi), ptr(g), ptr(o), ptr(do), ptr(lse), ptr(dtheta), ptr(dS), ptr(dsum),
         B, Q, Kk, d, dv, stream_ptr())
    dphi, dg = torch.empty_like(phi), torch.empty_like(g)
    call("icgan_attn_bwd_kv", ptr(theta), ptr(phi), ptr(g), ptr(do), ptr(lse), ptr(dsum), ptr(dphi), ptr(dg),
         B, Q, Kk, d, dv, stream_ptr())
    torch.cuda.synchronize()
    dP_ref = do.float() @ g.float().transpose(1, 2)
    dS_ref = P_ref * (dP_ref - (dP_ref * P_ref).sum(-1, keepdim=True))
    report["dS"] = _rel(dS, dS_ref)
    report["dtheta"] = _rel(dtheta, dt_ref)
    report["dsum"] = _rel(dsum, (do.float() * o.float()).sum(-1))
    report["dphi"] = _rel(dphi, dp_ref)
    report["dg"] = _rel(dg, dg_ref)

    # the autograd function the block calls, fused and unfused
    errs = {}
    for fused in (True, False):
        ops.FUSED_ATTENTION = fused
        try:
            t, p, v = (x.clone().requires_grad_(True) for x in (theta, phi, g))
            out = ops.AttentionCoreFn.apply(t, p, v)
            out.backward(do)
            errs[fused] = {"o": _rel(out, o_ref), "dtheta": _rel(t.grad, dt_ref), "dphi": _rel(p.grad, dp_ref),
                           "dg": _rel(v.grad, dg_ref)}
        finally:
            ops.FUSED_ATTENTION = True
    print(f"\n[attention {shape}] entry points: " + ", ".join(f"{k} {v:.3e}" for k, v in report.items()))
    for fused, e in errs.items():
        print(f"[attention {shape}] {'fused  ' if fused else 'unfused'}: " + ", ".join(f"{k} {v:.3e}" for k, v in e.items()))
    assert report["lse max abs"] < 2e-3
    assert report["o without lse == o"] == 0.0
    assert report["o"] < 6e-3, report                                 # bf16 rounding of the probabilities and of o
    assert report["dS"] < 1.5e-2 and report["dtheta"] < 1.5e-2 and report["dsum"] < 1e-5, report
    assert report["dphi"] < 1.5e-2 and report["dg"] < 1.5e-2, report
    for k, v in errs[True].items():
        assert v < max(1.5e-2, 1.5 * errs[False][k]), (k, errs)        # no worse than the unfused tensor-core path
