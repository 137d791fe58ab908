"""TEST INFRASTRUCTURE ONLY: a CPU emulation of the libicgan_b200 entry points the StyleGAN2 path (and the fused
non-local block of the BigGAN path) calls, so that the HOST
logic above the C ABI -- autograd closures (first and second order), tap tables of the strided / transposed tensor-core
convolutions, layout handling, split-bf16 sequencing -- can be exercised in the CPU test suite, where no GPU exists.

Each emulator is a direct restatement of the contract written in include/icgan_b200.h (NOT of the kernels), operating on
the same flat buffers the kernels would receive.  `emulated()` patches `_lib.call/ptr/stream_ptr` for the duration of a
test; the product never imports this module, and without the patch every op still raises on CPU tensors."""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

from ic_gan_b200 import _lib

_DT = {0: torch.float32, 1: torch.bfloat16, 2: torch.float16}
_REG = {}


def _ptr(t):
    if t is None:
        return None
    if not (t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))):
        raise RuntimeError(f"emulator: non-dense tensor {tuple(t.shape)} {t.stride()}")
    _REG[t.data_ptr()] = t
    return t.data_ptr()


def _flat(p, dtype=None):
    """Flat view of the buffer behind handle p (from its first element to the end of the tensor that was registered)."""
    if p is None:
        return None
    t = _REG[p]
    out = torch.empty(0, dtype=t.dtype).set_(t.untyped_storage(), t.storage_offset(), (t.numel(),), (1,))
    if dtype is not None:
        assert out.dtype == dtype, (out.dtype, dtype)
    return out


def _ints(a):
    return [int(v) for v in a]


# ------------------------------------------------------------------------------------------------- elementwise
def icgan_modulate(x, s, y, N, hw, C, in_dt, out_dt, stream):
    xv, sv, yv = _flat(x, _DT[in_dt]), _flat(s, torch.float32), _flat(y, _DT[out_dt])
    yv[:N * hw * C].view(N, hw, C).copy_((xv[:N * hw * C].view(N, hw, C).float() * sv[:N * C].view(N, 1, C)).to(yv.dtype))


def icgan_chan_dot(a, b, out, N, hw, C, a_dt, b_dt, stream):
    av, bv = _flat(a, _DT[a_dt])[:N * hw * C].view(N, hw, C).float(), _flat(b, _DT[b_dt])[:N * hw * C].view(N, hw, C).float()
    _flat(out, torch.float32)[:N * C].view(N, C).copy_((av * bv).sum(1))


def icgan_bias_act_nhwc(x, yref, y, bias, pre, noise, ns, noise_per_sample, N, hw, C, grad, act, alpha, gain, clamp, dtype,
                        stream):
    n = N * hw * C
    xv = _flat(x, _DT[dtype])[:n].view(N, hw, C).float()
    pv = None if pre is None else _flat(pre, torch.float32)[:N * C].view(N, 1, C)
    if grad == 0:
        t = xv if pv is None else xv * pv
        if noise is not None:
            nz = _flat(noise, torch.float32)
            nz = nz[:N * hw].view(N, hw, 1) if noise_per_sample else nz[:hw].view(1, hw, 1)
            t = t + nz * (1.0 if ns is None else float(_flat(ns)[0]))
        if bias is not None:
            t = t + _flat(bias, torch.float32)[:C].view(1, 1, C)
        if act == 3:
            t = torch.where(t > 0, t, t * alpha)
        t = t * gain
        if clamp >= 0:
            t = t.clamp(-clamp, clamp)
    else:
        yr = None if yref is None else _flat(yref, _DT[dtype])[:n].view(N, hw, C).float()
        t = xv
        if act == 3:
            t = torch.where(yr > 0, t, t * alpha)
        t = t * gain
        if clamp >= 0:
            t = torch.where((yr > -clamp) & (yr < clamp), t, torch.zeros_like(t))
        if pv is not None:
            t = t * pv
    _flat(y, _DT[dtype])[:n].view(N, hw, C).copy_(t.to(_DT[dtype]))


def icgan_bias_act(x, b, xref, yref, dy, y, n, step_b, size_b, grad, act, alpha, gain, clamp, dtype, stream):
    assert act in (1, 3) and grad in (0, 1), "emulator: linear / lrelu, first order only"
    xv = _flat(x, _DT[dtype])[:n].float()
    idx = (torch.arange(n) // step_b) % size_b
    bv = _flat(b, _DT[dtype])[:size_b].float()[idx] if b is not None else 0.0
    if grad == 0:
        t = xv + bv
        if act == 3:
            t = torch.where(t > 0, t, t * alpha)
        t = t * gain
        if clamp >= 0:
            t = t.clamp(-clamp, clamp)
    else:
        yr = _flat(yref, _DT[dtype])[:n].float() if yref is not None else torch.zeros(n)
        t = xv
        if act == 3:
            t = torch.where(yr > 0, t, t * alpha)
        t = t * gain
        if clamp >= 0:
            t = torch.where((yr > -clamp) & (yr < clamp), t, torch.zeros_like(t))
    _flat(y, _DT[dtype])[:n].copy_(t.to(_DT[dtype]))


# ------------------------------------------------------------------------------------------------- upfirdn2d
def _upfirdn_ref(x, f, upx, upy, downx, downy, px0, px1, py0, py1, flip, gain):
    """x [N,C,H,W] float32; direct definition (zero-stuff, pad/crop, convolve, decimate)."""
    N, C, H, W = x.shape
    x = x.reshape(N, C, H, 1, W, 1)
    x = F.pad(x, [0, upx - 1, 0, 0, 0, upy - 1]).reshape(N, C, H * upy, W * upx)
    x = F.pad(x, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    x = x[:, :, max(-py0, 0): x.shape[2] - max(-py1, 0), max(-px0, 0): x.shape[3] - max(-px1, 0)]
    f = f * gain
    if not flip:
        f = f.flip([0, 1])
    x = F.conv2d(x, f[None, None].repeat(C, 1, 1, 1), groups=C)
    return x[:, :, ::downy, ::downx]


def icgan_upfirdn2d(x, f, y, N, C, inH, inW, fh, fw, upx, upy, downx, downy, px0, px1, py0, py1, flip, gain, cl, dtype,
                    stream):
    xv = _flat(x, _DT[dtype])[:N * C * inH * inW].float()
    xv = xv.view(N, inH, inW, C).permute(0, 3, 1, 2) if cl else xv.view(N, C, inH, inW)
    out = _upfirdn_ref(xv, _flat(f, torch.float32)[:fh * fw].view(fh, fw), upx, upy, downx, downy, px0, px1, py0, py1, flip, gain)
    out = out.permute(0, 2, 3, 1) if cl else out
    _flat(y, _DT[dtype])[:out.numel()].copy_(out.reshape(-1).to(_DT[dtype]))


def icgan_upfirdn2d_nhwc(x, f, y, N, C, inH, inW, up, down, px0, px1, py0, py1, flip, gain, pre, noise, ns, nps, bias, act,
                         alpha, act_gain, clamp, s2, y2, fx_host, fy_host, dtype, stream):
    assert act == 0 and y2 is None, "emulator: the fused epilogue of icgan_upfirdn2d_nhwc is exercised on the GPU only"
    icgan_upfirdn2d(x, f, y, N, C, inH, inW, 4, 4, up, up, down, down, px0, px1, py0, py1, flip, gain, 1, dtype, stream)


# ------------------------------------------------------------------------------------------------- convolutions
def _conv_ex(xv, wk, taps_dh, taps_dw, taps_w, in_stride, Hd, Wd):
    """xv [B,H,W,Ci] float, wk [Co,wt,Ci] float -> [B,Hd,Wd,Co]: the tap-list definition of icgan_conv2d_tc_ex."""
    B, H, W, Ci = xv.shape
    out = torch.zeros(B, Hd, Wd, wk.shape[0])
    pad = 80
    xp = F.pad(xv, [0, 0, pad, pad, pad, pad])
    for dh, dw, wi in zip(taps_dh, taps_dw, taps_w):
        hs = [pad + h * in_stride + dh for h in range(Hd)]
        ws = [pad + w * in_stride + dw for w in range(Wd)]
        hs_c = [min(max(h, 0), xp.shape[1] - 1) for h in hs]
        ws_c = [min(max(w, 0), xp.shape[2] - 1) for w in ws]
        sl = xp[:, hs_c][:, :, ws_c]
        ok_h = torch.tensor([0 <= h < xp.shape[1] for h in hs]).view(1, -1, 1, 1)
        ok_w = torch.tensor([0 <= w < xp.shape[2] for w in ws]).view(1, 1, -1, 1)
        out += torch.einsum("bhwc,oc->bhwo", sl * ok_h * ok_w, wk[:, wi])
    return out


def icgan_conv2d_tc_ex(x, wk, alpha, bias, res, y, B, H, W, Ci, Co, wt, nt, tdh, tdw, tw, in_stride, Hd, Wd, OH, OW, osy, ooy, osx,
                       oox, out_dt, res_dt, res_mask, stream):
    assert Ci % 16 == 0 and Co % 8 == 0 and nt <= 16 and alpha is None and not res_mask
    xv = _flat(x, torch.bfloat16)[:B * H * W * Ci].view(B, H, W, Ci).float()
    wv = _flat(wk, torch.bfloat16)[:Co * wt * Ci].view(Co, wt, Ci).float()
    val = _conv_ex(xv, wv, _ints(tdh)[:nt], _ints(tdw)[:nt], _ints(tw)[:nt], in_stride, Hd, Wd)
    yv = _flat(y, _DT[out_dt])[:B * OH * OW * Co].view(B, OH, OW, Co)
    hs = [h * osy + ooy for h in range(Hd) if h * osy + ooy < OH]
    ws = [w * osx + oox for w in range(Wd) if w * osx + oox < OW]
    val = val[:, :len(hs), :len(ws)]
    if bias is not None:
        val = val + _flat(bias, torch.float32)[:Co]
    if res is not None:
        rv = _flat(res, _DT[res_dt])[:B * OH * OW * Co].view(B, OH, OW, Co).float()
        val = val + rv[:, hs][:, :, ws]
    tmp = yv.clone()
    idx_h, idx_w = torch.tensor(hs), torch.tensor(ws)
    tmp[:, idx_h[:, None], idx_w[None, :]] = val.to(tmp.dtype)
    yv.copy_(tmp)


def icgan_conv2d_tc(x, wk, alpha, bias, res, y, stats, B, H, W, Ci, Co, k, out_dt, res_dt, res_shift, act, stream):
    assert alpha is None and stats is None and res_shift == 0 and act == 0
    p = k // 2
    taps = [(kh - p, kw - p, kh * k + kw) for kh in range(k) for kw in range(k)]
    icgan_conv2d_tc_ex(x, wk, None, bias, res, y, B, H, W, Ci, Co, k * k, len(taps), [t[0] for t in taps], [t[1] for t in taps],
                       [t[2] for t in taps], 1, H, W, H, W, 1, 0, 1, 0, out_dt, res_dt, 0, stream)


def icgan_conv2d_wgrad_tc_ex(a, b, out, B, Ha, Wa, Ca, Hb, Wb, Cb, nt, tdh, tdw, in_stride, stream):
    av = _flat(a, torch.bfloat16)[:B * Ha * Wa * Ca].view(B, Ha, Wa, Ca).float()
    bv = _flat(b, torch.bfloat16)[:B * Hb * Wb * Cb].view(B, Hb, Wb, Cb).float()
    ov = _flat(out, torch.float32)[:Ca * nt * Cb].view(Ca, nt, Cb)
    eye = torch.eye(Cb).view(Cb, 1, Cb)
    for t, (dh, dw) in enumerate(zip(_ints(tdh)[:nt], _ints(tdw)[:nt])):
        shifted = _conv_ex(bv, eye, [dh], [dw], [0], in_stride, Ha, Wa)  # b at the tap-shifted positions
        ov[:, t] += torch.einsum("bhwa,bhwc->ac", av, shifted)


def icgan_conv2d_wgrad_tc(x, dy, dwk, B, H, W, Ci, Co, k, stream):
    p = k // 2
    taps = [(kh - p, kw - p) for kh in range(k) for kw in range(k)]
    icgan_conv2d_wgrad_tc_ex(dy, x, dwk, B, H, W, Co, H, W, Ci, len(taps), [t[0] for t in taps], [t[1] for t in taps], 1,
                             stream)


def _conv_generic(x, wk, y, B, H, W, Ci, Co, k, stride, pad, in_dt, out_dt):
    xv = _flat(x, _DT[in_dt])[:B * H * W * Ci].view(B, H, W, Ci).permute(0, 3, 1, 2).float()
    wv = _flat(wk, torch.float32)[:Co * k * k * Ci].view(Co, k, k, Ci).permute(0, 3, 1, 2)
    out = F.conv2d(xv, wv, stride=stride, padding=pad).permute(0, 2, 3, 1)
    _flat(y, _DT[out_dt])[:out.numel()].copy_(out.reshape(-1).to(_DT[out_dt]))


def icgan_conv2d_simt(x, wk, alpha, bias, res, y, B, H, W, Ci, Co, k, stride, pad, in_dt, out_dt, res_dt, res_shift, act, stream):
    assert alpha is None and bias is None and res is None and act == 0
    _conv_generic(x, wk, y, B, H, W, Ci, Co, k, stride, pad, in_dt, out_dt)


def icgan_conv2d_small(x, wk, alpha, bias, y, B, H, W, Ci, Co, k, in_dt, out_dt, act, stream):
    assert alpha is None and bias is None and act == 0
    _conv_generic(x, wk, y, B, H, W, Ci, Co, k, 1, k // 2, in_dt, out_dt)


def _wgrad_generic(x, dy, dwk, B, H, W, Ci, Co, k, stride, pad, x_dt, dy_dt):
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    xv = _flat(x, _DT[x_dt])[:B * H * W * Ci].view(B, H, W, Ci).permute(0, 3, 1, 2).float().requires_grad_(False)
    gv = _flat(dy, _DT[dy_dt])[:B * Ho * Wo * Co].view(B, Ho, Wo, Co).permute(0, 3, 1, 2).float()
    with torch.enable_grad():
        w = torch.zeros(Co, Ci, k, k, requires_grad=True)
        (F.conv2d(xv, w, stride=stride, padding=pad) * gv).sum().backward()
    _flat(dwk, torch.float32)[:Co * k * k * Ci].view(Co, k, k, Ci).add_(w.grad.permute(0, 2, 3, 1))


def icgan_conv2d_wgrad_simt(x, dy, dwk, B, H, W, Ci, Co, k, stride, pad, in_dt, stream):
    _wgrad_generic(x, dy, dwk, B, H, W, Ci, Co, k, stride, pad, in_dt, in_dt)


def icgan_conv2d_wgrad_small(x, dy, dwk, B, H, W, Ci, Co, k, x_dt, dy_dt, stream):
    _wgrad_generic(x, dy, dwk, B, H, W, Ci, Co, k, 1, k // 2, x_dt, dy_dt)

# ------------------------------------------------------------------------------------------------- fused non-local block
_LOG2E = 1.4426950408889634


def _attn_views(theta, phi, g, B, Q, Kk, d, dv):
    t = _flat(theta, torch.bfloat16)[:B * Q * d].view(B, Q, d).float()
    p = _flat(phi, torch.bfloat16)[:B * Kk * d].view(B, Kk, d).float()
    v = _flat(g, torch.bfloat16)[:B * Kk * dv].view(B, Kk, dv).float()
    return t, p, v


def icgan_attn_fwd(theta, phi, g, o, lse2, B, Q, Kk, d, dv, stream):
    t, p, v = _attn_views(theta, phi, g, B, Q, Kk, d, dv)
    S = t @ p.transpose(1, 2)
    _flat(o, torch.bfloat16)[:B * Q * dv].view(B, Q, dv).copy_((torch.softmax(S, -1) @ v).to(torch.bfloat16))
    if lse2 is not None:
        _flat(lse2, torch.float32)[:B * Q].view(B, Q).copy_(torch.logsumexp(S, -1) * _LOG2E)


def _attn_ds(theta, phi, g, dout, lse2, dsum, B, Q, Kk, d, dv):
    t, p, v = _attn_views(theta, phi, g, B, Q, Kk, d, dv)
    do = _flat(dout, torch.bfloat16)[:B * Q * dv].view(B, Q, dv).float()
    lse = _flat(lse2, torch.float32)[:B * Q].view(B, Q, 1)
    P = torch.exp2((t @ p.transpose(1, 2)) * _LOG2E - lse)
    dS = P * (do @ v.transpose(1, 2) - _flat(dsum, torch.float32)[:B * Q].view(B, Q, 1))
    return t, p, do, P, dS


def icgan_attn_bwd_q(theta, phi, g, o, dout, lse2, dtheta, ds, dsum, B, Q, Kk, d, dv, stream):
    ov = _flat(o, torch.bfloat16)[:B * Q * dv].view(B, Q, dv).float()
    do = _flat(dout, torch.bfloat16)[:B * Q * dv].view(B, Q, dv).float()
    _flat(dsum, torch.float32)[:B * Q].view(B, Q).copy_((do * ov).sum(-1))     # the pre-pass of this call
    t, p, do, P, dS = _attn_ds(theta, phi, g, dout, lse2, dsum, B, Q, Kk, d, dv)
    _flat(dtheta, torch.bfloat16)[:B * Q * d].view(B, Q, d).copy_((dS @ p).to(torch.bfloat16))
    if ds is not None:
        _flat(ds, torch.bfloat16)[:B * Q * Kk].view(B, Q, Kk).copy_(dS.to(torch.bfloat16))


def icgan_attn_bwd_kv(theta, phi, g, dout, lse2, dsum, dphi, dg, B, Q, Kk, d, dv, stream):
    t, p, do, P, dS = _attn_ds(theta, phi, g, dout, lse2, dsum, B, Q, Kk, d, dv)
    _flat(dphi, torch.bfloat16)[:B * Kk * d].view(B, Kk, d).copy_((dS.transpose(1, 2) @ t).to(torch.bfloat16))
    _flat(dg, torch.bfloat16)[:B * Kk * dv].view(B, Kk, dv).copy_((P.transpose(1, 2) @ do).to(torch.bfloat16))


EMULATED = {f.__name__: f for f in (
    icgan_modulate, icgan_chan_dot, icgan_bias_act_nhwc, icgan_bias_act, icgan_upfirdn2d, icgan_upfirdn2d_nhwc,
    icgan_conv2d_tc_ex, icgan_conv2d_tc, icgan_conv2d_wgrad_tc_ex, icgan_conv2d_wgrad_tc, icgan_conv2d_simt,
    icgan_conv2d_small, icgan_conv2d_wgrad_simt, icgan_conv2d_wgrad_small, icgan_attn_fwd, icgan_attn_bwd_q,
    icgan_attn_bwd_kv)}


@contextlib.contextmanager
def emulated(monkeypatch):
    """Route `_lib.call` to the emulators above (every module that imported call/ptr/stream_ptr by name is patched)."""
    import ic_gan_b200.stylegan2.ops.bias_act as m1
    import ic_gan_b200.stylegan2.ops.conv2d_gradfix as m2
    import ic_gan_b200.stylegan2.ops.elementwise as m3
    import ic_gan_b200.stylegan2.ops.upfirdn2d as m4
    import ic_gan_b200.ops as m5

    def call(name, *args):
        if name not in EMULATED:
            raise RuntimeError(f"emulator: {name} is not emulated")
        with torch.no_grad():
            EMULATED[name](*args)

    for m in (_lib, m1, m2, m3, m4, m5):
        for attr, fn in (("call", call), ("ptr", _ptr), ("stream_ptr", lambda: None)):
            if hasattr(m, attr):
                monkeypatch.setattr(m, attr, fn)
    try:
        yield
    finally:
        _REG.clear()
