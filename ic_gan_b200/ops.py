"""torch.autograd glue over the C ABI: every forward/backward below is one or more libicgan_b200 kernel launches.

Activations are explicit NHWC tensors ``[B, H, W, C]`` (contiguous) of dtype float32 ("parity mode", CUDA-core fp32
kernels) or bfloat16 ("throughput mode", tcgen05 tensor-core kernels, fp32 accumulate/statistics).  Statistics, master
weights, weight gradients and everything of shape [B, C] stay float32.  PyTorch supplies memory, streams and the
autograd tape only.
"""
from __future__ import annotations

from typing import List, Optional

import ctypes
import torch

from . import _lib as L
from ._lib import call, dt, ptr, stream_ptr

Tensor = torch.Tensor


# When bench.py sets PROFILE to a list, every tensor-core conv launch is bracketed by CUDA events on the launching
# stream and recorded as (kernel, algorithmic FLOPs, start, end) — the live source of the roofline numbers.
PROFILE = None


def _timed(kernel: str, flops: float, fn):
    if PROFILE is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    PROFILE.append((kernel, flops, e0, e1))
    return out


def _tc_ok(x: Tensor, cin: int, cout: int, k: int) -> bool:
    return x.dtype == torch.bfloat16 and cin % 16 == 0 and cout % 8 == 0 and k in (1, 3)


# ===================================================================================== spectral-norm state
class SNState:
    """Per-layer device state for spectral norm + the operand copies of the (scaled) weight.

    Restates SN.W_ (BigGAN_PyTorch/layers.py:98-112): one power iteration per forward, sigma = u'^T W v, weight / sigma.
    `weight` is the float32 master parameter in the reference layout (OIHW / [out,in]); `u`, `sv` are the module's
    registered buffers `u0`, `sv0` (same names/shapes as the reference, so checkpoints load unchanged)."""

    def __init__(self, module, kind: str, use_sn: bool = True):
        self.module = module
        self.kind = kind  # "conv" | "linear" | "embed"
        self.use_sn = use_sn
        self.fresh = False
        self.aux = None  # [v(cols) | u_new(rows) | sigma(2) | scratch(2)]
        self.wk_fwd = self.wk_dgrad = self.wk_fwd32 = None
        self.key = None

    # -- buffers -------------------------------------------------------------------------------------------------
    def _ensure(self, compute_dtype):
        w = self.module.weight
        rows, cols = w.shape[0], w[0].numel()
        key = (w.device, w.data_ptr(), compute_dtype, self.module.u0.data_ptr() if self.use_sn else 0)
        if self.key == key:
            return
        self.key = key
        self.rows, self.cols = rows, cols
        self.aux = torch.zeros(cols + rows + 4, device=w.device, dtype=torch.float32)
        self.v = self.aux[:cols]
        self.u_new = self.aux[cols:cols + rows]
        self.sigma = self.aux[cols + rows:cols + rows + 2]
        self.scratch = self.aux[cols + rows + 2:]
        self.sigma[0] = 1.0
        self.sigma[1] = 1.0
        if self.kind == "conv":
            co, ci, k, _ = w.shape
            self.wk_fwd = torch.empty(co, k, k, ci, device=w.device, dtype=compute_dtype)
            self.wk_dgrad = torch.empty(ci, k, k, co, device=w.device, dtype=compute_dtype)
            # image-side layers (Cin=3 / Cout=3) run on the CUDA-core kernel, which reads float32 weights
            self.need32 = compute_dtype != torch.float32 and not (ci % 16 == 0 and co % 8 == 0)
            self.need32d = compute_dtype != torch.float32 and not (co % 16 == 0 and ci % 8 == 0)
            self.wk_fwd32 = torch.empty(co, k, k, ci, device=w.device, dtype=torch.float32) if self.need32 else None
            self.wk_dgrad32 = torch.empty(ci, k, k, co, device=w.device, dtype=torch.float32) if self.need32d else None
        else:
            self.wk_fwd = torch.empty(rows, cols, device=w.device, dtype=torch.float32)

    def descriptor(self) -> L.IcganSnLayer:
        m = self.module
        return L.IcganSnLayer(ptr(m.weight), ptr(m.u0), ptr(self.v), ptr(self.u_new), ptr(self.sigma), ptr(self.scratch),
                              self.rows, self.cols)

    def prepare(self):
        """Scaled operand copies from the master weight (after sigma is known)."""
        w = self.module.weight
        inv = ptr(self.sigma[1:]) if self.use_sn else None
        if self.kind == "conv":
            co, ci, k, _ = w.shape
            call("icgan_sn_prepare_weight", ptr(w), inv, ptr(self.wk_fwd), ptr(self.wk_dgrad), co, ci, k,
                 dt(self.wk_fwd), stream_ptr())
            if self.wk_fwd32 is not None or self.wk_dgrad32 is not None:
                call("icgan_sn_prepare_weight", ptr(w), inv, ptr(self.wk_fwd32), ptr(self.wk_dgrad32), co, ci, k, L.F32,
                     stream_ptr())
        else:
            call("icgan_sn_prepare_weight", ptr(w), inv, ptr(self.wk_fwd), None, self.rows, self.cols, 1, L.F32,
                 stream_ptr())

    def weight_grad(self, G: Tensor) -> Tensor:
        """dL/dW (master layout) from G = dL/d(W/sigma) given in operand layout (float32)."""
        w = self.module.weight
        dW = torch.empty_like(w)
        if self.kind == "conv":
            co, ci, k, _ = w.shape
        else:
            co, ci, k = self.rows, self.cols, 1
        if self.use_sn:
            call("icgan_sn_weight_grad", ptr(G), ptr(w), ptr(self.u_new), ptr(self.v), ptr(self.sigma),
                 ptr(self.scratch), ptr(dW), co, ci, k, stream_ptr())
        else:
            call("icgan_sn_weight_grad", ptr(G), None, None, None, None, None, ptr(dW), co, ci, k, stream_ptr())
        return dW


def refresh_sn(states: List[SNState], training: bool, eps: float, compute_dtype, table_cache: dict) -> None:
    """One batched power-iteration step + operand preparation for all given layers (start of every forward)."""
    if not states:
        return
    for s in states:
        s._ensure(compute_dtype)
    sn_states = [s for s in states if s.use_sn]
    if sn_states:
        key = tuple(s.key for s in sn_states)
        if table_cache.get("key") != key:
            arr = (L.IcganSnLayer * len(sn_states))(*[s.descriptor() for s in sn_states])
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            table_cache["table"] = host.to(sn_states[0].module.weight.device)
            table_cache["key"] = key
            table_cache["max_rows"] = max(s.rows for s in sn_states)
            table_cache["max_cols"] = max(s.cols for s in sn_states)
        call("icgan_sn_power_iteration", ptr(table_cache["table"]), len(sn_states), table_cache["max_rows"],
             table_cache["max_cols"], float(eps), 1 if training else 0, stream_ptr())
        if training:  # sv0 is a log-only buffer (layers.py:108-111)
            with torch.no_grad():
                torch._foreach_copy_([s.module.sv0 for s in sn_states], [s.sigma[:1] for s in sn_states])
    for s in states:
        s.prepare()
        s.fresh = True


# ===================================================================================== convolution
def _conv_forward(x: Tensor, st: SNState, bias, residual, res_shift: int, act: int, out_dtype, dgrad: bool = False):
    B, H, W, cin = x.shape
    if not dgrad:
        wk, wk32 = st.wk_fwd, st.wk_fwd32
    else:
        wk, wk32 = st.wk_dgrad, st.wk_dgrad32
    cout, k = wk.shape[0], wk.shape[1]
    y = torch.empty(B, H, W, cout, device=x.device, dtype=out_dtype)
    rdt = dt(residual) if residual is not None else L.F32
    if _tc_ok(x, cin, cout, k) and wk.dtype == torch.bfloat16:
        _timed("tc_conv_kernel", 2.0 * B * H * W * cout * cin * k * k,
               lambda: call("icgan_conv2d_tc", ptr(x), ptr(wk), ptr(bias), ptr(residual), ptr(y), B, H, W, cin, cout, k,
                            dt(y), rdt, res_shift, act, stream_ptr()))
    elif min(cin, cout) <= 4 and residual is None:
        w32 = wk if wk.dtype == torch.float32 else wk32
        call("icgan_conv2d_small", ptr(x), ptr(w32), ptr(bias), ptr(y), B, H, W, cin, cout, k, dt(x), dt(y), act,
             stream_ptr())
    else:
        w32 = wk if wk.dtype == torch.float32 else wk32
        if w32 is None:
            raise RuntimeError(f"no float32 operand copy for conv {cin}->{cout} (dtype {x.dtype})")
        call("icgan_conv2d_simt", ptr(x), ptr(w32), ptr(bias), ptr(residual), ptr(y), B, H, W, cin, cout, k, 1, k // 2,
             dt(x), dt(y), rdt, res_shift, act, stream_ptr())
    return y


class SNConvFn(torch.autograd.Function):
    """y = act(conv(x, W/sigma) + bias + residual) with stride 1, pad k//2 (layers.SNConv2d.forward, layers.py:144-153).
    `residual` may live at half resolution (res_shift=1): the nearest-upsampled shortcut of GBlock (layers.py:545-552),
    using conv1x1(up(x)) == up(conv1x1(x))."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, st: SNState, res_shift: int, act: int, out_dtype):
        x = x.contiguous()
        y = _conv_forward(x, st, bias, residual, res_shift, act, out_dtype)
        ctx.st, ctx.res_shift, ctx.act = st, res_shift, act
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        ctx.res_dtype = residual.dtype if residual is not None else None
        ctx.save_for_backward(x, y if act != L.ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        st: SNState = ctx.st
        dy = dy.contiguous()
        B, H, W, cout = dy.shape
        cin = x.shape[3]
        k = st.wk_fwd.shape[1] if st.kind == "conv" else 1
        n = dy.numel()
        if ctx.act == L.ACT_RELU:
            g = torch.empty_like(dy)
            call("icgan_relu_bwd", ptr(dy), ptr(y), ptr(g), n, dt(y), dt(g), stream_ptr())
            dy = g
        elif ctx.act == L.ACT_TANH:
            g = torch.empty_like(dy)
            call("icgan_tanh_bwd", ptr(dy), ptr(y), ptr(g), n, dt(y), dt(g), stream_ptr())
            dy = g
        dx = dW = db = dres = None
        if ctx.has_res and ctx.needs_input_grad[3]:
            if ctx.res_shift:
                dres = torch.empty(B, H // 2, W // 2, cout, device=dy.device, dtype=dy.dtype)
                call("icgan_pool2", ptr(dy), None, ptr(dres), B, H // 2, W // 2, cout, 1.0, 0, dt(dy), stream_ptr())
            else:
                dres = dy
            if dres.dtype != ctx.res_dtype:
                dres = dres.to(ctx.res_dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(cout, device=dy.device, dtype=torch.float32)
            call("icgan_channel_sum", ptr(dy), ptr(db), B * H * W, cout, dt(dy), stream_ptr())
        if ctx.needs_input_grad[0]:
            gdt = x.dtype
            dyc = dy if dy.dtype == x.dtype else dy.to(x.dtype)
            dx = _conv_forward(dyc, st, None, None, 0, L.ACT_NONE, gdt, dgrad=True)
        if ctx.needs_input_grad[1]:
            G = torch.zeros(cout, k, k, cin, device=dy.device, dtype=torch.float32)
            dyc = dy if dy.dtype == x.dtype else dy.to(x.dtype)
            if x.dtype == torch.bfloat16 and cin % 16 == 0 and cout % 8 == 0:
                _timed("tc_wgrad_kernel", 2.0 * B * H * W * cout * cin * k * k,
                       lambda: call("icgan_conv2d_wgrad_tc", ptr(x), ptr(dyc), ptr(G), B, H, W, cin, cout, k,
                                    stream_ptr()))
            elif min(cin, cout) <= 4:
                call("icgan_conv2d_wgrad_small", ptr(x), ptr(dy), ptr(G), B, H, W, cin, cout, k, dt(x), dt(dy),
                     stream_ptr())
            else:
                call("icgan_conv2d_wgrad_simt", ptr(x), ptr(dyc), ptr(G), B, H, W, cin, cout, k, 1, k // 2, dt(x),
                     stream_ptr())
            dW = st.weight_grad(G)
        return dx, dW, db, dres, None, None, None, None


# ===================================================================================== linear / embedding
def _gemm(A, B_, Cm, M, N, K, sa, sb, sc, alpha=1.0, alpha_dev=None, beta=0.0, bias=None, batch=1, bstrides=(0, 0, 0)):
    call("icgan_gemm", ptr(A), ptr(B_), ptr(Cm), M, N, K, batch, sa[0], sa[1], bstrides[0], sb[0], sb[1], bstrides[1],
         sc[0], sc[1], bstrides[2], float(alpha), ptr(alpha_dev), float(beta), ptr(bias), dt(A), dt(B_), dt(Cm),
         stream_ptr())


class SNLinearFn(torch.autograd.Function):
    """y = x (W/sigma)^T + b  (layers.SNLinear.forward, layers.py:164-165); all float32."""

    @staticmethod
    def forward(ctx, x, weight, bias, st: SNState):
        x = x.contiguous().float()
        Bn, K = x.shape
        N = weight.shape[0]
        y = torch.empty(Bn, N, device=x.device, dtype=torch.float32)
        # A = x [Bn,K]; B[k][n] = wk[n][k]
        _gemm(x, st.wk_fwd, y, Bn, N, K, (K, 1), (1, K), (N, 1), bias=bias)
        ctx.st, ctx.has_bias = st, bias is not None
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        st: SNState = ctx.st
        dy = dy.contiguous().float()
        Bn, K = x.shape
        N = dy.shape[1]
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _gemm(dy, st.wk_fwd, dx, Bn, K, N, (N, 1), (K, 1), (K, 1))  # dx = dy @ W~
        if ctx.needs_input_grad[1]:
            G = torch.empty(N, K, device=x.device, dtype=torch.float32)
            _gemm(dy, x, G, N, K, Bn, (1, N), (K, 1), (K, 1))  # G = dy^T @ x
            dW = st.weight_grad(G)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(N, device=x.device, dtype=torch.float32)
            call("icgan_channel_sum", ptr(dy), ptr(db), Bn, N, L.F32, stream_ptr())
        return dx, dW, db, None


class SNEmbedFn(torch.autograd.Function):
    """rows of W/sigma (layers.SNEmbedding.forward, layers.py:199-200)."""

    @staticmethod
    def forward(ctx, idx, weight, st: SNState):
        ctx.st = st
        ctx.save_for_backward(idx)
        return st.wk_fwd.index_select(0, idx)

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        st: SNState = ctx.st
        G = torch.zeros(st.rows, st.cols, device=dy.device, dtype=torch.float32)
        G.index_add_(0, idx, dy.float())
        return None, st.weight_grad(G), None


# ===================================================================================== batch norm (+ReLU, +up x2)
class BNActFn(torch.autograd.Function):
    """y = [up2]([relu](batch_norm(x) * gain + bias)) — layers.ccbn.forward (layers.py:398-437, per-sample gain/bias
    [B,C]) and layers.bn.forward (:485-503, shared [C]); F.batch_norm semantics incl. running-stat update."""

    @staticmethod
    def forward(ctx, x, gain, bias, running_mean, running_var, training: bool, eps: float, momentum: float, relu: bool,
                up: bool, out_dtype):
        x = x.contiguous()
        B, H, W, Cc = x.shape
        gain = gain.contiguous().float()
        bias = bias.contiguous().float()
        gstride = Cc if gain.dim() == 2 else 0
        dev = x.device
        if training:
            ws = torch.empty(2 * Cc, device=dev, dtype=torch.float32)
            mean = torch.empty(Cc, device=dev, dtype=torch.float32)
            invstd = torch.empty(Cc, device=dev, dtype=torch.float32)
            call("icgan_bn_train_stats", ptr(x), B * H * W, Cc, dt(x), ptr(ws), ptr(running_mean), ptr(running_var),
                 ptr(mean), ptr(invstd), float(eps), float(momentum), stream_ptr())
        else:
            mean = running_mean.float()
            invstd = torch.rsqrt(running_var.float() + eps)
        s = 2 if up else 1
        y = torch.empty(B, H * s, W * s, Cc, device=dev, dtype=out_dtype)
        call("icgan_bn_apply", ptr(x), ptr(y), ptr(mean), ptr(invstd), ptr(gain), ptr(bias), gstride, B, H, W, Cc,
             int(relu), int(up), dt(x), dt(y), stream_ptr())
        ctx.cfg = (training, relu, up, gstride)
        ctx.save_for_backward(x, gain, bias, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gain, bias, mean, invstd = ctx.saved_tensors
        training, relu, up, gstride = ctx.cfg
        dy = dy.contiguous()
        B, H, W, Cc = x.shape
        dev = x.device
        s1 = torch.empty(B, Cc, device=dev, dtype=torch.float32)
        s2 = torch.empty(B, Cc, device=dev, dtype=torch.float32)
        call("icgan_bn_bwd_reduce", ptr(x), ptr(dy), ptr(mean), ptr(invstd), ptr(gain), ptr(bias), gstride, ptr(s1),
             ptr(s2), B, H, W, Cc, int(relu), int(up), dt(x), dt(dy), stream_ptr())
        dgain = dbias = dx = None
        if ctx.needs_input_grad[1]:
            dgain = s2 if gstride else s2.sum(0)
        if ctx.needs_input_grad[2]:
            dbias = s1 if gstride else s1.sum(0)
        if ctx.needs_input_grad[0]:
            if training:
                inv_p = 1.0 / float(B * H * W)
                m1 = ((gain * s1).sum(0) * inv_p).contiguous()
                m2 = ((gain * s2).sum(0) * inv_p).contiguous()
            else:
                m1 = torch.zeros(Cc, device=dev, dtype=torch.float32)
                m2 = m1
            dx = torch.empty(B, H, W, Cc, device=dev, dtype=dy.dtype)
            call("icgan_bn_bwd_apply", ptr(x), ptr(dy), ptr(dx), ptr(mean), ptr(invstd), ptr(gain), ptr(bias), gstride,
                 ptr(m1), ptr(m2), B, H, W, Cc, int(relu), int(up), dt(x), dt(dy), stream_ptr())
            if dx.dtype != x.dtype:
                dx = dx.to(x.dtype)
        return dx, dgain, dbias, None, None, None, None, None, None, None, None


# ===================================================================================== small NHWC ops
class ReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        y = torch.empty_like(x)
        call("icgan_relu", ptr(x), ptr(y), x.numel(), dt(x), stream_ptr())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        call("icgan_relu_bwd", ptr(dy), ptr(y), ptr(dx), dy.numel(), dt(y), dt(dy), stream_ptr())
        return dx


class Pool2Fn(torch.autograd.Function):
    """mode 0: scale * sum of each 2x2 window (+ add) — nn.AvgPool2d(2) with scale=0.25 (BigGAN.py:528);
    mode 1: F.max_pool2d(x, 2) (layers.py:230-231)."""

    @staticmethod
    def forward(ctx, x, add, scale: float, mode: int):
        x = x.contiguous()
        B, H, W, Cc = x.shape
        y = torch.empty(B, H // 2, W // 2, Cc, device=x.device, dtype=x.dtype)
        if add is not None:
            add = add.contiguous()
        call("icgan_pool2", ptr(x), ptr(add), ptr(y), B, H // 2, W // 2, Cc, float(scale), mode, dt(x), stream_ptr())
        ctx.cfg = (scale, mode, add is not None)
        ctx.save_for_backward(x if mode == 1 else None)
        ctx.xshape = x.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        scale, mode, has_add = ctx.cfg
        dy = dy.contiguous()
        B, H, W, Cc = ctx.xshape
        dx = torch.empty(B, H, W, Cc, device=dy.device, dtype=dy.dtype)
        call("icgan_unpool2", ptr(dy), ptr(x), ptr(dx), B, H // 2, W // 2, Cc, float(scale), mode,
             dt(x) if x is not None else dt(dy), dt(dy), stream_ptr())
        return dx, (dy if has_add else None), None, None


class ReluSumPoolFn(torch.autograd.Function):
    """h[n,c] = sum_hw relu(x[n,hw,c])  (Discriminator.forward, BigGAN.py:624); float32 out."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        B, H, W, Cc = x.shape
        out = torch.empty(B, Cc, device=x.device, dtype=torch.float32)
        call("icgan_relu_sumpool", ptr(x), ptr(out), B, H * W, Cc, dt(x), stream_ptr())
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dh):
        (x,) = ctx.saved_tensors
        B, H, W, Cc = x.shape
        dx = torch.empty_like(x)
        call("icgan_relu_sumpool_bwd", ptr(x), ptr(dh.contiguous().float()), ptr(dx), B, H * W, Cc, dt(x), stream_ptr())
        return dx


class ScaleAddFn(torch.autograd.Function):
    """y = gamma * o + x   (layers.Attention.forward, layers.py:244); gamma is a 0-d float32 parameter."""

    @staticmethod
    def forward(ctx, o, x, gamma):
        o, x = o.contiguous(), x.contiguous()
        y = torch.empty_like(x)
        g = gamma.detach().reshape(1).float().contiguous()
        call("icgan_axpby", ptr(o), ptr(x), ptr(y), 1.0, ptr(g), 1.0, None, x.numel(), dt(x), stream_ptr())
        ctx.save_for_backward(o, g)
        return y

    @staticmethod
    def backward(ctx, dy):
        o, g = ctx.saved_tensors
        dy = dy.contiguous()
        do = dgamma = None
        if ctx.needs_input_grad[0]:
            do = torch.empty_like(dy)
            call("icgan_axpby", ptr(dy), None, ptr(do), 1.0, ptr(g), 0.0, None, dy.numel(), dt(dy), stream_ptr())
        if ctx.needs_input_grad[2]:
            acc = torch.empty(1, device=dy.device, dtype=torch.float32)
            call("icgan_dot", ptr(dy), ptr(o), ptr(acc), dy.numel(), dt(dy), stream_ptr())
            dgamma = acc.reshape(())
        return do, dy, dgamma


def _gemm_tc(A, B_, Cm, M, N, K, batch, a_mn, b_mn, lda, sab, ldb, sbb, ldc, scb, alpha=1.0):
    _timed("tc_gemm_kernel", 2.0 * batch * M * N * K,
           lambda: call("icgan_gemm_tc", ptr(A), ptr(B_), ptr(Cm), M, N, K, batch, int(a_mn), int(b_mn), lda, sab, ldb,
                        sbb, ldc, scb, float(alpha), dt(Cm), stream_ptr()))


class AttentionCoreFn(torch.autograd.Function):
    """o[b,q,:] = sum_k softmax_k(theta[b,q,:] . phi[b,k,:]) * g[b,k,:]   (layers.py:233-243) on NHWC-flattened
    theta [B,Q,d], pooled phi [B,Kk,d], pooled g [B,Kk,dv].  bfloat16 inputs run on the batched tcgen05 GEMM (logits and
    their gradient kept in float32, probabilities in bf16); float32 inputs on the CUDA-core GEMM."""

    @staticmethod
    def forward(ctx, theta, phi, g):
        theta, phi, g = theta.contiguous(), phi.contiguous(), g.contiguous()
        B, Q, d = theta.shape
        Kk, dv = phi.shape[1], g.shape[2]
        tc = theta.dtype == torch.bfloat16 and d % 8 == 0 and dv % 8 == 0 and Kk % 8 == 0
        S = torch.empty(B, Q, Kk, device=theta.device, dtype=torch.float32)
        if tc:
            _gemm_tc(theta, phi, S, Q, Kk, d, B, 0, 0, d, Q * d, d, Kk * d, Kk, Q * Kk)
        else:
            _gemm(theta, phi, S, Q, Kk, d, (d, 1), (1, d), (Kk, 1), batch=B, bstrides=(Q * d, Kk * d, Q * Kk))
        P = torch.empty(B, Q, Kk, device=theta.device, dtype=theta.dtype)
        call("icgan_softmax_rows", ptr(S), ptr(P), B * Q, Kk, dt(S), dt(P), stream_ptr())
        del S
        o = torch.empty(B, Q, dv, device=theta.device, dtype=theta.dtype)
        if tc:
            _gemm_tc(P, g, o, Q, dv, Kk, B, 0, 1, Kk, Q * Kk, dv, Kk * dv, dv, Q * dv)
        else:
            _gemm(P, g, o, Q, dv, Kk, (Kk, 1), (dv, 1), (dv, 1), batch=B, bstrides=(Q * Kk, Kk * dv, Q * dv))
        ctx.tc = tc
        ctx.save_for_backward(theta, phi, g, P)
        return o

    @staticmethod
    def backward(ctx, do):
        theta, phi, g, P = ctx.saved_tensors
        do = do.contiguous()
        B, Q, d = theta.shape
        Kk, dv = phi.shape[1], g.shape[2]
        tc = ctx.tc
        dP = torch.empty(B, Q, Kk, device=P.device, dtype=torch.float32)
        dg = torch.empty_like(g)
        if tc:
            _gemm_tc(do, g, dP, Q, Kk, dv, B, 0, 0, dv, Q * dv, dv, Kk * dv, Kk, Q * Kk)
            _gemm_tc(P, do, dg, Kk, dv, Q, B, 1, 1, Kk, Q * Kk, dv, Q * dv, dv, Kk * dv)
        else:
            _gemm(do, g, dP, Q, Kk, dv, (dv, 1), (1, dv), (Kk, 1), batch=B, bstrides=(Q * dv, Kk * dv, Q * Kk))
            _gemm(P, do, dg, Kk, dv, Q, (1, Kk), (dv, 1), (dv, 1), batch=B, bstrides=(Q * Kk, Q * dv, Kk * dv))
        dS = dP if P.dtype == torch.float32 else torch.empty_like(P)
        call("icgan_softmax_rows_bwd", ptr(P), ptr(dP), ptr(dS), B * Q, Kk, dt(P), dt(dP), dt(dS), stream_ptr())
        dtheta = torch.empty_like(theta)
        dphi = torch.empty_like(phi)
        if tc:
            del dP
            _gemm_tc(dS, phi, dtheta, Q, d, Kk, B, 0, 1, Kk, Q * Kk, d, Kk * d, d, Q * d)
            _gemm_tc(dS, theta, dphi, Kk, d, Q, B, 1, 1, Kk, Q * Kk, d, Q * d, d, Kk * d)
        else:
            _gemm(dS, phi, dtheta, Q, d, Kk, (Kk, 1), (d, 1), (d, 1), batch=B, bstrides=(Q * Kk, Kk * d, Q * d))
            _gemm(dS, theta, dphi, Kk, d, Q, (1, Kk), (d, 1), (d, 1), batch=B, bstrides=(Q * Kk, Q * d, Kk * d))
        return dtheta, dphi, dg
