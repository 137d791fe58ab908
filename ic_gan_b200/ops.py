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
# Accumulating batch-norm statistics in the producing conv's epilogue is implemented (icgan_conv2d_tc bn_stats) but OFF by
# default: measured on B200 it lowers the conv kernels more (730 -> 654 TFLOP/s average, they are epilogue/L2-bound at
# high resolution) than the two saved statistics passes gain.
FUSE_BN_STATS = False


def _timed(kernel: str, flops: float, fn):
    if PROFILE is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    PROFILE.append((kernel, flops, e0, e1, _SHAPE[0]))
    return out


_SHAPE = [None]  # set by the conv wrappers so bench.py can attribute time to layer shapes


# ===================================================================================== spectral-norm state
# Operand copies of a conv weight are rebuilt when the master weight may have changed.  Signals: the tensor version
# counter (load_state_dict, in-place ops), a per-parameter counter bumped by a post-step hook on every optimiser that
# owns it (fused optimisers do not bump tensor versions), eval mode (always rebuild: EMA copies are written through
# .data), or invalidate_operands() with no argument (everything).
_WEIGHT_EPOCH = [0]
_PARAM_EPOCH = {}  # id(parameter) -> number of optimiser steps that have touched it


def invalidate_operands(*args, **_kwargs) -> None:
    """Optimiser post-step hook (args[0] = the optimiser: only ITS parameters are stale -- G's operands survive D's step)
    or, called with no optimiser, a global invalidation."""
    opt = args[0] if args and hasattr(args[0], "param_groups") else None
    if opt is None:
        _WEIGHT_EPOCH[0] += 1
        return
    for group in opt.param_groups:
        for p in group["params"]:
            _PARAM_EPOCH[id(p)] = _PARAM_EPOCH.get(id(p), 0) + 1


def weight_stamp(w) -> tuple:
    return (w._version, _WEIGHT_EPOCH[0], _PARAM_EPOCH.get(id(w), 0))


from torch.optim.optimizer import register_optimizer_step_post_hook as _register_step_hook  # noqa: E402

_register_step_hook(invalidate_operands)


class SNState:
    """Per-layer device state for spectral norm + the operand copies of the weight the conv kernels read.

    Restates SN.W_ (BigGAN_PyTorch/layers.py:98-112): one power iteration per forward, sigma = u'^T W v, weight / sigma.
    The division is folded into the kernels' epilogues as a device scalar alpha = 1/sigma, so the (unscaled) operand
    copies are rebuilt only when the master weight actually changed (optimizer step / load_state_dict), tracked with
    the tensor version counter.  `weight` is the float32 master parameter in the reference layout (OIHW / [out,in]);
    `u0`, `sv0` are the module's registered buffers (same names/shapes as the reference)."""

    def __init__(self, module, kind: str, use_sn: bool = True):
        self.module = module
        self.kind = kind  # "conv" | "linear" | "embed"
        self.use_sn = use_sn
        self.fresh = False
        self.key = None
        self.version = None
        self.co_pad = 0

    # -- buffers -------------------------------------------------------------------------------------------------
    def _ensure(self, compute_dtype):
        w = self.module.weight
        rows, cols = w.shape[0], w[0].numel()
        key = (w.device, w.data_ptr(), compute_dtype, self.module.u0.data_ptr() if self.use_sn else 0)
        if self.key == key:
            return
        self.key, self.version = key, None
        self.rows, self.cols = rows, cols
        self.bind_aux(torch.zeros(cols + rows + 4, device=w.device, dtype=torch.float32))
        self.sigma.fill_(1.0)
        self.wk_fwd = self.wk_dgrad = self.wk_fwd32 = self.wk_dgrad32 = None
        self.mode = self.mode_d = "plain"
        if self.kind != "conv":
            return
        co, ci, k, _ = w.shape
        dev, bf = w.device, compute_dtype == torch.bfloat16
        # `pad_out_to` (set by Attention on its theta / phi convs): the layer computes extra all-zero output channels so
        # that Cout is a multiple of 8 -- ch/8 = 12 at ch = 96 would otherwise push the conv AND the attention products
        # onto the CUDA-core kernels (38 % of the ic-128 step, profiles/r02_ic128_kernel_table.txt).  Zero channels change
        # nothing in theta.phi^T; the master weight and its gradient keep their [Cout, Cin] shape.
        self.co_pad = ((-co) % int(getattr(self.module, "pad_out_to", 0))) if (bf and getattr(self.module, "pad_out_to", 0)) else 0
        co = co + self.co_pad
        # forward operand: how the layer runs (see _conv_forward)
        #   "tc"    tensor-core implicit GEMM on [Cout,k,k,Cin] bf16
        #   "col"   Cin<=4 (RGB in): explicit im2col to KP=16/32 columns, then the tensor-core kernel as a 1x1 conv
        #   "pad8"  Cout<=4 (RGB out): weight rows zero-padded to 8, output sliced back
        #   "f32"   CUDA-core fp32 kernels (parity mode, or shapes the tensor-core path does not take)
        def pick(cin, cout):
            if bf and cin % 16 == 0 and cout % 8 == 0:
                return "tc"
            if bf and cin <= 4 and cout % 8 == 0:
                return "col"
            if bf and cout <= 4 and cin % 16 == 0:
                return "pad8"
            return "f32"
        self.mode, self.mode_d = pick(ci, co), pick(co, ci)
        self.kp = (k * k * ci + 15) // 16 * 16   # im2col width of the forward / of the dgrad ("col" modes)
        self.kp_d = (k * k * co + 15) // 16 * 16

    def bind_aux(self, aux: Tensor):
        """aux = v[cols] | u'[rows] | sigma, 1/sigma | scratch[2]: the LIVE spectral-norm state the power iteration writes."""
        rows, cols = self.rows, self.cols
        self.aux = aux
        self.v = aux[:cols]
        self.u_new = aux[cols:cols + rows]
        self.sigma = aux[cols + rows:cols + rows + 2]
        self.scratch = aux[cols + rows + 2:cols + rows + 4]
        self.alpha = self.sigma[1:] if self.use_sn else None  # device scalar 1/sigma
        self.bind_snapshot(aux)

    def bind_snapshot(self, snap: Tensor):
        """Per-forward copy of (v, u', sigma, 1/sigma): what THIS forward's graph is differentiated with.  The live buffers
        are overwritten by the next forward (G_D with split_D=True runs D twice before one backward; the reference keeps
        sigma in each graph, layers.py:98-112)."""
        rows, cols = self.rows, self.cols
        self.snap = (snap[:cols], snap[cols:cols + rows], snap[cols + rows:cols + rows + 2])

    def descriptor(self) -> L.IcganSnLayer:
        m = self.module
        return L.IcganSnLayer(ptr(m.weight), ptr(m.u0), ptr(self.v), ptr(self.u_new), ptr(self.sigma), ptr(self.scratch),
                              self.rows, self.cols)

    def _operand(self, wk: Tensor, mode: str, kp: int):
        """[Cout,k,k,Cin] float32 -> operand tensor for `mode`."""
        co, k, _, ci = wk.shape
        if mode == "tc":
            return wk.to(torch.bfloat16)
        if mode == "col":
            out = torch.zeros(co, 1, 1, kp, device=wk.device, dtype=torch.bfloat16)
            out.view(co, kp)[:, :k * k * ci] = wk.reshape(co, -1)
            return out
        if mode == "pad8":
            out = torch.zeros(8, k, k, ci, device=wk.device, dtype=torch.bfloat16)
            out[:co] = wk
            return out
        return wk

    def _keep(self, name: str, value: Tensor) -> Tensor:
        """Operand copies are PERSISTENT buffers rewritten in place: a captured CUDA graph (biggan/graphs.py) bakes their
        addresses, and the eager path stops allocating per rebuild."""
        cur = getattr(self, name, None)
        if cur is not None and cur.shape == value.shape and cur.dtype == value.dtype and cur.device == value.device:
            cur.copy_(value)
            return cur
        value = value.contiguous()
        setattr(self, name, value)
        return value

    def _slot(self, name: str, shape, dtype, device) -> Tensor:
        cur = getattr(self, name, None)
        if cur is None or tuple(cur.shape) != tuple(shape) or cur.dtype != dtype or cur.device != device:
            cur = torch.empty(*shape, device=device, dtype=dtype)
            setattr(self, name, cur)
        return cur

    def prepare(self):
        """(Re)build the operand copies iff the master weight changed since the last build."""
        w = self.module.weight
        if self.kind != "conv":
            return
        stamp = weight_stamp(w)
        if self.version == stamp and self.module.training:
            return
        self.version = stamp
        co, ci, k, _ = w.shape
        if getattr(self.module, "sub_pixel_up", False) and k == 3:
            self.build_up_operands()
        if getattr(self.module, "pooled_down", False) and k == 3:
            self.build_down_operands()
        if self.co_pad:  # zero-padded output channels: relayout in float32, pad, narrow (small 1x1 weights only)
            f32 = torch.empty(co, k, k, ci, device=w.device, dtype=torch.float32)
            d32 = torch.empty(ci, k, k, co, device=w.device, dtype=torch.float32)
            call("icgan_sn_prepare_weight", ptr(w), None, ptr(f32), ptr(d32), co, ci, k, L.F32, stream_ptr())
            self._keep("wk_fwd", torch.nn.functional.pad(f32, (0, 0, 0, 0, 0, 0, 0, self.co_pad)).to(torch.bfloat16))
            self._keep("wk_dgrad", torch.nn.functional.pad(d32, (0, self.co_pad)).to(torch.bfloat16))
            return
        if self.mode == "tc" and self.mode_d == "tc":  # both operands are plain bf16 relayouts: write them directly
            fwd = self._slot("wk_fwd", (co, k, k, ci), torch.bfloat16, w.device)
            dgr = self._slot("wk_dgrad", (ci, k, k, co), torch.bfloat16, w.device)
            call("icgan_sn_prepare_weight", ptr(w), None, ptr(fwd), ptr(dgr), co, ci, k, L.BF16, stream_ptr())
            return
        f32 = torch.empty(co, k, k, ci, device=w.device, dtype=torch.float32)
        d32 = torch.empty(ci, k, k, co, device=w.device, dtype=torch.float32)
        call("icgan_sn_prepare_weight", ptr(w), None, ptr(f32), ptr(d32), co, ci, k, L.F32, stream_ptr())
        self._keep("wk_fwd", self._operand(f32, self.mode, self.kp))
        self._keep("wk_dgrad", self._operand(d32, self.mode_d, self.kp_d))

    def build_down_operands(self):
        """4x4 stride-2 kernel of avgpool2(conv3x3(.)) (DownConvFn), merged from the float32 master weight."""
        w = self.module.weight
        co, ci, k, _ = w.shape
        w9 = w.detach().permute(0, 2, 3, 1).reshape(co, 9, ci)
        dn = torch.einsum("tk,okc->otc", down_merge_matrix(w.device), w9)
        self._keep("wk_down", dn.to(torch.bfloat16))                      # [Co, 16, Ci]
        self._keep("wk_down_d", dn.permute(2, 1, 0).to(torch.bfloat16))   # [Ci, 16, Co]
        self.down_version = weight_stamp(w)

    def build_up_operands(self):
        """Merged 2x2-phase slices of the sub-pixel up-convolution (UpConvFn), from the float32 master weight."""
        w = self.module.weight
        co, ci, k, _ = w.shape
        w9 = w.detach().permute(0, 2, 3, 1).reshape(co, 9, ci)
        up = torch.einsum("tk,okc->otc", up_merge_matrix(w.device), w9)
        self._keep("wk_up", up.to(torch.bfloat16))                      # [Co, 16, Ci]
        self._keep("wk_up_d", up.permute(2, 1, 0).to(torch.bfloat16))   # [Ci, 16, Co]
        self.up_version = weight_stamp(w)

    def weight_grad(self, G: Tensor, snap=None) -> Tensor:
        """dL/dW (master layout) from G = dL/d(W/sigma) given in operand layout (float32); `snap` = the (v, u', sigma)
        snapshot of the forward being differentiated."""
        v, u_new, sigma = snap if snap is not None else (self.v, self.u_new, self.sigma)
        w = self.module.weight
        dW = torch.empty_like(w)
        if self.kind == "conv":
            co, ci, k, _ = w.shape
        else:
            co, ci, k = self.rows, self.cols, 1
        if self.use_sn:
            call("icgan_sn_weight_grad", ptr(G), ptr(w), ptr(u_new), ptr(v), ptr(sigma),
                 ptr(self.scratch), ptr(dW), co, ci, k, stream_ptr())
        else:
            call("icgan_sn_weight_grad", ptr(G), None, None, None, None, None, ptr(dW), co, ci, k, stream_ptr())
        return dW


def refresh_sn(states: List[SNState], training: bool, eps: float, compute_dtype, table_cache: dict) -> None:
    """One batched power-iteration step for all given layers + operand rebuild where weights changed
    (start of every forward)."""
    if not states:
        return
    for s in states:
        s._ensure(compute_dtype)
    sn_states = [s for s in states if s.use_sn]
    if sn_states:
        key = tuple(s.key for s in sn_states)
        if table_cache.get("key") != key:
            # all layers' (v, u', sigma, scratch) live in ONE flat buffer so that a forward can snapshot them in one copy
            sizes = [s.cols + s.rows + 4 for s in sn_states]
            flat = torch.zeros(sum(sizes), device=sn_states[0].module.weight.device, dtype=torch.float32)
            for s, view in zip(sn_states, flat.split(sizes)):
                s.bind_aux(view)
                s.sigma.fill_(1.0)
            table_cache["flat"], table_cache["sizes"] = flat, sizes
            arr = (L.IcganSnLayer * len(sn_states))(*[s.descriptor() for s in sn_states])
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            table_cache["table"] = host.to(sn_states[0].module.weight.device)
            table_cache["key"] = key
            table_cache["max_rows"] = max(s.rows for s in sn_states)
            table_cache["max_cols"] = max(s.cols for s in sn_states)
            table_cache["items"] = sum(((s.cols + 127) // 128) * ((s.rows + 63) // 64) for s in sn_states)
        call("icgan_sn_power_iteration", ptr(table_cache["table"]), len(sn_states), table_cache["max_rows"],
             table_cache["max_cols"], table_cache["items"], float(eps), 1 if training else 0, stream_ptr())
        if training:  # sv0 is a log-only buffer (layers.py:108-111)
            with torch.no_grad():
                torch._foreach_copy_([s.module.sv0 for s in sn_states], [s.sigma[:1] for s in sn_states])
        snap = table_cache["flat"].clone()
        for s, view in zip(sn_states, snap.split(table_cache["sizes"])):
            s.bind_snapshot(view)
    for s in states:
        s.prepare()
        s.fresh = True


# ===================================================================================== convolution
def _tc_conv(x, wk, alpha, bias, residual, y, B, H, W, cin, cout, k, res_shift, act, stats=None):
    rdt = dt(residual) if residual is not None else L.F32
    _SHAPE[0] = (B, H, W, cin, cout, k)
    _timed("tc_conv_kernel", 2.0 * B * H * W * cout * cin * k * k,
           lambda: call("icgan_conv2d_tc", ptr(x), ptr(wk), ptr(alpha), ptr(bias), ptr(residual), ptr(y), ptr(stats), B,
                        H, W, cin, cout, k, dt(y), rdt, res_shift, act, stream_ptr()))


def _im2col(x: Tensor, k: int, kp: int) -> Tensor:
    B, H, W, cs = x.shape
    out = torch.empty(B, H, W, kp, device=x.device, dtype=torch.bfloat16)
    call("icgan_im2col_small", ptr(x), ptr(out), B, H, W, cs, k, kp, dt(x), stream_ptr())
    return out


def _conv_forward(x: Tensor, st: SNState, bias, residual, res_shift: int, act: int, out_dtype, dgrad: bool = False,
                  keep: Optional[dict] = None, stats: Optional[Tensor] = None, snap=None):
    """act(alpha * conv(x, operand) + bias + residual); `keep` receives tensors worth saving for the backward.
    `snap`: (v, u', sigma) snapshot whose 1/sigma is used instead of the live one (backward of an earlier forward)."""
    B, H, W, cin = x.shape
    alpha = st.alpha if (snap is None or not st.use_sn) else snap[2][1:]
    wk, mode, kp = (st.wk_fwd, st.mode, st.kp) if not dgrad else (st.wk_dgrad, st.mode_d, st.kp_d)
    w = st.module.weight
    cout, k = (w.shape[0] + st.co_pad, w.shape[2]) if not dgrad else (w.shape[1], w.shape[2])
    if stats is not None and mode != "tc":
        raise RuntimeError("fused batch-norm statistics need the tensor-core conv path")
    if mode == "tc":
        y = torch.empty(B, H, W, cout, device=x.device, dtype=out_dtype)
        _tc_conv(x, wk, alpha, bias, residual, y, B, H, W, cin, cout, k, res_shift, act, stats)
        return y
    if mode == "col" and kp == 32 and cin <= 3 and cout <= 256 and residual is None and x.dtype == torch.bfloat16:
        # RGB-side input, im2col fused into the kernel; the backward rebuilds the column buffer only if it needs it
        y = torch.empty(B, H, W, cout, device=x.device, dtype=out_dtype)
        _SHAPE[0] = (B, H, W, cin, cout, k)
        _timed("tc_conv_kernel", 2.0 * B * H * W * cout * cin * k * k,
               lambda: call("icgan_conv2d_rgb_tc", ptr(x), ptr(wk), ptr(alpha), ptr(bias), ptr(y), B, H, W, cin, cout, k,
                            dt(y), act, stream_ptr()))
        return y
    if mode == "col":  # RGB-side input: im2col (27 -> 32 columns) + tensor-core 1x1
        xcol = _im2col(x, k, kp)
        if keep is not None:
            keep["xcol"] = xcol
        y = torch.empty(B, H, W, cout, device=x.device, dtype=out_dtype)
        _tc_conv(xcol, wk, alpha, bias, residual, y, B, H, W, kp, cout, 1, res_shift, act)
        return y
    if mode == "pad8":  # RGB-side output: 8 padded output channels, first `cout` kept
        if residual is not None:
            raise RuntimeError("residual is not supported on <=4-channel outputs")
        b8 = None
        if bias is not None:
            b8 = torch.zeros(8, device=x.device, dtype=torch.float32)
            b8[:cout] = bias
        y8 = torch.empty(B, H, W, 8, device=x.device, dtype=out_dtype)
        _tc_conv(x, wk, alpha, b8, None, y8, B, H, W, cin, 8, k, 0, act)
        return y8[..., :cout].contiguous()
    # float32 CUDA-core kernels
    y = torch.empty(B, H, W, cout, device=x.device, dtype=out_dtype)
    if min(cin, cout) <= 4 and residual is None:
        call("icgan_conv2d_small", ptr(x), ptr(wk), ptr(alpha), ptr(bias), ptr(y), B, H, W, cin, cout, k, dt(x), dt(y),
             act, stream_ptr())
    else:
        rdt = dt(residual) if residual is not None else L.F32
        call("icgan_conv2d_simt", ptr(x), ptr(wk), ptr(alpha), ptr(bias), ptr(residual), ptr(y), B, H, W, cin, cout, k,
             1, k // 2, dt(x), dt(y), rdt, res_shift, act, stream_ptr())
    return y


def _wgrad_tc(x, dy, G, B, H, W, cin, cout, k):
    _SHAPE[0] = (B, H, W, cin, cout, k)
    _timed("tc_wgrad_kernel", 2.0 * B * H * W * cout * cin * k * k,
           lambda: call("icgan_conv2d_wgrad_tc", ptr(x), ptr(dy), ptr(G), B, H, W, cin, cout, k, stream_ptr()))


class SNConvFn(torch.autograd.Function):
    """y = act(conv(x, W/sigma) + bias + residual) with stride 1, pad k//2 (layers.SNConv2d.forward, layers.py:144-153).
    `residual` may live at half resolution (res_shift=1): the nearest-upsampled shortcut of GBlock (layers.py:545-552),
    using conv1x1(up(x)) == up(conv1x1(x))."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, st: SNState, res_shift: int, act: int, out_dtype, stats=None,
                mask_input: bool = False, act_bwd_in_consumer: bool = False):
        """mask_input: x is the output of a ReLU whose backward THIS op performs -- the returned input gradient is already
        gated with (x > 0), inside the dgrad kernel's epilogue on the tensor-core path (the producer passes gradients
        through: ReluPassFn, or a conv with act_bwd_in_consumer). act_bwd_in_consumer: this op's fused ReLU is
        differentiated by its (single) consumer, so the incoming gradient is used as is."""
        x = x.contiguous()
        keep = {}
        y = _conv_forward(x, st, bias, residual, res_shift, act, out_dtype, keep=keep, stats=stats)
        ctx.st, ctx.res_shift, ctx.act = st, res_shift, act
        ctx.snap = st.snap  # this forward's (v, u', sigma): the live buffers may be overwritten before the backward runs
        ctx.mask_input, ctx.act_bwd_in_consumer = mask_input, act_bwd_in_consumer
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        ctx.res_dtype = residual.dtype if residual is not None else None
        ctx.save_for_backward(x, y if act != L.ACT_NONE else None, keep.get("xcol"))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, xcol = ctx.saved_tensors
        st: SNState = ctx.st
        dy = dy.contiguous()
        B, H, W, cout = dy.shape
        cin = x.shape[3]
        k = st.module.weight.shape[2]
        n = dy.numel()
        if ctx.act == L.ACT_RELU and not ctx.act_bwd_in_consumer:
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
        dyc = dy if dy.dtype == x.dtype else dy.to(x.dtype)  # gradients travel in the activation dtype
        if ctx.needs_input_grad[0]:
            if ctx.mask_input and st.mode_d == "tc":
                dx = _conv_forward(dyc, st, None, x, 2, L.ACT_NONE, x.dtype, dgrad=True, snap=ctx.snap)  # ReLU gate
            else:
                dx = _conv_forward(dyc, st, None, None, 0, L.ACT_NONE, x.dtype, dgrad=True, snap=ctx.snap)
                if ctx.mask_input:
                    g = torch.empty_like(dx)
                    call("icgan_relu_bwd", ptr(dx), ptr(x), ptr(g), dx.numel(), dt(x), dt(g), stream_ptr())
                    dx = g
        if ctx.needs_input_grad[1]:
            if st.mode == "tc":
                G = torch.zeros(cout, k, k, cin, device=dy.device, dtype=torch.float32)
                _wgrad_tc(x, dyc, G, B, H, W, cin, cout, k)
                if st.co_pad:
                    G = G[:cout - st.co_pad].contiguous()
            elif st.mode == "col":
                Gc = torch.zeros(cout, st.kp, device=dy.device, dtype=torch.float32)
                _wgrad_tc(xcol if xcol is not None else _im2col(x, k, st.kp), dyc, Gc, B, H, W, st.kp, cout, 1)
                G = Gc[:, :k * k * cin].reshape(cout, k, k, cin).contiguous()
            elif st.mode == "pad8":
                dy8 = torch.zeros(B, H, W, 8, device=dy.device, dtype=x.dtype)
                dy8[..., :cout] = dyc
                G8 = torch.zeros(8, k, k, cin, device=dy.device, dtype=torch.float32)
                _wgrad_tc(x, dy8, G8, B, H, W, cin, 8, k)
                G = G8[:cout].contiguous()
            else:
                G = torch.zeros(cout, k, k, cin, device=dy.device, dtype=torch.float32)
                if min(cin, cout) <= 4:
                    call("icgan_conv2d_wgrad_small", ptr(x), ptr(dy), ptr(G), B, H, W, cin, cout, k, dt(x), dt(dy),
                         stream_ptr())
                else:
                    call("icgan_conv2d_wgrad_simt", ptr(x), ptr(dyc), ptr(G), B, H, W, cin, cout, k, 1, k // 2, dt(x),
                         stream_ptr())
            dW = st.weight_grad(G, ctx.snap)
        return dx, dW, db, dres, None, None, None, None, None, None, None


# ===================================================================================== sub-pixel up-convolution
# conv3x3(nearest_up2(x)) -- every GBlock's conv1 (BigGAN.py:256-262, layers.py:542-546) -- without the upsampled tensor:
# output pixel (2i+a, 2j+b) only ever sees the 2x2 low-resolution neighbourhood of (i, j), so each of the four output
# parity classes is a 2x2-tap convolution of x with kernel rows / columns of the 3x3 filter merged:
#   a = 0: source rows (i-1, i) with (w[0], w[1]+w[2]);   a = 1: source rows (i, i+1) with (w[0]+w[1], w[2])
# 16 MACs per low-resolution pixel instead of 36, no [B,2H,2W,C] operand written by the batch norm or read by the conv.
# All three products run on the tap-table tensor-core kernels (icgan_conv2d_tc_ex / icgan_conv2d_wgrad_tc_ex).
# default ON: measured +8.3 % on the cc-256 step (619.6 -> 670.8 img/s, same box), all parity tests green with it
SUBPIXEL_UP = bool(int(__import__("os").environ.get("ICGAN_SUBPIXEL_UP", "1")))
_UP_SRC = {0: ((-1, (0,)), (0, (1, 2))), 1: ((0, (0, 1)), (1, (2,)))}  # parity -> ((source offset, merged kernel rows), ...)


def _up_slices():
    """[(a, b, dh, dw, rows, cols)] for the 16 merged weight slices, index t = ((a*2+b)*2+u)*2+v."""
    out = []
    for a in (0, 1):
        for b in (0, 1):
            for dh, rows in _UP_SRC[a]:
                for dw, cols in _UP_SRC[b]:
                    out.append((a, b, dh, dw, rows, cols))
    return out


_UP_SLICES = _up_slices()


_MERGE_CACHE = {}  # (kind, device) -> constant matrix, built once (and never inside a CUDA-graph capture)


def up_merge_matrix(device):
    """M[t, kh*3+kw] = 1 where 3x3 tap (kh, kw) is merged into slice t (float32 [16, 9])."""
    if ("up", device) in _MERGE_CACHE:
        return _MERGE_CACHE[("up", device)]
    M = torch.zeros(16, 9)
    for t, (_, _, _, _, rows, cols) in enumerate(_UP_SLICES):
        for kh in rows:
            for kw in cols:
                M[t, kh * 3 + kw] = 1.0
    _MERGE_CACHE[("up", device)] = M.to(device)
    return _MERGE_CACHE[("up", device)]


class UpConvFn(torch.autograd.Function):
    """y[B,2H,2W,Co] = conv3x3(nearest_up2(x), W/sigma) + bias, x [B,H,W,Ci] bf16."""

    @staticmethod
    def forward(ctx, x, weight, bias, st: "SNState"):
        x = x.contiguous()
        B, H, W, ci = x.shape
        co = weight.shape[0]
        y = torch.empty(B, 2 * H, 2 * W, co, device=x.device, dtype=x.dtype)
        for a in (0, 1):
            for b in (0, 1):
                taps = [(dh, dw, t) for t, (pa, pb, dh, dw, _, _) in enumerate(_UP_SLICES) if (pa, pb) == (a, b)]
                _SHAPE[0] = ("up", B, H, W, ci, co, a, b)
                _timed("tc_conv_kernel", 2.0 * B * H * W * co * ci * 4, lambda: call(
                    "icgan_conv2d_tc_ex", ptr(x), ptr(st.wk_up), ptr(st.alpha), ptr(bias), None, ptr(y), B, H, W, ci, co, 16, 4,
                    L.int_array([t[0] for t in taps]), L.int_array([t[1] for t in taps]), L.int_array([t[2] for t in taps]),
                    1, H, W, 2 * H, 2 * W, 2, a, 2, b, dt(y), L.F32, 0, stream_ptr()))
        ctx.st, ctx.snap, ctx.has_bias = st, st.snap, bias is not None
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        st: SNState = ctx.st
        dy = dy.contiguous()
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        B, H, W, ci = x.shape
        co = dy.shape[3]
        dx = dW = db = None
        # the adjoint reads dy at (2i + a - 2dh, 2j + b - 2dw): one stride-2 launch with all 16 merged slices
        offs = [(a - 2 * dh, b - 2 * dw) for (a, b, dh, dw, _, _) in _UP_SLICES]
        alpha = ctx.snap[2][1:] if st.use_sn else None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, H, W, ci, device=x.device, dtype=x.dtype)
            _SHAPE[0] = ("up_dgrad", B, H, W, co, ci)
            _timed("tc_conv_kernel", 2.0 * B * H * W * co * ci * 16, lambda: call(
                "icgan_conv2d_tc_ex", ptr(dy), ptr(st.wk_up_d), ptr(alpha), None, None, ptr(dx), B, 2 * H, 2 * W, co, ci, 16, 16,
                L.int_array([o[0] for o in offs]), L.int_array([o[1] for o in offs]), L.int_array(list(range(16))),
                2, H, W, H, W, 1, 0, 1, 0, dt(dx), L.F32, 0, stream_ptr()))
        if ctx.needs_input_grad[1]:
            G16 = torch.zeros(4, ci, 4, co, device=x.device, dtype=torch.float32)  # [phase][ci][slice in phase][co]
            for ph in range(4):
                o4 = offs[4 * ph:4 * ph + 4]
                _SHAPE[0] = ("up_wgrad", B, H, W, ci, co, ph)
                _timed("tc_wgrad_kernel", 2.0 * B * H * W * co * ci * 4, lambda: call(
                    "icgan_conv2d_wgrad_tc_ex", ptr(x), ptr(dy), ptr(G16[ph]), B, H, W, ci, 2 * H, 2 * W, co, 4,
                    L.int_array([o[0] for o in o4]), L.int_array([o[1] for o in o4]), 2, stream_ptr()))
            # un-merge: dW[co, kh, kw, ci] = sum of the slices that contain tap (kh, kw)
            M = up_merge_matrix(x.device)
            G9 = torch.einsum("tk,itc->ikc", M, G16.permute(1, 0, 2, 3).reshape(ci, 16, co))
            dW = st.weight_grad(G9.permute(2, 1, 0).reshape(co, 3, 3, ci).contiguous(), ctx.snap)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(co, device=dy.device, dtype=torch.float32)
            call("icgan_channel_sum", ptr(dy), ptr(db), B * 4 * H * W, co, dt(dy), stream_ptr())
        return dx, dW, db, None


# ===================================================================================== pooled strided down-convolution
# avgpool2(conv3x3(h)) + shortcut -- the tail of every down-sampling DBlock (BigGAN.py:587-613, layers.py:603-613) -- as ONE
# stride-2 convolution with a 4x4 kernel whose taps are the average of the four shifted 3x3 kernels:
#   W4[r][s] = 1/4 * sum_{a,b in {0,1}} w[r-a][s-b]          (16 MACs per output pixel instead of 36, no full-resolution
# conv output written, no pooling pass, bias / shortcut added in the epilogue).  Its dgrad is the stride-2 transposed
# form (four parity classes x 2x2 taps, with the ReLU gate of the block's inner activation in the epilogue), its wgrad
# the stride-2 tap-table weight gradient, un-merged to the 3x3 layout.
# default ON: measured +10.4 % on the cc-256 step (668.8 -> 738.4 img/s, same box), all parity tests green with it
POOLED_DOWN = bool(int(__import__("os").environ.get("ICGAN_POOLED_DOWN", "1")))


def down_merge_matrix(device):
    """M[t = r*4+s, kh*3+kw] = 1/4 where tap (kh, kw) of the 3x3 kernel contributes to tap (r, s) of the 4x4 one."""
    if ("down", device) in _MERGE_CACHE:
        return _MERGE_CACHE[("down", device)]
    M = torch.zeros(16, 9)
    for r in range(4):
        for s_ in range(4):
            for kh in range(3):
                for kw in range(3):
                    if 0 <= r - kh <= 1 and 0 <= s_ - kw <= 1:
                        M[r * 4 + s_, kh * 3 + kw] = 0.25
    _MERGE_CACHE[("down", device)] = M.to(device)
    return _MERGE_CACHE[("down", device)]


_DOWN_ADJ = {0: ((0, 1), (-1, 3)), 1: ((1, 0), (0, 2))}  # output parity -> ((source offset in the pooled grid, 4x4 tap row), ...)


class DownConvFn(torch.autograd.Function):
    """y[B,H/2,W/2,Co] = avgpool2(conv3x3(x, W/sigma) + bias) + residual; x [B,H,W,Ci] bf16 is the output of a ReLU whose
    backward this op performs when mask_input (as SNConvFn does for the plain path)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, st: "SNState", mask_input: bool):
        x = x.contiguous()
        B, H, W, ci = x.shape
        co = weight.shape[0]
        y = torch.empty(B, H // 2, W // 2, co, device=x.device, dtype=x.dtype)
        if residual is not None:
            residual = residual.contiguous()
        taps = [(r - 1, s_ - 1, r * 4 + s_) for r in range(4) for s_ in range(4)]
        _SHAPE[0] = ("down", B, H, W, ci, co)
        _timed("tc_conv_kernel", 2.0 * B * (H // 2) * (W // 2) * co * ci * 16, lambda: call(
            "icgan_conv2d_tc_ex", ptr(x), ptr(st.wk_down), ptr(st.alpha), ptr(bias), ptr(residual), ptr(y), B, H, W, ci, co,
            16, 16, L.int_array([t[0] for t in taps]), L.int_array([t[1] for t in taps]), L.int_array([t[2] for t in taps]),
            2, H // 2, W // 2, H // 2, W // 2, 1, 0, 1, 0, dt(y), dt(residual) if residual is not None else L.F32, 0,
            stream_ptr()))
        ctx.st, ctx.snap, ctx.has_bias, ctx.has_res, ctx.mask_input = st, st.snap, bias is not None, residual is not None, mask_input
        ctx.res_dtype = residual.dtype if residual is not None else None
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        st: SNState = ctx.st
        dy = dy.contiguous()
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        B, H, W, ci = x.shape
        Hh, Wh, co = H // 2, W // 2, dy.shape[3]
        alpha = ctx.snap[2][1:] if st.use_sn else None
        dx = dW = db = dres = None
        if ctx.has_res and ctx.needs_input_grad[3]:
            dres = dy if dy.dtype == ctx.res_dtype else dy.to(ctx.res_dtype)
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, H, W, ci, device=x.device, dtype=x.dtype)
            for a in (0, 1):
                for b in (0, 1):
                    taps = [(di, dj, r * 4 + s_) for di, r in _DOWN_ADJ[a] for dj, s_ in _DOWN_ADJ[b]]
                    _SHAPE[0] = ("down_dgrad", B, Hh, Wh, co, ci, a, b)
                    _timed("tc_conv_kernel", 2.0 * B * Hh * Wh * co * ci * 4, lambda: call(
                        "icgan_conv2d_tc_ex", ptr(dy), ptr(st.wk_down_d), ptr(alpha), None, ptr(x) if ctx.mask_input else None,
                        ptr(dx), B, Hh, Wh, co, ci, 16, 4, L.int_array([t[0] for t in taps]),
                        L.int_array([t[1] for t in taps]), L.int_array([t[2] for t in taps]), 1, Hh, Wh, H, W, 2, a, 2, b,
                        dt(dx), dt(x), 1 if ctx.mask_input else 0, stream_ptr()))
        if ctx.needs_input_grad[1]:
            G16 = torch.zeros(4, co, 4, ci, device=x.device, dtype=torch.float32)  # [tap row r][co][tap col s][ci]
            for r in range(4):
                _SHAPE[0] = ("down_wgrad", B, Hh, Wh, ci, co, r)
                _timed("tc_wgrad_kernel", 2.0 * B * Hh * Wh * co * ci * 4, lambda: call(
                    "icgan_conv2d_wgrad_tc_ex", ptr(dy), ptr(x), ptr(G16[r]), B, Hh, Wh, co, H, W, ci, 4,
                    L.int_array([r - 1] * 4), L.int_array([s_ - 1 for s_ in range(4)]), 2, stream_ptr()))
            M = down_merge_matrix(x.device)
            G9 = torch.einsum("tk,otc->okc", M, G16.permute(1, 0, 2, 3).reshape(co, 16, ci))
            dW = st.weight_grad(G9.reshape(co, 3, 3, ci).contiguous(), ctx.snap)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(co, device=dy.device, dtype=torch.float32)
            call("icgan_channel_sum", ptr(dy), ptr(db), B * Hh * Wh, co, dt(dy), stream_ptr())
        return dx, dW, db, dres, None, None


# ===================================================================================== linear / embedding
def _gemm(A, B_, Cm, M, N, K, sa, sb, sc, alpha=1.0, alpha_dev=None, beta=0.0, bias=None, batch=1, bstrides=(0, 0, 0)):
    call("icgan_gemm", ptr(A), ptr(B_), ptr(Cm), M, N, K, batch, sa[0], sa[1], bstrides[0], sb[0], sb[1], bstrides[1],
         sc[0], sc[1], bstrides[2], float(alpha), ptr(alpha_dev), float(beta), ptr(bias), dt(A), dt(B_), dt(Cm),
         stream_ptr())


class SNLinearFn(torch.autograd.Function):
    """y = x (W/sigma)^T + b  (layers.SNLinear.forward, layers.py:164-165); all float32; 1/sigma is the GEMM's alpha."""

    @staticmethod
    def forward(ctx, x, weight, bias, st: SNState):
        x = x.contiguous().float()
        Bn, K = x.shape
        N = weight.shape[0]
        y = torch.empty(Bn, N, device=x.device, dtype=torch.float32)
        _gemm(x, weight, y, Bn, N, K, (K, 1), (1, K), (N, 1), alpha_dev=st.alpha, bias=bias)  # B[k][n] = W[n][k]
        ctx.st, ctx.has_bias, ctx.snap = st, bias is not None, st.snap
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        st: SNState = ctx.st
        dy = dy.contiguous().float()
        Bn, K = x.shape
        N = dy.shape[1]
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _gemm(dy, weight, dx, Bn, K, N, (N, 1), (K, 1), (K, 1),
                  alpha_dev=ctx.snap[2][1:] if st.use_sn else None)  # dx = dy @ (W/sigma)
        if ctx.needs_input_grad[1]:
            G = torch.empty(N, K, device=x.device, dtype=torch.float32)
            _gemm(dy, x, G, N, K, Bn, (1, N), (K, 1), (K, 1))  # G = dy^T @ x = dL/d(W/sigma)
            dW = st.weight_grad(G, ctx.snap)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(N, device=x.device, dtype=torch.float32)
            call("icgan_channel_sum", ptr(dy), ptr(db), Bn, N, L.F32, stream_ptr())
        return dx, dW, db, None


class SNEmbedFn(torch.autograd.Function):
    """rows of W/sigma (layers.SNEmbedding.forward, layers.py:199-200)."""

    @staticmethod
    def forward(ctx, idx, weight, st: SNState):
        ctx.st, ctx.snap = st, st.snap
        ctx.save_for_backward(idx)
        rows = weight.detach().index_select(0, idx)
        return rows * st.alpha if st.alpha is not None else rows

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        st: SNState = ctx.st
        G = torch.zeros(st.rows, st.cols, device=dy.device, dtype=torch.float32)
        G.index_add_(0, idx, dy.float())
        return None, st.weight_grad(G, ctx.snap), None


# ===================================================================================== batch norm (+ReLU, +up x2)
class BNActFn(torch.autograd.Function):
    """y = [up2]([relu](batch_norm(x) * gain + bias)) — layers.ccbn.forward (layers.py:398-437, per-sample gain/bias
    [B,C]) and layers.bn.forward (:485-503, shared [C]); F.batch_norm semantics incl. running-stat update."""

    @staticmethod
    def forward(ctx, x, gain, bias, running_mean, running_var, training: bool, eps: float, momentum: float, relu: bool,
                up: bool, out_dtype, sums=None, shift=None, hint=None):
        x = x.contiguous()
        B, H, W, Cc = x.shape
        gain = gain.contiguous().float()
        bias = bias.contiguous().float()
        gstride = Cc if gain.dim() == 2 else 0
        dev = x.device
        if training and sums is not None:  # statistics already accumulated by the producing conv's epilogue
            mean = torch.empty(Cc, device=dev, dtype=torch.float32)
            invstd = torch.empty(Cc, device=dev, dtype=torch.float32)
            call("icgan_bn_stats_from_sums", ptr(sums), ptr(shift), B * H * W, Cc, ptr(running_mean), ptr(running_var),
                 ptr(mean), ptr(invstd), float(eps), float(momentum), stream_ptr())
        elif training:
            ws = torch.empty(2 * Cc, device=dev, dtype=torch.float32)
            mean = torch.empty(Cc, device=dev, dtype=torch.float32)
            invstd = torch.empty(Cc, device=dev, dtype=torch.float32)
            prev = hint.get("mean") if hint is not None else None  # last step's batch mean centres the one-pass moments
            call("icgan_bn_train_stats", ptr(x), B * H * W, Cc, dt(x), ptr(ws), ptr(prev), ptr(running_mean),
                 ptr(running_var), ptr(mean), ptr(invstd), float(eps), float(momentum), stream_ptr())
            if hint is not None:  # a persistent buffer: a captured CUDA graph keeps reading (and refreshing) this address
                if prev is None or prev.shape != mean.shape:
                    hint["mean"] = mean.clone()
                else:
                    prev.copy_(mean)
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
        return dx, dgain, dbias, None, None, None, None, None, None, None, None, None, None, None


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


class ReluPassFn(torch.autograd.Function):
    """relu(x) whose backward is the identity: its consumer (SNConvFn with mask_input=True) applies the gate."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        y = torch.empty_like(x)
        call("icgan_relu", ptr(x), ptr(y), x.numel(), dt(x), stream_ptr())
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy


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
        dh = dh.contiguous().float()
        call("icgan_relu_sumpool_bwd", ptr(x), ptr(dh), ptr(dx), B, H * W, Cc, dt(x), stream_ptr())
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
    _SHAPE[0] = ("gemm", batch, M, N, K)
    _timed("tc_gemm_kernel", 2.0 * batch * M * N * K,
           lambda: call("icgan_gemm_tc", ptr(A), ptr(B_), ptr(Cm), M, N, K, batch, int(a_mn), int(b_mn), lda, sab, ldb,
                        sbb, ldc, scb, float(alpha), dt(Cm), stream_ptr()))


# 1: the non-local block's logits stay in tensor memory (csrc/tc_attn.cu); 0: three batched GEMMs around fp32 logits in HBM.
FUSED_ATTENTION = __import__("os").environ.get("ICGAN_FUSED_ATTENTION", "1") != "0"


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
        ctx.fused = (tc and FUSED_ATTENTION and Q % 128 == 0 and Kk % 128 == 0 and d <= 64 and dv % 16 == 0
                     and dv <= 192)
        if ctx.fused:  # logits live in tensor memory only (csrc/tc_attn.cu)
            o = torch.empty(B, Q, dv, device=theta.device, dtype=theta.dtype)
            grad = any(ctx.needs_input_grad)
            lse = torch.empty(B, Q, device=theta.device, dtype=torch.float32) if grad else None
            call("icgan_attn_fwd", ptr(theta), ptr(phi), ptr(g), ptr(o), ptr(lse), B, Q, Kk, d, dv, stream_ptr())
            ctx.tc = True
            if grad:
                ctx.save_for_backward(theta, phi, g, o, lse)
            return o
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
        do = do.contiguous()
        if ctx.fused:
            theta, phi, g, o, lse = ctx.saved_tensors
            B, Q, d = theta.shape
            Kk, dv = phi.shape[1], g.shape[2]
            dtheta, dphi, dg = torch.empty_like(theta), torch.empty_like(phi), torch.empty_like(g)
            dsum = torch.empty_like(lse)
            call("icgan_attn_bwd_q", ptr(theta), ptr(phi), ptr(g), ptr(o), ptr(do), ptr(lse), ptr(dtheta), None,
                 ptr(dsum), B, Q, Kk, d, dv, stream_ptr())
            call("icgan_attn_bwd_kv", ptr(theta), ptr(phi), ptr(g), ptr(do), ptr(lse), ptr(dsum), ptr(dphi), ptr(dg),
                 B, Q, Kk, d, dv, stream_ptr())
            return dtheta, dphi, dg
        theta, phi, g, P = ctx.saved_tensors
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
