"""Adam + generator EMA as ONE kernel launch per network.

The reference builds ``torch.optim.Adam`` per network (BigGAN_PyTorch/trainer.py:158-171, BigGAN.py:323-331), steps it in
``train_fns.py:115,177`` and then runs ``utils.ema.update`` (utils.py:1055-1067) -- on a GPU that is a multi-tensor Adam
launch train plus two ``_foreach`` passes over the EMA copy.  Here the parameters of a network, their gradients, both
Adam moments and the EMA copy each live in one flat float32 buffer (the module's tensors become views), and
``icgan_adam_ema_step`` performs the whole update -- optional 1/world gradient scale, Adam, EMA -- in a single HBM pass.
Same hyper-parameter meaning and state names (``exp_avg``, ``exp_avg_sq``, ``step``) as ``torch.optim.Adam``.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from ._lib import call, ptr, stream_ptr


def _flatten_into(tensors: List[torch.Tensor]) -> torch.Tensor:
    """Move `tensors` (float32, same device) into one flat buffer; each tensor's .data becomes a view of it.
    Offsets are rounded up to 4 elements so every view stays 16-byte aligned."""
    dev = tensors[0].device
    offs, n = [], 0
    for t in tensors:
        if t.dtype != torch.float32 or t.device != dev:
            raise TypeError("flat buffers hold float32 tensors of one device")
        offs.append(n)
        n += (t.numel() + 3) // 4 * 4
    flat = torch.zeros(n, device=dev, dtype=torch.float32)
    for t, o in zip(tensors, offs):
        view = flat[o:o + t.numel()].view(t.shape)
        view.copy_(t.data)
        t.data = view
    return flat


class FusedAdamEMA:
    """Drop-in for the ``torch.optim.Adam`` instances of the IC-GAN trainers (``step``, ``zero_grad``, ``param_groups``,
    ``state_dict``), restricted to what every IC-GAN config uses: one group, no weight decay, no amsgrad.

    ``ema_params``: the parameters of the EMA copy of the same network, in the same order; when given, ``step()`` also
    applies ``ema = ema*decay + p*(1-decay)`` with the decay set through ``set_ema_decay`` (None = leave the copy alone).
    ``grad_scale`` (``set_grad_scale``) multiplies the gradients before use: the 1/world of a data-parallel mean."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr=2e-4, betas=(0.0, 0.999), eps=1e-8,
                 ema_params: Optional[Iterable[torch.nn.Parameter]] = None):
        self.params = [p for p in params]
        if not self.params or not all(p.is_cuda for p in self.params):
            raise RuntimeError("FusedAdamEMA needs CUDA parameters (ic_gan_b200 has no CPU optimiser path)")
        self.param_groups = [dict(params=self.params, lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False)]
        self.flat_p = _flatten_into(self.params)
        # gradients: one flat buffer in the same layout (dist.FlatGrads uses the same class so its all-reduce sees it)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        off = 0
        self._views = []
        for p in self.params:
            p.grad = self.flat_g[off:off + p.numel()].view_as(p)
            self._views.append((off, p.numel(), p.shape))
            off += (p.numel() + 3) // 4 * 4
        self.flat_ema = None
        if ema_params is not None:
            ema_params = [p for p in ema_params]
            if [tuple(p.shape) for p in ema_params] != [tuple(p.shape) for p in self.params]:
                raise ValueError("EMA copy must have the same parameters in the same order")
            self.flat_ema = _flatten_into(ema_params)
        self.step_count = 0
        self.ema_decay: Optional[float] = None
        self.grad_scale = 1.0

    # ---- torch.optim.Optimizer surface used by the trainers ---------------------------------------------------------
    def zero_grad(self, set_to_none: bool = False):
        self.flat_g.zero_()
        for p, (off, n, shape) in zip(self.params, self._views):  # a .grad that autograd replaced is re-pointed
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * off:
                p.grad = self.flat_g[off:off + n].view(shape)

    def set_ema_decay(self, decay: Optional[float]):
        self.ema_decay = decay

    def set_grad_scale(self, scale: float):
        self.grad_scale = float(scale)

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("closures are not used by the IC-GAN trainers")
        g = self.param_groups[0]
        self.step_count += 1
        use_ema = self.flat_ema is not None and self.ema_decay is not None
        call("icgan_adam_ema_step", ptr(self.flat_p), ptr(self.flat_g), ptr(self.exp_avg), ptr(self.exp_avg_sq),
             ptr(self.flat_ema) if use_ema else None, self.flat_p.numel(), float(g["lr"]), float(g["betas"][0]),
             float(g["betas"][1]), float(g["eps"]), self.step_count, float(self.grad_scale),
             float(self.ema_decay) if use_ema else -1.0, stream_ptr())
        self.grad_scale = 1.0
        from . import ops
        ops.invalidate_operands(self)  # the conv operand copies of THESE master weights are stale now

    @property
    def state(self):
        """Per-parameter views of the moments under torch.optim.Adam's names."""
        out = {}
        for p, (off, n, shape) in zip(self.params, self._views):
            out[p] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.exp_avg[off:off + n].view(shape),
                      "exp_avg_sq": self.exp_avg_sq[off:off + n].view(shape)}
        return out

    def state_dict(self):
        st = self.state
        return {"state": {i: {k: v.clone() for k, v in st[p].items()} for i, p in enumerate(self.params)},
                "param_groups": [{**{k: v for k, v in self.param_groups[0].items() if k != "params"},
                                  "params": list(range(len(self.params)))}]}

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        self.param_groups[0].update({k: v for k, v in g.items() if k != "params"})
        for i, (off, n, shape) in enumerate(self._views):
            s = sd["state"].get(i)
            if s is None:
                continue
            self.exp_avg[off:off + n].view(shape).copy_(s["exp_avg"])
            self.exp_avg_sq[off:off + n].view(shape).copy_(s["exp_avg_sq"])
            self.step_count = int(s["step"])


class FlatBufferEMA:
    """The non-parameter half of utils.ema (BN running statistics, SN u0/sv0: every floating-point state entry that is not
    a parameter): source and target buffers flattened, one ``icgan_ema_lerp`` launch per update."""

    def __init__(self, source: torch.nn.Module, target: torch.nn.Module):
        src = [b for b in source.buffers() if b.dtype == torch.float32]
        tgt = [b for b in target.buffers() if b.dtype == torch.float32]
        if [tuple(b.shape) for b in src] != [tuple(b.shape) for b in tgt]:
            raise ValueError("EMA copy must have the same buffers in the same order")
        self.n = len(src)
        if self.n:
            self.flat_src, self.flat_tgt = _flatten_into(src), _flatten_into(tgt)

    @torch.no_grad()
    def update(self, decay: float):
        if self.n:
            call("icgan_ema_lerp", ptr(self.flat_tgt), ptr(self.flat_src), self.flat_src.numel(), float(decay),
                 stream_ptr())
