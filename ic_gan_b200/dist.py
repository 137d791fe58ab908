"""Data-parallel plumbing for the G+D step: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The batch shards across ranks with replicated weights; the only data-path collectives are ONE mean
all-reduce of D's gradients and ONE of G's per optimiser step, each over a single flat float32 buffer that the
parameters' ``.grad`` tensors alias (no bucketing, no copies), plus a rank-0 broadcast of the small buffers
(BN running statistics, SN ``u``) that the reference's DDP performs at every forward (trainer.py:196-210).
"""
from __future__ import annotations

from typing import Dict, Iterable, List

import torch
import torch.distributed as dist


class FlatGrads:
    """Gives every parameter a ``.grad`` that is a view into one flat float32 buffer."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()


class _Existing:
    def __init__(self, flat: torch.Tensor):
        self.flat = flat

    def zero(self):
        self.flat.zero_()


class GradSync:
    """grad_sync object for GAN_training_function: sync("D") / sync("G") -> one all-reduce (mean) each."""

    def __init__(self, nets: Dict[str, torch.nn.Module], world_size: int, optimizers: Dict[str, object] = None):
        """optimizers[k] may be an ic_gan_b200.optim.FusedAdamEMA: its flat gradient buffer is the one all-reduced and the
        1/world of the mean is folded into its step kernel (no separate scaling pass over the buffer)."""
        self.world = world_size
        self.opt = {k: (optimizers or {}).get(k) for k in nets}
        self.flat = {k: (_Existing(self.opt[k].flat_g) if hasattr(self.opt[k], "flat_g") else FlatGrads(net.parameters()))
                     for k, net in nets.items()}
        self.nets = nets
        self.calls = 0

    def sync(self, which: str):
        if self.world > 1:
            buf = self.flat[which].flat
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            if hasattr(self.opt[which], "set_grad_scale"):
                self.opt[which].set_grad_scale(1.0 / self.world)
            else:
                buf.mul_(1.0 / self.world)
            self.calls += 1

    def broadcast_buffers(self):
        """rank 0's buffers -> all ranks (DDP broadcast_buffers=True semantics), coalesced into one message per net."""
        if self.world <= 1:
            return
        for net in self.nets.values():
            bufs = [b for b in net.buffers() if b.dtype.is_floating_point]
            if not bufs:
                continue
            flat = torch.cat([b.reshape(-1) for b in bufs])
            dist.broadcast(flat, src=0)
            torch._foreach_copy_(bufs, [t.view_as(b) for t, b in zip(flat.split([b.numel() for b in bufs]), bufs)])

    def broadcast_params(self):
        if self.world <= 1:
            return
        for net in self.nets.values():
            for p in net.parameters():
                dist.broadcast(p.data, src=0)
        self.broadcast_buffers()
