"""CUDA-graph replay of the micro-steps of ``GAN_training_function``.

One G+D step of BigGAN 256x256 issues ~4 400 launches; most of the device time is in a few hundred tensor-core
convolutions, but the spectral-norm bookkeeping, batch-norm statistics, bias gradients and the optimiser glue are
thousands of launches of a few microseconds each, and between them the B200 waits for Python: the launch list of a step
shows ~9 % of the wall time with no kernel running (profiles/r02_cc256_kernel_table.txt: 298 ms of kernels in a 328 ms
step).  Each micro-step (one discriminator accumulation = G forward without grad + D forward/backward on fake+real; one
generator accumulation = G and D forward/backward) has static shapes, draws no random numbers and reads nothing back,
so it is captured ONCE into a CUDA graph and replayed; conditioning draws, optimiser steps, gradient all-reduce and the
EMA stay outside (train_fns.py).

Exactness: capture needs warm-up executions, which would otherwise leave their marks (gradients accumulated, batch-norm
running statistics and the power-iteration vector advanced).  ``run`` therefore stashes every gradient and buffer of both
networks before warming up, restores them after capture and only then replays, so the capturing call has exactly the
effect of one eager micro-step.  The operand copies of the weights (ops.SNState.prepare: bf16 re-layouts, merged up- /
down-sampling kernels) are persistent buffers rewritten in place, and they are NOT part of the graphs: ``run`` rebuilds
them eagerly before a replay iff an optimiser step happened since (a replay runs no Python that could notice)."""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from .. import _lib, ops


class GraphedMicroSteps:
    def __init__(self, nets: Dict[str, torch.nn.Module], warmup: int = 1):
        self.nets, self.warmup = nets, warmup
        self.graphs, self.static, self.outputs, self.launches = {}, {}, {}, {}
        self.replayed_launches = 0  # kernels of this package replayed so far (counted once at capture, per graph)

    def _state(self):
        out = []
        for net in self.nets.values():
            for p in net.parameters():
                if p.requires_grad:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                    out.append(p.grad)
            out.extend(net.buffers())
        return out

    def _capture(self, key, fn: Callable, inputs: Dict[str, Optional[torch.Tensor]]):
        if ops.PROFILE is not None:
            raise RuntimeError("per-launch event timing (ops.PROFILE) cannot be captured into a graph")
        st = {k: (v.clone() if v is not None else None) for k, v in inputs.items()}
        state = self._state()
        stash = [t.clone() for t in state]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # allocator pool, autograd buffers, first-use kernel attributes
            for _ in range(self.warmup):
                fn(**st)
        torch.cuda.current_stream().wait_stream(side)
        self._refresh_operands()  # nothing to rebuild right after the warm-up: no re-layout kernel lands in the graph
        before = _lib.LAUNCHES
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn(**st)
        with torch.no_grad():
            torch._foreach_copy_(state, stash)
        self.graphs[key], self.static[key], self.outputs[key] = g, st, out
        self.launches[key] = _lib.LAUNCHES - before

    def _refresh_operands(self):
        for net in self.nets.values():
            net.prepare_operands()

    def run(self, name: str, fn: Callable, **inputs):
        """fn(**inputs) -> tuple of tensors; later calls with inputs of the same shapes replay the captured graph and
        return the same (static) output tensors."""
        key = (name,) + tuple((k, tuple(v.shape), v.dtype) if v is not None else (k, None) for k, v in inputs.items())
        if key not in self.graphs:
            self._capture(key, fn, inputs)
        self._refresh_operands()
        st = self.static[key]
        for k, v in inputs.items():
            if v is not None:
                st[k].copy_(v, non_blocking=True)
        self.graphs[key].replay()
        self.replayed_launches += self.launches[key]
        return self.outputs[key]
