"""CUDA-graph capture of the StyleGAN2 loss phases.

One training iteration of StyleGAN2 at 256x256 issues ~5 000 small launches (14 modulated layers x {affine, demodulation,
modulate, 1-4 tensor-core convolutions, FIR, activation} x forward / backward / second backward), and at 64 images per
GPU the B200 finishes them faster than Python can issue them: measured 139 ms per iteration with the SM clock at its
maximum and no power-cap events, i.e. the device idles.  The reference hides the same problem behind cuDNN autotuning
and large batches; the B200-native answer is the one the hardware offers -- capture each phase
(``StyleGAN2Loss.accumulate_gradients`` = forward + backward [+ double backward]) ONCE into a CUDA graph and replay it.

``GraphedLoss`` wraps a ``StyleGAN2Loss``: per phase it owns static input buffers, warms the phase up on a side stream,
captures it, and afterwards ``accumulate_gradients`` is a few ``copy_`` calls plus ``graph.replay()``.  Gradients land in
the parameters' ``.grad`` (which must be persistent buffers, e.g. the flat gradient buffer of ``optim.FusedAdamEMA``);
random draws (noise inputs, style mixing, path-length noise) stay random across replays (PyTorch registers its
generator with the graph).  The style-mixing cutoff is drawn on the device in this mode (``loss.device_side_mixing``).
Optimiser steps, gradient all-reduce and EMA stay outside the graphs."""
from __future__ import annotations

from typing import Dict

import torch


class GraphedLoss:
    def __init__(self, loss, modules: Dict[str, torch.nn.Module], warmup: int = 3):
        """modules: {"G": generator, "D": discriminator} -- whose requires_grad each phase toggles (training_loop.py:491-512)."""
        self.loss, self.modules, self.warmup = loss, modules, warmup
        self.graphs, self.static, self.launches = {}, {}, {}
        self.replayed_launches = 0  # kernels of this package replayed so far (counted once at capture, per phase)
        loss.device_side_mixing = True

    def _toggle(self, phase, on):
        self.modules["G" if phase.startswith("G") else "D"].requires_grad_(on)

    def _capture(self, phase, inputs, gain):
        from .. import ops
        if ops.PROFILE is not None:
            raise RuntimeError("per-launch event timing (ops.PROFILE) cannot be captured into a graph")
        st = {k: v.clone() for k, v in inputs.items()}
        self._toggle(phase, True)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up: first-use kernel attributes, allocator pool, autograd buffers
            for _ in range(self.warmup):
                self.loss.accumulate_gradients(phase=phase, sync=True, gain=gain, **st)
        torch.cuda.current_stream().wait_stream(side)
        from .. import _lib
        before = _lib.LAUNCHES
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.loss.accumulate_gradients(phase=phase, sync=True, gain=gain, **st)
        self._toggle(phase, False)
        self.graphs[(phase, gain)], self.static[(phase, gain)] = g, st
        self.launches[(phase, gain)] = _lib.LAUNCHES - before

    def accumulate_gradients(self, phase, real_img, real_c, real_h, gen_z, gen_c, gen_h, sync=True, gain=1):
        """Same call as StyleGAN2Loss.accumulate_gradients (loss.py:85).  NOTE: capturing a phase runs it warmup+1 times,
        so its gradients from the capturing call are (warmup+1)x accumulated -- zero the gradient buffer before the call
        that follows the first one of each phase (bench.py does; trainers call one warm iteration and discard it)."""
        inputs = dict(real_img=real_img, real_c=real_c, real_h=real_h, gen_z=gen_z, gen_c=gen_c, gen_h=gen_h)
        key = (phase, gain)
        if key not in self.graphs:
            self._capture(phase, inputs, gain)
            return
        for k, v in inputs.items():
            self.static[key][k].copy_(v)
        self.graphs[key].replay()
        self.replayed_launches += self.launches[key]

    @property
    def stats(self):
        return self.loss.stats
