"""StyleGAN2-ADA loss of IC-GAN on the B200 networks: the four phases of a training
iteration -- Gmain, Greg (path-length regulariser: a double backward through the synthesis network w.r.t. the latents),
Dmain, Dreg (R1: a double backward through the discriminator w.r.t. the real images).

Mirrors stylegan2_ada_pytorch/training/loss.py (`Loss` :15-20, `StyleGAN2Loss` :26-194): same constructor keywords,
`run_G`, `run_D`, `accumulate_gradients(phase, real_img, real_c, real_h, gen_z, gen_c, gen_h, sync, gain)`.  `sync` is
accepted and ignored unless the modules are DistributedDataParallel (then `no_sync()` is used exactly as
`misc.ddp_sync` does); this package's own data-parallel path all-reduces flat gradient buffers once per phase instead.
The style-mixing cutoff is drawn on the host (it only selects a slice boundary)."""
from __future__ import annotations

import contextlib
import math

import torch

from .ops import conv2d_gradfix


@contextlib.contextmanager
def _ddp_sync(module, sync):
    if sync or not isinstance(module, torch.nn.parallel.DistributedDataParallel):
        yield
    else:
        with module.no_sync():
            yield


class Loss:
    def accumulate_gradients(self, phase, real_img, real_c, real_h, gen_z, gen_c, gen_h, sync, gain):
        raise NotImplementedError()


class StyleGAN2Loss(Loss):
    def __init__(self, device, G_mapping, G_synthesis, D, augment_pipe=None, style_mixing_prob=0.9, r1_gamma=10,
                 pl_batch_shrink=2, pl_decay=0.01, pl_weight=2):
        super().__init__()
        self.device = device
        self.G_mapping, self.G_synthesis, self.D = G_mapping, G_synthesis, D
        self.augment_pipe = augment_pipe
        self.style_mixing_prob, self.r1_gamma = style_mixing_prob, r1_gamma
        self.pl_batch_shrink, self.pl_decay, self.pl_weight = pl_batch_shrink, pl_decay, pl_weight
        self.pl_mean = torch.zeros([], device=device)
        self.device_side_mixing = False  # True: draw the style-mixing cutoff on the device (CUDA-graph capturable)
        self.stats = {}  # last value of each reported loss term (0-d device tensors; the reference hands these to
        #                  training_stats.report, loss.py:101-103,120-122,...): what a trainer reads back per iteration

    def run_G(self, z, c, h, sync):
        with _ddp_sync(self.G_mapping, sync):
            ws = self.G_mapping(z, c, h)
            if self.style_mixing_prob > 0 and self.device_side_mixing:
                # the same distribution (loss.py:66-70) with no host decision: cutoff ~ U{1..num_ws-1} with probability
                # style_mixing_prob, else num_ws; latents at and after the cutoff come from a second mapping pass
                n_ws = ws.shape[1]
                cutoff = torch.randint(1, n_ws, [], device=ws.device)
                cutoff = torch.where(torch.rand([], device=ws.device) < self.style_mixing_prob, cutoff,
                                     torch.full_like(cutoff, n_ws))
                ws2 = self.G_mapping(torch.randn_like(z), c, h, skip_w_avg_update=True)
                keep = (torch.arange(n_ws, device=ws.device) < cutoff).reshape(1, n_ws, 1)
                ws = torch.where(keep, ws, ws2)
            elif self.style_mixing_prob > 0:
                cutoff = int(torch.empty([], dtype=torch.int64).random_(1, ws.shape[1]))
                if not bool(torch.rand([]) < self.style_mixing_prob):
                    cutoff = ws.shape[1]
                ws[:, cutoff:] = self.G_mapping(torch.randn_like(z), c, h, skip_w_avg_update=True)[:, cutoff:]
        with _ddp_sync(self.G_synthesis, sync):
            img = self.G_synthesis(ws)
        return img, ws

    def run_D(self, img, c, h, sync):
        if self.augment_pipe is not None:
            img = self.augment_pipe(img)
        with _ddp_sync(self.D, sync):
            return self.D(img, c, h)

    def accumulate_gradients(self, phase, real_img, real_c, real_h, gen_z, gen_c, gen_h, sync, gain):
        assert phase in ["Gmain", "Greg", "Gboth", "Dmain", "Dreg", "Dboth"]
        g_main, d_main = phase in ["Gmain", "Gboth"], phase in ["Dmain", "Dboth"]
        g_pl = phase in ["Greg", "Gboth"] and self.pl_weight != 0
        d_r1 = phase in ["Dreg", "Dboth"] and self.r1_gamma != 0
        softplus = torch.nn.functional.softplus

        if g_main:  # maximise the logits of generated images: -log sigmoid(D(G(z)))
            gen_img, _ = self.run_G(gen_z, gen_c, gen_h, sync=(sync and not g_pl))
            loss = softplus(-self.run_D(gen_img, gen_c, gen_h, sync=False))
            self.stats["Loss/G/loss"] = loss.detach().mean()
            loss.mean().mul(gain).backward()

        if g_pl:  # path length: |J_w^T y| should stay near its running mean
            n = gen_z.shape[0] // self.pl_batch_shrink
            gen_img, gen_ws = self.run_G(gen_z[:n], gen_c[:n], gen_h[:n], sync=sync)
            pl_noise = torch.randn_like(gen_img) / math.sqrt(gen_img.shape[2] * gen_img.shape[3])
            with conv2d_gradfix.no_weight_gradients():
                pl_grads = torch.autograd.grad(outputs=[(gen_img * pl_noise).sum()], inputs=[gen_ws], create_graph=True,
                                               only_inputs=True)[0]
            pl_lengths = pl_grads.square().sum(2).mean(1).sqrt()
            pl_mean = self.pl_mean.lerp(pl_lengths.mean(), self.pl_decay)
            self.pl_mean.copy_(pl_mean.detach())
            loss_pl = (pl_lengths - pl_mean).square() * self.pl_weight
            (gen_img[:, 0, 0, 0] * 0 + loss_pl).mean().mul(gain).backward()

        loss_gen = 0
        if d_main:  # minimise the logits of generated images: -log(1 - sigmoid(D(G(z))))
            gen_img, _ = self.run_G(gen_z, gen_c, gen_h, sync=False)
            loss_gen = softplus(self.run_D(gen_img, gen_c, gen_h, sync=False))
            self.stats["Loss/D/loss_gen"] = loss_gen.detach().mean()
            loss_gen.mean().mul(gain).backward()

        if d_main or d_r1:  # real images: logistic loss and/or R1 gradient penalty
            real_tmp = real_img.detach().requires_grad_(d_r1)
            real_logits = self.run_D(real_tmp, real_c, real_h, sync=sync)
            loss_real = softplus(-real_logits) if d_main else 0
            loss_r1 = 0
            if d_r1:
                with conv2d_gradfix.no_weight_gradients():
                    r1_grads = torch.autograd.grad(outputs=[real_logits.sum()], inputs=[real_tmp], create_graph=True,
                                                   only_inputs=True)[0]
                loss_r1 = r1_grads.square().sum([1, 2, 3]) * (self.r1_gamma / 2)
            (real_logits * 0 + loss_real + loss_r1).mean().mul(gain).backward()
