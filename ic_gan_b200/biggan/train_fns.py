"""The G+D training step with the reference's calling convention.

``GAN_training_function`` mirrors ``BigGAN_PyTorch/train_fns.py:28-193`` (same arguments, same schedule: zero grads,
toggle requires_grad, ``num_D_steps`` x ``num_D_accumulations`` D micro-batches with hinge loss, optimizer_D.step, then
``num_G_accumulations`` G micro-batches, optimizer_G.step, EMA) so ``trainer.py`` can call it unchanged.  The optional
``grad_sync`` implements the data-parallel reduction B200-style: ONE NCCL all-reduce of D's flat gradient buffer after
the last D micro-batch and ONE for G (the reference's DDP all-reduces on every micro-batch, train_fns.py:68-107).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def toggle_grad(model, on_or_off: bool) -> None:
    for p in model.parameters():
        p.requires_grad = on_or_off


def loss_hinge_dis(dis_fake, dis_real):  # BigGAN_PyTorch/losses.py:24-27
    return torch.mean(F.relu(1.0 - dis_real)), torch.mean(F.relu(1.0 + dis_fake))


def loss_hinge_gen(dis_fake):  # losses.py:36-38
    return -torch.mean(dis_fake)


class ema(object):
    """Exponential moving average over EVERY state_dict entry (buffers included), decay 0 before ``start_itr``
    (BigGAN_PyTorch/utils.py:1039-1067), as two fused multi-tensor launches."""

    def __init__(self, source, target, decay=0.9999, start_itr=0):
        self.source, self.target, self.decay, self.start_itr = source, target, decay, start_itr
        self.source_dict = self.source.state_dict()
        self.target_dict = self.target.state_dict()
        with torch.no_grad():
            for k in self.source_dict:
                self.target_dict[k].data.copy_(self.source_dict[k].data)
        self.fused_optimizer = None  # set by fuse_into(): the parameter half then rides in the optimiser's kernel
        self.buffer_ema = None

    def decay_at(self, itr=None):
        return 0.0 if (itr and itr < self.start_itr) else self.decay  # utils.py:1058: itr == 0 uses the real decay

    def fuse_into(self, optimizer):
        """Let `optimizer` (ic_gan_b200.optim.FusedAdamEMA built with ema_params=target.parameters()) apply the parameter
        half of the average inside its step kernel; update() then only averages the buffers (one launch)."""
        from ..optim import FlatBufferEMA
        self.fused_optimizer = optimizer
        self.buffer_ema = FlatBufferEMA(self.source, self.target)
        self.source_dict, self.target_dict = self.source.state_dict(), self.target.state_dict()

    def update(self, itr=None):
        decay = self.decay_at(itr)
        if self.fused_optimizer is not None:  # parameters were averaged by the optimiser step that just ran
            self.buffer_ema.update(decay)
            return
        with torch.no_grad():
            keys = [k for k in self.source_dict if self.target_dict[k].dtype.is_floating_point]
            tgt = [self.target_dict[k].data for k in keys]
            src = [self.source_dict[k].data for k in keys]
            torch._foreach_mul_(tgt, decay)
            torch._foreach_add_(tgt, src, alpha=1.0 - decay)


def GAN_training_function(G, D, GD, ema, state_dict, config, sample_conditionings, embedded_optimizers=True,
                          device="cuda", batch_size=0, grad_sync=None, lazy_losses=False, graphs=False):
    """`grad_sync` / `lazy_losses` / `graphs` are extensions (defaults = reference behaviour): see the module docstring;
    with lazy_losses=True train() returns 0-d device tensors instead of the reference's Python floats
    (train_fns.py:183-187), leaving the host free to queue the next step; with graphs=True every micro-step (forward +
    loss + backward of one accumulation) is captured into a CUDA graph on first use and replayed afterwards
    (biggan/graphs.py; pass a GraphedMicroSteps to share
    captured graphs between training functions over the same networks); the object is exposed as train.graphs."""
    graphed = None
    if graphs:
        from .graphs import GraphedMicroSteps
        graphed = graphs if isinstance(graphs, GraphedMicroSteps) else GraphedMicroSteps({"G": G, "D": D})
    n_d_acc, n_g_acc = float(config["num_D_accumulations"]), float(config["num_G_accumulations"])

    def _d_micro(z, gy, fg, x, dy, f):
        D_fake, D_real = GD(z, gy, fg, x, dy, f, train_G=False, split_D=config["split_D"],
                            policy=config.get("DiffAugment", False), DA=config.get("DA", False))
        D_loss_real, D_loss_fake = loss_hinge_dis(D_fake, D_real)
        ((D_loss_real + D_loss_fake) / n_d_acc).backward()
        return D_loss_real.detach(), D_loss_fake.detach()

    def _g_micro(z, gy, fg):
        D_fake = GD(z, gy, fg, train_G=True, split_D=config["split_D"], policy=config.get("DiffAugment", False),
                    DA=config.get("DA", False))
        G_loss = loss_hinge_gen(D_fake) / n_g_acc
        G_loss.backward()
        return (G_loss.detach(),)

    def _opt(net, name):
        return net.optim if embedded_optimizers else getattr(GD, name)

    def _zero(opt):
        opt.zero_grad(set_to_none=False)

    def _draw(have_y, have_f, n):
        cond = sample_conditionings()
        labels_g = f_g = None
        if have_f and have_y:
            z_, labels_g, f_g = cond
        elif have_y:
            z_, labels_g = cond
        elif have_f:
            z_, f_g = cond
        else:
            z_ = cond[0] if isinstance(cond, (tuple, list)) else cond
        if n is not None:  # the D phase slices the draw to the micro-batch (train_fns.py:82-88); the G phase does not (:144-150)
            z_ = z_[:n]
            labels_g = labels_g[:n] if labels_g is not None else None
            f_g = f_g[:n] if f_g is not None else None
        if labels_g is not None:
            labels_g = labels_g.to(device, non_blocking=True).long()
        if f_g is not None:
            f_g = f_g.to(device, non_blocking=True)
        return z_.to(device, non_blocking=True), labels_g, f_g

    def train(x, y=None, features=None):
        opt_G, opt_D = _opt(G, "optimizer_G"), _opt(D, "optimizer_D")
        _zero(opt_G)
        _zero(opt_D)
        xs = torch.split(x, batch_size)
        ys = torch.split(y, batch_size) if y is not None else None
        fs = torch.split(features, batch_size) if features is not None else None
        counter = 0
        if config["toggle_grads"]:
            toggle_grad(D, True)
            toggle_grad(G, False)
        for _ in range(config["num_D_steps"]):
            _zero(opt_D)
            for _ in range(config["num_D_accumulations"]):
                z_, labels_g, f_g = _draw(y is not None, features is not None, batch_size)
                args = dict(z=z_, gy=labels_g, fg=f_g, x=xs[counter], dy=ys[counter] if ys is not None else None,
                            f=fs[counter] if fs is not None else None)
                D_loss_real, D_loss_fake = graphed.run("D", _d_micro, **args) if graphed else _d_micro(**args)
                counter += 1
            if config.get("D_ortho", 0.0) > 0.0:
                raise NotImplementedError("ortho regularisation is off in every IC-GAN config")
            if grad_sync is not None:
                grad_sync.sync("D")
            opt_D.step()
        if config["toggle_grads"]:
            toggle_grad(D, False)
            toggle_grad(G, True)
        _zero(opt_G)
        for _ in range(config["num_G_accumulations"]):
            z_, labels_g, f_g = _draw(y is not None, features is not None, None)
            args = dict(z=z_, gy=labels_g, fg=f_g)
            (G_loss,) = graphed.run("G", _g_micro, **args) if graphed else _g_micro(**args)
        if config.get("G_ortho", 0.0) > 0.0:
            raise NotImplementedError("ortho regularisation is off in every IC-GAN config")
        if grad_sync is not None:
            grad_sync.sync("G")
        if config["ema"] and getattr(ema, "fused_optimizer", None) is opt_G:
            opt_G.set_ema_decay(ema.decay_at(state_dict["itr"]))
        opt_G.step()
        if config["ema"]:
            ema.update(state_dict["itr"])
        # (with graphs the three tensors are the graphs' static outputs: read or clone them before the next step)
        out = {"G_loss": G_loss.detach(), "D_loss_real": D_loss_real.detach(), "D_loss_fake": D_loss_fake.detach()}
        if lazy_losses:
            return out
        vals = torch.stack([out["G_loss"].float(), out["D_loss_real"].float(), out["D_loss_fake"].float()]).tolist()
        return {"G_loss": vals[0], "D_loss_real": vals[1], "D_loss_fake": vals[2]}  # one device->host read (12 bytes)

    train.graphs = graphed
    return train
