"""TEST INFRASTRUCTURE. Golden vectors for oracle.stylegan_nets_oracle.accumulate_gradients from the LIVE reference's
training/loss.py StyleGAN2Loss (CPU, this container only): the four phases of a training iteration (Gmain, Greg = path
length, Dmain, Dreg = R1) incl. style mixing and random noise, with fixed torch seeds.

    python oracle/make_golden_stylegan_loss.py   ->  tests/golden/stylegan_loss.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import stylegan_nets_oracle as O  # noqa: E402
from oracle.make_golden_stylegan_nets import CFG, GRAD_KEYS_D, GRAD_KEYS_G, inputs  # noqa: E402

PHASES = [("Gmain", 31, 1.0), ("Greg", 32, 4.0), ("Dmain", 33, 1.0), ("Dreg", 34, 16.0)]  # (phase, seed, gain = interval)


def main():
    sys.path.insert(0, "/root/reference/stylegan2_ada_pytorch")
    sys.path.insert(0, "/root/reference")
    from training import networks as N
    from training.loss import StyleGAN2Loss

    G = N.Generator(z_dim=CFG["z_dim"], c_dim=0, h_dim=CFG["h_dim"], w_dim=CFG["w_dim"], img_resolution=CFG["img_resolution"],
                    img_channels=3, mapping_kwargs=dict(num_layers=CFG["map_layers"]),
                    synthesis_kwargs=dict(channel_base=CFG["channel_base"], channel_max=CFG["channel_max"], num_fp16_res=0,
                                          conv_clamp=CFG["conv_clamp"]))
    D = N.Discriminator(c_dim=0, h_dim=CFG["h_dim"], img_resolution=CFG["img_resolution"], img_channels=3,
                        channel_base=CFG["channel_base"], channel_max=CFG["channel_max"], num_fp16_res=0,
                        conv_clamp=CFG["conv_clamp"], mapping_kwargs=dict(num_layers=CFG["d_map_layers"]),
                        epilogue_kwargs=dict(mbstd_group_size=CFG["mbstd_group_size"]))
    g_shapes = {k: list(v.shape) for k, v in G.state_dict().items()}
    d_shapes = {k: list(v.shape) for k, v in D.state_dict().items()}
    G.load_state_dict(O.synth_state_dict(g_shapes, 21))
    D.load_state_dict(O.synth_state_dict(d_shapes, 22))
    G.train(); D.train()
    z, h, x = inputs()
    c0 = torch.zeros(z.shape[0], 0)  # c_dim = 0: the training loop passes empty label tensors (training_loop.py:372-376)
    loss = StyleGAN2Loss(torch.device("cpu"), G.mapping, G.synthesis, D, augment_pipe=None, style_mixing_prob=0.9,
                         r1_gamma=10.0, pl_batch_shrink=2, pl_decay=0.01, pl_weight=2.0)
    loss.pl_mean.fill_(0.05)
    out = {"pl_mean_before": np.array([0.05], dtype=np.float32)}
    for phase, seed, gain in PHASES:
        G.zero_grad(set_to_none=True); D.zero_grad(set_to_none=True)
        G.requires_grad_(phase.startswith("G")); D.requires_grad_(phase.startswith("D"))
        w_before = G.mapping.w_avg.clone()
        torch.manual_seed(seed)
        loss.accumulate_gradients(phase=phase, real_img=x, real_c=c0, real_h=h, gen_z=z, gen_c=c0, gen_h=h, sync=True, gain=gain)
        net, keys = (G, GRAD_KEYS_G) if phase.startswith("G") else (D, GRAD_KEYS_D)
        params = dict(net.named_parameters())
        for k in keys:
            g = params[k].grad
            out[f"{phase}/grad/{k}"] = (g if g is not None else torch.zeros_like(params[k])).numpy().copy()
        out[f"{phase}/w_avg_before"] = w_before.numpy().copy()
        out[f"{phase}/w_avg_after"] = G.mapping.w_avg.numpy().copy()
        out[f"{phase}/pl_mean_after"] = np.array([loss.pl_mean.item()], dtype=np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "stylegan_loss.npz"), **out)
    print("wrote tests/golden/stylegan_loss.npz:", len(out), "arrays; pl_mean", out["Greg/pl_mean_after"])
    for phase, _, _ in PHASES:
        ks = [k for k in out if k.startswith(phase + "/grad/")]
        print(phase, {k.split("/grad/")[1]: float(np.abs(out[k]).max()) for k in ks[:4]})


if __name__ == "__main__":
    main()
