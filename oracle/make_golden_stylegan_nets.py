"""TEST INFRASTRUCTURE. Golden vectors for oracle/stylegan_nets_oracle.py from the LIVE reference (this container only):
builds stylegan2_ada_pytorch/training/networks.Generator / Discriminator (IC-GAN variant: h_dim > 0) at a tiny
configuration on the CPU (bias_act / upfirdn2d take their own reference implementations there), loads deterministic
synthetic weights and records images, latents, logits, the w_avg update and loss gradients.

    python oracle/make_golden_stylegan_nets.py   ->  tests/golden/stylegan_nets.{npz,json}"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import stylegan_nets_oracle as O  # noqa: E402

CFG = dict(z_dim=16, c_dim=0, h_dim=24, w_dim=32, img_resolution=32, img_channels=3, channel_base=256, channel_max=32,
           map_layers=2, d_map_layers=2, conv_clamp=256.0, mbstd_group_size=4)
GRAD_KEYS_G = ["mapping.embed_feats.weight", "mapping.fc1.bias", "synthesis.b4.const", "synthesis.b8.conv0.weight",
               "synthesis.b8.conv0.noise_strength", "synthesis.b16.conv1.affine.weight", "synthesis.b32.torgb.weight",
               "synthesis.b32.conv1.bias"]
GRAD_KEYS_D = ["b32.fromrgb.weight", "b32.conv1.weight", "b16.skip.weight", "b8.conv0.bias", "mapping.embed_feats.weight",
               "mapping.fc1.weight", "b4.conv.weight", "b4.fc.weight", "b4.out.bias"]


def inputs(B=4):
    g = torch.Generator().manual_seed(11)
    z = torch.randn(B, CFG["z_dim"], generator=g)
    h = F.normalize(torch.randn(B, CFG["h_dim"], generator=g), dim=1)
    x = torch.rand(B, 3, CFG["img_resolution"], CFG["img_resolution"], generator=g) * 2 - 1
    return z, h, x


def main():
    sys.path.insert(0, "/root/reference/stylegan2_ada_pytorch")
    sys.path.insert(0, "/root/reference")
    from training import networks as N  # the reference, unmodified

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
    g_sd, d_sd = O.synth_state_dict(g_shapes, 21), O.synth_state_dict(d_shapes, 22)
    G.load_state_dict(g_sd, strict=True)
    D.load_state_dict(d_sd, strict=True)
    z, h, x = inputs()
    out = {}
    G.eval(); D.eval()
    with torch.no_grad():
        out["ws"] = G.mapping(z, None, h).numpy()
        out["ws_trunc"] = G.mapping(z, None, h, truncation_psi=0.7, truncation_cutoff=3).numpy()
        out["img_const"] = G(z, None, h, noise_mode="const").numpy()
        out["img_none"] = G(z, None, h, noise_mode="none").numpy()
        out["img_trunc"] = G(z, None, h, truncation_psi=0.5, noise_mode="const").numpy()
        torch.manual_seed(5)
        out["img_random"] = G(z, None, h, noise_mode="random").numpy()
        out["d_real"] = D(x, None, h).numpy()
        out["d_fake"] = D(torch.from_numpy(out["img_const"]), None, h).numpy()
    # training mode: non-fused modulated conv (networks.py:589-594), w_avg tracking (:330-335)
    G.train(); D.train()
    img = G(z, None, h, noise_mode="const")
    out["img_train"] = img.detach().numpy()
    out["w_avg_after"] = G.mapping.w_avg.detach().numpy().copy()
    # loss.py:96-100 (Gmain) and :126-150 (Dmain): non-saturating logistic
    loss_g = F.softplus(-D(img, None, h)).mean()
    G.zero_grad(); D.zero_grad()
    loss_g.backward()
    out["loss_g"] = np.array([loss_g.item()])
    for k in GRAD_KEYS_G:
        out["G_grad/" + k] = dict(G.named_parameters())[k].grad.numpy().copy()
    G.zero_grad(); D.zero_grad()
    loss_d = F.softplus(D(img.detach(), None, h)).mean() + F.softplus(-D(x, None, h)).mean()
    loss_d.backward()
    out["loss_d"] = np.array([loss_d.item()])
    for k in GRAD_KEYS_D:
        out["D_grad/" + k] = dict(D.named_parameters())[k].grad.numpy().copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "stylegan_nets.npz"), **out)
    with open(os.path.join(ROOT, "tests", "golden", "stylegan_nets.json"), "w") as f:
        json.dump({"cfg": CFG, "g_shapes": g_shapes, "d_shapes": d_shapes, "g_seed": 21, "d_seed": 22,
                   "grad_keys_g": GRAD_KEYS_G, "grad_keys_d": GRAD_KEYS_D}, f, indent=1)
    print("wrote tests/golden/stylegan_nets.{npz,json}:", len(out), "arrays; |img| max", float(np.abs(out["img_const"]).max()),
          "d_real", out["d_real"].ravel())


if __name__ == "__main__":
    main()
