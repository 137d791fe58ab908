"""Round-2 pins from the LIVE reference (build container only; needs /root/reference):

  python oracle/make_golden_r2.py [width] [bound] [step]

* ``width``  -- one real-width slice of BASELINE config 3: cc-256, ch 96 (1536-channel layers), batch 2: G eval image, D phase
  and G phase, frozen exactly like the tiny cases of make_golden.py (oracle asserted equal to the reference first).
* ``bound``  -- what bf16 costs the REFERENCE ITSELF: the unmodified reference modules run under
  ``torch.autocast('cpu', torch.bfloat16)`` against their own fp32 run on the same weights and inputs (image max-abs, logit
  max-abs, worst per-parameter gradient rel-L2, 0-d parameters).  tests/test_biggan_gpu.py derives the tolerances of
  the bf16 tensor-core mode from these measured numbers instead of from hand-picked constants.
* ``step``   -- two full ``train_fns.GAN_training_function(...).train`` calls (BigGAN_PyTorch/train_fns.py:40-191) with two
  accumulations each, embedded Adam optimisers and ``utils.ema``: weights, Adam moments, EMA copy and BN/SN buffers
  afterwards.  Pins ``oracle.biggan_oracle.train_step`` (asserted here) and is what the -m gpu full-step test checks
  ``ic_gan_b200.biggan.train_fns`` against.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ICGAN_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import biggan_oracle as O  # noqa: E402
from oracle import make_golden as MG  # noqa: E402

CC256 = dict(resolution=256, G_ch=96, D_ch=96, G_attn="64", D_attn="64", n_classes=1000, shared_dim=128, hier=True,
             class_cond=True, instance_cond=True)
TINY = {
    "ic64_tiny": (dict(resolution=64, G_ch=16, D_ch=16, G_attn="32", D_attn="32", class_cond=False,
                       instance_cond=True), 4, 12),
    "cc32_tiny": (dict(resolution=32, G_ch=16, D_ch=16, G_attn="16", D_attn="16", n_classes=10, shared_dim=32,
                       shared_dim_feat=64, class_cond=True, instance_cond=True), 4, 21),
    "cc256_w96": (CC256, 2, 33),
}


def rel_l2(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# ------------------------------------------------------------------------------------------------- bf16 bound
def bf16_bound(RB, RL, name):
    kw, B, seed = TINY[name]
    cfg = O.BigGANConfig(**kw)
    G, D = MG._ref_models(RB, cfg)
    gs, ds = MG._check_shapes(G, D, cfg)
    g_sd0, d_sd0 = O.synth_state_dict(gs, seed), O.synth_state_dict(ds, seed + 1)
    z, feats, feats_r, x, lab, lab_r = MG._inputs(cfg, B, seed + 2)
    GD = RB.G_D(G, D)

    def fresh():
        G.load_state_dict({k: v.clone() for k, v in g_sd0.items()})
        D.load_state_dict({k: v.clone() for k, v in d_sd0.items()})
        for p in list(G.parameters()) + list(D.parameters()):
            p.grad = None

    def run(bf16):
        ctx = (lambda: torch.autocast("cpu", dtype=torch.bfloat16)) if bf16 else contextlib.nullcontext
        out = {}
        fresh()
        G.eval()
        with torch.no_grad(), ctx():
            out["img"] = G(z, lab, feats).float()
        fresh()
        G.train(); D.train()
        for p in G.parameters():
            p.requires_grad_(False)
        for p in D.parameters():
            p.requires_grad_(True)
        with ctx():
            d_fake, d_real = GD(z, lab, feats, x, lab_r, feats_r, train_G=False)
            a, b = RL.loss_hinge_dis(d_fake.float(), d_real.float())
        (a + b).backward()
        out["d_logits"] = torch.cat([d_fake, d_real]).detach().float()
        out["d_grads"] = {k: p.grad.clone() for k, p in D.named_parameters()}
        fresh()
        for p in G.parameters():
            p.requires_grad_(True)
        for p in D.parameters():
            p.requires_grad_(False)
        with ctx():
            d_fake = GD(z, lab, feats, train_G=True)
            gl = RL.loss_hinge_gen(d_fake.float())
        gl.backward()
        out["g_logits"] = d_fake.detach().float()
        out["g_grads"] = {k: p.grad.clone() for k, p in G.named_parameters() if p.grad is not None}
        return out

    f32, b16 = run(False), run(True)

    def grads(tag, params):
        worst, worst_k, zero_d = 0.0, "", 0.0
        for k, ref in f32[tag].items():
            scale = max(1.0, params[k].abs().max().item())
            if ref.abs().max().item() < 1e-5 * scale:
                continue
            e = rel_l2(b16[tag][k], ref)
            if ref.dim() == 0:
                zero_d = max(zero_d, e)
            elif e > worst:
                worst, worst_k = e, k
        return {"worst_rel_l2": worst, "worst_param": worst_k, "zero_dim_rel": zero_d}

    rep = {
        "image_max_abs": float((b16["img"] - f32["img"]).abs().max()),
        "d_logit_max_abs_over_scale": float((b16["d_logits"] - f32["d_logits"]).abs().max()
                                            / max(1.0, f32["d_logits"].abs().max().item())),
        "g_logit_max_abs_over_scale": float((b16["g_logits"] - f32["g_logits"]).abs().max()
                                            / max(1.0, f32["g_logits"].abs().max().item())),
        "d_phase_grads": grads("d_grads", d_sd0),
        "g_phase_grads": grads("g_grads", g_sd0),
        "how": "unmodified reference modules, CPU, torch.autocast('cpu', bfloat16) vs their own fp32 run",
    }
    print(f"[bound] {name}: {json.dumps(rep)}")
    return rep


# ------------------------------------------------------------------------------------------------- full step
from oracle.step_fixture import STEP, STEP_CFG, sample_of, step_inputs  # noqa: E402


def full_step(RB):
    sys.path.insert(0, os.path.join(REF, "BigGAN_PyTorch"))
    with contextlib.redirect_stdout(io.StringIO()):
        import train_fns as RT  # noqa  (the reference's own file)
        import utils as RU  # noqa
    cfg, hp = O.BigGANConfig(**STEP_CFG), STEP
    common = dict(resolution=cfg.resolution, n_classes=cfg.n_classes, SN_eps=cfg.SN_eps, class_cond=True,
                  instance_cond=True, skip_init=True, adam_eps=hp["adam_eps"])
    gkw = dict(G_ch=cfg.G_ch, dim_z=cfg.dim_z, G_attn=cfg.G_attn, G_shared=True, shared_dim=cfg.shared_dim,
               hier=cfg.hier, BN_eps=cfg.BN_eps, G_shared_feat=True, shared_dim_feat=cfg.shared_dim_feat,
               G_lr=hp["G_lr"], G_B1=hp["B1"], G_B2=hp["B2"], **common)
    with contextlib.redirect_stdout(io.StringIO()):
        G = RB.Generator(**gkw)
        G_ema = RB.Generator(**{**gkw, "no_optim": True})
        D = RB.Discriminator(D_ch=cfg.D_ch, D_attn=cfg.D_attn, D_lr=hp["D_lr"], D_B1=hp["B1"], D_B2=hp["B2"], **common)
    gs, ds = MG._check_shapes(G, D, cfg)
    g_sd0, d_sd0 = O.synth_state_dict(gs, hp["seed"]), O.synth_state_dict(ds, hp["seed"] + 1)
    G.load_state_dict({k: v.clone() for k, v in g_sd0.items()})
    D.load_state_dict({k: v.clone() for k, v in d_sd0.items()})
    G.train(); D.train(); G_ema.eval()
    GD = RB.G_D(G, D)
    with contextlib.redirect_stdout(io.StringIO()):
        ema = RU.ema(G, G_ema, hp["ema_decay"], hp["ema_start"])
    calls, pool = step_inputs(cfg, hp)
    it = iter(pool)
    config = dict(toggle_grads=True, num_D_steps=1, num_D_accumulations=hp["n_acc"],
                  num_G_accumulations=hp["n_acc"], split_D=False, DiffAugment=False, DA=False, D_ortho=0.0,
                  G_ortho=0.0, ema=True)
    state = {"itr": 0}
    train = RT.GAN_training_function(G, D, GD, ema, state, config, lambda: next(it), embedded_optimizers=True,
                                     device="cpu", batch_size=hp["batch_size"])
    losses = []
    for (x, y, f) in calls:
        out = train(x, y, f)
        losses.append([out["G_loss"], out["D_loss_real"], out["D_loss_fake"]])
        state["itr"] += 1

    # the oracle's train_step on the same inputs must land on the same state
    st = O.make_step_state({k: v.clone() for k, v in g_sd0.items()}, {k: v.clone() for k, v in d_sd0.items()},
                           G_lr=hp["G_lr"], D_lr=hp["D_lr"], B1=hp["B1"], B2=hp["B2"], adam_eps=hp["adam_eps"], ema=True)
    it2 = iter(pool)
    o_losses = []
    for (x, y, f) in calls:
        out = O.train_step(st, cfg, x, y, f, lambda: next(it2), hp["batch_size"], num_D_acc=hp["n_acc"],
                           num_G_acc=hp["n_acc"], ema_decay=hp["ema_decay"], ema_start=hp["ema_start"])
        o_losses.append([out["G_loss"], out["D_loss_real"], out["D_loss_fake"]])
    # (adam_eps is deliberately large in this fixture, see oracle/step_fixture.py)
    worst = 0.0
    t = hp["n_steps"]
    for tag, net, mine in (("G", G, st.g_sd), ("D", D, st.d_sd)):
        lr = hp["D_lr"] if tag == "D" else hp["G_lr"]
        for k, p in net.named_parameters():
            e = (mine[k].detach() - p.detach()).abs().max().item()
            assert e <= 0.1 * lr, f"oracle train_step vs reference: {tag}.{k} differs by {e:.3e} (lr {lr})"
            worst = max(worst, e / lr)
    for tag, ref_sd, mine in (("G", G.state_dict(), st.g_sd), ("D", D.state_dict(), st.d_sd)):
        for k, v in ref_sd.items():
            if not O.is_param(k, v):
                err = (mine[k].detach() - v).abs().max().item()
                assert err <= 2e-4 * max(1.0, v.abs().max().item()), f"oracle vs reference buffer {tag}.{k}: {err:.3e}"
    for k, v in G_ema.state_dict().items():
        err = (st.ema_sd[k].detach() - v).abs().max().item()
        assert err <= 0.1 * hp["G_lr"] + 2e-4 * max(1.0, v.abs().max().item()), f"oracle vs reference EMA {k}: {err:.3e}"
    assert np.allclose(np.array(losses), np.array(o_losses), atol=2e-4), (losses, o_losses)
    print(f"[step] oracle.train_step == reference train_fns over {hp['n_steps']} steps x {hp['n_acc']} accumulations: "
          f"worst weight difference {worst:.2e} x lr; losses {losses}")

    fx = {"losses": torch.tensor(losses)}
    for tag, sd in (("G", G.state_dict()), ("D", D.state_dict()), ("G_ema", G_ema.state_dict())):
        for k, v in sd.items():
            fx[f"{tag}/{k}"] = sample_of(v.float())
    for tag, net in (("G", G), ("D", D)):
        names = {id(p): k for k, p in net.named_parameters()}
        for p, s in net.optim.state.items():
            fx[f"{tag}_exp_avg/{names[id(p)]}"] = sample_of(s["exp_avg"])
            fx[f"{tag}_exp_avg_sq/{names[id(p)]}"] = sample_of(s["exp_avg_sq"])
    np.savez_compressed(os.path.join(GOLD, "biggan_step_cc32.npz"), **{k: v.numpy() for k, v in fx.items()})
    with open(os.path.join(GOLD, "biggan_step_cc32.json"), "w") as f:
        json.dump({"config": STEP_CFG, "hp": hp, "losses": losses, "oracle_vs_reference_worst_over_lr": worst,
                   "sample_limit": 1024, "g_shapes": gs, "d_shapes": ds}, f, indent=1, sort_keys=True)


def main():
    what = set(sys.argv[1:]) or {"width", "bound", "step"}
    torch.manual_seed(0)
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    RB, RL = MG._import_reference()
    if "step" in what:
        full_step(RB)
    if "width" in what:
        kw, B, seed = TINY["cc256_w96"]
        MG.biggan_case(RB, RL, "cc256_w96", O.BigGANConfig(**kw), B=B, seed=seed)
    if "bound" in what:
        path = os.path.join(GOLD, "biggan_bf16_reference_bound.json")
        rep = json.load(open(path)) if os.path.exists(path) else {}
        for name in [n for n in TINY if n in what or not (what & set(TINY))]:
            rep[name] = bf16_bound(RB, RL, name)
        with open(path, "w") as f:
            json.dump(rep, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
