"""Eager vs CUDA-graph micro-steps on two identical copies of the cc32 fixture networks: prints the largest difference
of gradients / parameters / buffers after every phase of three steps.  python scripts/debug/graph_vs_eager.py [f32|bf16]"""
import copy
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ic_gan_b200.biggan import G_D, Discriminator, Generator, train_fns  # noqa: E402
from oracle import biggan_oracle as O  # noqa: E402  (test infrastructure: fixture shapes and inputs only)
from oracle.step_fixture import step_inputs  # noqa: E402
from tests.helpers import GOLD, model_kwargs  # noqa: E402


def build(meta, cdt, dev):
    cfg, hp = O.BigGANConfig(**meta["config"]), meta["hp"]
    kw = model_kwargs(cfg)
    okw = dict(adam_eps=hp["adam_eps"], compute_dtype=cdt)
    G = Generator(G_lr=hp["G_lr"], G_B1=hp["B1"], G_B2=hp["B2"], **okw, **kw)
    D = Discriminator(D_lr=hp["D_lr"], D_B1=hp["B1"], D_B2=hp["B2"], **okw, **kw)
    G_ema = Generator(no_optim=True, **okw, **kw)
    G.load_state_dict(O.synth_state_dict(meta["g_shapes"], hp["seed"]))
    D.load_state_dict(O.synth_state_dict(meta["d_shapes"], hp["seed"] + 1))
    G, D, G_ema = G.to(dev), D.to(dev), G_ema.to(dev)
    G.train(); D.train(); G_ema.eval()
    return cfg, hp, G, D, G_ema


def diff(tag, a, b):
    worst, name = 0.0, ""
    for (k, x), (_, y) in zip(a, b):
        if x is None or y is None:
            if (x is None) != (y is None):
                print(f"   {tag} {k}: one side is None")
            continue
        d = float((x.float() - y.float()).abs().max() / (y.float().abs().max() + 1e-12))
        if d > worst:
            worst, name = d, k
    print(f"   {tag:28s} worst relative difference {worst:.3e}  ({name})")


def main():
    cdt = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float32
    dev = torch.device("cuda:0")
    with open(os.path.join(GOLD, "biggan_step_cc32.json")) as f:
        meta = json.load(f)
    runs = {}
    for mode in ("eager", "graphs"):
        cfg, hp, G, D, G_ema = build(meta, cdt, dev)
        GD = G_D(G, D)
        ema = train_fns.ema(G, G_ema, hp["ema_decay"], hp["ema_start"])
        calls, pool = step_inputs(cfg, hp)
        it = iter(pool)
        config = dict(toggle_grads=True, num_D_steps=1, num_D_accumulations=hp["n_acc"], num_G_accumulations=hp["n_acc"],
                      split_D=False, DiffAugment=False, DA=False, D_ortho=0.0, G_ortho=0.0, ema=True)
        state = {"itr": 0}
        train = train_fns.GAN_training_function(G, D, GD, ema, state, config, lambda: next(it), embedded_optimizers=True,
                                                device=dev, batch_size=hp["batch_size"], graphs=mode.startswith("graphs"))
        snaps = []
        for (x, y, f) in calls:
            out = train(x.to(dev), y.to(dev), f.to(dev))
            state["itr"] += 1
            torch.cuda.synchronize()
            snaps.append({
                "losses": out,
                "G.grad": [(k, p.grad.clone() if p.grad is not None else None) for k, p in G.named_parameters()],
                "D.grad": [(k, p.grad.clone() if p.grad is not None else None) for k, p in D.named_parameters()],
                "G.param": [(k, p.detach().clone()) for k, p in G.named_parameters()],
                "D.param": [(k, p.detach().clone()) for k, p in D.named_parameters()],
                "G.buffer": [(k, b.clone()) for k, b in G.named_buffers()],
                "D.buffer": [(k, b.clone()) for k, b in D.named_buffers()],
                "G_ema": [(k, v.clone()) for k, v in G_ema.state_dict().items()],
                "G.exp_avg": [(k, G.optim.state[p]["exp_avg"].clone()) for k, p in G.named_parameters()],
                "D.exp_avg": [(k, D.optim.state[p]["exp_avg"].clone()) for k, p in D.named_parameters()],
            })
        runs[mode] = snaps
    for mode in ("graphs",):
      for i, (a, b) in enumerate(zip(runs[mode], runs["eager"])):
        print(f"{mode} step {i}: losses {a['losses']} eager {b['losses']}")
        for tag in ("D.grad", "G.grad", "D.param", "G.param", "D.buffer", "G.buffer", "G_ema", "G.exp_avg", "D.exp_avg"):
            diff(tag, a[tag], b[tag])
        diff("G.exp_avg vs own G.grad", a["G.exp_avg"], a["G.grad"])
        if i == 0 and False:
            rows = []
            for (k, x), (_, y) in zip(a["G.grad"], b["G.grad"]):
                rows.append((float((x - y).norm() / (y.norm() + 1e-30)), float(x.norm() / (y.norm() + 1e-30)), float(y.norm()), k))
            for e, r, n, k in sorted(rows, reverse=True)[:40]:
                print(f"      G.grad {k:40s} rel-L2 {e:.3e}  |graph|/|eager| {r:.4f}  |eager| {n:.3e}")


if __name__ == "__main__":
    main()
