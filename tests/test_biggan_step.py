"""Full-step parity: three ``GAN_training_function(...).train`` calls with two gradient
accumulations each -- D update, G update, embedded Adam optimisers, EMA over every state entry with its start_itr quirk
-- against tests/golden/biggan_step_cc32.*, frozen from the LIVE reference's own train_fns.py / utils.ema by
oracle/make_golden_r2.py.  The CPU test pins ``oracle.train_step``; the GPU test checks ``ic_gan_b200.biggan.train_fns``."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import biggan_oracle as O
from oracle.step_fixture import sample_of, step_inputs
from tests.helpers import GOLD, model_kwargs


def _load():
    with open(os.path.join(GOLD, "biggan_step_cc32.json")) as f:
        meta = json.load(f)
    data = np.load(os.path.join(GOLD, "biggan_step_cc32.npz"))
    return meta, {k: torch.from_numpy(data[k]) for k in data.files}


def _check(tag, got_sd, fx, lr, param_tol_lr, buf_tol=2e-4, hp=None):
    """Parameters: |w - w_ref| <= param_tol_lr x lr on every element whose gradient is well above adam_eps (second moment
    from the reference's own optimiser state); elements with |g| ~ adam_eps or below sit where the Adam update is steep
    in g (d step / d g ~ lr / eps), so fp32 summation-order noise alone moves them by a sizeable fraction of lr
    (the class embedding `shared.weight`, |g| ~ 1e-5: 0.03 x lr between the CPU oracle and the reference, 0.35 x lr on
    the GPU, 1.3 x lr in another run) -- those are held to n_steps x lr, i.e. only to "finite and the right magnitude"."""
    worst, bad = 0.0, []
    for k, v in got_sd.items():
        ref = fx[f"{tag}/{k}"]
        got = sample_of(v.detach().float().cpu())
        diff = (got - ref).abs()
        err = diff.max().item()
        if O.is_param(k, v):
            vkey = f"{'G' if tag == 'G_ema' else tag}_exp_avg_sq/{k}"
            if hp is not None and vkey in fx:
                vhat = (fx[vkey] / (1 - hp["B2"] ** hp["n_steps"])).sqrt()
                good = vhat > 30 * hp["adam_eps"]
                err = diff[good].max().item() if good.any() else 0.0
                loose = diff[~good].max().item() if (~good).any() else 0.0
                if loose > 3.0 * lr:  # at most one full Adam step per call in the fixture
                    bad.append(f"{tag}.{k}: ill-conditioned elements off by {loose / lr:.3f} x lr")
            worst = max(worst, err / lr)
            if err > param_tol_lr * lr:
                bad.append(f"{tag}.{k}: |w - w_ref| = {err / lr:.3f} x lr")
        elif err > buf_tol * max(1.0, ref.abs().max().item()):
            bad.append(f"{tag}.{k} (buffer): {err:.3e}")
    assert not bad, f"{len(bad)} entries off: " + "; ".join(bad[:12])
    return worst


def test_oracle_train_step_matches_reference_golden():
    meta, fx = _load()
    cfg, hp = O.BigGANConfig(**meta["config"]), meta["hp"]
    g_sd, d_sd = O.synth_state_dict(meta["g_shapes"], hp["seed"]), O.synth_state_dict(meta["d_shapes"], hp["seed"] + 1)
    st = O.make_step_state(g_sd, d_sd, G_lr=hp["G_lr"], D_lr=hp["D_lr"], B1=hp["B1"], B2=hp["B2"],
                           adam_eps=hp["adam_eps"], ema=True)
    calls, pool = step_inputs(cfg, hp)
    it = iter(pool)
    losses = []
    for (x, y, f) in calls:
        out = O.train_step(st, cfg, x, y, f, lambda: next(it), hp["batch_size"], num_D_acc=hp["n_acc"],
                           num_G_acc=hp["n_acc"], ema_decay=hp["ema_decay"], ema_start=hp["ema_start"])
        losses.append([out["G_loss"], out["D_loss_real"], out["D_loss_fake"]])
    assert np.allclose(np.array(losses), fx["losses"].numpy(), atol=2e-4)
    _check("G", st.g_sd, fx, hp["G_lr"], 0.1)
    _check("D", st.d_sd, fx, hp["D_lr"], 0.1)
    _check("G_ema", st.ema_sd, fx, hp["G_lr"], 0.1)
    for tag, opt, sd in (("G", st.opt_g, st.g_sd), ("D", st.opt_d, st.d_sd)):
        names = {id(v): k for k, v in sd.items()}
        for p, s in opt.state.items():
            k = names[id(p)]
            for mom in ("exp_avg", "exp_avg_sq"):
                ref = fx[f"{tag}_{mom}/{k}"]
                got = sample_of(s[mom])
                assert (got - ref).norm() <= 5e-3 * ref.norm() + 1e-6 * ref.numel() ** 0.5, f"{tag} {mom} {k}"  # floor: noise-only grads


@pytest.mark.gpu
@pytest.mark.parametrize("graphs", [False, True], ids=["eager", "cuda_graphs"])
@pytest.mark.parametrize("cdt", [torch.float32, torch.bfloat16])
def test_gpu_training_function_matches_reference_golden(cuda_device, cdt, graphs):
    """graphs=True: every micro-step is captured into a CUDA graph on first use (biggan/graphs.py) -- the very first call
    of the test captures, so the stash / restore around the warm-up executions is held to the same golden."""
    from ic_gan_b200.biggan import G_D, Discriminator, Generator, train_fns
    meta, fx = _load()
    cfg, hp = O.BigGANConfig(**meta["config"]), meta["hp"]
    dev = cuda_device
    kw = model_kwargs(cfg)
    okw = dict(adam_eps=hp["adam_eps"], compute_dtype=cdt)
    G = Generator(G_lr=hp["G_lr"], G_B1=hp["B1"], G_B2=hp["B2"], **okw, **kw)
    G_ema = Generator(no_optim=True, **okw, **kw)
    D = Discriminator(D_lr=hp["D_lr"], D_B1=hp["B1"], D_B2=hp["B2"], **okw, **kw)
    g_sd0 = O.synth_state_dict(meta["g_shapes"], hp["seed"])
    d_sd0 = O.synth_state_dict(meta["d_shapes"], hp["seed"] + 1)
    G.load_state_dict(g_sd0, strict=True)
    D.load_state_dict(d_sd0, strict=True)
    G, D, G_ema = G.to(dev), D.to(dev), G_ema.to(dev)
    # the optimisers were built on the CPU parameters; .to() keeps Parameter identity for nn.Module, state is still empty
    G.train(); D.train(); G_ema.eval()
    GD = G_D(G, D)
    ema = train_fns.ema(G, G_ema, hp["ema_decay"], hp["ema_start"])
    calls, pool = step_inputs(cfg, hp)
    it = iter(pool)
    config = dict(toggle_grads=True, num_D_steps=1, num_D_accumulations=hp["n_acc"], num_G_accumulations=hp["n_acc"],
                  split_D=False, DiffAugment=False, DA=False, D_ortho=0.0, G_ortho=0.0, ema=True)
    state = {"itr": 0}
    train = train_fns.GAN_training_function(G, D, GD, ema, state, config, lambda: next(it), embedded_optimizers=True,
                                            device=dev, batch_size=hp["batch_size"], graphs=graphs)
    losses = []
    for (x, y, f) in calls:
        out = train(x.to(dev), y.to(dev), f.to(dev))
        assert all(isinstance(v, float) for v in out.values())  # the reference returns Python floats (train_fns.py:183-187)
        losses.append([out["G_loss"], out["D_loss_real"], out["D_loss_fake"]])
        state["itr"] += 1
    if graphs:
        assert len(train.graphs.graphs) == 2 and train.graphs.replayed_launches > 0  # one D graph, one G graph, replayed
    ref_losses = fx["losses"].numpy()
    print(f"step {cdt}: losses {losses} vs reference {ref_losses.tolist()}")
    def update_error(tag, net_sd, sd0, lr):
        """worst per-tensor rel-L2 of (w_after - w_before) against the reference's update; tensors whose reference
        update is rounding noise (e.g. conv biases feeding a batch norm) are skipped."""
        worst, worst_k, num, den = 0.0, "", 0.0, 0.0
        for k, w0 in sd0.items():
            if f"{tag}/{k}" not in fx or not O.is_param(k, net_sd[k]):
                continue
            w0s = sample_of(w0)
            upd_ref = fx[f"{tag}/{k}"] - w0s
            if upd_ref.norm() < 0.05 * lr * upd_ref.numel() ** 0.5:
                continue
            err = float((sample_of(net_sd[k].detach().float().cpu()) - w0s - upd_ref).norm())
            num, den = num + err ** 2, den + float(upd_ref.norm()) ** 2
            if err / float(upd_ref.norm()) > worst:
                worst, worst_k = err / float(upd_ref.norm()), f"{tag}.{k}"
        return worst, worst_k, (num / max(den, 1e-30)) ** 0.5, num, max(den, 1e-30)

    if cdt == torch.float32:
        assert np.allclose(np.array(losses), ref_losses, atol=2e-3 * max(1.0, np.abs(ref_losses).max()))
        # The GPU's float32 summation orders (atomics in wgrad / embedding backward) differ run to run, and this fixture
        # (lr 1e-3 / 2e-3 against adam_eps 1e-4) amplifies that: D's first Adam step turns 4e-6 relative gradient noise
        # into 1.4e-5 in the weights, G's gradient through that D differs by 2e-4, and from there the trajectory lands in
        # one of a few discrete branches.  Sixteen runs, eager and CUDA-graph mode alike (profiles/r02_step_parity_spread.txt):
        # worst well-conditioned element 0.003 .. 0.28 x lr in eleven of them, 1.1 .. 2.0 x lr in five; worst per-tensor
        # update rel-L2 9e-4 .. 9.1e-2 in the eight runs that recorded it.  So: every tensor's UPDATE within 0.5 rel-L2 of
        # the reference's and all tensors together within 0.25 (a skipped update is 1.0, a wrong schedule -- EMA start,
        # accumulation, toggling, optimiser order -- changes updates by O(1)), and every element within the 3 x lr that
        # three Adam steps can move it.
        wg = _check("G", G.state_dict(), fx, hp["G_lr"], 3.0, buf_tol=5e-3, hp=hp)
        wd = _check("D", D.state_dict(), fx, hp["D_lr"], 3.0, buf_tol=5e-3, hp=hp)
        we = _check("G_ema", G_ema.state_dict(), fx, hp["G_lr"], 3.0, buf_tol=5e-3, hp=hp)
        print(f"step fp32: worst |w - w_ref| / lr: G {wg:.3e}, D {wd:.3e}, G_ema {we:.3e}")
        ug = update_error("G", G.state_dict(), g_sd0, hp["G_lr"])
        ud = update_error("D", D.state_dict(), d_sd0, hp["D_lr"])
        ue = update_error("G_ema", G_ema.state_dict(), g_sd0, hp["G_lr"])
        print(f"step fp32: worst per-tensor update rel-L2: G {ug[0]:.3e} ({ug[1]}), D {ud[0]:.3e} ({ud[1]}), "
              f"G_ema {ue[0]:.3e} ({ue[1]}); all tensors together G {ug[2]:.3e}, D {ud[2]:.3e}, G_ema {ue[2]:.3e}")
        assert max(ug[0], ud[0], ue[0]) <= 0.5 and max(ug[2], ud[2], ue[2]) <= 0.25, (ug, ud, ue)
        for tag, net in (("G", G), ("D", D)):
            for k, p in net.named_parameters():
                for mom in ("exp_avg", "exp_avg_sq"):
                    ref = fx[f"{tag}_{mom}/{k}"]
                    got = sample_of(net.optim.state[p][mom].float().cpu())
                    # (halved gradients -- a dropped accumulation -- would be 0.5 / 0.75 off; the trajectory branches above
                    # move the last step's gradients by a few per cent)
                    assert (got - ref).norm() <= 0.35 * ref.norm() + 1e-5 * ref.numel() ** 0.5, f"{tag} {mom} {k}"
    else:
        # bf16 tensor-core mode: per-element agreement of an Adam update is not a meaningful bar (a gradient element whose
        # bf16 noise exceeds adam_eps moves by a different fraction of lr); hold the losses and the per-tensor UPDATE
        # direction instead: rel-L2 of (w_after - w_before) against the reference's update, worst tensor printed.
        assert np.allclose(np.array(losses), ref_losses, atol=0.05 * max(1.0, np.abs(ref_losses).max()))
        ug = update_error("G", G.state_dict(), g_sd0, hp["G_lr"])
        ud = update_error("D", D.state_dict(), d_sd0, hp["D_lr"])
        together = ((ug[3] + ud[3]) / (ug[4] + ud[4])) ** 0.5  # every parameter of G and D as one vector
        print(f"step bf16: update rel-L2: worst tensor G {ug[0]:.3e} ({ug[1]}), D {ud[0]:.3e} ({ud[1]}); "
              f"G {ug[2]:.3e}, D {ud[2]:.3e}, all tensors together {together:.3e}")
        # Ten runs (profiles/r02_step_parity_spread.txt, eager and graph mode): all tensors together 0.169 .. 0.176; the
        # worst single tensor is always a handful of elements (a bias of 16, the attention gamma scalar) and moves between
        # 0.40 and 0.59 from run to run -- it is held below "update skipped" (1.0), the aggregate to 0.3.
        assert together <= 0.3 and max(ug[0], ud[0]) <= 0.9, (together, ug, ud)
