"""StyleGAN2-ADA network modules and loss (ic_gan_b200/stylegan2/{networks,loss}.py).

* CPU (host logic): state_dict layout equals the reference's; with the package's CUDA ops swapped -- IN THE TEST ONLY --
  for the CPU op oracles, the modules reproduce the golden vectors recorded from the live reference (latent
  bookkeeping, gains, clamps, noise order, truncation, w_avg, minibatch-std, projection, every loss phase incl. both
  double backwards and the RNG call order).  The product path has no such swap: on a CUDA-less host the ops raise.
* GPU: the same modules on the real kernels against the same golden vectors (first GPU run pending: the round's GPU
  budget was spent before these modules existed, hence xfail(strict=False) -- an XPASS is the expected outcome)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from ic_gan_b200.stylegan2 import loss as b200_loss
from ic_gan_b200.stylegan2 import networks as N
from ic_gan_b200.stylegan2.ops import bias_act as m_bias_act
from ic_gan_b200.stylegan2.ops import conv2d_gradfix as m_gradfix
from ic_gan_b200.stylegan2.ops import elementwise as m_elem
from ic_gan_b200.stylegan2.ops import fma as m_fma
from ic_gan_b200.stylegan2.ops import upfirdn2d as m_upfirdn2d
from oracle import stylegan_nets_oracle as O
from oracle import stylegan_ops_oracle as oops
from oracle.make_golden_stylegan_nets import inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load():
    with open(os.path.join(GOLD, "stylegan_nets.json")) as f:
        meta = json.load(f)
    return meta, dict(np.load(os.path.join(GOLD, "stylegan_nets.npz"))), dict(np.load(os.path.join(GOLD, "stylegan_loss.npz")))


def _build(meta, device="cpu", num_fp16_res=0):
    c = meta["cfg"]
    G = N.Generator(z_dim=c["z_dim"], c_dim=0, h_dim=c["h_dim"], w_dim=c["w_dim"], img_resolution=c["img_resolution"],
                    img_channels=3, mapping_kwargs=dict(num_layers=c["map_layers"]),
                    synthesis_kwargs=dict(channel_base=c["channel_base"], channel_max=c["channel_max"],
                                          num_fp16_res=num_fp16_res, conv_clamp=c["conv_clamp"]))
    D = N.Discriminator(c_dim=0, h_dim=c["h_dim"], img_resolution=c["img_resolution"], img_channels=3,
                        channel_base=c["channel_base"], channel_max=c["channel_max"], num_fp16_res=num_fp16_res,
                        conv_clamp=c["conv_clamp"], mapping_kwargs=dict(num_layers=c["d_map_layers"]),
                        epilogue_kwargs=dict(mbstd_group_size=c["mbstd_group_size"]))
    G.load_state_dict(O.synth_state_dict(meta["g_shapes"], meta["g_seed"]), strict=True)
    D.load_state_dict(O.synth_state_dict(meta["d_shapes"], meta["d_seed"]), strict=True)
    return G.to(device), D.to(device)


@pytest.fixture
def cpu_ops(monkeypatch):
    """Swap the package's CUDA ops for the CPU op oracles (test-only; exercises the host logic of the modules)."""
    monkeypatch.setattr(m_bias_act, "bias_act",
                        lambda x, b=None, dim=1, act="linear", alpha=None, gain=None, clamp=None, impl="cuda":
                        oops.bias_act(x, b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp))
    monkeypatch.setattr(m_upfirdn2d, "upfirdn2d",
                        lambda x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl="cuda":
                        oops.upfirdn2d(x, f, up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain))
    monkeypatch.setattr(m_gradfix, "conv2d", F.conv2d)
    monkeypatch.setattr(m_gradfix, "conv_transpose2d", F.conv_transpose2d)
    monkeypatch.setattr(m_fma, "fma", lambda a, b, c: a * b + c)
    monkeypatch.setattr(m_elem, "modulate", lambda x, s, out_dtype=None: x * s.to(x.dtype).reshape(x.shape[0], -1, 1, 1))
    monkeypatch.setattr(m_elem, "chan_dot", lambda a, b: (a * b).sum([2, 3]))

    def mod_bias_act(x, pre=None, noise=None, bias=None, act="linear", alpha=0.2, gain=1.0, clamp=None):
        if pre is not None:
            x = x * pre.to(x.dtype).reshape(x.shape[0], -1, 1, 1)
        if noise is not None:
            x = x + noise.to(x.dtype).reshape(-1, 1, x.shape[2], x.shape[3])
        return oops.bias_act(x, None if bias is None else bias.to(x.dtype), dim=1, act=act, alpha=None, gain=gain, clamp=clamp)
    monkeypatch.setattr(m_elem, "mod_bias_act", mod_bias_act)


def _close(got, ref, what, tol=2e-5):
    got = got.detach().float().cpu().numpy() if isinstance(got, torch.Tensor) else got
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = np.abs(got - ref).max()
    assert err <= tol * max(1.0, np.abs(ref).max()), f"{what}: max-abs err {err:.3e}"


def test_state_dict_layout_matches_reference():
    meta, _, _ = _load()
    G, D = _build(meta)
    assert [(k, list(v.shape)) for k, v in G.state_dict().items()] == [(k, v) for k, v in meta["g_shapes"].items()]
    assert [(k, list(v.shape)) for k, v in D.state_dict().items()] == [(k, v) for k, v in meta["d_shapes"].items()]
    assert G.num_ws == G.synthesis.num_ws == 8 and G.mapping.w_avg.shape == (meta["cfg"]["w_dim"],)


def test_ops_have_no_cpu_fallback():
    meta, _, _ = _load()
    G, _ = _build(meta)
    z, h, _ = inputs()
    with pytest.raises(Exception):
        G(z, None, h, noise_mode="const")


def test_module_logic_on_cpu_oracle_ops(cpu_ops):
    meta, fx, _ = _load()
    G, D = _build(meta)
    z, h, x = inputs()
    G.eval(); D.eval()
    with torch.no_grad():
        _close(G.mapping(z, None, h), fx["ws"], "ws")
        _close(G.mapping(z, None, h, truncation_psi=0.7, truncation_cutoff=3), fx["ws_trunc"], "ws_trunc")
        _close(G(z, None, h, noise_mode="const"), fx["img_const"], "img_const")  # fused modconv (eval)
        _close(G(z, None, h, noise_mode="none"), fx["img_none"], "img_none")
        _close(G(z, None, h, truncation_psi=0.5, noise_mode="const"), fx["img_trunc"], "img_trunc")
        torch.manual_seed(5)
        _close(G(z, None, h, noise_mode="random"), fx["img_random"], "img_random")
        _close(D(x, None, h), fx["d_real"], "d_real")
        _close(D(torch.from_numpy(fx["img_const"]), None, h), fx["d_fake"], "d_fake")
    G.train(); D.train()
    img = G(z, None, h, noise_mode="const")  # non-fused modconv (training)
    _close(img, fx["img_train"], "img_train")
    _close(G.mapping.w_avg, fx["w_avg_after"], "w_avg")
    loss_g = F.softplus(-D(img, None, h)).mean()
    loss_g.backward()
    params = dict(G.named_parameters())
    for k in meta["grad_keys_g"]:
        ref = fx["G_grad/" + k]
        rel = np.linalg.norm(params[k].grad.numpy() - ref) / max(np.linalg.norm(ref), 1e-30)
        assert rel <= 2e-4, f"G grad {k}: {rel:.3e}"


@pytest.mark.parametrize("phase,seed,gain", [("Gmain", 31, 1.0), ("Greg", 32, 4.0), ("Dmain", 33, 1.0), ("Dreg", 34, 16.0)])
def test_loss_phases_on_cpu_oracle_ops(cpu_ops, phase, seed, gain):
    meta, _, lx = _load()
    G, D = _build(meta)
    G.train(); D.train()
    G.mapping.w_avg.copy_(torch.from_numpy(lx[f"{phase}/w_avg_before"]))
    loss = b200_loss.StyleGAN2Loss(torch.device("cpu"), G.mapping, G.synthesis, D, style_mixing_prob=0.9, r1_gamma=10.0,
                                   pl_batch_shrink=2, pl_decay=0.01, pl_weight=2.0)
    loss.pl_mean.fill_(float(lx["pl_mean_before"][0] if phase in ("Gmain", "Greg") else lx["Greg/pl_mean_after"][0]))
    G.requires_grad_(phase.startswith("G")); D.requires_grad_(phase.startswith("D"))
    z, h, x = inputs()
    c0 = torch.zeros(z.shape[0], 0)
    torch.manual_seed(seed)
    loss.accumulate_gradients(phase=phase, real_img=x, real_c=c0, real_h=h, gen_z=z, gen_c=c0, gen_h=h, sync=True, gain=gain)
    net, keys = (G, meta["grad_keys_g"]) if phase.startswith("G") else (D, meta["grad_keys_d"])
    params = dict(net.named_parameters())
    for k in keys:
        ref = lx[f"{phase}/grad/{k}"]
        got = params[k].grad.numpy() if params[k].grad is not None else np.zeros_like(ref)
        rel = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
        assert rel <= 2e-3, f"{phase} grad {k}: rel-L2 {rel:.3e}"
    _close(loss.pl_mean.reshape(1), lx[f"{phase}/pl_mean_after"], "pl_mean", tol=1e-5)
    _close(G.mapping.w_avg, lx[f"{phase}/w_avg_after"], "w_avg", tol=1e-5)


# ------------------------------------------------------------------------------------------------------- GPU

@pytest.mark.gpu
def test_gpu_generator_and_discriminator_fp32(cuda_device):
    meta, fx, _ = _load()
    G, D = _build(meta, cuda_device)
    z, h, x = (t.to(cuda_device) for t in inputs())
    G.eval(); D.eval()
    with torch.no_grad():
        _close(G(z, None, h, noise_mode="const"), fx["img_const"], "img_const", tol=1e-3)
        _close(G(z, None, h, noise_mode="none"), fx["img_none"], "img_none", tol=1e-3)
        _close(D(x, None, h), fx["d_real"], "d_real", tol=1e-3)
    G.train(); D.train()
    img = G(z, None, h, noise_mode="const")
    _close(img, fx["img_train"], "img_train", tol=1e-3)
    F.softplus(-D(img, None, h)).mean().backward()
    params = dict(G.named_parameters())
    for k in meta["grad_keys_g"]:
        ref = fx["G_grad/" + k]
        rel = np.linalg.norm(params[k].grad.cpu().numpy() - ref) / max(np.linalg.norm(ref), 1e-30)
        assert rel <= 5e-3, f"G grad {k}: {rel:.3e}"


@pytest.mark.gpu
def test_gpu_generator_bf16_blocks(cuda_device):
    """num_fp16_res=2: the two highest resolutions compute in bfloat16 (tensor-core conv path)."""
    meta, fx, _ = _load()
    G, D = _build(meta, cuda_device, num_fp16_res=2)
    z, h, x = (t.to(cuda_device) for t in inputs())
    G.eval(); D.eval()
    with torch.no_grad():
        img = G(z, None, h, noise_mode="const")
        assert img.dtype == torch.float32
        _close(img, fx["img_const"], "img_const bf16", tol=8e-2)
        _close(D(x, None, h), fx["d_real"], "d_real bf16", tol=5e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("phase,gain", [("Gmain", 1.0), ("Greg", 4.0), ("Dmain", 1.0), ("Dreg", 16.0)])
def test_gpu_loss_phases_run_and_match_cpu_oracle(cuda_device, phase, gain):
    _phase_vs_oracle(cuda_device, phase, gain, 2e-2)


@pytest.mark.parametrize("phase,gain", [("Gmain", 1.0), ("Greg", 4.0), ("Dmain", 1.0), ("Dreg", 16.0)])
def test_phase_harness_on_cpu_oracle_ops(cpu_ops, phase, gain):
    """The harness of the GPU test above (shared seeded noise for oracle and modules), exercised on the CPU op oracles."""
    _phase_vs_oracle(torch.device("cpu"), phase, gain, 2e-3)


def _phase_vs_oracle(cuda_device, phase, gain, tol):
    """Device RNG streams differ from the CPU's, so every random draw (per-layer synthesis noise, path-length noise) comes
    from ONE seeded CPU generator for both runs (torch.randn / randn_like patched in this test only); style mixing is
    off (its cutoff only selects a slice).  Gradients -- including noise_strength, which is sum(dy * noise) -- are
    compared with the CPU oracle's on the same inputs.  (Round 1 zeroed the noise strengths but still compared the
    noise_strength gradient, which depends on the draw itself: that, not a kernel, made Gmain/Greg fail on the GPU.)"""
    meta, _, _ = _load()
    g_sd = O.synth_state_dict(meta["g_shapes"], meta["g_seed"])
    d_sd = O.synth_state_dict(meta["d_shapes"], meta["d_seed"])
    cfg = O.StyleGANConfig(**meta["cfg"])
    G, D = _build(meta, cuda_device)
    G.load_state_dict(g_sd); G.train(); D.train()
    z, h, x = inputs()
    is_g = phase.startswith("G")
    for k, v in (g_sd if is_g else d_sd).items():
        if v.dtype.is_floating_point and not k.endswith(("resample_filter", "noise_const", "w_avg")):
            v.requires_grad_(True)
    orig, orig_like = torch.randn, torch.randn_like
    gen = [None]

    def fake_randn(*size, **kw):
        shape = size[0] if len(size) == 1 and isinstance(size[0], (list, tuple, torch.Size)) else size
        out = orig(list(shape), generator=gen[0])
        return out.to(kw["device"]) if kw.get("device") is not None else out

    def fake_randn_like(t, **kw):
        return orig(list(t.shape), generator=gen[0]).to(device=t.device, dtype=t.dtype)
    try:
        torch.randn, torch.randn_like = fake_randn, fake_randn_like
        gen[0] = torch.Generator().manual_seed(9)
        O.accumulate_gradients(phase, g_sd, d_sd, cfg, x, None, h, z, None, h, gain, pl_mean=torch.tensor(0.05),
                               style_mixing_prob=0.0)
        loss = b200_loss.StyleGAN2Loss(cuda_device, G.mapping, G.synthesis, D, style_mixing_prob=0.0)
        loss.pl_mean.fill_(0.05)
        G.requires_grad_(is_g); D.requires_grad_(not is_g)
        c0 = torch.zeros(z.shape[0], 0, device=cuda_device)
        gen[0] = torch.Generator().manual_seed(9)
        loss.accumulate_gradients(phase=phase, real_img=x.to(cuda_device), real_c=c0, real_h=h.to(cuda_device),
                                  gen_z=z.to(cuda_device), gen_c=c0, gen_h=h.to(cuda_device), sync=True, gain=gain)
    finally:
        torch.randn, torch.randn_like = orig, orig_like
    net, keys, sd = (G, meta["grad_keys_g"], g_sd) if is_g else (D, meta["grad_keys_d"], d_sd)
    params = dict(net.named_parameters())
    report = []
    for k in sorted(params):  # every parameter, not only the golden's key subset
        ref = sd[k].grad
        if ref is None or ref.abs().max() == 0:
            continue
        got = params[k].grad
        assert got is not None, f"{phase}: no gradient for {k}"
        rel = float((got.cpu() - ref).norm() / ref.norm().clamp_min(1e-30))
        report.append((rel, k))
    report.sort(reverse=True)
    print(f"{phase}: worst gradient rel-L2 vs CPU oracle: " + ", ".join(f"{k} {r:.2e}" for r, k in report[:4]))
    bad = [(r, k) for r, k in report if r > tol]
    assert not bad, f"{phase}: {bad[:6]}"


# ------------------------------------------------------------------------------------------------------- emulated kernels
@pytest.fixture
def emu(monkeypatch):
    """The REAL op stack of the package (autograd Functions, tap tables, layouts, split-bf16 sequencing) over a CPU
    emulation of the C ABI (tests/kernel_emulator.py): exercises everything above the kernels without a GPU."""
    from tests.kernel_emulator import emulated
    with emulated(monkeypatch):
        yield


def test_modules_on_emulated_kernels(emu):
    meta, fx, _ = _load()
    G, D = _build(meta)
    z, h, x = inputs()
    G.eval(); D.eval()
    with torch.no_grad():
        _close(G(z, None, h, noise_mode="const"), fx["img_const"], "img_const", tol=3e-4)
        _close(G(z, None, h, noise_mode="none"), fx["img_none"], "img_none", tol=3e-4)
        _close(D(x, None, h), fx["d_real"], "d_real", tol=3e-4)
    G.train(); D.train()
    img = G(z, None, h, noise_mode="const")
    _close(img, fx["img_train"], "img_train", tol=3e-4)
    F.softplus(-D(img, None, h)).mean().backward()
    params = dict(G.named_parameters())
    for k in meta["grad_keys_g"]:
        ref = fx["G_grad/" + k]
        rel = np.linalg.norm(params[k].grad.numpy() - ref) / max(np.linalg.norm(ref), 1e-30)
        assert rel <= 2e-3, f"G grad {k}: {rel:.3e}"


def test_bf16_blocks_on_emulated_kernels(emu):
    meta, fx, _ = _load()
    G, D = _build(meta, num_fp16_res=2)
    z, h, x = inputs()
    G.eval(); D.eval()
    with torch.no_grad():
        _close(G(z, None, h, noise_mode="const"), fx["img_const"], "img_const bf16", tol=8e-2)
        _close(D(x, None, h), fx["d_real"], "d_real bf16", tol=5e-2)


@pytest.mark.parametrize("phase,gain", [("Gmain", 1.0), ("Greg", 4.0), ("Dmain", 1.0), ("Dreg", 16.0)])
def test_loss_phases_on_emulated_kernels(emu, phase, gain):
    """All four phases -- both double backwards included -- through the package's own Functions."""
    _phase_vs_oracle(torch.device("cpu"), phase, gain, 5e-3)


@pytest.mark.gpu
def test_gpu_graphed_phase_matches_eager(cuda_device):
    """A loss phase replayed from a CUDA graph leaves the same gradients as the eager call (random inputs of the phase
    switched off: zero noise strengths, no style mixing), and different inputs through the SAME graph give the eager
    result for those inputs."""
    from ic_gan_b200.stylegan2.graphs import GraphedLoss
    meta, _, _ = _load()
    g_sd = O.synth_state_dict(meta["g_shapes"], meta["g_seed"])
    for k in g_sd:
        if k.endswith("noise_strength"):
            g_sd[k] = torch.zeros_like(g_sd[k])
    G, D = _build(meta, cuda_device, num_fp16_res=2)
    G.load_state_dict(g_sd); G.train(); D.train()
    G.requires_grad_(False); D.requires_grad_(False)
    for p in list(G.parameters()) + list(D.parameters()):
        p.grad = torch.zeros_like(p)
    z, h, x = (t.to(cuda_device) for t in inputs())
    c0 = torch.zeros(z.shape[0], 0, device=cuda_device)
    eager = b200_loss.StyleGAN2Loss(cuda_device, G.mapping, G.synthesis, D, style_mixing_prob=0.0)
    graphed = GraphedLoss(b200_loss.StyleGAN2Loss(cuda_device, G.mapping, G.synthesis, D, style_mixing_prob=0.0), {"G": G, "D": D})

    def grads(runner, zz, toggle):
        for p in G.parameters():
            p.grad.zero_()
        if toggle:
            G.requires_grad_(True)
        runner.accumulate_gradients(phase="Gmain", real_img=x, real_c=c0, real_h=h, gen_z=zz, gen_c=c0, gen_h=h, sync=True, gain=1.0)
        G.requires_grad_(False)
        return {k: p.grad.clone() for k, p in G.named_parameters() if not k.endswith("noise_strength")}

    graphed.accumulate_gradients(phase="Gmain", real_img=x, real_c=c0, real_h=h, gen_z=z, gen_c=c0, gen_h=h, sync=True, gain=1.0)
    for zz in (z, z.flip(0) * 0.5):
        want, got = grads(eager, zz, True), grads(graphed, zz, True)
        for k in want:
            if want[k].abs().max() > 0:
                rel = float((got[k] - want[k]).norm() / want[k].norm())
                assert rel <= 1e-2, f"{k}: graphed vs eager rel-L2 {rel:.3e}"
