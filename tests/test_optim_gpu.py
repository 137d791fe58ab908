"""Fused Adam + EMA kernel against torch.optim.Adam and utils.ema's update rule."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _err_over_lr(a: torch.Tensor, b: torch.Tensor, lr: float) -> float:
    """max |a - b| in units of one optimiser step: the ulp of a weight near zero says nothing about the update."""
    return float((a - b).abs().max()) / lr


@pytest.mark.parametrize("betas,eps", [((0.0, 0.999), 1e-6), ((0.5, 0.99), 1e-8)])
def test_fused_adam_matches_torch_adam(cuda_device, betas, eps):
    from ic_gan_b200.optim import FusedAdamEMA
    g = torch.Generator(device=cuda_device).manual_seed(3)
    shapes = [(96, 48, 3, 3), (7,), (), (33, 5), (1536, 657), (3, 96, 3, 3)]
    ref_p = [torch.randn(s, device=cuda_device, generator=g).requires_grad_(True) for s in shapes]
    my_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ema_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ema_ref = [p.detach().clone() for p in ref_p]
    ref = torch.optim.Adam(ref_p, lr=2e-3, betas=betas, eps=eps, weight_decay=0, foreach=False, fused=False)
    mine = FusedAdamEMA(my_p, lr=2e-3, betas=betas, eps=eps, ema_params=ema_p)
    decay = 0.9
    worst, same, total = 0.0, 0, 0
    lr = 2e-3
    for step in range(4):
        for rp, mp in zip(ref_p, my_p):
            gr = torch.randn(rp.shape, device=cuda_device, generator=g) * (10.0 ** (step - 2))
            rp.grad = gr.clone()
            mp.grad.copy_(gr)  # the flat gradient buffer the parameter's .grad aliases
        ref.step()
        mine.set_ema_decay(decay if step != 1 else 0.0)
        mine.step()
        d = decay if step != 1 else 0.0
        for e, rp in zip(ema_ref, ref_p):
            e.copy_(e * d + rp.detach() * (1 - d))  # utils.ema.update (utils.py:1062-1066)
        for rp, mp, e, er in zip(ref_p, my_p, ema_p, ema_ref):
            st = ref.state[rp]
            e_m = float((mine.state[mp]["exp_avg"] - st["exp_avg"]).abs().max() / st["exp_avg"].abs().max().clamp_min(1e-30))
            e_v = float((mine.state[mp]["exp_avg_sq"] - st["exp_avg_sq"]).abs().max() / st["exp_avg_sq"].abs().max().clamp_min(1e-30))
            e_p = _err_over_lr(mp.detach(), rp.detach(), lr)
            e_e = _err_over_lr(e.detach(), er, lr)
            worst = max(worst, e_p)
            same += int((mp.detach() == rp.detach()).sum())
            total += rp.numel()
            assert e_m <= 1e-6 and e_v <= 1e-6, f"step {step} shape {tuple(rp.shape)}: moments differ {e_m:.2e} {e_v:.2e}"
            assert e_p <= 1e-3, f"step {step} shape {tuple(rp.shape)}: |p - p_torch| = {e_p:.3e} x lr"
            assert e_e <= 1e-3, f"step {step} shape {tuple(rp.shape)}: EMA differs by {e_e:.3e} x lr"
    print(f"fused Adam vs torch.optim.Adam(foreach=False) betas={betas}: worst |p - p_torch| = {worst:.2e} x lr, "
          f"{100.0 * same / total:.2f}% of the weights bit-identical")


def test_grad_scale_and_zero_grad(cuda_device):
    from ic_gan_b200.optim import FusedAdamEMA
    p = torch.nn.Parameter(torch.ones(1000, device=cuda_device))
    q = torch.nn.Parameter(torch.ones(1000, device=cuda_device))
    a, b = FusedAdamEMA([p], lr=1e-2, eps=1e-3), FusedAdamEMA([q], lr=1e-2, eps=1e-3)
    gr = torch.linspace(-1, 1, 1000, device=cuda_device)
    p.grad.copy_(gr * 4)
    a.set_grad_scale(0.25)
    a.step()
    q.grad.copy_(gr)
    b.step()
    assert torch.equal(p.detach(), q.detach())
    a.zero_grad()
    assert float(p.grad.abs().max()) == 0 and p.grad.data_ptr() == a.flat_g.data_ptr()


def test_buffer_ema(cuda_device):
    from ic_gan_b200.optim import FlatBufferEMA
    src = torch.nn.BatchNorm2d(5).to(cuda_device)
    tgt = torch.nn.BatchNorm2d(5).to(cuda_device)
    src.running_mean.normal_(); src.running_var.uniform_(0.5, 2)
    want_m = tgt.running_mean * 0.9 + src.running_mean * (1 - 0.9)
    e = FlatBufferEMA(src, tgt)
    e.update(0.9)
    assert torch.allclose(tgt.running_mean, want_m, rtol=1e-6)
    assert int(tgt.num_batches_tracked) == 0  # integer entries are left alone
