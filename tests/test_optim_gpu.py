"""Fused Adam + EMA kernel against torch.optim.Adam and utils.ema's update rule."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ulp_diff(a: torch.Tensor, b: torch.Tensor) -> int:
    ia, ib = a.contiguous().view(torch.int32).long(), b.contiguous().view(torch.int32).long()
    ia = torch.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = torch.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return int((ia - ib).abs().max())


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
    worst_ulp = 0
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
            worst_ulp = max(worst_ulp, _ulp_diff(mp.detach(), rp.detach()))
            assert torch.allclose(mp.detach(), rp.detach(), rtol=2e-6, atol=1e-9)
            st = ref.state[rp]
            assert torch.allclose(mine.state[mp]["exp_avg"], st["exp_avg"], rtol=1e-6, atol=0)
            assert torch.allclose(mine.state[mp]["exp_avg_sq"], st["exp_avg_sq"], rtol=1e-6, atol=0)
            assert torch.allclose(e.detach(), er, rtol=2e-6, atol=1e-9)
    print(f"fused Adam vs torch.optim.Adam(foreach=False) betas={betas}: worst parameter difference {worst_ulp} ulp")
    assert worst_ulp <= 4


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
