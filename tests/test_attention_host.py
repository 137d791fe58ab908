"""Host logic of the fused non-local block (ops.AttentionCoreFn -> icgan_attn_fwd / icgan_attn_bwd_q / icgan_attn_bwd_kv):
argument order, the nullable log-sum-exp of the no-grad forward, the rowsum hand-over between the two backward calls --
on a CPU emulation of the three entry points written from the header's contracts (tests/kernel_emulator.py).  The
kernels themselves are checked on the GPU (tests/test_attention_gpu.py)."""
import pytest
import torch

from tests.kernel_emulator import emulated


@pytest.mark.parametrize("shape", [(2, 256, 128, 16, 32), (1, 128, 128, 8, 16)], ids=lambda s: "x".join(map(str, s)))
def test_fused_attention_function_on_emulated_kernels(monkeypatch, shape):
    from ic_gan_b200 import ops
    B, Q, Kk, d, dv = shape
    gen = torch.Generator().manual_seed(3)
    mk = lambda *s: torch.randn(*s, generator=gen).bfloat16()
    theta, phi, g, do = mk(B, Q, d), mk(B, Kk, d), mk(B, Kk, dv), mk(B, Q, dv)
    t, p, v = (x.float().requires_grad_(True) for x in (theta, phi, g))
    ref = torch.softmax(t @ p.transpose(1, 2), -1) @ v
    ref.backward(do.float())
    with emulated(monkeypatch):
        monkeypatch.setattr(ops, "FUSED_ATTENTION", True)
        with torch.no_grad():
            o_ng = ops.AttentionCoreFn.apply(theta, phi, g)          # no log-sum-exp requested
        a, b, c = (x.clone().requires_grad_(True) for x in (theta, phi, g))
        o = ops.AttentionCoreFn.apply(a, b, c)
        o.backward(do)
    rel = lambda x, y: float((x.detach().float() - y.detach()).norm() / y.detach().norm())
    assert torch.equal(o_ng, o.detach())
    assert rel(o, ref.detach()) < 6e-3
    assert rel(a.grad, t.grad) < 1e-2 and rel(b.grad, p.grad) < 1e-2 and rel(c.grad, v.grad) < 1e-2
