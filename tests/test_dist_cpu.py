"""Host-side data-parallel logic on CPU with the gloo backend (world_size 2): flat-gradient all-reduce (one per network
per step), rank-0 buffer broadcast, and the GAN_training_function schedule calling them exactly twice per step."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class ToyG(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = nn.Linear(4, 12)
        self.register_buffer("stat", torch.zeros(3))
        self.dim_z = 4

    def forward(self, z, label=None, feats=None):
        self.stat += 1
        return torch.tanh(self.fc(z + feats[:, :4])).view(-1, 3, 2, 2)


class ToyD(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = nn.Linear(12, 1)
        self.register_buffer("u0", torch.randn(1, 5))

    def forward(self, x, y=None, feat=None):
        return self.fc(x.reshape(x.shape[0], -1)) + feat[:, :1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ic_gan_b200.biggan import train_fns
    from ic_gan_b200.biggan.model import G_D
    from ic_gan_b200.dist import GradSync
    torch.manual_seed(100 + rank)  # deliberately different initial weights / buffers per rank
    G, D = ToyG(), ToyD()
    sync = GradSync({"G": G, "D": D}, world)
    sync.broadcast_params()
    w0 = torch.cat([p.detach().flatten() for p in list(G.parameters()) + list(D.parameters())])
    gathered = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(gathered, w0)
    assert torch.equal(gathered[0], gathered[1]), "broadcast_params did not replicate rank 0"
    ub = [torch.zeros_like(D.u0) for _ in range(world)]
    dist.all_gather(ub, D.u0)
    assert torch.equal(ub[0], ub[1])
    opt_G = torch.optim.SGD(G.parameters(), lr=0.1)
    opt_D = torch.optim.SGD(D.parameters(), lr=0.1)
    GD = G_D(G, D, opt_G, opt_D)
    cfg = {"toggle_grads": True, "num_D_steps": 1, "num_D_accumulations": 2, "num_G_accumulations": 2, "split_D": False,
           "ema": False, "D_ortho": 0.0, "G_ortho": 0.0}
    gen = torch.Generator().manual_seed(7 + rank)  # each rank draws its own shard of the batch
    import copy
    G0, D0 = copy.deepcopy(G), copy.deepcopy(D)  # the replicated starting point, for the single-process replay below
    draws = []

    def sample():
        d = (torch.randn(3, 4, generator=gen), torch.randn(3, 8, generator=gen))
        draws.append(d)
        return d
    train = train_fns.GAN_training_function(G, D, GD, None, {"itr": 0}, cfg, sample, embedded_optimizers=False,
                                            device="cpu", batch_size=3, grad_sync=sync)
    x = torch.randn(6, 3, 2, 2, generator=gen)
    f = torch.randn(6, 8, generator=gen)
    train(x, None, f)
    # --- the synchronised 2-rank step must equal ONE process stepping over both ranks' micro-batches (4 accumulations):
    # mean over ranks of the per-rank accumulated gradient == mean over all micro-batches (train_fns.py:103-106, 159-162)
    everything = [None] * world
    dist.all_gather_object(everything, {"draws": draws, "x": x, "f": f})
    if rank == 0:
        assert all(len(e["draws"]) == 4 for e in everything)  # D acc 0,1 then G acc 0,1 on every rank
        replay = [everything[r]["draws"][i] for r in range(world) for i in (0, 1)] + \
                 [everything[r]["draws"][i] for r in range(world) for i in (2, 3)]
        it = iter(replay)
        o_G, o_D = torch.optim.SGD(G0.parameters(), lr=0.1), torch.optim.SGD(D0.parameters(), lr=0.1)
        cfg1 = dict(cfg, num_D_accumulations=2 * world, num_G_accumulations=2 * world)
        single = train_fns.GAN_training_function(G0, D0, G_D(G0, D0, o_G, o_D), None, {"itr": 0}, cfg1, lambda: next(it),
                                                 embedded_optimizers=False, device="cpu", batch_size=3, grad_sync=None)
        single(torch.cat([e["x"] for e in everything]), None, torch.cat([e["f"] for e in everything]))
        ws = torch.cat([p.detach().flatten() for p in list(G0.parameters()) + list(D0.parameters())])
        wd = torch.cat([p.detach().flatten() for p in list(G.parameters()) + list(D.parameters())])
        assert torch.allclose(ws, wd, atol=1e-6, rtol=1e-5), f"2-rank step != single-process step: {(ws - wd).abs().max()}"
    assert sync.calls == 2, f"expected exactly 2 all-reduces per step, got {sync.calls}"
    w1 = torch.cat([p.detach().flatten() for p in list(G.parameters()) + list(D.parameters())])
    gathered = [torch.zeros_like(w1) for _ in range(world)]
    dist.all_gather(gathered, w1)
    assert torch.allclose(gathered[0], gathered[1], atol=0, rtol=0), "replicas diverged after a synchronised step"
    assert not torch.equal(w0, w1)
    # the flat buffers really alias .grad
    p0 = next(D.parameters())
    assert p0.grad.data_ptr() == sync.flat["D"].flat.data_ptr()
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


def test_two_rank_gloo_step():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"


def test_flat_grads_alias_and_mean():
    from ic_gan_b200.dist import FlatGrads
    m = nn.Linear(3, 2)
    fg = FlatGrads(m.parameters())
    m(torch.ones(1, 3)).sum().backward()
    assert fg.flat.abs().sum() > 0 and m.weight.grad.data_ptr() == fg.flat.data_ptr()
    m.zero_grad(set_to_none=False)
    assert fg.flat.abs().sum() == 0


def test_ema_covers_every_state_entry_and_start_itr():
    """utils.py:1039-1067: EMA over ALL state_dict entries (buffers such as BN statistics and SN `u` too); plain copy
    (decay 0) while itr < start_itr."""
    from ic_gan_b200.biggan.train_fns import ema

    class M(nn.Module):
        def __init__(self, v):
            super().__init__()
            self.w = nn.Parameter(torch.full((3,), float(v)))
            self.register_buffer("stored_mean", torch.full((2,), float(v) * 10))
            self.register_buffer("u0", torch.full((1, 4), float(v) * 100))

    src, tgt = M(1.0), M(-5.0)
    e = ema(src, tgt, decay=0.9, start_itr=5)
    assert all(torch.equal(tgt.state_dict()[k], src.state_dict()[k]) for k in src.state_dict())  # initial copy
    with torch.no_grad():
        src.w.fill_(2.0); src.stored_mean.fill_(20.0); src.u0.fill_(200.0)
    e.update(itr=1)  # before start_itr: decay 0 -> target := source
    assert torch.equal(tgt.w.data, src.w.data) and torch.equal(tgt.stored_mean, src.stored_mean) and torch.equal(tgt.u0, src.u0)
    with torch.no_grad():
        src.w.fill_(3.0); src.stored_mean.fill_(30.0); src.u0.fill_(300.0)
    e.update(itr=7)
    assert torch.allclose(tgt.w.data, torch.full((3,), 0.9 * 2.0 + 0.1 * 3.0))
    assert torch.allclose(tgt.stored_mean, torch.full((2,), 0.9 * 20.0 + 0.1 * 30.0))
    assert torch.allclose(tgt.u0, torch.full((1, 4), 0.9 * 200.0 + 0.1 * 300.0))
    # the reference tests `if itr and itr < start_itr` (utils.py:1058): itr == 0 is falsy and uses the REAL decay
    # (pinned against the live reference by tests/golden/biggan_step_cc32.*)
    before = tgt.w.data.clone()
    with torch.no_grad():
        src.w.fill_(4.0)
    e.update(itr=0)
    assert torch.allclose(tgt.w.data, 0.9 * before + 0.1 * 4.0)
