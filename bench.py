#!/usr/bin/env python
"""bench.py — images/sec of one IC-GAN BigGAN G+D training step (BASELINE.json metric) on N B200 GPUs of one node.

    python bench.py --gpus 1 --steps K --warmup W                      (N=1)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus N --steps K --warmup W     (the reference algorithm on the host CPU cores)

A "step" is one GAN_training_function.train call (BigGAN_PyTorch/train_fns.py:40-191 semantics: D update on B fake +
B real, G update on B fake, Adam x2, EMA) over a per-GPU batch of synthetic images and instance features, realised as
micro-batches with gradient accumulation (BN statistics are per micro-batch, as in the reference).  Weak scaling: every
rank processes the same per-GPU batch; gradients are averaged with one NCCL all-reduce for D and one for G per step.
One JSON line is printed by rank 0 (contract in the task statement; extra keys: roofline, cpu_baseline, e2e, clocks).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# forward GFLOP per image (conv + linear + attention bmm, 2*MAC) measured on the reference modules: SURVEY.md §8(d)
WORKLOADS = {
    "cc256": dict(name="cc-IC-GAN BigGAN 256x256 (ch96, attn@64, class+instance cond)", resolution=256, ch=96,
                  attn="64", class_cond=True, G_f=146.43, D_f=74.67, per_gpu_batch=256, micro_batch=128),
    "ic128": dict(name="IC-GAN BigGAN 128x128 (ch96, attn@64, instance cond)", resolution=128, ch=96, attn="64",
                  class_cond=False, G_f=42.26, D_f=21.68, per_gpu_batch=256, micro_batch=256),
    "ic64": dict(name="IC-GAN BigGAN 64x64 (ch64, attn@32, instance cond)", resolution=64, ch=64, attn="32",
                 class_cond=False, G_f=14.52, D_f=2.23, per_gpu_batch=256, micro_batch=128),
}
METRIC = "images/sec G+D step, IC-GAN BigGAN"
JSON_OUT = [None]  # where the one JSON line goes (the original stdout when fd 1 has been pointed at stderr)


T0 = time.perf_counter()


def note(msg):
    """Progress on stderr (stdout carries exactly one JSON line)."""
    print(f"[bench {time.perf_counter() - T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def cpu_leg(flag, timeout_s=240):
    """Run `bench.py <flag>` (a CPU baseline leg) in a subprocess with bounded threads and wall time; returns its JSON
    or None.  A CPU leg that overruns must never take the GPU numbers down with it."""
    threads = min(32, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), MKL_NUM_THREADS=str(threads), CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__)] + flag, env=env, capture_output=True, text=True,
                             timeout=timeout_s)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:  # timeout, crash, no output
        note(f"cpu leg {flag} unavailable: {type(e).__name__}")
        return None


def nccl_to_stderr():
    """NCCL's communicator lines ("... nranks N ...") are evidence the driver reads: make sure they are emitted (INFO
    unless the caller asked for more) and send them to stderr -- fd 1 is pointed at stderr for the whole run and the one
    JSON line is written to the saved original stdout."""
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
        os.environ["NCCL_DEBUG"] = os.environ.get("ICGAN_NCCL_DEBUG", "INFO")
    os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
    sys.stdout.flush()
    JSON_OUT[0] = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = JSON_OUT[0] or sys.stdout
    print(json.dumps(line), file=out, flush=True)


def model_kwargs(w):
    return dict(G_ch=w["ch"], D_ch=w["ch"], dim_z=120, resolution=w["resolution"], G_attn=w["attn"], D_attn=w["attn"],
                n_classes=1000, G_shared=True, shared_dim=128, hier=True, BN_eps=1e-5, SN_eps=1e-6,
                class_cond=w["class_cond"], instance_cond=True, G_shared_feat=True, shared_dim_feat=512,
                G_init="ortho", D_init="ortho")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------- reference arm
def oracle_step_rate(w, batch, steps, warmup, threads):
    """images/s of the oracle's train_step (CPU restatement of the reference algorithm) at micro-batch `batch`."""
    from oracle import biggan_oracle as O
    torch.set_num_threads(threads)
    cfg = O.BigGANConfig(resolution=w["resolution"], G_ch=w["ch"], D_ch=w["ch"], G_attn=w["attn"], D_attn=w["attn"],
                         class_cond=w["class_cond"], instance_cond=True)
    gs, ds = O.state_shapes(cfg)
    st = O.make_step_state(O.synth_state_dict(gs, 1), O.synth_state_dict(ds, 2), G_lr=4e-5, D_lr=1e-4, ema=True)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(batch, 3, cfg.resolution, cfg.resolution, generator=g) * 2 - 1
    feats = torch.nn.functional.normalize(torch.randn(batch, 2048, generator=g), dim=1)
    y = torch.randint(0, 1000, (batch,), generator=g) if cfg.class_cond else None

    def sample_cond():
        z = torch.randn(batch, cfg.eff_dim_z, generator=g)
        f = torch.nn.functional.normalize(torch.randn(batch, 2048, generator=g), dim=1)
        lab = torch.randint(0, 1000, (batch,), generator=g) if cfg.class_cond else None
        return z, lab, f
    for _ in range(warmup):
        O.train_step(st, cfg, x, y, feats, sample_cond, batch)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(st, cfg, x, y, feats, sample_cond, batch)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, dt


ORACLE_THREADS = 16  # per worker process: a batch-1 256x256 step does not scale past ~16-32 threads (measured: 128
#                      threads in one process = 89.6 s per step on the GPU box's host, 16 threads = ~9 s)


def oracle_pool_rate(workload, batch, steps, warmup):
    """The oracle on ALL host cores: cpu_count // ORACLE_THREADS independent worker processes (independent samples, as the
    reference's data parallelism would place them), each timing `steps` oracle train_steps at micro-batch `batch`.
    Returns (aggregate images/s, mean seconds per step, workers, threads per worker)."""
    import subprocess
    # the cores this process may actually run on (cgroup cpusets make cpu_count() lie), each worker pinned to its own
    # disjoint set so that two hosts with the same core count time the same thing (round 1: 5x host-to-host spread)
    try:
        avail = sorted(os.sched_getaffinity(0))
    except AttributeError:
        avail = list(range(os.cpu_count() or 1))
    cores = len(avail)
    threads = min(ORACLE_THREADS, cores)
    workers = max(1, cores // threads)
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), MKL_NUM_THREADS=str(threads), CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), "--oracle-worker", "--workload", workload, "--steps", str(steps),
           "--warmup", str(warmup), "--cpu-batch", str(batch), "--threads", str(threads)]
    procs = []
    for i in range(workers):
        mine = avail[i * threads:(i + 1) * threads]
        procs.append(subprocess.Popen(cmd + ["--affinity", ",".join(map(str, mine))], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    dts = []
    for pr in procs:
        out, err = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"oracle worker failed: {err[-2000:]}")
        dts.append(json.loads(out.strip().splitlines()[-1])["dt"])
    return sum(batch / d for d in dts), sum(dts) / len(dts), workers, threads


def run_oracle_worker(args, w):
    if args.affinity:
        try:
            os.sched_setaffinity(0, {int(c) for c in args.affinity.split(",")})
        except (AttributeError, OSError):
            pass
    ips, dt = oracle_step_rate(w, args.cpu_batch, args.steps, args.warmup, args.threads)
    print(json.dumps({"dt": dt, "ips": ips}), flush=True)


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 1
    ips, dt, workers, threads = oracle_pool_rate(args.workload, batch, args.steps, args.warmup)
    sample = (f"oracle train_step (CPU restatement of train_fns.py:40-191), fp32, {workers} worker processes x {threads} "
              f"threads, each stepping its own micro-batch of {batch}")
    threads = workers * threads
    line = {"impl": "reference", "metric": f"{METRIC} {w['resolution']}x{w['resolution']}", "value": ips,
            "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": w["name"], "micro_batch": batch, "host_threads": threads},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------- kNN build (config 5)
KNN_D, KNN_K = 2048, 50


def knn_cpu_rate(n_sample):
    """queries/s of the oracle (exact float64 restatement of _obtain_nns: BLAS X.X^T blocks + argpartition + exact re-rank)
    on an n_sample-row database with all host threads; the caller extrapolates to N by the N^2 work law."""
    import numpy as np
    from oracle import knn_oracle as K
    rng = np.random.default_rng(6)
    x = K.normalize_features(rng.standard_normal((n_sample, KNN_D)))
    t0 = time.perf_counter()
    K.obtain_nns(x, KNN_K)
    return n_sample / (time.perf_counter() - t0)


def run_knn(args):
    """BASELINE config 5: brute-force k-NN build over N x 2048 float32 instance features, k = 50.  A "step" is one complete
    build of this rank's query-row shard (split-bf16 tensor-core distance sweep with fused top-64 selection, exact float64
    re-rank + certification, device fallback for uncertified rows); database replicated, query rows sharded, one
    all_gather of the results at the end.  value = queries/s of the whole job."""
    import numpy as np
    import torch.distributed as dist
    from ic_gan_b200 import _lib, knn
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    N = args.knn_n
    if args.impl == "reference":
        if rank:
            return
        ns = min(N, 16384)
        qps = knn_cpu_rate(ns)
        full = qps * ns / N
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
        sample = (f"oracle obtain_nns (float64 BLAS X.X^T blocks + argpartition + exact re-rank) on a {ns}-row database, "
                  f"{qps:.1f} queries/s there, extrapolated x {ns}/{N} (work per query is proportional to N)")
        emit({"impl": "reference", "metric": f"queries/sec kNN build (N={N}, d={KNN_D}, k={KNN_K})", "value": full,
              "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": N / full * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": "f64", "data": "synthetic", "config": {"workload": f"kNN build N={N}", "host_threads": cores},
              "cpu_baseline": {"value": full, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
              "e2e": {"value": full, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
              "gpu_launches": 0})
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        nccl_to_stderr()
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    gen = torch.Generator(device=dev).manual_seed(6)  # same table on every rank (replicated database)
    X = torch.empty(N, KNN_D, device=dev, dtype=torch.float32)
    for c0 in range(0, N, 65536):  # randn -> float64 normalise -> float32 (datasets_common.py:422-428), chunked
        blk = torch.randn(min(65536, N - c0), KNN_D, device=dev, generator=gen, dtype=torch.float64)
        X[c0:c0 + blk.shape[0]] = (blk / blk.norm(dim=1, keepdim=True)).float()
    q0, q1 = rank * N // world, (rank + 1) * N // world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    coarse_ev = []

    def build(rows, gather=True):
        res = knn.obtain_nns(X, KNN_K, rows=rows, timing=coarse_ev)
        if world > 1 and gather:  # [N/W, 50] int64 + [N/W] float64 per rank, padded to the largest shard
            m = (N + world - 1) // world
            nn_pad = torch.full((m, KNN_K), -1, device=dev, dtype=torch.int64)
            rd_pad = torch.zeros(m, device=dev, dtype=torch.float64)
            nn_pad[:rows[1] - rows[0]] = res.sample_nns
            rd_pad[:rows[1] - rows[0]] = res.sample_nns_radius
            all_nn = torch.empty(world * m, KNN_K, device=dev, dtype=torch.int64)
            all_rd = torch.empty(world * m, device=dev, dtype=torch.float64)
            dist.all_gather_into_tensor(all_nn, nn_pad)
            dist.all_gather_into_tensor(all_rd, rd_pad)
        return res

    for _ in range(max(1, args.warmup)):  # warm-up on a 1/16 slice of the shard (same kernels, same table)
        build((q0, q0 + max(256, (q1 - q0) // 16)), gather=True)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    del coarse_ev[:]
    launches0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = build((q0, q1))
    e1.record()
    barrier()
    launches = _lib.LAUNCHES - launches0
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    coarse_ms = sum(a.elapsed_time(b) for a, b in coarse_ev) / args.steps
    # spot check inside the bench: rows of this shard against a float64 brute force on the device
    rows = torch.randint(q0, q1, (16,), generator=torch.Generator().manual_seed(rank)).tolist()
    bad = 0
    for r in rows:
        d2 = torch.empty(N, device=dev, dtype=torch.float64)
        for c0 in range(0, N, 131072):
            diff = X[c0:c0 + 131072].double() - X[r].double()[None, :]
            d2[c0:c0 + 131072] = (diff * diff).sum(1)
        m = KNN_K + 17
        vals, idx = torch.topk(d2, m, largest=False)
        order = np.lexsort((idx.cpu().numpy(), vals.cpu().numpy()))
        cand = idx.cpu().numpy()[order][:KNN_K + 1]
        keep = cand[cand != r][:KNN_K]
        bad += int(not np.array_equal(keep, res.sample_nns[r - q0].cpu().numpy()))
    # ---- e2e: host features in (pinned), neighbour tables out to the host
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty(N, KNN_D, dtype=torch.float32).pin_memory()
        Xh.copy_(X)
        out_nn = torch.empty(q1 - q0, KNN_K, dtype=torch.int64).pin_memory()
        out_rd = torch.empty(q1 - q0, dtype=torch.float64).pin_memory()
        barrier()
        e0.record()
        for _ in range(args.steps):
            X.copy_(Xh, non_blocking=True)
            r2 = knn.obtain_nns(X, KNN_K, rows=(q0, q1))
            out_nn.copy_(r2.sample_nns, non_blocking=True)
            out_rd.copy_(r2.sample_nns_radius, non_blocking=True)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": N / (float(t.item()) * 1e-3), "unit": "queries/s", "ms_per_step": float(t.item()),
               "h2d_bytes_per_step": N * KNN_D * 4, "d2h_bytes_per_step": (q1 - q0) * (KNN_K * 8 + 8)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    nq = q1 - q0
    alg = 2.0 * nq * N * KNN_D  # 2*N^2*d for the whole job; this rank's share
    ach = alg / (coarse_ms * 1e-3) * 1e-12
    line = {"metric": f"queries/sec kNN build (N={N}, d={KNN_D}, k={KNN_K})", "value": N / (ms * 1e-3),
            "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16x3 (split) + f64 re-rank",
            "data": "synthetic",
            "config": {"workload": f"kNN conditioning build, N={N} x {KNN_D} float32, k={KNN_K}, bit-exact indices",
                       "query_rows_per_gpu": nq, "parallelism": f"query-sharded x{world}, database replicated",
                       "l2": "inputs larger than L2", "seconds_per_build": ms * 1e-3,
                       "uncertified_rows": res.stats["uncertified_rows"], "spot_check_rows": len(rows),
                       "spot_check_mismatches": bad,
                       "peak_hbm_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1)},
            "roofline": {"kernel": "knn_coarse_kernel (split-bf16 distance sweep + fused top-64 selection)",
                         "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                         "frac": ach / pk["tf_sustained"], "traffic": None,
                         "note": "achieved = ALGORITHMIC 2*nq*N*d / kernel time; the kernel executes 3 bf16 products per "
                                 f"algorithmic one, i.e. {3 * ach:.0f} TFLOP/s of tensor-pipe work",
                         "kernel_ms_per_step": coarse_ms, "share_of_step": coarse_ms / ms,
                         "peak_source": pk["source"] + ", sustained"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
    if world == 1 and not args.no_cpu_baseline:
        ns = 16384
        qps = knn_cpu_rate(ns)
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
        line["cpu_baseline"] = {"value": qps * ns / N, "unit": "queries/s", "cores": cores, "kind": "port",
                                "sample": f"oracle obtain_nns on a {ns}-row database ({qps:.1f} queries/s), extrapolated "
                                          f"x {ns}/{N} (work per query proportional to N)"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------- StyleGAN2 (config 4)
SG = dict(z_dim=512, w_dim=512, h_dim=2048, res=256, channel_base=16384, channel_max=512, map_layers=2, num_fp16_res=4,
          conv_clamp=256, mbstd=4, G_f=29.79, D_f=30.75, G_reg=4, D_reg=16, lr=0.0025)
# conv GFLOP per image and iteration: Gmain 3 G_f + 2 D_f, Dmain G_f + 6 D_f, + lazy regularisers
SG_STEP_GF = 4 * SG["G_f"] + 8 * SG["D_f"] + (2.5 * SG["G_f"]) / 4 + (5 * SG["D_f"]) / 16


def sg_build(dev):
    from ic_gan_b200.stylegan2 import networks as N
    G = N.Generator(z_dim=SG["z_dim"], c_dim=0, h_dim=SG["h_dim"], w_dim=SG["w_dim"], img_resolution=SG["res"], img_channels=3,
                    mapping_kwargs=dict(num_layers=SG["map_layers"]),
                    synthesis_kwargs=dict(channel_base=SG["channel_base"], channel_max=SG["channel_max"],
                                          num_fp16_res=SG["num_fp16_res"], conv_clamp=SG["conv_clamp"]))
    D = N.Discriminator(c_dim=0, h_dim=SG["h_dim"], img_resolution=SG["res"], img_channels=3, channel_base=SG["channel_base"],
                        channel_max=SG["channel_max"], num_fp16_res=SG["num_fp16_res"], conv_clamp=SG["conv_clamp"],
                        epilogue_kwargs=dict(mbstd_group_size=SG["mbstd"]))
    return G.to(dev), D.to(dev)


def sg_cpu_rate(batch=2):
    """images/s of the oracle's StyleGAN2 iteration (CPU restatement of loss.py:85-194 over networks.py) at 256x256:
    one Gmain + one Dmain + Greg/4 + Dreg/16, each phase timed once at `batch`, all host threads."""
    from oracle import stylegan_nets_oracle as O
    G, D = sg_build("cpu")
    cfg = O.StyleGANConfig(z_dim=SG["z_dim"], h_dim=SG["h_dim"], w_dim=SG["w_dim"], img_resolution=SG["res"],
                           channel_base=SG["channel_base"], channel_max=SG["channel_max"], map_layers=SG["map_layers"],
                           d_map_layers=8, conv_clamp=float(SG["conv_clamp"]), mbstd_group_size=min(SG["mbstd"], batch))
    g_sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    d_sd = {k: v.detach().clone() for k, v in D.state_dict().items()}
    g = torch.Generator().manual_seed(0)
    z, h = torch.randn(batch, SG["z_dim"], generator=g), torch.randn(batch, SG["h_dim"], generator=g)
    x = torch.rand(batch, 3, SG["res"], SG["res"], generator=g) * 2 - 1
    times = {}
    for phase in ("Gmain", "Dmain", "Greg", "Dreg"):
        sd = g_sd if phase.startswith("G") else d_sd
        for k, v in sd.items():
            if v.dtype.is_floating_point and not k.endswith(("resample_filter", "noise_const", "w_avg")):
                v.requires_grad_(True)
        t0 = time.perf_counter()
        O.accumulate_gradients(phase, g_sd, d_sd, cfg, x, None, h, z, None, h, 1.0, pl_mean=torch.tensor(0.0))
        times[phase] = time.perf_counter() - t0
        for v in sd.values():
            v.requires_grad_(False)
            v.grad = None
    it = times["Gmain"] + times["Dmain"] + times["Greg"] / SG["G_reg"] + times["Dreg"] / SG["D_reg"]
    return batch / it, it, times


def run_sg256(args):
    """BASELINE config 4: one StyleGAN2-ADA IC-GAN training iteration at 256x256, 64 images per GPU (training_loop.py:
    395-535 schedule: Gmain and Dmain every iteration, Greg every 4th, Dreg every 16th, lazy-regularisation-scaled Adam,
    G_ema) on the B200 networks / loss.  --steps should be a multiple of 16 so that every regulariser is averaged in."""
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "images/sec G+D iteration, IC-GAN StyleGAN2-ADA 256x256"
    if args.sg_cpu_worker:
        torch.set_num_threads(int(os.environ.get("OMP_NUM_THREADS", "16")))
        ips, it, times = sg_cpu_rate(2)
        print(json.dumps({"ips": ips, "it": it, "times": times, "threads": torch.get_num_threads()}), flush=True)
        return
    if args.impl == "reference":
        if rank:
            return
        res = cpu_leg(["--workload", "sg256", "--sg-cpu-worker"], 400)
        if res is None:
            emit({"impl": "reference", "unavailable": "oracle StyleGAN2 iteration did not finish within 400 s on this host"})
            return
        ips, it, times, cores = res["ips"], res["it"], res["times"], res["threads"]
        sample = ("oracle iteration (CPU restatement of training/loss.py:85-194 over training/networks.py), fp32, batch 2: "
                  + ", ".join(f"{k} {v:.1f} s" for k, v in times.items()) + f"; iteration = Gmain + Dmain + Greg/4 + Dreg/16 = {it:.1f} s")
        emit({"impl": "reference", "metric": metric, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
              "warmup": args.warmup, "ms_per_step": it * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "f32", "data": "synthetic", "config": {"workload": "IC-GAN StyleGAN2-ADA 256x256", "batch": 2,
                                                               "host_threads": cores},
              "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
              "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return
    from ic_gan_b200 import _lib, ops
    from ic_gan_b200.optim import FusedAdamEMA
    from ic_gan_b200.stylegan2 import loss as sg_loss
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        nccl_to_stderr()
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    torch.manual_seed(4321)
    Bg = args.per_gpu_batch or 64
    G, D = sg_build(dev)
    G_ema, _ = sg_build(dev)
    G_ema.load_state_dict(G.state_dict())
    G.train(); D.train(); G_ema.eval()
    G.requires_grad_(False); D.requires_grad_(False)  # training_loop.py:226-238: each phase switches its own module on
    if world > 1:
        for net in (G, D):
            for t in list(net.parameters()) + list(net.buffers()):
                dist.broadcast(t.data, src=0)

    def lazy(interval):  # training_loop.py:332-339: Adam hyper-parameters rescaled for lazy regularisation
        r = interval / (interval + 1)
        return dict(lr=SG["lr"] * r, betas=(0.0 ** r, 0.99 ** r), eps=1e-8)
    opt_G = FusedAdamEMA(G.parameters(), **lazy(SG["G_reg"]))
    opt_D = FusedAdamEMA(D.parameters(), **lazy(SG["D_reg"]))
    loss = sg_loss.StyleGAN2Loss(dev, G.mapping, G.synthesis, D, augment_pipe=None, style_mixing_prob=0.9,
                                 r1_gamma=0.0002 * SG["res"] ** 2 / (Bg * world), pl_batch_shrink=2, pl_decay=0.01, pl_weight=2)
    ema_beta = 0.5 ** (Bg * world / max(Bg * world * 10 / 32 * 1000, 1e-8))  # training_loop.py:528-531, cfg auto ema
    eager_loss = loss
    if not args.no_graphs and not args.ncu:
        from ic_gan_b200.stylegan2.graphs import GraphedLoss
        loss = GraphedLoss(eager_loss, {"G": G, "D": D})
    g_params, e_params = list(G.parameters()), list(G_ema.parameters())
    g_bufs, e_bufs = list(G.buffers()), list(G_ema.buffers())
    gen = torch.Generator(device=dev).manual_seed(300 + rank)
    real_dev = torch.randint(0, 256, (Bg, 3, SG["res"], SG["res"]), device=dev, generator=gen, dtype=torch.uint8)
    h_dev = torch.randn(Bg, SG["h_dim"], device=dev, generator=gen)
    c0 = torch.zeros(Bg, 0, device=dev)
    it_count = [0]

    def iteration(real_u8, real_h):
        it = it_count[0]
        it_count[0] += 1
        real_img = real_u8.to(torch.float32) / 127.5 - 1  # training_loop.py:436
        gen_z = torch.randn(Bg, SG["z_dim"], device=dev)
        gen_h = real_h[torch.randperm(Bg, device=dev)]
        phases = [("Gmain", G, opt_G, 1), ("Greg", G, opt_G, SG["G_reg"]), ("Dmain", D, opt_D, 1), ("Dreg", D, opt_D, SG["D_reg"])]
        for name, module, opt, interval in phases:
            if it % interval:
                continue
            opt.zero_grad()
            module.requires_grad_(True)
            loss.accumulate_gradients(phase=name, real_img=real_img, real_c=c0, real_h=real_h, gen_z=gen_z, gen_c=c0,
                                      gen_h=gen_h, sync=True, gain=interval)
            module.requires_grad_(False)
            if world > 1:
                dist.all_reduce(opt.flat_g)
                opt.set_grad_scale(1.0 / world)
            torch.nan_to_num(opt.flat_g, nan=0, posinf=1e5, neginf=-1e5, out=opt.flat_g)  # training_loop.py:516-521
            opt.step()
        with torch.no_grad():  # G_ema (training_loop.py:527-535)
            torch._foreach_lerp_(e_params, g_params, 1.0 - ema_beta)
            torch._foreach_copy_(e_bufs, g_bufs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.ncu:
        args.no_e2e = args.no_cpu_baseline = True
    if args.torch_profile:
        from torch.profiler import ProfilerActivity, profile
        loss = eager_loss
        for _ in range(3):
            iteration(real_dev, h_dev)
        it_count[0] = 0
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof_t:
            for _ in range(16):
                iteration(real_dev, h_dev)
            torch.cuda.synchronize()
        rows = sorted(prof_t.key_averages(), key=lambda e: -e.device_time_total)
        tot = sum(e.device_time_total for e in rows)
        with open(args.torch_profile, "w") as f:
            f.write(f"# torch.profiler, 16 eager sg256 iterations (Gmain x16, Dmain x16, Greg x4, Dreg x1), {tot / 16e3:.2f} ms of "
                    f"kernel time per iteration\n# kernel, launches, total ms, share\n")
            for e in rows[:70]:
                f.write(f"{e.key[:110]:110s} {e.count:7d} {e.device_time_total / 1e3:10.2f} {100 * e.device_time_total / tot:6.2f}%\n")
        note(f"profile written: {tot / 16e3:.2f} ms kernel time per iteration")
        return
    note("sg256: networks built, warm-up")
    graphed = loss is not eager_loss
    prof_eager = None
    if graphed:
        # per-launch CUDA events cannot live inside a graph: the tensor-core kernel statistics of the roofline come from
        # one instrumented EAGER cycle of 16 iterations (same kernels, same shapes), the throughput from graph replays
        loss_g, loss = loss, eager_loss
        for i in range(2):
            iteration(real_dev, h_dev)
        it_count[0] = 0
        torch.cuda.synchronize()
        ops.PROFILE = []
        for i in range(16):
            iteration(real_dev, h_dev)
        torch.cuda.synchronize()
        prof_eager, ops.PROFILE = ops.PROFILE, None
        loss = loss_g
        it_count[0] = 0
        note("sg256: eager instrumented cycle done; capturing graphs")
        iteration(real_dev, h_dev)  # it = 0 runs all four phases: captures them
        torch.cuda.synchronize()
        note("sg256: graphs captured")
    for i in range(args.warmup if args.ncu else max(args.warmup, 3)):
        iteration(real_dev, h_dev)
        torch.cuda.synchronize()
        note(f"sg256: warm-up iteration {i} done")
    it_count[0] = 0
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.PROFILE = [] if (rank == 0 and not graphed) else None
    launches0 = _lib.LAUNCHES + (loss.replayed_launches if graphed else 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        iteration(real_dev, h_dev)
    e1.record()
    barrier()
    launches = _lib.LAUNCHES + (loss.replayed_launches if graphed else 0) - launches0
    prof, ops.PROFILE = ops.PROFILE, None
    prof_steps = args.steps
    if graphed:
        prof, prof_steps = (prof_eager or []), 16
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    note(f"sg256: {ms:.1f} ms per iteration (device-resident inputs)")
    clocks = sampler.stop() if rank == 0 else None
    e2e = None
    if not args.no_e2e:
        cpu_gen = torch.Generator().manual_seed(400 + rank)
        real_host = torch.randint(0, 256, (Bg, 3, SG["res"], SG["res"]), generator=cpu_gen, dtype=torch.uint8).pin_memory()
        h_host = torch.randn(Bg, SG["h_dim"], generator=cpu_gen).pin_memory()
        it_count[0] = 0

        def step_host():
            iteration(real_host.to(dev, non_blocking=True), h_host.to(dev, non_blocking=True))
            return float(loss.stats["Loss/G/loss"]) + float(loss.stats["Loss/D/loss_gen"])  # host read of the losses
        step_host()
        it_count[0] = 0
        barrier()
        e0.record()
        for _ in range(args.steps):
            step_host()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": Bg * world / (float(t.item()) * 1e-3), "unit": "images/s", "ms_per_step": float(t.item()),
               "h2d_bytes_per_step": real_host.numel() + h_host.numel() * 4, "d2h_bytes_per_step": 8}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    per = {}
    for name, flops, a, b, _ in prof:
        d = per.setdefault(name, [0.0, 0.0, 0])
        d[0] += flops; d[1] += a.elapsed_time(b) * 1e-3; d[2] += 1
    kinfo = {k: {"launches": v[2], "avg_ms": v[1] / v[2] * 1e3, "tflops": v[0] / v[1] * 1e-12,
                 "share_of_step": v[1] / (ms * 1e-3 * prof_steps)} for k, v in per.items()}
    dom = max(per, key=lambda k: per[k][1]) if per else None
    roofline = None
    if dom:
        ach = per[dom][0] / per[dom][1] * 1e-12
        roofline = {"kernel": {"sg2_conv": "tc_conv_halo_kernel + tc_conv_kernel (stride 1/2, transposed phases; forward and dgrad)",
                               "sg2_wgrad": "tc_wgrad_kernel + tc_wgrad_halo_kernel"}[dom], "bound": "tensor", "achieved": ach,
                    "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"], "traffic": None,
                    "peak_source": pk["source"] + ", sustained", "flops_per_launch": per[dom][0] / per[dom][2],
                    "kernels": kinfo}
    value = Bg * world / (ms * 1e-3)
    step_tf = SG_STEP_GF * 1e9 * value / world * 1e-12
    line = {"metric": metric, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "IC-GAN StyleGAN2-ADA 256x256 (cfg auto: channel_base 16384, map 2, num_fp16_res 4 -> bf16, "
                                   "conv_clamp 256, mbstd 4, h_dim 2048; fp32 blocks on split-bf16 tensor-core passes)",
                       "per_gpu_batch": Bg, "global_batch": Bg * world, "parallelism": f"dp{world}",
                       "schedule": "Gmain + Dmain every iteration, Greg every 4th, Dreg every 16th (lazy regularisation), "
                                   "fused Adam, G_ema", "l2": "inputs larger than L2",
                       "launch": ("each phase replayed from a CUDA graph; roofline kernel statistics from an instrumented "
                                  "eager cycle of the same 16 iterations") if graphed else "eager launches",
                       "step_gflop_per_image": SG_STEP_GF, "peak_hbm_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1)},
            "roofline": roofline,
            "step_roofline": {"achieved_tflops_per_gpu": step_tf, "frac_of_sustained_peak": step_tf / pk["tf_sustained"]},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
    if world == 1 and not args.no_cpu_baseline:
        note("cpu baseline leg")
        res = cpu_leg(["--workload", "sg256", "--sg-cpu-worker"], 240)
        if res is not None:
            line["cpu_baseline"] = {"value": res["ips"], "unit": "images/s", "cores": res["threads"], "kind": "port",
                                    "sample": "oracle iteration at batch 2 (Gmain + Dmain + Greg/4 + Dreg/16 = %.1f s)" % res["it"]}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cc256", choices=sorted(WORKLOADS) + ["knn", "sg256"])
    ap.add_argument("--knn-n", type=int, default=1281167, help="database rows of --workload knn")
    ap.add_argument("--per-gpu-batch", type=int, default=0)
    ap.add_argument("--micro-batch", type=int, default=0)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--oracle-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-batch", type=int, default=1, help=argparse.SUPPRESS)
    ap.add_argument("--threads", type=int, default=ORACLE_THREADS, help=argparse.SUPPRESS)
    ap.add_argument("--affinity", default="", help=argparse.SUPPRESS)
    ap.add_argument("--sg-cpu-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--torch-profile", default="", help="sg256: write a per-kernel device-time table (torch.profiler, 16 eager "
                                                       "iterations) to this path and exit; shares only, never a bench value")
    ap.add_argument("--no-graphs", action="store_true", help="issue every launch from Python instead of replaying "
                                                             "one CUDA graph per loss phase")
    ap.add_argument("--ncu", action="store_true", help="profiling run under ncu: short warm-up allowed, no e2e/cpu legs "
                                                       "(a number printed by such a run is never a bench value)")
    args = ap.parse_args()
    if args.workload == "knn":
        return run_knn(args)
    if args.workload == "sg256":
        return run_sg256(args)
    w = dict(WORKLOADS[args.workload])
    if args.per_gpu_batch:
        w["per_gpu_batch"] = args.per_gpu_batch
    if args.micro_batch:
        w["micro_batch"] = args.micro_batch
    if args.oracle_worker:
        return run_oracle_worker(args, w)
    if args.impl == "reference":
        return run_reference(args, w)
    if args.ncu:
        args.no_e2e = args.no_cpu_baseline = True
    elif args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    from ic_gan_b200 import _lib, ops
    from ic_gan_b200.biggan import G_D, Discriminator, Generator
    from ic_gan_b200.biggan import train_fns
    from ic_gan_b200.dist import GradSync
    from ic_gan_b200.optim import FusedAdamEMA

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device; there is no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        nccl_to_stderr()
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    cdt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    m, Bg = w["micro_batch"], w["per_gpu_batch"]
    assert Bg % m == 0, "per-GPU batch must be a multiple of the micro-batch"
    acc = Bg // m
    kw = model_kwargs(w)
    torch.manual_seed(1234)
    # random-init weights of the named architecture, drawn on the device (N(0, 0.02) = the reference's "N02" init style)
    G = Generator(no_optim=True, skip_init=True, compute_dtype=cdt, **kw).to(dev)
    D = Discriminator(embedded_optimizer=False, skip_init=True, compute_dtype=cdt, **kw).to(dev)
    G_ema = Generator(no_optim=True, skip_init=True, compute_dtype=cdt, **kw).to(dev)
    for net in (G, D):
        net.init = "N02"
        net.init_weights()
    # Adam + EMA: one kernel launch per network over flat parameter / gradient / moment / EMA buffers (ic_gan_b200/optim.py);
    # the data-parallel all-reduce runs on the same flat gradient buffers and its 1/world rides in the step kernel
    opt_G = FusedAdamEMA(G.parameters(), lr=4e-5, betas=(0.0, 0.999), eps=1e-6, ema_params=G_ema.parameters())
    opt_D = FusedAdamEMA(D.parameters(), lr=1e-4, betas=(0.0, 0.999), eps=1e-6)
    sync = GradSync({"G": G, "D": D}, world, optimizers={"G": opt_G, "D": opt_D})
    sync.broadcast_params()
    GD = G_D(G, D, opt_G, opt_D)
    ema = train_fns.ema(G, G_ema, 0.9999, 20000)
    ema.fuse_into(opt_G)
    state = {"itr": 0}
    config = {"toggle_grads": True, "num_D_steps": 1, "num_D_accumulations": acc, "num_G_accumulations": acc,
              "split_D": False, "ema": True, "D_ortho": 0.0, "G_ortho": 0.0, "DA": False, "DiffAugment": False}
    R, dim_z, cc = w["resolution"], G.dim_z, w["class_cond"]

    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    # device-resident real batch + a pool of pre-drawn conditionings (the `value` leg)
    x_dev = torch.rand(Bg, 3, R, R, device=dev, generator=gen) * 2 - 1
    f_dev = torch.nn.functional.normalize(torch.randn(Bg, 2048, device=dev, generator=gen), dim=1)
    y_dev = torch.randint(0, 1000, (Bg,), device=dev, generator=gen) if cc else None
    pool = 2 * acc
    z_pool = torch.randn(pool, m, dim_z, device=dev, generator=gen)
    fg_pool = torch.nn.functional.normalize(torch.randn(pool, m, 2048, device=dev, generator=gen), dim=2)
    yg_pool = torch.randint(0, 1000, (pool, m), device=dev, generator=gen) if cc else None
    cursor = [0]

    def sample_dev():
        i = cursor[0] % pool
        cursor[0] += 1
        return (z_pool[i], yg_pool[i], fg_pool[i]) if cc else (z_pool[i], fg_pool[i])

    # micro-steps replayed from CUDA graphs (ic_gan_b200/biggan/graphs.py) unless --no-graphs / profiling: ~4 400 launches
    # per step, ~9 % of the eager step is the device waiting for Python between the small ones
    graphed = not (args.no_graphs or args.ncu or args.torch_profile)
    train_dev = train_fns.GAN_training_function(G, D, GD, ema, state, config, sample_dev, embedded_optimizers=False,
                                                device=dev, batch_size=m, grad_sync=sync, lazy_losses=True,
                                                graphs=graphed)
    train_eager = train_dev if not graphed else train_fns.GAN_training_function(
        G, D, GD, ema, state, config, sample_dev, embedded_optimizers=False, device=dev, batch_size=m, grad_sync=sync,
        lazy_losses=True)

    def step_dev(fn=None):
        sync.broadcast_buffers()
        out = (fn or train_dev)(x_dev, y_dev, f_dev)
        state["itr"] += 1
        return out

    def replayed():
        return train_dev.graphs.replayed_launches if graphed else 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    G.train(); D.train()
    warm = args.warmup
    if graphed and warm > 0:
        try:  # the first step captures the two micro-step graphs
            step_dev()
            torch.cuda.synchronize()
        except Exception as exc:  # a capture problem must not take the measurement down: eager launches measure the same step
            print(f"[bench] CUDA-graph capture failed ({type(exc).__name__}: {exc}); continuing with eager launches",
                  file=sys.stderr, flush=True)
            graphed, train_dev = False, train_eager
        warm -= 1
    for _ in range(warm):
        step_dev()
    barrier()
    if args.torch_profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof_t:
            for _ in range(2):
                step_dev()
            torch.cuda.synchronize()
        rows = sorted(prof_t.key_averages(), key=lambda e: -e.device_time_total)
        tot = sum(e.device_time_total for e in rows)
        with open(args.torch_profile, "w") as f:
            f.write(f"# torch.profiler, 2 steps of {w['name']} (per-GPU batch {Bg}, micro-batch {m}): {tot / 2e3:.2f} ms of kernel "
                    f"time per step\n# kernel, launches, total ms, share\n")
            for e in rows[:70]:
                f.write(f"{e.key[:110]:110s} {e.count:7d} {e.device_time_total / 1e3:10.2f} {100 * e.device_time_total / tot:6.2f}%\n")
        return
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.PROFILE = [] if (rank == 0 and not graphed) else None
    launches0 = _lib.LAUNCHES + replayed()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    barrier()
    launches = _lib.LAUNCHES + replayed() - launches0
    prof = ops.PROFILE
    ops.PROFILE = None
    ms = e0.elapsed_time(e1) / args.steps
    prof_steps, prof_ms = args.steps, ms
    if graphed:
        # per-kernel CUDA events cannot be recorded inside a graph: the tensor-core launches are timed in an eager,
        # instrumented pass over the same step right after the timed region (same kernels, shapes and data).  Every rank
        # runs it (the steps contain collectives); only rank 0 records events.
        prof_steps = min(args.steps, 3)
        step_dev(train_eager)  # the eager path's allocator pool is cold after the replays: one untimed step first
        ops.PROFILE = [] if rank == 0 else None
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(prof_steps):
            step_dev(train_eager)
        p1.record()
        barrier()
        prof, ops.PROFILE = ops.PROFILE, None
        prof_ms = p0.elapsed_time(p1) / prof_steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    value = Bg * world / (ms * 1e-3)

    # ---------------- e2e: the same step through the public call with HOST inputs (pinned) and a host read of the losses
    e2e = None
    if not args.no_e2e:
        cpu_gen = torch.Generator().manual_seed(200 + rank)
        x_host = (torch.rand(Bg, 3, R, R, generator=cpu_gen) * 2 - 1).pin_memory()
        f_host = torch.nn.functional.normalize(torch.randn(Bg, 2048, generator=cpu_gen), dim=1).pin_memory()
        y_host = torch.randint(0, 1000, (Bg,), generator=cpu_gen).pin_memory() if cc else None
        zbuf = torch.empty(m, dim_z).pin_memory()
        fg_host = torch.nn.functional.normalize(torch.randn(64, m, 2048, generator=cpu_gen), dim=2).pin_memory()
        yg_host = torch.randint(0, 1000, (64, m), generator=cpu_gen).pin_memory() if cc else None
        h2d = [0]

        def sample_host():  # host-side draw like data_utils.sample_conditioning_values: z ~ N(0,1) in place, rows of feats
            i = cursor[0] % 64
            cursor[0] += 1
            zbuf.normal_(generator=cpu_gen)
            h2d[0] += zbuf.numel() * 4 + fg_host[i].numel() * 4 + (yg_host[i].numel() * 8 if cc else 0)
            return (zbuf, yg_host[i], fg_host[i]) if cc else (zbuf, fg_host[i])

        train_host = train_fns.GAN_training_function(G, D, GD, ema, state, config, sample_host,
                                                     embedded_optimizers=False, device=dev, batch_size=m,
                                                     grad_sync=sync, graphs=train_dev.graphs if graphed else False)

        def step_host():
            sync.broadcast_buffers()
            xd = x_host.to(dev, non_blocking=True)
            fd = f_host.to(dev, non_blocking=True)
            yd = y_host.to(dev, non_blocking=True) if cc else None
            out = train_host(xd, yd, fd)
            state["itr"] += 1
            return out["G_loss"] + out["D_loss_real"] + out["D_loss_fake"]  # Python floats: train() read them back (12 B)

        step_host()
        barrier()
        h2d[0] = 0
        e0.record()
        for _ in range(args.steps):
            step_host()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        per_step_in = x_host.numel() * 4 + f_host.numel() * 4 + (y_host.numel() * 8 if cc else 0) + h2d[0] // args.steps
        e2e = {"value": Bg * world / (e2e_ms * 1e-3), "unit": "images/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(per_step_in), "d2h_bytes_per_step": 12}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel from the events recorded INSIDE the timed region
    pk = peaks()
    per_kernel = {}
    per_shape = {}
    for name, flops, a, b, shape in prof:
        t = a.elapsed_time(b) * 1e-3
        d = per_kernel.setdefault(name, [0.0, 0.0, 0])
        d[0] += flops
        d[1] += t
        d[2] += 1
        e = per_shape.setdefault((name, shape), [0.0, 0.0, 0])
        e[0] += flops
        e[1] += t
        e[2] += 1
    if os.environ.get("ICGAN_BENCH_SHAPES"):
        tot = sum(v[1] for v in per_shape.values())
        print("# kernel (B,H,W,Cin,Cout,k): launches, ms total, share of TC time, TFLOP/s", file=sys.stderr)
        for (name, shape), v in sorted(per_shape.items(), key=lambda kv: -kv[1][1])[:40]:
            print(f"# {name:16s} {str(shape):34s} {v[2]:5d} {v[1]*1e3:9.2f} {100*v[1]/tot:5.1f}% {v[0]/v[1]*1e-12:8.1f}",
                  file=sys.stderr)
    kinfo = {k: {"launches": v[2], "avg_ms": v[1] / v[2] * 1e3, "tflops": v[0] / v[1] * 1e-12,
                 "share_of_step": v[1] / (prof_ms * 1e-3 * prof_steps)} for k, v in per_kernel.items()}
    dom = max(per_kernel, key=lambda k: per_kernel[k][1]) if per_kernel else None
    roofline = None
    if dom:
        ach = per_kernel[dom][0] / per_kernel[dom][1] * 1e-12
        label = {"tc_conv_kernel": "tc_conv_halo_kernel + tc_conv_kernel + tc_conv_rgb_kernel (conv forward and dgrad)",
                 "tc_wgrad_kernel": "tc_wgrad_kernel + tc_wgrad_halo_kernel"}.get(dom, dom)
        # DRAM bytes of one representative launch of the dominant kernel from the committed `ncu --set full` capture
        traffic, traffic_launch = None, None
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f).get(dom)
            if tj:
                traffic, traffic_launch = tj["dram_bytes_per_launch"], tj["launch"]
        roofline = {"kernel": label, "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["tf_sustained"], "traffic": traffic, "traffic_launch": traffic_launch,
                    "peak_source": pk["source"] + ", sustained",
                    "flops_per_launch": per_kernel[dom][0] / per_kernel[dom][2], "kernels": kinfo,
                    "timed_in": (f"eager instrumented pass of {prof_steps} steps ({prof_ms:.1f} ms each) right after the "
                                 "graph-replayed timed region" if graphed else "the timed region")}
    f_step = 4 * w["G_f"] + 8 * w["D_f"]  # GF per image, reference step model
    step_tf = f_step * 1e9 * value / world * 1e-12
    line = {"metric": f"{METRIC} {R}x{R}", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": w["name"], "per_gpu_batch": Bg, "micro_batch": m, "accumulations": acc,
                       "global_batch": Bg * world, "parallelism": f"dp{world}", "l2": "inputs larger than L2",
                       "optimizer": "icgan_adam_ema_step: Adam + EMA fused, one launch per network", "step_gflop_per_image": f_step,
                       "cuda_graphs": graphed,
                       "peak_hbm_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1)},
            "roofline": roofline,
            "step_roofline": {"achieved_tflops_per_gpu": step_tf, "frac_of_sustained_peak": step_tf / pk["tf_sustained"]},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
    if world == 1 and not args.no_cpu_baseline:
        cb = 1 if w["resolution"] >= 256 else 2
        ips, dt, workers, threads = oracle_pool_rate(args.workload, cb, 1, 1)
        line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": workers * threads, "kind": "port",
                                "sample": f"1 warm oracle G+D step (CPU restatement of train_fns.py:40-191) in each of "
                                          f"{workers} worker processes x {threads} threads, micro-batch {cb}, fp32, "
                                          f"{dt:.1f} s per step"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
