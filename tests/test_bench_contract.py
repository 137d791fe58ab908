"""bench.py contract pieces that can be checked without a GPU: the reference arm prints exactly one JSON line with the
agreed keys (oracle on the host cores, here on the cheap ic64 workload), and the B200 arm refuses to run without CUDA
instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", ""))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run("--impl", "reference", "--workload", "ic64", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["steps"] == 1 and d["warmup"] == 0


def test_reference_arm_is_rank0_only_under_torchrun_env():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""  # non-zero ranks exit 0 without work or output


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a host without CUDA")
def test_b200_arm_has_no_cpu_path():
    out = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-e2e", timeout=300)
    assert out.returncode != 0
    assert "no CPU path" in (out.stderr + out.stdout)
