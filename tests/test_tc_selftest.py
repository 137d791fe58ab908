"""Runs the standalone tcgen05/TMA kernel self-test (tests/cuda/tc_selftest.cu) on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{}, {"ICGAN_TC_WGRAD_HALO": "2"}, {"ICGAN_TC_HALO": "0", "ICGAN_TC_WGRAD_HALO": "0"}],
                         ids=["default", "wgrad-halo-everywhere", "per-tap-kernels"])
def test_tc_selftest_binary(cuda_device, env):
    """All conv / wgrad cases against a CPU reference; the env variants force each kernel family onto every eligible
    shape (by default the halo weight-gradient kernel only takes layers with >= 192 channels)."""
    exe = os.path.join(ROOT, "tests", "cuda", "tc_selftest")
    assert os.path.exists(exe), "build first: python __graft_entry__.py"
    out = subprocess.run([exe, "all"], capture_output=True, text=True, timeout=300, env={**os.environ, **env})
    print(out.stdout[-4000:])
    assert out.returncode == 0 and "TC_SELFTEST PASSED" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.gpu
def test_tc_selftest_real_shapes(cuda_device):
    """Mode `perf`: everything above plus the layer shapes of BASELINE config 3 at real widths and batch (384@64x64 B=32,
    96@256x256 B=8 and B=256, 1536@8x8/16x16 B=64, 768@32x32, 192@128x128 ...), each verified against the host fp64
    reference on 6000 sampled outputs and then timed (the timings are informational; the assertion is correctness)."""
    exe = os.path.join(ROOT, "tests", "cuda", "tc_selftest")
    out = subprocess.run([exe, "perf"], capture_output=True, text=True, timeout=1200)
    print(out.stdout[-6000:])
    assert out.returncode == 0 and "TC_SELFTEST PASSED" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
