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
