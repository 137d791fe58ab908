"""Runs the standalone tcgen05/TMA kernel self-test (tests/cuda/tc_selftest.cu) on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_tc_selftest_binary(cuda_device):
    exe = os.path.join(ROOT, "tests", "cuda", "tc_selftest")
    assert os.path.exists(exe), "build first: python __graft_entry__.py"
    out = subprocess.run([exe, "all"], capture_output=True, text=True, timeout=300)
    print(out.stdout[-4000:])
    assert out.returncode == 0 and "TC_SELFTEST PASSED" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
