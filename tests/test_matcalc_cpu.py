"""The matrix-calculus helpers of the per-Gaussian backward kernels (r2_gaussian_b200/csrc/r2x_matcalc.cuh) are
__host__ __device__: compile tests/host/matcalc_check.cu for the host with nvcc and check every identity against
finite differences on the CPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_matcalc_identities_against_finite_differences(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "matcalc_check")
    src = os.path.join(ROOT, "tests", "host", "matcalc_check.cu")
    r = subprocess.run([nvcc, "-std=c++17", "-O1", "-o", exe, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "matcalc_check: ok" in r.stdout
