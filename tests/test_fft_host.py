"""CPU check of the FFT building blocks of the CUDA kernel (everyvoice_b200/csrc/evfeat_fft.cuh).

The butterfly / first-stage-fusion functions are ``__host__ __device__``; tests/native/fft_host_check.cu
simulates the 32 lanes of a warp on the host (window-fused first pass, transpose, twiddle-fused second
pass) and compares the 1024-point result with a float64 DFT.  This pins the index arithmetic
(bit-reversed register positions, twiddle table layout) without a GPU."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_warp_fft_index_arithmetic_on_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    exe = tmp_path / "fft_host_check"
    subprocess.run(
        [nvcc, "-O2", "-std=c++17", "-I", str(ROOT / "everyvoice_b200" / "csrc"), "-I", str(ROOT / "include"),
         "-o", str(exe), str(ROOT / "tests" / "native" / "fft_host_check.cu")],
        check=True, capture_output=True, text=True,
    )
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "fft1024 max abs err" in r.stdout
