"""The arrival key that picks a pixel's supporting surfel (csrc/sm_kernels.cuh) is a bijection on the
slots: host-side round trip over 5.7 M (slot, kind, pixel) triples per wave size, against the plain
division formulas."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_tiebreak_key_roundtrip(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("needs nvcc (host compilation of the shared header)")
    exe = tmp_path / "tiebreak_check"
    build = subprocess.run([nvcc, "-std=c++17", "-O2", "-I", str(ROOT / "surfelmeshing_b200" / "csrc"), "-o", str(exe),
                            str(ROOT / "tests" / "stubs" / "tiebreak_check.cu")], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout
    assert run.stdout.count(" 0 bad") == 8, run.stdout
