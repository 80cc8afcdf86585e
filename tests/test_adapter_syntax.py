"""The header-only C++ adapter (include/surfel_b200_adapter.h: class vis::CUDASurfelReconstruction on
top of the C ABI) is compiled - syntax and types - against minimal stand-ins of the reference's libvis
headers (tests/stubs/), with a translation unit that makes the calls APP/main.cc makes. The real
headers need Eigen / Sophus / Qt, which this image does not have (verified absent)."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
CUDA_INCLUDE = Path("/usr/local/cuda/include")


@pytest.mark.parametrize("defines", [[], ["-DSURFEL_B200_DELTA_TRANSFER"], ["-DSURFEL_B200_NO_GL_INTEROP"]])
def test_adapter_compiles_against_stub_libvis(defines, tmp_path):
    gxx = shutil.which("g++")
    if gxx is None or not CUDA_INCLUDE.exists():
        pytest.skip("needs g++ and the CUDA headers")
    obj = tmp_path / "adapter_check.o"
    cmd = [gxx, "-std=c++14", "-Wall", "-Wextra", "-Werror", "-c", str(ROOT / "tests" / "stubs" / "adapter_check.cc"),
           "-I", str(ROOT / "tests" / "stubs"), "-I", str(ROOT / "include"), "-isystem", str(CUDA_INCLUDE), "-o", str(obj),
           *defines]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    # every sm_* function the adapter calls is declared by the header and exported by the library
    syms = subprocess.run(["nm", "-u", str(obj)], capture_output=True, text=True, check=True).stdout
    wanted = {line.split()[-1] for line in syms.splitlines() if " sm_" in line}
    assert {"sm_create", "sm_integrate", "sm_regularize", "sm_export_vertices"} <= wanted
    from surfelmeshing_b200 import _lib
    assert wanted <= set(_lib.EXPORTED_SYMBOLS), wanted - set(_lib.EXPORTED_SYMBOLS)
