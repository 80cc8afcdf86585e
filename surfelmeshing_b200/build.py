"""Build recipe for libsurfel_b200.so (hand-written sm_100a kernels + the C ABI).

`python -m surfelmeshing_b200.build` compiles every .cu under csrc/ with nvcc for
sm_100a only (no fallback architectures) and links them in-tree into
surfelmeshing_b200/libsurfel_b200.so, so that the library travels to the GPU box
with the repository snapshot.

Flags: -ftz=true -fmad=false. The kernels spell out every fp32 operation (csrc/sm_math.cuh)
in the order of the reference's -use_fast_math SASS; -fmad=false guarantees the compiler
contracts nothing on its own.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libsurfel_b200.so"
SOURCES = ["api.cu", "pipeline.cu", "transfer.cu", "preprocess.cu", "integrate.cu", "regularize.cu", "knn.cu"]
HEADERS = ["sm_math.cuh", "sm_kernels.cuh", "sm_handle.cuh", "../../include/surfel_b200.h"]

NVCC_FLAGS = [
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo",
    "-ftz=true", "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    nvcc = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found: libsurfel_b200.so cannot be built (there is no CPU fallback)")
    return nvcc


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu -> libsurfel_b200.so (incremental). Returns the library path."""
    nvcc = _nvcc()
    obj_dir = PKG_DIR / "build"
    obj_dir.mkdir(exist_ok=True)
    headers = [CSRC / h for h in HEADERS] + [Path(__file__)]
    objects = []
    for src in SOURCES:
        obj = obj_dir / (src.replace(".cu", ".o"))
        objects.append(obj)
        if force or _stale(obj, [CSRC / src] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
            (obj_dir / (src + ".ptxas.log")).write_text(res.stderr)
    if force or _stale(LIB_PATH, objects):
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB_PATH), *map(str, objects)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
