"""Build the in-tree CUDA library for sm_100a:  python -m mbgdml_b200.build [--force]"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmbgdml_b200.so")
SOURCES = ["api.cu"]
HEADERS = ["common.cuh", "geom.cuh", "enumerate.cuh", "matern.cuh", "hostmath.h", "matern_mma.cuh", "mbe_accum.cuh", "kernel_mat.cuh",
           os.path.join("..", "..", "include", "mbgdml_b200.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libmbgdml_b200.so (nvcc cross-compiles without a GPU)."""
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = ["-DMBG_TUNING_VARIANTS"] if os.environ.get("MBG_TUNING_VARIANTS") else []
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [
        os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building " + OUT)
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
