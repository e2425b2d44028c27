"""Build the in-tree CUDA library for sm_100a:  python -m mbgdml_b200.build [--force] [-v]"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmbgdml_b200.so")
STAMP = OUT + ".srchash"  # hash of the sources the library was built from (travels with it, not tracked)
SOURCES = ["api.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    # the CUDA runtime as a shared library (the process usually has one already: torch's): the shipped binary then
    # carries only the runtime entry points it calls
    "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64",
]


def _source_files():
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    return files + [os.path.join(HERE, "..", "include", "mbgdml_b200.h"), os.path.abspath(__file__)]


def source_hash():
    """SHA-256 over every source the library is compiled from (+ the flags)."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in _source_files():
        with open(path, "rb") as f:
            h.update(os.path.basename(path).encode() + b"\0" + f.read())
    return h.hexdigest()


def built_hash():
    try:
        with open(STAMP) as f:
            return f.read().strip()
    except OSError:
        return None


def _stale():
    return not os.path.exists(OUT) or built_hash() != source_hash()


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libmbgdml_b200.so (nvcc cross-compiles without a GPU).  The library is rebuilt whenever the
    hash of its sources differs from the hash recorded next to it, so a shipped binary can never silently be stale."""
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = ["-DMBG_TUNING_VARIANTS"] if os.environ.get("MBG_TUNING_VARIANTS") else []
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [
        os.path.join(CSRC, s) for s in SOURCES]
    want = source_hash()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building " + OUT)
    if verbose:
        sys.stderr.write(res.stderr)
    with open(STAMP, "w") as f:
        f.write(want + "\n")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
