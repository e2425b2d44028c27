"""Install the UNMODIFIED reference package into the git-ignored ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY.  The reference (keithgroup/mbGDML) is pure Python, so "building" it is a
``pip install --no-deps --target``; nothing is copied into the tracked tree (``oracle/_ref/`` is listed in
.gitignore, not in .gpurunignore, so the installed package travels to the GPU box with the snapshot).
Its missing third-party imports (ray, ase, cclib, qcelemental) are the stubs of ``oracle/refshim``;
``oracle/reference.py`` puts both on ``sys.path``.

    python oracle/fetch_ref.py            # needs /root/reference (build container only)

The checker (tests/, smoke()) and the CPU arm of bench.py (`--impl reference`, `cpu_baseline`) use it when
present; ``__graft_entry__.build()`` calls this script when /root/reference exists.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("MBGDML_REFERENCE", "/root/reference")


def fetch(force=False):
    """Returns the install directory, or None when the reference source tree is not available."""
    marker = os.path.join(DEST, "mbgdml", "mbe.py")
    if os.path.exists(marker) and not force:
        return DEST
    if not os.path.isdir(os.path.join(SRC, "mbgdml")):
        return None
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "src")  # the build writes egg-info next to setup.py; /root/reference is read-only
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns(".git", "docs", "tests"))
        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DEST, work]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0 or not os.path.exists(marker):
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("installing the reference into oracle/_ref failed")
    return DEST


if __name__ == "__main__":
    print(fetch(force="--force" in sys.argv))
