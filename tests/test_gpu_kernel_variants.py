"""The launch-size dispatch of the library picks kernels the small parity cases never reach on their own:
`k_matern_small` (fragments of 2-4 atoms) only runs from 48 k queries on, `k_kernel_mat_d` (tensor-pipe kernel matrix)
only for 5- to 10-atom fragments, the persistent mode of the CTA-cooperative contraction kernel for 9- to 12-atom
fragments only on mid-size launches (and always from 13 atoms on, where the per-warp kernel is then never exercised).  These tests re-run the small parity cases with the developer overrides that force
them (MBG_MMA_CFG is read when the context is created, hence a fresh interpreter), so every feature the big cases do
not touch -- periodic cells, virial, decomposed outputs, lattice models, tiny training sets, energy constraints -- is
checked on those kernels too."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rerun(cfg, files):
    env = dict(os.environ, MBG_MMA_CFG=str(cfg))
    res = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider", *files],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]


@pytest.mark.gpu
def test_thread_per_query_kernel_forced_on_small_cases():
    _rerun(8, ["tests/test_gpu_parity.py", "tests/test_gpu_edge_cases.py", "tests/test_gpu_lattice.py", "tests/test_gpu_steps.py"])


@pytest.mark.gpu
def test_tensor_pipe_kernel_matrix_forced_on_all_fragment_sizes():
    _rerun(7, ["tests/test_train_kernel.py"])


@pytest.mark.gpu
def test_persistent_cta_kernel_forced_from_9_atoms_on():
    _rerun(10, ["tests/test_gpu_parity.py", "tests/test_gpu_edge_cases.py", "tests/test_gpu_steps.py"])


@pytest.mark.gpu
def test_per_warp_kernel_forced_for_wide_fragments():
    _rerun(9, ["tests/test_gpu_parity.py", "tests/test_gpu_edge_cases.py"])
