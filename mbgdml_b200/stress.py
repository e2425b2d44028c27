"""Stress helpers (reference mbgdml/stress.py:31-152).

The group virial sum_k r_k (x) f_k of every fragment (stress.virial_atom_loop, called from
gdml_predict.py:266-267) is accumulated inside the contraction kernel; this module keeps
the host-side post-processing and the finite-difference driver, which is just 18 more
``predict`` calls on the GPU.
"""
from __future__ import annotations

import numpy as np

from .periodic import Cell

VOIGT_INDICES = [[0, 0], [1, 1], [2, 2], [1, 0], [0, 2], [0, 1]]


def to_voigt(arr):
    """3x3 -> (xx, yy, zz, yz, xz, xy) (stress.py:59-72)."""
    return np.array([arr[i, j] for i, j in VOIGT_INDICES])


def virial_finite_diff(Z, R, entity_ids, comp_ids, cell_vectors, mbe_pred, dh=1e-4, only_normal_stress=False):
    """Finite-difference virial W_ij = (E(h_ij + dh/2) - E(h_ij - dh/2)) / dh (stress.py:75-152)."""
    compute_stress_original = mbe_pred.compute_stress
    cell_original = mbe_pred.periodic_cell
    mbe_pred.compute_stress = False  # prevents recursion
    R_fractional = np.dot(R, np.linalg.inv(cell_vectors))
    dEdh = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            if only_normal_stress and (i != j):
                continue
            cell_vectors_ = np.array(cell_vectors, dtype=np.float64)
            cell_vectors_[i, j] -= dh / 2
            mbe_pred.periodic_cell = Cell(cell_vectors_.copy(), pbc=True)
            E_minus, _ = mbe_pred.predict(Z, np.dot(R_fractional, cell_vectors_), entity_ids, comp_ids)
            cell_vectors_[i, j] += dh
            mbe_pred.periodic_cell = Cell(cell_vectors_.copy(), pbc=True)
            E_plus, _ = mbe_pred.predict(Z, np.dot(R_fractional, cell_vectors_), entity_ids, comp_ids)
            dEdh[i, j] = (E_plus[0] - E_minus[0]) / dh
    mbe_pred.compute_stress = compute_stress_original
    # NOTE: the reference leaves the last perturbed Cell installed (stress.py:140-151); it is restored here.
    mbe_pred.periodic_cell = cell_original
    return dEdh
