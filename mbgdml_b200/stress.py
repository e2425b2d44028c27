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
    """Finite-difference virial W_ij = (E(h_ij + dh/2) - E(h_ij - dh/2)) / dh (stress.py:75-152).

    The reference runs 18 (6 with ``only_normal_stress``) ``predict`` calls, each under its own strained ``Cell``.  With
    this package's ``predict_gdml`` plugin they are ONE native call: the strained geometries form an 18-frame batch
    and every frame carries its own cell (``mbg_system_desc.frame_cells``); repeated calls (MD) replay one CUDA graph.
    Any other plugin is driven call by call like the reference."""
    from .predictors import predict_gdml

    cell_vectors = np.array(cell_vectors, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64)
    if R.ndim == 3:
        if R.shape[0] != 1:
            raise ValueError("virial_finite_diff takes one structure")
        R = R[0]
    R_fractional = np.dot(R, np.linalg.inv(cell_vectors))
    pairs = [(i, j) for i in range(3) for j in range(3) if not (only_normal_stress and i != j)]
    cells, frames = [], []
    for i, j in pairs:  # backward, then forward, exactly the reference's two perturbed cells
        for sign in (-1.0, 1.0):
            cell_vectors_ = cell_vectors.copy()
            cell_vectors_[i, j] -= dh / 2
            if sign > 0:
                cell_vectors_[i, j] += dh
            cells.append(cell_vectors_)
            frames.append(np.dot(R_fractional, cell_vectors_))
    dEdh = np.zeros((3, 3))
    if mbe_pred.predict_model is predict_gdml:
        cutoffs = [Cell(c, pbc=True).cutoff for c in cells]  # Cell(cell_vectors_tmp, pbc=True): the default cut-off
        E, _, _ = mbe_pred._predict_batched(Z, np.array(frames), entity_ids, comp_ids, False,
                                            frame_cells=(np.array(cells), np.array(cutoffs)))
        for k, (i, j) in enumerate(pairs):
            dEdh[i, j] = (E[2 * k + 1] - E[2 * k]) / dh
        return dEdh
    compute_stress_original = mbe_pred.compute_stress
    cell_original = mbe_pred.periodic_cell
    mbe_pred.compute_stress = False  # prevents recursion
    try:
        for k, (i, j) in enumerate(pairs):
            mbe_pred.periodic_cell = Cell(cells[2 * k].copy(), pbc=True)
            E_minus, _ = mbe_pred.predict(Z, frames[2 * k], entity_ids, comp_ids)
            mbe_pred.periodic_cell = Cell(cells[2 * k + 1].copy(), pbc=True)
            E_plus, _ = mbe_pred.predict(Z, frames[2 * k + 1], entity_ids, comp_ids)
            dEdh[i, j] = (E_plus[0] - E_minus[0]) / dh
    finally:
        mbe_pred.compute_stress = compute_stress_original
        # NOTE: the reference leaves the last perturbed Cell installed (stress.py:140-151); it is restored here.
        mbe_pred.periodic_cell = cell_original
    return dEdh
