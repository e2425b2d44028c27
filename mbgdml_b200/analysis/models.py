"""Matern-5/2 covariances of structures against a GDML training set (reference mbgdml/analysis/models.py:31-139).

    k(x, x_j) = (1 + sqrt(5) d / sig + 5 d^2 / (3 sig^2)) exp(-sqrt(5) d / sig),   d = |x - x_j|

The reference loops over structures on the host against the permutation-expanded ``R_desc_perms``; here one
kernel launch (``k_mat52``, csrc/kernel_mat.cuh) evaluates all structures against the un-expanded device model,
permuting the query descriptor instead.  Output rows keep the reference's order (row ``t * n_perms + p``).
"""
from __future__ import annotations

import numpy as np

from .. import _native


def gdml_mat52_wrk(r_desc, sig, n_perms, R_desc_perms):  # pylint: disable=unused-argument
    """Covariances of ONE descriptor ``r_desc`` (D,) with explicit rows ``R_desc_perms`` (n_rows, D)
    (analysis/models.py:31-81).  ``n_perms`` is accepted for signature parity; the rows are already expanded."""
    R_desc_perms = np.asarray(R_desc_perms, dtype=np.float64)
    r_desc = np.asarray(r_desc, dtype=np.float64).reshape(-1)
    if R_desc_perms.ndim != 2 or R_desc_perms.shape[1] != r_desc.shape[0]:
        raise ValueError("R_desc_perms must be (n_rows, D) with D = len(r_desc)")
    return _native.get_context().mat52_rows(r_desc[None, :], sig, R_desc_perms)[0]


def gdml_mat52(model, Z, R):
    """Covariances between all structures in ``R`` (n_atoms, 3) or (n_R, n_atoms, 3) and the training set of
    ``model`` (a :class:`mbgdml_b200.gdmlModel`): (n_train * n_perms,) or (n_R, n_train * n_perms)
    (analysis/models.py:84-139)."""
    R = np.asarray(R, dtype=np.float64)
    if R.ndim == 2:
        R = R[None, ...]
    n_atoms = len(Z)
    if R.ndim != 3 or R.shape[1:] != (n_atoms, 3) or n_atoms != model.n_atoms:
        raise ValueError(f"R must be (n_R, {model.n_atoms}, 3) for this model")
    ctx = _native.get_context()
    out = ctx.mat52(model.native_handle(ctx), R, int(model.n_train) * int(model.n_perms))
    return out[0] if out.shape[0] == 1 else out
