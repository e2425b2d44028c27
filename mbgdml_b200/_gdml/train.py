"""Force-field kernel matrix of the closed-form GDML trainer
(reference mbgdml/_gdml/train.py:92-292 ``_assemble_kernel_mat_wkr``, 1105-1337 ``GDMLTrain._assemble_kernel_mat``).

The reference fills K column block by column block in a multiprocessing pool (or through a torch module); here
one launch of ``k_kernel_mat`` (csrc/kernel_mat.cuh) computes every (i, j) block, one CTA per ordered pair of
training points.  Only the full matrix (``col_idxs = np.s_[:]``) is built on the device; column subsets
(the iterative solver's preconditioner path) raise ``NotImplementedError`` -- there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from .. import _native


def assemble_kernel_mat(R_desc, R_d_desc, tril_perms_lin, sig, desc=None, use_E_cstr=False, col_idxs=np.s_[:],
                        alloc_extra_rows=0):  # pylint: disable=unused-argument
    """Same arguments and return value as ``GDMLTrain._assemble_kernel_mat`` (train.py:1105-1115):
    ``R_desc`` (M, D), ``R_d_desc`` (M, D, 3) compressed Jacobians (``Desc.from_R``), ``tril_perms_lin`` (P*D,).
    Returns K of shape (M 3N [+ M] + alloc_extra_rows, M 3N [+ M]); the extra rows are left zero-filled."""
    if not (isinstance(col_idxs, slice) and col_idxs == slice(None)):
        raise NotImplementedError("only the full kernel matrix (col_idxs = np.s_[:]) is built on the device")
    if alloc_extra_rows < 0:
        raise ValueError("alloc_extra_rows must be >= 0")
    K = _native.get_context().kernel_mat(R_desc, R_d_desc, tril_perms_lin, sig, use_E_cstr)
    if alloc_extra_rows:
        K = np.vstack([K, np.zeros((alloc_extra_rows, K.shape[1]))])
    return K


class GDMLTrain:
    """The one method of the reference's trainer that runs on the device here
    (``GDMLTrain._assemble_kernel_mat``, train.py:1105-1337); the rest of the trainer is out of scope."""

    def __init__(self, max_memory=None, max_processes=None, use_torch=False):  # pylint: disable=unused-argument
        pass

    def _assemble_kernel_mat(self, R_desc, R_d_desc, tril_perms_lin, sig, desc=None, use_E_cstr=False,
                             col_idxs=np.s_[:], alloc_extra_rows=0):
        return assemble_kernel_mat(R_desc, R_d_desc, tril_perms_lin, sig, desc, use_E_cstr, col_idxs, alloc_extra_rows)
