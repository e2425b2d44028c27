"""Force-field kernel matrix of the closed-form GDML trainer
(reference mbgdml/_gdml/train.py:92-292 ``_assemble_kernel_mat_wkr``, 1105-1337 ``GDMLTrain._assemble_kernel_mat``).

The reference fills K column block by column block in a multiprocessing pool (or through a torch module); here
one launch (csrc/kernel_mat.cuh: tensor-pipe kernel for 5- to 10-atom fragments, warp / CTA per pair otherwise)
computes every (i, j) block.  Column subsets (``col_idxs``, what the iterative solver's preconditioner asks for,
train.py:1187-1245) are computed on the device too: only the column BLOCKS (training points) the subset touches are
assembled (``mbg_kernel_mat_cols``), the requested columns are gathered from them on the device and only those
travel to the host.  With energy constraints the subset is gathered from the full matrix, which then has to fit the
device (rows^2 doubles).  There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from .. import _native


def _torch():
    import torch  # the tensor container for device-side gathers

    return torch


def _subset_columns(n_rows, col_idxs):
    """Column indices of a ``col_idxs`` argument (slice or sorted index array), as the reference resolves them
    (train.py:1169-1185)."""
    if isinstance(col_idxs, slice):
        return np.arange(n_rows)[col_idxs]
    cols = np.asarray(col_idxs, dtype=np.int64)
    assert len(cols) == len(set(cols.tolist()))  # train.py:1175: no duplicate indices
    assert np.array_equal(cols, np.sort(cols))   # train.py:1178: ascending
    return cols


def assemble_kernel_mat(R_desc, R_d_desc, tril_perms_lin, sig, desc=None, use_E_cstr=False, col_idxs=np.s_[:],
                        alloc_extra_rows=0):  # pylint: disable=unused-argument
    """Same arguments and return value as ``GDMLTrain._assemble_kernel_mat`` (train.py:1105-1115):
    ``R_desc`` (M, D), ``R_d_desc`` (M, D, 3) compressed Jacobians (``Desc.from_R``), ``tril_perms_lin`` (P*D,).
    Returns K[:, col_idxs] of shape (M 3N [+ M] + alloc_extra_rows, n_cols); the extra rows are zero-filled
    (uninitialised in the reference)."""
    if alloc_extra_rows < 0:
        raise ValueError("alloc_extra_rows must be >= 0")
    ctx = _native.get_context()
    full = isinstance(col_idxs, slice) and col_idxs == slice(None)
    if full:
        K = ctx.kernel_mat(R_desc, R_d_desc, tril_perms_lin, sig, use_E_cstr)
    else:
        R_desc = _native.as_f64(R_desc)
        T, D = R_desc.shape
        n_atoms = int((1 + np.sqrt(8 * D + 1)) / 2)
        dim_i = 3 * n_atoms
        n_rows = T * dim_i + (T if use_E_cstr else 0)
        cols = _subset_columns(n_rows, col_idxs)
        if cols.size > n_rows or (cols.size and (cols.min() < 0 or cols.max() >= n_rows)):
            raise ValueError("Columns indexed beyond range.")  # train.py:1183-1184
        torch = _torch()
        dev = torch.device("cuda", ctx.device)
        if use_E_cstr:  # energy columns mix all training points: gather from the full matrix on the device
            Kd = torch.empty((n_rows, n_rows), dtype=torch.float64, device=dev)
            ctx.kernel_mat(R_desc, R_d_desc, tril_perms_lin, sig, True, out=Kd.data_ptr())
            pick = torch.from_numpy(cols).to(dev)
        else:
            blocks = np.unique(cols // dim_i)
            Kd = torch.empty((n_rows, blocks.size * dim_i), dtype=torch.float64, device=dev)
            ctx.kernel_mat_cols(R_desc, R_d_desc, tril_perms_lin, sig, blocks, out=Kd.data_ptr())
            pick = torch.from_numpy(np.searchsorted(blocks, cols // dim_i) * dim_i + cols % dim_i).to(dev)
        ctx.synchronize()
        K = Kd.index_select(1, pick).cpu().numpy()
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
