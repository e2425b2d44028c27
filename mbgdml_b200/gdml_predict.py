"""sGDML bulk predictor (reference mbgdml/_gdml/predict.py:246-1182, ``GDMLPredict``): energies and forces
of M geometries of ONE molecule with one GDML model.

It is the same Matern contraction as the many-body path with every geometry treated as a frame that holds a
single one-entity fragment, so it runs through the same kernels: the M geometries go to the GPU as M frames in
one native call.  The reference's CPU tuning knobs (``batch_size``, ``num_workers``, ``prepare_parallel`` ...)
are accepted and ignored.
"""
from __future__ import annotations

import numpy as np

from . import _engine, _native
from .models import gdmlModel
from .utils import gen_combs


def geometries_from_desc(R_desc, R_d_desc):
    """Cartesian geometries (up to a translation, which the model does not see) of structures given by their
    descriptors ``R_desc`` (M, D) and compressed descriptor Jacobians ``R_d_desc`` (M, D, 3) as ``Desc.from_R``
    produces them (_gdml/desc.py:160-233): entry (i, j) of the Jacobian is (r_i - r_j) / d_ij^3 and the descriptor is
    1 / d_ij, so r_i - r_0 = R_d_desc[(i, 0)] / R_desc[(i, 0)]^3.  This is how predictions "by descriptor" (the
    training-error mode of ``GDMLPredict.predict()`` / ``GDMLTorchPredict.forward(train_idxs)``) reach the device
    kernels, which build descriptor and Jacobian from coordinates themselves."""
    R_desc = np.asarray(R_desc, dtype=np.float64)
    R_d_desc = np.asarray(R_d_desc, dtype=np.float64)
    if R_desc.ndim == 1:
        R_desc, R_d_desc = R_desc[None], R_d_desc[None]
    M, D = R_desc.shape
    n_atoms = int((1 + np.sqrt(8 * D + 1)) / 2)
    if R_d_desc.shape != (M, D, 3):
        raise ValueError(f"R_d_desc must be ({M}, {D}, 3)")
    first = np.array([i * (i - 1) // 2 for i in range(1, n_atoms)], dtype=np.int64)  # pairs (i, 0) in tril order
    R = np.zeros((M, n_atoms, 3))
    R[:, 1:, :] = R_d_desc[:, first, :] / (R_desc[:, first] ** 3)[:, :, None]
    return R


class GDMLPredict:
    """Query a trained sGDML force field (``predict`` mirrors _gdml/predict.py:1031-1182)."""

    def __init__(self, model, batch_size=None, num_workers=None, max_memory=None, max_processes=None,
                 use_torch=False):  # pylint: disable=unused-argument
        model = dict(model)
        self.n_atoms = int(np.asarray(model["z"] if "z" in model else model["Z"]).shape[0])
        model.setdefault("r_unit", np.array("Angstrom"))
        self._model_dict = model
        self.n_train = int(np.asarray(model["R_desc"]).shape[1])
        self.R_desc = None    # training-mode caches (predict.py:485-523)
        self.R_d_desc = None
        self._model = gdmlModel(model, comp_ids=np.array(["mol"]))
        self.std = float(self._model.stdev)
        self.c = float(self._model.integ_c)
        self._Z = np.asarray(self._model.Z)
        self._eids = np.zeros(self.n_atoms, dtype=np.int64)

    def prepare_parallel(self, n_bulk=1, n_reps=1, return_is_from_cache=False):  # pylint: disable=unused-argument
        """No-op: there are no CPU workers / chunk sizes to tune (predict.py:707-920)."""
        return (1, False) if return_is_from_cache else 1

    def set_R_desc(self, R_desc):  # pylint: disable=invalid-name
        """Reference to the training geometry descriptors (M, D) (predict.py:485-497)."""
        self.R_desc = R_desc

    def set_R_d_desc(self, R_d_desc):  # pylint: disable=invalid-name
        """Reference to the training geometry descriptor Jacobians (M, D, 3) (predict.py:500-523)."""
        self.R_d_desc = R_d_desc

    def set_alphas(self, alphas_F, alphas_E=None):  # pylint: disable=invalid-name
        """Reconfigure the model with new regression parameters (predict.py:525-575; iterative training):
        ``R_d_desc_alpha = Desc.d_desc_dot_vec(R_d_desc, alphas_F)`` (desc.py:341-358), uploaded as a new device model."""
        assert self.R_d_desc is not None
        a = np.asarray(alphas_F, dtype=np.float64).reshape(-1, self.n_atoms, 3)
        i, j = np.tril_indices(self.n_atoms, k=-1)
        model = dict(self._model_dict)
        model["R_d_desc_alpha"] = np.einsum("tkc,tkc->tk", np.asarray(self.R_d_desc, dtype=np.float64), a[:, j, :] - a[:, i, :])
        if alphas_E is not None:
            model["alphas_E"] = np.asarray(alphas_E, dtype=np.float64)
        self._model_dict = model
        self._model = gdmlModel(model, comp_ids=np.array(["mol"]))

    def predict(self, R=None, return_E=True):
        """R: (M, 3N) -> (E (M,), F (M, 3N)) or (F,) when ``return_E`` is False.  Without ``R`` the cached training
        descriptors and Jacobians (``set_R_desc`` / ``set_R_d_desc``) are evaluated (predict.py:1108-1120)."""
        if R is None:
            if self.R_desc is None or self.R_d_desc is None:
                raise ValueError("A reference to the training geometry descriptors and Jacobians needs to be set "
                                 "for this function to work without arguments.")  # predict.py:1112-1117
            R = geometries_from_desc(self.R_desc, self.R_d_desc).reshape(len(self.R_desc), -1)
        R = np.asarray(R, dtype=np.float64)
        if R.ndim == 1:
            R = R[None, :]
        if R.shape[1] != 3 * self.n_atoms:
            raise ValueError(f"R must be (M, {3 * self.n_atoms})")
        ctx = _native.get_context()
        E, F, _, _ = _engine.predict_total(ctx, self._Z, R.reshape(-1, self.n_atoms, 3), self._eids, [self._model],
                                           [gen_combs([np.array([0])])], [[]], None)
        F = F.reshape(R.shape[0], -1)
        return (E, F) if return_E else (F,)
