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


class GDMLPredict:
    """Query a trained sGDML force field (``predict`` mirrors _gdml/predict.py:1031-1182)."""

    def __init__(self, model, batch_size=None, num_workers=None, max_memory=None, max_processes=None,
                 use_torch=False):  # pylint: disable=unused-argument
        model = dict(model)
        self.n_atoms = int(np.asarray(model["z"] if "z" in model else model["Z"]).shape[0])
        model.setdefault("r_unit", np.array("Angstrom"))
        self._model = gdmlModel(model, comp_ids=np.array(["mol"]))
        self.std = float(self._model.stdev)
        self.c = float(self._model.integ_c)
        self._Z = np.asarray(self._model.Z)
        self._eids = np.zeros(self.n_atoms, dtype=np.int64)

    def prepare_parallel(self, n_bulk=1, n_reps=1, return_is_from_cache=False):  # pylint: disable=unused-argument
        """No-op: there are no CPU workers / chunk sizes to tune (predict.py:707-920)."""
        return (1, False) if return_is_from_cache else 1

    def predict(self, R=None, return_E=True):
        """R: (M, 3N) -> (E (M,), F (M, 3N)) or (F,) when ``return_E`` is False."""
        if R is None:
            raise NotImplementedError("predict() without R (training-set error) needs set_R(), which is training-side")
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
