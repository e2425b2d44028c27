"""GDML predictor plugin (reference mbgdml/predictors/gdml_predict.py:170-364).

Same call signatures as the reference's ``predict_gdml`` / ``predict_gdml_decomp``;
every fragment is enumerated, culled, contracted and scattered by the CUDA kernels
behind the C ABI.  No NumPy arithmetic happens here.
"""
from __future__ import annotations

import numpy as np

from . import _engine, _native


def predict_gdml(Z, R, entity_ids, entity_combs, model, periodic_cell=None, **kwargs):
    """Total n-body energy and forces of a single structure (gdml_predict.py:170-272).

    Returns ``(E, F)`` or ``(E, F, virial)`` when ``compute_virial`` is requested under a
    periodic cell."""
    R = np.asarray(R)
    assert R.ndim == 2
    alchemy_scalers = kwargs.get("alchemy_scalers", None)
    compute_virial = bool(bool(periodic_cell) and kwargs.get("compute_virial", False))
    ctx = _native.get_context()
    E, F, V, stats = _engine.predict_total(
        ctx, Z, R, entity_ids, [model], [entity_combs], [alchemy_scalers],
        periodic_cell if periodic_cell else None, compute_virial=compute_virial)
    predict_gdml.last_stats = stats
    if compute_virial:
        return float(E[0]), F[0], V[0]
    return float(E[0]), F[0]


def predict_gdml_decomp(Z, R, entity_ids, entity_combs, model, **kwargs):
    """All individual n-body energies and forces of a single structure
    (gdml_predict.py:275-364); rows of rejected fragments are NaN."""
    R = np.asarray(R)
    assert R.ndim == 2
    entity_combs = np.asarray(entity_combs)
    if entity_combs.ndim == 1:
        entity_combs_2d = entity_combs.reshape(-1, 1)
    else:
        entity_combs_2d = entity_combs
    ctx = _native.get_context()
    E, F = _engine.predict_decomp(ctx, Z, R, entity_ids, model, entity_combs_2d,
                                  alchemy_scalers=kwargs.get("alchemy_scalers", None))
    return E, F, entity_combs


def accept_mask(Z, R, entity_ids, entity_combs, model, periodic_cell=None):
    """Which candidate fragments ``predict_gdml`` evaluates (gdml_predict.py:214-235: slice -> ``Cell.r_mic`` ->
    ``Criteria.accept``), decided on device for one or several frames: uint8 (n_frames, n_candidates) with
    0 = tuple not ascending (never yielded by ``gen_combs``), 1 = accepted, 2 = rejected.  ``entity_combs`` is a
    ``gen_combs`` generator (candidates = the product space of its sets) or explicit tuples."""
    return _engine.accept_mask(_native.get_context(), Z, R, entity_ids, model, entity_combs,
                               periodic_cell if periodic_cell else None)
