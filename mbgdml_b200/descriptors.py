"""Fragment acceptance criteria (reference mbgdml/descriptors.py:32-201).

``com_distance_sum`` and ``max_atom_pair_dist`` are evaluated on the GPU
(``mbg_criteria_eval``); inside predictions the same device code runs fused into
the fragment-culling kernel.  ``Criteria`` recognises these two callables by
identity -- an arbitrary Python ``desc`` cannot run on device and raises.
"""
from __future__ import annotations

import numpy as np

from . import _native

# Most-abundant-isotope masses, as qcelemental.periodictable.to_mass returns
# (reference descriptors.py:125; H and O pinned bit-exactly by the reference's
# tests/test_descriptors.py:40-49).
MASSES = {
    1: 1.00782503223, 2: 4.00260325413, 3: 7.0160034366, 4: 9.012183065,
    5: 11.00930536, 6: 12.0, 7: 14.00307400443, 8: 15.99491461957,
    9: 18.99840316273, 10: 19.9924401762, 11: 22.9897692820, 12: 23.985041697,
    13: 26.98153853, 14: 27.97692653465, 15: 30.97376199842, 16: 31.9720711744,
    17: 34.968852682, 18: 39.9623831237,
}


_MASS_TABLE = np.full(max(MASSES) + 1, np.nan)
for _z, _m in MASSES.items():
    _MASS_TABLE[_z] = _m


def masses_of(Z, strict=True):
    """Per-atom masses for atomic numbers ``Z`` (KeyError for an unknown element if ``strict``)."""
    Z = np.asarray(Z)
    if Z.size and (Z.min() < 0 or Z.max() >= len(_MASS_TABLE)):
        if strict:
            raise KeyError(f"no mass tabulated for atomic number {int(Z.max() if Z.min() >= 0 else Z.min())}")
        out = np.full(Z.shape, np.nan)
        known = (Z >= 0) & (Z < len(_MASS_TABLE))
        out[known] = _MASS_TABLE[Z[known]]
        return out
    out = _MASS_TABLE[Z]
    if strict and np.isnan(out).any():
        raise KeyError(f"no mass tabulated for atomic number {int(Z[np.isnan(out)][0])}")
    return out


def _as_batch(R):
    R = np.asarray(R, dtype=np.float64)
    return R[None, ...] if R.ndim == 2 else R


def max_atom_pair_dist(Z, R):
    """Largest atomic pair distance of each structure (descriptors.py:131-156)."""
    return _native.get_context().criteria_eval(_native.CRIT_MAX_ATOM_PAIR_DIST, None, None, _as_batch(R))


def com_distance_sum(Z, R, entity_ids):
    """Sum over entities of |CM_entity - CM_structure| (descriptors.py:159-201)."""
    return _native.get_context().criteria_eval(
        _native.CRIT_COM_DISTANCE_SUM, masses_of(Z), np.asarray(entity_ids), _as_batch(R)
    )


class Criteria:
    """Accept or reject a structure from a descriptor and a cutoff (descriptors.py:32-103)."""

    def __init__(self, desc, desc_kwargs, cutoff, bound="upper"):
        self.desc = desc
        self.desc_kwargs = desc_kwargs
        self.cutoff = cutoff
        if isinstance(self.cutoff, (tuple, list)):
            self.cutoff = sorted(self.cutoff)  # always a list
        bound = bound.lower()
        assert bound in ["upper", "lower"]
        self.bound = bound

    def accept(self, Z, R, **kwargs):
        R = np.asarray(R, dtype=np.float64)
        if R.ndim == 2:
            R = R[None, ...]
        n_R = R.shape[0]
        if self.cutoff is None:
            return np.full(n_R, True), None
        desc_v = self.desc(Z, R, **self.desc_kwargs, **kwargs)
        if isinstance(self.cutoff, list):
            accept_r = (self.cutoff[0] < desc_v) & (desc_v < self.cutoff[1])
        elif self.bound == "upper":
            accept_r = desc_v < self.cutoff
        else:
            accept_r = desc_v > self.cutoff
        if n_R == 1:
            accept_r = accept_r[0]
            desc_v = desc_v[0]
        return accept_r, desc_v

    def native_spec(self, n_atoms):
        """(kind, bound, lo, hi, entity_ids) for mbg_model_desc."""
        if self.cutoff is None:
            return _native.CRIT_NONE, _native.BOUND_UPPER, 0.0, 0.0, None
        if self.desc is com_distance_sum:
            kind = _native.CRIT_COM_DISTANCE_SUM
            extra = set(self.desc_kwargs) - {"entity_ids"}
            if extra or "entity_ids" not in self.desc_kwargs:
                raise NotImplementedError(f"com_distance_sum takes exactly entity_ids (got {sorted(self.desc_kwargs)})")
            eids = np.ascontiguousarray(self.desc_kwargs["entity_ids"], dtype=np.int32)
            if eids.shape != (n_atoms,):
                raise ValueError("criteria entity_ids must have one entry per fragment atom")
        elif self.desc is max_atom_pair_dist:
            kind = _native.CRIT_MAX_ATOM_PAIR_DIST
            if self.desc_kwargs:
                raise NotImplementedError("max_atom_pair_dist takes no keyword arguments")
            eids = None
        else:
            raise NotImplementedError(
                f"criteria descriptor {self.desc!r} has no device implementation "
                "(supported: mbgdml_b200.descriptors.com_distance_sum, max_atom_pair_dist)"
            )
        if isinstance(self.cutoff, list):
            if len(self.cutoff) != 2:
                raise ValueError("a cutoff range needs exactly two values")
            return kind, _native.BOUND_RANGE, float(self.cutoff[0]), float(self.cutoff[1]), eids
        if self.bound == "upper":
            return kind, _native.BOUND_UPPER, 0.0, float(self.cutoff), eids
        return kind, _native.BOUND_LOWER, float(self.cutoff), 0.0, eids
