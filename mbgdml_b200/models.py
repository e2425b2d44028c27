"""GDML model container (reference mbgdml/models/gdml_model.py:35-191, base.py:28-50)."""
from __future__ import annotations

import ctypes as C
import itertools

import numpy as np

from . import _native

_uid = itertools.count()


class Model:
    """Parent class of model objects (models/base.py:28-50)."""

    def __init__(self, criteria=None):
        self.criteria = criteria


class gdmlModel(Model):  # pylint: disable=invalid-name
    """Loads an sGDML ``.npz``/dict.  The reference expands the training descriptors under
    every permutation on the host (gdml_model.py:109-118); here the un-expanded arrays are
    uploaded once per GPU context and the kernels permute the query instead."""

    def __init__(self, model, comp_ids=None, criteria=None, for_predict=True):
        super().__init__(criteria)
        self.type = "gdml"
        self._uid = next(_uid)
        self._load(model, comp_ids)
        if for_predict:
            self._unpack()

    def _load(self, model, comp_ids):
        if isinstance(model, str):
            model = dict(np.load(model, allow_pickle=True))
        elif not isinstance(model, dict):
            raise TypeError(f"{type(model)} is not string or dict")
        self.model_dict = model
        if "z" in model:
            self.Z = model["z"]
        elif "Z" in model:
            self.Z = model["Z"]
        else:
            raise KeyError("z or Z not found in model")
        self.r_unit = model["r_unit"].item()
        self.entity_ids = model["entity_ids"] if "entity_ids" in model.keys() else None
        self.comp_ids = comp_ids if comp_ids is not None else model["comp_ids"]
        self.nbody_order = len(self.comp_ids)

    def _unpack(self):
        model = self.model_dict
        self.sig = model["sig"]
        self.n_atoms = self.Z.shape[0]
        if "lattice" in model:
            raise NotImplementedError("models carrying their own `lattice` (sGDML periodic descriptors, "
                                      "desc.py:42-76) have no device implementation; use periodic.Cell")
        self.lat_and_inv = None
        self.n_train = model["R_desc"].shape[1]
        self.stdev = model["std"] if "std" in model else 1.0
        self.integ_c = model["c"]
        self.n_perms = model["perms"].shape[0]
        self.tril_perms_lin = model["tril_perms_lin"]
        self.alphas_E = model["alphas_E"] if "alphas_E" in model.keys() else None

    # -- device side --------------------------------------------------------------------
    def _criteria_key(self):
        c = self.criteria
        if c is None:
            return None
        cut = tuple(c.cutoff) if isinstance(c.cutoff, list) else c.cutoff
        eids = c.desc_kwargs.get("entity_ids") if isinstance(c.desc_kwargs, dict) else None
        return (id(c.desc), cut, c.bound, None if eids is None else tuple(np.asarray(eids).tolist()))

    def native_handle(self, ctx):
        """Device copy of this model (+ its criteria) in context ``ctx``; cached."""
        key = (self._uid, self._criteria_key())

        def build():
            m = self.model_dict
            n_atoms = int(self.Z.shape[0])
            dim_d = n_atoms * (n_atoms - 1) // 2
            R_desc = _native.as_f64(m["R_desc"])
            RdA = _native.as_f64(m["R_d_desc_alpha"])
            tpl = np.ascontiguousarray(m["tril_perms_lin"], dtype=np.int64)
            T, P = R_desc.shape[1], int(m["perms"].shape[0])
            if R_desc.shape != (dim_d, T) or RdA.shape != (T, dim_d) or tpl.shape != (P * dim_d,):
                raise ValueError("model arrays have inconsistent shapes")
            aE = _native.as_f64(self.alphas_E) if self.alphas_E is not None else None
            if self.criteria is None:
                kind, bound, lo, hi, eids = _native.CRIT_NONE, _native.BOUND_UPPER, 0.0, 0.0, None
            else:
                kind, bound, lo, hi, eids = self.criteria.native_spec(n_atoms)
            d = _native.ModelDesc(
                n_atoms=n_atoms, order=int(self.nbody_order), n_train=T, n_perms=P,
                sig=float(self.sig), c=float(self.integ_c), std=float(self.stdev),
                R_desc=_native.ptr(R_desc, _native._dp), R_d_desc_alpha=_native.ptr(RdA, _native._dp),
                tril_perms_lin=_native.ptr(tpl, _native._lp), alphas_E=_native.ptr(aE, _native._dp),
                criteria_kind=kind, criteria_bound=bound, criteria_lo=lo, criteria_hi=hi,
                criteria_entity_ids=_native.ptr(eids, _native._ip),
            )
            return d, (R_desc, RdA, tpl, aE, eids)

        return ctx.model_handle(key, build)

    def save(self, path):
        np.savez_compressed(path, **self.model_dict)
