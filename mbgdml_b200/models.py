"""GDML model container (reference mbgdml/models/gdml_model.py:35-191, base.py:28-50)."""
from __future__ import annotations

import ctypes as C
import hashlib
import itertools
import weakref

import numpy as np

from . import _native

_uid = itertools.count()


def _from_r(r, lat_and_inv=None):
    """Descriptor and (compressed) Jacobian of one geometry, evaluated on the GPU (``mbg_from_r``): the callable the
    reference stores as ``gdmlModel.desc_func`` (gdml_model.py:95, _gdml/desc.py:203-233).  ``r`` is the flattened
    (3N,) geometry; ``lat_and_inv`` the model's (lattice, inverse) pair or None."""
    r = np.asarray(r, dtype=np.float64).reshape(1, -1, 3)
    lattice = None if lat_and_inv is None else lat_and_inv[0]
    x, J = _native.get_context().from_r(r, lattice)
    return x, J[0]


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
        self.desc_func = _from_r
        # model["lattice"] (vectors as columns): the kernels clamp pair differences to it (_gdml/desc.py:42-76)
        self.lat_and_inv = ((np.asarray(model["lattice"], dtype=np.float64), np.linalg.inv(model["lattice"]))
                            if "lattice" in model else None)
        self.n_train = model["R_desc"].shape[1]
        self.stdev = model["std"] if "std" in model else 1.0
        self.integ_c = model["c"]
        self.n_perms = model["perms"].shape[0]
        self.tril_perms_lin = model["tril_perms_lin"]
        self.alphas_E = model["alphas_E"] if "alphas_E" in model.keys() else None
        self.use_torch = False
        self._perm_cache = {}

    # The reference expands the training set under every permutation at load time (gdml_model.py:109-125: two
    # (T P, D) arrays, 27.6 MB for a water trimer model).  The kernels never need them -- they permute the query --
    # so here they are built only if a foreign predictor asks (pure indexing, no arithmetic), then cached.
    def _expanded(self, key):
        if key not in self._perm_cache:
            src = self.model_dict["R_desc"].T if key == "R_desc" else self.model_dict["R_d_desc_alpha"]
            self._perm_cache[key] = (
                np.tile(src, self.n_perms)[:, self.tril_perms_lin]
                .reshape(self.n_train, self.n_perms, -1, order="F")
                .reshape(self.n_train * self.n_perms, -1))
        return self._perm_cache[key]

    @property
    def R_desc_perms(self):
        return self._expanded("R_desc")

    @property
    def R_d_desc_alpha_perms(self):
        return self._expanded("R_d_desc_alpha")

    @property
    def alphas_E_lin(self):
        if self.alphas_E is None:
            return None
        return np.tile(np.asarray(self.alphas_E)[:, None], (1, self.n_perms)).ravel()

    @property
    def code_version(self):
        if hasattr(self, "model_dict"):
            return self.model_dict["code_version"].item()
        raise AttributeError("No model is loaded.")

    @property
    def md5(self):
        """MD5 of the model (gdml_model.py:141-163 + utils.md5_data: the attributes present among md5_train, Z, R_desc,
        R_d_desc_alpha, sig, alphas_F, digested in sorted key order)."""
        keys = sorted(k for k in ("md5_train", "Z", "R_desc", "R_d_desc_alpha", "sig", "alphas_F") if hasattr(self, k))
        md5_hash = hashlib.md5()
        for key in keys:
            d = getattr(self, key)
            if isinstance(d, np.ndarray):
                d = d.ravel()
            md5_hash.update(hashlib.md5(d).digest())
        return md5_hash.hexdigest()

    def add_modifications(self, dset):
        """gdml_model.py:165-179: transfer entity_ids / comp_ids of a data set into the model dict."""
        for key in ("entity_ids", "comp_ids"):
            self.model_dict[key] = dset[key]
        self.model_dict["md5"] = np.array(self.md5)

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
            lat = _native.as_f64(self.lat_and_inv[0]).reshape(9) if self.lat_and_inv is not None else None
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
                criteria_entity_ids=_native.ptr(eids, _native._ip), lattice=_native.ptr(lat, _native._dp),
            )
            return d, (R_desc, RdA, tpl, aE, eids, lat)

        if key not in ctx._models:  # the device copy goes away with this object
            weakref.finalize(self, ctx.drop_model, key)
        return ctx.model_handle(key, build)

    def save(self, path):
        np.savez_compressed(path, **self.model_dict)
