"""Reference-side binding of the mbgdml_b200 C ABI: the file a keithgroup/mbGDML maintainer would add as
``mbgdml/predictors/gdml_b200_predict.py`` (option B).

It imports NOTHING from the ``mbgdml_b200`` Python package: only ``ctypes``, NumPy and the reference's own
``qcelemental`` dependency.  The objects it receives are the reference's (``mbgdml.models.gdmlModel``,
``mbgdml.descriptors.Criteria``, ``mbgdml.periodic.Cell``, ``mbgdml.alchemy.mbeAlchemyScale``), and it is used as
the ``predict_model`` plugin of the UNMODIFIED driver:

    from mbgdml.mbe import mbePredict
    from gdml_b200_predict import predict_gdml_b200, predict_gdml_decomp_b200
    E, F = mbePredict(models, predict_gdml_b200, periodic_cell=cell).predict(Z, R, entity_ids, comp_ids)

Call sites served: mbe.py:809-811 (totals, one frame x one model per call, the Python ``gen_combs`` generator
still producing the tuples) and mbe.py:949-951 (decomposed).  tests/test_gpu_dropin.py runs exactly this against
the reference's own outputs.  Library location: $MBGDML_B200_LIB, else ../mbgdml_b200/libmbgdml_b200.so.
"""
import ctypes as C
import os

import numpy as np
from qcelemental.periodictable import to_mass  # the reference's mass table (descriptors.py:125)

_dp, _ip, _lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)


class _ModelDesc(C.Structure):  # struct mbg_model_desc (include/mbgdml_b200.h), field for field
    _fields_ = [("n_atoms", C.c_int32), ("order", C.c_int32), ("n_train", C.c_int32), ("n_perms", C.c_int32),
                ("sig", C.c_double), ("c", C.c_double), ("std", C.c_double),
                ("R_desc", _dp), ("R_d_desc_alpha", _dp), ("tril_perms_lin", _lp), ("alphas_E", _dp),
                ("criteria_kind", C.c_int32), ("criteria_bound", C.c_int32),
                ("criteria_lo", C.c_double), ("criteria_hi", C.c_double), ("criteria_entity_ids", _ip), ("lattice", _dp)]


class _SystemDesc(C.Structure):  # struct mbg_system_desc
    _fields_ = [("n_atoms", C.c_int32), ("n_entities", C.c_int32), ("entity_offsets", _ip), ("entity_atoms", _ip),
                ("masses", _dp), ("cell", _dp), ("cell_cutoff", C.c_double), ("frame_cells", _dp), ("frame_cutoffs", _dp), ("pbc", _ip)]


class _Job(C.Structure):  # struct mbg_job
    _fields_ = [("model", C.c_void_p), ("set_offsets", _ip), ("set_entities", _ip), ("combs", _ip),
                ("n_combs", C.c_int64), ("n_alchemy", C.c_int32), ("alchemy_entity", _ip), ("alchemy_factor", _dp)]


_MBG_FLAG_VIRIAL = 4
_CRIT_NONE, _CRIT_COM, _CRIT_PAIR = 0, 1, 2
_BOUND_UPPER, _BOUND_LOWER, _BOUND_RANGE = 0, 1, 2
_ERRORS = {-1: ValueError, -3: NotImplementedError, -4: MemoryError}

_lib = None
_ctx = C.c_void_p()


def _library():
    global _lib
    if _lib is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.environ.get("MBGDML_B200_LIB", os.path.join(here, "..", "mbgdml_b200", "libmbgdml_b200.so"))
        lib = C.CDLL(os.path.abspath(path))
        lib.mbg_last_error.restype = C.c_char_p
        lib.mbg_context_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        lib.mbg_model_create.argtypes = [C.c_void_p, C.POINTER(_ModelDesc), C.POINTER(C.c_void_p)]
        lib.mbg_predict.argtypes = [C.c_void_p, C.POINTER(_SystemDesc), C.c_void_p, C.c_int32, C.POINTER(_Job), C.c_int32,
                                    C.c_uint32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mbg_predict_decomp.argtypes = [C.c_void_p, C.POINTER(_SystemDesc), C.c_void_p, C.POINTER(_Job), C.c_uint32,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
        _check(lib, lib.mbg_context_create(int(os.environ.get("LOCAL_RANK", "0")), C.byref(_ctx)))
        _lib = lib
    return _lib


def _check(lib, rc):
    if rc:
        raise _ERRORS.get(rc, RuntimeError)(lib.mbg_last_error().decode("utf-8", "replace"))


def _criteria_spec(criteria, n_atoms):
    """(kind, bound, lo, hi, entity_ids) from a reference ``Criteria`` (descriptors.py:32-103).  Only the
    descriptors the reference ships run on device; anything else raises (there is no CPU fallback)."""
    if criteria is None or criteria.cutoff is None:
        return _CRIT_NONE, _BOUND_UPPER, 0.0, 0.0, None
    name = getattr(criteria.desc, "__name__", "")
    eids = None
    if name == "com_distance_sum":
        kind = _CRIT_COM
        eids = np.ascontiguousarray(criteria.desc_kwargs["entity_ids"], dtype=np.int32)
        if eids.shape != (n_atoms,) or set(criteria.desc_kwargs) != {"entity_ids"}:
            raise NotImplementedError("com_distance_sum takes exactly entity_ids, one per fragment atom")
    elif name == "max_atom_pair_dist" and not criteria.desc_kwargs:
        kind = _CRIT_PAIR
    else:
        raise NotImplementedError(f"criteria descriptor {criteria.desc!r} has no device implementation")
    if isinstance(criteria.cutoff, list):
        return kind, _BOUND_RANGE, float(criteria.cutoff[0]), float(criteria.cutoff[1]), eids
    if criteria.bound == "upper":
        return kind, _BOUND_UPPER, 0.0, float(criteria.cutoff), eids
    return kind, _BOUND_LOWER, float(criteria.cutoff), 0.0, eids


def _handle(model):
    """Upload a reference gdmlModel once (its raw .npz arrays are ``model.model_dict``, gdml_model.py:69)."""
    cached = getattr(model, "_b200_handle", None)
    if cached is not None:
        return cached
    lib = _library()
    m = model.model_dict
    lat = np.ascontiguousarray(m["lattice"], dtype=np.float64).reshape(9) if "lattice" in m else None
    n_atoms = len(model.Z)
    R_desc = np.ascontiguousarray(m["R_desc"], dtype=np.float64)
    RdA = np.ascontiguousarray(m["R_d_desc_alpha"], dtype=np.float64)
    tpl = np.ascontiguousarray(m["tril_perms_lin"], dtype=np.int64)
    aE = np.ascontiguousarray(m["alphas_E"], dtype=np.float64) if "alphas_E" in m else None
    kind, bound, lo, hi, eids = _criteria_spec(model.criteria, n_atoms)
    d = _ModelDesc(n_atoms, len(model.comp_ids), R_desc.shape[1], int(m["perms"].shape[0]), float(m["sig"]),
                   float(m["c"]), float(m["std"]) if "std" in m else 1.0,
                   R_desc.ctypes.data_as(_dp), RdA.ctypes.data_as(_dp), tpl.ctypes.data_as(_lp),
                   aE.ctypes.data_as(_dp) if aE is not None else _dp(), kind, bound, lo, hi,
                   eids.ctypes.data_as(_ip) if eids is not None else _ip(),
                   lat.ctypes.data_as(_dp) if lat is not None else _dp())
    h = C.c_void_p()
    _check(lib, lib.mbg_model_create(_ctx, C.byref(d), C.byref(h)))
    model._b200_handle = h
    return h


def _system(Z, entity_ids, periodic_cell, keep):
    entity_ids = np.asarray(entity_ids)
    order = np.ascontiguousarray(np.argsort(entity_ids, kind="stable"), dtype=np.int32)  # np.where(ids == e)[0], all e
    offs = np.zeros(int(entity_ids.max()) + 2, dtype=np.int32)
    offs[1:] = np.cumsum(np.bincount(entity_ids))
    masses = np.array([to_mass(int(z)) for z in Z], dtype=np.float64)
    cell, cutoff, pbc = None, 0.0, None
    if periodic_cell:
        cell = np.ascontiguousarray(periodic_cell.cell_v, dtype=np.float64).reshape(9)
        cutoff = float(periodic_cell.cutoff)
        pbc = np.ascontiguousarray(np.broadcast_to(np.asarray(periodic_cell.pbc, dtype=bool), (3,)), dtype=np.int32)
    keep += [order, offs, masses, cell, pbc]
    return _SystemDesc(len(Z), len(offs) - 1, offs.ctypes.data_as(_ip), order.ctypes.data_as(_ip),
                       masses.ctypes.data_as(_dp), cell.ctypes.data_as(_dp) if cell is not None else _dp(), cutoff,
                       _dp(), _dp(), pbc.ctypes.data_as(_ip) if pbc is not None else _ip())


def _job(model, entity_combs, scalers, keep):
    order = len(model.comp_ids)
    combs = np.ascontiguousarray(np.array([tuple(c) for c in entity_combs], dtype=np.int64).reshape(-1, order),
                                 dtype=np.int32)
    a_ent = np.array([s.entity_id for s in scalers], dtype=np.int32)
    for s in scalers:
        if getattr(s.switching_func, "__name__", "") != "linear_switching":
            raise NotImplementedError("only linear_switching runs on device")
        assert 0 <= s.switching_kwargs["factor"] <= 1  # switching.py:47
    a_fac = np.array([s.switching_kwargs["factor"] for s in scalers], dtype=np.float64)
    pad = np.zeros(1, dtype=np.int32)  # a non-NULL pointer marks the explicit-tuple source even when it is empty
    keep += [combs, a_ent, a_fac, pad]
    return _Job(_handle(model), _ip(), _ip(), (combs if combs.size else pad).ctypes.data_as(_ip), len(combs),
                len(scalers), a_ent.ctypes.data_as(_ip), a_fac.ctypes.data_as(_dp)), combs


def predict_gdml_b200(Z, R, entity_ids, entity_combs, model, periodic_cell=None, **kwargs):
    """Drop-in for ``mbgdml.predictors.predict_gdml`` (gdml_predict.py:170-272)."""
    assert R.ndim == 2
    lib = _library()
    keep = []
    sysd = _system(Z, entity_ids, periodic_cell, keep)
    job, _ = _job(model, entity_combs, kwargs.get("alchemy_scalers") or [], keep)
    want_virial = bool(periodic_cell) and bool(kwargs.get("compute_virial", False))
    Rc = np.ascontiguousarray(R, dtype=np.float64)
    E, F, V = np.zeros(1), np.zeros(Rc.shape), np.zeros((3, 3))
    _check(lib, lib.mbg_predict(_ctx, C.byref(sysd), Rc.ctypes.data, 1, C.byref(job), 1,
                                _MBG_FLAG_VIRIAL if want_virial else 0, 0, 1, E.ctypes.data, F.ctypes.data,
                                V.ctypes.data, None))
    return (float(E[0]), F, V) if want_virial else (float(E[0]), F)


def predict_gdml_decomp_b200(Z, R, entity_ids, entity_combs, model, **kwargs):
    """Drop-in for ``mbgdml.predictors.predict_gdml_decomp`` (gdml_predict.py:275-364)."""
    assert R.ndim == 2
    lib = _library()
    keep = []
    sysd = _system(Z, entity_ids, None, keep)
    entity_combs = np.asarray(entity_combs)
    job, combs = _job(model, entity_combs.reshape(len(entity_combs), -1), [], keep)
    n, n_at = len(combs), len(model.Z)
    E, F = np.full(n, np.nan), np.full((n, n_at, 3), np.nan)
    if n:
        Rc = np.ascontiguousarray(R, dtype=np.float64)
        _check(lib, lib.mbg_predict_decomp(_ctx, C.byref(sysd), Rc.ctypes.data, C.byref(job), 0, E.ctypes.data,
                                           F.ctypes.data, None))
    return E, F, entity_combs
