"""Host glue between the reference-shaped Python surface and the C ABI: turns
(Z, entity_ids, models, combs, cell, scalers) into mbg_system_desc / mbg_job arrays."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native
from .descriptors import masses_of
from .utils import CombGenerator


class System:
    """mbg_system_desc plus the arrays that back its pointers."""

    def __init__(self, Z, entity_ids, periodic_cell=None, need_masses=True, frame_cells=None):
        Z = np.asarray(Z)
        entity_ids = np.asarray(entity_ids)
        if entity_ids.shape != Z.shape:
            raise ValueError("entity_ids must have one entry per atom")
        if entity_ids.size and entity_ids.min() < 0:
            raise ValueError("entity_ids must be non-negative")
        self.n_atoms = int(Z.shape[0])
        self.n_entities = int(entity_ids.max()) + 1 if entity_ids.size else 0
        # np.where(entity_ids == e)[0] for every e: a stable sort keeps atoms ascending within an entity
        # (already-sorted ids, the usual case, need no sort)
        if entity_ids.size and np.all(entity_ids[1:] >= entity_ids[:-1]):
            self.entity_atoms = np.arange(self.n_atoms, dtype=np.int32)
        else:
            self.entity_atoms = _native.as_i32(np.argsort(entity_ids, kind="stable"))
        counts = np.bincount(entity_ids, minlength=self.n_entities)
        self.entity_offsets = np.zeros(self.n_entities + 1, dtype=np.int32)
        np.cumsum(counts, out=self.entity_offsets[1:])
        self.masses = _native.as_f64(masses_of(Z, strict=need_masses))
        self.cell = None
        self.pbc = None  # int32[3] when a direction is not periodic (Cell.pbc)
        cutoff = 0.0
        if periodic_cell is not None:
            self.cell, cutoff = periodic_cell.native_cell()
            self.pbc = periodic_cell.native_pbc()
        self.frame_cells = self.frame_cutoffs = None
        if frame_cells is not None:  # one cell per frame: (cells (n, 3, 3), cutoffs (n,)[, pbc int32[3]])
            self.frame_cells = _native.as_f64(frame_cells[0]).reshape(-1, 9)
            self.frame_cutoffs = _native.as_f64(frame_cells[1]).reshape(-1)
            if len(self.frame_cells) != len(self.frame_cutoffs):
                raise ValueError("one cutoff per frame cell")
            if len(frame_cells) > 2 and frame_cells[2] is not None:  # Cell.pbc shared by all frames
                self.pbc = _native.as_i32(frame_cells[2])
        self.desc = _native.SystemDesc(
            n_atoms=self.n_atoms, n_entities=self.n_entities,
            entity_offsets=_native.ptr(self.entity_offsets, _native._ip),
            entity_atoms=_native.ptr(self.entity_atoms, _native._ip),
            masses=_native.ptr(self.masses, _native._dp),
            cell=_native.ptr(self.cell, _native._dp), cell_cutoff=cutoff,
            frame_cells=_native.ptr(self.frame_cells, _native._dp), frame_cutoffs=_native.ptr(self.frame_cutoffs, _native._dp),
            pbc=_native.ptr(self.pbc, _native._ip),
        )


_last_system = None  # (Z, entity_ids, cell key, need_masses, System) of the previous call


def cached_system(Z, entity_ids, periodic_cell, need_masses):
    """MD drivers and batched evaluations call with the same (Z, entity_ids, cell) over and over: keep the last
    ``System`` (entity CSR, masses, cell tables) and reuse it when the inputs compare equal -- an O(n_atoms)
    memcmp instead of re-deriving it."""
    global _last_system
    Z = np.asarray(Z)
    entity_ids = np.asarray(entity_ids)
    cell_key = None
    if periodic_cell is not None:
        cell, cutoff = periodic_cell.native_cell()
        pbc = periodic_cell.native_pbc()
        cell_key = (cell.tobytes(), float(cutoff), None if pbc is None else pbc.tobytes())
    hit = _last_system
    if (hit is not None and hit[2] == cell_key and hit[3] == bool(need_masses) and hit[0].shape == Z.shape
            and hit[1].shape == entity_ids.shape and np.array_equal(hit[0], Z) and np.array_equal(hit[1], entity_ids)):
        return hit[4]
    system = System(Z, entity_ids, periodic_cell, need_masses=need_masses)
    _last_system = (Z.copy(), entity_ids.copy(), cell_key, bool(need_masses), system)
    return system


def _needs_masses(models):
    from .descriptors import com_distance_sum

    return any(m.criteria is not None and m.criteria.cutoff is not None and m.criteria.desc is com_distance_sum
               for m in models)


def make_job(ctx, model, entity_combs, alchemy_scalers, keep):
    """One mbg_job.  ``entity_combs`` is a CombGenerator (device enumeration from its sets) or any
    iterable / array of entity tuples (explicit)."""
    job = _native.Job()
    job.model = model.native_handle(ctx)
    order = int(model.nbody_order)
    if isinstance(entity_combs, CombGenerator) and entity_combs._iter is None:
        sets = entity_combs.sets
        if len(sets) != order:
            raise ValueError("number of entity sets differs from the model's n-body order")
        offs = np.zeros(order + 1, dtype=np.int32)
        offs[1:] = np.cumsum([len(s) for s in sets])
        vals = _native.as_i32(np.concatenate(sets) if offs[-1] else np.zeros(0))
        keep += [offs, vals]
        job.set_offsets = _native.ptr(offs, _native._ip)
        job.set_entities = _native.ptr(vals, _native._ip)
        job.n_combs = 0
    else:
        if isinstance(entity_combs, np.ndarray):
            combs = entity_combs
        else:
            combs = np.array([tuple(c) for c in entity_combs], dtype=np.int64)
        combs = _native.as_i32(combs.reshape(-1, order))
        keep.append(combs)
        # a non-NULL pointer marks the explicit source even when it is empty
        job.combs = _native.ptr(combs if combs.size else np.zeros(1, dtype=np.int32), _native._ip)
        if not combs.size:
            keep.append(np.zeros(1, dtype=np.int32))
            job.combs = _native.ptr(keep[-1], _native._ip)
        job.n_combs = combs.shape[0]
    scalers = list(alchemy_scalers or [])
    if scalers:
        ents = _native.as_i32([s.entity_id for s in scalers])
        facs = _native.as_f64([s.native_factor() for s in scalers])
        keep += [ents, facs]
        job.n_alchemy = len(scalers)
        job.alchemy_entity = _native.ptr(ents, _native._ip)
        job.alchemy_factor = _native.ptr(facs, _native._dp)
    return job


def _same_array(kept, new):
    """Content comparison at memcmp speed (``kept`` is the cache's own copy, so in-place edits of the caller's array are
    seen)."""
    if kept.shape != new.shape or kept.dtype != new.dtype:
        return False
    if new.flags.c_contiguous and kept.dtype.kind in "iufUS" and kept.size:
        return bool(np.array_equal(kept.reshape(-1).view(np.uint8), new.reshape(-1).view(np.uint8)))
    return bool(np.array_equal(kept, new))


class StepCache:
    """Prepared steps of one predictor.  A step is built the second time a description is seen (a one-off evaluation
    does not pay for the capture) and at most ``capacity`` of them are kept.  What a step fixes: system (Z, entity_ids,
    comp_ids), models (+ criteria), scalers, frame count, kind of periodicity, virial flag -- not the cell values."""

    def __init__(self, capacity=4):
        self.capacity, self.entries = capacity, []

    @staticmethod
    def small_key(models, scalers_per_model, n_frames, periodic_cell, frame_cells, compute_virial):
        cell_kind = "frames" if frame_cells is not None else ("cell" if periodic_cell is not None else None)
        if cell_kind == "cell" and periodic_cell.native_pbc() is not None:  # the step keeps Cell.pbc from its creation
            cell_kind = ("cell", periodic_cell.native_pbc().tobytes())
        return (tuple((m._uid, m._criteria_key()) for m in models),
                tuple(tuple((s.entity_id, s.order, s.native_factor()) for s in (sc or [])) for sc in scalers_per_model),
                int(n_frames), cell_kind, bool(compute_virial))

    def get(self, Z, entity_ids, comp_ids, small, build):
        Z, entity_ids, comp_ids = np.asarray(Z), np.asarray(entity_ids), np.asarray(comp_ids)
        for e in self.entries:
            if e["small"] == small and _same_array(e["Z"], Z) and _same_array(e["eids"], entity_ids) \
                    and _same_array(e["cids"], comp_ids):
                if e["step"] is None:
                    e["step"] = build()  # second sighting
                    live = [x for x in self.entries if x["step"] is not None]
                    if len(live) > self.capacity:
                        live[0]["step"].close()
                        self.entries.remove(live[0])
                return e["step"]
        self.entries.append({"small": small, "Z": np.ascontiguousarray(Z).copy(), "eids": np.ascontiguousarray(entity_ids).copy(),
                             "cids": np.ascontiguousarray(comp_ids).copy(), "step": None})
        if len(self.entries) > 16:
            old = self.entries.pop(0)
            if old["step"] is not None:
                old["step"].close()
        return None

    @property
    def steps(self):
        return [e["step"] for e in self.entries if e["step"] is not None]

    def clear(self):
        for st in self.steps:
            st.close()
        self.entries = []


def make_step(ctx, Z, entity_ids, models, comb_sources, scalers_per_model, periodic_cell, n_frames, compute_virial=False,
              frame_cells=None, shard=(0, 1)):
    """mbg_step_create for (system, models, scalers, frame count): see include/mbgdml_b200.h."""
    if frame_cells is not None:
        system = System(Z, entity_ids, None, need_masses=_needs_masses(models), frame_cells=frame_cells)
    else:
        system = System(Z, entity_ids, periodic_cell, need_masses=_needs_masses(models))
    keep = []
    jobs = (_native.Job * len(models))()
    for k, (m, src, sc) in enumerate(zip(models, comb_sources, scalers_per_model)):
        jobs[k] = make_job(ctx, m, src, sc, keep)
    return _native.Step(ctx, system.desc, int(n_frames), system.n_atoms, jobs, len(models), bool(compute_virial), shard=shard)


def predict_total(ctx, Z, R, entity_ids, models, comb_sources, scalers_per_model, periodic_cell,
                  compute_virial=False, shard=(0, 1), device_out=None, frame_cells=None):
    """All frames x all jobs in one native call.  Returns (E, F, virial-or-None, stats).

    ``R`` is a host array, or a contiguous float64 torch CUDA tensor (n_frames, n_atoms, 3) that
    is already resident in HBM (MBG_FLAG_R_ON_DEVICE)."""
    r_on_device = hasattr(R, "data_ptr")
    if r_on_device:
        assert R.is_cuda and R.is_contiguous() and R.dim() == 3 and R.element_size() == 8
        n_frames, n_atoms = int(R.shape[0]), int(R.shape[1])
        r_ptr = R.data_ptr()
    else:
        R = _native.as_f64(R)
        if R.ndim == 2:
            R = R[None]
        n_frames, n_atoms = R.shape[0], R.shape[1]
        r_ptr = R.ctypes.data
    if frame_cells is not None:
        system = System(Z, entity_ids, None, need_masses=_needs_masses(models), frame_cells=frame_cells)
        if len(system.frame_cells) != n_frames:
            raise ValueError("one cell per frame")
    else:
        system = cached_system(Z, entity_ids, periodic_cell, _needs_masses(models))
    if system.n_atoms != n_atoms:
        raise ValueError("R and Z disagree on the number of atoms")
    keep = []
    jobs = (_native.Job * len(models))()
    for k, (m, src, sc) in enumerate(zip(models, comb_sources, scalers_per_model)):
        jobs[k] = make_job(ctx, m, src, sc, keep)
    stats = (_native.JobStats * len(models))()
    flags = _native.FLAG_VIRIAL if compute_virial else 0
    if r_on_device:
        flags |= _native.FLAG_R_ON_DEVICE
    if device_out is not None:  # torch CUDA tensors (E, F, virial) for the multi-GPU path
        E_t, F_t, V_t = device_out
        flags |= _native.FLAG_OUT_ON_DEVICE
        _native.check(ctx.lib.mbg_predict(
            ctx.handle, C.byref(system.desc), r_ptr, n_frames, jobs, len(models), flags, shard[0], shard[1],
            E_t.data_ptr(), F_t.data_ptr(), V_t.data_ptr() if V_t is not None else None, stats))
        return None, None, None, [(_s.n_candidates, _s.n_enumerated, _s.n_accepted) for _s in stats]
    E = np.zeros(n_frames)
    F = np.zeros((n_frames, n_atoms, 3))
    V = np.zeros((n_frames, 3, 3)) if compute_virial else None
    _native.check(ctx.lib.mbg_predict(
        ctx.handle, C.byref(system.desc), r_ptr, n_frames, jobs, len(models), flags, shard[0], shard[1],
        E.ctypes.data, F.ctypes.data, V.ctypes.data if V is not None else None, stats))
    return E, F, V, [(_s.n_candidates, _s.n_enumerated, _s.n_accepted) for _s in stats]


def predict_decomp(ctx, Z, R, entity_ids, model, entity_combs, alchemy_scalers=None):
    """One frame, one model, explicit combs -> per-fragment rows (NaN when rejected)."""
    R = _native.as_f64(R)
    system = System(Z, entity_ids, None, need_masses=_needs_masses([model]))
    keep = []
    combs = np.asarray(entity_combs)
    job = make_job(ctx, model, combs.reshape(-1, int(model.nbody_order)), alchemy_scalers, keep)
    n = int(job.n_combs)
    n_frag_atoms = int(model.Z.shape[0])
    E = np.full(n, np.nan)
    F = np.full((n, n_frag_atoms, 3), np.nan)
    st = _native.JobStats()
    if n:
        _native.check(ctx.lib.mbg_predict_decomp(ctx.handle, C.byref(system.desc), R.ctypes.data, C.byref(job), 0,
                                                 E.ctypes.data, F.ctypes.data, C.byref(st)))
    return E, F


def accept_mask(ctx, Z, R, entity_ids, model, entity_combs, periodic_cell=None):
    """Accept decision of every candidate fragment (0 not ascending / 1 accepted / 2 rejected), all frames:
    uint8 array (n_frames, candidates per frame) in the job's candidate order (gdml_predict.py:214-235)."""
    R = _native.as_f64(R)
    if R.ndim == 2:
        R = R[None]
    system = cached_system(Z, entity_ids, periodic_cell, _needs_masses([model]))
    if system.n_atoms != R.shape[1]:
        raise ValueError("R and Z disagree on the number of atoms")
    keep = []
    job = make_job(ctx, model, entity_combs, None, keep)
    n = C.c_int64(0)
    _native.check(ctx.lib.mbg_accept_mask(ctx.handle, C.byref(system.desc), R.ctypes.data, R.shape[0], C.byref(job), 0,
                                          None, C.byref(n)))
    mask = np.zeros((R.shape[0], n.value), dtype=np.uint8)
    if mask.size:
        _native.check(ctx.lib.mbg_accept_mask(ctx.handle, C.byref(system.desc), R.ctypes.data, R.shape[0], C.byref(job),
                                              0, mask.ctypes.data, C.byref(n)))
    return mask
