"""Many-body expansion driver (reference mbgdml/mbe.py:511-1181).

``mbePredict(models, predict_model, ...)`` keeps the reference's constructor, public
attributes and methods.  When ``predict_model`` is this package's ``predict_gdml`` the
whole ``R`` batch x all models goes to the GPU in ONE native call (fragment
enumeration, culling, contraction and scatter all on device); any other callable is
driven per frame x model exactly as the reference does (mbe.py:1048-1058).

Multi-GPU: one process per GPU, opt-in (``pred.shard_group = torch.distributed.group.WORLD``):
every rank of the NCCL group evaluates its block-cyclic share of the fragments and the per-atom
forces, energies and virials are summed with a single all-reduce.
"""
from __future__ import annotations

import itertools

import numpy as np

from . import _engine, _native
from .predictors import predict_gdml, predict_gdml_decomp
from .stress import to_voigt, virial_finite_diff
from .utils import gen_combs


def gen_r_entity_combs(r_prov_spec, entity_ids_lower):
    """mbe.py:36-70: entity-id combinations (without replacement) of one structure's provenance spec."""
    n_entities_lower = len(set(np.asarray(entity_ids_lower).tolist()))
    for comb in itertools.combinations(r_prov_spec[2:], n_entities_lower):
        yield comb


def gen_r_idxs_worker(r_prov_specs, r_prov_ids_lower, n_workers):
    """mbe.py:200-233: structure indices whose r_prov_id appears in the lower-order data, split over workers."""
    r_idxs = []
    for r_prov_id_lower in r_prov_ids_lower.keys():
        r_idxs.extend(np.where(r_prov_specs[:, 0] == r_prov_id_lower)[0].tolist())
    task_size = int(np.ceil(len(r_idxs) / n_workers)) if r_idxs else 1
    for i in range(0, len(r_idxs), task_size):
        yield r_idxs[i:i + task_size]


def _mbe_pairs(entity_ids, r_prov_specs, r_idxs, entity_ids_lower, r_prov_specs_lower):
    """Host side of mbe_worker (mbe.py:73-197): which lower-order row belongs to which (structure, entity
    combination).  The reference scans all lower rows per combination (np.all(...) , mbe.py:154-156: O(n^2));
    this is a sort join.  Returns (pair_offsets, lower_idx, slot_entity) in the reference's loop order:
    structures in r_idxs order, combinations in itertools.combinations order, missing fragments skipped."""
    specs = np.asarray(r_prov_specs)[np.asarray(r_idxs, dtype=np.int64)].astype(np.int64)
    lower = np.asarray(r_prov_specs_lower).astype(np.int64)
    n_sel, n_ent = specs.shape[0], specs.shape[1] - 2
    n_low = len(set(np.asarray(entity_ids_lower).tolist()))
    if lower.shape[1] != 2 + n_low:
        raise ValueError("r_prov_specs_lower has the wrong number of columns for entity_ids_lower")
    pos = np.array(list(itertools.combinations(range(n_ent), n_low)), dtype=np.int64).reshape(-1, n_low)
    n_comb = pos.shape[0]
    ents = specs[:, 2:][:, pos]  # (n_sel, n_comb, n_low) provenance entity ids of every combination
    keys = np.concatenate([np.broadcast_to(specs[:, None, :2], (n_sel, n_comb, 2)), ents], axis=2).reshape(-1, 2 + n_low)
    # first occurrence of every key among the lower rows (np.where(...)[0][0])
    both = np.concatenate([lower, keys], axis=0)
    _, inv = np.unique(both, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    first = np.full(inv.max() + 1 if inv.size else 0, -1, dtype=np.int64)
    n_lower_rows = lower.shape[0]
    first[inv[:n_lower_rows][::-1]] = np.arange(n_lower_rows - 1, -1, -1)
    match = first[inv[n_lower_rows:]].reshape(n_sel, n_comb)
    found = match >= 0
    # entity id (position in the structure's spec) each slot lands on: first occurrence of the provenance id
    # (np.where(r_prov_spec[2:] == entity_id_prov)[0][0], mbe.py:183)
    slot_entity = (specs[:, None, None, 2:] == ents[..., None]).argmax(axis=-1)
    pair_offsets = np.zeros(n_sel + 1, dtype=np.int64)
    np.cumsum(found.sum(axis=1), out=pair_offsets[1:])
    return pair_offsets, match[found], slot_entity[found].astype(np.int32)


def mbe_worker(r_shape, entity_ids, r_prov_specs, r_idxs, E_lower, Deriv_lower, entity_ids_lower, r_prov_specs_lower):
    """mbe.py:73-197: summed lower-order contributions of the structures ``r_idxs`` (no sign applied)."""
    E = np.zeros(len(r_idxs))
    Deriv = np.zeros((len(r_idxs), *r_shape))
    _mbe_accumulate(E, Deriv, np.arange(len(r_idxs)), entity_ids, r_prov_specs, r_idxs, E_lower, Deriv_lower,
                    entity_ids_lower, r_prov_specs_lower, 1.0)
    return r_idxs, E, Deriv


def _mbe_accumulate(E, Deriv, rows, entity_ids, r_prov_specs, r_idxs, E_lower, Deriv_lower, entity_ids_lower,
                    r_prov_specs_lower, sign):
    entity_ids = np.asarray(entity_ids)
    entity_ids_lower = np.asarray(entity_ids_lower)
    if len(r_idxs) == 0:
        return
    pair_offsets, lower_idx, slot_entity = _mbe_pairs(entity_ids, r_prov_specs, r_idxs, entity_ids_lower,
                                                      r_prov_specs_lower)
    # atoms of the reference structure: entity id and index within the entity (np.where(entity_ids == e)[0] order)
    atom_rank = np.zeros(len(entity_ids), dtype=np.int32)
    for e in np.unique(entity_ids):
        idx = np.where(entity_ids == e)[0]
        atom_rank[idx] = np.arange(len(idx))
    n_low = len(set(entity_ids_lower.tolist()))
    lower_atoms = [np.where(entity_ids_lower == k)[0] for k in range(n_low)]
    lower_off = np.zeros(n_low + 1, dtype=np.int32)
    lower_off[1:] = np.cumsum([len(x) for x in lower_atoms])
    _native.get_context().mbe_accumulate(
        rows, pair_offsets, lower_idx, slot_entity, entity_ids, atom_rank, lower_off,
        np.concatenate(lower_atoms) if lower_atoms else np.zeros(0), E_lower, Deriv_lower, sign, E, Deriv)


def mbe_contrib(E, Deriv, entity_ids, r_prov_ids, r_prov_specs, E_lower, Deriv_lower, entity_ids_lower,
                r_prov_ids_lower, r_prov_specs_lower, operation="remove", use_ray=False, ray_address="auto",
                n_workers=1):
    """Add or remove lower-order energy / derivative contributions (mbe.py:237-429).  ``E`` and ``Deriv`` are
    updated in place and returned, like the reference.  The matching of lower-order rows is a host sort join;
    the accumulation is a device segmented scatter in the reference's summation order (bit-identical)."""
    if use_ray:
        raise NotImplementedError("use_ray=True: the ray CPU fan-out (mbe.py:399-427) is replaced by one GPU kernel")
    if operation not in ("add", "remove"):
        raise ValueError(f'{operation} is not "add" or "remove"')
    if (r_prov_ids is not None) and (r_prov_ids != {}):
        for r_prov_id_lower, r_prov_md5_lower in r_prov_ids_lower.items():
            assert r_prov_md5_lower == r_prov_ids[r_prov_id_lower]
    else:  # assume that all lower models apply (mbe.py:367-377)
        assert r_prov_specs is None or r_prov_specs.shape == (1, 0)
        assert len(set(np.asarray(r_prov_specs_lower)[:, 0].tolist())) == 1
        n_r = len(E)
        n_entities = len(set(np.asarray(entity_ids).tolist()))
        r_prov_specs = np.empty((n_r, 2 + n_entities), dtype=int)
        r_prov_specs[:, 0] = np.asarray(r_prov_specs_lower)[0][0]
        r_prov_specs[:, 1] = np.arange(n_r)
        for entity_id in range(n_entities):
            r_prov_specs[:, entity_id + 2] = entity_id
    r_idxs = next(gen_r_idxs_worker(r_prov_specs, r_prov_ids_lower, 1), [])
    if not (isinstance(E, np.ndarray) and E.dtype == np.float64 and E.flags.c_contiguous
            and isinstance(Deriv, np.ndarray) and Deriv.dtype == np.float64 and Deriv.flags.c_contiguous):
        raise TypeError("E and Deriv must be C-contiguous float64 arrays (they are updated in place)")
    _mbe_accumulate(E, Deriv, np.asarray(r_idxs, dtype=np.int64), entity_ids, r_prov_specs, r_idxs, E_lower, Deriv_lower,
                    entity_ids_lower, r_prov_specs_lower, 1.0 if operation == "add" else -1.0)
    return E, Deriv


def decomp_to_total(E_nbody, F_nbody, entity_ids, entity_combs, use_ray=False, n_workers=1, ray_address="auto"):
    """Sum decomposed n-body energies and forces into totals (mbe.py:432-508); NaN rows count as zero."""
    entity_ids = np.asarray(entity_ids)
    entity_combs = np.asarray(entity_combs)
    n_atoms = len(entity_ids)
    n_r = len(np.unique(entity_combs[:, 0]))
    E = np.zeros((n_r,))
    F = np.zeros((n_r, n_atoms, 3))
    entity_ids_lower = []
    for i in range(len(entity_combs[0][1:])):
        entity_ids_lower.extend([i for j in range(np.count_nonzero(entity_ids == i))])
    entity_ids_lower = np.array(entity_ids_lower)
    r_prov_specs_lower = np.hstack((np.zeros((len(entity_combs), 1)), entity_combs))
    E_nbody = np.nan_to_num(E_nbody, copy=False, nan=0.0)
    F_nbody = np.nan_to_num(F_nbody, copy=False, nan=0.0)
    return mbe_contrib(E, F, entity_ids, None, None, E_nbody, F_nbody, entity_ids_lower, {0: ""}, r_prov_specs_lower,
                       operation="add", use_ray=use_ray, n_workers=n_workers, ray_address=ray_address)


def _shard_group(requested):
    """The process group a prediction is sharded over, or None for a purely local evaluation (the default, like
    the reference's ``mbePredict.predict``).  Sharding is OPT-IN -- ``pred.shard_group = dist.group.WORLD`` (or the
    environment variable MBGDML_B200_SHARD=1 for the default group) -- because it is an SPMD contract: every rank of
    the group must call ``predict`` with the same (Z, R, entity_ids, comp_ids) the same number of times."""
    import os

    if requested is None and os.environ.get("MBGDML_B200_SHARD", "0") not in ("1", "true", "True"):
        return None
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        if requested is not None:
            raise RuntimeError("shard_group is set but torch.distributed is not initialised")
        return None
    group = dist.group.WORLD if requested is None or requested is True else requested
    if dist.get_world_size(group) <= 1:
        return None
    if dist.get_backend(group) != "nccl":
        raise RuntimeError("fragment sharding all-reduces a CUDA buffer: the process group must use the NCCL backend")
    return group


class mbePredict:  # pylint: disable=invalid-name
    """Predict energies and forces of structures with many-body GDML models on B200 GPUs."""

    def __init__(
        self,
        models,
        predict_model,
        use_ray=False,
        n_workers=1,
        ray_address="auto",
        wkr_chunk_size=None,
        alchemy_scalers=None,
        periodic_cell=None,
        compute_stress=False,
    ):
        if use_ray:
            raise NotImplementedError(
                "use_ray=True: the ray CPU fan-out (mbe.py:814-868) is replaced by GPU kernels; "
                "launch one process per GPU (torchrun) for multi-GPU runs"
            )
        self.models = models
        self.predict_model = predict_model
        self.use_ray = False
        self.n_workers = n_workers
        self.ray_address = ray_address
        self.wkr_chunk_size = wkr_chunk_size
        self.alchemy_scalers = [] if alchemy_scalers is None else alchemy_scalers
        self.periodic_cell = periodic_cell
        self.compute_stress = compute_stress
        self.virial_form = "group"
        self.finite_diff_dh = 1e-4
        self.use_voigt = False
        self.only_normal_stress = False
        self.box_scaling_type = "anisotropic"
        self.last_stats = None
        self.shard_group = None  # opt-in multi-GPU fragment sharding (see _shard_group)
        # Latency mode (MD: interfaces/ase.py:65-108 calls predict once per integrator step with the same system):
        # from the second evaluation of the same (system, models, frame count) on, the whole call is one replay of
        # a captured CUDA graph (mbg_step_create / mbg_step_run).  Set to False to always take the plain call.
        self.use_graph = True
        self._steps = _engine.StepCache()

    @property
    def periodic_cell(self):
        return getattr(self, "_periodic_cell", None)

    @periodic_cell.setter
    def periodic_cell(self, var):
        self._periodic_cell = var

    def get_avail_entities(self, comp_ids_r, comp_ids_model):
        """mbe.py:713-736: entity ids of the structure that carry each model component id."""
        comp_ids_r = np.asarray(comp_ids_r)
        return [np.where(comp_id == comp_ids_r)[0] for comp_id in comp_ids_model]

    def _scalers_for(self, model):
        order = len(model.comp_ids)
        return [s for s in (self.alchemy_scalers or []) if s.order == order]  # mbe.py:789-796

    # ------------------------------------------------------------------------------
    def compute_nbody(self, Z, R, entity_ids, comp_ids, model):
        """All n-body contributions of one structure with one model (mbe.py:739-875)."""
        R = np.asarray(R)
        if R.ndim == 3:
            if R.shape[0] == 1:
                R = R[0]
            else:
                raise ValueError("R.ndim is not 2 (only one structure is allowed)")
        avail_entity_ids = self.get_avail_entities(comp_ids, model.comp_ids)
        nbody_gen = gen_combs(avail_entity_ids)
        kwargs_pred = {"alchemy_scalers": self._scalers_for(model)}
        periodic_cell = self.periodic_cell
        compute_stress = self.compute_stress and bool(periodic_cell) and self.virial_form == "group"
        if compute_stress:
            kwargs_pred["compute_virial"] = True
        E, F, *virial_model = self.predict_model(Z, R, entity_ids, nbody_gen, model, periodic_cell, **kwargs_pred)
        if compute_stress:
            return E, F, virial_model[0]
        return E, F

    def compute_nbody_decomp(self, Z, R, entity_ids, comp_ids, model):
        """Per-fragment n-body energies and forces of one structure (mbe.py:878-1007)."""
        R = np.asarray(R)
        if R.ndim == 3:
            if R.shape[0] == 1:
                R = R[0]
            else:
                raise ValueError("R.ndim is not 2 (only one structure is allowed)")
        avail_entity_ids = self.get_avail_entities(comp_ids, model.comp_ids)
        entity_combs = gen_combs(avail_entity_ids).to_array()  # enumerated on device, reference order
        E, F, entity_combs = self.predict_model(Z, R, entity_ids, entity_combs, model)
        return E, F, entity_combs

    # ------------------------------------------------------------------------------
    def _predict_batched(self, Z, R, entity_ids, comp_ids, want_virial, frame_cells=None):
        """Fast path: every frame and every model in one native call (+ one all-reduce).  ``frame_cells`` =
        (cells (n_frames, 3, 3), cutoffs (n_frames,)) evaluates every frame under its own periodic cell."""
        ctx = _native.get_context()
        scalers = [self._scalers_for(m) for m in self.models]
        cell = None if frame_cells is not None else self.periodic_cell

        def sources():  # one gen_combs per model: the device enumerates from its sets
            return [gen_combs(self.get_avail_entities(comp_ids, m.comp_ids)) for m in self.models]

        group = _shard_group(self.shard_group)
        if group is None:
            step = None
            if self.use_graph and not hasattr(R, "data_ptr"):
                try:
                    small = _engine.StepCache.small_key(self.models, scalers, R.shape[0], cell, frame_cells, want_virial)
                    step = self._steps.get(Z, entity_ids, comp_ids, small, lambda: _engine.make_step(
                        ctx, Z, entity_ids, self.models, sources(), scalers, cell, R.shape[0], compute_virial=want_virial,
                        frame_cells=frame_cells))
                except NotImplementedError:  # e.g. the FMA-pipe / wave-launched kernels are selected
                    self.use_graph, step = False, None
            if step is not None:
                if frame_cells is not None:
                    cells, cutoffs = frame_cells
                elif cell is not None:
                    cells, cutoffs = cell.native_cell()
                    cutoffs = np.array([cutoffs])
                else:
                    cells = cutoffs = None
                E, F, V, stats = step.run(R, cells, cutoffs)
            else:
                E, F, V, stats = _engine.predict_total(ctx, Z, R, entity_ids, self.models, sources(), scalers, cell,
                                                       compute_virial=want_virial, frame_cells=frame_cells)
            self.last_stats = stats
            return E, F, V
        import torch
        import torch.distributed as dist

        dev = torch.device("cuda", ctx.device)
        n_frames, n_atoms = R.shape[0], R.shape[1]
        # one flat buffer [E | F | virial] so that a single all-reduce carries everything; mbg_predict zeroes it on
        # the stream it runs on (torch.empty: no fill kernel of ours on a stream the library does not order against)
        n_E, n_F, n_V = n_frames, n_frames * n_atoms * 3, n_frames * 9 if want_virial else 0
        buf = torch.empty(n_E + n_F + n_V, dtype=torch.float64, device=dev)
        E_t, F_t = buf[:n_E], buf[n_E:n_E + n_F]
        V_t = buf[n_E + n_F:] if want_virial else None
        # run the library on torch's current stream (NCCL orders the all-reduce after it); the legacy default stream
        # is addressed by its explicit handle cudaStreamLegacy = 0x1, a NULL handle would select the context's own
        cur = torch.cuda.current_stream(dev).cuda_stream
        prev = ctx.swap_stream(cur if cur != 0 else 1)
        try:
            _, _, _, stats = _engine.predict_total(ctx, Z, R, entity_ids, self.models, sources(), scalers,
                                                   cell, compute_virial=want_virial,
                                                   shard=(dist.get_rank(group), dist.get_world_size(group)),
                                                   device_out=(E_t, F_t, V_t), frame_cells=frame_cells)
        finally:
            ctx.swap_stream(prev)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        host = buf.cpu().numpy()
        self.last_stats = stats
        E = host[:n_E].copy()
        F = host[n_E:n_E + n_F].reshape(n_frames, n_atoms, 3).copy()
        V = host[n_E + n_F:].reshape(n_frames, 3, 3).copy() if want_virial else None
        return E, F, V

    def predict(self, Z, R, entity_ids, comp_ids):
        """Energies and forces (and stress) of one or several structures (mbe.py:1009-1104)."""
        R = np.asarray(R, dtype=np.float64)
        if R.ndim == 2:
            R = np.array([R])
        n_R = R.shape[0]
        compute_stress = self.compute_stress and bool(self.periodic_cell)
        if compute_stress:
            assert self.virial_form in ("group", "finite_diff")
        group_virial = compute_stress and self.virial_form == "group"

        if self.predict_model is predict_gdml:
            E, F, V = self._predict_batched(Z, R, entity_ids, comp_ids, group_virial)
            stress = V if group_virial else (np.zeros((n_R, 3, 3)) if compute_stress else None)
        else:
            E = np.zeros((n_R,), dtype=np.float64)
            F = np.zeros(R.shape, dtype=np.float64)
            stress = np.zeros((n_R, 3, 3), dtype=np.float64) if compute_stress else None
            for i, r in enumerate(R):
                for model in self.models:
                    E_nbody, F_nbody, *virial_nbody = self.compute_nbody(Z, r, entity_ids, comp_ids, model)
                    E[i] += E_nbody
                    F[i] += F_nbody
                    if group_virial:
                        stress[i] += virial_nbody[0]

        if compute_stress and self.virial_form == "finite_diff":
            cell_v = self.periodic_cell.cell_v
            for i, r in enumerate(R):
                # like the reference (mbe.py:1065-1073) the call does not forward self.finite_diff_dh: dh stays 1e-4
                stress[i] = virial_finite_diff(Z, r, entity_ids, comp_ids, cell_v, self,
                                               only_normal_stress=self.only_normal_stress)

        if compute_stress:
            stress = stress / self.periodic_cell.volume  # mbe.py:1079
            if self.only_normal_stress:  # mbe.py:1081-1091
                for n in range(stress.shape[0]):
                    for i in range(3):
                        for j in range(3):
                            if i != j:
                                stress[n][i][j] = 0.0
                    if self.box_scaling_type == "isotropic":
                        stress_average = np.trace(stress[n]) / 3
                        for i in range(3):
                            stress[n][i][i] = stress_average
            if self.use_voigt:
                stress = np.array([to_voigt(r_stress) for r_stress in stress])
            return E, F, stress
        return E, F

    def predict_decomp(self, Z, R, entity_ids, comp_ids):
        """Per-fragment predictions grouped by n-body order (mbe.py:1106-1181)."""
        R = np.asarray(R, dtype=np.float64)
        if R.ndim == 2:
            R = np.array([R])
        E_data, F_data, entity_comb_data = [], [], []
        model_nbody_orders = [model.nbody_order for model in self.models]
        nbody_orders = sorted(set(model_nbody_orders))
        for nbody_order in nbody_orders:
            E_nbody, F_nbody, entity_comb_nbody = [], [], []
            for model in (m for m in self.models if m.nbody_order == nbody_order):
                for i, r in enumerate(R):
                    E, F, entity_combs = self.compute_nbody_decomp(Z, r, entity_ids, comp_ids, model)
                    entity_combs = np.hstack((np.full((len(entity_combs), 1), i, dtype=np.int64), entity_combs))
                    E_nbody.extend(E)
                    F_nbody.extend(F)
                    entity_comb_nbody.extend(entity_combs)
            E_data.append(np.array(E_nbody))
            F_data.append(np.array(F_nbody))
            entity_comb_data.append(np.array(entity_comb_nbody))
        return E_data, F_data, entity_comb_data, nbody_orders


__all__ = ["mbePredict", "predict_gdml", "predict_gdml_decomp", "gen_combs", "mbe_contrib", "decomp_to_total",
           "mbe_worker", "gen_r_entity_combs", "gen_r_idxs_worker"]
