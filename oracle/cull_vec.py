"""Vectorised restatement of the per-fragment accept decision (slice -> MIC -> pair cut-off -> criteria).

TEST INFRASTRUCTURE ONLY (tests/, bench.py's CPU legs).  Same arithmetic, operation by operation, as the scalar
path of ``mbgdml_oracle._fragment`` (reference gdml_predict.py:214-235, periodic.py:61-107 + ase find_mic,
descriptors.py:65-201), evaluated for all candidate tuples of a frame at once so that FULL-frame accept masks
(447,580 candidate trimers of the 140-H2O box) can be compared bit for bit with the device's.  Elementwise
IEEE operations give the same bits whether NumPy applies them to one fragment or to a batch; the two places
where the scalar path calls LAPACK/BLAS (``solve(cell.T, v.T)`` and ``f @ cell``) are reproduced with the same
calls on the stacked vectors.  ``tests/test_oracle_cull_vec.py`` checks mask == scalar oracle on samples.
"""
from __future__ import annotations

import itertools

import numpy as np

from . import mbgdml_oracle as orc

ase_mic = orc.ase_mic


def atom_table(entity_ids, combs):
    """(n_comb, n_frag_atoms) atom indices: r_slice of every tuple (gdml_predict.py:214-216).  All entities of a
    slot must have the same atom count (one model = one fragment composition)."""
    entity_ids = np.asarray(entity_ids)
    combs = np.asarray(combs, dtype=np.int64)
    n_ent = int(entity_ids.max()) + 1
    order = np.argsort(entity_ids, kind="stable")
    counts = np.bincount(entity_ids, minlength=n_ent)
    starts = np.concatenate(([0], np.cumsum(counts)[:-1]))
    cols = []
    for s in range(combs.shape[1]):
        sizes = counts[combs[:, s]]
        if len(sizes) and not np.all(sizes == sizes[0]):
            raise ValueError("entities of one slot differ in size")
        k = int(sizes[0]) if len(sizes) else 0
        cols.append(order[starts[combs[:, s]][:, None] + np.arange(k)[None, :]])
    return np.concatenate(cols, axis=1) if cols else np.zeros((len(combs), 0), dtype=np.int64)


def _norm_rows(v):
    """np.linalg.norm(v, axis=-1) for 3-vectors: sqrt((x*x + y*y) + z*z) (add.reduce of 3 elements is sequential)."""
    return np.sqrt((v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2])


def _pdist_max_ge(r, cutoff):
    """any(pdist(r) >= cutoff) per fragment; r (n, NA, 3).  pdist arithmetic: orc._pdist_euclid."""
    n, na, _ = r.shape
    bad = np.zeros(n, dtype=bool)
    for i in range(na):
        for j in range(i + 1, na):
            d = r[:, i] - r[:, j]
            bad |= np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]) >= cutoff
    return bad


def find_mic_batch(d, cell, pbc=True):
    """ase.geometry.find_mic (oracle/ase_mic.py:find_mic) applied independently to each fragment's vector set d[k]
    (n, NA, 3).  Returns the minimum-image vectors (n, NA, 3).  With a non-periodic direction the naive form is
    never tried (find_mic: dim < 3) and the image search runs over the periodic directions only."""
    cell = np.asarray(cell, dtype=float)
    pbc = np.broadcast_to(np.asarray(pbc, dtype=bool), (3,))
    n, na, _ = d.shape
    if not pbc.any():
        return d.copy()
    if pbc.all():
        flat = d.reshape(-1, 3)
        f = np.linalg.solve(cell.T, flat.T).T
        f -= np.floor(f + 0.5)
        vmin = (f @ cell).reshape(n, na, 3)
        vlen = _norm_rows(vmin)
        half = 0.5 * min(np.linalg.norm(cell, axis=1))
        naive_ok = (vlen < half).all(axis=1)
    else:
        vmin = np.zeros_like(d)
        naive_ok = np.zeros(n, dtype=bool)
    redo = np.where(~naive_ok)[0]
    if redo.size:
        rcell, _ = ase_mic.minkowski_reduce(cell, pbc)
        hkls = np.array([(0, 0, 0)] + list(itertools.product(*[np.arange(-1 * p, p + 1) for p in pbc.astype(int)])))
        vrvecs = hkls @ rcell
        for lo in range(0, redo.size, 20000):  # 28 images x 20000 x NA x 3 doubles per chunk
            sel = redo[lo:lo + 20000]
            v = d[sel].reshape(-1, 3)
            pos = ase_mic.wrap_positions_eps0(v, rcell, pbc)
            x = pos[None, :, :] + vrvecs[:, None, :]
            lengths = _norm_rows(x)
            idx = np.argmin(lengths, axis=0)
            vmin[sel] = x[idx, np.arange(len(pos)), :].reshape(len(sel), na, 3)
    return vmin


def com_distance_sum_batch(masses, r, groups):
    """descriptors.py:159-201 on (n, NA, 3): sum over entity groups (ascending label) of |CM_total - CM_group|;
    centres of mass as orc.center_of_mass (sequential sums over atoms, one division)."""
    n, na, _ = r.shape

    def com(idx):
        num = np.zeros((n, 3))
        den = 0.0
        for a in idx:
            num = num + r[:, a, :] * masses[a]
            den = den + masses[a]
        return num / den

    cm_all = com(range(na))
    out = np.zeros(n)
    for g in sorted(set(np.asarray(groups).tolist())):
        dvec = cm_all - com(np.where(np.asarray(groups) == g)[0])
        out += np.sqrt((dvec[:, 0] * dvec[:, 0] + dvec[:, 1] * dvec[:, 1]) + dvec[:, 2] * dvec[:, 2])
    return out


def accept_mask(Z, R, entity_ids, combs, model, periodic_cell=None, return_coords=False):
    """Accept decision of every tuple in ``combs`` (n, order) for ONE frame R (n_atoms, 3).  ``model`` is an
    ``orc.Model`` (its ``criteria`` decides); ``periodic_cell`` an ``orc.Cell`` or None."""
    Z = np.asarray(Z)
    R = np.asarray(R, dtype=np.float64)
    combs = np.asarray(combs, dtype=np.int64).reshape(-1, len(model.comp_ids))
    atoms = atom_table(entity_ids, combs)
    r = R[atoms]  # (n, NA, 3)
    ok = np.ones(len(combs), dtype=bool)
    if periodic_cell is not None:
        r = find_mic_batch(r - r[:, :1, :], periodic_cell.cell_v, periodic_cell.pbc)
        ok &= ~_pdist_max_ge(r, periodic_cell.cutoff)
    crit = model.criteria
    if crit is not None and crit.cutoff is not None and len(combs):
        z = Z[atoms[0]]
        if crit.desc is orc.com_distance_sum:
            m = np.array([orc.MASSES[int(a)] for a in z])
            v = com_distance_sum_batch(m, r, crit.desc_kwargs["entity_ids"])
        elif crit.desc is orc.max_atom_pair_dist:
            v = np.zeros(len(combs))
            for i in range(r.shape[1]):
                for j in range(i + 1, r.shape[1]):
                    dd = r[:, i] - r[:, j]
                    v = np.maximum(v, np.sqrt((dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]) + dd[:, 2] * dd[:, 2]))
        else:
            raise NotImplementedError("vectorised culling: unknown criteria descriptor")
        if isinstance(crit.cutoff, list):
            ok &= (crit.cutoff[0] < v) & (v < crit.cutoff[1])
        elif crit.bound == "upper":
            ok &= v < crit.cutoff
        else:
            ok &= v > crit.cutoff
    if return_coords:
        return ok, atoms, r
    return ok
