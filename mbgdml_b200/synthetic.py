"""Synthetic structures and random-init GDML models of the shapes BASELINE.json names.

Pure NumPy data generation (no prediction arithmetic): used by bench.py, the
tests and oracle/make_golden.py so that every arm sees the same seeded inputs.
Recipe: SURVEY.md section 8(d).  Model dict keys follow what the reference's
``gdmlModel._load/_unpack`` consumes (reference mbgdml/models/gdml_model.py:60-125).
"""
from __future__ import annotations

import itertools

import numpy as np

# Rigid monomer templates (Angstrom): atoms listed as the reference's data sets do (heavy atom first).
_OH, _HOH = 0.9572, np.deg2rad(104.52)
WATER_Z = np.array([8, 1, 1])
WATER_R = np.array(
    [
        [0.0, 0.0, 0.0],
        [_OH * np.sin(_HOH / 2), _OH * np.cos(_HOH / 2), 0.0],
        [-_OH * np.sin(_HOH / 2), _OH * np.cos(_HOH / 2), 0.0],
    ]
)
# methanol: C, O, H(O), 3 x H(C)
MEOH_Z = np.array([6, 8, 1, 1, 1, 1])
MEOH_R = np.array(
    [
        [0.000, 0.000, 0.000],
        [1.427, 0.000, 0.000],
        [1.748, 0.903, 0.000],
        [-0.363, 1.028, 0.000],
        [-0.363, -0.514, 0.890],
        [-0.363, -0.514, -0.890],
    ]
)
# small extra species so that every fragment size the kernels are instantiated for (2..9 atoms) can be built
HF_Z = np.array([9, 1])
HF_R = np.array([[0.0, 0.0, 0.0], [0.917, 0.0, 0.0]])
NH3_Z = np.array([7, 1, 1, 1])
NH3_R = np.array([[0.0, 0.0, 0.116], [0.0, 0.939, -0.272], [0.813, -0.470, -0.272], [-0.813, -0.470, -0.272]])
CH4_Z = np.array([6, 1, 1, 1, 1])
CH4_R = 0.629 * np.array([[0.0, 0.0, 0.0], [1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]])
MONOMERS = {"h2o": (WATER_Z, WATER_R), "meoh": (MEOH_Z, MEOH_R), "hf": (HF_Z, HF_R), "nh3": (NH3_Z, NH3_R),
            "ch4": (CH4_Z, CH4_R)}
# Atom permutations that leave a monomer invariant (identity first).
MONOMER_PERMS = {
    "h2o": [[0, 1, 2], [0, 2, 1]],
    "meoh": [[0, 1, 2, 3, 4, 5], [0, 1, 2, 4, 5, 3], [0, 1, 2, 5, 3, 4]],
    "hf": [[0, 1]],
    "nh3": [[0, 1, 2, 3], [0, 2, 3, 1], [0, 3, 1, 2]],
    "ch4": [[0, 1, 2, 3, 4], [0, 2, 1, 4, 3], [0, 3, 4, 1, 2], [0, 4, 3, 2, 1]],
}


def random_rotations(rng, n):
    """n uniformly random rotation matrices (QR of a Gaussian matrix, det fixed to +1)."""
    q, r = np.linalg.qr(rng.normal(size=(n, 3, 3)))
    q = q * np.sign(np.diagonal(r, axis1=1, axis2=2))[:, None, :]
    q[:, :, 0] *= np.linalg.det(q)[:, None]
    return q


def fragment_perms(comp_ids, cap=100):
    """Symmetric group of a fragment: monomer-internal permutations x permutations of
    identical monomers, identity first, capped like reference _gdml/perm.py:275."""
    comp_ids = list(comp_ids)
    sizes = [len(MONOMERS[c][0]) for c in comp_ids]
    offs = np.concatenate(([0], np.cumsum(sizes)))
    perms = []
    for mol_perm in itertools.permutations(range(len(comp_ids))):
        if any(comp_ids[a] != comp_ids[b] for a, b in enumerate(mol_perm)):
            continue
        for inner in itertools.product(*[MONOMER_PERMS[c] for c in comp_ids]):
            p = []
            for slot, src in enumerate(mol_perm):
                p.extend(offs[src] + np.asarray(inner[slot]))
            perms.append(p)
    perms = np.array(perms, dtype=np.int64)
    ident = np.arange(offs[-1])
    order = sorted(range(len(perms)), key=lambda k: (not np.array_equal(perms[k], ident), k))
    return perms[order][:cap]


def _tril_desc(r):
    """1/|r_i - r_j| in np.tril_indices(n, -1) order, for rows of geometries (m, n, 3)."""
    n = r.shape[1]
    i, j = np.tril_indices(n, k=-1)
    d = r[:, i, :] - r[:, j, :]
    return 1.0 / np.sqrt((d * d).sum(-1))


def _desc_perm(perm):
    n = len(perm)
    rest = np.zeros((n, n))
    rest[np.tril_indices(n, -1)] = np.arange((n * n - n) // 2)
    rest = rest + rest.T
    rest = rest[perm, :][:, perm]
    return rest[np.tril_indices(n, -1)].astype(np.int64)


def fragment_geometries(rng, comp_ids, n, spread=3.0, jitter=0.05):
    """n geometries of a fragment: rigid monomers, random rotation, centres offset ~spread, xyz noise."""
    parts = []
    for k, cid in enumerate(comp_ids):
        _, r0 = MONOMERS[cid]
        rot = random_rotations(rng, n)
        r = np.einsum("mij,aj->mai", rot, r0 - r0.mean(0))
        centre = rng.normal(scale=0.6, size=(n, 1, 3))
        if k > 0:
            direction = rng.normal(size=(n, 1, 3))
            direction /= np.linalg.norm(direction, axis=-1, keepdims=True)
            centre = centre + direction * (spread + 0.4 * k)
        parts.append(r + centre)
    r = np.concatenate(parts, axis=1)
    return r + rng.normal(scale=jitter, size=r.shape)


def make_model(comp_ids, n_train=1000, sig=100.0, seed=0, c=0.0, std=1.0, use_E=False):
    """Random-init GDML model dict of the shape a trained model of these components has."""
    comp_ids = [str(c_) for c_ in comp_ids]
    rng = np.random.default_rng(seed)
    z = np.concatenate([MONOMERS[cid][0] for cid in comp_ids])
    n_atoms = len(z)
    dim_d = n_atoms * (n_atoms - 1) // 2
    r_train = fragment_geometries(rng, comp_ids, n_train)
    perms = fragment_perms(comp_ids)
    tril_perms = np.array([_desc_perm(p) for p in perms])
    n_perms = len(perms)
    tril_perms_lin = (tril_perms + np.arange(n_perms)[:, None] * dim_d).flatten("F")
    entity_ids = np.concatenate([np.full(len(MONOMERS[cid][0]), k) for k, cid in enumerate(comp_ids)])
    model = {
        "z": z.astype(np.int64),
        "r_unit": np.array("Angstrom"),
        "e_unit": np.array("kcal/mol"),
        "comp_ids": np.array(comp_ids),
        "entity_ids": entity_ids.astype(np.int64),
        "sig": float(sig),
        "R_desc": np.ascontiguousarray(_tril_desc(r_train).T),  # (D, T)
        "R_d_desc_alpha": rng.normal(size=(n_train, dim_d)),  # (T, D)
        "c": float(c),
        "std": float(std),
        "perms": perms,
        "tril_perms_lin": tril_perms_lin.astype(np.int64),
    }
    if use_E:
        model["alphas_E"] = rng.normal(size=(n_train,))
    return model


def _place(rng, comp_list, centres, jitter=0.0):
    z, r, eids = [], [], []
    rots = random_rotations(rng, len(comp_list))
    for k, cid in enumerate(comp_list):
        z0, r0 = MONOMERS[cid]
        z.append(z0)
        r.append((r0 - r0.mean(0)) @ rots[k].T + centres[k])
        eids.append(np.full(len(z0), k))
    r = np.concatenate(r)
    if jitter:
        r = r + rng.normal(scale=jitter, size=r.shape)
    return np.concatenate(z).astype(np.int64), r, np.concatenate(eids).astype(np.int64), np.array(comp_list)


def make_cluster(n_mol, seed=0, spacing=3.1, comp="h2o", grid_jitter=0.25):
    """Non-periodic cluster: molecules on a jittered cubic grid, random orientations.
    Returns Z, R (n_atoms, 3), entity_ids, comp_ids."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(n_mol ** (1 / 3) - 1e-9))
    sites = np.array(list(itertools.product(range(side), repeat=3)), dtype=float)[:n_mol]
    centres = sites * spacing + rng.uniform(-grid_jitter, grid_jitter, size=(n_mol, 3))
    comp_list = [comp] * n_mol if isinstance(comp, str) else list(comp)
    return _place(rng, comp_list, centres)


def make_periodic_box(n_mol=140, box=16.12, n_frames=1, seed=0, comp="h2o", frame_sigma=0.1):
    """Cubic periodic box: molecules on a jittered grid, per-frame Gaussian displacement of
    whole atoms, then wrapped into [0, L).  Returns Z, R (n_frames, n_atoms, 3), entity_ids,
    comp_ids, cell_v (3, 3)."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(n_mol ** (1 / 3) - 1e-9))
    sites = np.array(list(itertools.product(range(side), repeat=3)), dtype=float)
    pick = np.sort(rng.choice(len(sites), size=n_mol, replace=False))
    a = box / side
    centres = (sites[pick] + 0.5) * a + rng.uniform(-0.15, 0.15, size=(n_mol, 3))
    comp_list = [comp] * n_mol if isinstance(comp, str) else list(comp)
    if not isinstance(comp, str):  # mixtures: species are listed in blocks (entity ids), but mixed in space
        centres = centres[rng.permutation(n_mol)]
    z, r0, eids, comp_ids = _place(rng, comp_list, centres)
    R = r0[None] + rng.normal(scale=frame_sigma, size=(n_frames,) + r0.shape)
    R = np.mod(R, box)
    return z, R, eids, comp_ids, np.eye(3) * box


def jitter_frames(R0, n_frames, sigma=0.02, seed=0):
    """n_frames copies of one structure with N(0, sigma) noise (BASELINE config 1)."""
    rng = np.random.default_rng(seed)
    return R0[None] + rng.normal(scale=sigma, size=(n_frames,) + R0.shape)
