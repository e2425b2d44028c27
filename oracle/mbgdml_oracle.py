"""CPU oracle: a NumPy restatement of mbGDML's many-body GDML prediction path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mbgdml_b200/`` may import, call or
execute this file; it is the checker for ``tests/``, ``__graft_entry__.smoke()``
and the CPU-baseline / ``--impl reference`` legs of ``bench.py``.

Every function cites the reference lines (paths relative to /root/reference)
it follows.  The reference is pure Python/NumPy, so this restatement is
validated here, in the build container, against the *unmodified* reference
imported from /root/reference (``oracle/make_golden.py`` asserts oracle ==
reference on every fixture it writes) and against the reference's own
known-answer tests; outputs of that run are committed under ``tests/golden/``.

Pinned:    gen_combs, get_avail_entities, r_slice, descriptor/Jacobian,
           Matern contraction, model permutation expansion, Criteria /
           com_distance_sum (bit-exact known answer), predict_gdml(+decomp),
           mbePredict.predict -- all against the live reference.
Unpinned:  the periodic path rests on ``oracle/ase_mic.py`` (third-party
           ASE find_mic, restated, no real ASE available) -- "parity unpinned".

The arithmetic follows the reference's operation order where a discrete
decision depends on it (cut-offs, acceptance), so accept masks are
reproducible bit for bit.
"""
from __future__ import annotations

import itertools
import os
import importlib.util

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("_oracle_ase_mic", os.path.join(_here, "ase_mic.py"))
ase_mic = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(ase_mic)

# Most-abundant-isotope masses as returned by qcelemental.periodictable.to_mass
# (reference descriptors.py:125; H and O pinned by tests/test_descriptors.py:40-49).
MASSES = {
    1: 1.00782503223, 2: 4.00260325413, 3: 7.0160034366, 4: 9.012183065,
    5: 11.00930536, 6: 12.0, 7: 14.00307400443, 8: 15.99491461957,
    9: 18.99840316273, 10: 19.9924401762, 11: 22.9897692820, 12: 23.985041697,
    13: 26.98153853, 14: 27.97692653465, 15: 30.97376199842, 16: 31.9720711744,
    17: 34.968852682, 18: 39.9623831237,
}


# --------------------------------------------------------------------------
# Fragment enumeration
# --------------------------------------------------------------------------
def gen_combs(sets):
    """utils.py:449-489 (replacement=False): tuples of itertools.product(*sets)
    that have no repeated element and are already ascending, in product order."""
    for comb in itertools.product(*sets):
        ok = True
        for a, b in zip(comb, comb[1:]):
            if not a < b:  # "no repeats" + "sorted(comb) == list(comb)"  <=>  strictly ascending
                ok = False
                break
        if ok:
            yield comb


def gen_combs_array(sets):
    """Same sequence as :func:`gen_combs`, materialised as an (n, order) int64 array
    (what compute_nbody_decomp builds, mbe.py:936-944)."""
    sets = [np.asarray(s, dtype=np.int64) for s in sets]
    order = len(sets)
    if any(len(s) == 0 for s in sets):
        return np.zeros((0, order), dtype=np.int64)
    grids = np.meshgrid(*sets, indexing="ij")
    combs = np.stack([g.ravel() for g in grids], axis=1)  # C-order ravel == product order
    keep = np.ones(len(combs), dtype=bool)
    for k in range(order - 1):
        keep &= combs[:, k] < combs[:, k + 1]
    return combs[keep]


def get_avail_entities(comp_ids_r, comp_ids_model):
    """mbe.py:713-736: for each model component id, the entity ids of the structure carrying it."""
    comp_ids_r = np.asarray(comp_ids_r)
    return [np.where(cid == comp_ids_r)[0] for cid in comp_ids_model]


def r_slice(entity_ids, comb):
    """gdml_predict.py:214-216: atom indices of the fragment (entity order of comb,
    atoms ascending within each entity)."""
    out = []
    for e in comb:
        out.extend(np.where(entity_ids == e)[0])
    return np.asarray(out, dtype=np.int64)


# --------------------------------------------------------------------------
# Periodic cell (periodic.py:32-140)
# --------------------------------------------------------------------------
def _pdist_euclid(r):
    """scipy.spatial.distance.pdist(r, 'euclidean') order and arithmetic:
    pairs (i<j) row-major, sqrt(((dx*dx)+(dy*dy))+(dz*dz))."""
    n = len(r)
    iu, ju = np.triu_indices(n, k=1)
    d = r[iu] - r[ju]
    return np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])


class Cell:
    """periodic.py:32-140."""

    def __init__(self, cell_v, cutoff=None, pbc=True):
        self.cell_v = cell_v  # also sets the default cutoff (periodic.py:121-126)
        if cutoff is not None:
            self.cutoff = cutoff
        self.pbc = pbc

    @property
    def cell_v(self):
        return self._cell_v

    @cell_v.setter
    def cell_v(self, var):
        var = np.array(var)
        self._cell_v = var
        self.cutoff = np.min(np.linalg.norm(var, axis=1)) / 2.0

    @property
    def volume(self):
        v = self.cell_v
        return np.dot(v[2], np.cross(v[0], v[1]))

    def d_mic(self, d, check_cutoff=True):
        """periodic.py:61-84."""
        d_periodic, _ = ase_mic.find_mic(d, self.cell_v, pbc=self.pbc)
        if check_cutoff:
            if np.any(_pdist_euclid(d_periodic) >= self.cutoff):
                return None
        return d_periodic

    def r_mic(self, r):
        """periodic.py:86-107: coordinates relative to the first atom, minimum-imaged."""
        assert r.ndim == 2
        return self.d_mic(np.subtract(r, r[0, :]))


# --------------------------------------------------------------------------
# Acceptance criteria (descriptors.py:32-201)
# --------------------------------------------------------------------------
def center_of_mass(Z, R):
    """descriptors.py:106-128 (np.average with per-atom masses; sums run over atoms in order)."""
    R = np.asarray(R, dtype=np.float64)
    if R.ndim == 2:
        R = R[None]
    m = np.array([MASSES[int(z)] for z in Z], dtype=np.float64)
    num = np.zeros((R.shape[0], 3))
    den = 0.0
    for a in range(R.shape[1]):  # sequential accumulation == numpy's reduction over axis 1
        num = num + R[:, a, :] * m[a]
        den = den + m[a]
    return num / den


def com_distance_sum(Z, R, entity_ids):
    """descriptors.py:159-201: sum over entities of |CM_entity - CM_fragment|."""
    R = np.asarray(R, dtype=np.float64)
    if R.ndim == 2:
        R = R[None]
    Z = np.asarray(Z)
    entity_ids = np.asarray(entity_ids)
    cm_all = center_of_mass(Z, R)
    d_sum = np.zeros(R.shape[0])
    for eid in set(entity_ids.tolist()):
        idx = np.where(entity_ids == eid)[0]
        d = cm_all - center_of_mass(Z[idx], R[:, idx])
        d_sum += np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
    return d_sum


def max_atom_pair_dist(Z, R):
    """descriptors.py:131-156: largest atom-pair distance."""
    R = np.asarray(R, dtype=np.float64)
    if R.ndim == 2:
        R = R[None]
    out = np.empty(R.shape[0])
    for s in range(R.shape[0]):
        out[s] = np.max(_pdist_euclid(R[s])) if R.shape[1] > 1 else 0.0
    return out


class Criteria:
    """descriptors.py:32-103."""

    def __init__(self, desc, desc_kwargs, cutoff, bound="upper"):
        self.desc = desc
        self.desc_kwargs = desc_kwargs
        self.cutoff = cutoff
        if isinstance(self.cutoff, (tuple, list)):
            self.cutoff = sorted(self.cutoff)
        bound = bound.lower()
        assert bound in ["upper", "lower"]
        self.bound = bound

    def accept(self, Z, R, **kwargs):
        if R.ndim == 2:
            R = R[None, ...]
        n_R = R.shape[0]
        if self.cutoff is None:
            return np.full(n_R, True), None
        v = self.desc(Z, R, **self.desc_kwargs, **kwargs)
        if isinstance(self.cutoff, list):
            acc = (self.cutoff[0] < v) & (v < self.cutoff[1])
        elif self.bound == "upper":
            acc = v < self.cutoff
        else:
            acc = v > self.cutoff
        if n_R == 1:
            return acc[0], v[0]
        return acc, v


# --------------------------------------------------------------------------
# Descriptor (desc.py:42-233) and model expansion (gdml_model.py:90-125)
# --------------------------------------------------------------------------
def from_r(r, lat_and_inv=None):
    """desc.py:203-233 -> _pdist 79-108, _r_to_desc 137-158, _r_to_d_desc 161-200.

    x_k = 1/|r_i - r_j|, J_k = (r_i - r_j)/|r_i - r_j|^3 for (i, j) = np.tril_indices(n, -1)."""
    r = np.asarray(r, dtype=np.float64).reshape(-1, 3)
    n = r.shape[0]
    i, j = np.tril_indices(n, k=-1)
    pdiff = r[i] - r[j]
    if lat_and_inv is not None:  # desc.py:42-76 (model-level lattice; not the Cell path)
        lat, lat_inv = lat_and_inv
        c = lat_inv.dot(pdiff.T)
        pdiff = pdiff - lat.dot(np.around(c)).T
    pd = np.sqrt((pdiff[:, 0] * pdiff[:, 0] + pdiff[:, 1] * pdiff[:, 1]) + pdiff[:, 2] * pdiff[:, 2])
    return 1.0 / pd, pdiff / (pd**3)[:, None]


def desc_perm(perm):
    """desc.py:482-510: atom permutation -> descriptor (tril) permutation."""
    perm = np.asarray(perm)
    n = len(perm)
    rest = np.zeros((n, n))
    rest[np.tril_indices(n, -1)] = list(range((n * n - n) // 2))
    rest = rest + rest.T
    rest = rest[perm, :][:, perm]
    return rest[np.tril_indices(n, -1)].astype(int)


def tril_perms_lin_from_perms(perms):
    """_gdml/train.py:841-847: tril_perms_lin = (tril_perms + p*D).flatten('F')."""
    perms = np.asarray(perms)
    n_perms, n = perms.shape
    dim_d = n * (n - 1) // 2
    tril_perms = np.array([desc_perm(p) for p in perms])
    perm_offsets = np.arange(n_perms)[:, None] * dim_d
    return (tril_perms + perm_offsets).flatten("F")


class Model:
    """gdml_model.py:35-125: the attributes the predictor reads (gdml_predict.py:231-252)."""

    def __init__(self, model, comp_ids=None, criteria=None):
        if isinstance(model, str):
            model = dict(np.load(model, allow_pickle=True))
        elif not isinstance(model, dict):
            raise TypeError(f"{type(model)} is not string or dict")
        self.model_dict = model
        self.type = "gdml"
        self.criteria = criteria
        if "z" in model:
            self.Z = model["z"]
        elif "Z" in model:
            self.Z = model["Z"]
        else:
            raise KeyError("z or Z not found in model")
        self.comp_ids = comp_ids if comp_ids is not None else model["comp_ids"]
        self.nbody_order = len(self.comp_ids)
        self.sig = model["sig"]
        self.n_atoms = self.Z.shape[0]
        self.lat_and_inv = (
            (model["lattice"], np.linalg.inv(model["lattice"])) if "lattice" in model else None
        )
        self.n_train = model["R_desc"].shape[1]
        self.stdev = model["std"] if "std" in model else 1.0
        self.integ_c = model["c"]
        self.n_perms = model["perms"].shape[0]
        tpl = model["tril_perms_lin"]
        T, P = self.n_train, self.n_perms
        # gdml_model.py:109-118
        self.R_desc_perms = (
            np.tile(model["R_desc"].T, P)[:, tpl].reshape(T, P, -1, order="F").reshape(T * P, -1)
        )
        self.R_d_desc_alpha_perms = (
            np.tile(model["R_d_desc_alpha"], P)[:, tpl].reshape(T, P, -1, order="F").reshape(T * P, -1)
        )
        # gdml_model.py:120-125
        if "alphas_E" in model:
            self.alphas_E_lin = np.tile(model["alphas_E"][:, None], (1, P)).ravel()
        else:
            self.alphas_E_lin = None


# --------------------------------------------------------------------------
# Matern-5/2 contraction (gdml_predict.py:32-165)
# --------------------------------------------------------------------------
def predict_gdml_wkr(r_desc, r_d_desc, n_atoms, sig, n_perms, R_desc_perms,
                     R_d_desc_alpha_perms, alphas_E_lin=None, chunk_rows=None):
    """gdml_predict.py:32-165.  Returns out (1+3n,) = [E_raw, F_raw...].

    diff_j = x - x_j ; n_j = sqrt(5)|diff_j| ; a_j = diff_j . A_j
    c1_j = exp(-n_j/sig) * 5/(3 sig^3) ; c2_j = c1_j (n_j + sig)
    F_d = (5/sig) sum_j a_j c1_j diff_j - sum_j c2_j A_j ; E_raw = sum_j a_j c2_j
    """
    sig = float(sig)
    n_rows, dim_d = R_desc_perms.shape
    sig_inv = 1.0 / sig
    base_fact = 5.0 / (3 * sig**3)
    diag_fact = 5.0 / sig
    sqrt5 = np.sqrt(5.0)
    Fd = np.zeros(dim_d)
    E = 0.0
    step = n_rows if chunk_rows is None else int(chunk_rows)
    for lo in range(0, n_rows, step):
        Xj = R_desc_perms[lo:lo + step]
        Aj = R_d_desc_alpha_perms[lo:lo + step]
        diff = r_desc[None, :] - Xj
        nrm = sqrt5 * np.linalg.norm(diff, axis=1)
        c1 = np.exp(-nrm * sig_inv) * base_fact
        a = np.einsum("ji,ji->j", diff, Aj)
        Fd += (a * c1).dot(diff) * diag_fact
        c2 = c1 * (nrm + sig)
        Fd -= c2.dot(Aj)
        E += a.dot(c2)
        if alphas_E_lin is not None:  # gdml_predict.py:132-140
            aE = alphas_E_lin[lo:lo + step]
            Fd += aE.dot(diff * c2[:, None])
            K_ee = (1 + (nrm * sig_inv) * (1 + nrm / (3 * sig))) * np.exp(-nrm * sig_inv)
            E += K_ee.dot(aE)
    # J^T F_d (gdml_predict.py:151-162): k <-> (i>j): f_i += J_k Fd_k ; f_j -= J_k Fd_k
    i, j = np.tril_indices(n_atoms, k=-1)
    pair = np.zeros((n_atoms, n_atoms, 3))
    pair[i, j, :] = r_d_desc * Fd[:, None]
    pair[j, i, :] = -pair[i, j, :]
    out = np.empty(1 + 3 * n_atoms)
    out[0] = E
    out[1:] = pair.sum(axis=0).reshape(-1)
    return out


def virial_atom_loop(R, F):
    """stress.py:34-56."""
    v = np.zeros((3, 3))
    for ra, fa in zip(R, F):
        v += np.multiply.outer(ra, fa)
    return v


class AlchemyScale:
    """alchemy.py:30-62 with switching.linear_switching (switching.py:30-48)."""

    def __init__(self, entity_id, order, factor):
        assert 0 <= factor <= 1
        self.entity_id = entity_id
        self.order = order
        self.factor = factor

    def scale(self, data):
        return self.factor * data


def _fragment(Z, R, entity_ids, comb, model, periodic_cell, check_criteria_on_mic=True):
    """Slice -> MIC -> criteria for one fragment (gdml_predict.py:214-235).
    Returns (idx, z, r) or None when rejected."""
    idx = r_slice(entity_ids, comb)
    z = Z[idx]
    r = R[idx]
    if periodic_cell is not None:
        r = periodic_cell.r_mic(r)
        if r is None:
            return None
    if model.criteria is not None:
        ok, _ = model.criteria.accept(z, r)
        if not ok:
            return None
    return idx, z, r


def predict_gdml(Z, R, entity_ids, entity_combs, model, periodic_cell=None,
                 alchemy_scalers=None, compute_virial=False):
    """gdml_predict.py:170-272: total n-body E, F (and group virial) of one structure."""
    assert R.ndim == 2
    E = 0.0
    F = np.zeros(R.shape)
    compute_virial = bool(periodic_cell is not None and compute_virial)
    virial = np.zeros((3, 3))
    n_eval = 0
    for comb in entity_combs:
        frag = _fragment(Z, R, entity_ids, comb, model, periodic_cell)
        if frag is None:
            continue
        idx, z, r = frag
        x, J = from_r(r.flatten(), model.lat_and_inv)
        out = predict_gdml_wkr(x, J, len(z), model.sig, model.n_perms, model.R_desc_perms,
                               model.R_d_desc_alpha_perms, model.alphas_E_lin)
        out *= model.stdev
        e = out[0] + model.integ_c
        f = out[1:].reshape(len(z), 3)
        for sc in alchemy_scalers or []:
            if sc.entity_id in comb:
                e = sc.scale(e)
                f = sc.scale(f)
        E += e
        F[idx] += f
        n_eval += 1
        if compute_virial:
            virial += virial_atom_loop(r, f)
    predict_gdml.last_n_eval = n_eval
    if compute_virial:
        return E, F, virial
    return E, F


def predict_gdml_decomp(Z, R, entity_ids, entity_combs, model, alchemy_scalers=None):
    """gdml_predict.py:275-364: per-fragment E, F with NaN rows for rejected fragments.
    (No periodic support, as in the reference, mbe.py:565.)"""
    assert R.ndim == 2
    entity_combs = np.asarray(entity_combs)
    first = entity_combs[0] if entity_combs.ndim > 1 else entity_combs[:1]
    n_atoms = int(sum(np.count_nonzero(entity_ids == e) for e in np.atleast_1d(first)))
    E = np.full(len(entity_combs), np.nan)
    F = np.full((len(entity_combs), n_atoms, 3), np.nan)
    for k, comb in enumerate(entity_combs):
        comb = np.atleast_1d(comb)
        frag = _fragment(Z, R, entity_ids, comb, model, None)
        if frag is None:
            continue
        _, z, r = frag
        x, J = from_r(r.flatten(), model.lat_and_inv)
        out = predict_gdml_wkr(x, J, n_atoms, model.sig, model.n_perms, model.R_desc_perms,
                               model.R_d_desc_alpha_perms, model.alphas_E_lin)
        out *= model.stdev
        E[k] = out[0] + model.integ_c
        F[k] = out[1:].reshape(n_atoms, 3)
        for sc in alchemy_scalers or []:
            if sc.entity_id in comb:
                E[k] = sc.scale(E[k])
                F[k] = sc.scale(F[k])
    return E, F, entity_combs


# --------------------------------------------------------------------------
# MBE driver (mbe.py:739-811, 1009-1104)
# --------------------------------------------------------------------------
def mbe_predict(models, Z, R, entity_ids, comp_ids, periodic_cell=None, alchemy_scalers=None,
                compute_stress=False, use_voigt=False, only_normal_stress=False,
                box_scaling_type="anisotropic"):
    """mbePredict.predict (mbe.py:1009-1104) with virial_form == 'group'."""
    Z = np.asarray(Z)
    R = np.asarray(R, dtype=np.float64)
    if R.ndim == 2:
        R = R[None]
    entity_ids = np.asarray(entity_ids)
    n_R = R.shape[0]
    compute_stress = bool(compute_stress and periodic_cell is not None)
    E = np.zeros(n_R)
    F = np.zeros(R.shape)
    stress = np.zeros((n_R, 3, 3))
    for i in range(n_R):
        for model in models:
            avail = get_avail_entities(comp_ids, model.comp_ids)
            order = len(model.comp_ids)
            scalers = [s for s in (alchemy_scalers or []) if s.order == order]  # mbe.py:789-796
            res = predict_gdml(Z, R[i], entity_ids, gen_combs(avail), model, periodic_cell,
                               alchemy_scalers=scalers, compute_virial=compute_stress)
            E[i] += res[0]
            F[i] += res[1]
            if compute_stress:
                stress[i] += res[2]
    if not compute_stress:
        return E, F
    stress /= periodic_cell.volume  # mbe.py:1079
    if only_normal_stress:  # mbe.py:1081-1091
        for s in stress:
            diag = np.diag(s).copy()
            s[:] = 0.0
            if box_scaling_type == "isotropic":
                diag[:] = diag.sum() / 3
            s[np.diag_indices(3)] = diag
    if use_voigt:  # stress.py:31,59-72
        vi = [[0, 0], [1, 1], [2, 2], [1, 0], [0, 2], [0, 1]]
        stress = np.array([[s[a, b] for a, b in vi] for s in stress])
    return E, F, stress


def mbe_predict_decomp(models, Z, R, entity_ids, comp_ids):
    """mbePredict.predict_decomp (mbe.py:1106-1181).  compute_nbody_decomp passes no
    kwargs to the predictor (mbe.py:949-951), so alchemical scaling never applies here."""
    Z = np.asarray(Z)
    R = np.asarray(R, dtype=np.float64)
    if R.ndim == 2:
        R = R[None]
    entity_ids = np.asarray(entity_ids)
    orders = sorted({m.nbody_order for m in models})
    E_data, F_data, comb_data = [], [], []
    for order in orders:
        Es, Fs, Cs = [], [], []
        for model in (m for m in models if m.nbody_order == order):
            for i in range(R.shape[0]):
                combs = gen_combs_array(get_avail_entities(comp_ids, model.comp_ids))
                e, f, c = predict_gdml_decomp(Z, R[i], entity_ids, combs, model)
                Es.extend(e)
                Fs.extend(f)
                Cs.extend(np.hstack((np.full((len(c), 1), i, dtype=np.int64), c)))
        E_data.append(np.array(Es))
        F_data.append(np.array(Fs))
        comb_data.append(np.array(Cs))
    return E_data, F_data, comb_data, orders


# --------------------------------------------------------------------------
# MBE recombination (add / remove lower-order contributions)
# --------------------------------------------------------------------------
def mbe_worker(r_shape, entity_ids, r_prov_specs, r_idxs, E_lower, Deriv_lower, entity_ids_lower,
               r_prov_specs_lower):
    """mbe.py:73-197: for every structure, every combination of its entities of the lower order
    (itertools.combinations of r_prov_spec[2:]), first lower-order row with that provenance spec;
    missing rows are skipped; derivatives land on the atoms of the matching entity."""
    entity_ids = np.asarray(entity_ids)
    entity_ids_lower = np.asarray(entity_ids_lower)
    E = np.zeros(len(r_idxs))
    Deriv = np.zeros((len(r_idxs), *r_shape))
    n_lower = len(set(entity_ids_lower.tolist()))
    for i, i_r in enumerate(r_idxs):
        spec = r_prov_specs[i_r]
        for comb in itertools.combinations(spec[2:], n_lower):
            spec_lower = np.block([spec[:2], np.array(comb)])
            hits = np.where(np.all(r_prov_specs_lower == spec_lower, axis=1))[0]
            if hits.size == 0:
                continue
            i_low = hits[0]
            E[i] += E_lower[i_low]
            for prov in spec_lower[2:]:
                entity_id = np.where(spec[2:] == prov)[0][0]
                i_z = np.where(entity_ids == entity_id)[0]
                entity_id_lower = np.where(spec_lower[2:] == prov)[0][0]
                i_z_lower = np.where(entity_ids_lower == entity_id_lower)[0]
                Deriv[i][i_z] += Deriv_lower[i_low][i_z_lower]
    return r_idxs, E, Deriv


def mbe_contrib(E, Deriv, entity_ids, r_prov_ids, r_prov_specs, E_lower, Deriv_lower, entity_ids_lower,
                r_prov_ids_lower, r_prov_specs_lower, operation="remove"):
    """mbe.py:237-429 (use_ray=False branch); updates E and Deriv in place."""
    if operation not in ("add", "remove"):
        raise ValueError(f'{operation} is not "add" or "remove"')
    if (r_prov_ids is not None) and (r_prov_ids != {}):
        for k, md5 in r_prov_ids_lower.items():
            assert md5 == r_prov_ids[k]
    else:
        n_r = len(E)
        n_entities = len(set(np.asarray(entity_ids).tolist()))
        r_prov_specs = np.empty((n_r, 2 + n_entities), dtype=int)
        r_prov_specs[:, 0] = r_prov_specs_lower[0][0]
        r_prov_specs[:, 1] = np.arange(n_r)
        for e in range(n_entities):
            r_prov_specs[:, e + 2] = e
    r_idxs = []
    for k in r_prov_ids_lower.keys():
        r_idxs.extend(np.where(r_prov_specs[:, 0] == k)[0].tolist())
    _, E_w, D_w = mbe_worker(Deriv.shape[1:], entity_ids, r_prov_specs, r_idxs, E_lower, Deriv_lower,
                             entity_ids_lower, r_prov_specs_lower)
    if operation == "remove":
        E[r_idxs] -= E_w
        Deriv[r_idxs] -= D_w
    else:
        E[r_idxs] += E_w
        Deriv[r_idxs] += D_w
    return E, Deriv


def decomp_to_total(E_nbody, F_nbody, entity_ids, entity_combs):
    """mbe.py:432-508."""
    entity_ids = np.asarray(entity_ids)
    n_r = len(np.unique(entity_combs[:, 0]))
    E = np.zeros((n_r,))
    F = np.zeros((n_r, len(entity_ids), 3))
    entity_ids_lower = []
    for i in range(len(entity_combs[0][1:])):
        entity_ids_lower.extend([i] * np.count_nonzero(entity_ids == i))
    r_prov_specs_lower = np.hstack((np.zeros((len(entity_combs), 1)), entity_combs))
    E_nbody = np.nan_to_num(E_nbody, copy=True, nan=0.0)
    F_nbody = np.nan_to_num(F_nbody, copy=True, nan=0.0)
    return mbe_contrib(E, F, entity_ids, None, None, E_nbody, F_nbody, np.array(entity_ids_lower), {0: ""},
                       r_prov_specs_lower, operation="add")


# --------------------------------------------------------------------------
# Matern-5/2 covariance of structures vs. the training set, and the
# force-field kernel matrix of the closed-form trainer
# --------------------------------------------------------------------------
def gdml_mat52(model, Z, R):
    """analysis/models.py:31-139: k_5/2(x, x_j) = (1 + s d + 5 d^2 / (3 sig^2)) exp(-s d), s = sqrt(5)/sig,
    d = |x - x_j| over the T*P rows of ``R_desc_perms``; (n_R, T*P), or (T*P,) for one structure."""
    R = np.asarray(R, dtype=np.float64)
    if R.ndim == 2:
        R = R[None, ...]
    sig = float(model.sig)
    out = np.empty((R.shape[0], model.R_desc_perms.shape[0]))
    for m in range(R.shape[0]):
        x, _ = from_r(R[m].reshape(-1), model.lat_and_inv)
        d = np.linalg.norm(x[None, :] - model.R_desc_perms, axis=1)
        s = np.sqrt(5.0) / sig
        out[m] = (1.0 + s * d + 5.0 / (3 * sig**2) * d**2) * np.exp(-s * d)
    return out[0] if out.shape[0] == 1 else out


def d_desc_from_comp(r_d_desc, n_atoms):
    """desc.py:397-444: compressed Jacobian (D, 3) -> full (D, 3N); row k <-> (i>j): +g at atom j, -g at atom i."""
    D = n_atoms * (n_atoms - 1) // 2
    i, j = np.tril_indices(n_atoms, k=-1)
    out = np.zeros((D, n_atoms, 3))
    out[np.arange(D), j, :] = r_d_desc
    out[np.arange(D), i, :] = -r_d_desc
    return out.reshape(D, 3 * n_atoms)


def assemble_kernel_mat(R_desc, R_d_desc, tril_perms_lin, sig, use_E_cstr=False):
    """_gdml/train.py:92-292 (worker) and 1105-1337 (driver, ``col_idxs = np.s_[:]``): the Hessian-of-Matern
    force-field kernel matrix, (T 3N [+T]) square.  R_desc (T, D); R_d_desc (T, D, 3) compressed.

    Block (i, j), j < T (train.py:183-218), with pi_p(k) = tril_perms_lin[k P + p] mod D:
        diff_p = X_i - X_j[pi_p],  n_p = sqrt(5)|diff_p|,  b_p = exp(-n_p/sig) 5/(3 sig^4),  Jj_p = J_j[pi_p, :]
        K_ij = J_i^T sum_p [ 5 b_p diff_p (diff_p^T Jj_p) - (sig^2 + sig n_p) b_p Jj_p ]
    Energy rows/columns (use_E_cstr, train.py:224-289) as written there."""
    R_desc = np.asarray(R_desc, dtype=np.float64)
    R_d_desc = np.asarray(R_d_desc, dtype=np.float64)
    sig = float(sig)
    T, D = R_desc.shape
    n = int((1 + np.sqrt(8 * D + 1)) / 2)
    dim_i = 3 * n
    P = len(tril_perms_lin) // D
    pi = (np.asarray(tril_perms_lin).reshape(D, P) % D).T  # pi[p][k]
    J = np.stack([d_desc_from_comp(R_d_desc[t], n) for t in range(T)])  # (T, D, 3N)
    rows = T * dim_i + (T if use_E_cstr else 0)
    K = np.zeros((rows, rows))
    sqrt5 = np.sqrt(5.0)
    for j in range(T):
        Xjp = R_desc[j][pi]  # (P, D)
        Jjp = J[j][pi]  # (P, D, 3N)
        for i in range(T):
            diff = R_desc[i][None, :] - Xjp
            nrm = sqrt5 * np.linalg.norm(diff, axis=1)
            b = np.exp(-nrm / sig) / (3 * sig**4) * 5
            M = np.einsum("pk,pc->kc", diff * (5 * b)[:, None], np.einsum("pk,pkc->pc", diff, Jjp))
            M -= np.einsum("pkc,p->kc", Jjp, (sig**2 + sig * nrm) * b)
            K[i * dim_i:(i + 1) * dim_i, j * dim_i:(j + 1) * dim_i] = J[i].T.dot(M)
            if use_E_cstr:  # train.py:224-239
                kfe = 5 * diff / (3 * sig**3) * (nrm[:, None] + sig) * np.exp(-nrm / sig)[:, None]
                K[T * dim_i + i, j * dim_i:(j + 1) * dim_i] = -np.einsum("pk,pkc->c", kfe, Jjp)
    if use_E_cstr:  # train.py:241-289: columns T 3N + j
        for j in range(T):
            for i in range(T):
                Xip = R_desc[i][pi]
                Jip = J[i][pi]
                diff = R_desc[j][None, :] - Xip
                nrm = sqrt5 * np.linalg.norm(diff, axis=1)
                kfe = 5 * diff / (3 * sig**3) * (nrm[:, None] + sig) * np.exp(-nrm / sig)[:, None]
                K[i * dim_i:(i + 1) * dim_i, T * dim_i + j] = -np.einsum("pk,pkc->c", kfe, Jip)
                K[T * dim_i + i, T * dim_i + j] = -(1 + (nrm / sig) * (1 + nrm / (3 * sig))).dot(np.exp(-nrm / sig))
    return K
