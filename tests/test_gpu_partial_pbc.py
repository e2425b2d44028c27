"""Partial periodicity -- ``Cell(cell_v, cutoff, pbc=(True, True, False))`` (reference periodic.py:40-59; the flags go
to ``ase.geometry.find_mic(d, cell_v, pbc=self.pbc)``, periodic.py:78): slabs and wires.  With a non-periodic direction
find_mic never tries the naive image; it Gauss-reduces the periodic pair of cell vectors, wraps the periodic fractional
coordinates and searches the images of the periodic directions only.  Device (geom.cuh mic_general with per-direction
limits, hostmath minkowski_reduce_pbc) against the oracle's restatement (oracle/ase_mic.py; parity unpinned like all of
find_mic, checked there against a brute-force image search)."""
import numpy as np
import pytest

import mbgdml_b200 as mb
from helpers import b200_models, cutoffs_from, load, oracle_models, relerr, unpack_models
from mbgdml_b200 import _native
from mbgdml_b200.predictors import accept_mask
from oracle import cull_vec
from oracle import mbgdml_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9

CELLS = {
    "cubic": np.eye(3) * 12.4,
    "ortho": np.diag([9.0, 11.5, 14.2]),
    "triclinic": np.array([[12.6, 0.0, 0.0], [2.8, 12.1, 0.0], [-1.9, 3.1, 11.8]]),
    "skewed": np.array([[10.0, 0.0, 0.0], [7.0, 8.0, 0.0], [3.0, -5.0, 9.0]]),
}
PBCS = [(True, True, False), (True, False, True), (False, True, True), (False, True, False), (True, False, False),
        (False, False, False)]


@pytest.mark.parametrize("name", list(CELLS))
@pytest.mark.parametrize("pbc", PBCS)
def test_r_mic_and_d_mic_operators(name, pbc):
    cell_v = CELLS[name]
    rng = np.random.default_rng(11)
    for cutoff in (None, 1000.0):  # the default (half the shortest vector) rejects some; a wide one accepts all
        cell, ocell = mb.Cell(cell_v, cutoff=cutoff, pbc=pbc), orc.Cell(cell_v, cutoff=cutoff, pbc=pbc)
        n_none = 0
        for _ in range(12):
            r = rng.uniform(-1.2, 1.2, size=(6, 3)) @ cell_v
            if cutoff is None:
                r = r[:1] + 0.25 * (r - r[:1])
            got, want = cell.r_mic(r), ocell.r_mic(r)
            assert (got is None) == (want is None)
            n_none += got is None
            if got is not None:
                assert np.abs(got - want).max() <= 1e-11
            d = rng.uniform(-1.5, 1.5, size=(5, 3)) @ cell_v
            got, want = cell.d_mic(d, check_cutoff=False), ocell.d_mic(d, check_cutoff=False)
            assert np.abs(got - want).max() <= 1e-11
            # no translation along a non-periodic direction
            frac = np.linalg.solve(cell_v.T, (got - d).T).T
            assert np.abs(frac[:, ~np.array(pbc)]).max(initial=0.0) <= 1e-9
        if cutoff is not None:
            assert n_none == 0


@pytest.mark.parametrize("pbc", [(True, True, False), (False, True, True), (True, False, False)])
@pytest.mark.parametrize("tag", ["cub", "tri"])
def test_slab_accept_masks_and_prediction(pbc, tag):
    """40-H2O box as a slab / wire: accept masks of every candidate bit for bit against the vectorised oracle culling
    (full enumeration and pair list), energies and forces of the whole frame against the oracle."""
    g = load("periodic_40h2o.npz")
    mds = unpack_models(g)
    cuts = cutoffs_from(g["crit_cutoffs"])
    models, omodels = b200_models(mds, cuts), oracle_models(mds, cuts)
    Z, eids, cids = g["cub_Z"], g["cub_eids"], g["cub_cids"]
    R = g[tag + "_R"][:1]
    cell, ocell = mb.Cell(g[tag + "_cell"], pbc=pbc), orc.Cell(g[tag + "_cell"], pbc=pbc)
    pred = mb.mbePredict(models, mb.predict_gdml, periodic_cell=cell)
    ctx = _native.get_context()
    E_o, F_o = 0.0, np.zeros_like(R[0])
    try:
        for m, om in zip(models, omodels):
            sets = pred.get_avail_entities(cids, m.comp_ids)
            combs = orc.gen_combs_array(sets)
            want = cull_vec.accept_mask(Z, R[0], eids, combs, om, ocell)
            for kind in ("full", "pairlist"):
                ctx.set_culling(kind)
                got = accept_mask(Z, R[0], eids, mb.gen_combs(sets), m, cell)[0]
                assert np.array_equal(got[got != 0] == 1, want), (kind, m.comp_ids)
            e, f = orc.predict_gdml(Z, R[0], eids, combs[want], om, ocell)[:2]
            E_o, F_o = E_o + e, F_o + f
    finally:
        ctx.set_culling("auto")
    for _ in range(2):  # plain call, then the prepared step
        E, F = pred.predict(Z, R, eids, cids)
        assert abs(E[0] - E_o) <= TOL * abs(E_o) and relerr(F[0], F_o) <= TOL
    # the fully periodic box gives something else: the flags are not ignored
    E_full, _ = mb.mbePredict(models, mb.predict_gdml, periodic_cell=mb.Cell(g[tag + "_cell"], pbc=True)).predict(Z, R, eids, cids)
    assert abs(E_full[0] - E_o) > 1e-6 * abs(E_o)


def test_slab_and_wire_golden_from_the_live_reference():
    """Energies and forces of the UNMODIFIED reference (`mbePredict` + `predict_gdml` + `Cell(pbc=...)`, its find_mic
    the restated one) for the 40-H2O box as slab / wire, cubic and triclinic: tests/golden/slab_40h2o.npz."""
    g = load("slab_40h2o.npz")
    mds = unpack_models(g)
    models = b200_models(mds, cutoffs_from(g["crit_cutoffs"]))
    for tag in ("cub", "tri"):
        for k in range(len(g["pbcs"])):
            pbc = tuple(bool(x) for x in g["pbcs"][k])
            pred = mb.mbePredict(models, mb.predict_gdml, periodic_cell=mb.Cell(g[tag + "_cell"], pbc=pbc))
            E, F = pred.predict(g["Z"], g[tag + "_R"][:1], g["eids"], g["cids"])
            assert relerr(E, g[f"{tag}_E{k}"][:1]) <= TOL and relerr(F, g[f"{tag}_F{k}"][:1]) <= TOL, (tag, pbc)
