"""The vectorised oracle culling (oracle/cull_vec.py) takes the scalar oracle's decisions (== the reference's
gdml_predict.py:214-235 loop) bit for bit: cubic and triclinic cells, both criteria descriptors, explicit cut-off."""
import numpy as np
import pytest

from helpers import cutoffs_from, load, oracle_models, unpack_models
from oracle import cull_vec
from oracle import mbgdml_oracle as orc


@pytest.mark.parametrize("tag,cutoff_key", [("cub", None), ("tri", None), ("tri", "tribig_cutoff")])
@pytest.mark.parametrize("desc", ["com", "pair"])
def test_vectorised_mask_equals_scalar_oracle(tag, cutoff_key, desc):
    g = load("periodic_40h2o.npz")
    mds = unpack_models(g)
    cuts = cutoffs_from(g["crit_cutoffs"])
    oms = oracle_models(mds, cuts, desc=desc)
    Z, eids, cids = g["cub_Z"], g["cub_eids"], g["cub_cids"]
    R = g[tag + "_R"][0]
    cell = orc.Cell(g[tag + "_cell"], cutoff=None if cutoff_key is None else float(g[cutoff_key]))
    rng = np.random.default_rng(0)
    for m in oms:
        combs = orc.gen_combs_array(orc.get_avail_entities(cids, m.comp_ids))
        vec = cull_vec.accept_mask(Z, R, eids, combs, m, cell)
        sel = rng.choice(len(combs), min(len(combs), 600), replace=False)
        scalar = np.array([orc._fragment(Z, R, eids, combs[i], m, cell) is not None for i in sel])
        assert np.array_equal(vec[sel], scalar)
        assert 0 < vec.sum() <= len(combs)


def test_non_periodic_and_range_criteria():
    g = load("cluster_6h2o.npz")
    mds = unpack_models(g)
    Z, R, eids, cids = g["Z"], g["R"][0], g["entity_ids"], g["comp_ids"]
    for bound, cut in (("upper", 6.0), ("lower", 4.0), ("upper", [3.0, 7.5])):
        oms = oracle_models(mds, [None, cut, cut], bounds=[bound] * 3)
        for m in oms:
            combs = orc.gen_combs_array(orc.get_avail_entities(cids, m.comp_ids))
            vec = cull_vec.accept_mask(Z, R, eids, combs, m, None)
            scalar = np.array([orc._fragment(Z, R, eids, c, m, None) is not None for c in combs])
            assert np.array_equal(vec, scalar)


@pytest.mark.parametrize("pbc", [(True, True, False), (False, True, True), (True, False, False)])
def test_vectorised_mask_equals_scalar_oracle_partial_pbc(pbc):
    """Cell(pbc=...) with a non-periodic direction (periodic.py:40-59 -> find_mic(..., pbc)): slab / wire geometries."""
    g = load("periodic_40h2o.npz")
    oms = oracle_models(unpack_models(g), cutoffs_from(g["crit_cutoffs"]))
    Z, eids, cids = g["cub_Z"], g["cub_eids"], g["cub_cids"]
    rng = np.random.default_rng(1)
    for tag in ("cub", "tri"):
        R = g[tag + "_R"][0]
        cell = orc.Cell(g[tag + "_cell"], pbc=pbc)
        for m in oms[1:]:
            combs = orc.gen_combs_array(orc.get_avail_entities(cids, m.comp_ids))
            vec = cull_vec.accept_mask(Z, R, eids, combs, m, cell)
            sel = rng.choice(len(combs), min(len(combs), 400), replace=False)
            scalar = np.array([orc._fragment(Z, R, eids, combs[i], m, cell) is not None for i in sel])
            assert np.array_equal(vec[sel], scalar)
            assert 0 < vec.sum() < len(combs)


def test_vectorised_mask_equals_scalar_oracle_four_body():
    """4-body fragments (the reference point of tests/test_gpu_pairlist.py::test_four_body_...): periodic box."""
    from mbgdml_b200 import synthetic

    md = synthetic.make_model(["h2o"] * 4, n_train=4, sig=500.0, seed=31)
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=24, box=9.0, n_frames=1, seed=6, frame_sigma=0.1)
    om = oracle_models([md], [13.0])[0]
    cell = orc.Cell(cell_v, cutoff=6.5)  # above L/2: the general branch of find_mic decides
    combs = orc.gen_combs_array([list(range(24))] * 4)
    vec = cull_vec.accept_mask(Z, R[0], eids, combs, om, cell)
    sel = np.random.default_rng(2).choice(len(combs), 500, replace=False)
    scalar = np.array([orc._fragment(Z, R[0], eids, combs[i], om, cell) is not None for i in sel])
    assert np.array_equal(vec[sel], scalar) and 0 < vec.sum() < len(combs)
