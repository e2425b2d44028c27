"""Full-frame parity of the BENCHMARK configurations (BASELINE configs 2, 4 and 5) against the oracle.

(a) accept masks of EVERY candidate fragment of a frame, bit for bit, against the vectorised oracle culling
    (oracle/cull_vec.py == reference gdml_predict.py:214-235, periodic.py:61-107, descriptors.py:65-201);
(b) energy and forces of a whole frame (every accepted fragment, every model, T = 1000) against the oracle
    fanned over the host cores, <= 1e-9 relative (north_star);
(c) config 2: 100,000 monomers, energies/forces of a sample against the oracle, per-structure decomposition.
"""
import math
import multiprocessing as mp
import os

import numpy as np
import pytest

import mbgdml_b200 as mb
from helpers import b200_models, oracle_models, relerr
from mbgdml_b200 import synthetic
from mbgdml_b200.predictors import accept_mask
from oracle import cull_vec
from oracle import mbgdml_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9

BOX = dict(
    families=[["h2o"], ["h2o", "h2o"], ["h2o", "h2o", "h2o"]], sigs=[10.0, 100.0, 400.0], cutoffs=[None, 6.0, 10.0],
    alchemy=[(0, 3, 0.5), (0, 2, 0.5)], box=16.12, comp="h2o")
MIXED = dict(
    families=[["h2o"], ["meoh"], ["h2o", "h2o"], ["h2o", "meoh"], ["meoh", "meoh"], ["h2o", "h2o", "h2o"],
              ["h2o", "h2o", "meoh"], ["h2o", "meoh", "meoh"], ["meoh", "meoh", "meoh"]],
    sigs=[10.0, 20.0, 100.0, 120.0, 150.0, 400.0, 420.0, 450.0, 500.0],
    cutoffs=[None, None, 6.0, 6.5, 7.0, 10.0, 10.5, 11.0, 11.5], alchemy=[(100, 3, 0.5)], box=17.8,
    comp=["h2o"] * 100 + ["meoh"] * 40)


def _workload(cfg, n_train, n_frames=1):
    dicts = [synthetic.make_model(f, n_train=n_train, sig=s, seed=k)
             for k, (f, s) in enumerate(zip(cfg["families"], cfg["sigs"]))]
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=140, box=cfg["box"], n_frames=n_frames, seed=0,
                                                           comp=cfg["comp"], frame_sigma=0.1)
    return dicts, Z, R, eids, cids, cell_v


@pytest.mark.parametrize("name", ["box140", "mixed140"])
def test_full_frame_accept_masks_bit_exact(name):
    """Every candidate of a frame: device decision == oracle decision (21,368 of 447,580 trimers accepted in the
    140-H2O box; product-space enumeration with mixed tuples in the water-methanol box)."""
    cfg = BOX if name == "box140" else MIXED
    dicts, Z, R, eids, cids, cell_v = _workload(cfg, n_train=4)  # the accept decision does not depend on T
    models, omodels = b200_models(dicts, cfg["cutoffs"]), oracle_models(dicts, cfg["cutoffs"])
    cell, ocell = mb.Cell(cell_v, pbc=True), orc.Cell(cell_v)
    pred = mb.mbePredict(models, mb.predict_gdml, periodic_cell=cell)
    n_acc = []
    for m, om in zip(models, omodels):
        sets = pred.get_avail_entities(cids, m.comp_ids)
        got = accept_mask(Z, R[0], eids, mb.gen_combs(sets), m, cell)[0]
        combs = orc.gen_combs_array(sets)
        assert int((got != 0).sum()) == len(combs)  # the tuples gen_combs yields, utils.py:449-489
        want = cull_vec.accept_mask(Z, R[0], eids, combs, om, ocell)
        assert np.array_equal(got[got != 0] == 1, want), f"{name} {m.comp_ids}: accept masks differ"
        # explicit tuples take the same decisions as the enumerated candidates
        sel = np.random.default_rng(1).choice(len(combs), min(len(combs), 5000), replace=False)
        got_x = accept_mask(Z, R[0], eids, combs[sel], m, cell)[0]
        assert np.array_equal(got_x == 1, want[sel])
        n_acc.append(int(want.sum()))
    if name == "box140":
        assert n_acc == [140, 1997, 21368]  # SURVEY 8(d) sizing, frame 0 of the benchmark workload
    else:
        assert all(n > 0 for n in n_acc)


# ---- (b) whole-frame energies and forces ---------------------------------------------------------------
_W = None


def _oracle_chunk(args):
    from threadpoolctl import threadpool_limits

    k, combs = args
    Z, R0, eids, omodels, ocell, scalers = _W
    order = len(omodels[k].comp_ids)
    with threadpool_limits(limits=1):
        E, F = orc.predict_gdml(Z, R0, eids, combs, omodels[k], ocell,
                                alchemy_scalers=[s for s in scalers if s.order == order])[:2]
    return E, F


def _oracle_frame(Z, R0, eids, cids, omodels, ocell, scalers):
    """mbePredict.predict of one frame with the oracle: culling vectorised, accepted fragments over all cores."""
    global _W
    tasks = []
    for k, om in enumerate(omodels):
        combs = orc.gen_combs_array(orc.get_avail_entities(cids, om.comp_ids))
        acc = combs[cull_vec.accept_mask(Z, R0, eids, combs, om, ocell)]
        per = max(1, min(400, len(acc) // 64 + 1))
        tasks += [(k, acc[i:i + per]) for i in range(0, len(acc), per)]
    tasks.sort(key=lambda t: -len(t[1]) * len(omodels[t[0]].R_desc_perms))  # longest first
    _W = (Z, R0, eids, omodels, ocell, scalers)
    with mp.get_context("fork").Pool(os.cpu_count() or 1) as pool:
        parts = pool.map(_oracle_chunk, tasks, chunksize=1)
    _W = None
    return sum(p[0] for p in parts), sum(p[1] for p in parts), sum(len(t[1]) for t in tasks)


@pytest.mark.parametrize("name,n_train", [("box140", 1000), ("mixed140", 500)])
def test_full_frame_energy_forces_vs_oracle(name, n_train):
    """One complete frame of the benchmark workload -- all models, all accepted fragments, alchemical scaling,
    minimum image -- against the oracle: the number the bench times is the number the reference computes."""
    cfg = BOX if name == "box140" else MIXED
    dicts, Z, R, eids, cids, cell_v = _workload(cfg, n_train=n_train)
    models, omodels = b200_models(dicts, cfg["cutoffs"]), oracle_models(dicts, cfg["cutoffs"])
    al = [mb.mbeAlchemyScale(e, o, mb.linear_switching, {"factor": f}) for e, o, f in cfg["alchemy"]]
    oal = [orc.AlchemyScale(e, o, f) for e, o, f in cfg["alchemy"]]
    pred = mb.mbePredict(models, mb.predict_gdml, alchemy_scalers=al, periodic_cell=mb.Cell(cell_v, pbc=True))
    E, F = pred.predict(Z, R[:1], eids, cids)
    Eo, Fo, n_eval = _oracle_frame(Z, R[0], eids, cids, omodels, orc.Cell(cell_v), oal)
    assert n_eval == sum(s[2] for s in pred.last_stats)
    assert abs(E[0] - Eo) <= TOL * abs(Eo), (E[0], Eo)
    assert relerr(F[0], Fo) <= TOL, relerr(F[0], Fo)


def test_config2_100k_monomers_sample_vs_oracle():
    """BASELINE config 2: 100,000 monomer structures through the 1-body model; every structure's energy and forces
    are separable (one fragment each), a random sample is checked against the oracle."""
    md = synthetic.make_model(["h2o"], n_train=1000, sig=10.0, seed=0)
    rng = np.random.default_rng(0)
    n = 100000
    r = synthetic.WATER_R[None] + rng.normal(scale=0.05, size=(n, 3, 3))
    Z, eids, cids = np.tile(synthetic.WATER_Z, n), np.repeat(np.arange(n), 3), np.array(["h2o"] * n)
    model, om = mb.gdmlModel(dict(md)), orc.Model(dict(md))
    pred = mb.mbePredict([model], mb.predict_gdml)
    E, F = pred.predict(Z, r.reshape(1, 3 * n, 3), eids, cids)
    assert pred.last_stats[0][1] == n and pred.last_stats[0][2] == n
    sel = rng.choice(n, 300, replace=False)
    Eo, Fo, _ = orc.predict_gdml_decomp(Z, r.reshape(3 * n, 3), eids, sel.reshape(-1, 1), om)
    Fg = F[0].reshape(n, 3, 3)[sel]
    assert relerr(Fg, Fo) <= TOL
    # the total energy is the sum over monomers: check it through the decomposed GPU rows of the sample, and the
    # whole sum through a second, independent grouping (per-thousand partial sums of explicit tuple batches)
    Ed, Fd, _ = mb.predict_gdml_decomp(Z, r.reshape(3 * n, 3), eids, sel.reshape(-1, 1), model)
    assert relerr(Ed, Eo) <= TOL and relerr(Fd, Fo) <= TOL
    E_parts = sum(mb.predict_gdml(Z, r.reshape(3 * n, 3), eids, np.arange(lo, lo + 25000).reshape(-1, 1), model)[0]
                  for lo in range(0, n, 25000))
    assert abs(E_parts - E[0]) <= 1e-11 * abs(E[0])
    assert math.isfinite(E[0])
