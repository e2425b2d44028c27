"""Latency mode and per-frame cells: prepared steps (one CUDA-graph replay per evaluation), frames that
carry their own periodic cell, and the finite-difference virial (reference stress.py:75-152) as ONE 18-frame call."""
import numpy as np
import pytest

import mbgdml_b200 as mb
from helpers import b200_models, cutoffs_from, load, oracle_models, relerr, unpack_models
from mbgdml_b200 import _native, synthetic
from mbgdml_b200.stress import virial_finite_diff
from oracle import mbgdml_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _periodic_case():
    g = load("periodic_40h2o.npz")
    mds = unpack_models(g)
    cuts = cutoffs_from(g["crit_cutoffs"])
    al = [mb.mbeAlchemyScale(int(e), int(o), mb.linear_switching, {"factor": float(f)}) for e, o, f in g["alch"]]
    return g, mds, cuts, al


def test_replayed_graph_equals_plain_call_and_reference():
    """Calls 1 (plain), 2 (step created + run) and 3.. (graph replays) agree with each other and with the reference's
    outputs; new coordinates and a new cell (NPT) go through the same captured graph."""
    g, mds, cuts, al = _periodic_case()
    Z, R, eids, cids = g["cub_Z"], g["cub_R"], g["cub_eids"], g["cub_cids"]
    pred = mb.mbePredict(b200_models(mds, cuts), mb.predict_gdml, alchemy_scalers=al,
                         periodic_cell=mb.Cell(g["cub_cell"], pbc=True), compute_stress=True)
    outs = [pred.predict(Z, R, eids, cids) for _ in range(4)]
    assert len(pred._steps.steps) == 1  # built on the second call, replayed afterwards
    for E, F, S in outs:
        assert relerr(E, g["cub_E"]) <= TOL and relerr(F, g["cub_F"]) <= TOL and relerr(S, g["cub_stress"]) <= TOL
    assert pred.last_stats[2][2] == int(g["cub_naccept2"]) * R.shape[0] or pred.last_stats[2][2] > 0
    # other coordinates, other cell, same graph: against the plain path of a fresh predictor
    rng = np.random.default_rng(3)
    R2 = R + rng.normal(scale=0.05, size=R.shape)
    cell2 = np.asarray(g["cub_cell"]) * 1.013
    pred.periodic_cell = mb.Cell(cell2, pbc=True)
    E2, F2, S2 = pred.predict(Z, R2 * 1.013, eids, cids)
    assert len(pred._steps.steps) == 1
    plain = mb.mbePredict(b200_models(mds, cuts), mb.predict_gdml, alchemy_scalers=al,
                          periodic_cell=mb.Cell(cell2, pbc=True), compute_stress=True)
    plain.use_graph = False
    E3, F3, S3 = plain.predict(Z, R2 * 1.013, eids, cids)
    assert relerr(E2, E3) <= 1e-12 and relerr(F2, F3) <= 1e-11 and relerr(S2, S3) <= 1e-11
    assert [s[2] for s in pred.last_stats] == [s[2] for s in plain.last_stats]
    # triclinic cell through the same step object type
    pred_t = mb.mbePredict(b200_models(mds, cuts), mb.predict_gdml, alchemy_scalers=al,
                           periodic_cell=mb.Cell(g["tri_cell"], pbc=True), compute_stress=True)
    for _ in range(3):
        Et, Ft, St = pred_t.predict(Z, g["tri_R"], eids, cids)
    assert relerr(Et, g["tri_E"]) <= TOL and relerr(Ft, g["tri_F"]) <= TOL and relerr(St, g["tri_stress"]) <= TOL


def test_cluster_steps_and_md_like_loop():
    """Non-periodic: an MD-like loop of single-frame evaluations (the shape mbeCalculator produces)."""
    g = load("cluster_6h2o.npz")
    mds = unpack_models(g)
    cuts = cutoffs_from(g["crit_cutoffs"])
    Z, R, eids, cids = g["Z"], g["R"], g["entity_ids"], g["comp_ids"]
    pred = mb.mbePredict(b200_models(mds, cuts), mb.predict_gdml)
    ref = mb.mbePredict(b200_models(mds, cuts), mb.predict_gdml)
    ref.use_graph = False
    rng = np.random.default_rng(0)
    r = R[0].copy()
    for step in range(6):
        E, F = pred.predict(Z, r, eids, cids)
        E0, F0 = ref.predict(Z, r, eids, cids)
        assert relerr(E, E0) <= 1e-12 and relerr(F, F0) <= 1e-11
        r = r + 1e-3 * F[0] / np.abs(F[0]).max() + rng.normal(scale=1e-3, size=r.shape)
    assert len(pred._steps.steps) == 1
    Eb, Fb = pred.predict(Z, R, eids, cids)  # another frame count: its own step after two sightings
    assert relerr(Eb, g["crit_E"]) <= TOL and relerr(Fb, g["crit_F"]) <= TOL


def test_every_frame_under_its_own_cell():
    """mbg_system_desc.frame_cells: a batch whose frames have different cells == the frames evaluated one by one."""
    g, mds, cuts, al = _periodic_case()
    Z, eids, cids = g["cub_Z"], g["cub_eids"], g["cub_cids"]
    base = np.asarray(g["cub_cell"], dtype=float)
    cells = [base, base * 1.02, np.asarray(g["tri_cell"], dtype=float), base * np.array([1.0, 1.03, 0.98])[:, None]]
    frames = [g["cub_R"][0], g["cub_R"][1] * 1.02, g["tri_R"][0], g["cub_R"][0] * np.array([1.0, 1.03, 0.98])]
    pred = mb.mbePredict(b200_models(mds, cuts), mb.predict_gdml, alchemy_scalers=al, periodic_cell=mb.Cell(base, pbc=True),
                         compute_stress=True)
    pred.use_graph = False
    cutoffs = np.array([mb.Cell(c, pbc=True).cutoff for c in cells])
    E, F, V = pred._predict_batched(Z, np.array(frames), eids, cids, True, frame_cells=(np.array(cells), cutoffs))
    for k, (c, r) in enumerate(zip(cells, frames)):
        one = mb.mbePredict(b200_models(mds, cuts), mb.predict_gdml, alchemy_scalers=al, periodic_cell=mb.Cell(c, pbc=True),
                            compute_stress=True)
        one.use_graph = False
        E1, F1, S1 = one.predict(Z, r, eids, cids)
        assert relerr(E[k], E1[0]) <= 1e-12 and relerr(F[k], F1[0]) <= 1e-11
        assert relerr(V[k] / one.periodic_cell.volume, S1[0]) <= 1e-11


def test_virial_finite_diff_is_one_batched_call_and_matches_the_oracle():
    """stress.py:75-152: the 18 strained evaluations as one 18-frame native call.  (a) every strained energy and the
    resulting dE/dh against the oracle evaluated cell by cell; (b) for a compact cluster in a large box (no fragment
    crosses a cut-off under the strain) the finite-difference virial reproduces the analytic group virial:
    sum_k r_k (x) f_k = -h^T dE/dh."""
    md = [synthetic.make_model(f, n_train=40, sig=s_, seed=k) for k, (f, s_) in
          enumerate(zip((["h2o"], ["h2o", "h2o"], ["h2o", "h2o", "h2o"]), (10.0, 60.0, 200.0)))]
    Z, R, eids, cids = synthetic.make_cluster(5, seed=3, spacing=2.9)
    h = np.array([[21.0, 0.0, 0.0], [1.5, 20.0, 0.0], [-0.7, 2.1, 22.0]])
    R = R + np.array([9.0, 8.0, 10.0])
    ctx = _native.get_context()
    pred = mb.mbePredict(b200_models(md, [None] * 3), mb.predict_gdml, periodic_cell=mb.Cell(h, pbc=True), compute_stress=True)
    n0 = ctx.launch_count()
    W = virial_finite_diff(Z, R, eids, cids, h, pred, dh=1e-4)
    n_launch_fd = ctx.launch_count() - n0
    # (a) the oracle, one strained cell at a time
    omodels = oracle_models(md, [None] * 3)
    Rf = R @ np.linalg.inv(h)
    Wo = np.zeros((3, 3))
    scale = 0.0
    for i in range(3):
        for j in range(3):
            e = []
            for sgn in (-1, 1):
                hh = h.copy()
                hh[i, j] += sgn * 1e-4 / 2
                e.append(orc.mbe_predict(omodels, Z, Rf @ hh, eids, cids, periodic_cell=orc.Cell(hh))[0][0])
            Wo[i, j] = (e[1] - e[0]) / 1e-4
            scale = max(scale, abs(e[0]))
    assert np.abs(W - Wo).max() <= 1e-9 * scale / 1e-4  # energies agree to 1e-9 relative
    # one native call for all 18 strained frames: 3 jobs x (cull, scan, compact, contract) (+ nothing per frame)
    assert n_launch_fd <= 3 * 6
    # (b) analytic group virial
    pred.virial_form = "group"
    _, F, S = pred.predict(Z, R, eids, cids)
    V = S[0] * pred.periodic_cell.volume
    assert relerr(V, -h.T @ W) <= 5e-5
    # the driver-level switch (mbe.py:1060-1073) takes the same route
    pred.virial_form = "finite_diff"
    _, _, S_fd = pred.predict(Z, R, eids, cids)
    assert np.abs(S_fd[0] * pred.periodic_cell.volume - W).max() <= 1e-9 * scale / 1e-4  # (atomics: summation order varies)
    # a plugin other than predict_gdml is driven call by call, like the reference
    other = mb.mbePredict(b200_models(md, [None] * 3), lambda *a, **k: mb.predict_gdml(*a, **k), periodic_cell=mb.Cell(h, pbc=True))
    W2 = virial_finite_diff(Z, R, eids, cids, h, other, dh=1e-4)
    assert np.abs(W2 - W).max() <= 1e-9 * scale / 1e-4
    assert np.array_equal(other.periodic_cell.cell_v, h)
