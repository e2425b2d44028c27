"""Pair-list candidate generation (csrc/pairlist.cuh, MBG_CULL_PAIRLIST) against the full enumeration (k_cull, which
tests/test_gpu_fullframe.py pins bit for bit to the oracle culling == reference gdml_predict.py:214-235,
periodic.py:61-107, descriptors.py:65-201): the accepted set, its order and the fragment records must not depend on
how the candidates were generated.

* accept masks of EVERY candidate, byte for byte: the 140-H2O benchmark box, a 512-H2O box at the same density with
  a physical cut-off, a triclinic cell (general minimum image in the pair test), clusters with com_distance_sum and
  max_atom_pair_dist (the two non-periodic pair conditions), lower / range bounds;
* energies, forces, virial and the per-job counts of whole predictions, plain calls, sharded calls and prepared steps.
"""
import numpy as np
import pytest

import mbgdml_b200 as mb
from helpers import b200_models, oracle_models, relerr
from mbgdml_b200 import _engine, _native, synthetic
from mbgdml_b200.predictors import accept_mask
from oracle import cull_vec
from oracle import mbgdml_oracle as orc

pytestmark = pytest.mark.gpu

FAMILIES = [["h2o", "h2o"], ["h2o", "h2o", "h2o"]]


@pytest.fixture()
def ctx():
    c = _native.get_context()
    yield c
    c.set_culling("auto")


def _dicts(n_train=4):
    return [synthetic.make_model(f, n_train=n_train, sig=s, seed=k + 1) for k, (f, s) in enumerate(zip(FAMILIES, [100.0, 400.0]))]


def _masks(ctx, Z, R, eids, cids, models, cell):
    pred = mb.mbePredict(models, mb.predict_gdml, periodic_cell=cell)
    out = {}
    for kind in ("full", "pairlist"):
        ctx.set_culling(kind)
        out[kind] = [accept_mask(Z, R, eids, mb.gen_combs(pred.get_avail_entities(cids, m.comp_ids)), m, cell)
                     for m in models]
    return out


def _assert_same(masks, min_accepted=1):
    for full, nl in zip(masks["full"], masks["pairlist"]):
        assert full.shape == nl.shape
        assert int((full == 1).sum()) >= min_accepted
        assert int((full == 2).sum()) > 0  # something was culled
        assert np.array_equal(full, nl)


def test_box140_masks_equal_and_match_the_oracle(ctx):
    """Two frames of the benchmark box (dimers < 6.0, trimers < 10.0, cut-off L/2), and frame 0 against the oracle."""
    dicts = _dicts()
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=140, box=16.12, n_frames=2, seed=0, frame_sigma=0.1)
    cell = mb.Cell(cell_v, pbc=True)
    masks = _masks(ctx, Z, R, eids, cids, b200_models(dicts, [6.0, 10.0]), cell)
    _assert_same(masks)
    assert [int((m[0] == 1).sum()) for m in masks["pairlist"]] == [1997, 21368]  # SURVEY 8(d) sizing
    om = oracle_models(dicts, [6.0, 10.0])[1]
    combs = orc.gen_combs_array([list(range(140))] * 3)
    want = cull_vec.accept_mask(Z, R[0], eids, combs, om, orc.Cell(cell_v))
    assert np.array_equal(masks["pairlist"][1][0] == 1, want)


@pytest.mark.parametrize("cutoff", [8.0, None])
def test_box512_masks_equal(ctx, cutoff):
    """512 H2O at the density of the benchmark box: C(512, 3) = 22.2 M candidates per frame."""
    box = 16.12 * (512 / 140) ** (1 / 3)
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=512, box=box, n_frames=1, seed=3, frame_sigma=0.1)
    cell = mb.Cell(cell_v, cutoff=cutoff, pbc=True)
    masks = _masks(ctx, Z, R, eids, cids, b200_models(_dicts(), [6.0, 10.0]), cell)
    _assert_same(masks, min_accepted=1000)


def test_triclinic_cell_masks_equal(ctx):
    """Non-orthorhombic cell: the pair test takes the Minkowski-reduced 27-image minimum; cut-offs below and above
    half the shortest cell vector (naive and general branches of find_mic in the full evaluation)."""
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=60, box=12.2, n_frames=2, seed=5, frame_sigma=0.1)
    tri = np.array([[12.2, 0.0, 0.0], [3.1, 11.8, 0.0], [-2.0, 2.6, 12.5]])
    for cutoff in (5.5, 7.5):
        cell = mb.Cell(tri, cutoff=cutoff, pbc=True)
        masks = _masks(ctx, Z, R, eids, cids, b200_models(_dicts(), [6.0, 10.0]), cell)
        _assert_same(masks)


@pytest.mark.parametrize("desc,cut", [("com", [5.0, 9.0]), ("pair", [5.5, 6.5])])
def test_cluster_masks_equal(ctx, desc, cut):
    """Non-periodic: |CM_a - CM_b| < hi (com_distance_sum) and |p_a - p_b| < hi (max_atom_pair_dist) as pair tests."""
    Z, R, eids, cids = synthetic.make_cluster(125, seed=2)
    masks = _masks(ctx, Z, R, eids, cids, b200_models(_dicts(), cut, desc=desc), None)
    _assert_same(masks)
    # a range bound keeps its upper end as the pair test; a lower bound has none (full enumeration either way)
    masks = _masks(ctx, Z, R, eids, cids, b200_models(_dicts(), [(3.0, c) for c in cut], desc=desc), None)
    _assert_same(masks)
    masks = _masks(ctx, Z, R, eids, cids, b200_models(_dicts(), cut, desc=desc, bounds=["lower"] * 2), None)
    for full, nl in zip(masks["full"], masks["pairlist"]):
        assert np.array_equal(full, nl)


def test_periodic_max_atom_pair_dist_masks_equal(ctx):
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=140, box=16.12, n_frames=1, seed=7, frame_sigma=0.1)
    masks = _masks(ctx, Z, R, eids, cids, b200_models(_dicts(), [5.0, 6.0], desc="pair"), mb.Cell(cell_v, pbc=True))
    _assert_same(masks)


def _predict(ctx, kind, Z, R, eids, cids, models, cell, al, shard=(0, 1), virial=True):
    ctx.set_culling(kind)
    pred = mb.mbePredict(models, mb.predict_gdml, alchemy_scalers=al, periodic_cell=cell)
    srcs = [mb.gen_combs(pred.get_avail_entities(cids, m.comp_ids)) for m in models]
    return _engine.predict_total(ctx, Z, R, eids, models, srcs, [pred._scalers_for(m) for m in models], cell,
                                 compute_virial=virial, shard=shard)


def test_predictions_equal_plain_sharded_and_step(ctx):
    """Whole predictions of a 3-frame periodic box: counts identical, E / F / virial equal to rounding (the FP64
    atomics of the scatter are the only thing whose order can differ); two shards sum to the whole; a prepared step
    created under MBG_CULL_PAIRLIST (device-side reduced count) replays to the same numbers."""
    dicts = [synthetic.make_model(f, n_train=40, sig=s, seed=k + 1) for k, (f, s) in enumerate(zip(FAMILIES, [100.0, 400.0]))]
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=140, box=16.12, n_frames=3, seed=1, frame_sigma=0.1)
    cell = mb.Cell(cell_v, pbc=True)
    models = b200_models(dicts, [6.0, 10.0])
    al = [mb.mbeAlchemyScale(0, 3, mb.linear_switching, {"factor": 0.5})]
    E0, F0, V0, st0 = _predict(ctx, "full", Z, R, eids, cids, models, cell, al)
    E1, F1, V1, st1 = _predict(ctx, "pairlist", Z, R, eids, cids, models, cell, al)
    assert st0 == st1 and st1[1][2] > 3 * 20000
    assert relerr(E1, E0) <= 1e-12 and relerr(F1, F0) <= 1e-12 and relerr(V1, V0) <= 1e-12
    parts = [_predict(ctx, "pairlist", Z, R, eids, cids, models, cell, al, shard=(r, 2)) for r in range(2)]
    assert sum(p[3][1][2] for p in parts) == st0[1][2] and min(p[3][1][2] for p in parts) > 0
    assert sum(p[3][1][0] for p in parts) == st0[1][0] and sum(p[3][1][1] for p in parts) == st0[1][1]
    assert relerr(parts[0][0] + parts[1][0], E0) <= 1e-12 and relerr(parts[0][1] + parts[1][1], F0) <= 1e-12
    # prepared step (CUDA graph): the choice in force at creation
    pred = mb.mbePredict(models, mb.predict_gdml, alchemy_scalers=al, periodic_cell=cell)
    srcs = [mb.gen_combs(pred.get_avail_entities(cids, m.comp_ids)) for m in models]
    step = _engine.make_step(ctx, Z, eids, models, srcs, [pred._scalers_for(m) for m in models], cell, 3, compute_virial=True)
    cells, cutoff = cell.native_cell()
    for Rk in (R, R[::-1].copy()):
        Es, Fs, Vs, sts = step.run(Rk, cells, np.array([cutoff]))
        Ek, Fk, Vk, stk = _predict(ctx, "full", Z, Rk, eids, cids, models, cell, al)
        assert [s[2] for s in sts] == [s[2] for s in stk]
        assert relerr(Es, Ek) <= 1e-12 and relerr(Fs, Fk) <= 1e-12 and relerr(Vs, Vk) <= 1e-12


# ---- heterogeneous product spaces ([h2o, meoh], [h2o, h2o, meoh], ...): one bit row per slot ------------------------
MIX_FAMILIES = [["h2o", "meoh"], ["h2o", "h2o", "meoh"], ["h2o", "meoh", "meoh"], ["meoh", "meoh", "meoh"]]
MIX_COMP = ["h2o", "meoh"] * 30 + ["h2o"] * 50 + ["meoh"] * 10  # interleaved ids: the product spaces hold non-ascending tuples


def _mix_dicts(n_train=4):
    return [synthetic.make_model(f, n_train=n_train, sig=100.0 + 50 * k, seed=k + 11) for k, f in enumerate(MIX_FAMILIES)]


def test_mixed_box_masks_equal(ctx):
    """Water-methanol box (config 5's kind of system), entity ids interleaved: 0 (not ascending) / 1 / 2 of every
    tuple of every product space, two frames."""
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=len(MIX_COMP), box=17.8, n_frames=2, seed=0, comp=MIX_COMP,
                                                           frame_sigma=0.1)
    masks = _masks(ctx, Z, R, eids, cids, b200_models(_mix_dicts(), [6.5, 10.5, 11.0, 11.5]), mb.Cell(cell_v, cutoff=7.5, pbc=True))
    _assert_same(masks)
    assert all(int((m == 0).sum()) > 0 for m in masks["pairlist"][:3])


@pytest.mark.parametrize("desc,cut", [("com", [6.0, 9.5, 10.0, 10.5]), ("pair", [6.5, 8.0, 8.5, 9.0])])
def test_mixed_cluster_masks_equal(ctx, desc, cut):
    Z, R, eids, cids = synthetic.make_cluster(len(MIX_COMP), seed=4, spacing=3.6, comp=MIX_COMP)
    masks = _masks(ctx, Z, R, eids, cids, b200_models(_mix_dicts(), cut, desc=desc), None)
    _assert_same(masks)


def test_mixed_predictions_equal_plain_sharded_and_step(ctx):
    dicts = _mix_dicts(n_train=24)
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=len(MIX_COMP), box=17.8, n_frames=2, seed=2, comp=MIX_COMP,
                                                           frame_sigma=0.1)
    cell = mb.Cell(cell_v, pbc=True)
    models = b200_models(dicts, [6.5, 10.5, 11.0, 11.5])
    al = [mb.mbeAlchemyScale(1, 3, mb.linear_switching, {"factor": 0.5})]
    E0, F0, V0, st0 = _predict(ctx, "full", Z, R, eids, cids, models, cell, al)
    E1, F1, V1, st1 = _predict(ctx, "pairlist", Z, R, eids, cids, models, cell, al)
    assert st0 == st1 and all(s[2] > 0 for s in st1)
    assert relerr(E1, E0) <= 1e-12 and relerr(F1, F0) <= 1e-12 and relerr(V1, V0) <= 1e-12
    parts = [_predict(ctx, "pairlist", Z, R, eids, cids, models, cell, al, shard=(r, 3)) for r in range(3)]
    for k in range(len(models)):
        assert [sum(p[3][k][i] for p in parts) for i in range(3)] == list(st0[k])
    assert relerr(sum(p[0] for p in parts), E0) <= 1e-12 and relerr(sum(p[1] for p in parts), F0) <= 1e-12
    pred = mb.mbePredict(models, mb.predict_gdml, alchemy_scalers=al, periodic_cell=cell)
    srcs = [mb.gen_combs(pred.get_avail_entities(cids, m.comp_ids)) for m in models]
    step = _engine.make_step(ctx, Z, eids, models, srcs, [pred._scalers_for(m) for m in models], cell, 2, compute_virial=True)
    cells, cutoff = cell.native_cell()
    Es, Fs, Vs, sts = step.run(R, cells, np.array([cutoff]))
    assert [tuple(s) for s in sts] == [tuple(s) for s in st0]
    assert relerr(Es, E0) <= 1e-12 and relerr(Fs, F0) <= 1e-12 and relerr(Vs, V0) <= 1e-12


def test_public_api_512_h2o_auto_switch():
    """MBG_CULL_AUTO through the public API: a 512-H2O frame has 22.2 M trimer candidates (> 2^21), so the plain call
    (first `predict`) and the prepared step (second `predict`: graph replay, reduced count on the device) both go
    through the pair list; the same prediction with the full enumeration forced is the reference point."""
    ctx = _native.get_context()
    box = 16.12 * (512 / 140) ** (1 / 3)
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=512, box=box, n_frames=1, seed=9, frame_sigma=0.1)
    dicts = [synthetic.make_model(["h2o"] * k, n_train=16, sig=s, seed=k) for k, s in ((1, 10.0), (2, 100.0), (3, 400.0))]
    models = b200_models(dicts, [None, 6.0, 10.0])
    cell = mb.Cell(cell_v, cutoff=8.0, pbc=True)
    try:
        ctx.set_culling("full")
        ref = mb.mbePredict(models, mb.predict_gdml, periodic_cell=cell)
        ref.use_graph = False
        E0, F0 = ref.predict(Z, R, eids, cids)
        st0 = list(ref.last_stats)
        ctx.set_culling("auto")
        pred = mb.mbePredict(models, mb.predict_gdml, periodic_cell=cell)
        for _ in range(3):
            E, F = pred.predict(Z, R, eids, cids)
            assert [tuple(s) for s in pred.last_stats] == [tuple(s) for s in st0]
            assert relerr(E, E0) <= 1e-12 and relerr(F, F0) <= 1e-12
    finally:
        ctx.set_culling("auto")
    assert st0[2][0] == 22238720 and 50000 < st0[2][2] < 100000


def test_four_body_masks_equal_and_match_the_oracle(ctx):
    """Homogeneous 4-body job (12-atom water tetramers): a row's reduced candidates are the C(deg, 3) triples of its
    neighbours, unranked with the job's binomial table; periodic box (vs the oracle culling too) and cluster."""
    md = synthetic.make_model(["h2o"] * 4, n_train=4, sig=500.0, seed=31)
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=48, box=11.3, n_frames=2, seed=6, frame_sigma=0.1)
    cell = mb.Cell(cell_v, pbc=True)
    masks = _masks(ctx, Z, R, eids, cids, b200_models([md], [13.0]), cell)
    _assert_same(masks)
    assert masks["full"][0].shape == (2, 194580)
    om = oracle_models([md], [13.0])[0]
    combs = orc.gen_combs_array([list(range(48))] * 4)
    want = cull_vec.accept_mask(Z, R[1], eids, combs, om, orc.Cell(cell_v))
    assert np.array_equal(masks["pairlist"][0][1] == 1, want) and 0 < want.sum() < len(combs)
    Zc, Rc, ec, cc = synthetic.make_cluster(40, seed=8)
    for desc, cut in (("com", 11.0), ("pair", 7.0)):
        _assert_same(_masks(ctx, Zc, Rc, ec, cc, b200_models([md], [cut], desc=desc), None))
