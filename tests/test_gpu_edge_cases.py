"""GPU parity on the edge cases of the path: nothing accepted, no frames, a model whose component is absent,
interleaved (non-contiguous) entity atoms, a one-row training set, a single-entity system.  Oracle on the same
seeded inputs; tolerance 1e-9 relative (max|delta| / max|ref|), accept counts exact."""
import numpy as np
import pytest

import mbgdml_b200 as mb
from helpers import b200_models, oracle_models, relerr
from mbgdml_b200 import synthetic
from oracle import mbgdml_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _water_models(T=24, seed=0):
    return [synthetic.make_model(["h2o"] * k, n_train=T, sig=[10.0, 60.0, 150.0][k - 1], seed=seed + k) for k in (1, 2, 3)]


def test_nothing_accepted_gives_zeros_and_all_nan_rows():
    mds = _water_models()
    Z, R, eids, cids = synthetic.make_cluster(5, seed=3)
    cuts = [None, 1e-3, 1e-3]  # no dimer / trimer can pass
    pred = mb.mbePredict(b200_models(mds, cuts), mb.predict_gdml)
    E, F = pred.predict(Z, R, eids, cids)
    Eo, Fo = orc.mbe_predict(oracle_models(mds, cuts), Z, R, eids, cids)
    assert relerr(E, Eo) <= TOL and relerr(F, Fo) <= TOL
    assert [s[2] for s in pred.last_stats] == [5, 0, 0]
    Ed, Fd, Cd, _ = mb.mbePredict(b200_models(mds, cuts), mb.predict_gdml_decomp).predict_decomp(Z, R, eids, cids)
    assert np.isnan(Ed[1]).all() and np.isnan(Fd[2]).all() and not np.isnan(Ed[0]).any()


def test_zero_frames_and_single_entity_system():
    mds = _water_models()
    Z, R, eids, cids = synthetic.make_cluster(4, seed=1)
    pred = mb.mbePredict(b200_models(mds, [None] * 3), mb.predict_gdml)
    E, F = pred.predict(Z, np.zeros((0,) + R.shape), eids, cids)
    assert E.shape == (0,) and F.shape == (0,) + R.shape
    # one molecule: only the 1-body model has fragments (gen_combs of a 2- or 3-fold product of one id is empty)
    Z1, R1, e1, c1 = synthetic.make_cluster(1, seed=2)
    E, F = pred.predict(Z1, R1, e1, c1)
    Eo, Fo = orc.mbe_predict(oracle_models(mds, [None] * 3), Z1, R1, e1, c1)
    assert relerr(E, Eo) <= TOL and relerr(F, Fo) <= TOL
    assert [s[1] for s in pred.last_stats] == [1, 0, 0]


def test_model_component_absent_from_the_structure():
    mds = _water_models() + [synthetic.make_model(["meoh"], n_train=16, sig=20.0, seed=9),
                             synthetic.make_model(["h2o", "meoh"], n_train=16, sig=50.0, seed=10)]
    Z, R, eids, cids = synthetic.make_cluster(4, seed=5)  # water only: the methanol models find no entities
    pred = mb.mbePredict(b200_models(mds, [None] * 5), mb.predict_gdml)
    E, F = pred.predict(Z, R, eids, cids)
    Eo, Fo = orc.mbe_predict(oracle_models(mds, [None] * 5), Z, R, eids, cids)
    assert relerr(E, Eo) <= TOL and relerr(F, Fo) <= TOL
    assert [s[1] for s in pred.last_stats] == [4, 6, 4, 0, 0]


def test_interleaved_entity_atoms():
    """entity_ids need not be contiguous: r_slice = np.where(entity_ids == e)[0] per entity, in comb order
    (gdml_predict.py:214-223).  Atoms of the cluster are shuffled molecule-wise so that every entity's atoms
    keep their internal order (O, H, H) but are spread over the array."""
    mds = _water_models()
    Z, R, eids, cids = synthetic.make_cluster(6, seed=7)
    rng = np.random.default_rng(0)
    # a random interleaving that preserves the order of atoms inside each entity
    order = np.argsort(rng.permutation(len(Z)) // 1e9 + np.tile(np.arange(3), 6) + rng.uniform(0, 0.9, len(Z)), kind="stable")
    Zs, Rs, es = Z[order], R[order], eids[order]
    assert not np.all(np.diff(es) >= 0)
    cuts = [None, 7.0, 11.0]
    E, F = mb.mbePredict(b200_models(mds, cuts), mb.predict_gdml).predict(Zs, Rs, es, cids)
    Eo, Fo = orc.mbe_predict(oracle_models(mds, cuts), Zs, Rs, es, cids)
    assert relerr(E, Eo) <= TOL and relerr(F, Fo) <= TOL
    # and the result is the un-shuffled one, permuted
    E0, F0 = mb.mbePredict(b200_models(mds, cuts), mb.predict_gdml).predict(Z, R, eids, cids)
    assert relerr(E, E0) <= TOL and relerr(F[0], F0[0][order]) <= TOL


@pytest.mark.parametrize("T", [1, 7, 33])
def test_tiny_training_sets(T):
    mds = _water_models(T=T, seed=20)
    Z, R, eids, cids = synthetic.make_cluster(4, seed=8)
    E, F = mb.mbePredict(b200_models(mds, [None] * 3), mb.predict_gdml).predict(Z, R, eids, cids)
    Eo, Fo = orc.mbe_predict(oracle_models(mds, [None] * 3), Z, R, eids, cids)
    assert relerr(E, Eo) <= TOL and relerr(F, Fo) <= TOL


def test_tail_split_launch_shape():
    """A launch of a little more than one wave: 969 water trimers x 48 permutations = 162 CTAs of 288 queries on
    148 SMs.  The 14 CTAs of the last wave are cut into training-row slices (api.cu, tail splitting); the result
    must not change.  Oracle on the same inputs (3-body model only, T = 250: 8 row tiles -> 4 tail slices)."""
    md = synthetic.make_model(["h2o"] * 3, n_train=250, sig=150.0, seed=5)
    Z, R, eids, cids = synthetic.make_cluster(19, seed=4)
    pred = mb.mbePredict(b200_models([md], [None]), mb.predict_gdml)
    E, F = pred.predict(Z, R, eids, cids)
    assert pred.last_stats[0][2] == 969
    Eo, Fo = orc.mbe_predict(oracle_models([md], [None]), Z, R, eids, cids)
    assert relerr(E, Eo) <= TOL and relerr(F, Fo) <= TOL


def test_d_mic_operator_matches_reference_semantics():
    """Cell.d_mic (periodic.py:61-84): the cut-off test runs over the pair distances of the imaged vectors only -- two
    vectors longer than the cut-off but close to each other are accepted, and check_cutoff=False always returns them."""
    from oracle import mbgdml_oracle as orc

    rng = np.random.default_rng(0)
    for cell_v in (np.eye(3) * 14.0, np.array([[12.6, 0.0, 0.0], [2.8, 12.1, 0.0], [-1.9, 3.1, 11.8]])):
        cell, ocell = mb.Cell(cell_v, cutoff=5.0, pbc=True), orc.Cell(cell_v, cutoff=5.0)
        far_but_close = np.array([[6.0, 0.5, 0.2], [6.2, 0.1, -0.4]])  # |d| > cutoff, |d0 - d1| < cutoff (6.3 would tie two images)
        for d in (far_but_close, far_but_close + cell_v[0] - 2 * cell_v[2], rng.uniform(-20, 20, (4, 3)), rng.uniform(-2, 2, (3, 3))):
            got, want = cell.d_mic(d), ocell.d_mic(d)
            assert (got is None) == (want is None)
            if want is not None:
                assert np.abs(got - want).max() <= 1e-12
            free = cell.d_mic(d, check_cutoff=False)
            assert np.abs(free - ocell.d_mic(d, check_cutoff=False)).max() <= 1e-12
        assert cell.d_mic(far_but_close) is not None
