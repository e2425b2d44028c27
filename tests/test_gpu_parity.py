"""GPU parity: the CUDA path (through the C ABI / reference-shaped Python surface) against the
reference outputs committed in tests/golden/ and against the NumPy oracle on the same seeded inputs.

Tolerance (north_star): energies/forces within 1e-9 relative on random-init models, measured as
max|delta| / max|reference|; fragment index sets and accept masks bit-exact.  The reference's trained
models are ill-conditioned, so for them the reference's own tolerance applies.
"""
import numpy as np
import pytest

import mbgdml_b200 as mb
from helpers import b200_models, cutoffs_from, load, oracle_models, relerr, unpack_models
from mbgdml_b200 import _native, synthetic
from oracle import mbgdml_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9  # north_star bar; observed errors are ~1e-13


@pytest.fixture(autouse=True, params=["dmma", "fma"])
def contraction(request):
    """Every parity test runs against both contraction kernels (tensor-pipe GEMM recast and FMA-pipe
    difference form); the C ABI selects one per context (mbg_context_set_contraction)."""
    ctx = _native.get_context()
    ctx.set_contraction(request.param)
    yield request.param
    ctx.set_contraction("auto")


def alch(rows):
    return [mb.mbeAlchemyScale(int(e), int(o), mb.linear_switching, {"factor": float(f)}) for e, o, f in rows]


# ------------------------------------------------------------------------------------------
def test_enumerate_bit_exact(gpu_ctx):
    g = load("gen_combs.npz")
    for k in range(int(g["n"])):
        order = int(g[f"s{k}_order"])
        sets = [g[f"s{k}_set{j}"] for j in range(order)]
        got = gpu_ctx.enumerate(sets)
        assert got.shape == g[f"s{k}_combs"].shape
        assert np.array_equal(got, g[f"s{k}_combs"])
    # python-facing generator yields the same tuples lazily
    assert list(mb.gen_combs([[0, 2], [1, 3]])) == [(0, 1), (0, 3), (2, 3)]
    assert list(mb.gen_combs([[2, 3], [0, 1]])) == []


def test_enumerate_large_matches_oracle(gpu_ctx):
    sets = [np.arange(140)] * 3  # BASELINE config 4: C(140,3) = 447,580 trimers out of 2,744,000 tuples
    got = gpu_ctx.enumerate(sets)
    assert got.shape == (447580, 3)
    assert np.array_equal(got, orc.gen_combs_array(sets))
    sets = [np.arange(0, 100), np.arange(50, 100), np.arange(100, 140)]  # heterogeneous, overlapping
    assert np.array_equal(gpu_ctx.enumerate(sets), orc.gen_combs_array(sets))


@pytest.mark.parametrize("tag", ["1b", "2b", "3b"])
def test_single_fragment_kernel(tag):
    """_predict_gdml_wkr on single fragments (reference outputs), through predict_gdml."""
    g = load("fragment_kernel.npz")
    md = unpack_models(g, f"{tag}_m")[0]
    model = mb.gdmlModel(dict(md))
    n = len(md["z"])
    eids = np.asarray(md["entity_ids"])
    for r, out_ref in zip(g[f"{tag}_r"], g[f"{tag}_out"]):
        E, F = mb.predict_gdml(md["z"], r, eids, [tuple(range(model.nbody_order))], model)
        e_ref = out_ref[0] * md["std"] + md["c"]
        f_ref = out_ref[1:].reshape(n, 3) * md["std"]
        assert abs(E - e_ref) <= TOL * max(abs(e_ref), 1.0)
        assert relerr(F, f_ref) <= TOL


def test_cluster_variants():
    g = load("cluster_6h2o.npz")
    mds = unpack_models(g)
    Z, R, eids, cids = g["Z"], g["R"], g["entity_ids"], g["comp_ids"]
    E, F = mb.mbePredict(b200_models(mds, [None] * 3), mb.predict_gdml).predict(Z, R, eids, cids)
    assert relerr(E, g["plain_E"]) <= TOL and relerr(F, g["plain_F"]) <= TOL
    cuts = cutoffs_from(g["crit_cutoffs"])
    pred = mb.mbePredict(b200_models(mds, cuts), mb.predict_gdml)
    E, F = pred.predict(Z, R, eids, cids)
    assert relerr(E, g["crit_E"]) <= TOL and relerr(F, g["crit_F"]) <= TOL
    # accept masks: accepted counts equal the number of non-NaN rows of the reference's decomposition
    for k in range(3):
        assert pred.last_stats[k][2] == np.count_nonzero(~np.isnan(g[f"decomp_E{k}"]))
        assert pred.last_stats[k][1] == len(g[f"decomp_E{k}"])
    E, F = mb.mbePredict(b200_models(mds, cuts), mb.predict_gdml, alchemy_scalers=alch(g["alch"])).predict(
        Z, R, eids, cids)
    assert relerr(E, g["alch_E"]) <= TOL and relerr(F, g["alch_F"]) <= TOL
    E, F = mb.mbePredict(
        b200_models(mds, [None, (3.0, 6.0), 5.5], desc="pair", bounds=["upper", "upper", "lower"]),
        mb.predict_gdml).predict(Z, R, eids, cids)
    assert relerr(E, g["pair_E"]) <= TOL and relerr(F, g["pair_F"]) <= TOL
    mdsE = [dict(m) for m in mds]
    mdsE[1]["alphas_E"] = g["useE_alphas_E"]
    E, F = mb.mbePredict(b200_models(mdsE, cuts), mb.predict_gdml).predict(Z, R, eids, cids)
    assert relerr(E, g["useE_E"]) <= TOL and relerr(F, g["useE_F"]) <= TOL


def test_single_frame_2d_input_and_per_model_plugin_path():
    """R.ndim == 2 is accepted (mbe.py:1035-1036); compute_nbody drives the plugin per model like the
    reference (mbe.py:809-811) and sums to the batched result."""
    g = load("cluster_6h2o.npz")
    mds = unpack_models(g)
    Z, R, eids, cids = g["Z"], g["R"], g["entity_ids"], g["comp_ids"]
    models = b200_models(mds, cutoffs_from(g["crit_cutoffs"]))
    pred = mb.mbePredict(models, mb.predict_gdml)
    E, F = pred.predict(Z, R[1], eids, cids)
    assert E.shape == (1,) and F.shape == (1,) + R[1].shape
    assert relerr(E, g["crit_E"][1:2]) <= TOL and relerr(F, g["crit_F"][1:2]) <= TOL
    e_sum, f_sum = 0.0, np.zeros(R[1].shape)
    for m in models:
        e, f = pred.compute_nbody(Z, R[1], eids, cids, m)
        e_sum += e
        f_sum += f
    assert abs(e_sum - g["crit_E"][1]) <= TOL * abs(g["crit_E"][1]) and relerr(f_sum, g["crit_F"][1]) <= TOL
    with pytest.raises(ValueError):  # mbe.py:777
        pred.compute_nbody(Z, R, eids, cids, models[0])
    # explicit tuple chunk, as the reference's ray branch passes (mbe.py:823-828)
    e, f = mb.predict_gdml(Z, R[1], eids, ((0, 1), (2, 5), (3, 4)), models[1], None)
    eo, fo = orc.predict_gdml(Z, R[1], eids, ((0, 1), (2, 5), (3, 4)), oracle_models(mds, cutoffs_from(g["crit_cutoffs"]))[1])
    assert abs(e - eo) <= TOL * max(abs(eo), 1) and relerr(f, fo) <= TOL


def test_cluster_decomp():
    g = load("cluster_6h2o.npz")
    mds = unpack_models(g)
    Z, R, eids, cids = g["Z"], g["R"], g["entity_ids"], g["comp_ids"]
    pred = mb.mbePredict(b200_models(mds, cutoffs_from(g["crit_cutoffs"])), mb.predict_gdml_decomp)
    Ed, Fd, Cd, orders = pred.predict_decomp(Z, R, eids, cids)
    assert orders == [1, 2, 3]
    for k in range(3):
        assert np.array_equal(Cd[k], g[f"decomp_C{k}"])  # bit-exact fragment index sequence
        assert np.array_equal(np.isnan(Ed[k]), np.isnan(g[f"decomp_E{k}"]))  # bit-exact accept mask
        assert np.array_equal(np.isnan(Fd[k]), np.isnan(g[f"decomp_F{k}"]))
        assert relerr(np.nan_to_num(Ed[k]), np.nan_to_num(g[f"decomp_E{k}"])) <= TOL
        assert relerr(np.nan_to_num(Fd[k]), np.nan_to_num(g[f"decomp_F{k}"])) <= TOL


def test_periodic_boxes():
    g = load("periodic_40h2o.npz")
    mds = unpack_models(g)
    cuts = cutoffs_from(g["crit_cutoffs"])
    Z, eids, cids = g["cub_Z"], g["cub_eids"], g["cub_cids"]
    models = b200_models(mds, cuts)
    # cubic, default cutoff L/2, alchemy, group virial -> stress
    pred = mb.mbePredict(models, mb.predict_gdml, alchemy_scalers=alch(g["alch"]),
                         periodic_cell=mb.Cell(g["cub_cell"], pbc=True), compute_stress=True)
    E, F, S = pred.predict(Z, g["cub_R"], eids, cids)
    assert relerr(E, g["cub_E"]) <= TOL and relerr(F, g["cub_F"]) <= TOL and relerr(S, g["cub_stress"]) <= TOL
    # accept mask: same number of surviving fragments as the reference path for frame 0
    pred1 = mb.mbePredict(models, mb.predict_gdml, periodic_cell=mb.Cell(g["cub_cell"], pbc=True))
    pred1.predict(Z, g["cub_R"][0], eids, cids)
    for k in range(3):
        assert pred1.last_stats[k][2] == int(g[f"cub_naccept{k}"])
    # user cutoff, Voigt + isotropic normal stress post-processing (mbe.py:1081-1094)
    pred = mb.mbePredict(models, mb.predict_gdml, periodic_cell=mb.Cell(g["cub_cell"], cutoff=5.0, pbc=True),
                         compute_stress=True)
    pred.use_voigt, pred.only_normal_stress, pred.box_scaling_type = True, True, "isotropic"
    E, F, S = pred.predict(Z, g["cub_R"], eids, cids)
    assert S.shape == g["cut5_stress"].shape
    assert relerr(E, g["cut5_E"]) <= TOL and relerr(F, g["cut5_F"]) <= TOL and relerr(S, g["cut5_stress"]) <= TOL
    # triclinic cell
    pred = mb.mbePredict(models, mb.predict_gdml, alchemy_scalers=alch(g["alch"]),
                         periodic_cell=mb.Cell(g["tri_cell"], pbc=True), compute_stress=True)
    E, F, S = pred.predict(Z, g["tri_R"], eids, cids)
    assert relerr(E, g["tri_E"]) <= TOL and relerr(F, g["tri_F"]) <= TOL and relerr(S, g["tri_stress"]) <= TOL
    # cutoff above half the shortest cell vector: the general (Minkowski) minimum-image branch decides
    pred = mb.mbePredict(models, mb.predict_gdml,
                         periodic_cell=mb.Cell(g["tri_cell"], cutoff=float(g["tribig_cutoff"]), pbc=True))
    E, F = pred.predict(Z, g["tri_R"], eids, cids)
    assert relerr(E, g["tribig_E"]) <= TOL and relerr(F, g["tribig_F"]) <= TOL


def test_r_mic_operator():
    g = load("periodic_40h2o.npz")
    rng = np.random.default_rng(0)
    for cell_v, cutoff in ((g["cub_cell"], None), (g["tri_cell"], None), (g["tri_cell"], float(g["tribig_cutoff"]))):
        ocell, bcell = orc.Cell(cell_v, cutoff=cutoff), mb.Cell(cell_v, cutoff=cutoff)
        n_none = 0
        for _ in range(40):
            start = rng.integers(0, 38) * 3
            r = g["cub_R"][0][start:start + 9] if cell_v is g["cub_cell"] else g["tri_R"][0][start:start + 9]
            want, got = ocell.r_mic(r), bcell.r_mic(r)
            assert (want is None) == (got is None)
            if want is None:
                n_none += 1
            else:
                assert np.allclose(got, want, rtol=0, atol=1e-12)
        assert 0 < n_none < 40


def test_mixed_water_methanol(contraction):
    """BASELINE config 5 shape: heterogeneous comp_ids, 8 models with fragments of 3..15 atoms (D up to 105),
    alchemical 3-body factor; reference outputs from the golden fixture.  Fragments above 9 atoms exist on the
    tensor-pipe kernel only: with the FMA-pipe kernel selected they must raise (no CPU fallback)."""
    g = load("mixed_h2o_meoh.npz")
    mds = unpack_models(g)
    assert sorted(len(m["z"]) for m in mds) == [3, 6, 6, 9, 9, 12, 12, 15]
    for tag in ("", "2"):
        Z, R, eids, cids = g["Z" + tag], g["R" + tag], g["entity_ids" + tag], g["comp_ids" + tag]
        al_rows = g["alch"] if tag == "" else []
        pred = mb.mbePredict(b200_models(mds, [None] * len(mds)), mb.predict_gdml, alchemy_scalers=alch(al_rows))
        if contraction == "fma":
            with pytest.raises(NotImplementedError):
                pred.predict(Z, R, eids, cids)
            small = [m for m in mds if len(m["z"]) <= 9]
            E, F = mb.mbePredict(b200_models(small, [None] * 5), mb.predict_gdml, alchemy_scalers=alch(al_rows)).predict(
                Z, R, eids, cids)
            Eo, Fo = orc.mbe_predict(oracle_models(small, [None] * 5), Z, R, eids, cids,
                                     alchemy_scalers=[orc.AlchemyScale(int(e), int(o), float(f)) for e, o, f in al_rows])
            assert relerr(E, Eo) <= TOL and relerr(F, Fo) <= TOL
            continue
        E, F = pred.predict(Z, R, eids, cids)
        assert relerr(E, g["E" + tag]) <= TOL and relerr(F, g["F" + tag]) <= TOL  # live reference outputs


def test_widest_fragments_three_methanols(contraction):
    """18-atom fragments (D = 153, P capped at 100 like _gdml/perm.py:275): the widest kernel instantiation."""
    md = synthetic.make_model(["meoh", "meoh", "meoh"], n_train=21, sig=300.0, seed=12, c=0.7, std=1.3)
    assert md["perms"].shape == (100, 18)
    Z, R, eids, cids = synthetic.make_cluster(4, seed=3, spacing=4.2, comp=["meoh"] * 4)
    pred = mb.mbePredict([mb.gdmlModel(dict(md))], mb.predict_gdml)
    if contraction == "fma":
        with pytest.raises(NotImplementedError):
            pred.predict(Z, R, eids, cids)
        return
    E, F = pred.predict(Z, R, eids, cids)
    Eo, Fo = orc.mbe_predict([orc.Model(dict(md))], Z, R[None], eids, cids)
    assert pred.last_stats[0][2] == 4
    assert relerr(E, Eo) <= TOL and relerr(F, Fo) <= TOL


def test_reference_known_answers_on_device():
    """reference tests/test_descriptors.py:40-49 (bit-exact) and tests/test_predictsets.py:80-142."""
    g = load("ref_2h2o_criteria.npz")
    v = mb.com_distance_sum(g["z"], g["R"], g["entity_ids"])
    assert np.array_equal(v, g["desc_v"])  # 120 dimers, bit for bit
    crit = mb.Criteria(mb.com_distance_sum, {"entity_ids": g["entity_ids"]}, cutoff=5.3809148637976385)
    acc, val = crit.accept(g["z"], g["R"][42])
    assert not acc and val == 5.3809148637976385
    # NaN index set: each dimer is its own 2-entity structure predicted with the trained 2-body model
    md = unpack_models(load("ref_models_t500.npz"))[1]
    model = mb.gdmlModel(dict(md), criteria=mb.Criteria(mb.com_distance_sum, {"entity_ids": g["entity_ids"]}, 6.0))
    pred = mb.mbePredict([model], mb.predict_gdml_decomp)
    Ed, _, _, _ = pred.predict_decomp(g["z"], g["R"], g["entity_ids"], np.array(["h2o", "h2o"]))
    assert np.array_equal(np.where(np.isnan(Ed[0]))[0], g["nan_idx"])


def test_reference_16mer_and_6mer_with_trained_models():
    """reference tests/test_predict.py:40-128: pinned E and forces at the reference's own tolerance
    (the trained models are ill-conditioned: cond ~1e15)."""
    mds = unpack_models(load("ref_models_t500.npz"))
    models = b200_models(mds, [None, 6.0, 10.0])
    g = load("ref_16mer.npz")
    E, F = mb.mbePredict(models, mb.predict_gdml).predict(g["Z"], g["R"], g["entity_ids"], g["comp_ids"])
    assert np.allclose(E, g["E_pinned"])
    assert np.allclose(E, g["E_live"]) and np.allclose(F, g["F_live"], rtol=1e-4, atol=1e-2)
    g = load("ref_6h2o.npz")
    pred = mb.mbePredict(models, mb.predict_gdml)
    E, F = pred.predict(g["Z"], g["R"], g["entity_ids"], g["comp_ids"])
    assert np.allclose(E, g["E_live"]) and np.allclose(F, g["F_live"], rtol=1e-4, atol=1e-2)
    # tests/test_predictsets.py:40-77: decomposed predictions add up to the totals
    Ed, Fd, Cd, _ = mb.mbePredict(models, mb.predict_gdml_decomp).predict_decomp(
        g["Z"], g["R"], g["entity_ids"], g["comp_ids"])
    E_tot, F_tot = np.zeros_like(E), np.zeros_like(F)
    for k in range(3):
        for e, f, c in zip(Ed[k], Fd[k], Cd[k]):
            if not np.isnan(e):
                E_tot[c[0]] += e
                F_tot[c[0], orc.r_slice(g["entity_ids"], c[1:])] += f
    assert np.allclose(E_tot, E) and np.allclose(F_tot, F, rtol=1e-4, atol=1e-2)


def test_sharded_partial_sums_add_up(gpu_ctx):
    """Three shards evaluated one after the other on one GPU sum to the unsharded result, and
    every fragment is evaluated exactly once."""
    from mbgdml_b200 import _engine

    g = load("periodic_40h2o.npz")
    mds = unpack_models(g)
    models = b200_models(mds, cutoffs_from(g["crit_cutoffs"]))
    Z, R, eids, cids = g["cub_Z"], g["cub_R"], g["cub_eids"], g["cub_cids"]
    cell = mb.Cell(g["cub_cell"], pbc=True)
    pred = mb.mbePredict(models, mb.predict_gdml, periodic_cell=cell)
    sources = lambda: [mb.gen_combs(pred.get_avail_entities(cids, m.comp_ids)) for m in models]
    E1, F1, _, st1 = _engine.predict_total(gpu_ctx, Z, R, eids, models, sources(), [[]] * 3, cell)
    E_sum, F_sum, acc = 0, 0, np.zeros(3, dtype=int)
    for r in range(3):
        E, F, _, st = _engine.predict_total(gpu_ctx, Z, R, eids, models, sources(), [[]] * 3, cell, shard=(r, 3))
        E_sum, F_sum = E_sum + E, F_sum + F
        acc += np.array([s[2] for s in st])
    assert np.array_equal(acc, [s[2] for s in st1])
    assert relerr(E_sum, E1) <= 1e-12 and relerr(F_sum, F1) <= 1e-12
    assert relerr(E1, g["cub_E"] * 0 + E1) == 0  # sanity: finite


def test_full_size_properties_64_h2o_trimers():
    """BASELINE config 3 at full size (64 H2O, T = 1000, 41,664 trimers): size-independent checks.
    (a) internal forces sum to zero; (b) invariance under a rigid translation + rotation of the cluster
    (energies equal, forces co-rotate); (c) relabelling the molecules permutes the forces;
    (d) a random sample of trimers evaluated by the oracle matches the decomposed rows."""
    md = synthetic.make_model(["h2o"] * 3, n_train=1000, sig=100.0, seed=2)
    Z, R, eids, cids = synthetic.make_cluster(64, seed=0)
    model = mb.gdmlModel(dict(md))
    pred = mb.mbePredict([model], mb.predict_gdml)
    E, F = pred.predict(Z, R, eids, cids)
    assert pred.last_stats[0][1] == 41664 and pred.last_stats[0][2] == 41664
    scale = np.abs(F).max()
    assert np.abs(F[0].sum(axis=0)).max() <= 1e-9 * scale * 64  # (a)
    rot = synthetic.random_rotations(np.random.default_rng(1), 1)[0]
    E2, F2 = pred.predict(Z, R @ rot.T + np.array([3.0, -2.0, 0.5]), eids, cids)  # (b)
    assert abs(E2[0] - E[0]) <= 1e-9 * abs(E[0])
    assert relerr(F2[0], F[0] @ rot.T) <= 1e-9
    perm = np.random.default_rng(2).permutation(64)  # (c) molecule k of the new listing is old perm[k]
    atom_perm = (perm[:, None] * 3 + np.arange(3)).ravel()
    E3, F3 = pred.predict(Z[atom_perm], R[atom_perm], eids, cids)
    assert abs(E3[0] - E[0]) <= 1e-9 * abs(E[0]) and relerr(F3[0], F[0][atom_perm]) <= 1e-9
    rng = np.random.default_rng(5)  # (d)
    combs = np.sort(np.array([rng.choice(64, 3, replace=False) for _ in range(12)]), axis=1)
    Ed, Fd, _ = mb.predict_gdml_decomp(Z, R, eids, combs, model)
    Eo, Fo, _ = orc.predict_gdml_decomp(Z, R, eids, combs, orc.Model(dict(md)))
    assert relerr(Ed, Eo) <= TOL and relerr(Fd, Fo) <= TOL


def test_full_size_properties_140_h2o_periodic_box(contraction):
    """BASELINE config 4 at full size (140 H2O, periodic L = 16.12 A, 1+2+3-body, criteria, alchemy, T = 1000;
    447,580 candidate trimers per frame): size-independent checks.
    (a) fragment counts: enumerated = C(140, k); accepted masks reproduce a vectorised NumPy restatement of the
        first-atom / O-O pre-filter bound and, on a random sample of candidates, the oracle's decision exactly;
    (b) internal forces sum to zero per frame; (c) shifting every atom by a lattice vector changes nothing;
    (d) a frame predicted inside a batch equals the same frame predicted alone;
    (e) a random sample of accepted trimers evaluated by the oracle matches the decomposed rows;
    (f) [dmma only, slow] a 160-frame batch (71.6 M candidates: two culling chunks) returns the single-frame result
        for every copy of the frame."""
    import math

    fam = [["h2o"], ["h2o", "h2o"], ["h2o", "h2o", "h2o"]]
    cuts = [None, 6.0, 10.0]
    dicts = [synthetic.make_model(f, n_train=1000, sig=s_, seed=k) for k, (f, s_) in enumerate(zip(fam, (10.0, 100.0, 400.0)))]
    Z, R, eids, cids, cell_v = synthetic.make_periodic_box(n_mol=140, box=16.12, n_frames=3, seed=0, frame_sigma=0.1)
    models = b200_models(dicts, cuts)
    al = [mb.mbeAlchemyScale(0, 3, mb.linear_switching, {"factor": 0.5}), mb.mbeAlchemyScale(0, 2, mb.linear_switching, {"factor": 0.5})]
    pred = mb.mbePredict(models, mb.predict_gdml, alchemy_scalers=al, periodic_cell=mb.Cell(cell_v, pbc=True))
    E, F = pred.predict(Z, R, eids, cids)
    st = pred.last_stats
    assert [s[1] for s in st] == [3 * math.comb(140, k) for k in (1, 2, 3)]  # (a) enumerated
    assert st[0][2] == 3 * 140 and 0 < st[1][2] < st[1][1] and 0 < st[2][2] < st[2][1]
    ocell = orc.Cell(cell_v)
    omodels = oracle_models(dicts, cuts)
    rng = np.random.default_rng(3)
    sample = [rng.choice(140, 3, replace=False) for _ in range(30)]
    o_xyz = R[0][0::3]  # oxygen of every water; neighbour lists by minimum-image O-O distance
    d = o_xyz[:, None, :] - o_xyz[None, :, :]
    d -= 16.12 * np.round(d / 16.12)
    near = np.argsort(np.linalg.norm(d, axis=-1), axis=1)
    for i in rng.choice(140, 30, replace=False):  # half of the sample: a molecule with two of its 8 nearest
        sample.append(np.concatenate(([i], rng.choice(near[i, 1:9], 2, replace=False))))
    sample = np.sort(np.array(sample), axis=1)
    # accept decisions of the sample: NaN rows of the periodic decomposition are not offered by the reference
    # (mbe.py:565), so the decision is taken through total predictions of single fragments
    n_acc = 0
    for comb in sample:
        e_gpu, f_gpu = mb.predict_gdml(Z, R[0], eids, [tuple(comb)], models[2], mb.Cell(cell_v, pbc=True))
        e_o, f_o = orc.predict_gdml(Z, R[0], eids, [tuple(comb)], omodels[2], ocell)[:2]
        assert (e_gpu == 0.0 and not f_gpu.any()) == (e_o == 0.0 and not f_o.any())  # same accept decision
        if e_o != 0.0:
            n_acc += 1
            assert abs(e_gpu - e_o) <= TOL * max(abs(e_o), 1e-3) and relerr(f_gpu, f_o) <= TOL  # (e)
    assert n_acc >= 10
    scale = np.abs(F).max()
    for i in range(3):  # (b)
        assert np.abs(F[i].sum(axis=0)).max() <= 1e-9 * scale * 140
    shift = 2 * cell_v[0] - cell_v[1] + 3 * cell_v[2]  # (c)
    E2, F2 = pred.predict(Z, R + shift, eids, cids)
    assert relerr(E2, E) <= 1e-9 and relerr(F2, F) <= 1e-9
    E1, F1 = pred.predict(Z, R[1], eids, cids)  # (d)
    assert relerr(E1[0], E[1]) <= 1e-11 and relerr(F1[0], F[1]) <= 1e-11
    if contraction == "dmma":  # (f)
        Rb = np.repeat(R[:1], 160, axis=0)
        pred3 = mb.mbePredict(models[2:], mb.predict_gdml, alchemy_scalers=al, periodic_cell=mb.Cell(cell_v, pbc=True))
        Eb, Fb = pred3.predict(Z, Rb, eids, cids)
        n_batch = pred3.last_stats[0][2]
        E0, F0 = pred3.predict(Z, R[0], eids, cids)
        assert n_batch == 160 * pred3.last_stats[0][2]
        assert relerr(Eb, np.repeat(E0, 160)) <= 1e-11 and relerr(Fb, np.repeat(F0, 160, axis=0)) <= 1e-11


@pytest.mark.parametrize("family,n_train", [
    (["hf"], 5), (["h2o"], 1), (["nh3"], 8), (["hf", "hf"], 9), (["ch4"], 31), (["hf", "h2o"], 33),
    (["meoh"], 64), (["h2o", "h2o"], 65), (["nh3", "h2o"], 7), (["nh3", "nh3"], 40), (["ch4", "h2o"], 17),
    (["h2o", "h2o", "h2o"], 100), (["hf", "hf", "ch4"], 23), (["h2o", "meoh"], 129),
    (["ch4", "ch4"], 12), (["ch4", "meoh"], 19), (["meoh", "meoh"], 33), (["nh3", "nh3", "ch4"], 9),
    (["nh3", "ch4", "ch4"], 26), (["ch4", "ch4", "ch4"], 8), (["ch4", "ch4", "meoh"], 41), (["ch4", "meoh", "meoh"], 15),
])
def test_every_fragment_size_and_ragged_train_sizes(family, n_train, contraction):
    """Fragments of 2..17 atoms (every kernel instantiation; 18 atoms: test_widest_fragments_three_methanols) with training-set sizes that are not multiples of
    the 8-row block / 32- and 64-row tiles, with and without alphas_E, against the NumPy oracle."""
    comps = list(dict.fromkeys(family))  # entity ids ascend in the model's component order (gen_combs keeps only ascending tuples)
    n_each = 4 if len(family) < 3 else 3
    comp_list = [c for c in comps for _ in range(n_each)]
    Z, R, eids, cids = synthetic.make_cluster(len(comp_list), seed=11, spacing=3.4, comp=comp_list)
    for use_E in (False, True):
        md = synthetic.make_model(family, n_train=n_train, sig=37.0, seed=3, c=-1.5, std=2.25, use_E=use_E)
        pred = mb.mbePredict([mb.gdmlModel(dict(md))], mb.predict_gdml)
        if contraction == "fma" and len(md["z"]) > 9:  # the FMA-pipe kernel exists for 2..9 atoms only
            with pytest.raises(NotImplementedError):
                pred.predict(Z, R, eids, cids)
            continue
        E, F = pred.predict(Z, R, eids, cids)
        Eo, Fo = orc.mbe_predict([orc.Model(dict(md))], Z, R[None], eids, cids)
        assert pred.last_stats[0][2] > 0
        assert relerr(E, Eo) <= TOL and relerr(F, Fo) <= TOL


def test_extreme_kernel_widths():
    """sig far below / above the descriptor scale: the exp argument clamp and the exponent scaling paths."""
    Z, R, eids, cids = synthetic.make_cluster(5, seed=5, spacing=3.0)
    for sig in (0.05, 0.5, 5000.0):
        md = synthetic.make_model(["h2o", "h2o"], n_train=50, sig=sig, seed=9)
        E, F = mb.mbePredict([mb.gdmlModel(dict(md))], mb.predict_gdml).predict(Z, R, eids, cids)
        Eo, Fo = orc.mbe_predict([orc.Model(dict(md))], Z, R[None], eids, cids)
        assert np.all(np.isfinite(F))
        assert relerr(E, Eo) <= TOL and relerr(F, Fo) <= TOL


def test_errors_surface_as_reference_exceptions():
    md = synthetic.make_model(["h2o", "h2o"], n_train=8)
    Z, R, eids, cids = synthetic.make_cluster(3)
    model = mb.gdmlModel(dict(md))
    with pytest.raises(ValueError):  # entity with the wrong number of atoms for this model
        mb.predict_gdml(Z, R, np.array([0, 0, 0, 0, 1, 2, 2, 2, 2]), [(0, 1)], model)
    with pytest.raises(AssertionError):  # gdml_predict.py:197
        mb.predict_gdml(Z, R[None], eids, [(0, 1)], model)
    with pytest.raises(NotImplementedError):
        bad = mb.gdmlModel(dict(md), criteria=mb.Criteria(lambda Z, R: np.zeros(1), {}, 1.0))
        mb.predict_gdml(Z, R, eids, [(0, 1)], bad)
    E, F = mb.predict_gdml(Z, R, eids, [], model)  # empty fragment list -> zeros (gdml_predict.py:198-199)
    assert E == 0.0 and not F.any()


class _FakeAtoms:
    """The five ase.Atoms methods mbeCalculator.calculate uses (ASE is not part of the image)."""

    def __init__(self, Z, R, cell):
        self.Z, self.R, self.cell = np.array(Z), np.array(R, dtype=float), np.array(cell, dtype=float)
        self.wrapped = False

    def wrap(self):  # ase.Atoms.wrap for an orthorhombic cell
        L = np.diag(self.cell)
        self.R = self.R - L * np.floor(self.R / L)
        self.wrapped = True

    def copy(self):
        return _FakeAtoms(self.Z, self.R, self.cell)

    def get_cell(self):
        return self.cell

    def get_atomic_numbers(self):
        return self.Z

    def get_positions(self):
        return self.R


def test_ase_calculator_front_end():
    """interfaces/ase.py:65-108: wrap, refresh the cell (cutoff reset to L/2), force stress, unit conversion."""
    from mbgdml_b200.interfaces import mbeCalculator

    g = load("periodic_40h2o.npz")
    mds = unpack_models(g)
    models = b200_models(mds, cutoffs_from(g["crit_cutoffs"]))
    Z, R, eids, cids, cell_v = g["cub_Z"], g["cub_R"][0], g["cub_eids"], g["cub_cids"], g["cub_cell"]
    pred = mb.mbePredict(models, mb.predict_gdml, periodic_cell=mb.Cell(cell_v, cutoff=3.0, pbc=True))
    calc = mbeCalculator(pred, parameters={"entity_ids": eids, "comp_ids": cids}, e_conv=0.5, f_conv=2.0)
    atoms = _FakeAtoms(Z, R + 2 * cell_v[0], cell_v)  # outside the box: calculate() wraps
    calc.calculate(atoms)
    assert atoms.wrapped and pred.compute_stress and pred.use_voigt
    assert pred.periodic_cell.cutoff == np.min(np.linalg.norm(cell_v, axis=1)) / 2  # user cutoff silently reset
    ref = mb.mbePredict(models, mb.predict_gdml, periodic_cell=mb.Cell(cell_v, pbc=True), compute_stress=True)
    ref.use_voigt = True
    E, F, S = ref.predict(Z, R, eids, cids)
    assert abs(calc.results["energy"] - 0.5 * E[0]) <= 1e-9 * abs(E[0])
    assert relerr(calc.results["forces"], 2.0 * F[0]) <= 1e-9
    assert calc.results["stress"].shape == (6,) and relerr(calc.results["stress"], 0.5 * S[0]) <= 1e-9
    assert calc.todict()["comp_ids"] == list(cids)


def test_sgdml_bulk_predictor():
    """GDMLPredict.predict (_gdml/predict.py:1031-1182): M geometries of one molecule in one call."""
    from mbgdml_b200.gdml_predict import GDMLPredict

    md = synthetic.make_model(["h2o", "meoh"], n_train=77, sig=25.0, seed=4, c=-3.0, std=1.7, use_E=True)
    rng = np.random.default_rng(8)
    R = synthetic.fragment_geometries(rng, ["h2o", "meoh"], 40).reshape(40, -1)
    gp = GDMLPredict(dict(md))
    E, F = gp.predict(R)
    (F_only,) = gp.predict(R[3], return_E=False)
    om = orc.Model(dict(md))
    eids = np.zeros(9, dtype=int)
    for i in range(40):
        e_o, f_o = orc.predict_gdml(md["z"], R[i].reshape(9, 3), eids, [(0,)], _single_entity(om))[:2]
        assert abs(E[i] - e_o) <= TOL * max(abs(e_o), 1.0) and relerr(F[i], f_o.ravel()) <= TOL
    assert relerr(F_only[0], F[3]) <= 1e-12


def _single_entity(om):
    om.nbody_order = 1
    om.comp_ids = np.array(["mol"])
    return om


def test_sgdml_torch_predictor():
    """GDMLTorchPredict.forward (torchtools.py:999-1063): CUDA tensors in and out, CPU tensors in and out, both
    equal to the bulk predictor and the oracle."""
    import torch

    from mbgdml_b200._gdml.torchtools import GDMLTorchPredict
    from mbgdml_b200.gdml_predict import GDMLPredict

    md = synthetic.make_model(["h2o", "h2o"], n_train=50, sig=25.0, seed=13, c=-3.5, std=2.25)
    R = synthetic.fragment_geometries(np.random.default_rng(14), ["h2o", "h2o"], 9)
    om = orc.Model(dict(md))
    E_o = np.zeros(9)
    F_o = np.zeros((9, 6, 3))
    for k in range(9):
        x, J = orc.from_r(R[k].reshape(-1))
        out = orc.predict_gdml_wkr(x, J, 6, om.sig, om.n_perms, om.R_desc_perms, om.R_d_desc_alpha_perms)
        E_o[k], F_o[k] = out[0] * md["std"] + md["c"], out[1:].reshape(6, 3) * md["std"]
    net = GDMLTorchPredict(md)
    Es, Fs = net(torch.from_numpy(R).cuda())
    assert Es.is_cuda and Fs.shape == (9, 6, 3)
    assert relerr(Es.cpu().numpy(), E_o) <= TOL and relerr(Fs.cpu().numpy(), F_o) <= TOL
    (Fc,) = net(torch.from_numpy(R), return_E=False)
    assert not Fc.is_cuda and relerr(Fc.numpy(), F_o) <= TOL
    E_b, F_b = GDMLPredict(md).predict(R.reshape(9, -1))
    assert relerr(E_b, E_o) <= TOL and relerr(F_b.reshape(9, 6, 3), F_o) <= TOL
    with pytest.raises(ValueError):
        net(torch.arange(3))  # by training index: needs set_R_d_desc (torchtools.py:861-877)


def test_sgdml_predictors_training_mode():
    """The training-side entry points of the bulk predictors (predict.py:485-575, 1108-1120; torchtools.py:702-877):
    ``predict()`` without R evaluates the cached training descriptors / Jacobians, ``set_alphas`` reconfigures the model,
    ``forward(train_idxs)`` selects training points.  Checks: (a) against the oracle worker fed with the SAME
    descriptors and Jacobians; (b) the closed form the iterative solver relies on, F_train = std * K alpha, with K from
    the kernel-matrix assembly."""
    import torch

    from mbgdml_b200._gdml.torchtools import GDMLTorchPredict
    from mbgdml_b200._gdml.train import assemble_kernel_mat
    from mbgdml_b200.gdml_predict import GDMLPredict, geometries_from_desc

    comp, T, n = ["h2o", "h2o"], 30, 6
    md = synthetic.make_model(comp, n_train=T, sig=25.0, seed=21, c=-1.5, std=1.3)
    Rt = synthetic.fragment_geometries(np.random.default_rng(21), comp, T)  # what make_model derived R_desc from
    X, G = np.zeros((T, 15)), np.zeros((T, 15, 3))
    for t in range(T):
        X[t], G[t] = orc.from_r(Rt[t].reshape(-1))
    assert relerr(X, md["R_desc"].T) <= 1e-14
    assert relerr(geometries_from_desc(X, G), Rt - Rt[:, :1]) <= 1e-13
    gp = GDMLPredict(dict(md))
    with pytest.raises(ValueError):
        gp.predict()
    gp.set_R_desc(X), gp.set_R_d_desc(G)
    E, F = gp.predict()
    om = orc.Model(dict(md))
    for t in range(T):
        out = orc.predict_gdml_wkr(X[t], G[t], n, om.sig, om.n_perms, om.R_desc_perms, om.R_d_desc_alpha_perms)
        assert abs(E[t] - (out[0] * md["std"] + md["c"])) <= TOL * abs(E[t]) and relerr(F[t], out[1:] * md["std"]) <= TOL
    # (b) new alphas: forces on the training points are std * K alpha
    alpha = np.random.default_rng(22).normal(size=T * 3 * n)
    K = assemble_kernel_mat(X, G, md["tril_perms_lin"], md["sig"])
    gp.set_alphas(alpha)
    (F2,) = gp.predict(return_E=False)
    assert relerr(F2, md["std"] * (K @ alpha).reshape(T, -1)) <= 1e-9
    net = GDMLTorchPredict(dict(md))
    net.set_R_d_desc(torch.from_numpy(G))
    net.set_alphas(torch.from_numpy(alpha))
    idx = torch.tensor([4, 0, 17], device="cuda")
    Es, Fs = net(idx)
    assert Fs.is_cuda and relerr(Fs.cpu().numpy().reshape(3, -1), F2[[4, 0, 17]]) <= 1e-12
