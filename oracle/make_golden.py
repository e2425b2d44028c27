"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference with the stub packages in oracle/refshim) and cross-check the NumPy
oracle against it.  TEST INFRASTRUCTURE ONLY; runs in the build container only
(the GPU box has no /root/reference -- it uses the committed fixtures).

    python oracle/make_golden.py            # writes tests/golden/, asserts oracle == reference

Every fixture stores the inputs next to the reference's outputs, so the tests
need neither the reference nor this script at run time.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("MBGDML_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, REF)

from mbgdml.mbe import mbePredict  # noqa: E402  (the reference)
from mbgdml.models import gdmlModel  # noqa: E402
from mbgdml.predictors import predict_gdml, predict_gdml_decomp  # noqa: E402
from mbgdml.predictors.gdml_predict import _predict_gdml_wkr  # noqa: E402
from mbgdml.descriptors import Criteria, com_distance_sum, max_atom_pair_dist  # noqa: E402
from mbgdml.periodic import Cell  # noqa: E402
from mbgdml.alchemy import mbeAlchemyScale  # noqa: E402
from mbgdml.switching import linear_switching  # noqa: E402
from mbgdml.utils import gen_combs  # noqa: E402
from mbgdml._gdml.desc import _from_r, Desc  # noqa: E402

from oracle import mbgdml_oracle as orc  # noqa: E402
from mbgdml_b200 import synthetic as syn  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)

MODEL_KEYS = ["z", "sig", "R_desc", "R_d_desc_alpha", "c", "std", "perms", "tril_perms_lin", "comp_ids",
              "entity_ids", "alphas_E", "lattice"]


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = max(np.max(np.abs(b)), 1e-300) if b.size else 1.0
    return float(np.max(np.abs(a - b)) / scale) if b.size else 0.0


def slim(model):
    out = {k: np.asarray(model[k]) for k in MODEL_KEYS if k in model}
    out["r_unit"] = np.array("Angstrom")
    return out


def pack_models(prefix, models):
    out = {}
    for i, m in enumerate(models):
        for k, v in slim(m).items():
            out[f"{prefix}{i}_{k}"] = v
    out[f"{prefix}n"] = np.array(len(models))
    return out


def eids_of(model):
    return np.asarray(model["entity_ids"])


def ref_models(dicts, cutoffs, desc=com_distance_sum, bounds=None):
    out = []
    for k, (d, cut) in enumerate(zip(dicts, cutoffs)):
        kwargs = {"entity_ids": eids_of(d)} if desc is com_distance_sum else {}
        bound = "upper" if bounds is None else bounds[k]
        out.append(gdmlModel(dict(d), criteria=Criteria(desc, kwargs, cut, bound=bound)))
    return out


def orc_models(dicts, cutoffs, desc="com", bounds=None):
    out = []
    for k, (d, cut) in enumerate(zip(dicts, cutoffs)):
        bound = "upper" if bounds is None else bounds[k]
        if desc == "com":
            crit = orc.Criteria(orc.com_distance_sum, {"entity_ids": eids_of(d)}, cut, bound=bound)
        else:
            crit = orc.Criteria(orc.max_atom_pair_dist, {}, cut, bound=bound)
        out.append(orc.Model(dict(d), criteria=crit))
    return out


def check(name, got, want, tol=1e-12):
    err = relerr(got, want)
    print(f"  oracle vs reference  {name:32s} rel.err = {err:.2e}")
    assert err <= tol, (name, err)


# ----------------------------------------------------------------------------
def case_gen_combs():
    sets_list = [
        [[0, 1, 2, 3]],
        [[0, 1, 2, 3], [0, 1, 2, 3]],
        [list(range(7))] * 3,
        [[0, 1], [2, 3]],
        [[2, 3], [0, 1]],           # methanol ids before water ids -> empty (utils.py:486-488)
        [[0, 2], [1, 3]],
        [[0, 1, 2], [0, 1, 2], [3, 4]],
        [[0, 3, 5], [1, 3, 4, 6], [2, 4, 6, 7]],
        [[0], [1, 2], [1, 2, 3]],
    ]
    out = {"n": np.array(len(sets_list))}
    for k, sets in enumerate(sets_list):
        ref = np.array([tuple(int(x) for x in c) for c in gen_combs([np.array(s) for s in sets])], dtype=np.int64)
        ref = ref.reshape(-1, len(sets))
        got = orc.gen_combs_array(sets)
        assert np.array_equal(ref, got), sets
        lazy = np.array(list(orc.gen_combs(sets)), dtype=np.int64).reshape(-1, len(sets))
        assert np.array_equal(ref, lazy), sets
        out[f"s{k}_order"] = np.array(len(sets))
        for j, s in enumerate(sets):
            out[f"s{k}_set{j}"] = np.array(s, dtype=np.int64)
        out[f"s{k}_combs"] = ref
    np.savez_compressed(os.path.join(GOLD, "gen_combs.npz"), **out)
    print("gen_combs: %d set families bit-exact" % len(sets_list))


def case_desc_and_kernel():
    """_from_r, Desc.perm, model expansion and _predict_gdml_wkr on single fragments."""
    out = {}
    for tag, comps in (("1b", ["h2o"]), ("2b", ["h2o", "h2o"]), ("3b", ["h2o", "h2o", "h2o"])):
        md = syn.make_model(comps, n_train=48, sig={"1b": 10.0, "2b": 100.0, "3b": 400.0}[tag], seed=11,
                            c=-3.25, std=1.7)
        rm = gdmlModel(dict(md))
        om = orc.Model(dict(md))
        assert np.array_equal(rm.R_desc_perms, om.R_desc_perms)
        assert np.array_equal(rm.R_d_desc_alpha_perms, om.R_d_desc_alpha_perms)
        for p in md["perms"]:
            assert np.array_equal(Desc.perm(p), orc.desc_perm(p))
        rng = np.random.default_rng(5)
        r = syn.fragment_geometries(rng, comps, 4)
        xs, Js, outs = [], [], []
        for g in r:
            x, J = _from_r(g.flatten(), None)
            x = x.ravel()
            xo, Jo = orc.from_r(g.flatten())
            check(f"from_r x {tag}", xo, x, 1e-15)
            check(f"from_r J {tag}", Jo, J, 1e-15)
            o = _predict_gdml_wkr(x, J, len(md["z"]), rm.sig, rm.n_perms, rm.R_desc_perms,
                                  rm.R_d_desc_alpha_perms, rm.alphas_E_lin)
            oo = orc.predict_gdml_wkr(xo, Jo, len(md["z"]), om.sig, om.n_perms, om.R_desc_perms,
                                      om.R_d_desc_alpha_perms, om.alphas_E_lin)
            check(f"predict_gdml_wkr {tag}", oo, o, 1e-13)
            xs.append(x), Js.append(J), outs.append(o)
        out.update(pack_models(f"{tag}_m", [md]))
        out[f"{tag}_r"] = r
        out[f"{tag}_x"] = np.array(xs)
        out[f"{tag}_J"] = np.array(Js)
        out[f"{tag}_out"] = np.array(outs)
    np.savez_compressed(os.path.join(GOLD, "fragment_kernel.npz"), **out)


def synth_water_models(T, seed, use_E=False):
    return [
        syn.make_model(["h2o"], n_train=T, sig=10.0, seed=seed, c=-1.5, std=0.7),
        syn.make_model(["h2o", "h2o"], n_train=T, sig=100.0, seed=seed + 1, c=2.25, std=1.3, use_E=use_E),
        syn.make_model(["h2o", "h2o", "h2o"], n_train=T, sig=400.0, seed=seed + 2, c=-0.125, std=2.1),
    ]


def case_cluster():
    """6-H2O cluster: plain, with criteria, with alchemy, decomposed."""
    Z, R0, eids, cids = syn.make_cluster(6, seed=3)
    R = syn.jitter_frames(R0, 3, sigma=0.05, seed=4)
    mds = synth_water_models(40, seed=20)
    out = dict(Z=Z, R=R, entity_ids=eids, comp_ids=cids)
    out.update(pack_models("m", mds))

    # (a) no criteria
    ref = mbePredict(ref_models(mds, [None] * 3), predict_gdml).predict(Z, R, eids, cids)
    got = orc.mbe_predict(orc_models(mds, [None] * 3), Z, R, eids, cids)
    check("cluster plain E", got[0], ref[0]); check("cluster plain F", got[1], ref[1])
    out["plain_E"], out["plain_F"] = ref

    # (b) criteria (some rejected)
    cuts = [None, 4.2, 7.5]
    ref = mbePredict(ref_models(mds, cuts), predict_gdml).predict(Z, R, eids, cids)
    got = orc.mbe_predict(orc_models(mds, cuts), Z, R, eids, cids)
    check("cluster criteria E", got[0], ref[0]); check("cluster criteria F", got[1], ref[1])
    out["crit_cutoffs"] = np.array([np.nan, 4.2, 7.5])
    out["crit_E"], out["crit_F"] = ref

    # (c) criteria + alchemy (two scalers hit order 3, one order 2)
    al = [(2, 3, 0.5), (4, 3, 0.25), (2, 2, 0.75), (0, 1, 0.0)]
    ref_al = [mbeAlchemyScale(e, o, linear_switching, {"factor": f}) for e, o, f in al]
    orc_al = [orc.AlchemyScale(e, o, f) for e, o, f in al]
    ref = mbePredict(ref_models(mds, cuts), predict_gdml, alchemy_scalers=ref_al).predict(Z, R, eids, cids)
    got = orc.mbe_predict(orc_models(mds, cuts), Z, R, eids, cids, alchemy_scalers=orc_al)
    check("cluster alchemy E", got[0], ref[0]); check("cluster alchemy F", got[1], ref[1])
    out["alch"] = np.array(al)
    out["alch_E"], out["alch_F"] = ref

    # (d) decomposed (criteria on)
    refd = mbePredict(ref_models(mds, cuts), predict_gdml_decomp).predict_decomp(Z, R, eids, cids)
    gotd = orc.mbe_predict_decomp(orc_models(mds, cuts), Z, R, eids, cids)
    for k in range(3):
        assert np.array_equal(np.isnan(refd[0][k]), np.isnan(gotd[0][k]))
        assert np.array_equal(refd[2][k], gotd[2][k])
        check(f"cluster decomp E{k+1}", np.nan_to_num(gotd[0][k]), np.nan_to_num(refd[0][k]))
        check(f"cluster decomp F{k+1}", np.nan_to_num(gotd[1][k]), np.nan_to_num(refd[1][k]))
        out[f"decomp_E{k}"], out[f"decomp_F{k}"], out[f"decomp_C{k}"] = refd[0][k], refd[1][k], refd[2][k]
    print("  decomp: rejected dimers %d/%d, trimers %d/%d" % (
        np.isnan(refd[0][1]).sum(), len(refd[0][1]), np.isnan(refd[0][2]).sum(), len(refd[0][2])))

    # (e) max_atom_pair_dist criteria with a (lo, hi) range and a lower bound
    cuts2 = [None, (3.0, 6.0), 5.5]
    bounds2 = ["upper", "upper", "lower"]
    ref = mbePredict(ref_models(mds, cuts2, desc=max_atom_pair_dist, bounds=bounds2), predict_gdml).predict(
        Z, R, eids, cids)
    got = orc.mbe_predict(orc_models(mds, cuts2, desc="pair", bounds=bounds2), Z, R, eids, cids)
    check("cluster pairdist E", got[0], ref[0]); check("cluster pairdist F", got[1], ref[1])
    out["pair_E"], out["pair_F"] = ref

    # (f) alphas_E branch (gdml_predict.py:132-140) on the 2-body model
    mdsE = synth_water_models(40, seed=20, use_E=True)
    ref = mbePredict(ref_models(mdsE, cuts), predict_gdml).predict(Z, R, eids, cids)
    got = orc.mbe_predict(orc_models(mdsE, cuts), Z, R, eids, cids)
    check("cluster alphas_E E", got[0], ref[0]); check("cluster alphas_E F", got[1], ref[1])
    out["useE_alphas_E"] = mdsE[1]["alphas_E"]
    out["useE_E"], out["useE_F"] = ref
    np.savez_compressed(os.path.join(GOLD, "cluster_6h2o.npz"), **out)


def case_periodic():
    """Periodic boxes (cubic and triclinic): MIC + cutoff + criteria + alchemy + group virial."""
    out = {}
    mds = synth_water_models(32, seed=40)
    out.update(pack_models("m", mds))
    cuts = [None, 5.0, 8.5]
    al = [(1, 3, 0.5), (1, 2, 0.5)]
    ref_al = [mbeAlchemyScale(e, o, linear_switching, {"factor": f}) for e, o, f in al]
    orc_al = [orc.AlchemyScale(e, o, f) for e, o, f in al]
    out["alch"] = np.array(al)
    out["crit_cutoffs"] = np.array([np.nan, 5.0, 8.5])

    # cubic
    Z, R, eids, cids, cell = syn.make_periodic_box(n_mol=40, box=12.4, n_frames=2, seed=8, frame_sigma=0.08)
    pred = mbePredict(ref_models(mds, cuts), predict_gdml, alchemy_scalers=ref_al,
                      periodic_cell=Cell(cell, pbc=True), compute_stress=True)
    ref = pred.predict(Z, R, eids, cids)
    got = orc.mbe_predict(orc_models(mds, cuts), Z, R, eids, cids, periodic_cell=orc.Cell(cell),
                          alchemy_scalers=orc_al, compute_stress=True)
    for nm, g, r in zip(("E", "F", "stress"), got, ref):
        check(f"periodic cubic {nm}", g, r)
    out.update(cub_Z=Z, cub_R=R, cub_eids=eids, cub_cids=cids, cub_cell=cell,
               cub_E=ref[0], cub_F=ref[1], cub_stress=ref[2])
    # accepted counts (for information and for the GPU accept-mask test)
    for k, m in enumerate(orc_models(mds, cuts)):
        orc.predict_gdml(Z, R[0], eids, orc.gen_combs(orc.get_avail_entities(cids, m.comp_ids)), m, orc.Cell(cell))
        out[f"cub_naccept{k}"] = np.array(orc.predict_gdml.last_n_eval)
        print(f"  cubic box: model {k+1}-body accepted {orc.predict_gdml.last_n_eval}")

    # cubic with a user cutoff (5.0) below L/2 and Voigt / normal-stress post-processing
    pred = mbePredict(ref_models(mds, cuts), predict_gdml, periodic_cell=Cell(cell, cutoff=5.0, pbc=True),
                      compute_stress=True)
    pred.use_voigt = True
    pred.only_normal_stress = True
    pred.box_scaling_type = "isotropic"
    ref = pred.predict(Z, R, eids, cids)
    got = orc.mbe_predict(orc_models(mds, cuts), Z, R, eids, cids, periodic_cell=orc.Cell(cell, cutoff=5.0),
                          compute_stress=True, use_voigt=True, only_normal_stress=True,
                          box_scaling_type="isotropic")
    for nm, g, r in zip(("E", "F", "stress"), got, ref):
        check(f"periodic cutoff5 {nm}", g, r)
    out.update(cut5_E=ref[0], cut5_F=ref[1], cut5_stress=ref[2])

    # triclinic (naive branch where safe, general Minkowski branch otherwise)
    tri = np.array([[12.6, 0.0, 0.0], [2.8, 12.1, 0.0], [-1.9, 3.1, 11.8]])
    frac = np.linalg.solve(cell.T, R.reshape(-1, 3).T).T.reshape(R.shape)
    Rt = frac @ tri
    pred = mbePredict(ref_models(mds, cuts), predict_gdml, alchemy_scalers=ref_al,
                      periodic_cell=Cell(tri, pbc=True), compute_stress=True)
    ref = pred.predict(Z, Rt, eids, cids)
    got = orc.mbe_predict(orc_models(mds, cuts), Z, Rt, eids, cids, periodic_cell=orc.Cell(tri),
                          alchemy_scalers=orc_al, compute_stress=True)
    for nm, g, r in zip(("E", "F", "stress"), got, ref):
        check(f"periodic triclinic {nm}", g, r)
    out.update(tri_R=Rt, tri_cell=tri, tri_E=ref[0], tri_F=ref[1], tri_stress=ref[2])

    # triclinic with a cutoff above half the shortest cell vector (forces the general branch to matter)
    big = 0.5 * np.min(np.linalg.norm(tri, axis=1)) + 0.6
    pred = mbePredict(ref_models(mds, cuts), predict_gdml, periodic_cell=Cell(tri, cutoff=big, pbc=True))
    ref = pred.predict(Z, Rt, eids, cids)
    got = orc.mbe_predict(orc_models(mds, cuts), Z, Rt, eids, cids, periodic_cell=orc.Cell(tri, cutoff=big))
    check("periodic tri bigcut E", got[0], ref[0]); check("periodic tri bigcut F", got[1], ref[1])
    out.update(tribig_cutoff=np.array(big), tribig_E=ref[0], tribig_F=ref[1])
    np.savez_compressed(os.path.join(GOLD, "periodic_40h2o.npz"), **out)


def case_mixed():
    """Water-methanol cluster: heterogeneous comp_ids, n_atoms 3..12, ordering quirk."""
    comp = ["h2o", "h2o", "h2o", "meoh", "meoh"]
    Z, R, eids, cids = syn.make_cluster(5, seed=6, spacing=3.6, comp=comp)
    families = [["h2o"], ["meoh"], ["h2o", "h2o"], ["h2o", "meoh"], ["meoh", "meoh"],
                ["h2o", "h2o", "h2o"], ["h2o", "h2o", "meoh"], ["h2o", "meoh", "meoh"]]
    mds = [syn.make_model(f, n_train=24, sig=60.0 + 30 * k, seed=60 + k, c=0.5 * k, std=1.0 + 0.1 * k)
           for k, f in enumerate(families)]
    al = [(3, 3, 0.5), (3, 2, 0.0)]
    ref_al = [mbeAlchemyScale(e, o, linear_switching, {"factor": f}) for e, o, f in al]
    orc_al = [orc.AlchemyScale(e, o, f) for e, o, f in al]
    cuts = [None] * len(mds)
    ref = mbePredict(ref_models(mds, cuts), predict_gdml, alchemy_scalers=ref_al).predict(Z, R, eids, cids)
    got = orc.mbe_predict(orc_models(mds, cuts), Z, R, eids, cids, alchemy_scalers=orc_al)
    check("mixed E", got[0], ref[0]); check("mixed F", got[1], ref[1])
    out = dict(Z=Z, R=R, entity_ids=eids, comp_ids=cids, alch=np.array(al), E=ref[0], F=ref[1])
    out.update(pack_models("m", mds))
    # same molecules listed methanol-first: mixed tuples vanish (utils.py:486-488)
    comp2 = ["meoh", "meoh", "h2o", "h2o", "h2o"]
    Z2, R2, eids2, cids2 = syn.make_cluster(5, seed=6, spacing=3.6, comp=comp2)
    ref2 = mbePredict(ref_models(mds, cuts), predict_gdml).predict(Z2, R2, eids2, cids2)
    got2 = orc.mbe_predict(orc_models(mds, cuts), Z2, R2, eids2, cids2)
    check("mixed(meoh first) E", got2[0], ref2[0]); check("mixed(meoh first) F", got2[1], ref2[1])
    out.update(Z2=Z2, R2=R2, entity_ids2=eids2, comp_ids2=cids2, E2=ref2[0], F2=ref2[1])
    np.savez_compressed(os.path.join(GOLD, "mixed_h2o_meoh.npz"), **out)


def case_reference_fixtures():
    """The reference's own known-answer tests, with its trained T=500 water models."""
    mdir = os.path.join(REF, "tests/data/models")
    paths = [
        "140h2o.sphere.gfn2.md.500k.prod1.3h2o.dset.1h2o-model-train500.npz",
        "140h2o.sphere.gfn2.md.500k.prod1.3h2o.dset.2h2o.cm.6-model.mb-train500.npz",
        "140h2o.sphere.gfn2.md.500k.prod1.3h2o-model.mb-train500.npz",
    ]
    mds = [dict(np.load(os.path.join(mdir, p), allow_pickle=True)) for p in paths]
    for k, m in enumerate(mds):
        if "entity_ids" not in m or np.asarray(m["entity_ids"]).size != len(m["z"]):
            m["entity_ids"] = np.repeat(np.arange(k + 1), 3)
    cuts = [None, 6.0, 10.0]
    np.savez_compressed(os.path.join(GOLD, "ref_models_t500.npz"), **pack_models("m", mds))

    # tests/test_predict.py:40-128
    d16 = dict(np.load(os.path.join(REF, "tests/data/datasets/16h2o/16h2o.yoo.etal.boat.b-dset-mp2.def2tzvp.npz"),
                       allow_pickle=True))
    Z, R, eids, cids = d16["z"], d16["R"], d16["entity_ids"], d16["comp_ids"]
    ref = mbePredict(ref_models(mds, cuts), predict_gdml).predict(Z, R, eids, cids)
    got = orc.mbe_predict(orc_models(mds, cuts), Z, R, eids, cids)
    # trained models are ill-conditioned (hard part 1): compare at the reference's own tolerance
    print("  16mer: oracle vs live reference  E abs %.3e   F abs %.3e" % (
        abs(got[0][0] - ref[0][0]), np.abs(got[1] - ref[1]).max()))
    assert np.allclose(got[0], ref[0]) and np.allclose(got[1], ref[1], rtol=1e-4, atol=1e-2)
    E_pin = np.array([-766368.03399751])
    assert np.allclose(got[0], E_pin)
    np.savez_compressed(os.path.join(GOLD, "ref_16mer.npz"), Z=Z, R=R, entity_ids=eids, comp_ids=cids,
                        E_live=ref[0], F_live=ref[1], E_pinned=E_pin)

    # tests/test_predictsets.py:40-77 (decomp == total) inputs
    d6 = dict(np.load(os.path.join(REF, "tests/data/datasets/6h2o/6h2o.temelso.etal-dset.npz"), allow_pickle=True))
    Z, R, eids, cids = d6["z"], d6["R"], d6["entity_ids"], d6["comp_ids"]
    ref = mbePredict(ref_models(mds, cuts), predict_gdml).predict(Z, R, eids, cids)
    np.savez_compressed(os.path.join(GOLD, "ref_6h2o.npz"), Z=Z, R=R, entity_ids=eids, comp_ids=cids,
                        E_live=ref[0], F_live=ref[1])

    # tests/test_descriptors.py:40-49 and tests/test_predictsets.py:80-142
    d2 = dict(np.load(os.path.join(REF, "tests/data/datasets/2h2o/16h2o.yoo.etal.boat.b.2h2o-dset.mb.npz"),
                      allow_pickle=True))
    z, R2 = d2["z"], d2["R"]
    e2 = np.array([0, 0, 0, 1, 1, 1])
    v_ref = com_distance_sum(z, R2, e2)
    v_orc = orc.com_distance_sum(z, R2, e2)
    assert np.array_equal(v_ref, v_orc), "com_distance_sum must be bit-exact"
    assert v_orc[42] == 5.3809148637976385
    nan_idx = np.array([15, 16, 21, 25, 26, 31, 33, 35, 36, 40, 43, 45, 48, 52, 65, 67, 71, 72, 77, 78, 82, 88,
                        89, 92, 93, 97, 102, 107, 115, 117])
    assert np.array_equal(np.where(~(v_orc < 6.0))[0], nan_idx)
    np.savez_compressed(os.path.join(GOLD, "ref_2h2o_criteria.npz"), z=z, R=R2, entity_ids=e2,
                        desc_v=v_ref, nan_idx=nan_idx, known_answer_42=np.array(5.3809148637976385))
    print("  com_distance_sum: 120 dimers bit-exact; known answer and NaN index set reproduced")


def case_mbe_contrib():
    """mbe.mbe_contrib / decomp_to_total (reference tests/test_mbe.py:38-100, tests/test_predictsets.py:40-77)."""
    from mbgdml.mbe import mbe_contrib, decomp_to_total  # the reference

    mdir = os.path.join(REF, "tests/data/mbe")
    d4 = dict(np.load(os.path.join(mdir, "4h2o.npz"), allow_pickle=True))
    lows = [dict(np.load(os.path.join(mdir, f"4h2o-{k}body.npz"), allow_pickle=True)) for k in (1, 2, 3)]
    keys = [("energy_ele_mp2.def2tzvp_orca", "grads_mp2.def2tzvp_orca")] + \
           [("energy_ele_nbody_mp2.def2tzvp_orca", "grads_nbody_mp2.def2tzvp_orca")] * 2
    E_true, G_true = d4["energy_ele_mp2.def2tzvp_orca"], d4["grads_mp2.def2tzvp_orca"]
    entity_ids = d4["entity_ids"]
    out = {"entity_ids": entity_ids, "E_true": E_true, "G_true": G_true}
    E_ref, G_ref = np.zeros(E_true.shape), np.zeros(G_true.shape)
    E_orc, G_orc = np.zeros(E_true.shape), np.zeros(G_true.shape)
    for k, (low, (ek, gk)) in enumerate(zip(lows, keys)):
        args = (low[ek], low[gk], low["entity_ids"], low["r_prov_ids"].item(), low["r_prov_specs"])
        E_ref, G_ref = mbe_contrib(E_ref, G_ref, entity_ids, None, None, *args, operation="add", use_ray=False)
        E_orc, G_orc = orc.mbe_contrib(E_orc, G_orc, entity_ids, None, None, *args, operation="add")
        assert np.array_equal(E_ref, E_orc) and np.array_equal(G_ref, G_orc), "mbe_contrib must be bit-exact"
        out.update({f"low{k}_E": low[ek], f"low{k}_G": low[gk], f"low{k}_entity_ids": low["entity_ids"],
                    f"low{k}_r_prov_specs": low["r_prov_specs"], f"low{k}_r_prov_id": np.array(list(args[3].keys())),
                    f"add{k}_E": E_ref.copy(), f"add{k}_G": G_ref.copy()})
    assert np.allclose(E_ref, np.array([-305.30075391, -305.29900387, -305.29489548]))  # tests/test_mbe.py:99-100
    # "remove" with explicit provenance of the reference structures and a lower set that misses some fragments
    r_prov_ids = {0: "x"}
    specs = np.array([[0, i, 0, 1, 2, 3] for i in range(3)])
    low = lows[1]
    keep = np.ones(len(low["r_prov_specs"]), dtype=bool)
    keep[[1, 7, 8]] = False
    args = (low[keys[1][0]][keep], low[keys[1][1]][keep], low["entity_ids"], {0: "x"}, low["r_prov_specs"][keep])
    E_r, G_r = mbe_contrib(E_true.copy(), G_true.copy(), entity_ids, r_prov_ids, specs, *args, operation="remove",
                           use_ray=False)
    E_o, G_o = orc.mbe_contrib(E_true.copy(), G_true.copy(), entity_ids, r_prov_ids, specs, *args, operation="remove")
    assert np.array_equal(E_r, E_o) and np.array_equal(G_r, G_o)
    out.update({"rm_keep": keep, "rm_specs": specs, "rm_E": E_r, "rm_G": G_r})
    # decomp_to_total on a synthetic decomposition with NaN rows (rejected fragments), 2 structures
    rng = np.random.default_rng(4)
    eids6 = np.repeat(np.arange(5), 3)
    combs = np.array([(s, i, j) for s in range(2) for i in range(5) for j in range(i + 1, 5)])
    E_n, F_n = rng.normal(size=len(combs)), rng.normal(size=(len(combs), 6, 3))
    E_n[[3, 11]] = np.nan
    F_n[[3, 11]] = np.nan
    E_t, F_t = decomp_to_total(E_n.copy(), F_n.copy(), eids6, combs)
    E_to, F_to = orc.decomp_to_total(E_n.copy(), F_n.copy(), eids6, combs)
    assert np.array_equal(E_t, E_to) and np.array_equal(F_t, F_to)
    out.update({"d2t_entity_ids": eids6, "d2t_combs": combs, "d2t_E_nbody": E_n, "d2t_F_nbody": F_n,
                "d2t_E": E_t, "d2t_F": F_t})
    np.savez_compressed(os.path.join(GOLD, "mbe_contrib_4h2o.npz"), **out)
    print("  mbe_contrib add x3 (known answer reproduced), remove with missing fragments, decomp_to_total: bit-exact")


def case_train_kernel():
    """analysis.models.gdml_mat52 (analysis/models.py:31-139) and the force-field kernel matrix
    GDMLTrain._assemble_kernel_mat (_gdml/train.py:92-292, 1105-1337), outputs of the live reference."""
    from mbgdml._gdml.train import GDMLTrain  # the reference
    from mbgdml.analysis.models import gdml_mat52

    trainer = GDMLTrain(max_processes=1)
    out = {}
    cases = [("h2o", ["h2o"], 14, 7.0, 0), ("h2o2", ["h2o", "h2o"], 9, 20.0, 1), ("h2o3", ["h2o", "h2o", "h2o"], 4, 35.0, 2),
             ("meoh", ["meoh"], 6, 12.0, 3)]
    for name, comp, T, sig, seed in cases:
        rng = np.random.default_rng(100 + seed)
        Rt = syn.fragment_geometries(rng, comp, T)
        n = Rt.shape[1]
        desc = Desc(n, max_processes=1)
        R_desc, R_d_desc = desc.from_R(Rt.reshape(T, -1))
        perms = syn.fragment_perms(comp)
        tpl = orc.tril_perms_lin_from_perms(perms)
        out.update({f"{name}_R": Rt, f"{name}_R_desc": R_desc, f"{name}_R_d_desc": R_d_desc, f"{name}_perms": perms,
                    f"{name}_tril_perms_lin": tpl, f"{name}_sig": np.array(sig)})
        for use_E in (False, True):
            K_ref = np.array(trainer._assemble_kernel_mat(R_desc, R_d_desc, tpl, sig, desc, use_E_cstr=use_E))
            K_orc = orc.assemble_kernel_mat(R_desc, R_d_desc, tpl, sig, use_E_cstr=use_E)
            check(f"kernel_mat {name} use_E={use_E}", K_orc, K_ref, tol=1e-11)
            out[f"{name}_K_E{int(use_E)}"] = K_ref
    # covariances vs. the training set on the prediction models (monomer, dimer, trimer; 1 and 5 structures)
    for k, comp in enumerate((["h2o"], ["h2o", "h2o"], ["h2o", "h2o", "h2o"])):
        md = syn.make_model(comp, n_train=40, sig=[10.0, 100.0, 400.0][k], seed=20 + k)
        rng = np.random.default_rng(300 + k)
        Rq = syn.fragment_geometries(rng, comp, 5)
        ref_m, orc_m = gdmlModel(dict(md)), orc.Model(dict(md))
        for tag, R in (("1", Rq[0]), ("5", Rq)):
            want = gdml_mat52(ref_m, md["z"], R)
            check(f"mat52 {len(comp)}-body n_R={tag}", orc.gdml_mat52(orc_m, md["z"], R), want, tol=1e-13)
            out[f"mat52_{k}_R{tag}"] = R
            out[f"mat52_{k}_out{tag}"] = want
        out.update({f"mat52_{k}_{kk}": v for kk, v in slim(md).items()})
    # the reference's own kernel-matrix fixture (tests/data/train/1h2o/K.npy with R_desc.npy, tril_perms_lin.npy,
    # train_idxs.npy; sigma = 42, tests/test_train.py:51-88): the blocks of the first 20 training points
    tdir = os.path.join(REF, "tests/data/train/1h2o")
    K_fix = np.load(os.path.join(tdir, "K.npy"))
    idx = np.load(os.path.join(tdir, "train_idxs.npy"))
    tpl = np.load(os.path.join(tdir, "tril_perms_lin.npy"))
    dset = dict(np.load(os.path.join(REF, "tests/data/datasets/1h2o/140h2o.sphere.gfn2.md.500k.prod1.3h2o.dset.1h2o-dset.npz"),
                        allow_pickle=True))
    Rt = dset["R"][idx]
    R_desc, R_d_desc = Desc(3, max_processes=1).from_R(Rt.reshape(len(idx), -1))
    assert np.array_equal(R_desc.T, np.load(os.path.join(tdir, "R_desc.npy")))
    check("kernel_mat reference fixture 1h2o/K.npy", orc.assemble_kernel_mat(R_desc, R_d_desc, tpl, 42), K_fix, tol=1e-13)
    nk = 20
    out.update({"fix1h2o_R": Rt[:nk], "fix1h2o_R_desc": R_desc[:nk], "fix1h2o_R_d_desc": R_d_desc[:nk],
                "fix1h2o_tril_perms_lin": tpl, "fix1h2o_sig": np.array(42.0), "fix1h2o_K": K_fix[:nk * 9, :nk * 9].copy()})
    np.savez_compressed(os.path.join(GOLD, "train_kernel.npz"), **out)


def case_lattice():
    """Models that carry their own `lattice` (sGDML periodic descriptors): _from_r with _pbc_diff (desc.py:42-76) and
    predict_gdml through it, cubic and triclinic lattices, fragments that straddle the lattice boundary."""
    out = {}
    lattices = {"cub": np.eye(3) * 7.5, "tri": np.array([[7.9, 1.1, -0.6], [0.0, 7.2, 0.9], [0.0, 0.0, 8.4]])}  # columns
    for tag, lat in lattices.items():
        mds = []
        for k, comps in enumerate((["h2o"], ["h2o", "h2o"], ["h2o", "h2o", "h2o"])):
            md = syn.make_model(comps, n_train=40, sig=(10.0, 80.0, 300.0)[k], seed=31 + k, c=0.5 * k - 1.0, std=1.0 + 0.25 * k)
            md["lattice"] = lat.copy()
            mds.append(md)
        rng = np.random.default_rng(9)
        Z, R0, eids, cids = syn.make_cluster(5, seed=4, spacing=2.9)
        # move whole molecules by lattice vectors (columns of lat) so that pair differences need the clamp
        shifts = rng.integers(-1, 2, size=(5, 3)) @ lat.T
        R = np.array([R0 + np.repeat(shifts, 3, axis=0), R0 + rng.normal(scale=0.04, size=R0.shape)])
        rms = [gdmlModel(dict(md)) for md in mds]
        oms = [orc.Model(dict(md)) for md in mds]
        ref = mbePredict(rms, predict_gdml).predict(Z, R, eids, cids)
        got = orc.mbe_predict(oms, Z, R, eids, cids)
        check(f"lattice {tag} E", got[0], ref[0]); check(f"lattice {tag} F", got[1], ref[1])
        # clamped geometry reproduces the un-shifted one's energy (the lattice makes them equivalent)
        plain = mbePredict(rms, predict_gdml).predict(Z, R0[None], eids, cids)
        assert relerr(ref[0][:1], plain[0]) <= 1e-10
        frag = R[0][:9]
        x, J = _from_r(frag.flatten(), rms[2].lat_and_inv)
        xo, Jo = orc.from_r(frag.flatten(), oms[2].lat_and_inv)
        check(f"lattice {tag} from_r x", xo, x.ravel(), 1e-14); check(f"lattice {tag} from_r J", Jo, J, 1e-13)
        out.update(pack_models(f"{tag}_m", mds))
        out.update({f"{tag}_Z": Z, f"{tag}_R": R, f"{tag}_eids": eids, f"{tag}_cids": cids, f"{tag}_E": ref[0],
                    f"{tag}_F": ref[1], f"{tag}_frag": frag, f"{tag}_x": x.ravel(), f"{tag}_J": J, f"{tag}_lat": lat})
    np.savez_compressed(os.path.join(GOLD, "lattice_model.npz"), **out)


def case_train_kernel_cols():
    """Column subsets of the kernel matrix (`col_idxs`, _gdml/train.py:1105-1115, 1187-1245): the live reference's
    outputs for the index forms its callers use -- a leading slice of training points (exploit_sym + cols_m_limit), a
    slice of training points that does not start at 0, and sorted index arrays (with and without energy-constraint
    columns).  A slice that is not a whole number of training points makes the reference compare a Python list with
    an int (train.py:1217) and raise TypeError, so that form has no golden."""
    from mbgdml._gdml.train import GDMLTrain  # the reference

    trainer = GDMLTrain(max_processes=1)
    comp, T, sig = ["h2o", "h2o"], 7, 20.0
    rng = np.random.default_rng(411)
    Rt = syn.fragment_geometries(rng, comp, T)
    n = Rt.shape[1]
    desc = Desc(n, max_processes=1)
    R_desc, R_d_desc = desc.from_R(Rt.reshape(T, -1))
    tpl = orc.tril_perms_lin_from_perms(syn.fragment_perms(comp))
    out = {"R_desc": R_desc, "R_d_desc": R_d_desc, "tril_perms_lin": tpl, "sig": np.array(sig)}
    dim_i = 3 * n
    forms = {
        "lead": (np.s_[: 3 * dim_i], False, 0),
        "mid": (np.s_[2 * dim_i: 5 * dim_i], False, 0),
        "idx": (np.array([1, 5, 17, 18, 19, 20, 53, 100, 125]), False, 2),
        "idxE": (np.array([0, 2, 40, 41, 90, T * dim_i, T * dim_i + 3, T * dim_i + 6]), True, 0),
        "leadE": (np.s_[: 2 * dim_i], True, 1),
    }
    for name, (cols, use_E, extra) in forms.items():
        K_full = np.array(trainer._assemble_kernel_mat(R_desc, R_d_desc, tpl, sig, desc, use_E_cstr=use_E))
        try:
            K_sub = np.array(trainer._assemble_kernel_mat(R_desc, R_d_desc, tpl, sig, desc, use_E_cstr=use_E, col_idxs=cols,
                                                          alloc_extra_rows=extra))
        except (ValueError, TypeError, IndexError) as exc:
            print(f"  reference fails on col_idxs form {name!r}: {type(exc).__name__}: {exc}")
            continue
        want = K_full[:, cols]
        assert K_sub.shape == (K_full.shape[0] + extra, want.shape[1]), (name, K_sub.shape, want.shape)
        check(f"kernel_mat cols {name}: reference subset == reference full[:, cols]", K_sub[: K_full.shape[0]], want, tol=1e-13)
        if isinstance(cols, slice):
            out[f"{name}_slice"] = np.array([cols.start or 0, cols.stop])
        else:
            out[f"{name}_cols"] = cols
        out[f"{name}_use_E"] = np.array(int(use_E))
        out[f"{name}_extra"] = np.array(extra)
        out[f"{name}_K"] = K_sub[: K_full.shape[0]]
    np.savez_compressed(os.path.join(GOLD, "train_kernel_cols.npz"), **out)


def case_slab():
    """Partial periodicity (Cell(pbc=...) with a non-periodic direction, periodic.py:40-59 -> find_mic(d, cell_v, pbc)):
    the 40-H2O box of case_periodic as a slab and as a wire, cubic and triclinic, through the live reference."""
    out = {}
    mds = synth_water_models(32, seed=40)
    out.update(pack_models("m", mds))
    cuts = [None, 5.0, 8.5]
    out["crit_cutoffs"] = np.array([np.nan, 5.0, 8.5])
    Z, R, eids, cids, cell = syn.make_periodic_box(n_mol=40, box=12.4, n_frames=2, seed=8, frame_sigma=0.08)
    tri = np.array([[12.6, 0.0, 0.0], [2.8, 12.1, 0.0], [-1.9, 3.1, 11.8]])
    Rt = np.linalg.solve(cell.T, R.reshape(-1, 3).T).T.reshape(R.shape) @ tri
    out.update(Z=Z, eids=eids, cids=cids, cub_R=R, cub_cell=cell, tri_R=Rt, tri_cell=tri)
    pbcs = [(True, True, False), (False, True, True), (True, False, False)]
    out["pbcs"] = np.array(pbcs)
    for tag, cv, Rk in (("cub", cell, R), ("tri", tri, Rt)):
        for k, pbc in enumerate(pbcs):
            ref = mbePredict(ref_models(mds, cuts), predict_gdml, periodic_cell=Cell(cv, pbc=pbc)).predict(Z, Rk, eids, cids)
            got = orc.mbe_predict(orc_models(mds, cuts), Z, Rk, eids, cids, periodic_cell=orc.Cell(cv, pbc=pbc))
            check(f"slab {tag} {pbc} E", got[0], ref[0]); check(f"slab {tag} {pbc} F", got[1], ref[1])
            full = mbePredict(ref_models(mds, cuts), predict_gdml, periodic_cell=Cell(cv, pbc=True)).predict(Z, Rk, eids, cids)
            assert relerr(full[0], ref[0]) > 1e-6  # the flags matter
            out[f"{tag}_E{k}"], out[f"{tag}_F{k}"] = ref[0], ref[1]
    np.savez_compressed(os.path.join(GOLD, "slab_40h2o.npz"), **out)


if __name__ == "__main__":
    np.seterr(all="raise", under="ignore")
    if "--only-slab" in sys.argv:
        print("slab / wire"); case_slab()
        sys.exit(0)
    if "--only-lattice" in sys.argv:
        print("lattice models"); case_lattice()
        sys.exit(0)
    if "--only-train-cols" in sys.argv:
        print("train kernel, column subsets"); case_train_kernel_cols()
        sys.exit(0)
    print("gen_combs"); case_gen_combs()
    print("fragment kernel"); case_desc_and_kernel()
    print("cluster"); case_cluster()
    print("periodic"); case_periodic()
    print("mixed"); case_mixed()
    print("reference fixtures"); case_reference_fixtures()
    print("mbe_contrib"); case_mbe_contrib()
    print("train kernel"); case_train_kernel()
    print("lattice models"); case_lattice()
    print("train kernel, column subsets"); case_train_kernel_cols()
    print("slab / wire"); case_slab()
    total = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print("golden fixtures written: %.1f KB" % (total / 1024))
