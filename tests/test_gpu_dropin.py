"""The drop-in claim of SURVEY 8(b): the UNMODIFIED reference driver (mbgdml.mbe.mbePredict, reference gdmlModel /
Criteria / Cell / mbeAlchemyScale objects) with this library's predictor plugged in as ``predict_model``
(mbe.py:809-811 totals, mbe.py:949-951 decomposed) reproduces the reference's own outputs.

Two plugins are driven: (1) ``integration/gdml_b200_predict.py`` -- the stand-alone ctypes binding a reference
maintainer would add, which imports nothing from the mbgdml_b200 Python package -- and (2) the reference's own
predictor beside it on the same inputs (live, on the host).  The reference package comes from oracle/_ref
(oracle/fetch_ref.py; travels to the GPU box) with the third-party stubs of oracle/refshim."""
import importlib.util
import os

import numpy as np
import pytest

from helpers import cutoffs_from, load, relerr, unpack_models
from oracle import reference

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not reference.available(), reason="reference package not installed in oracle/_ref")]
TOL = 1e-9
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    return reference.load()


@pytest.fixture(scope="module")
def plugin(ref):
    spec = importlib.util.spec_from_file_location("gdml_b200_predict", os.path.join(ROOT, "integration", "gdml_b200_predict.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _ref_models(ref, dicts, cutoffs, desc="com", bounds=None):
    out = []
    for k, (d, cut) in enumerate(zip(dicts, cutoffs)):
        bound = "upper" if bounds is None else bounds[k]
        if desc == "com":
            crit = ref.Criteria(ref.com_distance_sum, {"entity_ids": np.asarray(d["entity_ids"])}, cut, bound=bound)
        else:
            crit = ref.Criteria(ref.max_atom_pair_dist, {}, cut, bound=bound)
        out.append(ref.gdmlModel(dict(d), criteria=crit))
    return out


def test_reference_driver_with_b200_plugin_cluster(ref, plugin):
    g = load("cluster_6h2o.npz")
    mds = unpack_models(g)
    Z, R, eids, cids = g["Z"], g["R"], g["entity_ids"], g["comp_ids"]
    # plain, with criteria, with alchemical scaling: the reference's outputs are in the fixture
    E, F = ref.mbePredict(_ref_models(ref, mds, [None] * 3), plugin.predict_gdml_b200).predict(Z, R, eids, cids)
    assert relerr(E, g["plain_E"]) <= TOL and relerr(F, g["plain_F"]) <= TOL
    cuts = cutoffs_from(g["crit_cutoffs"])
    E, F = ref.mbePredict(_ref_models(ref, mds, cuts), plugin.predict_gdml_b200).predict(Z, R, eids, cids)
    assert relerr(E, g["crit_E"]) <= TOL and relerr(F, g["crit_F"]) <= TOL
    al = [ref.mbeAlchemyScale(int(e), int(o), ref.linear_switching, {"factor": float(f)}) for e, o, f in g["alch"]]
    E, F = ref.mbePredict(_ref_models(ref, mds, cuts), plugin.predict_gdml_b200, alchemy_scalers=al).predict(Z, R, eids, cids)
    assert relerr(E, g["alch_E"]) <= TOL and relerr(F, g["alch_F"]) <= TOL
    pair = _ref_models(ref, mds, [None, (3.0, 6.0), 5.5], desc="pair", bounds=["upper", "upper", "lower"])  # make_golden (e)
    E, F = ref.mbePredict(pair, plugin.predict_gdml_b200).predict(Z, R, eids, cids)
    assert relerr(E, g["pair_E"]) <= TOL and relerr(F, g["pair_F"]) <= TOL
    # and live, side by side with the reference's own predictor on a perturbed geometry
    R2 = R + np.random.default_rng(0).normal(scale=0.03, size=R.shape)
    Er, Fr = ref.mbePredict(_ref_models(ref, mds, cuts), ref.predict_gdml).predict(Z, R2, eids, cids)
    Eb, Fb = ref.mbePredict(_ref_models(ref, mds, cuts), plugin.predict_gdml_b200).predict(Z, R2, eids, cids)
    assert relerr(Eb, Er) <= TOL and relerr(Fb, Fr) <= TOL


def test_reference_driver_decomposed(ref, plugin):
    g = load("cluster_6h2o.npz")
    mds = unpack_models(g)
    Z, R, eids, cids = g["Z"], g["R"], g["entity_ids"], g["comp_ids"]
    cuts = cutoffs_from(g["crit_cutoffs"])
    drv = ref.mbePredict(_ref_models(ref, mds, cuts), plugin.predict_gdml_decomp_b200)
    E_data, F_data, comb_data, orders = drv.predict_decomp(Z, R, eids, cids)
    assert list(orders) == [1, 2, 3]
    for k in range(3):
        Ek, Fk, Ck = g[f"decomp_E{k}"], g[f"decomp_F{k}"], g[f"decomp_C{k}"]
        n = len(E_data[k])
        assert np.array_equal(comb_data[k], Ck[:n])
        assert np.array_equal(np.isnan(E_data[k]), np.isnan(Ek[:n]))  # same fragments rejected
        ok = ~np.isnan(Ek[:n])
        assert relerr(E_data[k][ok], Ek[:n][ok]) <= TOL and relerr(F_data[k][ok], Fk[:n][ok]) <= TOL


def test_reference_driver_periodic_stress(ref, plugin):
    g = load("periodic_40h2o.npz")
    mds = unpack_models(g)
    cuts = cutoffs_from(g["crit_cutoffs"])
    al = [ref.mbeAlchemyScale(int(e), int(o), ref.linear_switching, {"factor": float(f)}) for e, o, f in g["alch"]]
    drv = ref.mbePredict(_ref_models(ref, mds, cuts), plugin.predict_gdml_b200, alchemy_scalers=al,
                         periodic_cell=ref.Cell(g["cub_cell"], pbc=True), compute_stress=True)
    E, F, S = drv.predict(g["cub_Z"], g["cub_R"], g["cub_eids"], g["cub_cids"])
    assert relerr(E, g["cub_E"]) <= TOL and relerr(F, g["cub_F"]) <= TOL and relerr(S, g["cub_stress"]) <= TOL


def test_unsupported_requests_raise_like_the_header_says(ref, plugin):
    g = load("cluster_6h2o.npz")
    mds = unpack_models(g)
    Z, R, eids, cids = g["Z"], g["R"], g["entity_ids"], g["comp_ids"]
    custom = ref.Criteria(lambda Z, R, **kw: np.zeros(len(R)), {}, 1.0)
    models = [ref.gdmlModel(dict(mds[1]), criteria=custom)]
    with pytest.raises(NotImplementedError):
        ref.mbePredict(models, plugin.predict_gdml_b200).predict(Z, R, eids, cids)
