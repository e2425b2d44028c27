"""Models that carry their own lattice (_gdml/desc.py:42-76 _pbc_diff, gdml_model.py:97-101): the device
descriptor / Jacobian clamp pair differences to the model's lattice.  Golden = outputs of the live reference
(oracle/make_golden.py:case_lattice)."""
import numpy as np
import pytest
import torch

import mbgdml_b200 as mb
from helpers import load, relerr, unpack_models
from mbgdml_b200 import _native
from mbgdml_b200._gdml.torchtools import GDMLTorchPredict
from mbgdml_b200.gdml_predict import GDMLPredict
from oracle import mbgdml_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("tag", ["cub", "tri"])
def test_mbe_prediction_with_lattice_models(tag):
    g = load("lattice_model.npz")
    mds = unpack_models(g, prefix=f"{tag}_m")
    assert all("lattice" in d for d in mds)
    models = [mb.gdmlModel(dict(d)) for d in mds]
    pred = mb.mbePredict(models, mb.predict_gdml)
    E, F = pred.predict(g[f"{tag}_Z"], g[f"{tag}_R"], g[f"{tag}_eids"], g[f"{tag}_cids"])
    assert relerr(E, g[f"{tag}_E"]) <= TOL and relerr(F, g[f"{tag}_F"]) <= TOL
    # without the lattice the shifted molecules are far apart: the clamp is what the golden answer depends on
    plain = [mb.gdmlModel({k: v for k, v in d.items() if k != "lattice"}) for d in mds]
    E0, _ = mb.mbePredict(plain, mb.predict_gdml).predict(g[f"{tag}_Z"], g[f"{tag}_R"][:1], g[f"{tag}_eids"], g[f"{tag}_cids"])
    assert abs(E0[0] - g[f"{tag}_E"][0]) > 1e-6 * abs(g[f"{tag}_E"][0])
    # the contract attribute a foreign predictor calls: model.desc_func(r.flatten(), model.lat_and_inv)
    x, J = models[2].desc_func(g[f"{tag}_frag"].flatten(), models[2].lat_and_inv)
    assert relerr(x.ravel(), g[f"{tag}_x"]) <= 1e-13 and relerr(J, g[f"{tag}_J"]) <= 1e-13
    x0, J0 = models[2].desc_func(g[f"{tag}_frag"].flatten(), None)
    xo, Jo = orc.from_r(g[f"{tag}_frag"].flatten())
    assert relerr(x0.ravel(), xo) <= 1e-13 and relerr(J0, Jo) <= 1e-13
    # decomposed rows under the lattice
    combs = np.array([[0, 1, 2], [0, 2, 4], [1, 3, 4]])
    Ed, Fd, _ = mb.predict_gdml_decomp(g[f"{tag}_Z"], g[f"{tag}_R"][0], g[f"{tag}_eids"], combs, models[2])
    Eo, Fo, _ = orc.predict_gdml_decomp(g[f"{tag}_Z"], g[f"{tag}_R"][0], g[f"{tag}_eids"], combs, orc.Model(dict(mds[2])))
    assert relerr(Ed, Eo) <= TOL and relerr(Fd, Fo) <= TOL


def test_bulk_predictors_with_lattice():
    """GDMLPredict / GDMLTorchPredict(lat_and_inv=...) (_gdml/predict.py, torchtools.py:401-1063) on single molecules."""
    g = load("lattice_model.npz")
    md = unpack_models(g, prefix="tri_m")[1]  # water dimer model
    lat = np.asarray(md["lattice"])
    om = orc.Model(dict(md))
    rng = np.random.default_rng(2)
    frag = g["tri_R"][0][:6]
    geoms = np.array([frag + rng.normal(scale=0.03, size=frag.shape) for _ in range(5)])
    geoms[1, 3:] += lat[:, 0]  # second molecule moved by a lattice vector
    want_E, want_F = [], []
    for r in geoms:
        x, J = orc.from_r(r.flatten(), om.lat_and_inv)
        o = orc.predict_gdml_wkr(x, J, 6, om.sig, om.n_perms, om.R_desc_perms, om.R_d_desc_alpha_perms, om.alphas_E_lin)
        o = o * om.stdev
        want_E.append(o[0] + om.integ_c), want_F.append(o[1:].reshape(6, 3))
    E, F = GDMLPredict(dict(md)).predict(geoms.reshape(5, -1))
    assert relerr(E, want_E) <= TOL and relerr(F.reshape(5, 6, 3), want_F) <= TOL
    no_lat = {k: v for k, v in md.items() if k != "lattice"}
    tp = GDMLTorchPredict(no_lat, lat_and_inv=(torch.from_numpy(lat), torch.from_numpy(np.linalg.inv(lat))))
    Es, Fs = tp(torch.from_numpy(geoms).cuda())
    assert relerr(Es.cpu().numpy(), want_E) <= TOL and relerr(Fs.cpu().numpy(), want_F) <= TOL


def test_lattice_models_need_the_default_kernel():
    g = load("lattice_model.npz")
    mds = unpack_models(g, prefix="cub_m")
    ctx = _native.get_context()
    pred = mb.mbePredict([mb.gdmlModel(dict(mds[1]))], mb.predict_gdml)
    for kind in ("fma", "dmma_wave"):
        ctx.set_contraction(kind)
        try:
            with pytest.raises(NotImplementedError):
                pred.predict(g["cub_Z"], g["cub_R"][:1], g["cub_eids"], g["cub_cids"])
        finally:
            ctx.set_contraction("auto")
