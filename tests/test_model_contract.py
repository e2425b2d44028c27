"""The model-object contract of SURVEY 8(b) (gdml_model.py:35-191): attributes a predictor reads from ``gdmlModel``.
Host-side only (no GPU): expansions are pure indexing and must equal the reference's / the oracle's bit for bit."""
import numpy as np
import pytest

import mbgdml_b200 as mb
from helpers import load, unpack_models
from oracle import mbgdml_oracle as orc
from oracle import reference


def _models():
    g = load("fragment_kernel.npz")
    return [unpack_models(g, prefix=f"{t}_m")[0] for t in ("1b", "2b", "3b")]


def test_expanded_training_sets_equal_the_oracle():
    for md in _models():
        m, om = mb.gdmlModel(dict(md)), orc.Model(dict(md))
        assert np.array_equal(m.R_desc_perms, om.R_desc_perms)
        assert np.array_equal(m.R_d_desc_alpha_perms, om.R_d_desc_alpha_perms)
        assert m.alphas_E_lin is None and m.lat_and_inv is None and callable(m.desc_func)
        assert m.R_desc_perms is m.R_desc_perms  # cached
        assert (m.n_train, m.n_perms, m.n_atoms, m.nbody_order) == (om.n_train, om.n_perms, om.n_atoms, om.nbody_order)
    g = load("cluster_6h2o.npz")
    md = dict(unpack_models(g)[1])
    md["alphas_E"] = g["useE_alphas_E"]
    m, om = mb.gdmlModel(md), orc.Model(dict(md))
    assert np.array_equal(m.alphas_E_lin, om.alphas_E_lin)


@pytest.mark.skipif(not reference.available(), reason="reference package not installed in oracle/_ref")
def test_attributes_equal_the_reference_objects():
    ref = reference.load()
    for md in _models():
        md = dict(md)
        md["sig"] = np.asarray(md["sig"])  # as loaded from an .npz (md5 digests the array's bytes)
        m, rm = mb.gdmlModel(dict(md)), ref.gdmlModel(dict(md))
        for name in ("R_desc_perms", "R_d_desc_alpha_perms"):
            assert np.array_equal(getattr(m, name), getattr(rm, name))
        for name in ("n_atoms", "n_train", "n_perms", "nbody_order", "type", "r_unit", "use_torch"):
            assert getattr(m, name) == getattr(rm, name), name
        assert m.stdev == rm.stdev and m.integ_c == rm.integ_c and np.array_equal(m.tril_perms_lin, rm.tril_perms_lin)
        assert m.md5 == rm.md5
    lat = dict(unpack_models(load("lattice_model.npz"), prefix="tri_m")[0])
    m, rm = mb.gdmlModel(dict(lat)), ref.gdmlModel(dict(lat))
    assert np.array_equal(m.lat_and_inv[0], rm.lat_and_inv[0]) and np.array_equal(m.lat_and_inv[1], rm.lat_and_inv[1])


def test_training_mode_helpers_on_the_host():
    """Host logic of the predictors' training mode (no GPU): geometries rebuilt from descriptor + compressed Jacobian
    (what ``GDMLPredict.predict()`` without R sends to the device) equal the geometries the descriptors came from, up to
    the translation; ``GDMLTorchAssemble`` maps the entries of ``J`` to columns like the reference's ``_forward``
    (torchtools.py:119-153)."""
    import torch

    from mbgdml_b200 import synthetic
    from mbgdml_b200._gdml.torchtools import GDMLTorchAssemble
    from mbgdml_b200.gdml_predict import geometries_from_desc

    for comp in (["h2o"], ["h2o", "meoh"], ["h2o", "h2o", "h2o"]):
        R = synthetic.fragment_geometries(np.random.default_rng(7), comp, 11)
        n = R.shape[1]
        X, G = np.zeros((11, n * (n - 1) // 2)), np.zeros((11, n * (n - 1) // 2, 3))
        for t in range(11):
            X[t], G[t] = orc.from_r(R[t].reshape(-1))
        got = geometries_from_desc(X, G)
        assert np.abs(got - (R - R[:, :1])).max() <= 1e-13
        x2, g2 = orc.from_r(got[3].reshape(-1))  # same descriptor and Jacobian again
        assert np.abs(x2 - X[3]).max() <= 1e-13 and np.abs(g2 - G[3]).max() <= 1e-12
    with pytest.raises(ValueError):
        geometries_from_desc(X, G[:, :-1])
    T, n, dim_i = 5, 3, 9
    asm = GDMLTorchAssemble([0, 3, 7, (4, 2, [1, 8]), (6, T * dim_i + 4, [0])], np.arange(6), 10.0, True,
                            torch.zeros(T, 3), torch.zeros(T, 3, 3), out=None)
    assert asm.n_atoms == n and asm.dim_i == dim_i and asm.n_perms == 2
    dst, src = asm._columns(asm.J[1])
    assert list(dst) == list(range(27, 36)) and list(src) == list(range(27, 36))
    assert [list(c) for c in asm._columns(asm.J[2])] == [[T * dim_i + 2], [T * dim_i + 2]]  # energy column j % n_train
    assert [list(c) for c in asm._columns(asm.J[3])] == [[4, 5], [19, 26]]
    assert [list(c) for c in asm._columns(asm.J[4])] == [[6], [T * dim_i + 4]]
