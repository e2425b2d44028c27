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
