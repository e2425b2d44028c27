"""Error of the CUDA path on the reference's trained (ill-conditioned) fixtures vs. the live-reference outputs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import mbgdml_b200 as mb
from helpers import b200_models, load, unpack_models
from mbgdml_b200 import _native
mds = unpack_models(load("ref_models_t500.npz"))
models = b200_models(mds, [None, 6.0, 10.0])
for kind in ("dmma", "fma"):
    _native.get_context().set_contraction(kind)
    for name in ("ref_16mer.npz", "ref_6h2o.npz"):
        g = load(name)
        E, F = mb.mbePredict(models, mb.predict_gdml).predict(g["Z"], g["R"], g["entity_ids"], g["comp_ids"])
        print(kind, name, "dE", np.abs(E - g["E_live"]).max(), "dF max", np.abs(F - g["F_live"]).max(), "Fmax", np.abs(g["F_live"]).max())
