"""Import the UNMODIFIED reference package (keithgroup/mbGDML) for checking and for the CPU arm.

TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's CPU legs.  Resolution order:
``oracle/_ref`` (installed by oracle/fetch_ref.py; travels to the GPU box) then ``/root/reference`` (build
container).  The reference's absent third-party imports resolve to ``oracle/refshim`` (ray: serial stubs;
cclib: unused names; qcelemental: isotope masses pinned by the reference's tests/test_descriptors.py:40-49;
ase: ``find_mic`` restated in oracle/ase_mic.py -- third-party, "parity unpinned").
"""
from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.path.join(HERE, "_ref"), os.environ.get("MBGDML_REFERENCE", "/root/reference")]


def reference_root():
    for root in _CANDIDATES:
        if root and os.path.exists(os.path.join(root, "mbgdml", "mbe.py")):
            return root
    return None


def available():
    return reference_root() is not None


_ns = None


def load():
    """Namespace with the reference's hot-path surface; raises RuntimeError when no reference is present."""
    global _ns
    if _ns is not None:
        return _ns
    root = reference_root()
    if root is None:
        raise RuntimeError("reference package not found: run `python oracle/fetch_ref.py` in the build container "
                           "(installs /root/reference into the git-ignored oracle/_ref/)")
    for p in (os.path.join(HERE, "refshim"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    ns = types.SimpleNamespace(root=root)
    ns.mbe = importlib.import_module("mbgdml.mbe")
    ns.mbePredict = ns.mbe.mbePredict
    ns.gdmlModel = importlib.import_module("mbgdml.models").gdmlModel
    pred = importlib.import_module("mbgdml.predictors")
    ns.predict_gdml, ns.predict_gdml_decomp = pred.predict_gdml, pred.predict_gdml_decomp
    desc = importlib.import_module("mbgdml.descriptors")
    ns.Criteria, ns.com_distance_sum, ns.max_atom_pair_dist = desc.Criteria, desc.com_distance_sum, desc.max_atom_pair_dist
    ns.Cell = importlib.import_module("mbgdml.periodic").Cell
    ns.mbeAlchemyScale = importlib.import_module("mbgdml.alchemy").mbeAlchemyScale
    ns.linear_switching = importlib.import_module("mbgdml.switching").linear_switching
    ns.gen_combs = importlib.import_module("mbgdml.utils").gen_combs
    ns.stress = importlib.import_module("mbgdml.stress")
    _ns = ns
    return ns
