"""CPU: the C-ABI library loads, exports every symbol include/mbgdml_b200.h declares, and the
host-side mirror fails loudly (no CPU fallback) and validates like the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

import mbgdml_b200 as mb
from mbgdml_b200 import _engine, _native, build, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "mbgdml_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mbg_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    lib = ctypes.CDLL(path)
    declared = header_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mbgdml_b200.h but not exported"
    assert sorted(_native.EXPORTED_SYMBOLS) == declared
    assert _native.load_library().mbg_abi_version() == 3


def test_struct_layouts_match_header_sizes():
    # sizes implied by the header on LP64
    assert ctypes.sizeof(_native.ModelDesc) == 16 + 24 + 32 + 8 + 16 + 8 + 8
    assert ctypes.sizeof(_native.SystemDesc) == 8 + 32 + 8 + 16 + 8
    assert ctypes.sizeof(_native.Job) == 8 + 24 + 8 + 8 + 16
    assert ctypes.sizeof(_native.JobStats) == 24


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_native.NativeUnavailable):
        _native.get_context()
    md = synthetic.make_model(["h2o"], n_train=8)
    Z, R, eids, cids = synthetic.make_cluster(2)
    pred = mb.mbePredict([mb.gdmlModel(md)], mb.predict_gdml)
    with pytest.raises(_native.NativeUnavailable):
        pred.predict(Z, R, eids, cids)


def test_system_csr_matches_np_where():
    Z = np.array([8, 1, 1, 6, 8, 1, 1, 1, 1, 8, 1, 1])
    eids = np.array([0, 0, 0, 2, 2, 1, 2, 2, 2, 1, 1, 2])  # non-contiguous entities
    s = _engine.System(Z, eids)
    for e in range(3):
        lo, hi = s.entity_offsets[e], s.entity_offsets[e + 1]
        assert np.array_equal(s.entity_atoms[lo:hi], np.where(eids == e)[0])
    assert s.masses[0] == 15.99491461957 and s.masses[1] == 1.00782503223 and s.masses[3] == 12.0


def test_criteria_native_spec_and_unsupported_callables():
    c = mb.Criteria(mb.com_distance_sum, {"entity_ids": np.array([0, 0, 0, 1, 1, 1])}, 6.0)
    kind, bound, lo, hi, eids = c.native_spec(6)
    assert (kind, bound, hi) == (_native.CRIT_COM_DISTANCE_SUM, _native.BOUND_UPPER, 6.0) and eids.dtype == np.int32
    c = mb.Criteria(mb.max_atom_pair_dist, {}, (7.0, 2.0))
    assert c.cutoff == [2.0, 7.0]
    assert c.native_spec(6)[:4] == (_native.CRIT_MAX_ATOM_PAIR_DIST, _native.BOUND_RANGE, 2.0, 7.0)
    c = mb.Criteria(mb.max_atom_pair_dist, {}, 3.0, bound="LOWER")
    assert c.native_spec(6)[:3] == (_native.CRIT_MAX_ATOM_PAIR_DIST, _native.BOUND_LOWER, 3.0)
    assert mb.Criteria(lambda Z, R: R, {}, None).native_spec(3)[0] == _native.CRIT_NONE
    with pytest.raises(NotImplementedError):
        mb.Criteria(lambda Z, R: R, {}, 1.0).native_spec(3)
    acc, val = mb.Criteria(mb.com_distance_sum, {}, None).accept(np.array([8, 1, 1]), np.zeros((3, 3)))
    assert acc.all() and val is None


def test_alchemy_and_reference_error_behaviour():
    s = mb.mbeAlchemyScale(3, 2, mb.linear_switching, {"factor": 0.25})
    assert s.native_factor() == 0.25 and s.scale(4.0) == 1.0
    with pytest.raises(AssertionError):  # switching.py:47
        mb.linear_switching(1.0, 1.5)
    with pytest.raises(NotImplementedError):
        mb.mbeAlchemyScale(3, 2, lambda d, k: d * k, {"k": 0.5}).native_factor()
    with pytest.raises(TypeError):  # gdml_model.py:67
        mb.gdmlModel(3.0)
    with pytest.raises(KeyError):  # gdml_model.py:76
        mb.gdmlModel({"r_unit": np.array("Angstrom")})
    with pytest.raises(NotImplementedError):
        mb.mbePredict([], mb.predict_gdml, use_ray=True)
    cell = mb.Cell(np.eye(3) * 10.0, cutoff=3.0)
    assert cell.cutoff == 3.0 and cell.volume == 1000.0
    cell.cell_v = np.eye(3) * 8.0  # the setter silently resets the cutoff (periodic.py:121-126)
    assert cell.cutoff == 4.0


def test_model_container_attributes():
    md = synthetic.make_model(["h2o", "h2o"], n_train=12, sig=55.0, c=1.5, std=2.0)
    m = mb.gdmlModel(md, criteria=mb.Criteria(mb.com_distance_sum, {"entity_ids": md["entity_ids"]}, 6.0))
    assert m.type == "gdml" and m.nbody_order == 2 and m.n_atoms == 6 and m.n_train == 12 and m.n_perms == 8
    assert m.sig == 55.0 and m.integ_c == 1.5 and m.stdev == 2.0 and list(m.comp_ids) == ["h2o", "h2o"]
    assert mb.gdmlModel(md, comp_ids=["a", "b"]).comp_ids == ["a", "b"]


def test_cached_system_reuse_and_invalidation():
    """_engine.cached_system: same (Z, entity_ids, cell) -> the same System object; any change -> a new one."""
    from mbgdml_b200 import _engine
    from mbgdml_b200.periodic import Cell

    Z = np.array([8, 1, 1, 8, 1, 1])
    eids = np.array([0, 0, 0, 1, 1, 1])
    s1 = _engine.cached_system(Z, eids, None, False)
    assert _engine.cached_system(Z.copy(), eids.copy(), None, False) is s1
    assert _engine.cached_system(Z, np.array([0, 0, 0, 1, 1, 2]), None, False) is not s1
    s2 = _engine.cached_system(Z, eids, None, True)  # masses become mandatory
    assert s2 is not s1
    cell = Cell(np.eye(3) * 10.0, pbc=True)
    s3 = _engine.cached_system(Z, eids, cell, True)
    assert s3 is not s2 and _engine.cached_system(Z, eids, Cell(np.eye(3) * 10.0, pbc=True), True) is s3
    assert _engine.cached_system(Z, eids, Cell(np.eye(3) * 11.0, pbc=True), True) is not s3
    assert list(s3.entity_offsets) == [0, 3, 6]


def test_constants_of_the_python_mirror_match_the_header():
    """#define values of include/mbgdml_b200.h against the constants _native.py passes through ctypes."""
    text = open(os.path.join(ROOT, "include", "mbgdml_b200.h")).read()
    defs = {k: int(v, 0) for k, v in re.findall(r"#define\s+(MBG_[A-Z0-9_]+)\s+\(?(-?(?:0x)?[0-9a-fA-F]+)\)?", text)}
    enums = dict(re.findall(r"\b(MBG_[A-Z0-9_]+)\s*=\s*(\d+)", text))
    for name in ("CONTRACT_AUTO", "CONTRACT_FMA", "CONTRACT_DMMA", "CONTRACT_DMMA_WAVE", "CULL_AUTO", "CULL_FULL",
                 "CULL_PAIRLIST"):
        assert defs["MBG_" + name] == getattr(_native, name), name
    for name in ("CRIT_NONE", "CRIT_COM_DISTANCE_SUM", "CRIT_MAX_ATOM_PAIR_DIST", "BOUND_UPPER", "BOUND_LOWER", "BOUND_RANGE"):
        assert int(enums["MBG_" + name]) == getattr(_native, name), name
    assert defs["MBG_ABI_VERSION"] == 3
