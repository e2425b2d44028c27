"""`ase.geometry` stand-in for importing the unmodified reference here.

TEST INFRASTRUCTURE ONLY.  Re-exports the restatement in oracle/ase_mic.py
(see its header: third-party algorithm, parity unpinned)."""
import importlib.util as _ilu
import os as _os

_p = _os.path.join(_os.path.dirname(__file__), "..", "..", "..", "ase_mic.py")
_spec = _ilu.spec_from_file_location("_oracle_ase_mic", _os.path.abspath(_p))
_m = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(_m)

find_mic = _m.find_mic
naive_find_mic = _m.naive_find_mic
general_find_mic = _m.general_find_mic
minkowski_reduce = _m.minkowski_reduce
