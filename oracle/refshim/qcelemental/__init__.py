"""Stub of `qcelemental` (absent from this image).  TEST INFRASTRUCTURE ONLY.

Only `periodictable.to_mass/to_symbol/to_atomic_number` are used by the
reference (descriptors.py:125, utils.py:267,284).  The masses are the
most-abundant-isotope masses qcelemental returns; H and O are pinned
bit-exactly by the reference's tests/test_descriptors.py:40-49.
"""
from . import periodictable  # noqa: F401
