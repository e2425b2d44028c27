"""Stub of `cclib` (absent from this image); only names the reference imports
at module load (utils.py:28, data/dataset.py:25).  TEST INFRASTRUCTURE ONLY."""
from . import parser, io  # noqa: F401
