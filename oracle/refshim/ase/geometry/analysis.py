"""`ase.geometry.analysis` stand-in (import-time placeholder only; never called on the paths we pin).
TEST INFRASTRUCTURE ONLY."""


class Analysis:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError("ase is not installed; this is an import shim")
