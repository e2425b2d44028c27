"""Stub of `ase` (absent from this image).  TEST INFRASTRUCTURE ONLY.

The only ASE arithmetic on the hot path is `ase.geometry.find_mic`
(reference periodic.py:78).  It is RESTATED in geometry/__init__.py from the
published algorithm (ASE >= 3.21, ase/geometry/geometry.py and
ase/geometry/minkowski_reduction.py); ASE is unpinned in the reference's
setup.py:12.  Parity of the periodic path is therefore "unpinned"
(see DESIGN.md): there is no real ASE here to check this restatement against.
"""


class Atoms:  # placeholder so `import ase; ase.Atoms` resolves (gap/schnet predictors)
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("ase stub")
