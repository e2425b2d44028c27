"""ASE calculator around ``mbePredict.predict`` (reference mbgdml/interfaces/ase.py:31-128).

One ``predict`` per MD / optimiser step: every step is a single native call (enumeration, culling,
contraction, scatter and group virial on the GPU).  ASE itself is optional at import time: with ASE installed
the class derives from ``ase.calculators.calculator.Calculator`` exactly like the reference; without it a
minimal stand-in base keeps ``calculate`` usable with any object that offers the five ``Atoms`` methods used
here (``wrap``, ``copy``, ``get_cell``, ``get_atomic_numbers``, ``get_positions``).
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - ASE is not part of the build image
    from ase.calculators.calculator import Calculator
except Exception:  # noqa: BLE001
    class Calculator:  # type: ignore[no-redef]
        """Stand-in with the attributes mbeCalculator relies on (results, atoms, parameters)."""

        implemented_properties: list = []
        default_parameters: dict = {}

        def __init__(self, restart=None, label=None, atoms=None, **kwargs):  # pylint: disable=unused-argument
            self.results = {}
            self.atoms = atoms
            self.parameters = {}

        def get_default_parameters(self):
            return dict(self.default_parameters)

        def get_potential_energy(self, atoms=None):
            self.calculate(atoms)
            return self.results["energy"]

        def get_forces(self, atoms=None):
            self.calculate(atoms)
            return self.results["forces"]

        def get_stress(self, atoms=None):
            self.calculate(atoms)
            return self.results["stress"]


class mbeCalculator(Calculator):  # pylint: disable=invalid-name
    """ASE calculator for a many-body expansion predictor."""

    implemented_properties = ["energy", "forces", "stress"]

    def __init__(self, mbe_pred, parameters=None, e_conv=1.0, f_conv=1.0, atoms=None, **kwargs):
        self.name = "mbGDML"
        Calculator.__init__(self, restart=None, label=None, atoms=atoms, **kwargs)
        self.mbe_pred = mbe_pred
        self.mbe_pred.use_voigt = True
        self.e_conv = e_conv
        self.f_conv = f_conv
        if parameters is None:
            parameters = {}
        self.parameters = parameters

    def calculate(self, atoms=None, *args, **kwargs):  # pylint: disable=unused-argument, keyword-arg-before-vararg
        """Predicts all available properties using the MBE predictor (interfaces/ase.py:65-108)."""
        is_periodic = bool(self.mbe_pred.periodic_cell)
        if atoms is not None:
            if is_periodic:
                atoms.wrap()
            self.atoms = atoms.copy()
        entity_ids = self.parameters["entity_ids"]
        comp_ids = self.parameters["comp_ids"]
        if is_periodic:
            # update the cell vectors; like the reference this also resets the cutoff to half the shortest
            # vector (periodic.py:121-126)
            periodic_cell = self.mbe_pred.periodic_cell
            periodic_cell.cell_v = atoms.get_cell()[:]
            self.mbe_pred.periodic_cell = periodic_cell
            if not self.mbe_pred.compute_stress:
                self.mbe_pred.compute_stress = True
        E, F, *stress = self.mbe_pred.predict(atoms.get_atomic_numbers(), atoms.get_positions(), entity_ids, comp_ids)
        E = E[0]
        F = F.reshape(-1, 3)
        E *= self.e_conv
        F *= self.f_conv
        self.results = {"energy": E, "forces": F}
        if is_periodic:
            self.results["stress"] = stress[0][0] * self.e_conv

    def todict(self, skip_default=True):
        """Serialisable copy of ``parameters`` for ASE's trajectory writers (interfaces/ase.py:110-128): entries that
        repeat a class default are dropped, nested objects are flattened through their own ``todict`` and string
        arrays become lists (ASE cannot reload NumPy string arrays)."""
        defaults = self.get_default_parameters() if hasattr(self, "get_default_parameters") else {}
        out = {}
        for key, value in self.parameters.items():
            if skip_default and key in defaults:
                continue
            if hasattr(value, "todict"):
                value = value.todict()
            elif isinstance(value, np.ndarray) and value.dtype.kind in "US":
                value = value.tolist()
            out[key] = value
        return out
