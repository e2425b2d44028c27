"""Front-ends that drive the prediction path (reference mbgdml/interfaces/)."""
from .ase import mbeCalculator

__all__ = ["mbeCalculator"]
