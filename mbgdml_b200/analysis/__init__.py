"""Analysis helpers on the device kernels (reference mbgdml/analysis/)."""
from .models import gdml_mat52, gdml_mat52_wrk

__all__ = ["gdml_mat52", "gdml_mat52_wrk"]
