"""mbgdml_b200: B200-native many-body GDML prediction path.

A from-scratch implementation of keithgroup/mbGDML's ``mbePredict`` ->
``predict_gdml`` hot path: on-device fragment enumeration, minimum-image culling,
inverse-distance descriptors, the FP64 Matern-5/2 contraction and the force scatter
run as hand-written sm_100a CUDA kernels behind a C ABI (include/mbgdml_b200.h).
The Python surface mirrors the reference (same names, arguments, error behaviour).
See DESIGN.md.
"""
from .alchemy import mbeAlchemyScale
from .descriptors import Criteria, com_distance_sum, max_atom_pair_dist
from .mbe import decomp_to_total, mbe_contrib, mbePredict
from .models import gdmlModel
from .periodic import Cell
from .predictors import predict_gdml, predict_gdml_decomp
from .switching import linear_switching
from .utils import gen_combs

__all__ = [
    "mbePredict", "gdmlModel", "predict_gdml", "predict_gdml_decomp", "Criteria", "com_distance_sum",
    "max_atom_pair_dist", "Cell", "mbeAlchemyScale", "linear_switching", "gen_combs", "mbe_contrib",
    "decomp_to_total",
]
