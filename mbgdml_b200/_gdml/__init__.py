"""Training-side kernels that reuse the prediction engine (reference mbgdml/_gdml/)."""
from .train import assemble_kernel_mat

__all__ = ["assemble_kernel_mat"]
