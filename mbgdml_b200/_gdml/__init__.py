"""Training-side kernels that reuse the prediction engine (reference mbgdml/_gdml/)."""
from .train import GDMLTrain, assemble_kernel_mat

__all__ = ["assemble_kernel_mat", "GDMLTrain"]  # torchtools.GDMLTorchPredict is imported on demand (needs torch)
