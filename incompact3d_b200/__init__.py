"""incompact3d_b200 -- B200-native hot path of Xcompact3d behind the reference's own
operator interface.  The compute path is the CUDA library libx3d_b200.so (C ABI in
include/x3d_b200.h); this package is the Python host mirror of that interface."""
from ._lib import DerivCoeffs, FilterCoeffs, LIB_PATH, load  # noqa: F401
from .api import AxisSchemes, X3D, X3DError, decomp_compute, nccl_unique_id, stretching, transpose_plan  # noqa: F401

__all__ = ["X3D", "X3DError", "AxisSchemes", "DerivCoeffs", "FilterCoeffs", "stretching", "load", "LIB_PATH"]
