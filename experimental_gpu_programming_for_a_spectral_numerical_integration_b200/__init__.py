"""B200-native batched Chebyshev-collocation integration of Cosserat-rod kinematics and statics.

Product path = libsri_cuda.so (hand-written sm_100a CUDA behind the C ABI of include/sri.h); this package is the
Python host-side mirror used by tests, bench.py and the Newton shape driver.  Importing it never touches oracle/.
"""
from .api import (  # noqa: F401
    ComputeChebyshevPoints,
    MultiDeviceIntegrator,
    GetCoefficients_c,
    Phi,
    SpectralRodIntegrator,
    ad,
    getDn,
    integratePosition,
    integrateQuaternions,
    kernel_launch_count,
    scale_for_length,
    shard_range,
    skew,
)
from ._lib import SriError  # noqa: F401

__all__ = [
    "ComputeChebyshevPoints", "GetCoefficients_c", "Phi", "SpectralRodIntegrator", "getDn",
    "integratePosition", "integrateQuaternions", "kernel_launch_count", "SriError", "skew", "ad", "MultiDeviceIntegrator", "shard_range", "scale_for_length",
]
