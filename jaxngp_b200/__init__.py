"""jaxngp_b200 -- B200-native (sm_100a) implementation of jaxngp's NeRF hot path.

Layout:
  csrc/        hand-written CUDA kernels + the C ABI (include/ngp_b200.h) -> lib/libngp_b200.so
  volrendjax/  host mirror of the reference's volume-rendering-jax op package
  jaxtcnn/     host mirror of the reference's jax-tcnn op package
  encoders.py  HashGridEncoder / TCNNHashGridEncoder modules (models/encoders.py:58-305)

The CUDA library is mandatory: nothing in this package computes on the CPU.
"""
from . import _lib, descriptors  # noqa: F401

__version__ = "0.1.0"
