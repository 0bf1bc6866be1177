"""Drop-in for the reference's compiled extension module ``jaxtcnn.tcnnutils`` (deps/jax-tcnn/lib/ffi.cc:17-55):
``get_hashgrid_registrations()`` and ``make_hashgrid_descriptor(...)`` backed by libngp_b200.so -- tiny-cuda-nn is no
longer a build dependency.  See ``volrendjax/volrendutils_cuda.py``."""
from .. import descriptors
from ..volrendjax.volrendutils_cuda import encapsulate_function


def get_hashgrid_registrations():  # jax-tcnn/lib/ffi.cc:25-30
    return {"hashgrid_encode": encapsulate_function("ngp_hashgrid_encode"),
            "hashgrid_encode_backward": encapsulate_function("ngp_hashgrid_encode_backward")}


make_hashgrid_descriptor = descriptors.make_hashgrid_descriptor  # jax-tcnn/lib/ffi.cc:34-47, tcnnutils.h:11-29
