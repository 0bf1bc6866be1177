"""Drop-in for the reference's compiled extension module ``volrendjax.volrendutils_cuda``
(deps/volume-rendering-jax/lib/ffi.cc:17-207, a pybind11 module): the same functions, returning the same things --
``get_*_registrations()`` dicts of ``PyCapsule(fn, "xla._CUSTOM_CALL_TARGET")`` under the reference's target names, and
``make_*_descriptor(...)`` returning the descriptor's raw bytes -- backed by libngp_b200.so instead of the reference's
kernels.  The reference's ``impl.py`` / ``lowering.py`` files import it as ``from .. import volrendutils_cuda``
(marching/impl.py:8-13, marching/lowering.py:4,40,140, ...) and need nothing else from native code, so with this file
in place of the extension they register and lower to this library's kernels unchanged.

jax is absent from this image, so the registration itself cannot be exercised here; the CPU tests check the module
surface against the reference's ffi.cc and that every capsule holds the address of the matching C symbol.
"""
import ctypes

from .. import _lib, descriptors

_CAPSULE_NAME = b"xla._CUSTOM_CALL_TARGET"  # ffi.cc:19 -- a C string that must outlive the capsules

_PyCapsule_New = ctypes.pythonapi.PyCapsule_New
_PyCapsule_New.restype = ctypes.py_object
_PyCapsule_New.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]


def encapsulate_function(symbol: str):
    """ffi.cc:17-20: the address of an ``extern "C"`` entry point of libngp_b200.so as an XLA custom-call capsule."""
    address = ctypes.cast(getattr(_lib.lib(), symbol), ctypes.c_void_p).value
    if not address:
        raise _lib.NgpError(f"libngp_b200.so does not export {symbol}")
    return _PyCapsule_New(address, _CAPSULE_NAME, None)


def get_packbits_registrations():  # ffi.cc:25-29
    return {"pack_density_into_bits": encapsulate_function("ngp_pack_density_into_bits")}


def get_marching_registrations():  # ffi.cc:31-36
    return {"march_rays": encapsulate_function("ngp_march_rays"),
            "march_rays_inference": encapsulate_function("ngp_march_rays_inference")}


def get_morton3d_registrations():  # ffi.cc:38-43
    return {"morton3d": encapsulate_function("ngp_morton3d"),
            "morton3d_invert": encapsulate_function("ngp_morton3d_invert")}


def get_integrating_registrations():  # ffi.cc:45-51
    return {"integrate_rays": encapsulate_function("ngp_integrate_rays"),
            "integrate_rays_backward": encapsulate_function("ngp_integrate_rays_backward"),
            "integrate_rays_inference": encapsulate_function("ngp_integrate_rays_inference")}


# descriptor factories (ffi.cc:55-207): positional arguments in the reference's order, bytes out, RuntimeError for the
# two values the reference rejects
make_packbits_descriptor = descriptors.make_packbits_descriptor
make_marching_descriptor = descriptors.make_marching_descriptor
make_marching_inference_descriptor = descriptors.make_marching_inference_descriptor
make_morton3d_descriptor = descriptors.make_morton3d_descriptor
make_integrating_descriptor = descriptors.make_integrating_descriptor
make_integrating_backward_descriptor = descriptors.make_integrating_backward_descriptor
make_integrating_inference_descriptor = descriptors.make_integrating_inference_descriptor
