"""Registration of libngp_b200.so's drop-in targets with a live JAX -- the file a box with jaxlib imports once.

The reference registers its ten compiled ops as XLA GPU custom calls from capsules of its pybind11 modules
(deps/volume-rendering-jax/src/volrendjax/marching/impl.py:12-13 and the sibling impl.py files;
deps/jax-tcnn/src/jaxtcnn/hashgrid_tcnn/impl.py) and lowers to them with
``custom_call(target, ..., backend_config=opaque)`` (marching/lowering.py:19-211, ...), whose default ``api_version`` is
XLA's status-returning form: XLA calls ``f(stream, buffers, opaque, opaque_len, XlaCustomCallStatus*)``.  ``register()``
puts this library's ``ngp_<target>_status`` entry points (include/ngp_b200.h: the same ops, failures reported through the
status object) under the reference's target names, so the reference's lowering rules reach them unchanged.

jax / jaxlib are not installed in this repo's build image: nothing here is imported by the rest of the package, and the
import of jax is deferred to ``register()``.  What CAN be checked without jax is (tests/test_bindings.py): the capsule
names, that every capsule holds the address of the matching symbol, and that a failing call reaches the status callback.
"""
import ctypes

from . import _lib
from .volrendjax.volrendutils_cuda import encapsulate_function

#: target name in XLA's registry -> exported symbol (status-returning form)
TARGETS = {t: f"ngp_{t}_status" for t in _lib.DROP_IN_TARGETS}


def capsules(status_form: bool = True) -> dict:
    """``{target: PyCapsule("xla._CUSTOM_CALL_TARGET")}`` for all ten ops; ``status_form=False`` gives the
    four-argument symbols the reference's own registration code would pick up from the drop-in extension modules."""
    return {t: encapsulate_function(sym if status_form else f"ngp_{t}") for t, sym in TARGETS.items()}


def install_status_failure_fn() -> bool:
    """Point the library at jaxlib's ``XlaCustomCallStatusSetFailure``.  The library looks the symbol up with
    ``dlsym(RTLD_DEFAULT)`` by itself; jaxlib is usually loaded RTLD_LOCAL, so this resolves it inside the loaded
    ``xla_extension`` module and hands the address over.  Returns False if the symbol cannot be found (failures are
    then only visible through ``ngp_b200_last_status()``)."""
    try:
        import jaxlib.xla_extension as xe  # noqa: F401
        handle = ctypes.CDLL(xe.__file__)
        fn = ctypes.cast(handle.XlaCustomCallStatusSetFailure, ctypes.c_void_p).value
    except Exception:
        return False
    _lib.lib().ngp_b200_set_status_failure_fn(ctypes.c_void_p(fn))
    return True


def register(platform: str = "CUDA") -> list:
    """Register every drop-in target with the running JAX and return their names.  Call once, before the first jit of a
    function that uses ``volrendjax`` / ``jaxtcnn`` ops (the reference does the same at import of its impl modules)."""
    import jax  # deferred: absent from the build image

    install_status_failure_fn()
    caps = capsules(status_form=True)
    ffi = getattr(jax, "ffi", None)
    for name, capsule in caps.items():
        if ffi is not None and hasattr(ffi, "register_ffi_target"):
            # api_version=0: the untyped custom-call ABI (stream, buffers, opaque, opaque_len[, status]) -- not the typed FFI
            ffi.register_ffi_target(name, capsule, platform=platform, api_version=0)
        else:  # jax < 0.4.31: the call the reference itself makes (marching/impl.py:12-13)
            from jax.lib import xla_client
            xla_client.register_custom_call_target(name, capsule, platform="gpu")
    return sorted(caps)
