"""ctypes binding of libngp_b200.so (include/ngp_b200.h).

Every op has the XLA legacy custom-call signature ``op(stream, void** buffers, opaque, len)``; this
module is the analogue of the reference's ``xla_client.register_custom_call_target`` plumbing
(deps/volume-rendering-jax/src/volrendjax/marching/impl.py:12-13) for a torch-driven host.

There is NO fallback: if the shared library is missing or a launch fails, calls raise.
"""
import ctypes as C
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NGP_B200_LIB") or os.path.join(_HERE, "lib", "libngp_b200.so")  # override: A/B builds
CSRC = os.path.join(_HERE, "csrc")

#: every symbol include/ngp_b200.h declares
OPS = (
    "ngp_pack_density_into_bits", "ngp_packbits_scalar", "ngp_march_rays", "ngp_march_rays_inference", "ngp_march_rays_skip_empty", "ngp_march_rays_inference_inplace", "ngp_integrate_rays_inference_inplace",
    "ngp_morton3d", "ngp_morton3d_invert", "ngp_integrate_rays", "ngp_integrate_rays_backward",
    "ngp_integrate_rays_inference", "ngp_hashgrid_encode", "ngp_hashgrid_encode_backward",
    "ngp_hashgrid_a1_forward", "ngp_hashgrid_a1_backward", "ngp_adam_step", "ngp_adam_step_exchange", "ngp_ogrid_sample_positions", "ngp_ogrid_decay_max", "ngp_ogrid_threshold", "ngp_nerf_mlp_forward", "ngp_nerf_mlp_backward", "ngp_nerf_mlp_backward_mma", "ngp_nerf_mlp_backward_tc", "ngp_nerf_fused_forward", "ngp_nerf_fused_forward_umma", "ngp_umma_selftest", "ngp_make_training_rays", "ngp_huber_loss_grad", "ngp_integrate_loss_fused",
    "ngp_make_training_rays_rng", "ngp_philox_uniform", "ngp_ogrid_draw_cells", "ngp_u32_axpy", "ngp_render_frame", "ngp_nerf_mlp_backward_acc", "ngp_hashgrid_a1_backward_acc", "ngp_nerf_mlp_backward_scatter",
)
#: the reference's ten registered targets (volume-rendering-jax lib/ffi.cc:25-51, jax-tcnn lib/ffi.cc:25-30)
DROP_IN_TARGETS = ("pack_density_into_bits", "march_rays", "march_rays_inference", "morton3d", "morton3d_invert",
                   "integrate_rays", "integrate_rays_backward", "integrate_rays_inference", "hashgrid_encode",
                   "hashgrid_encode_backward")
#: their status-returning forms (five arguments: + XlaCustomCallStatus*)
STATUS_FORMS = tuple(f"ngp_{t}_status" for t in DROP_IN_TARGETS)
STATUS_SYMBOLS = ("ngp_b200_abi_version", "ngp_b200_last_status", "ngp_b200_last_error", "ngp_b200_clear_error",
                  "ngp_b200_set_march_ctas_per_sm", "ngp_b200_set_status_failure_fn") + STATUS_FORMS

_lib = None
launch_count = 0  # number of custom calls issued through this binding (bench.py reports it)


class NgpError(RuntimeError):
    pass


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into lib/libngp_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise NgpError("building libngp_b200.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NgpError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or eager fallback for this path)")
        _lib = C.CDLL(LIB_PATH)
        for name in OPS:
            fn = getattr(_lib, name)
            fn.restype = None
            fn.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
        _lib.ngp_b200_abi_version.restype = C.c_int
        _lib.ngp_b200_last_status.restype = C.c_int
        _lib.ngp_b200_last_error.restype = C.c_char_p
        _lib.ngp_b200_clear_error.restype = None
        _lib.ngp_b200_set_march_ctas_per_sm.restype = None
        _lib.ngp_b200_set_march_ctas_per_sm.argtypes = [C.c_int]
        for name in STATUS_FORMS:
            fn = getattr(_lib, name)
            fn.restype = None
            fn.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t, C.c_void_p]
        _lib.ngp_b200_set_status_failure_fn.restype = None
        _lib.ngp_b200_set_status_failure_fn.argtypes = [C.c_void_p]
    return _lib


def _ptr(t):
    if isinstance(t, torch.Tensor):
        if not t.is_cuda:
            raise NgpError("ngp_b200 ops take CUDA tensors only (no CPU path)")
        if not t.is_contiguous():
            raise NgpError("ngp_b200 ops take contiguous (row-major) tensors")
        return t.data_ptr()
    return int(t)


def call(name, buffers, opaque, stream=None):
    """Enqueue custom call `name` on `stream` (default: torch's current stream)."""
    global launch_count
    L = lib()
    arr = (C.c_void_p * len(buffers))(*[_ptr(b) for b in buffers])
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    getattr(L, name)(C.c_void_p(stream), arr, opaque, len(opaque))
    launch_count += 1
    st = L.ngp_b200_last_status()
    if st != 0:
        msg = L.ngp_b200_last_error().decode()
        L.ngp_b200_clear_error()
        raise NgpError(f"{name} failed with status {st}: {msg}")
