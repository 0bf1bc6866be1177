"""TEST INFRASTRUCTURE: drives the reference's own CUDA ops, compiled UNMODIFIED from
/root/reference by oracle/build_ref.sh into oracle/_ref/libvolrend_ref.so, through ctypes with torch
CUDA tensors.  Buffer order = inputs then outputs exactly as each reference launcher reads them
(marching.cu:452-471,542-558; integrating.cu:336-348,393-409,462-476; packbits.cu:46-51).  The Python
semantics that live in the reference's wrappers (broadcasts, scatters, counter difference) are
restated here with the line they follow.  Never imported by the product.
"""
import ctypes as C
import os
import struct

import torch

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libvolrend_ref.so")
_SYMS = {
    "march_rays": "_ZN10volrendjax10march_raysEP11CUstream_stPPvPKcm",
    "march_rays_inference": "_ZN10volrendjax20march_rays_inferenceEP11CUstream_stPPvPKcm",
    "morton3d": "_ZN10volrendjax8morton3dEP11CUstream_stPPvPKcm",
    "morton3d_invert": "_ZN10volrendjax15morton3d_invertEP11CUstream_stPPvPKcm",
    "integrate_rays": "_ZN10volrendjax14integrate_raysEP11CUstream_stPPvPKcm",
    "integrate_rays_backward": "_ZN10volrendjax23integrate_rays_backwardEP11CUstream_stPPvPKcm",
    "integrate_rays_inference": "_ZN10volrendjax24integrate_rays_inferenceEP11CUstream_stPPvPKcm",
    "pack_density_into_bits": "_ZN10volrendjax22pack_density_into_bitsEP11CUstream_stPPvPKcm",
}
_lib = None


def available():
    return os.path.exists(_PATH)


def _call(name, buffers, opaque):
    global _lib
    if _lib is None:
        _lib = C.CDLL(_PATH)
    fn = getattr(_lib, _SYMS[name])
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    arr = (C.c_void_p * len(buffers))(*[b.data_ptr() for b in buffers])
    fn(C.c_void_p(torch.cuda.current_stream().cuda_stream), arr, opaque, len(opaque))


def _e(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


def morton3d(xyzs):
    out = _e(xyzs.shape[0], torch.int32, xyzs.device)
    _call("morton3d", [xyzs.contiguous(), out], struct.pack("<I", xyzs.shape[0]))
    return out


def morton3d_invert(idcs):
    out = _e((idcs.shape[0], 3), torch.int32, idcs.device)
    _call("morton3d_invert", [idcs.contiguous(), out], struct.pack("<I", idcs.shape[0]))
    return out


def packbits(density_threshold, density_grid):
    n = density_grid.shape[0]
    thr = torch.broadcast_to(torch.as_tensor(density_threshold, dtype=torch.float32, device=density_grid.device),
                             (n,)).contiguous()  # packbits/__init__.py:26
    mask, bits = _e(n, torch.bool, thr.device), _e(n // 8, torch.uint8, thr.device)
    _call("pack_density_into_bits", [thr, density_grid.contiguous(), mask, bits], struct.pack("<I", n // 8))
    return mask, bits


def march_rays(total_samples, diagonal_n_steps, K, G, bound, stepsize_portion, rays_o, rays_d, t_starts, t_ends,
               noises, occupancy_bitfield, raw=False):
    n, S, dev = rays_o.shape[0], total_samples, rays_o.device
    noises = torch.broadcast_to(torch.as_tensor(noises, dtype=torch.float32, device=dev), (n,)).contiguous()
    nxt, exc = _e(1, torch.int32, dev), _e(1, torch.int32, dev)
    valid, rn, rs = _e(n, torch.bool, dev), _e(n, torch.int32, dev), _e(n, torch.int32, dev)
    idcs, xyzs, dirs = _e(S, torch.int32, dev), _e((S, 3), torch.float32, dev), _e((S, 3), torch.float32, dev)
    dss, zs = _e(S, torch.float32, dev), _e(S, torch.float32, dev)
    _call("march_rays", [rays_o.contiguous(), rays_d.contiguous(), t_starts.contiguous(), t_ends.contiguous(), noises,
                         occupancy_bitfield.contiguous(), nxt, exc, valid, rn, rs, idcs, xyzs, dirs, dss, zs],
          struct.pack("<5I2f", n, S, diagonal_n_steps, K, G, bound, stepsize_portion))
    if raw:
        return nxt, exc, valid, rn, rs, idcs, xyzs, dirs, dss, zs
    return nxt[0] - exc[0], valid, rn, rs, idcs, xyzs, dirs, dss, zs  # marching/__init__.py:91


def march_rays_inference(diagonal_n_steps, K, G, march_steps_cap, bound, stepsize_portion, rays_o, rays_d, t_starts,
                         t_ends, occupancy_bitfield, next_ray_index_in, terminated, indices):
    N, n, cap, dev = rays_o.shape[0], terminated.shape[0], march_steps_cap, rays_o.device
    nri, idx_out, ns = _e(1, torch.int32, dev), _e(n, torch.int32, dev), _e(n, torch.int32, dev)
    tso = _e(n, torch.float32, dev)
    xyzs, dss, zs = _e((n, cap, 3), torch.float32, dev), _e((n, cap), torch.float32, dev), _e((n, cap), torch.float32, dev)
    _call("march_rays_inference",
          [rays_o.contiguous(), rays_d.contiguous(), t_starts.contiguous(), t_ends.contiguous(),
           occupancy_bitfield.contiguous(), next_ray_index_in.contiguous(), terminated.contiguous(),
           indices.contiguous(), nri, idx_out, ns, tso, xyzs, dss, zs],
          struct.pack("<6I2f", N, n, diagonal_n_steps, K, G, cap, bound, stepsize_portion))
    t_new = t_starts.clone()
    ok = idx_out.to(torch.int64) < N
    t_new[idx_out[ok].long()] = tso[ok]  # marching/__init__.py:156
    return nri, idx_out, ns, t_new, xyzs, dss, zs, tso


def integrate_rays(near_distance, rays_sample_startidx, rays_n_samples, bgs, dss, z_vals, drgbs):
    n, dev = rays_sample_startidx.shape[0], drgbs.device
    bgs = torch.broadcast_to(torch.as_tensor(bgs, dtype=torch.float32, device=dev), (n, 3)).contiguous()  # impl.py:60
    mbs, rgbd, opac = _e(1, torch.int32, dev), _e((n, 4), torch.float32, dev), _e(n, torch.float32, dev)
    _call("integrate_rays", [rays_sample_startidx.contiguous(), rays_n_samples.contiguous(), bgs, dss.contiguous(),
                             z_vals.contiguous(), drgbs.contiguous(), mbs, rgbd, opac],
          struct.pack("<2I", n, dss.shape[0]))
    return mbs[0], rgbd, opac


def integrate_rays_backward(near_distance, rays_sample_startidx, rays_n_samples, bgs, dss, z_vals, drgbs,
                            final_rgbds, final_opacities, dL_dfinal_rgbds):
    n, S, dev = rays_sample_startidx.shape[0], dss.shape[0], drgbs.device
    bgs = torch.broadcast_to(torch.as_tensor(bgs, dtype=torch.float32, device=dev), (n, 3)).contiguous()
    dbg, dz, dd = _e((n, 3), torch.float32, dev), _e(S, torch.float32, dev), _e((S, 4), torch.float32, dev)
    _call("integrate_rays_backward",
          [rays_sample_startidx.contiguous(), rays_n_samples.contiguous(), bgs, dss.contiguous(), z_vals.contiguous(),
           drgbs.contiguous(), final_rgbds.contiguous(), final_opacities.contiguous(), dL_dfinal_rgbds.contiguous(),
           dbg, dz, dd], struct.pack("<2If", n, S, near_distance))
    return dbg, dz, dd


def integrate_rays_inference(rays_bg, rays_rgbd, rays_T, n_samples, indices, dss, z_vals, drgbs, raw=False):
    N, (n, cap), dev = rays_rgbd.shape[0], dss.shape, drgbs.device
    cnt, term = _e(1, torch.int32, dev), _e(n, torch.bool, dev)
    rgbd_o, T_o = _e((n, 4), torch.float32, dev), _e(n, torch.float32, dev)
    _call("integrate_rays_inference",
          [rays_bg.contiguous(), rays_rgbd.contiguous(), rays_T.contiguous(), n_samples.contiguous(),
           indices.contiguous(), dss.contiguous(), z_vals.contiguous(), drgbs.contiguous(), cnt, term, rgbd_o, T_o],
          struct.pack("<3I", N, n, cap))
    if raw:
        return cnt, term, rgbd_o, T_o
    rgbd, T = rays_rgbd.clone(), rays_T.clone()
    ok = indices.to(torch.int64) < N
    rgbd[indices[ok].long()] = rgbd_o[ok]  # integrating/__init__.py:108-109
    T[indices[ok].long()] = T_o[ok]
    return cnt[0], term, rgbd, T
