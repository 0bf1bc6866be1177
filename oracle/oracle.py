"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy front-end of oracle/libngp_oracle.so.

Every function takes and returns numpy arrays and mirrors one reference op; the C side cites the
reference lines it follows.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libngp_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    """torchrun exports OMP_NUM_THREADS=1 to its workers: callers that want all host cores say so explicitly."""
    lib().orc_set_num_threads(int(n))
    return num_threads()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _u8(a):
    return np.ascontiguousarray(a).view(np.uint8) if np.asarray(a).dtype == np.bool_ else np.ascontiguousarray(a, dtype=np.uint8)


def morton3d(xyzs):
    xyzs = _u32(xyzs)
    out = np.empty(xyzs.shape[0], np.uint32)
    lib().orc_morton3d(C.c_uint32(xyzs.shape[0]), _p(xyzs), _p(out))
    return out


def morton3d_invert(idcs):
    idcs = _u32(idcs)
    out = np.empty((idcs.shape[0], 3), np.uint32)
    lib().orc_morton3d_invert(C.c_uint32(idcs.shape[0]), _p(idcs), _p(out))
    return out


def packbits(density_threshold, density_grid):
    density_grid = _f32(density_grid)
    thr = _f32(np.broadcast_to(np.asarray(density_threshold, np.float32), density_grid.shape))
    n = density_grid.shape[0]
    assert n % 8 == 0
    mask = np.empty(n, np.uint8)
    bits = np.empty(n // 8, np.uint8)
    lib().orc_packbits(C.c_uint32(n // 8), _p(thr), _p(density_grid), _p(mask), _p(bits))
    return mask.view(np.bool_), bits


def march_rays(total_samples, diagonal_n_steps, K, G, bound, stepsize_portion,
               rays_o, rays_d, t_starts, t_ends, noises, occupancy_bitfield, raw=False):
    rays_o, rays_d, t_starts, t_ends = map(_f32, (rays_o, rays_d, t_starts, t_ends))
    n = rays_o.shape[0]
    noises = _f32(np.broadcast_to(np.asarray(noises, np.float32), (n,)))
    bits = _u8(occupancy_bitfield)
    S = total_samples
    nxt, exc = np.zeros(1, np.uint32), np.zeros(1, np.uint32)
    valid = np.zeros(n, np.uint8)
    rn, rs = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    idcs = np.zeros(S, np.uint32)
    xyzs, dirs = np.zeros((S, 3), np.float32), np.zeros((S, 3), np.float32)
    dss, zs = np.zeros(S, np.float32), np.zeros(S, np.float32)
    lib().orc_march_rays(C.c_uint32(n), C.c_uint32(S), C.c_uint32(diagonal_n_steps), C.c_uint32(K),
                         C.c_uint32(G), C.c_float(bound), C.c_float(stepsize_portion),
                         _p(rays_o), _p(rays_d), _p(t_starts), _p(t_ends), _p(noises), _p(bits),
                         _p(nxt), _p(exc), _p(valid), _p(rn), _p(rs), _p(idcs), _p(xyzs), _p(dirs),
                         _p(dss), _p(zs))
    if raw:
        return nxt, exc, valid.view(np.bool_), rn, rs, idcs, xyzs, dirs, dss, zs
    return (int(nxt[0]) - int(exc[0]), valid.view(np.bool_), rn, rs, idcs, xyzs, dirs, dss, zs)


def march_rays_inference(diagonal_n_steps, K, G, march_steps_cap, bound, stepsize_portion,
                         rays_o, rays_d, t_starts, t_ends, occupancy_bitfield,
                         next_ray_index_in, terminated, indices):
    rays_o, rays_d, t_starts, t_ends = map(_f32, (rays_o, rays_d, t_starts, t_ends))
    N, n, cap = rays_o.shape[0], np.asarray(terminated).shape[0], march_steps_cap
    bits, term, idx_in = _u8(occupancy_bitfield), _u8(terminated), _u32(indices)
    nri_in = _u32(next_ray_index_in).reshape(1)
    nri, idx_out = np.zeros(1, np.uint32), np.zeros(n, np.uint32)
    ns, tso = np.zeros(n, np.uint32), np.zeros(n, np.float32)
    xyzs, dss, zs = np.zeros((n, cap, 3), np.float32), np.zeros((n, cap), np.float32), np.zeros((n, cap), np.float32)
    lib().orc_march_rays_inference(C.c_uint32(N), C.c_uint32(n), C.c_uint32(diagonal_n_steps),
                                   C.c_uint32(K), C.c_uint32(G), C.c_uint32(cap), C.c_float(bound),
                                   C.c_float(stepsize_portion), _p(rays_o), _p(rays_d), _p(t_starts),
                                   _p(t_ends), _p(bits), _p(nri_in), _p(term), _p(idx_in), _p(nri),
                                   _p(idx_out), _p(ns), _p(tso), _p(xyzs), _p(dss), _p(zs))
    # marching/__init__.py:156 -- t_starts.at[indices].set(t_starts_out); out-of-range dropped
    t_new = t_starts.copy()
    ok = idx_out < N
    t_new[idx_out[ok]] = tso[ok]
    return nri, idx_out, ns, t_new, xyzs, dss, zs, tso


def integrate_rays(near_distance, rays_sample_startidx, rays_n_samples, bgs, dss, z_vals, drgbs):
    start, ns = _u32(rays_sample_startidx), _u32(rays_n_samples)
    n = start.shape[0]
    bgs = _f32(np.broadcast_to(np.asarray(bgs, np.float32), (n, 3)))
    dss, z_vals, drgbs = map(_f32, (dss, z_vals, drgbs))
    mbs = np.zeros(1, np.uint32)
    rgbd, opac = np.zeros((n, 4), np.float32), np.zeros(n, np.float32)
    lib().orc_integrate_rays(C.c_uint32(n), _p(start), _p(ns), _p(bgs), _p(dss), _p(z_vals),
                             _p(drgbs), _p(mbs), _p(rgbd), _p(opac))
    return int(mbs[0]), rgbd, opac


def integrate_rays_backward(near_distance, rays_sample_startidx, rays_n_samples, bgs, dss, z_vals,
                            drgbs, final_rgbds, final_opacities, dL_dfinal_rgbds):
    start, ns = _u32(rays_sample_startidx), _u32(rays_n_samples)
    n, S = start.shape[0], np.asarray(dss).shape[0]
    bgs = _f32(np.broadcast_to(np.asarray(bgs, np.float32), (n, 3)))
    dss, z_vals, drgbs, final_rgbds, final_opacities, dL_dfinal_rgbds = map(
        _f32, (dss, z_vals, drgbs, final_rgbds, final_opacities, dL_dfinal_rgbds))
    dbg, dz, dd = np.zeros((n, 3), np.float32), np.zeros(S, np.float32), np.zeros((S, 4), np.float32)
    lib().orc_integrate_rays_backward(C.c_uint32(n), C.c_uint32(S), C.c_float(near_distance),
                                      _p(start), _p(ns), _p(bgs), _p(dss), _p(z_vals), _p(drgbs),
                                      _p(final_rgbds), _p(final_opacities), _p(dL_dfinal_rgbds),
                                      _p(dbg), _p(dz), _p(dd))
    return dbg, dz, dd


def integrate_rays_inference(rays_bg, rays_rgbd, rays_T, n_samples, indices, dss, z_vals, drgbs, raw=False):
    rays_bg, rays_rgbd, rays_T, dss, z_vals, drgbs = map(_f32, (rays_bg, rays_rgbd, rays_T, dss, z_vals, drgbs))
    ns, idx = _u32(n_samples), _u32(indices)
    N, n, cap = rays_rgbd.shape[0], ns.shape[0], dss.shape[1]
    cnt, term = np.zeros(1, np.uint32), np.zeros(n, np.uint8)
    rgbd_o, T_o = np.zeros((n, 4), np.float32), np.zeros(n, np.float32)
    lib().orc_integrate_rays_inference(C.c_uint32(N), C.c_uint32(n), C.c_uint32(cap), _p(rays_bg),
                                       _p(rays_rgbd), _p(rays_T), _p(ns), _p(idx), _p(dss),
                                       _p(z_vals), _p(drgbs), _p(cnt), _p(term), _p(rgbd_o), _p(T_o))
    if raw:
        return cnt, term.view(np.bool_), rgbd_o, T_o
    # integrating/__init__.py:108-109
    rgbd, T = rays_rgbd.copy(), rays_T.copy()
    ok = idx < N
    rgbd[idx[ok]] = rgbd_o[ok]
    T[idx[ok]] = T_o[ok]
    return int(cnt[0]), term.view(np.bool_), rgbd, T


def hashgrid_encode(levels, pos, bound, table, wrap="jaxngp"):
    """levels: dict from oracle.hashgrid_np.level_table(); pos [n, dim]; table [rows, F]."""
    pos, table = _f32(pos), _f32(table)
    n, dim = pos.shape
    L, F = levels["L"], table.shape[1]
    enc = np.zeros((n, L * F), np.float32)
    hashed = np.ascontiguousarray(levels["hashed"], np.uint8)
    lib().orc_hashgrid_encode(C.c_uint32(n), C.c_uint32(dim), C.c_uint32(L), C.c_uint32(F),
                              _p(_f32(levels["scales"])), _p(_u32(levels["res"])),
                              _p(_u32(levels["offsets"])), _p(hashed),
                              C.c_uint32(levels["T"] if wrap == "jaxngp" else 0), C.c_float(bound),
                              _p(pos), _p(table), _p(enc))
    return enc


def hashgrid_backward(levels, pos, bound, d_enc, F, wrap="jaxngp"):
    pos, d_enc = _f32(pos), _f32(d_enc)
    n, dim = pos.shape
    L = levels["L"]
    rows = int(levels["offsets"][-1])
    d_table = np.zeros((rows, F), np.float64)
    hashed = np.ascontiguousarray(levels["hashed"], np.uint8)
    lib().orc_hashgrid_backward(C.c_uint32(n), C.c_uint32(dim), C.c_uint32(L), C.c_uint32(F),
                                _p(_f32(levels["scales"])), _p(_u32(levels["res"])),
                                _p(_u32(levels["offsets"])), _p(hashed),
                                C.c_uint32(levels["T"] if wrap == "jaxngp" else 0), C.c_float(bound),
                                _p(pos), _p(d_enc), _p(d_table))
    return d_table
