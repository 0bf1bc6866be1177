"""march_rays / march_rays_inference -- mirrors volrendjax/marching/__init__.py:8-157."""
from typing import Tuple

import torch

from .. import _lib, descriptors
from ._check import as_u32, assert_shape, non_negative, positive, require_f32


def march_rays(
    # static
    total_samples: int,
    diagonal_n_steps: int,
    K: int,
    G: int,
    bound: float,
    stepsize_portion: float,

    # inputs
    rays_o: torch.Tensor,
    rays_d: torch.Tensor,
    t_starts: torch.Tensor,
    t_ends: torch.Tensor,
    noises,
    occupancy_bitfield: torch.Tensor,
    raw: bool = False,
) -> Tuple[torch.Tensor, ...]:
    """Same contract as the reference's ``march_rays`` (marching/__init__.py:8-93).

    Returns ``(measured_batch_size_before_compaction, ray_is_valid, rays_n_samples,
    rays_sample_startidx, idcs, xyzs, dirs, dss, z_vals)``; the first item is a 0-d int32 tensor
    on the device (``next_sample_write_location - number_of_exceeded_samples``, :91), never synced
    to the host here.  ``raw=True`` returns the 10 primitive outputs in abstract.py:61-72 order.
    """
    n_rays = rays_o.shape[0]
    # marching/abstract.py:24-47
    assert_shape(rays_o, (n_rays, 3), "rays_o")
    assert_shape(rays_d, (n_rays, 3), "rays_d")
    assert_shape(t_starts, (n_rays,), "t_starts")
    assert_shape(t_ends, (n_rays,), "t_ends")
    assert_shape(occupancy_bitfield, (K * G * G * G // 8,), "occupancy_bitfield")
    if occupancy_bitfield.dtype != torch.uint8:
        raise AssertionError(f"occupancy_bitfield must be uint8, got {occupancy_bitfield.dtype}")
    for v, nm in ((total_samples, "total_samples"), (diagonal_n_steps, "diagonal_n_steps"), (K, "K"), (G, "G"),
                  (bound, "bound")):
        positive(v, nm)
    non_negative(stepsize_portion, "stepsize_portion")
    require_f32(rays_o, "march_rays", "input coordinates")
    dev = rays_o.device
    # marching/__init__.py:71
    noises = torch.broadcast_to(torch.as_tensor(noises, dtype=torch.float32, device=dev), (n_rays,)).contiguous()

    S = total_samples
    counters = torch.empty(2, dtype=torch.int32, device=dev)
    ray_is_valid = torch.empty(n_rays, dtype=torch.bool, device=dev)
    rays_n_samples = torch.empty(n_rays, dtype=torch.int32, device=dev)
    rays_sample_startidx = torch.empty(n_rays, dtype=torch.int32, device=dev)
    idcs = torch.empty(S, dtype=torch.int32, device=dev)
    xyzs = torch.empty(S, 3, dtype=torch.float32, device=dev)
    dirs = torch.empty(S, 3, dtype=torch.float32, device=dev)
    dss = torch.empty(S, dtype=torch.float32, device=dev)
    z_vals = torch.empty(S, dtype=torch.float32, device=dev)
    _lib.call(
        "ngp_march_rays",
        [rays_o.contiguous(), rays_d.contiguous(), t_starts.contiguous(), t_ends.contiguous(), noises,
         occupancy_bitfield.contiguous(),
         counters[0:1], counters[1:2], ray_is_valid, rays_n_samples, rays_sample_startidx, idcs, xyzs, dirs, dss,
         z_vals],
        descriptors.make_marching_descriptor(n_rays, total_samples, diagonal_n_steps, K, G, bound, stepsize_portion),
    )
    if raw:
        return (counters[0:1], counters[1:2], ray_is_valid, rays_n_samples, rays_sample_startidx, idcs, xyzs, dirs,
                dss, z_vals)
    measured_batch_size_before_compaction = counters[0] - counters[1]  # marching/__init__.py:91
    return (measured_batch_size_before_compaction, ray_is_valid, rays_n_samples, rays_sample_startidx, idcs, xyzs,
            dirs, dss, z_vals)


def march_rays_inference(
    # static
    diagonal_n_steps: int,
    K: int,
    G: int,
    march_steps_cap: int,
    bound: float,
    stepsize_portion: float,

    # inputs
    rays_o: torch.Tensor,
    rays_d: torch.Tensor,
    t_starts: torch.Tensor,
    t_ends: torch.Tensor,
    occupancy_bitfield: torch.Tensor,
    next_ray_index_in: torch.Tensor,
    terminated: torch.Tensor,
    indices: torch.Tensor,
):
    """Same contract as the reference's ``march_rays_inference`` (marching/__init__.py:96-157).

    Returns ``(next_ray_index, indices, n_samples, t_starts, xyzs, dss, z_vals)`` where ``t_starts``
    is the full ``[n_total_rays]`` array with the advanced values scattered in (:156).
    """
    n_total_rays, n_rays = rays_o.shape[0], terminated.shape[0]
    # marching/abstract.py:93-101
    assert_shape(rays_o, (n_total_rays, 3), "rays_o")
    assert_shape(rays_d, (n_total_rays, 3), "rays_d")
    assert_shape(t_starts, (n_total_rays,), "t_starts")
    assert_shape(t_ends, (n_total_rays,), "t_ends")
    assert_shape(occupancy_bitfield, (K * G * G * G // 8,), "occupancy_bitfield")
    if occupancy_bitfield.dtype != torch.uint8:
        raise AssertionError(f"occupancy_bitfield must be uint8, got {occupancy_bitfield.dtype}")
    assert_shape(next_ray_index_in, (1,), "next_ray_index_in")
    assert_shape(indices, (n_rays,), "indices")
    if terminated.dtype != torch.bool:
        raise AssertionError(f"terminated must be bool, got {terminated.dtype}")
    dev = rays_o.device
    cap = march_steps_cap
    next_ray_index = torch.empty(1, dtype=torch.int32, device=dev)
    indices_out = torch.empty(n_rays, dtype=torch.int32, device=dev)
    n_samples = torch.empty(n_rays, dtype=torch.int32, device=dev)
    t_starts_out = torch.empty(n_rays, dtype=torch.float32, device=dev)
    xyzs = torch.empty(n_rays, cap, 3, dtype=torch.float32, device=dev)
    dss = torch.empty(n_rays, cap, dtype=torch.float32, device=dev)
    z_vals = torch.empty(n_rays, cap, dtype=torch.float32, device=dev)
    _lib.call(
        "ngp_march_rays_inference",
        [rays_o.contiguous(), rays_d.contiguous(), t_starts.contiguous(), t_ends.contiguous(),
         occupancy_bitfield.contiguous(), as_u32(next_ray_index_in, "next_ray_index_in").contiguous(),
         terminated.contiguous(), as_u32(indices, "indices").contiguous(),
         next_ray_index, indices_out, n_samples, t_starts_out, xyzs, dss, z_vals],
        descriptors.make_marching_inference_descriptor(n_total_rays, n_rays, diagonal_n_steps, K, G, cap, bound,
                                                       stepsize_portion),
    )
    # marching/__init__.py:156: t_starts.at[indices].set(t_starts_out); out-of-range indices are dropped
    t_new = scatter_rows(t_starts, indices_out, t_starts_out, n_total_rays)
    return next_ray_index, indices_out, n_samples, t_new, xyzs, dss, z_vals


def scatter_rows(dst, idx, src, n_total):
    """``dst.at[idx].set(src)`` with jax's drop-out-of-bounds semantics (idx are uint32 bits).
    No host synchronisation: out-of-range writers are routed to a scratch row that is cut off."""
    idx64 = idx.to(torch.int64) & 0xFFFFFFFF
    idx64 = torch.where(idx64 < n_total, idx64, torch.full_like(idx64, n_total))
    out = torch.cat([dst, dst.new_zeros((1,) + tuple(dst.shape[1:]))], dim=0)
    out.index_copy_(0, idx64, src)
    return out[:n_total]
