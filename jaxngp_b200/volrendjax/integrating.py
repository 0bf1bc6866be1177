"""integrate_rays / integrate_rays_inference -- mirrors volrendjax/integrating/__init__.py:11-110 and
the custom_vjp in integrating/impl.py:50-146."""
from typing import Tuple

import torch

from .. import _lib, descriptors
from ._check import as_u32, assert_shape, non_negative, require_f32
from .marching import scatter_rows


def _integrate_fwd(start, ns, bgs, dss, z_vals, drgbs):
    n_rays, total_samples = start.shape[0], dss.shape[0]
    dev = drgbs.device
    mbs = torch.empty(1, dtype=torch.int32, device=dev)
    final_rgbds = torch.empty(n_rays, 4, dtype=torch.float32, device=dev)
    final_opacities = torch.empty(n_rays, dtype=torch.float32, device=dev)
    _lib.call("ngp_integrate_rays", [start, ns, bgs, dss, z_vals, drgbs, mbs, final_rgbds, final_opacities],
              descriptors.make_integrating_descriptor(n_rays, total_samples))
    return mbs, final_rgbds, final_opacities


def _integrate_bwd(near_distance, start, ns, bgs, dss, z_vals, drgbs, final_rgbds, final_opacities, dL_dfinal_rgbds):
    n_rays, total_samples = start.shape[0], dss.shape[0]
    dev = drgbs.device
    dL_dbgs = torch.empty(n_rays, 3, dtype=torch.float32, device=dev)
    dL_dz_vals = torch.empty(total_samples, dtype=torch.float32, device=dev)
    dL_ddrgbs = torch.empty(total_samples, 4, dtype=torch.float32, device=dev)
    _lib.call("ngp_integrate_rays_backward",
              [start, ns, bgs, dss, z_vals, drgbs, final_rgbds, final_opacities, dL_dfinal_rgbds.contiguous(),
               dL_dbgs, dL_dz_vals, dL_ddrgbs],
              descriptors.make_integrating_backward_descriptor(n_rays, total_samples, near_distance))
    return dL_dbgs, dL_dz_vals, dL_ddrgbs


class _IntegrateRays(torch.autograd.Function):
    """custom_vjp of integrating/impl.py:50-146.  The reference's bwd binds ``dL_dbgs`` to ``dss`` and
    gives ``bgs`` no gradient (impl.py:126-141); here the cotangents go where they belong:
    (bgs, z_vals, drgbs) <- (dL_dbgs, dL_dz_vals, dL_ddrgbs), dss gets none."""

    @staticmethod
    def forward(ctx, near_distance, start, ns, bgs, dss, z_vals, drgbs):
        mbs, final_rgbds, final_opacities = _integrate_fwd(start, ns, bgs, dss, z_vals, drgbs)
        ctx.near_distance = near_distance
        ctx.save_for_backward(start, ns, bgs, dss, z_vals, drgbs, final_rgbds, final_opacities)
        ctx.mark_non_differentiable(mbs)
        return mbs, final_rgbds, final_opacities

    @staticmethod
    def backward(ctx, _g_mbs, dL_dfinal_rgbds, _dL_dfinal_opacities):
        start, ns, bgs, dss, z_vals, drgbs, final_rgbds, final_opacities = ctx.saved_tensors
        if dL_dfinal_rgbds is None:
            dL_dfinal_rgbds = torch.zeros_like(final_rgbds)
        dL_dbgs, dL_dz_vals, dL_ddrgbs = _integrate_bwd(ctx.near_distance, start, ns, bgs, dss, z_vals, drgbs,
                                                        final_rgbds, final_opacities, dL_dfinal_rgbds)
        return None, None, None, dL_dbgs, None, dL_dz_vals, dL_ddrgbs


def integrate_rays(
    near_distance: float,
    rays_sample_startidx: torch.Tensor,
    rays_n_samples: torch.Tensor,
    bgs: torch.Tensor,
    dss: torch.Tensor,
    z_vals: torch.Tensor,
    drgbs: torch.Tensor,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Same contract as the reference's ``integrate_rays`` (integrating/__init__.py:11-59); returns
    ``(measured_batch_size, final_rgbds, final_opacities)``, differentiable w.r.t. drgbs, z_vals, bgs."""
    start, ns = as_u32(rays_sample_startidx, "rays_sample_startidx"), as_u32(rays_n_samples, "rays_n_samples")
    n_rays, total_samples = start.shape[0], dss.shape[0]
    dev = drgbs.device
    bgs = torch.as_tensor(bgs, dtype=torch.float32, device=dev)
    try:
        bgs = torch.broadcast_to(bgs, (n_rays, 3))  # impl.py:60 (jnp.broadcast_to raises ValueError when it cannot)
    except RuntimeError as exc:
        raise ValueError(f"bgs of shape {tuple(bgs.shape)} cannot be broadcast to ({n_rays}, 3)") from exc
    # integrating/abstract.py:16-29,65-81
    assert_shape(ns, (n_rays,), "rays_n_samples")
    assert_shape(z_vals, (total_samples,), "z_vals")
    assert_shape(drgbs, (total_samples, 4), "drgbs")
    non_negative(near_distance, "near_distance")
    require_f32(drgbs, "integrate_rays", "input prediction (density, color)")
    mbs, final_rgbds, final_opacities = _IntegrateRays.apply(
        float(near_distance), start.contiguous(), ns.contiguous(), bgs.contiguous(), dss.contiguous(),
        z_vals.contiguous(), drgbs.contiguous())
    return mbs[0], final_rgbds, final_opacities


def integrate_rays_inference(
    rays_bg: torch.Tensor,
    rays_rgbd: torch.Tensor,
    rays_T: torch.Tensor,

    n_samples: torch.Tensor,
    indices: torch.Tensor,
    dss: torch.Tensor,
    z_vals: torch.Tensor,
    drgbs: torch.Tensor,
):
    """Same contract as the reference's ``integrate_rays_inference`` (integrating/__init__.py:62-110);
    returns ``(terminate_cnt, terminated, rays_rgbd, rays_T)`` with the per-slot results scattered
    back into the full arrays (:108-109)."""
    n_total_rays = rays_rgbd.shape[0]
    n_rays, cap = dss.shape
    # integrating/abstract.py:107-114
    assert_shape(rays_bg, (n_total_rays, 3), "rays_bg")
    assert_shape(rays_rgbd, (n_total_rays, 4), "rays_rgbd")
    assert_shape(rays_T, (n_total_rays,), "rays_T")
    assert_shape(n_samples, (n_rays,), "n_samples")
    assert_shape(indices, (n_rays,), "indices")
    assert_shape(z_vals, (n_rays, cap), "z_vals")
    assert_shape(drgbs, (n_rays, cap, 4), "drgbs")
    dev = drgbs.device
    terminate_cnt = torch.empty(1, dtype=torch.int32, device=dev)
    terminated = torch.empty(n_rays, dtype=torch.bool, device=dev)
    rgbd_out = torch.empty(n_rays, 4, dtype=torch.float32, device=dev)
    T_out = torch.empty(n_rays, dtype=torch.float32, device=dev)
    idx = as_u32(indices, "indices").contiguous()
    _lib.call("ngp_integrate_rays_inference",
              [rays_bg.contiguous(), rays_rgbd.contiguous(), rays_T.contiguous(),
               as_u32(n_samples, "n_samples").contiguous(), idx, dss.contiguous(), z_vals.contiguous(),
               drgbs.contiguous(), terminate_cnt, terminated, rgbd_out, T_out],
              descriptors.make_integrating_inference_descriptor(n_total_rays, n_rays, cap))
    rays_rgbd = scatter_rows(rays_rgbd, idx, rgbd_out, n_total_rays)
    rays_T = scatter_rows(rays_T, idx, T_out, n_total_rays)
    return terminate_cnt[0], terminated, rays_rgbd, rays_T
