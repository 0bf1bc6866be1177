"""Host-side mirror of models/renderers/cuda.py: ray generation, AABB near/far, the training renderer
(march -> NeRF -> integrate) and the inference renderer (slot-refill loop of march_rays_inference ->
NeRF -> integrate_rays_inference).  torch CUDA tensors throughout; no host synchronisation except
the loop-termination read the reference also has (cuda.py:326,361)."""
import math

import torch

from .volrendjax import integrate_rays, integrate_rays_inference, march_rays, march_rays_inference


def make_ray_directions(x, y, cam):
    """utils/types.py:398-439 for an undistorted PERSPECTIVE camera; x, y integer pixel coordinates."""
    dx = ((x.to(torch.float32) + 0.5) - cam["cx"]) / cam["fx"]
    dy = ((y.to(torch.float32) + 0.5) - cam["cy"]) / cam["fy"]
    d = torch.stack([dx, -dy, -torch.ones_like(dx)], dim=-1)  # CV -> CG axis flip
    return d / torch.linalg.norm(d, dim=-1, keepdim=True)


def make_rays_worldspace(cam, transform_cw):
    """cuda.py:22-53; transform_cw = [12] (R row-major, t)."""
    idx = torch.arange(cam["width"] * cam["height"], device=transform_cw.device)
    d_cam = make_ray_directions(idx % cam["width"], idx // cam["width"], cam)
    R = transform_cw[:9].reshape(3, 3)
    return transform_cw[9:].expand_as(d_cam).contiguous(), (d_cam @ R.T).contiguous()


def make_near_far_from_bound(bound: float, o: torch.Tensor, d: torch.Tensor):
    """cuda.py:57-97."""
    eps = 1e-15
    d = torch.where(torch.signbit(d), torch.clamp(d, max=-eps), torch.clamp(d, min=eps))
    t0, t1 = (-bound - o) / d, (bound - o) / d
    t_start = torch.minimum(t0, t1).amax(dim=-1)
    t_end = torch.maximum(t0, t1).amin(dim=-1)
    return torch.clamp(t_start, min=0.0), t_end


def render_rays_train(nerf, o_world, d_world, bg, total_samples, occupancy_bitfield, *, diagonal_n_steps=1024, K=1,
                      G=128, bound=1.0, stepsize_portion=0.0, near=0.3, noises=0.0):
    """cuda.py:101-162: returns (batch_metrics, final_rgbds, tv)."""
    t_starts, t_ends = make_near_far_from_bound(bound, o_world, d_world)
    mb, ray_is_valid, rays_n, rays_start, ray_idcs, xyzs, dirs, dss, z_vals = march_rays(
        total_samples=total_samples, diagonal_n_steps=diagonal_n_steps, K=K, G=G, bound=bound,
        stepsize_portion=stepsize_portion, rays_o=o_world, rays_d=d_world, t_starts=t_starts, t_ends=t_ends,
        noises=noises, occupancy_bitfield=occupancy_bitfield)
    drgbs, tv = nerf(xyzs, dirs, None)
    effective, final_rgbds, _ = integrate_rays(near_distance=near, rays_sample_startidx=rays_start,
                                               rays_n_samples=rays_n, bgs=bg, dss=dss, z_vals=z_vals, drgbs=drgbs)
    metrics = dict(n_valid_rays=ray_is_valid.sum(), ray_is_valid=ray_is_valid,
                   measured_batch_size_before_compaction=mb, measured_batch_size=effective)
    return metrics, final_rgbds, tv


@torch.no_grad()
def render_image_inference(nerf, cam, transform_cw, occupancy_bitfield, *, bg=(1.0, 1.0, 1.0), diagonal_n_steps=1024,
                           K=1, G=128, bound=1.0, stepsize_portion=0.0, march_steps_cap=8, n_rays=8192):
    """cuda.py:244-373: the reference's host loop, op for op.  Returns (rgb u8 [H, W, 3], depth f32 [H, W])."""
    o, d = make_rays_worldspace(cam, transform_cw)
    n_pixels = o.shape[0]
    dev = o.device
    t_starts, t_ends = make_near_far_from_bound(bound, o, d)
    rays_rgbd = torch.zeros(n_pixels, 4, device=dev)
    rays_T = torch.ones(n_pixels, device=dev)
    rays_bg = torch.tensor(bg, dtype=torch.float32, device=dev).expand(n_pixels, 3).contiguous()
    n_rays = min(n_rays, n_pixels)
    terminated = torch.ones(n_rays, dtype=torch.bool, device=dev)
    indices = torch.zeros(n_rays, dtype=torch.int32, device=dev)
    next_ray_index = torch.zeros(1, dtype=torch.int32, device=dev)
    n_rendered = 0
    while n_rendered < n_pixels:  # cuda.py:326
        iters = 2 ** (int(math.log2(max(1, (n_pixels - n_rendered) // n_rays))) + 1)  # cuda.py:327-328
        counts = []
        for _ in range(iters):
            next_ray_index, indices, n_samples, t_starts, xyzs, dss, z_vals = march_rays_inference(
                diagonal_n_steps=diagonal_n_steps, K=K, G=G, march_steps_cap=march_steps_cap, bound=bound,
                stepsize_portion=stepsize_portion, rays_o=o, rays_d=d, t_starts=t_starts, t_ends=t_ends,
                occupancy_bitfield=occupancy_bitfield, next_ray_index_in=next_ray_index, terminated=terminated,
                indices=indices)
            idx64 = (indices.to(torch.int64) & 0xFFFFFFFF).clamp(max=n_pixels - 1)
            dirs = d[idx64][:, None, :].expand(-1, march_steps_cap, -1)
            drgbs, _ = nerf(xyzs, dirs, None)
            cnt, terminated, rays_rgbd, rays_T = integrate_rays_inference(
                rays_bg=rays_bg, rays_rgbd=rays_rgbd, rays_T=rays_T, n_samples=n_samples, indices=indices, dss=dss,
                z_vals=z_vals, drgbs=drgbs.contiguous())
            counts.append(cnt)
        n_rendered += int(torch.stack(counts).sum())  # one host sync per batch of iterations
    rgb = (rays_rgbd[:, :3].clamp(0, 1) * 255 + 0.5).to(torch.uint8).reshape(cam["height"], cam["width"], 3)
    return rgb, rays_rgbd[:, 3].reshape(cam["height"], cam["width"])
