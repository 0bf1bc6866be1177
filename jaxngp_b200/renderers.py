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


def _max_loop_passes(n_pixels, n_slots, diagonal_n_steps, bound, cap):
    """Upper bound on slot-refill passes: every ray needs at most ceil(per-ray step cap / cap) + 1 passes in a slot and
    the slots work through ceil(n_pixels / n_slots) rounds of rays; doubled for the power-of-two batching."""
    per_ray = math.ceil(diagonal_n_steps * max(bound, 1.0) * 4 / cap) + 2
    return 4 * per_ray * (math.ceil(n_pixels / max(n_slots, 1)) + 1)


@torch.no_grad()
def render_image_inference(nerf, cam, transform_cw, occupancy_bitfield, *, bg=(1.0, 1.0, 1.0), diagonal_n_steps=1024,
                           K=1, G=128, bound=1.0, stepsize_portion=0.0, march_steps_cap=8, n_rays=8192, grouped=True):
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
    passes, max_passes = 0, _max_loop_passes(n_pixels, n_rays, diagonal_n_steps, bound, march_steps_cap)
    while n_rendered < n_pixels:  # cuda.py:326
        iters = 2 ** (int(math.log2(max(1, (n_pixels - n_rendered) // n_rays))) + 1)  # cuda.py:327-328
        passes += iters
        if passes > max_passes:
            raise RuntimeError(f"render_image_inference: {n_rendered}/{n_pixels} rays finished after {passes} passes "
                               f"(bound {max_passes}): the slot-refill loop is not making progress")
        counts = []
        for _ in range(iters):
            next_ray_index, indices, n_samples, t_starts, xyzs, dss, z_vals = march_rays_inference(
                diagonal_n_steps=diagonal_n_steps, K=K, G=G, march_steps_cap=march_steps_cap, bound=bound,
                stepsize_portion=stepsize_portion, rays_o=o, rays_d=d, t_starts=t_starts, t_ends=t_ends,
                occupancy_bitfield=occupancy_bitfield, next_ray_index_in=next_ray_index, terminated=terminated,
                indices=indices)
            idx64 = (indices.to(torch.int64) & 0xFFFFFFFF).clamp(max=n_pixels - 1)
            if grouped and hasattr(nerf, "forward_grouped"):
                drgbs = nerf.forward_grouped(xyzs, d[idx64], n_samples)  # skips the padding rows of every slot
            else:
                dirs = d[idx64][:, None, :].expand(-1, march_steps_cap, -1)  # cuda.py:222-228
                drgbs, _ = nerf(xyzs, dirs, None)
            cnt, terminated, rays_rgbd, rays_T = integrate_rays_inference(
                rays_bg=rays_bg, rays_rgbd=rays_rgbd, rays_T=rays_T, n_samples=n_samples, indices=indices, dss=dss,
                z_vals=z_vals, drgbs=drgbs.contiguous())
            counts.append(cnt)
        n_rendered += int(torch.stack(counts).sum())  # one host sync per batch of iterations
    rgb = (rays_rgbd[:, :3].clamp(0, 1) * 255 + 0.5).to(torch.uint8).reshape(cam["height"], cam["width"], 3)
    return rgb, rays_rgbd[:, 3].reshape(cam["height"], cam["width"])


class InferenceRenderer:
    """``render_image_inference`` (models/renderers/cuda.py:244-373) with the per-iteration work of the
    slot-refill loop (march_rays_inference -> NeRF -> integrate_rays_inference -> scatter back,
    cuda.py:180-241) captured ONCE in a CUDA graph over preallocated state, so a frame costs one ray
    generation launch plus one graph replay per loop iteration and one host read per batch of iterations
    (the reference's own `n_rendered_rays` check, cuda.py:326,361).  Same ops, same numbers: the image is
    bit-identical to ``render_image_inference``."""

    def __init__(self, nerf, cam, occupancy_bitfield, *, bg=(1.0, 1.0, 1.0), diagonal_n_steps=1024, K=1, G=128,
                 bound=1.0, stepsize_portion=0.0, march_steps_cap=16, n_rays=262144, pixel_indices=None, skip_empty=True,
                 persistent=None):
        from . import _lib, descriptors
        self.nerf, self.cam, self.bits = nerf, cam, occupancy_bitfield
        dev = occupancy_bitfield.device
        self.dev, self.bound, self.cap = dev, bound, march_steps_cap
        self.skip_empty = skip_empty
        if pixel_indices is None:
            pixel_indices = torch.arange(cam["width"] * cam["height"], dtype=torch.int32, device=dev)
        self.pixels = pixel_indices.to(torch.int32).contiguous()  # the rays this renderer owns (tile sharding)
        N = self.N = self.pixels.shape[0]
        n = self.n = min(n_rays, N)
        f32, i32 = torch.float32, torch.int32
        self.o, self.d = torch.empty(N, 3, dtype=f32, device=dev), torch.empty(N, 3, dtype=f32, device=dev)
        self.t_starts, self.t_ends = torch.empty(N + 1, dtype=f32, device=dev), torch.empty(N, dtype=f32, device=dev)
        self.rays_rgbd, self.rays_T = torch.empty(N + 1, 4, dtype=f32, device=dev), torch.empty(N + 1, dtype=f32, device=dev)
        self.rays_bg = torch.tensor(bg, dtype=f32, device=dev).expand(N, 3).contiguous()
        self.terminated = torch.empty(n, dtype=torch.bool, device=dev)
        self.indices = torch.empty(n, dtype=i32, device=dev)
        self.next_in = torch.empty(1, dtype=i32, device=dev)
        self.ray_dirs = torch.empty(n, 3, dtype=f32, device=dev)
        self.n_samples = torch.empty(n, dtype=i32, device=dev)
        self.xyzs = torch.empty(n, self.cap, 3, dtype=f32, device=dev)
        self.dss, self.z_vals = torch.empty(n, self.cap, dtype=f32, device=dev), torch.empty(n, self.cap, dtype=f32, device=dev)
        self.counters = torch.zeros(2, dtype=torch.int64, device=dev)  # rays terminated, samples marched (this frame)
        self.pose = torch.empty(1, 12, dtype=f32, device=dev)
        self._march_desc = descriptors.make_marching_inference_descriptor(N, n, diagonal_n_steps, K, G, self.cap, bound,
                                                                          stepsize_portion)
        self._integ_desc = descriptors.make_integrating_inference_descriptor(N, n, self.cap)
        self._rays_desc = descriptors.make_training_rays_descriptor(N, cam["width"], cam["height"], 1, cam["fx"], cam["fy"],
                                                                    cam["cx"], cam["cy"], bound)
        self._graph = None
        self._host_counters = self._host_events = None
        self._max_passes = 2 * _max_loop_passes(N, n, diagonal_n_steps, bound, self.cap)
        # The whole slot-refill loop as ONE persistent kernel (csrc/mlp_fwd_umma.cu nerf_render_frame_kernel, SURVEY 8
        # f4): default wherever it applies -- the fused tcgen05 encoder + MLP geometry, 16 samples per slot and pass.
        from . import encoders, nerf as nerf_mod
        enc_mod = getattr(nerf, "position_encoder", None)
        can = (enc_mod is not None and self.cap == 16 and getattr(nerf, "fused", False) and getattr(nerf, "grouped_impl", "umma") == "umma"
               and nerf_mod.fused_supported(enc_mod.levels, enc_mod.latents, enc_mod.wrap))
        self.persistent = can if persistent is None else bool(persistent)
        if self.persistent and not can:
            raise ValueError("the persistent frame kernel needs the fused hash-grid + MLP geometry (dim 3, L 16, F 2, power-of-two T), "
                             "march_steps_cap = 16 and grouped_impl = 'umma'")
        if self.persistent:
            self._frame_desc = lambda: (encoders._a1_descriptor(enc_mod.levels, 0, bound, enc_mod.wrap, enc_mod.latents.dtype)
                                        + self._march_desc)
            self._next_ray = torch.zeros(1, dtype=i32, device=dev)

    def _iteration(self):
        """One pass of the slot-refill loop (cuda.py:180-241): three custom calls, no torch glue -- the scatters back
        into the frame state are folded into the kernels (``*_inplace`` entry points, include/ngp_b200.h)."""
        from . import _lib
        _lib.call("ngp_march_rays_inference_inplace",
                  [self.o, self.d, self.t_starts, self.t_ends, self.bits, self.next_in, self.terminated, self.indices,
                   self.n_samples, self.xyzs, self.dss, self.z_vals, self.ray_dirs], self._march_desc)
        drgbs = self.nerf.forward_grouped(self.xyzs, self.ray_dirs, self.n_samples)
        _lib.call("ngp_integrate_rays_inference_inplace",
                  [self.rays_bg, self.rays_rgbd, self.rays_T, self.n_samples, self.indices, self.dss, self.z_vals, drgbs,
                   self.terminated, self.counters], self._integ_desc)

    @torch.no_grad()
    def render(self, transform_cw):
        """Returns (rgb u8 [N, 3], depth f32 [N]) for the renderer's pixels."""
        from . import _lib
        self.pose.copy_(transform_cw.reshape(1, 12))
        _lib.call("ngp_make_training_rays", [self.pixels, self.pose, self.o, self.d, self.t_starts, self.t_ends],
                  self._rays_desc)
        return self.render_current_rays()

    @torch.no_grad()
    def render_rays(self, o, d, t_starts, t_ends):
        """Same, for caller-supplied rays (parity tests feed the rays of ``make_rays_worldspace``)."""
        self.o.copy_(o)
        self.d.copy_(d)
        self.t_starts[: self.N].copy_(t_starts)
        self.t_ends.copy_(t_ends)
        return self.render_current_rays()

    @torch.no_grad()
    def render_current_rays(self):
        from . import _lib
        if self.skip_empty:  # walk every ray to its first occupied point once, thread per ray (same samples, bit for bit)
            _lib.call("ngp_march_rays_skip_empty", [self.o, self.d, self.t_starts, self.t_ends, self.bits, self.t_starts],
                      self._march_desc)
        if self.persistent:  # one launch: refill, march, encoder + MLP and compositing never leave the SM
            enc_mod = self.nerf.position_encoder
            _lib.call("ngp_render_frame",
                      [self.o, self.d, self.t_starts, self.t_ends, self.bits, self.rays_bg, enc_mod.latents.detach(),
                       self.nerf.mlp_flat.detach(), self._next_ray, self.counters, self.rays_rgbd], self._frame_desc())
            rgb = (self.rays_rgbd[: self.N, :3].clamp(0, 1) * 255 + 0.5).to(torch.uint8)
            return rgb, self.rays_rgbd[: self.N, 3]
        self.rays_rgbd.zero_()
        self.rays_T.fill_(1.0)
        self.terminated.fill_(True)
        self.indices.zero_()
        self.next_in.zero_()
        self.counters.zero_()
        if self._graph is None:
            side = torch.cuda.Stream(device=self.dev)
            # snapshot BEFORE the side stream is released: its warm-up iteration mutates this state, and a clone
            # enqueued after wait_stream could run behind it (then the "undo" below would restore the mutated state)
            state = [t.clone() for t in (self.t_starts, self.rays_rgbd, self.rays_T, self.terminated, self.indices, self.next_in)]
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                self._iteration()  # warm-up on the capture stream
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph, stream=side):
                    self._iteration()
            torch.cuda.current_stream(self.dev).wait_stream(side)
            for t, s in zip((self.t_starts, self.rays_rgbd, self.rays_T, self.terminated, self.indices, self.next_in), state):
                t.copy_(s)  # undo the two warm-up iterations
            self.counters.zero_()
        # cuda.py:326-361 with the host check one batch behind: the counter of batch k is copied to pinned memory and
        # read while batch k+1 already runs, so the GPU never waits for the host.  The one surplus batch enqueued
        # when the frame turns out to be complete runs on idle slots only (no ray index < N): no state changes.
        if self._host_counters is None:
            self._host_counters = [torch.zeros(2, dtype=torch.int64).pin_memory() for _ in range(2)]
            self._host_events = [torch.cuda.Event() for _ in range(2)]
        n_rendered, passes, batch, pending = 0, 0, 0, None
        while True:
            iters = 2 ** (int(math.log2(max(1, (self.N - n_rendered) // self.n))) + 1)
            passes += iters
            if passes > self._max_passes:
                raise RuntimeError(f"InferenceRenderer: {n_rendered}/{self.N} rays finished after {passes} passes (bound "
                                   f"{self._max_passes}, counters {self.counters.tolist()}): the loop is not making progress")
            for _ in range(iters):
                self._graph.replay()
            slot = batch & 1
            self._host_counters[slot].copy_(self.counters, non_blocking=True)
            self._host_events[slot].record()
            if pending is not None:
                self._host_events[pending].synchronize()
                n_rendered = int(self._host_counters[pending][0])
                if n_rendered >= self.N:
                    break
            pending = slot
            batch += 1
        rgb = (self.rays_rgbd[: self.N, :3].clamp(0, 1) * 255 + 0.5).to(torch.uint8)
        return rgb, self.rays_rgbd[: self.N, 3]

    @property
    def samples_done(self):
        """Samples marched for the last frame (device scalar)."""
        return self.counters[1]
