"""Device-resident training state and step for the reference's NeRF training path
(app/nerf/_utils.py:80-170 ``train_step``, app/nerf/train.py:28-96 ``train_epoch``), ray-sharded data
parallel over NCCL when launched with one process per GPU.

Per step (shapes at BASELINE config C2: n_rays = total_samples = 2^18):
  perm[n_rays] -> rays (o, d) -> near/far -> march_rays* -> hash-grid encode* -> MLP -> integrate_rays*
  -> Huber(0.1) over valid rays -> integrate_rays_backward* -> MLP backward -> hash-grid scatter*
  -> [all-reduce of the flat gradient buffer across ranks] -> Adam*            (* = this package's kernels)

All parameters live in ONE flat f32 buffer [hash table | MLP weights] with matching flat gradient and
Adam-moment buffers, so data parallelism is a single ``all_reduce`` and the optimizer a single launch.
"""
import math
import os

import torch

from . import _lib, descriptors, dp, encoders, exchange as exchange_mod, nerf as nerf_mod, ogrid, renderers, synthetic, trainops
from .volrendjax import integrate_rays, march_rays
from .volrendjax.integrating import _integrate_bwd, _integrate_fwd


def huber(pred, target, delta=0.1):
    """optax.huber_loss (app/nerf/_utils.py:153)."""
    err = (pred - target).abs()
    quad = torch.clamp(err, max=delta)
    return 0.5 * quad * quad + delta * (err - quad)


class Scene:
    """Device-resident training set: transforms [V, 12] f32 and RGBA pixels [V*H*W, 4] u8
    (utils/types.py:924-1116 SceneData), rendered from the procedural field of synthetic.py."""

    def __init__(self, device, n_views=100, width=synthetic.W, height=synthetic.H):
        self.cam = synthetic.camera()
        if (width, height) != (synthetic.W, synthetic.H):
            s = width / synthetic.W
            self.cam = dict(width=width, height=height, fx=self.cam["fx"] * s, fy=self.cam["fy"] * s, cx=width / 2,
                            cy=height / 2, near=synthetic.NEAR)
        self.n_views, self.width, self.height = n_views, width, height
        self.transforms = torch.from_numpy(synthetic.poses(n_views)).to(device)
        self.bitfield_gt = torch.from_numpy(synthetic.occupancy_bitfield()).to(device)
        self.rgbas_u8 = self._render_ground_truth(device)

    @property
    def n_pixels(self):
        return self.n_views * self.width * self.height

    def rays(self, perm):
        """app/nerf/_utils.py:93-115."""
        hw = self.width * self.height
        view, pix = perm // hw, perm % hw
        d_cam = renderers.make_ray_directions(pix % self.width, pix // self.width, self.cam)
        tf = self.transforms[view]
        d_world = (d_cam[:, None, :] * tf[:, :9].reshape(-1, 3, 3)).sum(-1)
        return tf[:, 9:].contiguous(), d_world.contiguous()

    @torch.no_grad()
    def _render_ground_truth(self, device, chunk=1 << 17, budget=1 << 25):
        """Renders every training pixel of the analytic field with this package's own march and
        integrate kernels (bg = 0) and stores straight-alpha RGBA like the blender PNGs."""
        out = torch.empty(self.n_pixels, 4, dtype=torch.uint8, device=device)
        lo = torch.tensor([-0.7, -0.3, -0.6], device=device)
        hi = torch.tensor([-0.2, 0.3, 0.1], device=device)
        for begin in range(0, self.n_pixels, chunk):
            perm = torch.arange(begin, min(begin + chunk, self.n_pixels), device=device)
            o, d = self.rays(perm)
            ts, te = renderers.make_near_far_from_bound(synthetic.BOUND, o, d)
            _, _, rn, rs, _, xyzs, _, dss, zs = march_rays(budget, synthetic.DIAGONAL_N_STEPS, synthetic.K, synthetic.G,
                                                           synthetic.BOUND, synthetic.STEPSIZE_PORTION, o, d, ts, te,
                                                           0.0, self.bitfield_gt)
            inside = ((xyzs ** 2).sum(-1) < 0.45 ** 2) | ((xyzs > lo) & (xyzs < hi)).all(-1)
            drgbs = torch.cat([torch.where(inside, synthetic.SIGMA_INSIDE, 0.0)[:, None],
                               0.5 + 0.5 * xyzs.clamp(-1, 1)], dim=-1).contiguous()
            _, rgbd, opac = integrate_rays(synthetic.NEAR, rs, rn, torch.zeros(3, device=device), dss, zs, drgbs)
            # integrate_rays renormalises terminated rays by opacity (integrating.cu:85-90): undo that to get
            # premultiplied colour, then store straight alpha
            premult = torch.where((opac >= 1 - 1e-4)[:, None], rgbd[:, :3] * opac[:, None], rgbd[:, :3])
            straight = premult / opac.clamp(min=1e-6)[:, None]
            rgba = torch.cat([straight.clamp(0, 1), opac[:, None]], dim=-1)
            out[begin:begin + perm.shape[0]] = (rgba * 255 + 0.5).to(torch.uint8)
        return out


class Trainer:
    def __init__(self, device="cuda:0", n_rays=1 << 18, total_samples=1 << 18, lr=1e-2, seed=1000000007, rank=0,
                 world_size=1, process_group=None, scene=None, T=1 << 19, use_graph=True, fused_mlp=True, fused_glue=None, fused_encoder=None,
                 exchange=None):
        self.device = torch.device(device)
        self.n_rays, self.total_samples = n_rays, total_samples
        self.rank, self.world_size, self.pg = rank, world_size, process_group
        # how the flat gradient is exchanged at world_size > 1: "peer" / "peer-p2p" (exchange.py: ONE fused kernel per rank
        # over NVLink peer / multicast memory, replayed as a captured graph; the default -- measured 0.594 vs 0.646 ms per
        # step at N=2, replicas bit-identical and equal to the NCCL arm) or "nccl" (dp.py: reduce-scatter, Adam,
        # all-gather; also what "auto" falls back to when the ranks cannot map each other's memory)
        self.exchange_mode = exchange_mod.requested_mode() if exchange is None else exchange
        self.peer_exchange = None
        torch.backends.cuda.matmul.allow_tf32 = True  # XLA's default f32 dot precision on Ampere+
        gen = torch.Generator(device=self.device).manual_seed(seed)  # same init on every rank
        self.nerf = nerf_mod.NeRF(bound=synthetic.BOUND, inference=False, device=self.device, generator=gen, T=T)
        self.levels = self.nerf.position_encoder.levels
        self.fused_mlp = fused_mlp
        self.fused_glue = fused_mlp if fused_glue is None else fused_glue
        # Persistent CTAs per SM of the prefetched march (0 = a full grid).  It runs beneath the step's other kernels and must be
        # done before the next step needs it: C2 at N = 1, ms per step: 1 -> 0.498 (the march was the critical path), 2 -> 0.444,
        # 3 -> 0.449, 4 -> 0.471, full grid -> 0.470 (`profiles/backward_knobs_r02.txt`).  More than one rank: full grid, as measured.
        self.prefetch_march_ctas_per_sm = int(os.environ.get("NGP_B200_MARCH_CTAS_PER_SM", "2" if world_size == 1 else "0"))
        self.fused_loss = True  # ngp_integrate_loss_fused instead of integrate_rays / huber_loss_grad / integrate_rays_backward
        # The backward, three arrangements (C2, ms per step at N = 1 / N = 2, `profiles/backward_knobs_r02.txt`):
        #   * fused scatter (default where the fused encoder applies): ONE kernel, `ngp_nerf_mlp_backward_scatter` -- the table
        #     scatter is issued from the MLP backward's own d_enc fragments, d_enc never exists in memory: 0.497 / 0.511;
        #   * one pass: `ngp_nerf_mlp_backward` then `ngp_hashgrid_a1_backward`: 0.517 / 0.519;
        #   * a two-stream pipeline over `bwd_chunks` slices of the sample array (the MLP backward of slice c+1 beside the
        #     table scatter of slice c): 0.545 / 0.546 with 2 slices, 0.547 with 3, 0.574 with 4.
        # Before the MLP backward got its issuing warp (csrc/mlp.cu) the order was the reverse (pipeline 0.533, one pass
        # 0.562, fused 0.528): the kernel now keeps 288 x 160 registers, so fewer scatter CTAs fit beside it, and the
        # 37 us it no longer spends are worth more than the overlap was.
        self.bwd_fused_scatter = os.environ.get("NGP_B200_BWD_FUSED_SCATTER", "1") == "1"
        self.bwd_chunks = int(os.environ.get("NGP_B200_BWD_CHUNKS", "1"))
        # optional explicit split in "waves" of 148 x 128 samples (one block per SM of the MLP kernel), e.g. "9,5"
        self.bwd_waves = [int(w) for w in os.environ.get("NGP_B200_BWD_WAVES", "").split(",") if w]
        self._bwd_side = None
        self._flatten_parameters()
        self.fused_encoder = (fused_mlp and nerf_mod.fused_supported(self.levels, self.table)) if fused_encoder is None else fused_encoder
        self.scene = scene if scene is not None else Scene(self.device)
        # occupancy grid state (utils/types.py:93-144): all-ones bitfield at step 0
        self.grid = ogrid.OccupancyDensityGrid(synthetic.K, synthetic.G, device=self.device)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.step = 0
        # per-rank random streams (march perturbations + backgrounds; grid-update draws): Philox counters in device
        # memory, advanced by the launches that consume them -- graph replays and eager launches draw the same numbers
        self.rng_seed = seed + 1 + rank
        self.rng_state = trainops.new_rng_state(self.device)
        self.grid.seed = self.rng_seed
        # weight decay applies to the MLP weights only (_utils.py:45-77): index >= table_numel, shard relative
        decay_begin = min(max(self.table_numel - self.shard_lo, 0), self.shard_hi - self.shard_lo)
        self.adam_desc = descriptors.make_adam_descriptor(
            n=self.shard_hi - self.shard_lo, decay_begin=decay_begin, lr_init=lr, lr_end=lr / 100, decay_rate=1 / 3,
            transition_steps=10_000, transition_begin=10_000, staircase=True, b1=0.9, b2=0.99, eps=1e-15,
            eps_root=1e-15, weight_decay=1e-6, grad_scale=1.0 / world_size)
        self.use_graph = use_graph
        # the fused exchange kernel + step counter replay as one captured graph behind the compute graph
        # (NGP_B200_GRAPH_EXCHANGE=0: eager launches).  The NCCL arm stays eager: its collectives captured in a graph
        # measured no faster (0.655 vs 0.646 ms at N=2) and made process-group teardown hang.
        self.graph_exchange = os.environ.get("NGP_B200_GRAPH_EXCHANGE", "1") == "1"
        self._exchange_graph = None
        self._graph = self._march_graph = None
        self._static_perm = self._static_out = self._static_marched = None
        self._prefetched = None
        self._slot = 0
        self._shadow = None

    @property
    def occupancy(self):
        return self.grid.occupancy

    # -- the reference's update cadence (utils/types.py:1380-1396), for callers that run its training loop
    #    (app/nerf/train.py:46-88): every step at first, every 16 steps from step 240 on
    @property
    def update_ogrid_interval(self) -> int:
        return min(16, self.step // 16 + 1)

    @property
    def should_call_update_ogrid(self) -> bool:
        return self.step > 0 and self.step % self.update_ogrid_interval == 0

    @property
    def should_update_all_ogrid_cells(self) -> bool:
        return self.step < 256

    @property
    def occ_mask(self):
        return self.grid.occ_mask

    # -- flat parameter / gradient / moment buffers ---------------------------------------------
    def _flatten_parameters(self):
        enc = self.nerf.position_encoder
        self.table_numel = enc.latents.numel()
        assert self.table_numel % 4 == 0 and nerf_mod.MLP_NUMEL % 4 == 0
        self.n_params = self.table_numel + nerf_mod.MLP_NUMEL
        # padded so that every rank's shard is float4-aligned and the shards cover the buffer, for any world size
        quantum = math.lcm(32, 4 * self.world_size)
        total = -(-self.n_params // quantum) * quantum
        if self.world_size > 1 and self.exchange_mode != "nccl":  # symmetric buffers every rank maps (collective)
            try:
                self.peer_exchange = exchange_mod.PeerExchange(total, self.rank, self.world_size, self.device, self.pg,
                                                               mode="peer" if self.exchange_mode == "auto" else self.exchange_mode)
            except Exception as exc:  # no peer mapping on this box (every rank fails alike: same node, same driver)
                if self.exchange_mode != "auto":
                    raise
                import warnings
                warnings.warn(f"peer exchange unavailable ({exc}); using the NCCL reduce-scatter / all-gather exchange")
                self.exchange_mode = "nccl"
        if self.peer_exchange is not None:
            self.flat_params, self.flat_grads = self.peer_exchange.params, self.peer_exchange.grads
        else:
            self.flat_params = torch.zeros(total, dtype=torch.float32, device=self.device)
            self.flat_grads = torch.zeros_like(self.flat_params)
        # optimizer state: this rank's shard only (ZeRO-1, dp.py); the whole buffer on one GPU
        self.shard_lo, self.shard_hi = dp.shard_bounds(total, self.rank, self.world_size)
        self.adam_m = torch.zeros(self.shard_hi - self.shard_lo, dtype=torch.float32, device=self.device)
        self.adam_v = torch.zeros_like(self.adam_m)
        self.flat_params[: self.table_numel].copy_(enc.latents.detach().reshape(-1))
        self.flat_params[self.table_numel:self.n_params].copy_(self.nerf.mlp_flat.detach())
        enc.latents.data = self.flat_params[: self.table_numel].view_as(enc.latents)
        self.nerf.mlp_flat.data = self.flat_params[self.table_numel:self.n_params]
        self.table = enc.latents.data
        self.mlp_flat = self.nerf.mlp_flat.data
        self.table_grad = self.flat_grads[: self.table_numel].view_as(self.table)
        self.mlp_grad = self.flat_grads[self.table_numel:self.n_params]

    # -- one training step ------------------------------------------------------------------------
    def _march_body(self, perm, noises=None):
        """Ray generation + march_rays: the part of the step that does not depend on the parameters (it reads the
        rays and the occupancy bitfield only), so the next step's can overlap this step's gradient exchange and
        optimizer (train_step)."""
        sc = self.scene
        bg = None
        if noises is None:
            # both random inputs of the step leave the ray-generation kernel: the march perturbations (cuda.py:118-122)
            # and the random backgrounds the loss composites onto (_utils.py:134-136), which travel with the marched batch
            o, d, t_starts, t_ends, noises, bg = trainops.make_training_rays_rng(perm, sc.transforms, sc.cam, synthetic.BOUND,
                                                                                 self.rng_state, self.rng_seed)
        else:
            o, d, t_starts, t_ends = trainops.make_training_rays(perm, sc.transforms, sc.cam, synthetic.BOUND)
        return march_rays(self.total_samples, synthetic.DIAGONAL_N_STEPS, synthetic.K, synthetic.G, synthetic.BOUND,
                          synthetic.STEPSIZE_PORTION, o, d, t_starts, t_ends, noises, self.grid.occupancy, raw=True) + (bg,)

    def _compute_body(self, perm, marched, bg=None):
        """Encoder + MLP forward, integrate, loss, and the whole backward into the flat gradient buffer."""
        sc = self.scene
        nxt, exc, ray_is_valid, rays_n, rays_start, _, xyzs, dirs, dss, z_vals, marched_bg = marched
        if bg is None:
            if marched_bg is None:
                raise ValueError("supplied march perturbations need supplied backgrounds as well (the step draws both or none)")
            bg = marched_bg  # random_bg, _utils.py:134-136
        if self.fused_encoder:  # encoder gather feeding the MLP's first tensor-core fragments (enc written once, for the backward)
            drgbs, enc = nerf_mod.fused_forward(self.levels, xyzs, synthetic.BOUND, self.table, dirs, self.mlp_flat, want_enc=True)
        else:
            enc = encoders.hashgrid_forward(self.levels, xyzs, synthetic.BOUND, self.table)
            drgbs = nerf_mod.mlp_forward(enc, dirs, self.mlp_flat)
        if self.fused_loss:  # integrate, loss and integrate backward in one pass per ray (same bits as the three ops)
            effective, final_rgbds, final_opac, d_drgbs, loss, n_valid = trainops.integrate_loss_fused(
                synthetic.NEAR, rays_start, rays_n, bg, dss, z_vals, drgbs, ray_is_valid, perm, sc.rgbas_u8)
        else:
            effective, final_rgbds, final_opac = _integrate_fwd(rays_start, rays_n, bg, dss, z_vals, drgbs)
            d_final, loss, n_valid = trainops.huber_loss_grad(final_rgbds, ray_is_valid, perm, sc.rgbas_u8, bg)
            _, _, d_drgbs = _integrate_bwd(synthetic.NEAR, rays_start, rays_n, bg, dss, z_vals, drgbs, final_rgbds,
                                           final_opac, d_final)
        self._backward(enc, dirs, xyzs, d_drgbs)
        used = trainops.u32_axpy(nxt, exc, 1, -1, 0)  # next - exceeded, marching/__init__.py:91
        return dict(loss=loss[0], n_valid_rays=n_valid[0], measured_batch_size_before_compaction=used[0],
                    measured_batch_size=effective[0])

    def _backward(self, enc, dirs, xyzs, d_drgbs):
        """MLP backward + hash-table scatter into the flat gradient buffer, as one pass or as the pipeline of ``bwd_chunks``
        slices described in ``__init__`` (per-sample work: any split of the sample array gives the same sums, to atomic order)."""
        n, chunks = xyzs.shape[0], self.bwd_chunks
        if self.fused_mlp and self.bwd_fused_scatter and self.fused_encoder:
            nerf_mod.mlp_backward_scatter(self.levels, xyzs, synthetic.BOUND, enc, dirs, self.mlp_flat, d_drgbs, self.mlp_grad, self.table_grad)
            return
        if not self.fused_mlp or chunks <= 1 or n < 128 * 148 * chunks:
            d_enc, _ = nerf_mod.mlp_backward(enc, dirs, self.mlp_flat, d_drgbs, d_weights=self.mlp_grad)
            encoders.hashgrid_backward(self.levels, xyzs, synthetic.BOUND, d_enc, out=self.table_grad)
            return
        main = torch.cuda.current_stream(self.device)
        if self._bwd_side is None:
            self._bwd_side = torch.cuda.Stream(device=self.device)
        side = self._bwd_side
        d_enc = torch.empty(n, 32, dtype=torch.float32, device=self.device)
        if self.bwd_waves:  # explicit split in waves of 148 x 128 samples (one block per SM of the persistent MLP kernel)
            sizes = [w * 128 * 148 for w in self.bwd_waves]
        else:  # equal slices of whole 128-sample blocks
            sizes = [-(-n // chunks // 128) * 128] * chunks
        bounds, lo = [], 0
        for sz in sizes:
            if lo >= n:
                break
            bounds.append((lo, min(lo + sz, n)))
            lo += sz
        if lo < n:
            bounds[-1] = (bounds[-1][0], n)
        side.wait_stream(main)  # joins a stream capture in progress; orders the scatter behind whatever wrote table_grad last
        for c, (lo, hi) in enumerate(bounds):
            nerf_mod.mlp_backward(enc[lo:hi], dirs[lo:hi], self.mlp_flat, d_drgbs[lo:hi], d_weights=self.mlp_grad, d_enc=d_enc[lo:hi],
                                  accumulate=c > 0)
            done = torch.cuda.Event()
            done.record(main)
            side.wait_event(done)
            with torch.cuda.stream(side):
                encoders.hashgrid_backward(self.levels, xyzs[lo:hi], synthetic.BOUND, d_enc[lo:hi], out=self.table_grad, accumulate=c > 0)
        main.wait_stream(side)

    def _step_body(self, perm, noises=None, bg=None, apply=True):
        if not self.fused_glue:
            return self._step_body_torch(perm, noises, bg, apply)
        out = self._compute_body(perm, self._march_body(perm, noises), bg)
        if apply and self.world_size == 1:
            self._optimizer_step()
        return out

    def _step_body_torch(self, perm, noises=None, bg=None, apply=True):
        """Same step with the glue (ray generation, loss) in torch ops and autograd: the cross-check arm."""
        sc, dev = self.scene, self.device
        perm = perm.to(torch.int64)
        o, d = sc.rays(perm)
        t_starts, t_ends = renderers.make_near_far_from_bound(synthetic.BOUND, o, d)
        if noises is None:
            noises = torch.rand(self.n_rays, device=dev)  # cuda.py:118-122
        if bg is None:
            bg = torch.rand(self.n_rays, 3, device=dev)  # random_bg, _utils.py:134-136
        mb, ray_is_valid, rays_n, rays_start, _, xyzs, dirs, dss, z_vals = march_rays(
            self.total_samples, synthetic.DIAGONAL_N_STEPS, synthetic.K, synthetic.G, synthetic.BOUND,
            synthetic.STEPSIZE_PORTION, o, d, t_starts, t_ends, noises, self.grid.occupancy)
        # gather, MLP and compositing without autograd: each backward kernel writes straight into the flat
        # gradient buffer [table grad | MLP grads]
        n = self.nerf
        enc = encoders.hashgrid_forward(self.levels, xyzs, synthetic.BOUND, self.table)
        if self.fused_mlp:
            drgbs = nerf_mod.mlp_forward(enc, dirs, self.mlp_flat)
            drgbs_leaf = drgbs
        else:
            enc.requires_grad_(True)
            x = torch.relu(enc @ n.density_w0) @ n.density_w1
            density = nerf_mod.trunc_exp(x[:, :1])
            h = torch.cat([x, nerf_mod.sh4(dirs)], dim=-1)
            rgb = torch.sigmoid(torch.relu(torch.relu(h @ n.rgb_w0) @ n.rgb_w1) @ n.rgb_w2)
            drgbs = torch.cat([density, rgb], dim=-1)
            drgbs_leaf = drgbs.detach()
        drgbs_leaf = drgbs_leaf.requires_grad_(True)
        effective, final_rgbds, _ = integrate_rays(synthetic.NEAR, rays_start, rays_n, bg, dss, z_vals, drgbs_leaf)
        gt = sc.rgbas_u8[perm].to(torch.float32) / 255  # _utils.py:165
        gt_rgb = gt[:, :3] * gt[:, 3:] + bg * (1 - gt[:, 3:])  # utils/data.py:443-464
        n_valid = ray_is_valid.sum()
        loss = torch.where(ray_is_valid, huber(final_rgbds[:, :3], gt_rgb).mean(-1), 0.0).sum() / n_valid  # :151-156
        (d_drgbs,) = torch.autograd.grad(loss, [drgbs_leaf])
        if self.fused_mlp:
            d_enc, _ = nerf_mod.mlp_backward(enc, dirs, self.mlp_flat, d_drgbs.contiguous(), d_weights=self.mlp_grad)
        else:
            d_enc, d_w = torch.autograd.grad(drgbs, [enc, n.mlp_flat], d_drgbs)
            self.mlp_grad.copy_(d_w)
        encoders.hashgrid_backward(self.levels, xyzs, synthetic.BOUND, d_enc.contiguous(), out=self.table_grad)
        out = dict(loss=loss.detach(), n_valid_rays=n_valid, measured_batch_size_before_compaction=mb,
                   measured_batch_size=effective)
        if apply and self.world_size == 1:
            self._optimizer_step()
        return out

    def _optimizer_step(self):
        """Gradient exchange + Adam.  With more than one rank this part stays outside the CUDA graph: the
        collectives are issued eagerly on the current stream after the captured compute graph."""
        lo, hi = self.shard_lo, self.shard_hi
        if self.peer_exchange is not None:  # the three steps below as one kernel over NVLink (csrc/exchange.cu)
            self.peer_exchange.step(self.step_dev, self.adam_m, self.adam_v, self.adam_desc, lo)
            trainops.u32_axpy(self.step_dev, self.step_dev, 1, 0, 1, out=self.step_dev)
            return
        # reduce-scatter [table grad | MLP grads] -> Adam on this rank's shard -> all-gather the parameters
        g = dp.reduce_scatter_flat_gradients(self.flat_grads, self.rank, self.world_size, self.pg)
        _lib.call("ngp_adam_step", [self.step_dev, self.flat_params[lo:hi], g, self.adam_m, self.adam_v], self.adam_desc)
        dp.all_gather_flat_parameters(self.flat_params, self.rank, self.world_size, self.pg)
        trainops.u32_axpy(self.step_dev, self.step_dev, 1, 0, 1, out=self.step_dev)

    def train_step(self, perm, next_perm=None):
        """perm: int32 [n_rays] indices into the scene's pixels (device or pinned host tensor).  Returns
        device-side metrics (no host synchronisation).

        ``next_perm`` (optional) is the batch of the FOLLOWING call.  Ray generation and march depend on the batch
        and the occupancy bitfield only, not on the parameters, so they are enqueued on a side stream right away and
        run underneath this whole step (forward, backward, gradient exchange, optimizer) instead of on its critical
        path.  The marched batches are double-buffered (two march graphs, two compute graphs over their own static
        buffers): the slot the side stream writes was last read by the PREVIOUS step's backward.  Same numbers as
        marching afterwards; a committed density-grid update in between drops the prefetch."""
        self.step += 1
        if not self.use_graph or not self.fused_glue:
            out = self._step_body(perm)
            if self.world_size > 1:
                self._optimizer_step()
            return out
        main = torch.cuda.current_stream(self.device)
        if self._graph is None:
            self._capture_graphs(perm, main)
        cur = self._slot
        if self._prefetched == self._batch_key(perm):
            main.wait_event(self._ev_march[cur])  # marched on the side stream during the previous step
        else:
            if self._prefetched is not None:  # a different batch was prefetched into this slot: let it finish first
                main.wait_event(self._ev_march[cur])
                self.rng_state[:1] -= 1  # its draws are handed back (see drop_prefetch)
            self._static_perm[cur].copy_(perm, non_blocking=True)
            self._march_graph[cur].replay()
        self._prefetched = None
        # One GPU: the prefetch goes out before the compute graph and runs underneath it.  Several GPUs: it goes out
        # after the compute graph and runs underneath the gradient exchange (reduce-scatter, sharded Adam, all-gather),
        # which leaves the SMs almost idle -- there it hides completely, whereas under the compute graph it would slow
        # the backward down and leave the exchange uncovered (measured at N=8: 0.73 vs 0.80 ms per step).
        if next_perm is not None and self.world_size == 1:
            self._prefetch(next_perm, cur, main)
        self._graph[cur].replay()
        if next_perm is not None and self.world_size > 1:
            self._prefetch(next_perm, cur, main)
        if self._exchange_graph is not None:
            self._exchange_graph.replay()
        else:
            self._optimizer_step()
        self._slot = 1 - cur
        return self._static_out[cur]

    def _prefetch(self, next_perm, cur, main):
        nxt = 1 - cur
        self._ev_free.record(main)  # everything that read slot `nxt` (the previous step's backward) precedes this point
        self._side.wait_event(self._ev_free)
        with torch.cuda.stream(self._side):
            self._static_perm[nxt].copy_(next_perm, non_blocking=True)
            self._march_graph[nxt].replay()
            self._ev_march[nxt].record(self._side)
        self._prefetched = self._batch_key(next_perm)

    @staticmethod
    def _batch_key(perm):
        """Identity of a batch for the prefetch hit test: address, length and torch's in-place version counter, so a
        staging buffer refilled in place (``copy_``, ``random_`` ...) is a different batch.  Writes torch cannot see
        (a numpy view of pinned memory) are not detected: do not refill a buffer that way while it is prefetched."""
        return (perm.data_ptr(), perm.numel(), perm._version)

    def _capture_graphs(self, perm, main):
        self._side = torch.cuda.Stream(device=self.device)
        self._static_perm = [torch.zeros(self.n_rays, dtype=torch.int32, device=self.device) for _ in range(2)]
        self._static_perm[0].copy_(perm, non_blocking=True)
        self._static_perm[1].copy_(perm, non_blocking=True)
        self._side.wait_stream(main)
        self._march_graph, self._graph, self._static_marched, self._static_out = [], [], [], []
        with torch.cuda.stream(self._side):
            # warm-up on the capture stream (scratch blocks, lazy module loads, NCCL channels).  It runs real steps, so
            # everything they change is put back afterwards: the first train_step applies ONE optimizer update from its
            # batch, as the eager path and the reference do (every rank issues the same warm-up collectives).
            saved = [t.clone() for t in (self.flat_params, self.adam_m, self.adam_v, self.step_dev, self.rng_state)]
            for _ in range(2):
                self._compute_body(self._static_perm[0], self._march_body(self._static_perm[0]))
                self._optimizer_step()
            for t, v in zip((self.flat_params, self.adam_m, self.adam_v, self.step_dev, self.rng_state), saved):
                t.copy_(v)
            for slot in range(2):
                mg = torch.cuda.CUDAGraph()
                # the captured march runs underneath the step's other kernels: a small persistent grid keeps it from
                # crowding them out of the SMs (its tiles are taken by ticket, so any grid size does all the work)
                _lib.lib().ngp_b200_set_march_ctas_per_sm(self.prefetch_march_ctas_per_sm)
                try:
                    with torch.cuda.graph(mg, stream=self._side):
                        marched = self._march_body(self._static_perm[slot])
                finally:
                    _lib.lib().ngp_b200_set_march_ctas_per_sm(0)
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg, stream=self._side):
                    out = self._compute_body(self._static_perm[slot], marched)
                self._march_graph.append(mg)
                self._graph.append(cg)
                self._static_marched.append(marched)
                self._static_out.append(out)
            if self.graph_exchange and self.peer_exchange is not None:  # every rank captures the same launch
                eg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(eg, stream=self._side):
                    self._optimizer_step()
                self._exchange_graph = eg
        main.wait_stream(self._side)
        self._ev_march = [torch.cuda.Event(), torch.cuda.Event()]
        self._ev_free = torch.cuda.Event()
        self._prefetched = None
        self._slot = 0

    def drop_prefetch(self):
        """Forget a march prefetched on the side stream (its inputs are about to change): the main stream first waits
        for it, so that nothing still writes the slot's buffers when the next ``train_step`` marches into them."""
        if self._prefetched is not None:
            torch.cuda.current_stream(self.device).wait_event(self._ev_march[self._slot])
            self._prefetched = None
            self.rng_state[:1] -= 1  # the dropped march consumed a call of the random stream: hand it back

    def release_graph(self):
        """Drop the captured step graphs (and their private memory pools)."""
        self._graph = self._march_graph = self._exchange_graph = None
        self._static_out = self._static_marched = None
        self._prefetched = None

    def mark_untrained_density_grid(self):
        """utils/types.py:1241-1362, called by the reference once before training and after a checkpoint load
        (app/nerf/train.py:206): cells no training camera sees are culled for good.  Not part of ``__init__``: the
        procedural scene's cameras see every cell of its single cascade, and bench.py installs the analytic occupancy."""
        self.drop_prefetch()
        return ogrid.mark_untrained_density_grid(self.grid, self.scene.transforms, self.scene.cam, synthetic.BOUND,
                                                 synthetic.DIAGONAL_N_STEPS, self.step)

    # -- density grid update (utils/types.py:1149-1239) --------------------------------------------
    def _density_fn(self, xyz, out=None):
        if self.fused_encoder:
            return nerf_mod.fused_forward(self.levels, xyz.contiguous(), synthetic.BOUND, self.table, None, self.mlp_flat, out=out)
        enc = encoders.hashgrid_forward(self.levels, xyz.contiguous(), synthetic.BOUND, self.table)
        if self.fused_mlp:
            return nerf_mod.mlp_forward(enc, None, self.mlp_flat)
        x = torch.relu(enc @ self.nerf.density_w0) @ self.nerf.density_w1
        return torch.exp(x[:, 0])

    @torch.no_grad()
    def update_ogrid(self, update_all=None, commit=True):
        """train.py:75-85: update every cascade's densities, then re-threshold and re-pack the bitfield.
        ``commit=False`` does all the work into shadow buffers (bench: keeps the marching workload fixed)."""
        if update_all is None:
            update_all = self.step < 256  # utils/types.py:1391-1392
        if commit:  # a prefetched march saw the old bitfield: redo it (with the same random draws)
            self.drop_prefetch()
        g = self.grid
        if commit:
            shadow = mask_out = bits_out = None
        else:
            if self._shadow is None:
                self._shadow = (torch.empty_like(g.density), torch.empty_like(g.occ_mask), torch.empty_like(g.occupancy))
            shadow, mask_out, bits_out = self._shadow
        for cas in range(g.K):
            ogrid.update_ogrid_density(g, self._density_fn, cas, update_all, synthetic.BOUND, self.total_samples,
                                       out_density=shadow)
        if self.world_size > 1:  # every rank sampled its own cells: take the max
            dp.allreduce_density_grid(g.density if commit else shadow, self.pg)
        _, occ_mask, occupancy = ogrid.threshold_ogrid_(g, synthetic.DIAGONAL_N_STEPS, synthetic.BOUND, density=shadow,
                                                        occ_mask=mask_out, occupancy=bits_out)
        return occ_mask, occupancy
