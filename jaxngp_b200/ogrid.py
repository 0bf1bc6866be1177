"""Occupancy / density grid state and its update -- host mirror of ``OccupancyDensityGrid``
(utils/types.py:93-144) and ``NeRFState.update_ogrid_density`` / ``threshold_ogrid``
(utils/types.py:1149-1239).  Every array step runs in this package's kernels (csrc/ogrid.cu,
gridops.cu); the random draws are inputs, as the reference draws them with jax.random outside the ops.
"""

import inspect

import torch

from . import _lib, descriptors
from .volrendjax import packbits


class OccupancyDensityGrid:
    """density f32[K*G^3] (-1 = untrainable), occ_mask bool[K*G^3], occupancy u8[K*G^3/8] (all ones at
    creation, utils/types.py:121-141)."""

    def __init__(self, cascades: int, grid_resolution: int = 128, device=None):
        self.K, self.G = cascades, grid_resolution
        n = cascades * grid_resolution ** 3
        self.density = torch.zeros(n, dtype=torch.float32, device=device)
        self.occ_mask = torch.zeros(n, dtype=torch.bool, device=device)
        self.occupancy = torch.full((n // 8,), 255, dtype=torch.uint8, device=device)
        # cells some training camera sees (mark_untrained_density_grid); None = every cell, as at creation (types.py:139)
        self.alive_indices = None      # int32 [n_alive], global Morton indices, cascade by cascade
        self.alive_indices_offset = None  # python list, K + 1 entries (types.py:1353-1358)
        # random draws of the update (cell choice, jitter): Philox stream of csrc/common.cuh, counter on the device
        self.seed = 0
        self.rng_state = torch.zeros(2, dtype=torch.int32, device=device)

    def alive_in_cascade(self, cas: int):
        """Aligned (in-cascade) indices of the trainable cells of one cascade, or None when all of them are."""
        if self.alive_indices is None:
            return None
        lo, hi = self.alive_indices_offset[cas], self.alive_indices_offset[cas + 1]
        return self.alive_indices[lo:hi] % self.G3

    @property
    def G3(self):
        return self.G ** 3


def density_threshold_from_min_step_size(diagonal_n_steps: int, bound: float) -> float:
    """utils/types.py:1367-1369."""
    return 0.01 * diagonal_n_steps / (2 * min(bound, 1) * 3 ** 0.5)


def sample_positions(idx: torch.Tensor, uniforms: torch.Tensor, G: int, cas: int, bound: float) -> torch.Tensor:
    """utils/types.py:1193-1206: a random point inside each selected cell.  idx: int32 Morton indices in
    [0, G^3); uniforms: f32 [M, 3] draws in [0, 1)."""
    m = idx.shape[0]
    coords = torch.empty(m, 3, dtype=torch.float32, device=idx.device)
    if m:
        _lib.call("ngp_ogrid_sample_positions", [idx.contiguous(), uniforms.contiguous(), coords],
                  descriptors.make_ogrid_sample_descriptor(m, G, min(bound, 2.0 ** cas)))
    return coords


STREAM_OGRID = 2  # NgpRngDescriptor.stream_id of the grid update's draws


def draw_cells(grid: OccupancyDensityGrid, cas: int, update_all: bool, bound: float):
    """utils/types.py:1166-1206 in one op (csrc/ogrid.cu ``ngp_ogrid_draw_cells``): which cells of cascade ``cas`` this
    update evaluates, and a jittered point inside each.  Returns (idx int32 [M] Morton indices inside the cascade,
    coords f32 [M, 3]).  Draws come from ``grid.rng_state`` / ``grid.seed``; the launch advances the counter."""
    G3, dev = grid.G3, grid.density.device
    alive = grid.alive_in_cascade(cas)
    n_alive = G3 if alive is None else int(alive.shape[0])
    if update_all:
        n_first, n_second, m = 0, 0, n_alive
    else:
        n_first = n_second = max(1, max(1, n_alive // 2) // 2)  # M = max(1, n_grids // 2); max(1, M // 2) each (:1171-1190)
        m = n_first + n_second
    idx = torch.empty(m, dtype=torch.int32, device=dev)
    coords = torch.empty(m, 3, dtype=torch.float32, device=dev)
    bits = grid.occupancy[cas * G3 // 8:(cas + 1) * G3 // 8]
    _lib.call("ngp_ogrid_draw_cells", [bits, alive.contiguous() if alive is not None else 0, grid.rng_state, idx, coords],
              descriptors.make_ogrid_draw_descriptor(G3, grid.G, n_alive, alive is not None, update_all, n_first, n_second,
                                                     min(bound, 2.0 ** cas), grid.seed, STREAM_OGRID))
    return idx, coords


def decay_and_max(density: torch.Tensor, idx: torch.Tensor, new_density: torch.Tensor, decay: float = 0.95,
                  out: torch.Tensor = None) -> torch.Tensor:
    """utils/types.py:1162-1164,1219-1221 on one cascade's slice: alive cells decay, then
    ``density[idx] = max(density[idx], new_density)`` (atomic max)."""
    if out is None:
        out = torch.empty_like(density)
    _lib.call("ngp_ogrid_decay_max", [density, idx.contiguous(), new_density.contiguous(), out],
              descriptors.make_ogrid_update_descriptor(density.shape[0], idx.shape[0], decay))
    return out


def threshold(density_cascade0: torch.Tensor, thr_max: float) -> torch.Tensor:
    """utils/types.py:1229-1230: min(thr_max, mean over alive cells of cascade 0), a device scalar [1]."""
    thr = torch.empty(1, dtype=torch.float32, device=density_cascade0.device)
    _lib.call("ngp_ogrid_threshold", [density_cascade0, thr],
              descriptors.make_ogrid_threshold_descriptor(density_cascade0.shape[0], thr_max))
    return thr


def update_ogrid_density(grid: OccupancyDensityGrid, density_fn, cas: int, update_all: bool, bound: float,
                         max_inference: int, draws=None, generator=None, out_density=None):
    """``NeRFState.update_ogrid_density`` (utils/types.py:1149-1225) for one cascade.  ``density_fn(xyz)``
    evaluates the NeRF's density branch.  ``draws`` = dict(first, second, jitter) overrides the random
    selections (parity tests).  After ``mark_untrained_density_grid`` culled cells, only trainable cells are sampled.
    ``density_fn(xyz, out)`` may write the densities of ``xyz`` into ``out`` and return it (no concatenation afterwards)."""
    G3, dev = grid.G3, grid.density.device
    sl = slice(cas * G3, (cas + 1) * G3)
    if draws is None and generator is None:  # the product path: cell choice, jitter and positions in one op
        idx, coords = draw_cells(grid, cas, update_all, bound)
    else:  # supplied draws (parity tests against the reference's own update) or a torch generator
        alive = grid.alive_in_cascade(cas)  # None unless mark_untrained_density_grid culled cells (:1158-1160)
        n_alive = G3 if alive is None else int(alive.shape[0])
        if update_all:  # :1166-1169
            idx = torch.arange(G3, dtype=torch.int32, device=dev) if alive is None else alive
        else:  # :1170-1191
            M = max(1, n_alive // 2)
            half = max(1, M // 2)
            if draws is not None:
                first, second = draws["first"], draws["second"]
            else:
                first = torch.randint(0, n_alive, (half,), device=dev, generator=generator, dtype=torch.int32)
                if alive is not None:  # uniform among the trainable cells
                    first = alive[first.long()]
                # uniform over the currently occupied cells: inverse-CDF over the mask (jran.choice with p)
                csum = torch.cumsum(grid.occ_mask[sl].to(torch.int32), 0)
                total = csum[-1]
                u = torch.rand(half, device=dev, generator=generator)
                target = (u * total.to(torch.float32)).to(torch.int32).clamp(max=(total - 1).clamp(min=0)) + 1
                second = torch.searchsorted(csum, target).clamp(max=G3 - 1).to(torch.int32)
            idx = torch.cat([first, second])
        jitter = draws["jitter"] if draws is not None else torch.rand(idx.shape[0], 3, device=dev, generator=generator)
        coords = sample_positions(idx, jitter, grid.G, cas, bound)
    # :1208-1217: the points go through the model in chunks of `max_inference`; every chunk writes its slice of one buffer
    new_density = torch.empty(coords.shape[0], dtype=torch.float32, device=dev)
    chunk = max(1, max_inference)
    takes_out = len(inspect.signature(density_fn).parameters) >= 2
    for begin in range(0, coords.shape[0], chunk):
        part = new_density[begin:begin + chunk]
        res = density_fn(coords[begin:begin + chunk], part) if takes_out else density_fn(coords[begin:begin + chunk])
        if res is not part:  # a density function that ignores the output slice
            part.copy_(res.reshape(-1))
    target = grid.density[sl] if out_density is None else out_density[sl]
    decay_and_max(grid.density[sl], idx, new_density, 0.95, out=target)
    return idx, coords, new_density


def threshold_ogrid(grid: OccupancyDensityGrid, diagonal_n_steps: int, bound: float, density=None, commit=True):
    """``NeRFState.threshold_ogrid`` (utils/types.py:1227-1239)."""
    density = grid.density if density is None else density
    thr = threshold(density[: grid.G3], density_threshold_from_min_step_size(diagonal_n_steps, bound))
    occ_mask, occupancy = packbits(thr, density)
    if commit:  # in place: captured CUDA graphs keep reading these buffers
        grid.occ_mask.copy_(occ_mask)
        grid.occupancy.copy_(occupancy)
    return thr, occ_mask, occupancy


def threshold_ogrid_(grid: OccupancyDensityGrid, diagonal_n_steps: int, bound: float, density=None, occ_mask=None, occupancy=None):
    """``threshold_ogrid`` writing the mask and the bitfield straight into the given buffers (default: the grid's own,
    which captured CUDA graphs keep reading): no temporaries, no copies."""
    density = grid.density if density is None else density
    occ_mask = grid.occ_mask if occ_mask is None else occ_mask
    occupancy = grid.occupancy if occupancy is None else occupancy
    thr = threshold(density[: grid.G3], density_threshold_from_min_step_size(diagonal_n_steps, bound))
    _lib.call("ngp_packbits_scalar", [thr, density, occ_mask, occupancy], descriptors.make_packbits_descriptor(density.shape[0] // 8))
    return thr, occ_mask, occupancy


_CELL_CORNERS = ((0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1))


def _morton3d_invert_host(idx: torch.Tensor) -> torch.Tensor:
    """Morton index -> (x, y, z), torch integer ops on any device (marching.cu:70-77's bit compaction).  Used by the
    one-time culling below so that it does not depend on a stream or a launch."""
    def compact(x):
        x = x & 0x49249249
        x = (x | (x >> 2)) & 0xC30C30C3
        x = (x | (x >> 4)) & 0x0F00F00F
        x = (x | (x >> 8)) & 0xFF0000FF
        x = (x | (x >> 16)) & 0x0000FFFF
        return x
    idx = idx.to(torch.int64)
    return torch.stack([compact(idx), compact(idx >> 1), compact(idx >> 2)], dim=-1)


def visible_cells(K: int, G: int, bound: float, transforms: torch.Tensor, cam: dict) -> torch.Tensor:
    """``NeRFState.mark_untrained_density_grid`` (utils/types.py:1241-1326), undistorted cameras: bool [K * G^3], true
    where one of the cell's 8 corners is in front of some training camera and projects into its frame.
    ``transforms`` [V, 12]: rot_cw row-major then t_cw (trainer.Scene).  One-time setup, a few torch ops per view."""
    if cam.get("distortion") or cam.get("has_distortion"):
        raise NotImplementedError("camera distortion models are outside this path (utils/types.py:1286-1297)")
    dev = transforms.device
    G3 = G ** 3
    cells = _morton3d_invert_host(torch.arange(G3, device=dev)).to(torch.float32)
    corners = torch.tensor(_CELL_CORNERS, dtype=torch.float32, device=dev)
    f = torch.tensor([cam["fx"], cam["fy"]], dtype=torch.float32, device=dev)
    c = torch.tensor([cam["cx"], cam["cy"]], dtype=torch.float32, device=dev)
    wh = torch.tensor([cam["width"], cam["height"]], dtype=torch.float32, device=dev)
    alive = torch.zeros(K * G3, dtype=torch.bool, device=dev)
    for cas in range(K):
        mip_bound = float(min(2 ** cas, bound))
        xyz = (cells / G - 0.5) * (2 * mip_bound)                       # :1249-1252
        verts = xyz[:, None, :] + (2 * mip_bound / G) * corners[None]   # :1253-1263
        part = alive[cas * G3:(cas + 1) * G3]
        for tf in transforms.to(torch.float32):
            rot, t = tf[:9].reshape(3, 3), tf[9:]
            p_cam = ((verts - t)[..., None, :] * rot.T).sum(-1)          # world -> camera (:1275-1276)
            front = p_cam[..., 2] < 0                                    # the camera looks along -z
            uv = ((p_cam[..., :2] / (-p_cam[..., 2:])) * f + c) / wh     # :1281,1298-1302
            inside = ((uv >= 0) & (uv < 1)).all(-1)
            part |= (front & inside).any(-1)                             # :1310-1312
            if bool(part.all()):
                break
    return alive


def mark_untrained_density_grid(grid: OccupancyDensityGrid, transforms: torch.Tensor, cam: dict, bound: float,
                                diagonal_n_steps: int, step: int) -> torch.Tensor:
    """utils/types.py:1241-1362: cull the cells no training camera sees (density -1, never sampled or decayed again),
    re-threshold (-0.5 at step 0: every trainable cell starts occupied) and re-pack the bitfield, in place.
    Returns the alive marker.  Called once at start / after loading a checkpoint (app/nerf/train.py:206)."""
    alive = visible_cells(grid.K, grid.G, bound, transforms, cam)
    if step > 0:  # :1343 reads the grid BEFORE the culling: mean over the cells that were trainable until now
        thr = threshold(grid.density[: grid.G3], density_threshold_from_min_step_size(diagonal_n_steps, bound))
    else:
        thr = -0.5
    grid.density.copy_(torch.where(alive, grid.density, torch.full_like(grid.density, -1.0)))
    occ_mask, occupancy = packbits(thr, grid.density)
    grid.occ_mask.copy_(occ_mask)
    grid.occupancy.copy_(occupancy)
    if bool(alive.all()):
        grid.alive_indices = grid.alive_indices_offset = None
    else:
        grid.alive_indices = torch.nonzero(alive).reshape(-1).to(torch.int32)
        per_cascade = alive.view(grid.K, grid.G3).sum(dim=1).tolist()
        grid.alive_indices_offset = [0]
        for n in per_cascade:
            grid.alive_indices_offset.append(grid.alive_indices_offset[-1] + int(n))
    return alive
