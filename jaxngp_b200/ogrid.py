"""Occupancy / density grid state and its update -- host mirror of ``OccupancyDensityGrid``
(utils/types.py:93-144) and ``NeRFState.update_ogrid_density`` / ``threshold_ogrid``
(utils/types.py:1149-1239).  Every array step runs in this package's kernels (csrc/ogrid.cu,
gridops.cu); the random draws are inputs, as the reference draws them with jax.random outside the ops.
"""

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
    selections (parity tests); cells are assumed alive (no camera culling in synthetic scenes)."""
    G3, dev = grid.G3, grid.density.device
    sl = slice(cas * G3, (cas + 1) * G3)
    if update_all:  # :1166-1169
        idx = torch.arange(G3, dtype=torch.int32, device=dev)
    else:  # :1170-1191
        M = max(1, G3 // 2)
        half = max(1, M // 2)
        if draws is not None:
            first, second = draws["first"], draws["second"]
        else:
            first = torch.randint(0, G3, (half,), device=dev, generator=generator, dtype=torch.int32)
            # uniform over the currently occupied cells: inverse-CDF over the mask (jran.choice with p)
            csum = torch.cumsum(grid.occ_mask[sl].to(torch.int32), 0)
            total = csum[-1]
            u = torch.rand(half, device=dev, generator=generator)
            target = (u * total.to(torch.float32)).to(torch.int32).clamp(max=(total - 1).clamp(min=0)) + 1
            second = torch.searchsorted(csum, target).clamp(max=G3 - 1).to(torch.int32)
        idx = torch.cat([first, second])
    jitter = draws["jitter"] if draws is not None else torch.rand(idx.shape[0], 3, device=dev, generator=generator)
    coords = sample_positions(idx, jitter, grid.G, cas, bound)
    new_density = torch.cat([density_fn(part).reshape(-1) for part in coords.split(max(1, max_inference))])
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
