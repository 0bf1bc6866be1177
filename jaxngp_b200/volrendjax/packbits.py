"""packbits -- mirrors volrendjax/packbits/__init__.py:8-28."""
from typing import Tuple

import torch

from .. import _lib, descriptors
from ._check import require_f32


def packbits(density_threshold, density_grid: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns ``(occ_mask bool[N], occ_bitfield uint8[N//8])``; bit k of byte i is
    ``density_grid[8i+k] > density_threshold[8i+k]``.

    A scalar threshold (python float or 0-d/1-element tensor) is read by the kernel from device
    memory instead of being broadcast to an [N] array as the reference does
    (packbits/__init__.py:25-28); a full [N] threshold array takes the drop-in entry point.
    """
    # packbits/abstract.py:12-29
    if density_grid.dim() != 1:
        raise AssertionError(f"density_grid must have rank 1, got shape {tuple(density_grid.shape)}")
    n_bits = density_grid.shape[0]
    if n_bits % 8 != 0:
        raise ValueError(
            "pack_density_into_bits expects size of density grid to be divisible by 8, got {}".format(n_bits))
    require_f32(density_grid, "pack_density_into_bits", "densities")
    dev = density_grid.device
    thr = torch.as_tensor(density_threshold, dtype=torch.float32, device=dev)
    occ_mask = torch.empty(n_bits, dtype=torch.bool, device=dev)
    occ_bitfield = torch.empty(n_bits // 8, dtype=torch.uint8, device=dev)
    if n_bits == 0:
        return occ_mask, occ_bitfield
    desc = descriptors.make_packbits_descriptor(n_bits // 8)
    if thr.numel() == 1:
        _lib.call("ngp_packbits_scalar", [thr.reshape(1).contiguous(), density_grid.contiguous(), occ_mask, occ_bitfield], desc)
    else:
        thr = torch.broadcast_to(thr, density_grid.shape).contiguous()
        _lib.call("ngp_pack_density_into_bits", [thr, density_grid.contiguous(), occ_mask, occ_bitfield], desc)
    return occ_mask, occ_bitfield
