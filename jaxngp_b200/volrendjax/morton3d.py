"""morton3d / morton3d_invert -- mirrors volrendjax/morton3d/__init__.py:6-10."""
import torch

from .. import _lib, descriptors


def _require_u32(t, op, what):
    if t.dtype not in (torch.int32, torch.uint32):  # morton3d/abstract.py:12-18,33-39
        raise NotImplementedError(
            "{} is only implemented for input {} of type `uint32` (carried as torch.int32), got {}".format(op, what, t.dtype))
    return t.view(torch.int32) if t.dtype == torch.uint32 else t


def morton3d(xyzs: torch.Tensor) -> torch.Tensor:
    xyzs = _require_u32(xyzs, "morton3d", "coordinates")
    length = xyzs.shape[0]
    idcs = torch.empty(length, dtype=torch.int32, device=xyzs.device)
    if length:
        _lib.call("ngp_morton3d", [xyzs.contiguous(), idcs], descriptors.make_morton3d_descriptor(length))
    return idcs


def morton3d_invert(idcs: torch.Tensor) -> torch.Tensor:
    idcs = _require_u32(idcs, "morton3d_invert", "indices")
    (length,) = idcs.shape
    xyzs = torch.empty(length, 3, dtype=torch.int32, device=idcs.device)
    if length:
        _lib.call("ngp_morton3d_invert", [idcs.contiguous(), xyzs], descriptors.make_morton3d_descriptor(length))
    return xyzs
