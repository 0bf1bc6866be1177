"""Host-side mirror of the reference's ``jaxtcnn`` package (deps/jax-tcnn/src/jaxtcnn/__init__.py:1-7,
hashgrid_tcnn/__init__.py:6-19, impl.py:20-137): ``HashGridMetadata`` and ``hashgrid_encode``."""
from dataclasses import dataclass

import torch

from .. import _lib, descriptors


@dataclass(frozen=True)
class HashGridMetadata:  # hashgrid_tcnn/impl.py:20-36
    L: int
    F: int
    N_min: int
    per_level_scale: float


def _validate(desc, offset_table_data, coords_rm, params):
    # hashgrid_tcnn/abstract.py:18-59
    dim, n_coords = coords_rm.shape
    if dim != 3:
        raise NotImplementedError(
            "hashgrid encoding is only implemented for 3D coordinates, expected input coordinates to have shape "
            "({}, n_coords), but got shape {}".format(dim, tuple(coords_rm.shape)))
    if tuple(offset_table_data.shape) != (desc.L + 1,):
        raise AssertionError(f"offset_table_data must have shape ({desc.L + 1},), got {tuple(offset_table_data.shape)}")
    if params.dim() != 2 or params.shape[1] != desc.F:
        raise AssertionError(f"params must have shape (n_params, {desc.F}), got {tuple(params.shape)}")
    if not (isinstance(desc.L, int) and isinstance(desc.F, int) and isinstance(desc.N_min, int)):
        raise AssertionError("L, F, N_min must be ints")
    if not isinstance(desc.per_level_scale, float):
        raise AssertionError("per_level_scale must be a float")
    if offset_table_data.dtype not in (torch.int32, torch.uint32):
        raise RuntimeError(
            "hashgrid encoding expects `offset_table_data` (a prefix sum of the hash table sizes of each level) to be "
            "of type uint32, got {}".format(offset_table_data.dtype))
    if coords_rm.dtype != torch.float32:
        raise NotImplementedError(
            "hashgrid encoding is only implemented for input coordinates of type float32, got {}".format(coords_rm.dtype))
    if params.dtype != torch.float32:
        raise NotImplementedError(
            "hashgrid encoding is only implemented for parameters of type float32, got {}".format(params.dtype))
    if desc.F not in (2, 4):  # deps/jax-tcnn/lib/impl/hashgrid.cu:81-85
        raise RuntimeError("supported values of F (n_features_per_level) are [2, 4], got {}".format(desc.F))
    return n_coords


class _HashGridEncodeFn(torch.autograd.Function):
    """custom_vjp of hashgrid_tcnn/impl.py:75-137."""

    @staticmethod
    def forward(ctx, desc, offset_table_data, coords_rm, params):
        n = coords_rm.shape[1]
        dev = coords_rm.device
        encoded_rm = torch.empty(desc.L * desc.F, n, dtype=torch.float32, device=dev)
        dy_dcoords_rm = torch.empty(3 * desc.L * desc.F, n, dtype=torch.float32, device=dev)
        opaque = descriptors.make_hashgrid_descriptor(n, desc.L, desc.F, desc.N_min, desc.per_level_scale)
        if n:
            _lib.call("ngp_hashgrid_encode", [offset_table_data, coords_rm, params, encoded_rm, dy_dcoords_rm], opaque)
        ctx.desc, ctx.opaque, ctx.n_params = desc, opaque, params.shape[0]
        ctx.save_for_backward(offset_table_data, coords_rm, dy_dcoords_rm)
        ctx.mark_non_differentiable(dy_dcoords_rm)
        return encoded_rm, dy_dcoords_rm

    @staticmethod
    def backward(ctx, dL_dy_rm, _unused):
        offset_table_data, coords_rm, dy_dcoords_rm = ctx.saved_tensors
        desc = ctx.desc
        dev = coords_rm.device
        dL_dparams = torch.empty(ctx.n_params, desc.F, dtype=torch.float32, device=dev)
        dL_dcoords_rm = torch.empty_like(coords_rm)
        _lib.call("ngp_hashgrid_encode_backward",
                  [offset_table_data, coords_rm, dL_dy_rm.contiguous(), dy_dcoords_rm, dL_dparams, dL_dcoords_rm],
                  ctx.opaque)
        return None, None, dL_dcoords_rm, dL_dparams


def hashgrid_encode(desc: HashGridMetadata, offset_table_data: torch.Tensor, coords_rm: torch.Tensor,
                    params: torch.Tensor) -> torch.Tensor:
    """``[L*F, n]`` encodings of ``coords_rm [3, n]`` (coordinates already in [0, 1]^3), same contract
    as hashgrid_tcnn/__init__.py:6-19."""
    _validate(desc, offset_table_data, coords_rm, params)
    if offset_table_data.dtype == torch.uint32:
        offset_table_data = offset_table_data.view(torch.int32)
    encoded_rm, _ = _HashGridEncodeFn.apply(desc, offset_table_data.contiguous(), coords_rm.contiguous(),
                                            params.contiguous())
    return encoded_rm


__all__ = ["HashGridMetadata", "hashgrid_encode"]
