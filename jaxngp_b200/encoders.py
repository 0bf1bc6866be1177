"""HashGridEncoder / TCNNHashGridEncoder -- torch modules with the reference's flax module signature
(models/encoders.py:58-305): fields ``(L, T, F, N_min, N_max, tv_scale, param_dtype)``, call
``encoder(pos[n, dim], bound) -> (encodings[n, L*F], tv)``, one parameter
``"latent codes stored on grid vertices"`` of shape ``[rows, F]`` initialised U(-1e-4, 1e-4).

Forward and backward each run as one hand-written sm_100a kernel (csrc/hashgrid.cu); there is no
torch/eager fallback.
"""
import math
from typing import NamedTuple

import torch

from . import _lib, descriptors

PARAM_NAME = "latent codes stored on grid vertices"  # models/encoders.py:106


def next_multiple(value: int, multiple: int) -> int:  # utils/common.py:336-337
    return ((value + multiple - 1) // multiple) * multiple


class LevelTable(NamedTuple):
    L: int
    T: int
    F: int
    dim: int
    b: float
    scales: tuple   # python floats already rounded to f32
    res: tuple
    hashed: tuple
    offsets: tuple  # L + 1 entries

    @property
    def rows(self) -> int:
        return self.offsets[-1]


def _f32(x: float) -> float:
    return torch.tensor(x, dtype=torch.float64).to(torch.float32).item()


def make_level_table(L: int, T: int, F: int, N_min: int, N_max: int, dim: int, align: int = 8) -> LevelTable:
    """Level geometry exactly as models/encoders.py:76-80,89-103 computes it: double precision on the
    host, scale cast to f32 at :216.  ``align=8`` is HashGridEncoder (:96), ``align=1`` is
    TCNNHashGridEncoder (:275)."""
    b = math.exp((math.log(N_max) - math.log(N_min)) / (L - 1))
    scales, res, hashed, offsets = [], [], [], [0]
    for i in range(L):
        scale = N_min * (b ** i) - 1
        scales.append(_f32(scale))
        r = math.ceil(scale) + 1
        res.append(r)
        n_entries = next_multiple(r ** dim, align)
        if n_entries <= T:
            hashed.append(False)
        else:
            n_entries = T
            hashed.append(True)
        offsets.append(offsets[-1] + n_entries)
    return LevelTable(L, T, F, dim, b, tuple(scales), tuple(res), tuple(hashed), tuple(offsets))


def _a1_descriptor(lt: LevelTable, n_points: int, bound: float, wrap: str, table_dtype: torch.dtype,
                   rows_per_group: int = 0) -> bytes:
    return descriptors.make_hashgrid_a1_descriptor(
        n_points=n_points, dim=lt.dim, L=lt.L, F=lt.F, wrap_T=lt.T if wrap == "jaxngp" else 0,
        table_dtype={torch.float32: 0, torch.float16: 1}[table_dtype], bound=bound, hashed=lt.hashed,
        scales=lt.scales, res=lt.res, offsets=lt.offsets, rows_per_group=rows_per_group)


def hashgrid_forward(lt: LevelTable, pos: torch.Tensor, bound: float, table: torch.Tensor, wrap: str = "jaxngp",
                     group_counts: torch.Tensor = None, rows_per_group: int = 0):
    """enc[n, L*F] = HashGridEncoder gather (models/encoders.py:216-233); no autograd.  With
    ``group_counts`` (int32 [n / rows_per_group]) only the first ``group_counts[g]`` rows of each group of
    ``rows_per_group`` rows are encoded (the [n_rays, cap] sample layout of march_rays_inference); the
    padding rows are left unwritten."""
    n = pos.shape[0]
    enc = torch.empty(n, lt.L * lt.F, dtype=torch.float32, device=pos.device)
    if n:
        if group_counts is None:
            _lib.call("ngp_hashgrid_a1_forward", [pos, table, enc], _a1_descriptor(lt, n, bound, wrap, table.dtype))
        else:
            _lib.call("ngp_hashgrid_a1_forward", [pos, table, group_counts, enc],
                      _a1_descriptor(lt, n, bound, wrap, table.dtype, rows_per_group))
    return enc


def hashgrid_backward(lt: LevelTable, pos: torch.Tensor, bound: float, d_enc: torch.Tensor, wrap: str = "jaxngp",
                      out: torch.Tensor = None, accumulate: bool = False):
    """d_table[rows, F] = scatter-add of w_c * d_enc (the autodiff of the gather at encoders.py:226-231)."""
    n = pos.shape[0]
    if out is None:
        out = torch.empty(lt.rows, lt.F, dtype=torch.float32, device=pos.device)
    _lib.call("ngp_hashgrid_a1_backward_acc" if accumulate else "ngp_hashgrid_a1_backward", [pos, d_enc, out],
              _a1_descriptor(lt, n, bound, wrap, torch.float32))
    return out


class _HashGridFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, table, lt, bound, wrap):
        pos = pos.contiguous()
        ctx.lt, ctx.bound, ctx.wrap = lt, bound, wrap
        ctx.save_for_backward(pos)
        return hashgrid_forward(lt, pos, bound, table.contiguous(), wrap)

    @staticmethod
    def backward(ctx, d_enc):
        (pos,) = ctx.saved_tensors
        # positions are never differentiated on this path: march outputs are not
        # functions of the parameters
        return None, hashgrid_backward(ctx.lt, pos, ctx.bound, d_enc.contiguous(), ctx.wrap), None, None, None


class HashGridEncoder(torch.nn.Module):
    """models/encoders.py:58-256.  ``dim`` (2 or 3) fixes the table geometry at construction, which
    flax defers to the first call."""

    wrap = "jaxngp"  # `indices mod T` on every level, encoders.py:187
    align = 8        # encoders.py:96

    def __init__(self, L: int, T: int, F: int, N_min: int, N_max: int, tv_scale: float = 0.0,
                 param_dtype: torch.dtype = torch.float32, dim: int = 3, device=None, generator=None):
        super().__init__()
        if param_dtype != torch.float32:
            raise NotImplementedError("the hash table parameter is float32 (models/encoders.py:83, nerfs.py:332)")
        self.L, self.T, self.F, self.N_min, self.N_max, self.tv_scale, self.dim = L, T, F, N_min, N_max, tv_scale, dim
        self.levels = make_level_table(L, T, F, N_min, N_max, dim, self.align)
        latents = torch.empty(self.levels.rows, F, dtype=torch.float32, device=device)
        latents.uniform_(-1e-4, 1e-4, generator=generator)  # encoders.py:111
        self.latents = torch.nn.Parameter(latents)

    @property
    def b(self) -> float:  # encoders.py:76-80
        return self.levels.b

    def named_reference_params(self):
        return {PARAM_NAME: self.latents}

    def forward(self, pos: torch.Tensor, bound: float):
        if pos.shape[-1] != self.dim:
            raise NotImplementedError(
                "{} was built for {}-D inputs, got {}-D".format(type(self).__name__, self.dim, pos.shape[-1]))
        enc = _HashGridFn.apply(pos, self.latents, self.levels, float(bound), self.wrap)
        if self.tv_scale > 0:
            tv = self.tv_scale * self._total_variation(pos, float(bound))
        else:
            tv = 0
        return enc, tv

    # -- total variation regulariser (encoders.py:236-254); off by default (tv_scale = 0).  Built from
    #    torch index ops on the device: it is not on the hot path.
    def _indices(self, vert):  # vert: int64 [L, n, B, dim] holding uint32 values
        lt = self.levels
        M = 0xFFFFFFFF
        res = torch.tensor(lt.res, dtype=torch.int64, device=vert.device)[:, None, None]
        hashed = torch.tensor(lt.hashed, dtype=torch.bool, device=vert.device)[:, None, None]
        offs = torch.tensor(lt.offsets[:-1], dtype=torch.int64, device=vert.device)[:, None, None]
        x, y = vert[..., 0], vert[..., 1]
        if lt.dim == 3:
            z = vert[..., 2]
            dense = (x + ((y * res) & M) + ((z * ((res * res) & M)) & M)) & M
            hsh = x ^ ((y * 2654435761) & M) ^ ((z * 805459861) & M)
        else:
            dense = (x + ((y * res) & M)) & M
            hsh = x ^ ((y * 2654435761) & M)
        idx = torch.where(hashed, hsh, dense)
        wrap = lt.T if self.wrap == "jaxngp" else (torch.tensor(lt.offsets[1:], device=vert.device) - offs[:, 0, 0])[:, None, None]
        return idx % wrap + offs

    def _total_variation(self, pos, bound):
        lt = self.levels
        dim = lt.dim
        scales = torch.tensor(lt.scales, dtype=torch.float32, device=pos.device)
        p01 = (pos + bound) / (2 * bound)
        ps = p01[None] * scales[:, None, None] + 0.5
        fl = torch.floor(ps).to(torch.int64)
        eye = torch.eye(dim, dtype=torch.int64, device=pos.device)
        adj = torch.cat([eye.flip(0), -eye.flip(0)], 0)  # encoders.py:35-51
        first = (fl & 0xFFFFFFFF)[:, :, None, :]
        adjacent = ((fl[:, :, None, :] + adj[None, None]) & 0xFFFFFFFF)
        lat0 = self.latents[self._indices(first)]
        lat_adj = self.latents[self._indices(adjacent)]
        return torch.square(lat_adj - lat0).sum(dim=(-2, -1)).mean()


class TCNNHashGridEncoder(HashGridEncoder):
    """models/encoders.py:259-305: same encoding through the jax-tcnn custom call (3-D only, rows not
    8-aligned, tiny-cuda-nn indexing and f32-on-device level scale)."""

    wrap = "tcnn"
    align = 1  # encoders.py:275

    def forward(self, pos: torch.Tensor, bound: float):
        from .jaxtcnn import HashGridMetadata, hashgrid_encode

        pos = (pos + bound) / (2 * bound)  # encoders.py:265
        offsets = torch.tensor(self.levels.offsets, dtype=torch.int32, device=pos.device)
        enc_rm = hashgrid_encode(
            desc=HashGridMetadata(L=self.L, F=self.F, N_min=self.N_min, per_level_scale=self.b),
            offset_table_data=offsets, coords_rm=pos.T.contiguous(), params=self.latents)
        return enc_rm.T, 0
