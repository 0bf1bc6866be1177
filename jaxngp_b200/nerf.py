"""NeRF model of the reference's ``make_nerf_ngp`` (models/nerfs.py:27-128,216-238,422-454):

    hash-grid(32) -> Dense 64 -> ReLU -> Dense 16 (no bias) ; density = trunc_exp(x[0])
    [x(16) | SH deg 4 (16)] -> Dense 64 -> ReLU -> Dense 64 -> ReLU -> Dense 3 -> sigmoid

The hash-grid encoder runs on this package's CUDA kernels.  The dense layers here are the plain
library-GEMM arm (torch matmul, TF32 tensor cores); SURVEY 8(f1) ranks the fused tensor-core MLP as
the next component after the section-8 rows.
"""
import math

import torch

from . import encoders


def sh4(d: torch.Tensor) -> torch.Tensor:
    """Real spherical harmonics up to degree 4 (16 coefficients) of unit vectors; same basis, order
    and signs as models/encoders.py:365-406."""
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    xy, xz, yz = x * y, x * z, y * z
    x2, y2, z2 = x * x, y * y, z * z
    return torch.stack([
        torch.full_like(x, 0.28209479177387814),
        -0.48860251190291987 * y,
        0.48860251190291987 * z,
        -0.48860251190291987 * x,
        1.0925484305920792 * xy,
        -1.0925484305920792 * yz,
        0.94617469575755997 * z2 - 0.31539156525251999,
        -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2,
        0.59004358992664352 * y * (-3.0 * x2 + y2),
        2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2),
        0.3731763325901154 * z * (5.0 * z2 - 3.0),
        0.45704579946446572 * x * (1.0 - 5.0 * z2),
        1.4453057213202769 * z * (x2 - y2),
        0.59004358992664352 * x * (-x2 + 3.0 * y2),
    ], dim=-1)


class _TruncExp(torch.autograd.Function):
    """models/nerfs.py:222-238: exp forward, gradient uses exp(clip(x, -15, 15))."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return torch.exp(torch.clamp(x, -15, 15)) * g


trunc_exp = _TruncExp.apply


def glorot_uniform_(w: torch.Tensor, generator=None):
    fan_in, fan_out = w.shape
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return w.uniform_(-lim, lim, generator=generator)


class NeRF(torch.nn.Module):
    """``nerf(xyz, dir, appearance_embeddings) -> (drgbs[..., 4], tv)``; ``dir=None`` returns densities
    only (models/nerfs.py:40-86).  Weights are stored [in, out] like flax Dense kernels."""

    def __init__(self, bound: float, inference: bool = False, tv_scale: float = 0.0, device=None, generator=None,
                 T: int = 2 ** 19):
        super().__init__()
        self.bound = float(bound)
        Enc = encoders.TCNNHashGridEncoder if inference else encoders.HashGridEncoder  # nerfs.py:431-438
        self.position_encoder = Enc(L=16, T=T, F=2, N_min=2 ** 4, N_max=int(2 ** 11 * bound), tv_scale=tv_scale,
                                    device=device, generator=generator)

        def dense(i, o):
            return torch.nn.Parameter(glorot_uniform_(torch.empty(i, o, dtype=torch.float32, device=device), generator))

        self.density_w0, self.density_w1 = dense(32, 64), dense(64, 16)
        self.rgb_w0, self.rgb_w1, self.rgb_w2 = dense(32, 64), dense(64, 64), dense(64, 3)

    def mlp_parameters(self):
        return [self.density_w0, self.density_w1, self.rgb_w0, self.rgb_w1, self.rgb_w2]

    def forward(self, xyz, dir=None, appearance_embeddings=None):
        shape = xyz.shape[:-1]
        xyz = xyz.reshape(-1, 3)
        pos_enc, tv = self.position_encoder(xyz, self.bound)
        x = torch.relu(pos_enc @ self.density_w0) @ self.density_w1
        density = trunc_exp(x[:, :1])
        if dir is None:
            return density.reshape(*shape, 1), tv
        h = torch.cat([x, sh4(dir.reshape(-1, 3))], dim=-1)
        rgb = torch.sigmoid(torch.relu(torch.relu(h @ self.rgb_w0) @ self.rgb_w1) @ self.rgb_w2)
        return torch.cat([density, rgb], dim=-1).reshape(*shape, 4), tv
