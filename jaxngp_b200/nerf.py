"""NeRF model of the reference's ``make_nerf_ngp`` (models/nerfs.py:27-128,216-238,422-454):

    hash-grid(32) -> Dense 64 -> ReLU -> Dense 16 (no bias) ; density = trunc_exp(x[0])
    [x(16) | SH deg 4 (16)] -> Dense 64 -> ReLU -> Dense 64 -> ReLU -> Dense 3 -> sigmoid

The hash-grid encoder and the dense layers run on this package's CUDA kernels (csrc/hashgrid.cu,
csrc/mlp.cu: fully fused TF32 tensor-core MLP); a plain torch-matmul arm is kept as a
cross-check (``fused=False``).
"""
import math

import torch

from . import _lib, descriptors, encoders


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


MLP_SHAPES = (("density_w0", 32, 64), ("density_w1", 64, 16), ("rgb_w0", 32, 64), ("rgb_w1", 64, 64), ("rgb_w2", 64, 3))
MLP_NUMEL = sum(i * o for _, i, o in MLP_SHAPES)  # 9408


def mlp_forward(enc: torch.Tensor, dirs, weights: torch.Tensor, group_counts: torch.Tensor = None,
                rows_per_group: int = 0) -> torch.Tensor:
    """Fused tensor-core MLP forward (csrc/mlp.cu).  ``dirs=None``: densities [n] only (nerfs.py:70-72).
    Grouped layout (``group_counts`` int32 [n / rows_per_group]): ``dirs`` holds ONE direction per group and
    only the first ``group_counts[g]`` rows of each group are evaluated; padding rows are left unwritten."""
    n = enc.shape[0]
    density_only = dirs is None
    out = torch.empty((n,) if density_only else (n, 4), dtype=torch.float32, device=enc.device)
    if n:
        if group_counts is None:
            _lib.call("ngp_nerf_mlp_forward", [enc, enc if density_only else dirs, weights, out],
                      descriptors.make_nerf_mlp_descriptor(n, density_only))
        else:
            _lib.call("ngp_nerf_mlp_forward", [enc, enc if density_only else dirs, weights, group_counts, out],
                      descriptors.make_nerf_mlp_descriptor(n, density_only, rows_per_group))
    return out


def fused_supported(lt, table: torch.Tensor, wrap: str = "jaxngp") -> bool:
    """Shapes ngp_nerf_fused_forward covers (include/ngp_b200.h); anything else runs encoder and MLP as two ops."""
    pair_bytes = 16 if table.dtype == torch.float32 else 8
    return (lt.dim == 3 and lt.L == 16 and lt.F == 2 and wrap == "jaxngp" and lt.T & (lt.T - 1) == 0
            and table.dtype in (torch.float32, torch.float16) and table.data_ptr() % pair_bytes == 0 and lt.rows % 2 == 0)


def fused_forward(lt, pos: torch.Tensor, bound: float, table: torch.Tensor, dirs, weights: torch.Tensor, *,
                  wrap: str = "jaxngp", group_counts: torch.Tensor = None, rows_per_group: int = 0, want_enc: bool = False,
                  impl: str = "mma", out: torch.Tensor = None):
    """Hash-grid encoder fused in front of the MLP forward (csrc/mlp.cu ``nerf_fused_forward_kernel``): bit-identical
    to ``encoders.hashgrid_forward`` + ``mlp_forward``, without the [n, 32] round trip through HBM.
    Returns ``drgbs`` (``dirs=None``: densities [n]); with ``want_enc`` also the encoding (kept for the backward)."""
    n = pos.shape[0]
    density_only = dirs is None
    if out is None:  # `out`: a caller-owned slice to write into (the grid update evaluates its points in chunks)
        out = torch.empty((n,) if density_only else (n, 4), dtype=torch.float32, device=pos.device)
    enc = torch.empty(n, 32, dtype=torch.float32, device=pos.device) if want_enc else None
    if n:
        desc = encoders._a1_descriptor(lt, n, bound, wrap, table.dtype, rows_per_group) + \
            descriptors.struct.pack("<2I", int(density_only), int(want_enc))
        bufs = [pos, table, pos if density_only else dirs, weights]
        if group_counts is not None:
            bufs.append(group_counts)
        bufs.append(out)
        if want_enc:
            bufs.append(enc)
        _lib.call("ngp_nerf_fused_forward_umma" if impl == "umma" and not density_only else "ngp_nerf_fused_forward", bufs, desc)
    return (out, enc) if want_enc else out


BACKWARD_IMPLS = {"tc": "ngp_nerf_mlp_backward_tc", "umma": "ngp_nerf_mlp_backward", "mma": "ngp_nerf_mlp_backward_mma"}
DEFAULT_BACKWARD_IMPL = "umma"


def mlp_backward(enc, dirs, weights, d_drgbs, d_weights=None, impl=None, d_enc=None, accumulate=False):
    """Fused backward: returns (d_enc [n, 32], d_weights [9408]); recomputes the forward on chip.
    ``impl``: "tc" = every matrix product on tcgen05, chain operands in tensor memory (csrc/mlp_bwd_tc.cu);
    "umma" = mma.sync register chain + weight gradients on tcgen05; "mma" = all mma.sync (cross-check arm)."""
    impl = impl or DEFAULT_BACKWARD_IMPL
    n = enc.shape[0]
    if d_enc is None:
        d_enc = torch.empty(n, 32, dtype=torch.float32, device=enc.device)
    if d_weights is None:
        d_weights = torch.empty(MLP_NUMEL, dtype=torch.float32, device=enc.device)
    op = BACKWARD_IMPLS[impl]
    if accumulate:  # d_weights += ... (a batch processed in chunks; the hybrid kernel only)
        if impl != "umma":
            raise ValueError("accumulate=True is implemented for impl='umma'")
        op = "ngp_nerf_mlp_backward_acc"
    _lib.call(op, [enc, dirs, weights, d_drgbs, d_enc, d_weights], descriptors.make_nerf_mlp_descriptor(n))
    return d_enc, d_weights


def mlp_backward_scatter(lt, pos, bound: float, enc, dirs, weights, d_drgbs, d_weights, d_table, wrap: str = "jaxngp"):
    """MLP backward with the hash-table scatter fused behind it (``ngp_nerf_mlp_backward_scatter``): the same sums as
    ``mlp_backward`` followed by ``encoders.hashgrid_backward``, without d_enc ever leaving the SM.  Writes ``d_weights``
    [9408] and ``d_table`` [rows, 2] (both defined by the call)."""
    n = enc.shape[0]
    _lib.call("ngp_nerf_mlp_backward_scatter", [enc, dirs, weights, d_drgbs, pos, d_weights, d_table],
              encoders._a1_descriptor(lt, n, bound, wrap, torch.float32))
    return d_weights, d_table


class _FusedMLP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, dirs, weights):
        enc, dirs, weights = enc.contiguous(), dirs.contiguous(), weights.contiguous()
        ctx.save_for_backward(enc, dirs, weights)
        return mlp_forward(enc, dirs, weights)

    @staticmethod
    def backward(ctx, d_drgbs):
        enc, dirs, weights = ctx.saved_tensors
        d_enc, d_w = mlp_backward(enc, dirs, weights, d_drgbs.contiguous())
        return d_enc, None, d_w


class NeRF(torch.nn.Module):
    """``nerf(xyz, dir, appearance_embeddings) -> (drgbs[..., 4], tv)``; ``dir=None`` returns densities
    only (models/nerfs.py:40-86).  The five Dense kernels live in one flat parameter ``mlp_flat``
    ([in, out] row-major each, flax layout); ``density_w0`` ... ``rgb_w2`` are views of it.
    ``fused=True`` (default) runs the dense layers in the fused tensor-core kernels of csrc/mlp.cu,
    ``fused=False`` through plain torch matmuls (the library-GEMM arm, used as a cross-check)."""

    def __init__(self, bound: float, inference: bool = False, tv_scale: float = 0.0, device=None, generator=None,
                 T: int = 2 ** 19, fused: bool = True):
        super().__init__()
        self.bound = float(bound)
        self.fused = fused
        #: kernel behind ``forward_grouped`` (the inference renderer): "umma" = dense layers on tcgen05 / TMEM over
        #: tiles of live slots only, "mma" = the mma.sync kernel (bit-identical to the ungrouped ``forward``)
        self.grouped_impl = "umma"
        Enc = encoders.TCNNHashGridEncoder if inference else encoders.HashGridEncoder  # nerfs.py:431-438
        self.position_encoder = Enc(L=16, T=T, F=2, N_min=2 ** 4, N_max=int(2 ** 11 * bound), tv_scale=tv_scale,
                                    device=device, generator=generator)
        flat = torch.empty(MLP_NUMEL, dtype=torch.float32, device=device)
        off = 0
        for _, i, o in MLP_SHAPES:  # glorot-uniform, no bias (nerfs.py:98,115-119)
            glorot_uniform_(flat[off:off + i * o].view(i, o), generator)
            off += i * o
        self.mlp_flat = torch.nn.Parameter(flat)

    def _view(self, name):
        off = 0
        for nm, i, o in MLP_SHAPES:
            if nm == name:
                return self.mlp_flat[off:off + i * o].view(i, o)
            off += i * o
        raise KeyError(name)

    density_w0 = property(lambda self: self._view("density_w0"))
    density_w1 = property(lambda self: self._view("density_w1"))
    rgb_w0 = property(lambda self: self._view("rgb_w0"))
    rgb_w1 = property(lambda self: self._view("rgb_w1"))
    rgb_w2 = property(lambda self: self._view("rgb_w2"))

    def mlp_parameters(self):
        return [self.mlp_flat]

    @torch.no_grad()
    def forward_grouped(self, xyzs, ray_dirs, n_samples):
        """Inference fast path for the [n_rays, cap, 3] sample layout of march_rays_inference: one direction
        per ray, only the first ``n_samples[i]`` samples of ray i are evaluated (the rest of the output is
        unspecified, integrate_rays_inference never reads it).  Same numbers as ``forward`` on the live rows."""
        n, cap = xyzs.shape[0], xyzs.shape[1]
        enc_mod = self.position_encoder
        if self.fused and fused_supported(enc_mod.levels, enc_mod.latents, enc_mod.wrap):
            return fused_forward(enc_mod.levels, xyzs.reshape(-1, 3), self.bound, enc_mod.latents.detach(),
                                 ray_dirs.contiguous(), self.mlp_flat.detach(), wrap=enc_mod.wrap, group_counts=n_samples,
                                 rows_per_group=cap, impl=self.grouped_impl if cap <= 128 else "mma").reshape(n, cap, 4)
        enc = encoders.hashgrid_forward(enc_mod.levels, xyzs.reshape(-1, 3), self.bound, enc_mod.latents.detach(),
                                        enc_mod.wrap, group_counts=n_samples, rows_per_group=cap)
        drgbs = mlp_forward(enc, ray_dirs.contiguous(), self.mlp_flat.detach(), group_counts=n_samples, rows_per_group=cap)
        return drgbs.reshape(n, cap, 4)

    def forward(self, xyz, dir=None, appearance_embeddings=None):
        shape = xyz.shape[:-1]
        xyz = xyz.reshape(-1, 3)
        pos_enc, tv = self.position_encoder(xyz, self.bound)
        if self.fused:
            if dir is None:
                if torch.is_grad_enabled() and (pos_enc.requires_grad or self.mlp_flat.requires_grad):
                    raise NotImplementedError("the density-only branch is inference-only (utils/types.py:1208-1217)")
                return mlp_forward(pos_enc.contiguous(), None, self.mlp_flat.detach()).reshape(*shape, 1), tv
            drgbs = _FusedMLP.apply(pos_enc, dir.reshape(-1, 3), self.mlp_flat)
            return drgbs.reshape(*shape, 4), tv
        x = torch.relu(pos_enc @ self.density_w0) @ self.density_w1
        density = trunc_exp(x[:, :1])
        if dir is None:
            return density.reshape(*shape, 1), tv
        h = torch.cat([x, sh4(dir.reshape(-1, 3))], dim=-1)
        rgb = torch.sigmoid(torch.relu(torch.relu(h @ self.rgb_w0) @ self.rgb_w1) @ self.rgb_w2)
        return torch.cat([density, rgb], dim=-1).reshape(*shape, 4), tv
