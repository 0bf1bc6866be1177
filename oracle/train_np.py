"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- the reference's training step on the host cores:
march (C oracle) -> hash-grid encode (C oracle) -> MLP (numpy/BLAS restatement of models/nerfs.py:27-128,
216-238) -> integrate (C oracle) -> Huber loss (app/nerf/_utils.py:151-156) -> backward of all of it
-> Adam (app/nerf/_utils.py:19-77).  jax is absent from this image, so this stands in for the
"JAX-CPU path" named by BASELINE.json; bench.py times it (cpu_baseline, --impl reference) and
tests/ use it as the checker of the GPU training step.  Never imported by the product.
Pinned: the MLP, SH basis and trunc_exp reproduce the reference's own models/nerfs.py + models/encoders.py, run
unmodified on numpy stand-ins (oracle/ref_shim.py, oracle/make_golden_nerf.py -> tests/golden/nerf_reference.npz).
"""
import numpy as np

from . import hashgrid_np as H
from . import oracle as O

SH_C = [0.28209479177387814, 0.48860251190291987, 1.0925484305920792, 0.94617469575755997, 0.31539156525251999,
        0.54627421529603959, 0.59004358992664352, 2.8906114426405538, 0.45704579946446572, 0.3731763325901154,
        1.4453057213202769]


def sh4(d):  # models/encoders.py:365-406
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    c = SH_C
    return np.stack([
        np.full_like(x, c[0]), -c[1] * y, c[1] * z, -c[1] * x, c[2] * xy, -c[2] * yz, c[3] * z2 - c[4], -c[2] * xz,
        c[5] * x2 - c[5] * y2, c[6] * y * (-3.0 * x2 + y2), c[7] * xy * z, c[8] * y * (1.0 - 5.0 * z2),
        c[9] * z * (5.0 * z2 - 3.0), c[8] * x * (1.0 - 5.0 * z2), c[10] * z * (x2 - y2), c[6] * x * (-x2 + 3.0 * y2),
    ], axis=-1).astype(np.float32)


def mlp_forward(w, enc, dirs):
    """w = dict(density_w0 [32,64], density_w1 [64,16], rgb_w0 [32,64], rgb_w1 [64,64], rgb_w2 [64,3])."""
    a0 = enc @ w["density_w0"]
    h0 = np.maximum(a0, 0)
    x = h0 @ w["density_w1"]
    density = np.exp(x[:, :1])
    hin = np.concatenate([x, sh4(dirs)], axis=-1)
    a1 = hin @ w["rgb_w0"]
    h1 = np.maximum(a1, 0)
    a2 = h1 @ w["rgb_w1"]
    h2 = np.maximum(a2, 0)
    a3 = h2 @ w["rgb_w2"]
    rgb = 1 / (1 + np.exp(-a3))
    cache = dict(enc=enc, a0=a0, h0=h0, x=x, hin=hin, a1=a1, h1=h1, a2=a2, h2=h2, rgb=rgb)
    return np.concatenate([density, rgb], axis=-1).astype(np.float32), cache


def mlp_backward(w, cache, d_drgbs):
    g = {}
    d_a3 = d_drgbs[:, 1:] * cache["rgb"] * (1 - cache["rgb"])
    g["rgb_w2"] = cache["h2"].T @ d_a3
    d_a2 = (d_a3 @ w["rgb_w2"].T) * (cache["a2"] > 0)
    g["rgb_w1"] = cache["h1"].T @ d_a2
    d_a1 = (d_a2 @ w["rgb_w1"].T) * (cache["a1"] > 0)
    g["rgb_w0"] = cache["hin"].T @ d_a1
    d_x = (d_a1 @ w["rgb_w0"].T)[:, :16].copy()
    d_x[:, 0] += d_drgbs[:, 0] * np.exp(np.clip(cache["x"][:, 0], -15, 15))  # trunc_exp, nerfs.py:222-238
    g["density_w1"] = cache["h0"].T @ d_x
    d_a0 = (d_x @ w["density_w1"].T) * (cache["a0"] > 0)
    g["density_w0"] = cache["enc"].T @ d_a0
    return g, (d_a0 @ w["density_w0"].T).astype(np.float32)


def huber_grad(pred, target, valid, delta=0.1):
    """loss = sum_valid mean_c huber(pred - target) / n_valid; returns (loss, dL/dpred)."""
    err = pred - target
    a = np.abs(err)
    q = np.minimum(a, delta)
    per = (0.5 * q * q + delta * (a - q)).mean(-1)
    n_valid = max(int(valid.sum()), 1)
    loss = float(np.where(valid, per, 0).sum() / valid.sum()) if valid.sum() else float("nan")
    grad = np.clip(err, -delta, delta) / 3.0 / n_valid
    return loss, (grad * valid[:, None]).astype(np.float32)


class AdamNp:
    def __init__(self, lr=1e-2):
        self.lr, self.t, self.state = lr, 0, {}

    def lr_at(self, count):  # optax.exponential_decay, _utils.py:20-27
        if count <= 10_000:
            return self.lr
        return max(self.lr * (1 / 3) ** ((count - 10_000) // 10_000), self.lr / 100)

    def step(self, params, grads, decay_keys=()):
        lr = self.lr_at(self.t)
        self.t += 1
        for k, g in grads.items():
            m, v = self.state.setdefault(k, [np.zeros_like(params[k]), np.zeros_like(params[k])])
            m *= 0.9
            m += 0.1 * g
            v *= 0.99
            v += 0.01 * g * g
            upd = -lr * (m / (1 - 0.9 ** self.t)) / (np.sqrt(v / (1 - 0.99 ** self.t) + 1e-15) + 1e-15)
            if k in decay_keys:
                upd = upd + 1e-6 * params[k]  # optax.add_decayed_weights chained after adam, _utils.py:45-77
            params[k] += upd.astype(np.float32)


def train_step(params, opt, levels, bitfield, rays, gt_rgba, bg, total_samples, *, steps=1024, K=1, G=128, bound=1.0,
               portion=0.0, near=0.3, apply=True):
    """One reference training step on the CPU.  params = MLP dict + {"table"}.  Returns metrics and grads."""
    mb, valid, rn, rs, idcs, xyzs, dirs, dss, zs = O.march_rays(total_samples, steps, K, G, bound, portion,
                                                                rays["rays_o"], rays["rays_d"], rays["t_starts"],
                                                                rays["t_ends"], rays["noises"], bitfield)
    enc = O.hashgrid_encode(levels, xyzs, bound, params["table"])
    drgbs, cache = mlp_forward(params, enc, dirs)
    mbs, rgbd, opac = O.integrate_rays(near, rs, rn, bg, dss, zs, drgbs)
    gt_rgb = gt_rgba[:, :3] * gt_rgba[:, 3:] + bg * (1 - gt_rgba[:, 3:])
    loss, d_rgb = huber_grad(rgbd[:, :3], gt_rgb, valid)
    d_fin = np.concatenate([d_rgb, np.zeros((d_rgb.shape[0], 1), np.float32)], axis=-1)
    _, _, d_drgbs = O.integrate_rays_backward(near, rs, rn, bg, dss, zs, drgbs, rgbd, opac, d_fin)
    grads, d_enc = mlp_backward(params, cache, d_drgbs)
    grads["table"] = O.hashgrid_backward(levels, xyzs, bound, d_enc, params["table"].shape[1]).astype(np.float32)
    if apply:
        opt.step(params, grads, decay_keys=("density_w0", "density_w1", "rgb_w0", "rgb_w1", "rgb_w2"))
    return dict(loss=loss, n_valid_rays=int(valid.sum()), measured_batch_size_before_compaction=int(mb),
                measured_batch_size=int(mbs)), grads
