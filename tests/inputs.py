"""Seeded inputs shared by the oracle tests, the GPU parity tests and oracle/make_golden.py."""
import numpy as np

from jaxngp_b200 import synthetic as S


def march_case(name):
    """Returns (static kwargs, array kwargs) for march_rays."""
    if name == "scene":  # NeRF-synthetic-shaped, nothing overflows
        r = S.training_rays(2048, seed=1000000007)
        st = dict(total_samples=1 << 19, diagonal_n_steps=1024, K=1, G=128, bound=1.0, stepsize_portion=0.0)
        bits = S.occupancy_bitfield()
    elif name == "overflow":  # the budget fills after a few dozen rays
        r = S.training_rays(4096, seed=7)
        st = dict(total_samples=8192, diagonal_n_steps=1024, K=1, G=128, bound=1.0, stepsize_portion=0.0)
        bits = S.occupancy_bitfield()
    elif name == "cascades":  # K=3, bound=4, exponential stepping, random occupancy
        rng = np.random.Generator(np.random.PCG64(11))
        n = 2048
        o = rng.uniform(-3.5, 3.5, (n, 3)).astype(np.float32)
        d = rng.normal(size=(n, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        ts, te = S.near_far(o, d, 4.0)
        r = dict(rays_o=o, rays_d=d.astype(np.float32), t_starts=ts, t_ends=te, noises=rng.random(n, dtype=np.float32))
        st = dict(total_samples=1 << 19, diagonal_n_steps=1024, K=3, G=32, bound=4.0, stepsize_portion=1.0 / 256)
        bits = rng.integers(0, 256, size=3 * 32 ** 3 // 8, dtype=np.uint8) & rng.integers(0, 256, size=3 * 32 ** 3 // 8, dtype=np.uint8)
    elif name == "dense":  # all-ones bitfield (state at step 0, utils/types.py:123-126): per-ray cap 1024*bound bites
        r = S.training_rays(512, seed=3)
        # a few rays along the cube diagonal: 2*sqrt(3)/ds = 1024 steps, so the per-ray cap bites
        diag = np.float32(1 / np.sqrt(3))
        for k, sgn in enumerate(((1, 1, 1), (-1, 1, 1), (1, -1, -1), (-1, -1, 1))):
            sg = np.asarray(sgn, np.float32)
            r["rays_d"][k] = sg * diag
            r["rays_o"][k] = -sg * np.float32(1.3)
            r["noises"][k] = np.float32(0.01 * k)
        r["t_starts"], r["t_ends"] = S.near_far(r["rays_o"], r["rays_d"])
        st = dict(total_samples=1 << 19, diagonal_n_steps=1024, K=1, G=128, bound=1.0, stepsize_portion=0.0)
        bits = np.full(128 ** 3 // 8, 0xFF, np.uint8)
    elif name == "miss":  # rays that never enter the box, and an empty grid
        r = S.training_rays(256, seed=5)
        r["rays_o"] = r["rays_o"] + np.float32(10)
        r["t_starts"], r["t_ends"] = S.near_far(r["rays_o"], r["rays_d"])
        st = dict(total_samples=1024, diagonal_n_steps=1024, K=1, G=128, bound=1.0, stepsize_portion=0.0)
        bits = np.zeros(128 ** 3 // 8, np.uint8)
    else:
        raise KeyError(name)
    arrays = dict(rays_o=r["rays_o"], rays_d=r["rays_d"], t_starts=r["t_starts"], t_ends=r["t_ends"],
                  noises=r["noises"], occupancy_bitfield=bits)
    return st, arrays


MARCH_CASES = ("scene", "overflow", "cascades", "dense", "miss")


def drgbs_for(xyzs, seed, scale=1.0):
    """Random-but-structured (density, rgb) predictions for a sample array."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = xyzs.shape[0]
    sigma = np.exp(rng.normal(1.0, 2.0, n)).astype(np.float32) * np.float32(scale)
    rgb = rng.random((n, 3), dtype=np.float32)
    return np.concatenate([sigma[:, None], rgb], axis=-1).astype(np.float32)


def inference_case(width=96, height=96, n_rays=1024, cap=8):
    fr = S.frame_rays(view=3, width=width, height=height)
    st = dict(diagonal_n_steps=1024, K=1, G=128, march_steps_cap=cap, bound=1.0, stepsize_portion=0.0)
    return st, fr, S.occupancy_bitfield(), n_rays


def encoder_points(n, dim=3, seed=42):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.uniform(-1, 1, (n, dim)).astype(np.float32)


def encoder_table(rows, F, seed=43, amp=1e-4):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.uniform(-amp, amp, (rows, F)).astype(np.float32)
