"""Synthetic NeRF-synthetic-shaped scene: cameras, poses, a procedural density /
colour field and its occupancy bitfield.  Host-side numpy data generation used by bench.py, the
tests and smoke(); there is no dataset on the box (no network).

Shapes follow the reference's loaders: camera from ``camera_angle_x`` (utils/data.py:655-667),
``scene.transforms[view] = [R (9, row-major), t (3)]`` (app/nerf/_utils.py:107-111), rays through
pixel centres with CV->CG axis flip (utils/types.py:398-439), bound=1 => K=1, stepsize_portion=0
(utils/types.py:858-886).
"""
import math

import numpy as np

W = H = 800
CAMERA_ANGLE_X = 0.6911
NEAR = 0.3          # utils/args.py:150
BOUND = 1.0
G = 128
K = 1
DIAGONAL_N_STEPS = 1024
STEPSIZE_PORTION = 0.0
SIGMA_INSIDE = 64.0
DENSITY_THRESHOLD = 0.01 * DIAGONAL_N_STEPS / (2 * min(BOUND, 1.0) * 3 ** 0.5)  # utils/types.py:1367-1369


def camera():
    f = 0.5 * W / math.tan(CAMERA_ANGLE_X / 2)
    return dict(width=W, height=H, fx=f, fy=f, cx=W / 2, cy=H / 2, near=NEAR)


def poses(n=100, radius=4.03 / 3):
    """[n, 12] camera-to-world transforms on a golden-angle spiral, looking at the origin
    (CG convention: camera looks down -z, +y up)."""
    out = np.zeros((n, 12), np.float32)
    ga = math.pi * (3 - math.sqrt(5))
    for i in range(n):
        zc = 0.15 + 0.8 * (i + 0.5) / n  # upper hemisphere, like the blender scenes
        r = math.sqrt(max(0.0, 1 - zc * zc))
        c = np.array([r * math.cos(ga * i), r * math.sin(ga * i), zc]) * radius
        fwd = -c / np.linalg.norm(c)
        right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
        right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        R = np.stack([right, up, -fwd], axis=1)  # columns: camera x, y, z axes in world
        out[i, :9] = R.reshape(-1)
        out[i, 9:] = c
    return out


def pixel_rays(transforms, view_idcs, pixel_idcs, cam=None):
    """World-space rays for (view, pixel) pairs: app/nerf/_utils.py:97-115 +
    utils/types.py:398-439 (no distortion, PERSPECTIVE)."""
    cam = cam or camera()
    x = (pixel_idcs % cam["width"]).astype(np.float32)
    y = (pixel_idcs // cam["width"]).astype(np.float32)
    dx = ((x + np.float32(0.5)) - np.float32(cam["cx"])) / np.float32(cam["fx"])
    dy = ((y + np.float32(0.5)) - np.float32(cam["cy"])) / np.float32(cam["fy"])
    d_cam = np.stack([dx, -dy, -np.ones_like(dx)], axis=-1)
    d_cam = d_cam / np.linalg.norm(d_cam, axis=-1, keepdims=True)
    R = transforms[view_idcs, :9].reshape(-1, 3, 3)
    d_world = (d_cam[:, None, :] * R).sum(-1).astype(np.float32)
    o_world = transforms[view_idcs, 9:].astype(np.float32)
    return np.ascontiguousarray(o_world), np.ascontiguousarray(d_world)


def near_far(o, d, bound=BOUND):
    """models/renderers/cuda.py:57-97 (numpy f32)."""
    eps = np.float32(1e-15)
    d = np.where(np.signbit(d), np.minimum(d, -eps), np.maximum(d, eps)).astype(np.float32)
    b = np.float32(bound)
    t0 = (-b - o) / d
    t1 = (b - o) / d
    ts, te = np.minimum(t0, t1), np.maximum(t0, t1)
    t_start = np.maximum(ts.max(axis=-1), np.float32(0))
    t_end = te.min(axis=-1)
    return t_start.astype(np.float32), t_end.astype(np.float32)


def density(xyz):
    """sigma(x): union of a sphere (r=.45 at the origin) and a box, SIGMA_INSIDE inside, 0 outside."""
    xyz = np.asarray(xyz)
    sphere = (xyz ** 2).sum(-1) < 0.45 ** 2
    lo, hi = np.array([-0.7, -0.3, -0.6]), np.array([-0.2, 0.3, 0.1])
    box = np.all((xyz > lo) & (xyz < hi), axis=-1)
    return np.where(sphere | box, np.float32(SIGMA_INSIDE), np.float32(0)).astype(np.float32)


def colour(xyz):
    return (0.5 + 0.5 * np.clip(np.asarray(xyz, np.float32), -1, 1)).astype(np.float32)


def _compact_bits(x):
    x = x & 0x49249249
    x = (x | (x >> 2)) & 0xc30c30c3
    x = (x | (x >> 4)) & 0x0f00f00f
    x = (x | (x >> 8)) & 0xff0000ff
    x = (x | (x >> 16)) & 0x0000ffff
    return x


def grid_cell_centres(G_=G, cas=0, bound=BOUND):
    """Morton-ordered cell centres of cascade `cas` (utils/types.py:1193-1197 without the jitter)."""
    idx = np.arange(G_ ** 3, dtype=np.uint32)
    xyz = np.stack([_compact_bits(idx), _compact_bits(idx >> 1), _compact_bits(idx >> 2)], -1).astype(np.float32)
    mip_bound = min(2.0 ** cas, bound)
    return ((xyz + 0.5) / G_ * 2 - 1) * mip_bound


def density_grid(G_=G, K_=K, bound=BOUND):
    return np.concatenate([density(grid_cell_centres(G_, c, bound)) for c in range(K_)]).astype(np.float32)


def occupancy_bitfield(G_=G, K_=K, bound=BOUND, threshold=DENSITY_THRESHOLD):
    occ = density_grid(G_, K_, bound) > np.float32(threshold)
    return np.packbits(occ, bitorder="little")


def training_batch(n_rays, seed=1000000007, n_views=100):
    """C2 batch: `n_rays` pixels drawn uniformly over all views, perturbation noises U[0,1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    perm = rng.integers(0, n_views * W * H, size=n_rays, dtype=np.int64)
    noises = np.random.Generator(np.random.PCG64(seed + 1)).random(n_rays, dtype=np.float32)
    return perm, noises


def training_rays(n_rays, seed=1000000007, n_views=100):
    tf = poses(n_views)
    perm, noises = training_batch(n_rays, seed, n_views)
    view_idcs, pixel_idcs = perm // (W * H), perm % (W * H)
    o, d = pixel_rays(tf, view_idcs, pixel_idcs)
    ts, te = near_far(o, d)
    return dict(rays_o=o, rays_d=d, t_starts=ts, t_ends=te, noises=noises, view_idcs=view_idcs.astype(np.int32))


def frame_rays(view=0, width=W, height=H, n_views=100):
    """C3: every pixel-centre ray of one pose."""
    cam = camera()
    if (width, height) != (W, H):
        s = width / W
        cam = dict(width=width, height=height, fx=cam["fx"] * s, fy=cam["fy"] * s, cx=width / 2, cy=height / 2, near=NEAR)
    tf = poses(n_views)
    pix = np.arange(width * height, dtype=np.int64)
    o, d = pixel_rays(tf, np.full(pix.shape, view), pix, cam)
    ts, te = near_far(o, d)
    return dict(rays_o=o, rays_d=d, t_starts=ts, t_ends=te)
