"""TEST INFRASTRUCTURE ONLY -- golden vectors for ray generation from the reference's OWN code.

Executes, unmodified and on numpy (oracle/ref_shim.py): ``Camera.make_ray_directions_from_pixel_coordinates``
(utils/types.py:398-439), the ray construction nested in ``train_step`` (app/nerf/_utils.py:96-115),
``make_rays_worldspace`` and ``make_near_far_from_bound`` (models/renderers/cuda.py:22-97), on the synthetic
NeRF-synthetic-shaped camera and poses of jaxngp_b200/synthetic.py.  Writes tests/golden/rays_reference.npz.

    python oracle/make_golden_rays.py        # needs /root/reference; run in the build container only
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from jaxngp_b200 import synthetic as S
    from oracle import ref_shim
    ref = ref_shim.install_rays()
    cam = S.camera()
    rc = ref.make_camera(cam["width"], cam["height"], cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    tf = S.poses(100)
    rng = np.random.Generator(np.random.PCG64(77))
    perm = rng.integers(0, 100 * cam["width"] * cam["height"], 2048, dtype=np.int64).astype(np.uint32)
    perm[:4] = [0, cam["width"] - 1, cam["width"] * cam["height"] - 1, 99 * cam["width"] * cam["height"] + 400 * cam["width"] + 400]
    o, d = ref.train_rays(rc, tf, perm)
    ts, te = ref.make_near_far_from_bound(1.0, o, d)
    out = dict(perm=perm, train_o=np.asarray(o), train_d=np.asarray(d), train_t_starts=np.asarray(ts), train_t_ends=np.asarray(te))
    # a frame: 48 x 32 pixels of view 7 through make_rays_worldspace (rotation / translation of the [12] transform)
    small = ref.make_camera(48, 32, cam["fx"] * 48 / cam["width"], cam["fy"] * 48 / cam["width"], 24.0, 16.0)
    import types
    t7 = types.SimpleNamespace(rotation=ref.array(tf[7, :9].reshape(3, 3).astype(np.float32)), translation=ref.array(tf[7, 9:].astype(np.float32)))
    fo, fd = ref.make_rays_worldspace(small, t7)
    fts, fte = ref.make_near_far_from_bound(1.0, fo, fd)
    out.update(frame_o=np.asarray(fo), frame_d=np.asarray(fd), frame_t_starts=np.asarray(fts), frame_t_ends=np.asarray(fte))
    for k, v in out.items():
        assert v.dtype in (np.float32, np.uint32), (k, v.dtype)
    path = os.path.join(ROOT, "tests", "golden", "rays_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "hit fraction", float((out["train_t_starts"] < out["train_t_ends"]).mean()))


if __name__ == "__main__":
    main()
