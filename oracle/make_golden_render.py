"""TEST INFRASTRUCTURE ONLY -- a golden frame from the reference's OWN inference renderer.

Executes, unmodified and on numpy (oracle/ref_shim.py): ``render_image_inference`` with its slot-refill loop and
power-of-two batching, ``march_and_integrate_inference`` (models/renderers/cuda.py:165-373), the volume-rendering-jax
wrappers ``march_rays_inference`` / ``integrate_rays_inference`` (with their ``.at[indices].set`` scatters), ray
generation and ``f32_to_u8``; the two custom-call primitives are served by the C oracle (pinned to the reference's CUDA
kernels by tests/golden/inference_loop.npz) and the NeRF by the analytic field of jaxngp_b200/synthetic.py.  Writes
tests/golden/render_reference.npz (a 96 x 96 frame of view 3).

    python oracle/make_golden_render.py        # needs /root/reference; run in the build container only
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

W = H = 96
VIEW = 3


def analytic_nerf(params, xyzs, dirs, appearance):
    from jaxngp_b200 import synthetic as S
    shape = xyzs.shape[:-1]
    x = np.asarray(xyzs, np.float32).reshape(-1, 3)
    drgbs = np.concatenate([S.density(x)[:, None] * 0.5, S.colour(x)], -1).astype(np.float32)
    return drgbs.reshape(*shape, 4), None


def main():
    from jaxngp_b200 import synthetic as S
    from oracle import oracle as O
    from oracle import ref_shim
    O.build()
    ref = ref_shim.install_renderer(O)
    cam = S.camera()
    s = W / cam["width"]
    camera = ref.make_camera(W, H, cam["fx"] * s, cam["fy"] * s, W / 2, H / 2)
    tf = S.poses(100)[VIEW]
    transform = types.SimpleNamespace(rotation=ref.array(tf[:9].reshape(3, 3).astype(np.float32)), translation=ref.array(tf[9:].astype(np.float32)))
    state = types.SimpleNamespace(
        scene_meta=types.SimpleNamespace(camera=camera, bound=1.0, cascades=1, stepsize_portion=0.0),
        raymarch=types.SimpleNamespace(diagonal_n_steps=1024, density_grid_res=128),
        render=types.SimpleNamespace(random_bg=False, bg=(1.0, 1.0, 1.0)), use_background_model=False,
        locked_params={"nerf": None}, nerf_fn=analytic_nerf, ogrid=types.SimpleNamespace(occupancy=ref.array(S.occupancy_bitfield())))
    bg, image, distance, cost = ref.render_image_inference(None, transform, state)
    image, distance = np.asarray(image), np.asarray(distance)
    assert image.dtype == np.uint8 and image.shape == (H, W, 3) and distance.shape == (H, W) and cost is None
    path = os.path.join(ROOT, "tests", "golden", "render_reference.npz")
    np.savez_compressed(path, image=image, distance=distance, width=np.int64(W), height=np.int64(H), view=np.int64(VIEW))
    print("wrote", path, os.path.getsize(path), "bytes; mean level", float(image.mean()), "object pixels", int((image < 250).any(-1).sum()))


if __name__ == "__main__":
    main()
