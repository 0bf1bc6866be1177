"""TEST INFRASTRUCTURE ONLY -- the forward of one whole training step from the reference's OWN code.

perm -> rays -> march_rays -> HashGridEncoder -> NeRF MLP -> integrate_rays -> Huber loss, every Python / JAX line of it
the reference's (app/nerf/_utils.py:93-162, models/renderers/cuda.py:57-162, the volume-rendering-jax wrappers,
models/nerfs.py, models/encoders.py, utils/data.py), unmodified, on numpy through oracle/ref_shim.py; the two CUDA
primitives are served by the C oracle.  Writes tests/golden/train_forward_reference.npz: the inputs that are not
regenerated from seeds, the loss and the batch metrics.

    python oracle/make_golden_train_forward.py        # needs /root/reference; ~1 minute
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_RAYS, TOTAL_SAMPLES, N_VIEWS = 384, 12288, 100
SHAPES = (("density_w0", 32, 64), ("density_w1", 64, 16), ("rgb_w0", 32, 64), ("rgb_w1", 64, 64), ("rgb_w2", 64, 3))


def make_inputs():
    from jaxngp_b200 import synthetic as S
    from oracle import hashgrid_np as H
    from tests import inputs
    rng = np.random.Generator(np.random.PCG64(123))
    cam = S.camera()
    perm = rng.integers(0, N_VIEWS * cam["width"] * cam["height"], N_RAYS, dtype=np.int64).astype(np.uint32)
    noises = rng.random(N_RAYS, dtype=np.float32)
    bg = rng.random((N_RAYS, 3), dtype=np.float32)
    lv = H.level_table(16, 2 ** 19, 2, 16, 2048, 3)
    table = inputs.encoder_table(int(lv["offsets"][-1]), 2, amp=0.5)
    w = {}
    for name, i, o in SHAPES:
        lim = np.sqrt(6.0 / (i + o))
        w[name] = rng.uniform(-lim, lim, (i, o)).astype(np.float32)
    w["density_w1"][:, 0] += 1.5  # denser medium: some rays saturate, others do not
    # ground-truth pixels for the sampled rays only (the reference gathers scene.rgbas_u8[perm])
    rgba_rows = rng.integers(0, 256, (N_RAYS, 4), dtype=np.uint8)
    return dict(cam=cam, perm=perm, noises=noises, bg=bg, lv=lv, table=table, w=w, rgba_rows=rgba_rows)


def main():
    from jaxngp_b200 import synthetic as S
    from oracle import oracle as O
    from oracle import ref_shim
    O.build()
    d = make_inputs()
    cam = d["cam"]
    jr = ref_shim.ScriptedRandom([], [d["bg"], d["noises"]])  # loss_fn draws the backgrounds, render_rays_train the noises
    ref = ref_shim.install_train_forward(O, jr)
    camera = ref.make_camera(cam["width"], cam["height"], cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    camera.near = cam["near"]

    class Rows:  # scene.rgbas_u8[perm] without materialising 64 M pixels
        def __getitem__(self, idx):
            return d["rgba_rows"]

    loss, metrics = ref.forward(d["perm"], S.poses(N_VIEWS), camera, d["table"], d["w"], S.occupancy_bitfield(), Rows(), TOTAL_SAMPLES, N_VIEWS)
    assert not jr._uniforms
    out = dict(loss=np.float64(loss), n_valid_rays=np.int64(metrics["n_valid_rays"]),
               measured_batch_size_before_compaction=np.int64(metrics["measured_batch_size_before_compaction"]),
               measured_batch_size=np.int64(metrics["measured_batch_size"]), ray_is_valid=np.asarray(metrics["ray_is_valid"]).astype(bool))
    path = os.path.join(ROOT, "tests", "golden", "train_forward_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v if v.ndim == 0 else v.sum()) for k, v in out.items()})


if __name__ == "__main__":
    main()
