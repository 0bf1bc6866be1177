"""TEST INFRASTRUCTURE ONLY -- golden vectors for the NeRF model (f1) from the reference's OWN code.

Builds ``make_nerf_ngp(bound=1, inference=False)`` (models/nerfs.py:422-454) and calls the resulting ``NeRF`` module
(models/nerfs.py:27-86), its ``CoordinateBasedMLP``s (:89-128), ``trunc_exp`` (:222-238) and the
``SphericalHarmonicsEncoder`` (models/encoders.py:365-406), all unmodified, on numpy through oracle/ref_shim.py
(flax's bias-free Dense = ``x @ kernel`` and sigmoid are restated library code).  Writes tests/golden/nerf_reference.npz:
inputs, weights, densities / colours, the SH basis, and trunc_exp's backward rule evaluated on a grid.

    python oracle/make_golden_nerf.py        # needs /root/reference; run in the build container only
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = (("density_w0", 32, 64), ("density_w1", 64, 16), ("rgb_w0", 32, 64), ("rgb_w1", 64, 64), ("rgb_w2", 64, 3))
N = 192


def main():
    from oracle import hashgrid_np as H
    from oracle import ref_shim
    from tests import inputs
    nerfs = ref_shim.install_nerf()
    model = nerfs.make_nerf_ngp(bound=1.0, inference=False)
    rng = np.random.Generator(np.random.PCG64(99))
    lv = H.level_table(16, 2 ** 19, 2, 16, 2048, 3)
    rows = int(lv["offsets"][-1])
    table = inputs.encoder_table(rows, 2, amp=1.0)
    w = {}
    for name, i, o in SHAPES:
        lim = np.sqrt(6.0 / (i + o))
        w[name] = rng.uniform(-lim, lim, (i, o)).astype(np.float32)
    xyz = inputs.encoder_points(N, 3)
    dirs = rng.normal(size=(N, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    model.position_encoder.bind_params(**{"latent codes stored on grid vertices": table})
    model.density_mlp.bind_params(Dense_0=w["density_w0"], Dense_1=w["density_w1"])
    model.rgb_mlp.bind_params(Dense_0=w["rgb_w0"], Dense_1=w["rgb_w1"], Dense_2=w["rgb_w2"])
    drgbs, tv = model(xyz, dirs, np.zeros((0,), np.float32))
    density_only, _ = model(xyz, None, None)
    sh = model.direction_encoder(dirs)
    act = nerfs.make_activation("truncated_exponential")
    xs = np.linspace(-20, 20, 81).astype(np.float32)
    gs = rng.normal(size=xs.shape).astype(np.float32)
    (grad_x,) = act.bwd(xs, gs)
    out = dict(xyz=xyz, dirs=dirs, drgbs=np.asarray(drgbs), density_only=np.asarray(density_only), sh=np.asarray(sh),
               trunc_exp_x=xs, trunc_exp_g=gs, trunc_exp_grad=np.asarray(grad_x), trunc_exp_fwd=np.asarray(act(xs)), **w)
    for k, v in out.items():
        assert v.dtype == np.float32, (k, v.dtype)
    assert out["drgbs"].shape == (N, 4) and out["sh"].shape == (N, 16) and out["density_only"].shape == (N, 1) and tv == 0
    path = os.path.join(ROOT, "tests", "golden", "nerf_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; density range", float(out["drgbs"][:, 0].min()), float(out["drgbs"][:, 0].max()))


if __name__ == "__main__":
    main()
