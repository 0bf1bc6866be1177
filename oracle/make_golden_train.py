"""TEST INFRASTRUCTURE ONLY -- golden vectors for the loss and the optimizer from the reference's OWN code.

Executes, unmodified and on numpy (oracle/ref_shim.py): the ``loss_fn`` nested in ``train_step``
(app/nerf/_utils.py:117-162: random background, ``blend_rgba_image_array`` utils/data.py:443-464, Huber(0.1) averaged
over the channels, masked by ``ray_is_valid`` and normalised by ``n_valid_rays``) with the renderer replaced by its
outputs, and ``make_optimizer`` (app/nerf/_utils.py:19-77) applied for a few steps to a small parameter tree.  optax is
not on disk; its primitives are restated from their published definitions in the shim.  Writes
tests/golden/train_reference.npz.

    python oracle/make_golden_train.py        # needs /root/reference; run in the build container only
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import ref_shim
    rng = np.random.Generator(np.random.PCG64(31))
    n = 96
    bg = rng.random((n, 3), dtype=np.float32)
    pred = rng.random((n, 4), dtype=np.float32)
    pred[: n // 3, :3] = bg[: n // 3] + rng.normal(size=(n // 3, 3)).astype(np.float32) * 0.03  # errors inside the quadratic zone
    valid = rng.random(n) < 0.8
    gt = rng.random((n, 4), dtype=np.float32)
    gt[: n // 3, 3] = 0.0  # transparent ground truth: the target is the background
    out = dict(bg=bg, pred=pred, valid=valid, gt_rgba=gt)

    def ref_loss(p):
        jr = ref_shim.ScriptedRandom([], [bg])  # the one jran.uniform call of loss_fn draws the background
        return float(ref_shim.install_loss_and_optimizer(jr).loss(p, valid, gt)[0])

    out["loss"] = np.float64(ref_loss(pred))
    # central differences of the reference's own loss on a subset of entries (Huber is piecewise quadratic / linear)
    eps, fd = 2e-3, np.zeros((32, 3), np.float64)
    for i in range(32):
        for c in range(3):
            hi, lo = pred.copy(), pred.copy()
            hi[i, c] += eps
            lo[i, c] -= eps
            fd[i, c] = (ref_loss(hi) - ref_loss(lo)) / (float(hi[i, c]) - float(lo[i, c]))
    out["loss_fd_grad"] = fd
    blend = ref_shim.install_loss_and_optimizer(ref_shim.ScriptedRandom([], []))
    out["blend"] = np.asarray(blend.blend_rgba_image_array(imgarr=blend.array(gt), bg=blend.array(bg)))

    L = ref_shim.install_loss_and_optimizer(ref_shim.ScriptedRandom([], []))
    opt = L.make_optimizer(1e-2)
    out["lr_counts"] = np.array([0, 1, 9999, 10000, 10001, 19999, 20000, 20001, 30000, 50000, 60000, 200000], np.int64)
    out["lr_values"] = np.array([L.optax.last_schedule(int(c)) for c in out["lr_counts"]], np.float64)
    params = {"nerf": {"position_encoder": (rng.normal(size=(40, 2)) * 1e-4).astype(np.float32),
                       "density_mlp": rng.normal(size=(8, 4)).astype(np.float32) * 0.3,
                       "rgb_mlp": rng.normal(size=(4, 3)).astype(np.float32) * 0.3}}
    state = opt.init(params)
    for k, v in params["nerf"].items():
        out[f"opt_p0_{k}"] = v.copy()
    for step in range(4):
        grads = {"nerf": {k: (rng.normal(size=v.shape) * (1e-3 if k == "position_encoder" else 1e-2)).astype(np.float32)
                          for k, v in params["nerf"].items()}}
        if step == 2:
            grads["nerf"]["position_encoder"][::2] = 0  # sparse table gradients (most rows untouched in a step)
        updates, state = opt.update(grads, state, params)
        params = {"nerf": {k: (params["nerf"][k] + updates["nerf"][k]).astype(np.float32) for k in params["nerf"]}}  # optax.apply_updates
        for k in params["nerf"]:
            out[f"opt_g{step}_{k}"] = grads["nerf"][k]
            out[f"opt_p{step + 1}_{k}"] = params["nerf"][k].copy()
    path = os.path.join(ROOT, "tests", "golden", "train_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; loss", float(out["loss"]), "lr", out["lr_values"])


if __name__ == "__main__":
    main()
