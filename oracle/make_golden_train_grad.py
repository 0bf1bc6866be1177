"""TEST INFRASTRUCTURE ONLY -- gradients of one training step by central differences of the reference's OWN forward.

jax's autodiff is the one thing oracle/ref_shim.py cannot stand in for, but the loss of a training step is a function
of the parameters that the reference's unmodified source evaluates here (oracle/make_golden_train_forward.py), so its
derivative with respect to individual parameters is measured directly: for a handful of MLP weights and hash-table
entries, (loss(p + h) - loss(p - h)) / 2h through the whole reference chain (rays -> march -> encoder -> MLP ->
integrate -> Huber loss).  oracle/train_np.train_step's analytic gradients -- the checker of the CUDA backward pass --
are held to these numbers in tests/test_oracle_golden.py.  Writes tests/golden/train_grad_reference.npz (~3 minutes).

    python oracle/make_golden_train_grad.py        # needs /root/reference
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_RAYS, TOTAL_SAMPLES, N_VIEWS = 96, 2048, 100
# (parameter, flat index, step).  Colour-path parameters only -- rgb MLP weights and the density MLP's output columns
# 1..15, which feed the colour branch alone: there the reference's hand-written VJP is the true derivative.  Its density
# gradient is NOT (integrating.cu:199-225: the background term of a non-terminated ray is subtracted twice -- the final
# colour already contains T * bg -- and the result is scaled by min(z^2, 1) and offset by a near-plane penalty), so
# the density-path probes below are recorded for the record and only required to differ.
PROBES = [("rgb_w0", 39, 2e-2), ("rgb_w0", 20 * 64 + 33, 2e-2), ("rgb_w1", 1051, 2e-2), ("rgb_w1", 5 * 64 + 7, 2e-2),
          ("rgb_w2", 154, 2e-2), ("rgb_w2", 14 * 3 + 1, 2e-2), ("density_w1", 30 * 16 + 5, 2e-2), ("density_w1", 12 * 16 + 9, 2e-2),
          ("density_w1", 7 * 16 + 0, 2e-2), ("density_w0", 20 * 64 + 40, 2e-2)]
TABLE_STEP = 5e-2


def make_inputs():
    from jaxngp_b200 import synthetic as S
    from oracle import hashgrid_np as H
    from tests import inputs
    rng = np.random.Generator(np.random.PCG64(321))
    cam = S.camera()
    perm = rng.integers(0, N_VIEWS * cam["width"] * cam["height"], N_RAYS, dtype=np.int64).astype(np.uint32)
    noises = rng.random(N_RAYS, dtype=np.float32)
    bg = rng.random((N_RAYS, 3), dtype=np.float32)
    lv = H.level_table(16, 2 ** 19, 2, 16, 2048, 3)
    table = inputs.encoder_table(int(lv["offsets"][-1]), 2, amp=0.5)
    w = {}
    for name, i, o in (("density_w0", 32, 64), ("density_w1", 64, 16), ("rgb_w0", 32, 64), ("rgb_w1", 64, 64), ("rgb_w2", 64, 3)):
        lim = np.sqrt(6.0 / (i + o))
        w[name] = rng.uniform(-lim, lim, (i, o)).astype(np.float32)
    w["density_w1"][:, 0] += 0.6
    rgba_rows = rng.integers(0, 256, (N_RAYS, 4), dtype=np.uint8)
    return dict(cam=cam, perm=perm, noises=noises, bg=bg, lv=lv, table=table, w=w, rgba_rows=rgba_rows)


def touched_table_entries(d):
    """Table rows the batch actually reads: the heaviest-weighted corner of the first marched sample on three levels."""
    from jaxngp_b200 import synthetic as S
    from oracle import hashgrid_np as H
    from oracle import oracle as O
    cam, perm = d["cam"], d["perm"].astype(np.int64)
    hw = cam["width"] * cam["height"]
    o, dd = S.pixel_rays(S.poses(N_VIEWS), perm // hw, perm % hw)
    ts, te = S.near_far(o, dd)
    out = O.march_rays(TOTAL_SAMPLES, 1024, 1, 128, 1.0, 0.0, o, dd, ts, te, d["noises"], S.occupancy_bitfield(), raw=True)
    xyzs, used = out[6], int(out[0][0]) - int(out[1][0])
    idx, wts = H.indices_and_weights(d["lv"], xyzs[used // 2:used // 2 + 1], 1.0)  # [L, 1, 8]
    picks = []
    for level in (1, 7, 13):
        c = int(np.argmax(wts[level, 0]))
        picks.append((int(idx[level, 0, c]), level % 2))
    return picks


def main():
    from jaxngp_b200 import synthetic as S
    from oracle import oracle as O
    from oracle import ref_shim
    O.build()
    d = make_inputs()
    cam = d["cam"]
    state = {}

    def loss_at(table, w):
        jr = ref_shim.ScriptedRandom([], [d["bg"], d["noises"]])
        if "ref" not in state:
            state["ref"] = ref_shim.install_train_forward(O, jr)
            state["jr"] = jr
            camera = state["ref"].make_camera(cam["width"], cam["height"], cam["fx"], cam["fy"], cam["cx"], cam["cy"])
            camera.near = cam["near"]
            state["camera"] = camera
        else:
            state["jr"]._uniforms[:] = [d["bg"], d["noises"]]

        class Rows:
            def __getitem__(self, idx):
                return d["rgba_rows"]

        loss, metrics = state["ref"].forward(d["perm"], S.poses(N_VIEWS), state["camera"], table, w, S.occupancy_bitfield(), Rows(),
                                             TOTAL_SAMPLES, N_VIEWS)
        return float(loss), metrics

    base, metrics = loss_at(d["table"], d["w"])
    print("loss", base, {k: int(np.asarray(v).sum()) for k, v in metrics.items() if k != "loss"})
    names, flat, steps, grads = [], [], [], []
    probes = list(PROBES) + [("table", row * 2 + f, TABLE_STEP) for row, f in touched_table_entries(d)[:1]]
    for name, k, h in probes:
        vals = []
        for sign in (+1, -1):
            table, w = d["table"], {n: v for n, v in d["w"].items()}
            if name == "table":
                table = d["table"].copy()
                table.reshape(-1)[k] += np.float32(sign * h)
            else:
                w[name] = d["w"][name].copy()
                w[name].reshape(-1)[k] += np.float32(sign * h)
            moved = (table if name == "table" else w[name]).reshape(-1)[k]
            vals.append((loss_at(table, w)[0], float(moved)))
        g = (vals[0][0] - vals[1][0]) / (vals[0][1] - vals[1][1])
        print(name, k, "dL/dp ~", g)
        names.append(name); flat.append(k); steps.append(h); grads.append(g)
    path = os.path.join(ROOT, "tests", "golden", "train_grad_reference.npz")
    np.savez_compressed(path, names=np.array(names), flat=np.array(flat, np.int64), steps=np.array(steps), fd=np.array(grads, np.float64),
                        loss=np.float64(base))
    print("wrote", path)


if __name__ == "__main__":
    main()
