"""TEST INFRASTRUCTURE ONLY -- golden vectors for the density-grid update from the reference's OWN code.

Executes ``OccupancyDensityGrid`` (utils/types.py:93-144) and ``NeRFState.update_ogrid_density`` /
``threshold_ogrid`` (utils/types.py:1149-1239), unmodified, on numpy through oracle/ref_shim.py, with the random draws
scripted (the oracle takes them as inputs) and an analytic density function standing in for the NeRF.  Cells are drawn
without repetition: with repeated cells the reference's ``.at[idx].set(maximum(...))`` keeps an arbitrary writer
(SURVEY Q14), whereas this repo takes the true maximum.  Writes tests/golden/ogrid_reference.npz.

    python oracle/make_golden_ogrid.py        # needs /root/reference; run in the build container only
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

G, K, BOUND, STEPS = 16, 2, 2.0, 1024


def density_fn(xyz):
    """Smooth positive field oscillating around 2 (below the step-size threshold 2.956, so that the mean branch of
    threshold_ogrid decides), about half of the cells of every cascade above its mean."""
    xyz = np.asarray(xyz, np.float32)
    return (np.float32(2.0) + np.float32(1.5) * np.sin(np.float32(5.0) * xyz[:, 0]) * np.cos(np.float32(4.0) * xyz[:, 1])
            * np.sin(np.float32(3.0) * xyz[:, 2] + np.float32(1.0))).astype(np.float32)


def main():
    from oracle import oracle as O
    from oracle import ref_shim
    O.build()
    rng = np.random.Generator(np.random.PCG64(2024))
    G3 = G ** 3
    out = dict(G=np.int64(G), K=np.int64(K), bound=np.float64(BOUND), steps=np.int64(STEPS))
    script_choices, script_uniforms = [], []
    jran = ref_shim.ScriptedRandom(script_choices, script_uniforms)
    Grid, RefState = ref_shim.install_grid_update(O, jran)
    ogrid = Grid.create(cascades=K, grid_resolution=G)
    state = RefState(ogrid, lambda params, xyz, d, a: (density_fn(xyz)[:, None], None), G, STEPS, BOUND, K)
    # pass 1: update_all (every cell), pass 2: sampled cells (M/2 uniform + M/2 among the occupied), thresholding after each
    for tag, update_all in (("all", True), ("sampled", False)):
        for cas in range(K):
            half = max(1, max(1, G3 // 2) // 2)
            if update_all:
                n_upd = G3
            else:
                first = rng.permutation(G3)[:half].astype(np.uint32)
                occ = np.nonzero(np.asarray(state.ogrid.occ_mask[cas * G3:(cas + 1) * G3]))[0]
                rest = np.setdiff1d(occ, first)
                assert rest.size >= half, "not enough occupied cells for a repetition-free draw"
                second = rng.permutation(rest)[:half].astype(np.uint32)
                jran._choices += [first, second]
                out[f"{tag}_c{cas}_first"], out[f"{tag}_c{cas}_second"] = first, second
                n_upd = 2 * half
            jitter = rng.random((n_upd, 3), dtype=np.float32)
            jran._uniforms.append(jitter)
            out[f"{tag}_c{cas}_jitter"] = jitter
            state = state.update_ogrid_density(None, cas, update_all, max_inference=4096)
        out[f"{tag}_density"] = np.asarray(state.ogrid.density).copy()
        state = state.threshold_ogrid()
        out[f"{tag}_threshold"] = np.float32(min(state.density_threshold_from_min_step_size,
                                                 float(state.ogrid.mean_density_up_to_cascade(1))))
        out[f"{tag}_occ_mask"] = np.asarray(state.ogrid.occ_mask).copy()
        out[f"{tag}_occupancy"] = np.asarray(state.ogrid.occupancy).copy()
        print(tag, "occupied fraction", float(out[f"{tag}_occ_mask"].mean()), "threshold", float(out[f"{tag}_threshold"]))
    assert not jran._choices and not jran._uniforms
    # the update cadence of the training loop (utils/types.py:1380-1396), from the reference's own properties
    cadence = ref_shim.install_cadence()
    steps = np.arange(0, 600, dtype=np.int64)
    out["cadence_steps"] = steps
    out["cadence_interval"] = np.array([cadence(int(s)).update_ogrid_interval for s in steps], np.int64)
    out["cadence_call"] = np.array([bool(cadence(int(s)).should_call_update_ogrid) for s in steps])
    out["cadence_all"] = np.array([bool(cadence(int(s)).should_update_all_ogrid_cells) for s in steps])
    path = os.path.join(ROOT, "tests", "golden", "ogrid_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
