"""TEST INFRASTRUCTURE ONLY -- golden vectors for the one-time camera-visibility culling of the density grid, from the
reference's OWN ``NeRFState.mark_untrained_density_grid`` (utils/types.py:1241-1362), extracted by AST and run
unmodified on numpy through oracle/ref_shim.py (morton3d_invert / packbits served by the C oracle).

Two cascades at bound 2 seen by three of the synthetic poses through the NeRF-synthetic camera, so that a good part
of the outer cascade is outside every frustum.  Writes tests/golden/mark_untrained_reference.npz.

    python oracle/make_golden_mark_untrained.py        # needs /root/reference
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G, K, BOUND, STEPS, VIEWS = 32, 2, 2.0, 1024, (0, 7, 31)


def main():
    from jaxngp_b200 import synthetic as S
    from oracle import oracle as O
    from oracle import ref_shim
    O.build()
    jran = ref_shim.ScriptedRandom([], [])
    Grid, RefState = ref_shim.install_grid_update(O, jran, extra_methods={"mark_untrained_density_grid"})
    cam = S.camera()
    poses = S.poses(100)[list(VIEWS)]  # [V, 12]: rot_cw row-major, then t_cw
    frames = []
    for tf in poses:
        m = np.concatenate([tf[:9].reshape(3, 3), tf[9:].reshape(3, 1)], axis=1).astype(np.float32)
        frames.append(types.SimpleNamespace(transform_matrix_jax_array=ref_shim._j(m)))
    ogrid = Grid.create(cascades=K, grid_resolution=G)
    rng = np.random.Generator(np.random.PCG64(5))
    ogrid = ogrid.replace(density=ref_shim._j(rng.uniform(0, 4, K * G ** 3).astype(np.float32)))
    state = RefState(ogrid, None, G, STEPS, BOUND, K)
    state.scene_meta.camera = types.SimpleNamespace(has_distortion=False, fx=np.float32(cam["fx"]), fy=np.float32(cam["fy"]),
                                                    cx=np.float32(cam["cx"]), cy=np.float32(cam["cy"]), width=cam["width"],
                                                    height=cam["height"])
    state.scene_meta.frames = frames
    out = dict(G=np.int64(G), K=np.int64(K), bound=np.float64(BOUND), views=np.array(VIEWS, np.int64), poses=poses.astype(np.float32),
               density_in=np.asarray(ogrid.density).copy())
    for step in (0, 300):  # step 0: threshold -0.5 (every trainable cell occupied); later: min(step-size threshold, mean density)
        state.step = step
        marked = state.mark_untrained_density_grid()
        g = marked.ogrid
        tag = f"step{step}"
        out[tag + "_density"] = np.asarray(g.density).copy()
        out[tag + "_occ_mask"] = np.asarray(g.occ_mask).copy()
        out[tag + "_occupancy"] = np.asarray(g.occupancy).copy()
        out[tag + "_alive_indices"] = np.asarray(g.alive_indices).astype(np.uint32)
        out[tag + "_alive_indices_offset"] = np.asarray(g.alive_indices_offset, np.int64)
        print(tag, "alive per cascade", np.diff(out[tag + "_alive_indices_offset"]), "of", G ** 3, "occupied", int(out[tag + "_occ_mask"].sum()))
    assert 0 < out["step0_alive_indices"].size < K * G ** 3
    path = os.path.join(ROOT, "tests", "golden", "mark_untrained_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
