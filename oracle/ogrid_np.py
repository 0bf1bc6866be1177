"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the density-grid update
(utils/types.py:1149-1239) with the random draws supplied as inputs.  Pinned: the reference's own methods run
unmodified on numpy stand-ins (oracle/ref_shim.py, oracle/make_golden_ogrid.py -> tests/golden/ogrid_reference.npz)
and tests/test_oracle_golden.py requires this file to reproduce their densities, masks and bitfields exactly."""
import numpy as np

from . import oracle as O


def sample_positions(idx, uniforms, G, cas, bound):
    c = O.morton3d_invert(np.asarray(idx, np.uint32)).astype(np.float32)  # :1193
    c = c / np.float32(G - 1) * np.float32(2) - np.float32(1)             # :1194
    mip_bound = np.float32(min(bound, 2 ** cas))
    half = mip_bound / np.float32(G)
    c = c * (mip_bound - half)                                              # :1197
    jitter = np.maximum(-half, np.asarray(uniforms, np.float32) * (half + half) + (-half))  # jax.random.uniform
    return (c + jitter).astype(np.float32)                                  # :1199-1206


def decay_and_max(density, idx, new_density, decay=0.95):
    d = np.asarray(density, np.float32).copy()
    alive = d >= 0
    d[alive] = d[alive] * np.float32(decay)                                 # :1162-1164
    np.maximum.at(d, np.asarray(idx, np.int64), np.asarray(new_density, np.float32))  # :1219-1221 as a true max (Q14)
    return d


def threshold(density_cascade0, thr_max):
    d = np.asarray(density_cascade0, np.float32)
    return np.float32(min(np.float32(thr_max), d[d >= 0].astype(np.float64).mean()))  # :1229-1230


_CELL_CORNERS = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]], np.float32)


def visible_cells(K, G, bound, poses, cam):
    """``mark_untrained_density_grid`` (utils/types.py:1241-1326) for an undistorted camera: a cell is trainable iff one
    of its 8 corners lies in front of some training camera and projects inside its frame.  ``poses`` [V, 12] = rot_cw
    (row-major) then t_cw.  Pinned by tests/golden/mark_untrained_reference.npz (the reference's own method)."""
    G3 = G ** 3
    alive = np.zeros(K * G3, bool)
    cells = O.morton3d_invert(np.arange(G3, dtype=np.uint32)).astype(np.float32)
    for cas in range(K):
        mip_bound = np.float32(min(2 ** cas, bound))
        width = np.float32(2) * mip_bound / np.float32(G)
        xyz = (cells / np.float32(G) - np.float32(0.5)) * (np.float32(2) * mip_bound)                # :1249-1252
        verts = xyz[:, None, :] + width * _CELL_CORNERS[None]                                         # :1253-1263
        part = np.zeros(G3, bool)
        for tf in np.asarray(poses, np.float32):
            rot, t = tf[:9].reshape(3, 3), tf[9:]
            p_cam = ((verts - t)[..., None, :] * rot.T).sum(-1)                                        # :1275-1276
            front = p_cam[..., 2] < 0
            uv = p_cam[..., :2] / (-p_cam[..., 2:])
            uv = (uv * np.array([cam["fx"], cam["fy"]], np.float32) + np.array([cam["cx"], cam["cy"]], np.float32)) \
                / np.array([cam["width"], cam["height"]], np.float32)                                  # :1298-1302
            inside = ((uv >= 0) & (uv < 1)).all(-1)
            part |= (front & inside).any(-1)                                                            # :1310-1312
        alive[cas * G3:(cas + 1) * G3] = part
    return alive


def mark_untrained(density, alive, step, thr_max, G3):
    """utils/types.py:1341-1362: culled cells get density -1; re-threshold with -0.5 at step 0, else with the usual rule
    evaluated on the state BEFORE the culling (:1343 reads self.ogrid, whose alive set is the previous one)."""
    marked = np.where(alive, np.asarray(density, np.float32), np.float32(-1))
    thr = np.float32(-0.5) if step == 0 else threshold(np.asarray(density, np.float32)[:G3], thr_max)
    mask, bits = O.packbits(float(thr), marked)
    return marked, np.asarray(mask).astype(bool), np.asarray(bits), np.nonzero(alive)[0].astype(np.uint32)
