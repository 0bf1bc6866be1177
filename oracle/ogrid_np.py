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
