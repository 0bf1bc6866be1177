"""TEST INFRASTRUCTURE ONLY -- golden vectors for the hash-grid encoder from the reference's OWN code.

Runs ``HashGridEncoder.__call__`` of /root/reference/models/encoders.py (unmodified; executed on numpy through
oracle/ref_shim.py, since jax/flax are absent here) on the seeded inputs of tests/inputs.py and writes
tests/golden/encoder_reference.npz: per configuration the query points and the reference's encodings.  The tables are
not stored (48 MB): tests regenerate them with ``inputs.encoder_table(rows, 2, amp=1.0)`` (PCG64, fixed seed); the
reference itself checks their shape -- its ``self.param(..., (offsets[-1], F), ...)`` request must match the rows the
oracle's level table predicts, which pins the table geometry (6,098,120 rows at C2, 5,592,320 at C1) as well.

    python oracle/make_golden_encoder.py        # needs /root/reference; run in the build container only
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = (  # (dim, T, N_max): C2 / C4 shape, C1 (imagefit) shape, a small table that hashes from level 3 on
    (3, 2 ** 19, 2048),
    (2, 2 ** 19, 2 ** 19),
    (3, 2 ** 14, 512),
    (2, 2 ** 20, 2 ** 19),  # the ImageFitter's own encoder (models/imagefit.py:28-37): ~1 Mi entries per level
)
N_POINTS = 256
PARAM = "latent codes stored on grid vertices"  # models/encoders.py:106


def edge_points(dim):
    return np.array([[-1.0] * dim, [1.0] * dim, [0.0] * dim, [0.999999] * dim, [0.96] * dim, [-0.5] * dim,
                     [0.933334] * dim], np.float32)


def main():
    from oracle import hashgrid_np as H
    from oracle import ref_shim
    from tests import inputs
    ref = ref_shim.install()
    out = {}
    for dim, T, N_max in CONFIGS:
        lv = H.level_table(16, T, 2, 16, N_max, dim)
        rows = int(lv["offsets"][-1])
        pts = inputs.encoder_points(N_POINTS, dim)
        pts[:7] = edge_points(dim)
        table = inputs.encoder_table(rows, 2, amp=1.0)
        enc_mod = ref.HashGridEncoder(L=16, T=T, F=2, N_min=16, N_max=N_max, tv_scale=0.0)
        enc_mod.bind_params(**{PARAM: table})  # the reference validates (offsets[-1], F) against this shape
        enc, tv = enc_mod(pts, 1.0)
        assert enc.dtype == np.float32 and enc.shape == (N_POINTS, 32) and tv == 0
        key = f"d{dim}_T{T}_N{N_max}"
        out[key + "_pts"] = pts
        out[key + "_enc"] = enc
        out[key + "_rows"] = np.int64(rows)
        out[key + "_b"] = np.float64(enc_mod.b)
        print(key, "rows", rows, "b", enc_mod.b, "|enc| max", float(np.abs(enc).max()))
    # the total-variation branch (models/encoders.py:234-254, tv_scale > 0): the regulariser's value on 64 points
    for dim, T, N_max in ((3, 2 ** 14, 512), (2, 2 ** 12, 256)):
        lv = H.level_table(16, T, 2, 16, N_max, dim)
        rows = int(lv["offsets"][-1])
        pts = inputs.encoder_points(64, dim)
        table = inputs.encoder_table(rows, 2, amp=1.0)
        enc_mod = ref.HashGridEncoder(L=16, T=T, F=2, N_min=16, N_max=N_max, tv_scale=0.25)
        enc_mod.bind_params(**{PARAM: table})
        _, tv = enc_mod(pts, 1.0)
        out[f"tv_d{dim}_T{T}_N{N_max}"] = np.float32(tv)
        print("tv", dim, T, N_max, float(tv))
    # TCNNHashGridEncoder.__call__ (models/encoders.py:259-305), the Python half of the tiny-cuda-nn path: what it hands
    # to jaxtcnn.hashgrid_encode (level offsets WITHOUT the 8-alignment of the pure-JAX encoder, the per-level scale, the
    # transposed unit-cube coordinates, the parameter shape it requests).  The CUDA half is tiny-cuda-nn v1.6 (absent).
    import collections
    import sys as _sys
    calls = []
    tcnn = _sys.modules["jaxtcnn"]
    tcnn.HashGridMetadata = collections.namedtuple("HashGridMetadata", "L F N_min per_level_scale")

    def record(desc, offset_table_data, coords_rm, params):
        calls.append((desc, np.asarray(offset_table_data), np.asarray(coords_rm), tuple(params.shape)))
        return np.zeros((desc.L * desc.F, coords_rm.shape[1]), np.float32)

    ref.hashgrid_encode, ref.HashGridMetadata = record, tcnn.HashGridMetadata
    for T, N_max in ((2 ** 19, 2048), (2 ** 14, 512)):
        lv = H.level_table(16, T, 2, 16, N_max, 3, align=1)
        rows = int(lv["offsets"][-1])
        pts = inputs.encoder_points(16, 3)
        mod = ref.TCNNHashGridEncoder(L=16, T=T, F=2, N_min=16, N_max=N_max, tv_scale=0.0)
        mod.bind_params(**{PARAM: np.zeros((rows, 2), np.float32)})  # shape-checked by the reference's self.param
        enc, tv = mod(pts, 1.0)
        desc, offs, coords, pshape = calls.pop()
        assert enc.shape == (16, 32) and offs.dtype == np.uint32 and coords.shape == (3, 16)
        key = f"tcnn_T{T}_N{N_max}"
        out[key + "_offsets"] = offs
        out[key + "_desc"] = np.array([desc.L, desc.F, desc.N_min, desc.per_level_scale], np.float64)
        out[key + "_coords_rm"] = coords
        print(key, "rows", int(offs[-1]), "per_level_scale", desc.per_level_scale)
    path = os.path.join(ROOT, "tests", "golden", "encoder_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
