"""TEST INFRASTRUCTURE ONLY -- independent numpy restatement of the reference's pure-JAX
HashGridEncoder (models/encoders.py:58-256).  Pinned: jax is absent from this image, but the reference module
runs unmodified on numpy stand-ins (oracle/ref_shim.py); its outputs are committed as
tests/golden/encoder_reference.npz and tests/test_oracle_golden.py checks this file (<= 2e-6) and
oracle/ngp_oracle.c (exact) against them.
"""
import math

import numpy as np


def next_multiple(value, multiple):  # utils/common.py:336-337
    return ((value + multiple - 1) // multiple) * multiple


def level_table(L, T, F, N_min, N_max, dim, align=8):
    """encoders.py:76-80 (b), :89-103 (levels).  align=8 -> HashGridEncoder (:96);
    align=1 -> TCNNHashGridEncoder (:275)."""
    b = math.exp((math.log(N_max) - math.log(N_min)) / (L - 1))
    scales, res, hashed, offsets = [], [], [], [0]
    for i in range(L):
        scale = N_min * (b ** i) - 1
        scales.append(scale)
        r = math.ceil(scale) + 1
        res.append(r)
        n_entries = next_multiple(r ** dim, align)
        if n_entries <= T:
            hashed.append(0)
        else:
            n_entries = T
            hashed.append(1)
        offsets.append(offsets[-1] + n_entries)
    # encoders.py:98-101: first_hash_level counts levels that fit; dense levels are the FIRST
    # `first_hash_level` levels (:181-186), i.e. a prefix -- which `hashed` above already is
    # because level sizes are monotone.
    return dict(L=L, T=T, F=F, b=b, dim=dim,
                scales=np.asarray(scales, np.float32),  # :216 cast to f32
                res=np.asarray(res, np.uint32),
                hashed=np.asarray(hashed, np.uint8),
                offsets=np.asarray(offsets, np.uint32))


_VERTS = {  # encoders.py:16-33
    2: np.array([[0, 0], [0, 1], [1, 0], [1, 1]], np.float32),
    3: np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1],
                 [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]], np.float32),
}
_PRIMES = np.array([1, 2654435761, 805459861], np.uint32)  # encoders.py:169


def indices_and_weights(levels, pos, bound):
    """Returns (indices [L, n, 2^dim] u32 incl. offsets, weights [L, n, 2^dim] f32)."""
    pos = np.asarray(pos, np.float32)
    n, dim = pos.shape
    L, T = levels["L"], levels["T"]
    p01 = (pos + np.float32(bound)) / np.float32(2 * bound)  # :87
    ps = p01[None] * levels["scales"][:, None, None] + np.float32(0.5)  # :218
    fl = np.floor(ps)
    vert = (fl[:, :, None, :] + _VERTS[dim][None, None]).astype(np.int64).astype(np.uint32)  # :116-123
    idx = np.empty(vert.shape[:3], np.uint32)
    with np.errstate(over="ignore"):
        for l in range(L):
            v = vert[l]
            if levels["hashed"][l]:  # :157-177
                h = v[..., 0] ^ (v[..., 1] * _PRIMES[1])
                if dim == 3:
                    h = h ^ (v[..., 2] * _PRIMES[2])
            else:  # :134-155
                r = np.uint32(levels["res"][l])
                h = v[..., 0] + v[..., 1] * r
                if dim == 3:
                    h = h + v[..., 2] * (r * r)
            idx[l] = h % np.uint32(T) + levels["offsets"][l]  # :187-188
    fr = ps - fl  # jnp.modf fractional part, :204
    w = np.clip((1 - _VERTS[dim])[None, None] + (2 * _VERTS[dim] - 1)[None, None] * fr[:, :, None, :], 0, 1)
    return idx, np.prod(w, axis=-1).astype(np.float32)  # :205-213


def encode(levels, pos, bound, table):
    idx, w = indices_and_weights(levels, pos, bound)
    # :226 is NumPy-style jnp indexing: an out-of-range index (last level dense, outer half-cell) is CLAMPED by XLA
    lat = np.asarray(table, np.float32)[np.minimum(idx, np.uint32(levels["offsets"][-1] - 1))]  # [L, n, C, F]
    enc = (lat * w[..., None]).sum(axis=-2)  # :231
    return enc.transpose(1, 0, 2).reshape(enc.shape[1], -1)  # :233


def backward(levels, pos, bound, d_enc, F):
    idx, w = indices_and_weights(levels, pos, bound)
    L, n, Cn = idx.shape
    d = np.asarray(d_enc, np.float64).reshape(n, L, F).transpose(1, 0, 2)  # [L, n, F]
    upd = w[..., None].astype(np.float64) * d[:, :, None, :]
    rows = int(levels["offsets"][-1])
    out = np.zeros((rows, F), np.float64)
    flat_idx, flat_upd = idx.reshape(-1), upd.reshape(-1, F)
    inside = flat_idx < rows  # the transposed scatter-add of that gather DROPS out-of-range updates (jax's documented
    np.add.at(out, flat_idx[inside], flat_upd[inside])  # out-of-bounds semantics for indexing)
    return out
