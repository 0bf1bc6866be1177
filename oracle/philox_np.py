"""TEST INFRASTRUCTURE: numpy restatement of the Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw,
"Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 `philox.h`) and of the element -> counter mapping of
csrc/common.cuh (`philox_uniform4`).  The reference draws its random inputs with jax.random outside the compiled ops
(models/renderers/cuda.py:118-122, app/nerf/_utils.py:134-136, utils/types.py:1170-1206), so the numbers themselves are
not a parity target; this file pins the generator to its published known-answer vectors and lets the CPU checker
reproduce what the kernels drew.  Never imported by the product."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Counters c0..c3 and key k0, k1 as uint32 arrays (broadcastable) -> four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c1, c3, c0, c2 = p1 & MASK, p0 & MASK, n0, n2
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def bits_to_unit_float(bits):
    """jax.random.uniform's construction: 23 mantissa bits under exponent 0, minus 1 -> [0, 1)."""
    return ((np.asarray(bits, np.uint32) >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1)


def uniform4(n, counter, seed, stream_id):
    """f32 [n, 4]: what ngp_philox_uniform writes for (n, counter, seed, stream_id)."""
    seed = int(seed) & (2 ** 64 - 1)
    r = philox4x32_10(np.arange(n, dtype=np.uint32), np.uint32(counter), np.uint32(stream_id), np.uint32(0),
                      seed & 0xFFFFFFFF, seed >> 32)
    return np.stack([bits_to_unit_float(x) for x in r], axis=-1)


#: Random123 known-answer vectors (kat_vectors, philox4x32 10 rounds): (counter, key, expected)
KAT = (
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
)
