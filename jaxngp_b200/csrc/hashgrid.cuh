// hashgrid.cuh -- device code of the hash-grid gather shared by hashgrid.cu (stand-alone encoder ops) and
// mlp.cu (encoder fused in front of the NeRF MLP).  Follows models/encoders.py:82-233.
#pragma once
#include "common.cuh"

namespace ngp {
namespace hg {

constexpr uint32_t kPrime1 = 2654435761u, kPrime2 = 805459861u;  // encoders.py:169

struct LevelMeta {
    float scale;
    uint32_t res, offset, wrap, hashed;
    uint32_t last_row;  // rows of the whole table - 1: where an out-of-range row of a dense level lands (see grid_row)
};

// kPow2: every level's wrap is a power of two (always true for `mod T`, encoders.py:187): mask, no branch.
// Row as the reference computes it -- possibly past the table, see grid_row below.
template <int DIM, bool kPow2 = false>
__device__ __forceinline__ uint32_t grid_row_unclamped(const uint32_t (&v)[DIM], const LevelMeta &m) {
    uint32_t idx;
    if (m.hashed) {  // encoders.py:157-177
        idx = v[0] ^ (v[1] * kPrime1);
        if (DIM == 3) idx ^= v[2] * kPrime2;
    } else {  // encoders.py:134-155 (uint32 wrap-around arithmetic)
        idx = v[0] + v[1] * m.res;
        if (DIM == 3) idx += v[2] * m.res * m.res;
    }
    // `mod wrap` (encoders.py:187); power-of-two wraps are a mask, others rarely exceed the range
    if (kPow2 || (m.wrap & (m.wrap - 1u)) == 0u) idx &= m.wrap - 1u;
    else if (idx >= m.wrap) idx %= m.wrap;
    return idx + m.offset;
}

// Q1: with `mod T` a dense level's vertex index reaches past the level's own rows (a vertex coordinate equals `res` in
// the outer half-cell) and lands in the NEXT level's rows -- reproduced.  On the last dense level of a table with no
// hashed level behind it the row would lie past the table.  `latents[indices]` (encoders.py:225) is NumPy-style jnp
// indexing: XLA clamps an out-of-range gather index to the last row, and DROPS the out-of-range update of the
// transposed scatter-add (jax's documented out-of-bounds semantics for indexing; jax is not on disk to confirm, and the
// reference's default geometry -- hashed levels behind the dense ones -- never gets here).  Forward: this clamp;
// backward: the contribution is skipped (hashgrid_a1_backward_kernel).
template <int DIM, bool kPow2 = false>
__device__ __forceinline__ uint32_t grid_row(const uint32_t (&v)[DIM], const LevelMeta &m) {
    return min(grid_row_unclamped<DIM, kPow2>(v, m), m.last_row);
}

template <typename TT, int F>
struct RowIO;
template <>
struct RowIO<float, 2> {
    static __device__ __forceinline__ void load(const float *t, uint32_t row, float (&f)[2]) {
        float2 v = __ldg(reinterpret_cast<const float2 *>(t) + row);
        f[0] = v.x; f[1] = v.y;
    }
    // rows 2k and 2k+1 with one aligned 16-byte load
    static __device__ __forceinline__ void load_pair(const float *t, uint32_t even_row, float (&lo)[2], float (&hi)[2]) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(t) + (even_row >> 1));
        lo[0] = v.x; lo[1] = v.y; hi[0] = v.z; hi[1] = v.w;
    }
};
template <>
struct RowIO<float, 4> {
    static __device__ __forceinline__ void load(const float *t, uint32_t row, float (&f)[4]) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(t) + row);
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
    static __device__ __forceinline__ void load_pair(const float *t, uint32_t even_row, float (&lo)[4], float (&hi)[4]) {
        load(t, even_row, lo);
        load(t, even_row + 1u, hi);
    }
};
template <>
struct RowIO<__half, 2> {
    static __device__ __forceinline__ void load(const __half *t, uint32_t row, float (&f)[2]) {
        __half2 h = __ldg(reinterpret_cast<const __half2 *>(t) + row);
        float2 v = __half22float2(h);
        f[0] = v.x; f[1] = v.y;
    }
    static __device__ __forceinline__ void load_pair(const __half *t, uint32_t even_row, float (&lo)[2], float (&hi)[2]) {
        uint2 raw = __ldg(reinterpret_cast<const uint2 *>(t) + (even_row >> 1));
        float2 a = __half22float2(*reinterpret_cast<__half2 *>(&raw.x));
        float2 b = __half22float2(*reinterpret_cast<__half2 *>(&raw.y));
        lo[0] = a.x; lo[1] = a.y; hi[0] = b.x; hi[1] = b.y;
    }
};
template <>
struct RowIO<__half, 4> {
    static __device__ __forceinline__ void load(const __half *t, uint32_t row, float (&f)[4]) {
        uint2 raw = __ldg(reinterpret_cast<const uint2 *>(t) + row);
        float2 a = __half22float2(*reinterpret_cast<__half2 *>(&raw.x));
        float2 b = __half22float2(*reinterpret_cast<__half2 *>(&raw.y));
        f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
    }
    static __device__ __forceinline__ void load_pair(const __half *t, uint32_t even_row, float (&lo)[4], float (&hi)[4]) {
        uint4 raw = __ldg(reinterpret_cast<const uint4 *>(t) + (even_row >> 1));
        float2 a = __half22float2(*reinterpret_cast<__half2 *>(&raw.x));
        float2 b = __half22float2(*reinterpret_cast<__half2 *>(&raw.y));
        float2 c = __half22float2(*reinterpret_cast<__half2 *>(&raw.z));
        float2 d = __half22float2(*reinterpret_cast<__half2 *>(&raw.w));
        lo[0] = a.x; lo[1] = a.y; lo[2] = b.x; lo[3] = b.y;
        hi[0] = c.x; hi[1] = c.y; hi[2] = d.x; hi[3] = d.y;
    }
};

// predicated row load (forced @p LDG: a branch here would split the batch of independent gathers)
template <typename TT, int F>
__device__ __forceinline__ void load_row_if(const TT *t, uint32_t row, bool pred, float (&f)[F]);
template <>
__device__ __forceinline__ void load_row_if<float, 2>(const float *t, uint32_t row, bool pred, float (&f)[2]) {
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p ld.global.nc.v2.f32 {%0, %1}, [%2]; }"
                 : "+f"(f[0]), "+f"(f[1]) : "l"(reinterpret_cast<const float2 *>(t) + row), "r"((uint32_t)pred));
}
template <>
__device__ __forceinline__ void load_row_if<float, 4>(const float *t, uint32_t row, bool pred, float (&f)[4]) {
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %5, 0; @p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4]; }"
                 : "+f"(f[0]), "+f"(f[1]), "+f"(f[2]), "+f"(f[3])
                 : "l"(reinterpret_cast<const float4 *>(t) + row), "r"((uint32_t)pred));
}
template <>
__device__ __forceinline__ void load_row_if<__half, 2>(const __half *t, uint32_t row, bool pred, float (&f)[2]) {
    uint32_t raw = 0;
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p ld.global.nc.b32 %0, [%1]; }"
                 : "+r"(raw) : "l"(reinterpret_cast<const uint32_t *>(t) + row), "r"((uint32_t)pred));
    float2 v = __half22float2(*reinterpret_cast<__half2 *>(&raw));
    f[0] = pred ? v.x : f[0]; f[1] = pred ? v.y : f[1];
}
template <>
__device__ __forceinline__ void load_row_if<__half, 4>(const __half *t, uint32_t row, bool pred, float (&f)[4]) {
    uint32_t r0 = 0, r1 = 0;
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p ld.global.nc.v2.b32 {%0, %1}, [%2]; }"
                 : "+r"(r0), "+r"(r1) : "l"(reinterpret_cast<const uint2 *>(t) + row), "r"((uint32_t)pred));
    float2 a = __half22float2(*reinterpret_cast<__half2 *>(&r0));
    float2 b = __half22float2(*reinterpret_cast<__half2 *>(&r1));
    f[0] = pred ? a.x : f[0]; f[1] = pred ? a.y : f[1]; f[2] = pred ? b.x : f[2]; f[3] = pred ? b.y : f[3];
}

__device__ __forceinline__ LevelMeta a1_level(const NgpHashGridA1Descriptor &d, uint32_t level) {
    LevelMeta m;
    m.scale = d.scales[level];
    m.res = d.res[level];
    m.offset = d.offsets[level];
    m.wrap = d.wrap_T ? d.wrap_T : d.offsets[level + 1] - d.offsets[level];
    m.hashed = (d.hashed_mask >> level) & 1u;
    m.last_row = d.offsets[d.L] - 1u;
    return m;
}

// pos01 = (pos + bound) / (2 bound) (encoders.py:87), level independent.  For a power-of-two divisor the
// multiplication by its exact reciprocal rounds identically to the IEEE division (warp-uniform branch).
template <int DIM>
__device__ __forceinline__ void unit_pos(const float (&x)[DIM], float bound, float (&p01)[DIM]) {
    const float two_b = __fmul_rn(2.f, bound);
    const uint32_t tb_bits = __float_as_uint(two_b);
    if ((tb_bits & 0x007FFFFFu) == 0u && tb_bits > 0x00800000u && tb_bits < 0x7E800000u) {
        const float inv_two_b = __uint_as_float(0x7F000000u - tb_bits);
#pragma unroll
        for (int k = 0; k < DIM; ++k) p01[k] = __fmul_rn(__fadd_rn(x[k], bound), inv_two_b);
    } else {
#pragma unroll
        for (int k = 0; k < DIM; ++k) p01[k] = __fdiv_rn(__fadd_rn(x[k], bound), two_b);
    }
}

// cell base and fractional offsets of a point at one level (encoders.py:116-123,204,216-218)
template <int DIM>
__device__ __forceinline__ void a1_cell(const float (&p01)[DIM], float scale, uint32_t (&base)[DIM], float (&fr)[DIM]) {
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        float ps = __fadd_rn(__fmul_rn(p01[k], scale), .5f);  // mul then add, as the XLA-CPU reference path
        float fl = floorf(ps);
        base[k] = (uint32_t)(int)fl;
        fr[k] = ps - fl;
    }
}

// enc[level*F .. level*F+F) of one point: sum over the 2^DIM cell corners of w_c * table[row_c] (encoders.py:204-233).
// Corner c has bit (DIM-1-k) of c on axis k (last axis fastest, encoders.py:16-33), so corners c and c + NC/2 differ
// by +1 along x.  Their rows are neighbours (2k, 2k+1) whenever the x base is even on a hashed level
// (h(x|1) = h(x)^1) or the flat index is even on a dense one: then ONE aligned 2-row load serves both corners
// (kPaired; needs the table base aligned to two rows), the odd case issues the second row's load predicated --
// 6 instead of 8 L1 wavefronts per lane on average, the limiter of this gather.
template <int DIM, int F, typename TT, bool kPaired, bool kPow2 = false>
__device__ __forceinline__ void encode_point_level(const TT *__restrict__ table, const LevelMeta &m,
                                                   const float (&p01)[DIM], float (&acc)[F]) {
    constexpr int NC = 1 << DIM, H = NC / 2;
    uint32_t base[DIM];
    float fr[DIM];
    a1_cell<DIM>(p01, m.scale, base, fr);
    uint32_t rows[NC];
    float w[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        uint32_t v[DIM];
        float wc = 1.f;
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            const uint32_t bit = (c >> (DIM - 1 - k)) & 1;
            v[k] = base[k] + bit;
            wc *= bit ? fr[k] : 1.f - fr[k];  // encoders.py:204-213 (clip is a no-op for frac in [0,1))
        }
        rows[c] = grid_row<DIM, kPow2>(v, m);
        w[c] = wc;
    }
    float vals[NC][F];
    if (kPaired) {
#pragma unroll
        for (int c = 0; c < H; ++c) {
            const uint32_t ra = rows[c], rb = rows[c + H];
            const bool paired = (ra ^ rb) == 1u;
            float lo[F], hi[F];
            RowIO<TT, F>::load_pair(table, ra & ~1u, lo, hi);
            const bool a_hi = ra & 1u;
#pragma unroll
            for (int f = 0; f < F; ++f) {
                vals[c][f] = a_hi ? hi[f] : lo[f];
                vals[c + H][f] = a_hi ? lo[f] : hi[f];
            }
            load_row_if<TT, F>(table, rb, !paired, vals[c + H]);
        }
    } else {
#pragma unroll
        for (int c = 0; c < NC; ++c) RowIO<TT, F>::load(table, rows[c], vals[c]);
    }
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] = fmaf(w[c], vals[c][f], acc[f]);
}

// predicated two-row load (rows 2k, 2k+1), F = 2
template <typename TT>
__device__ __forceinline__ void load_pair_if(const TT *t, uint32_t even_row, bool pred, float (&lo)[2], float (&hi)[2]);
template <>
__device__ __forceinline__ void load_pair_if<float>(const float *t, uint32_t even_row, bool pred, float (&lo)[2], float (&hi)[2]) {
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %5, 0; @p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4]; }"
                 : "+f"(lo[0]), "+f"(lo[1]), "+f"(hi[0]), "+f"(hi[1])
                 : "l"(reinterpret_cast<const float4 *>(t) + (even_row >> 1)), "r"((uint32_t)pred));
}
template <>
__device__ __forceinline__ void load_pair_if<__half>(const __half *t, uint32_t even_row, bool pred, float (&lo)[2], float (&hi)[2]) {
    uint32_t r0 = 0, r1 = 0;
    asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p ld.global.nc.v2.b32 {%0, %1}, [%2]; }"
                 : "+r"(r0), "+r"(r1) : "l"(reinterpret_cast<const uint2 *>(t) + (even_row >> 1)), "r"((uint32_t)pred));
    const float2 a = __half22float2(*reinterpret_cast<__half2 *>(&r0));
    const float2 b = __half22float2(*reinterpret_cast<__half2 *>(&r1));
    lo[0] = a.x; lo[1] = a.y; hi[0] = b.x; hi[1] = b.y;
}

// encode_point_level for dim = 3, F = 2, power-of-two wrap, paired loads, with every load predicated on `active`
// (inactive lanes get zeros and touch no memory): branch-free, so the gathers of several (point, level) pairs
// of one thread can be in flight together.  Same arithmetic, same summation order => same bits.
template <typename TT>
__device__ __forceinline__ void encode_point_level_pred(const TT *__restrict__ table, const LevelMeta &m,
                                                        const float (&p01)[3], bool active, float (&acc)[2]) {
    uint32_t base[3];
    float fr[3];
    a1_cell<3>(p01, m.scale, base, fr);
    uint32_t rows[8];
    float w[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t v[3];
        float wc = 1.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t bit = (c >> (2 - k)) & 1;
            v[k] = base[k] + bit;
            wc *= bit ? fr[k] : 1.f - fr[k];
        }
        rows[c] = grid_row<3, true>(v, m);
        w[c] = wc;
    }
    float vals[8][2];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t ra = rows[c], rb = rows[c + 4];
        float lo[2] = {0.f, 0.f}, hi[2] = {0.f, 0.f};
        load_pair_if<TT>(table, ra & ~1u, active, lo, hi);
        const bool a_hi = ra & 1u;
        vals[c][0] = a_hi ? hi[0] : lo[0];
        vals[c][1] = a_hi ? hi[1] : lo[1];
        vals[c + 4][0] = a_hi ? lo[0] : hi[0];
        vals[c + 4][1] = a_hi ? lo[1] : hi[1];
        load_row_if<TT, 2>(table, rb, active && (ra ^ rb) != 1u, vals[c + 4]);
    }
    acc[0] = acc[1] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        acc[0] = fmaf(w[c], vals[c][0], acc[0]);
        acc[1] = fmaf(w[c], vals[c][1], acc[1]);
    }
}

}  // namespace hg
}  // namespace ngp
