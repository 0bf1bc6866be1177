// hashgrid.cu -- multiresolution hash-grid encoder, forward gather and gradient scatter, sm_100a.
//
// Two front-ends over the same device code:
//   * ngp_hashgrid_a1_{forward,backward}: the pure-JAX HashGridEncoder.__call__ of
//     models/encoders.py:82-256 (row-major [n, L*F] output, host-computed f32 level scales,
//     `index mod T` on every level, corner order irrelevant to the sum), 2-D and 3-D.
//   * ngp_hashgrid_encode{,_backward}: the jax-tcnn custom calls (deps/jax-tcnn/lib/impl/
//     hashgrid.cu:20-174), i.e. tiny-cuda-nn v1.6 kernel_grid semantics (SoA [L*F, n] output,
//     f32-on-device scale = exp2f(l*log2f(b))*N_min-1, pos = fma(scale, x, .5), `index % level
//     size`, dy_dx side output).  tiny-cuda-nn is not vendored by the reference: parity unpinned.
//
// Mapping: one thread = one (point, level).  With L = 16 a warp covers 2 points x 16 levels, so the
// row-major output row of a point is written by 16 consecutive lanes (full 128 B lines), the
// position loads are broadcasts, and the 8 corner fetches of each lane are independent 8-byte
// (F=2 f32) / 4-byte (F=2 f16) vector loads issued back to back: 256 gathers in flight per warp.
// The whole table (48.8 MB at T=2^19) is L2-resident on B200 (126 MB), so the kernel is bound by
// L2 sector throughput, not HBM; see DESIGN.md section 5 for the roofline arithmetic.
//
// Backward: level-major warps with run aggregation (see the kernel), `red.global.add.v2.f32` (one
// 8-byte L2 reduction per corner instead of two scalar atomics), after a zero-fill of the table grad.
#include "common.cuh"
#include "hashgrid.cuh"

namespace ngp {
namespace {

constexpr int kBlock = 256;
using hg::LevelMeta;
using hg::RowIO;
using hg::a1_level;
using hg::grid_row;
using hg::kPrime1;
using hg::kPrime2;

template <int DIM, int F, typename TT, bool kPaired, bool kPow2>
__global__ void __launch_bounds__(kBlock) hashgrid_a1_forward_kernel(const __grid_constant__ NgpHashGridA1Descriptor d,
                                                                      const float *__restrict__ pos,
                                                                      const TT *__restrict__ table,
                                                                      const uint32_t *__restrict__ group_counts,
                                                                      float *__restrict__ enc) {
    __shared__ LevelMeta s_meta[NGP_HG_MAX_LEVELS];
    if (threadIdx.x < d.L) s_meta[threadIdx.x] = a1_level(d, threadIdx.x);
    __syncthreads();
    const uint64_t tid = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    uint32_t point, level;
    if ((d.L & (d.L - 1u)) == 0u) {  // L = 16: shift and mask instead of a 64-bit division
        point = (uint32_t)(tid >> (31 - __clz(d.L)));
        level = (uint32_t)tid & (d.L - 1u);
    } else {
        point = (uint32_t)(tid / d.L);
        level = (uint32_t)(tid % d.L);
    }
    if (point >= d.n_points) return;
    // grouped (inference) layout: rows past the group's sample count are padding, skip them
    if (d.rows_per_group && point % d.rows_per_group >= __ldg(group_counts + point / d.rows_per_group)) return;
    const LevelMeta m = s_meta[level];
    float x[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) x[k] = __ldg(pos + (size_t)point * DIM + k);
    float p01[DIM], acc[F];
    hg::unit_pos<DIM>(x, d.bound, p01);
    hg::encode_point_level<DIM, F, TT, kPaired, kPow2>(table, m, p01, acc);
    float *out = enc + ((size_t)point * d.L + level) * F;  // [n, L*F], level-major / feature-minor (:233)
    if (F == 2) *reinterpret_cast<float2 *>(out) = make_float2(acc[0], acc[1]);
    else *reinterpret_cast<float4 *>(out) = make_float4(acc[0], acc[1], acc[F - 2], acc[F - 1]);
}


// ---------------------------------------------------------------- level-major passes (tables larger than L2)
// When the whole table no longer fits in L2 (T >= 2^20 at L = 16, F = 2: 87 MB .. 1 GB), the point-major kernels above
// keep EVERY level's rows live at once and each 8-byte gather / reduction misses to DRAM as a 32-byte sector.  The
// grouped kernels run the same arithmetic level group by level group (grid.y = group, scheduled after grid.x): the rows
// in flight are those of `kLPG` consecutive levels only, sized by the launcher to stay L2-resident, so a level's rows
// are fetched from DRAM once per pass instead of once per access.  Same per-(point, level) code => same bits.
//
// Forward, kLPG >= 4: a point's kLPG * F floats are >= 32 contiguous bytes of its output row: full sectors.
// Forward, kLPG < 4 : partial sectors would be read-modify-written at DRAM on every pass; the pass writes a level-major
// scratch [L][n][F] instead (coalesced) and hashgrid_transpose_kernel lays the rows out afterwards.
template <int DIM, int F, typename TT, bool kPaired, bool kPow2, int kLPG, bool kScratch>
__global__ void __launch_bounds__(kBlock) hashgrid_a1_forward_group_kernel(const __grid_constant__ NgpHashGridA1Descriptor d,
                                                                            const float *__restrict__ pos,
                                                                            const TT *__restrict__ table,
                                                                            float *__restrict__ out) {
    __shared__ LevelMeta s_meta[kLPG];
    const uint32_t l0 = blockIdx.y * kLPG;
    if (threadIdx.x < kLPG && l0 + threadIdx.x < d.L) s_meta[threadIdx.x] = a1_level(d, l0 + threadIdx.x);
    __syncthreads();
    const uint64_t tid = (uint64_t)blockIdx.x * kBlock + threadIdx.x;
    const uint32_t point = (uint32_t)(tid / kLPG), j = (uint32_t)(tid % kLPG), level = l0 + j;
    if (point >= d.n_points || level >= d.L) return;
    const LevelMeta m = s_meta[j];
    float x[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) x[k] = __ldg(pos + (size_t)point * DIM + k);
    float p01[DIM], acc[F];
    hg::unit_pos<DIM>(x, d.bound, p01);
    hg::encode_point_level<DIM, F, TT, kPaired, kPow2>(table, m, p01, acc);
    float *dst = kScratch ? out + ((size_t)level * d.n_points + point) * F : out + ((size_t)point * d.L + level) * F;
    // streaming stores (evict-first): the L2 is for the level group's rows
    if (F == 2) __stcs(reinterpret_cast<float2 *>(dst), make_float2(acc[0], acc[1]));
    else __stcs(reinterpret_cast<float4 *>(dst), make_float4(acc[0], acc[1], acc[F - 2], acc[F - 1]));
}

// scratch [L][n][F] -> enc [n][L*F] through a shared-memory tile: both sides coalesced
template <int F>
__global__ void __launch_bounds__(kBlock) hashgrid_transpose_kernel(uint32_t n, uint32_t L, const float *__restrict__ scratch,
                                                                     float *__restrict__ enc) {
    constexpr int kPts = 64;
    extern __shared__ float s_tile[];  // [kPts][L*F + 1]
    const uint32_t p0 = blockIdx.x * kPts, LF = L * F, pitch = LF + 1;
    for (uint32_t e = threadIdx.x; e < L * kPts * F; e += kBlock) {  // level-major read: kPts * F consecutive floats per level
        const uint32_t l = e / (kPts * F), r = e % (kPts * F), pt = r / F, f = r % F;
        if (p0 + pt < n) s_tile[pt * pitch + l * F + f] = __ldcs(scratch + ((size_t)l * n + p0 + pt) * F + f);
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < kPts * LF; e += kBlock) {
        const uint32_t pt = e / LF, c = e % LF;
        if (p0 + pt < n) __stcs(enc + (size_t)(p0 + pt) * LF + c, s_tile[pt * pitch + c]);
    }
}

// enc-shaped [n][L*F] -> level-major [L][n][F] (the backward's input for one- and two-level passes: read in place, a pass
// would pull a 32-byte sector through L2 for every 8 bytes it uses -- four times the stream, next to the level's rows)
template <int F>
__global__ void __launch_bounds__(kBlock) hashgrid_transpose_in_kernel(uint32_t n, uint32_t L, const float *__restrict__ rows,
                                                                        float *__restrict__ scratch) {
    constexpr int kPts = 64;
    extern __shared__ float s_tile[];  // [kPts][L*F + 1]
    const uint32_t p0 = blockIdx.x * kPts, LF = L * F, pitch = LF + 1;
    for (uint32_t e = threadIdx.x; e < kPts * LF; e += kBlock) {
        const uint32_t pt = e / LF, c = e % LF;
        if (p0 + pt < n) s_tile[pt * pitch + c] = __ldcs(rows + (size_t)(p0 + pt) * LF + c);
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < L * kPts * F; e += kBlock) {
        const uint32_t l = e / (kPts * F), r = e % (kPts * F), pt = r / F, f = r % F;
        if (p0 + pt < n) scratch[((size_t)l * n + p0 + pt) * F + f] = s_tile[pt * pitch + l * F + f];
    }
}

// Backward: level-major warps.  A CTA owns 256 consecutive points; warp w walks levels w, w+8 and its
// lanes are 32 CONSECUTIVE points at ONE level.  Samples arrive ray by ray (march_rays emits each
// ray's samples contiguously), so neighbouring lanes usually sit in the same grid cell on the coarse
// and middle levels: runs of lanes with an identical cell are summed with a segmented shuffle
// reduction and only the run's first lane issues the 8 reductions.  L2 reductions are issued per
// active lane (~1.3 cycles each), so this removes ~55 % of them on ray-ordered samples; on
// incoherent points the run detection costs three shuffles and a vote per level.
// `lpg` levels per pass (grid.y = level group, see "level-major passes" above): the CTA's work items are (32-point
// sub-tile, level of the group), dealt to its warps round-robin; lpg = L in one pass is the mapping just described.
template <int DIM, int F, bool kPaired>
__global__ void __launch_bounds__(kBlock) hashgrid_a1_backward_kernel(const __grid_constant__ NgpHashGridA1Descriptor d,
                                                                       const float *__restrict__ pos,
                                                                       const float *__restrict__ d_enc,
                                                                       float *__restrict__ d_table, uint32_t lpg, bool level_major_in) {
    __shared__ LevelMeta s_meta[NGP_HG_MAX_LEVELS];
    if (threadIdx.x < d.L) s_meta[threadIdx.x] = a1_level(d, threadIdx.x);
    __syncthreads();
    constexpr int NC = 1 << DIM;
    constexpr int kWarps = kBlock / 32;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t tile_begin = blockIdx.x * kBlock;

    const uint32_t l0 = blockIdx.y * lpg, l1 = min(l0 + lpg, d.L);
    uint32_t sub_loaded = 0xFFFFFFFFu;
    bool in_range = false;
    uint32_t point = 0;
    float p01[DIM];
    for (uint32_t task = warp; task < kWarps * lpg; task += kWarps) {
        const uint32_t sub = task / lpg, level = l0 + task % lpg;
        if (level >= l1) continue;
        if (sub != sub_loaded) {
            sub_loaded = sub;
            point = tile_begin + sub * 32u + lane;
            in_range = point < d.n_points;
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                const float x = in_range ? __ldg(pos + (size_t)point * DIM + k) : 0.f;
                p01[k] = __fdiv_rn(__fadd_rn(x, d.bound), __fmul_rn(2.f, d.bound));  // encoders.py:87
            }
        }
        {
            const LevelMeta m = s_meta[level];
            float g[F];
#pragma unroll
            for (int f = 0; f < F; ++f) g[f] = 0.f;
            if (in_range) {
                const float *gin = level_major_in ? d_enc + ((size_t)level * d.n_points + point) * F
                                                  : d_enc + ((size_t)point * d.L + level) * F;
                // streamed once (evict-first): the L2 is for the gradient rows the reductions below hit
                if (F == 2) {
                    const float2 v = __ldcs(reinterpret_cast<const float2 *>(gin));
                    g[0] = v.x; g[1] = v.y;
                } else {
                    const float4 v = __ldcs(reinterpret_cast<const float4 *>(gin));
                    g[0] = v.x; g[1] = v.y; g[F - 2] = v.z; g[F - 1] = v.w;
                }
            }
            bool active = false;  // padded / masked samples carry exact zeros: nothing to add
#pragma unroll
            for (int f = 0; f < F; ++f) active |= (g[f] != 0.f);

            uint32_t base[DIM];
            float fr[DIM];
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                const float ps = __fadd_rn(__fmul_rn(p01[k], m.scale), .5f);  // encoders.py:218
                const float fl = floorf(ps);
                base[k] = (uint32_t)(int)fl;
                fr[k] = ps - fl;
            }
            // runs of consecutive active lanes in the same cell
            bool same_as_prev = lane > 0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) same_as_prev &= (__shfl_up_sync(0xffffffffu, base[k], 1) == base[k]);
            const uint32_t A = __ballot_sync(0xffffffffu, active);
            const bool prev_active = lane > 0 && ((A >> (lane - 1)) & 1u);
            const bool head = active && !(same_as_prev && prev_active);
            const uint32_t H = __ballot_sync(0xffffffffu, head);
            const uint32_t above = (lane == 31u) ? 0u : (0xFFFFFFFFu << (lane + 1u));
            const uint32_t stops = (H | ~A) & above;
            const uint32_t seg_end = stops ? (uint32_t)__ffs(stops) - 1u : 32u;

            float v[NC][F];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                float wc = active ? 1.f : 0.f;
#pragma unroll
                for (int k = 0; k < DIM; ++k) wc *= ((c >> (DIM - 1 - k)) & 1) ? fr[k] : 1.f - fr[k];
#pragma unroll
                for (int f = 0; f < F; ++f) v[c][f] = wc * g[f];
            }
#pragma unroll
            for (uint32_t dist = 1; dist < 32; dist <<= 1) {
                const bool take = active && lane + dist < seg_end;
                if (!__any_sync(0xffffffffu, take)) break;  // no run in this warp is longer than `dist`
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int f = 0; f < F; ++f) {
                        const float o = __shfl_down_sync(0xffffffffu, v[c][f], dist);
                        if (take) v[c][f] += o;
                    }
            }
            if (head) {
                // corners c and c + NC/2 differ by +1 along x; when their rows are neighbours (2k, 2k+1) both
                // gradients go out in ONE 16-byte reduction (F = 2; needs d_table 16-byte aligned, kPaired)
#pragma unroll
                for (int c = 0; c < NC / 2; ++c) {
                    uint32_t va[DIM], vb[DIM];
#pragma unroll
                    for (int k = 0; k < DIM; ++k) {
                        va[k] = base[k] + ((c >> (DIM - 1 - k)) & 1);
                        vb[k] = base[k] + (((c + NC / 2) >> (DIM - 1 - k)) & 1);
                    }
                    uint32_t ra = hg::grid_row_unclamped<DIM>(va, m), rb = hg::grid_row_unclamped<DIM>(vb, m);
                    // a row past the table (see hg::grid_row): XLA's scatter-add drops the update
                    if (ra > m.last_row) {
                        ra = m.last_row;
#pragma unroll
                        for (int f = 0; f < F; ++f) v[c][f] = 0.f;
                    }
                    if (rb > m.last_row) {
                        rb = m.last_row;
#pragma unroll
                        for (int f = 0; f < F; ++f) v[c + NC / 2][f] = 0.f;
                    }
                    if (F == 2) {
                        if (kPaired && (ra ^ rb) == 1u) {
                            const bool a_hi = ra & 1u;
                            red_add_v4(d_table + (size_t)(ra & ~1u) * 2, a_hi ? v[c + NC / 2][0] : v[c][0],
                                       a_hi ? v[c + NC / 2][1] : v[c][1], a_hi ? v[c][0] : v[c + NC / 2][0],
                                       a_hi ? v[c][1] : v[c + NC / 2][1]);
                        } else {
                            red_add_v2(d_table + (size_t)ra * 2, v[c][0], v[c][1]);
                            red_add_v2(d_table + (size_t)rb * 2, v[c + NC / 2][0], v[c + NC / 2][1]);
                        }
                    } else {
                        red_add_v4(d_table + (size_t)ra * F, v[c][0], v[c][1], v[c][F - 2], v[c][F - 1]);
                        red_add_v4(d_table + (size_t)rb * F, v[c + NC / 2][0], v[c + NC / 2][1], v[c + NC / 2][F - 2],
                                   v[c + NC / 2][F - 1]);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------- tiny-cuda-nn compatible path
// kernel_grid<float,3,F,CoherentPrime> (tiny-cuda-nn v1.6 include/tiny-cuda-nn/encodings/grid.h),
// as launched at deps/jax-tcnn/lib/impl/hashgrid.cu:53-85: Linear interpolation, GridType::Hash,
// max_level = 1e3, quantize_threshold = 0.
struct TcnnLevel {
    float scale;
    uint32_t res, offset, size;
};

__device__ __forceinline__ TcnnLevel tcnn_level(const uint32_t *__restrict__ offset_table, uint32_t level,
                                                uint32_t N_min, float log2_per_level_scale) {
    TcnnLevel m;
    m.offset = __ldg(offset_table + level);
    m.size = __ldg(offset_table + level + 1) - m.offset;
    m.scale = exp2f((float)level * log2_per_level_scale) * (float)N_min - 1.0f;  // grid_scale()
    m.res = (uint32_t)ceilf(m.scale) + 1u;                                        // grid_resolution()
    return m;
}

__device__ __forceinline__ uint32_t tcnn_index(const uint32_t (&v)[3], const TcnnLevel &m) {  // grid_index()
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (stride <= m.size) {
            index += v[k] * stride;
            stride *= m.res;
        }
    }
    if (m.size < stride) index = v[0] ^ (v[1] * kPrime1) ^ (v[2] * kPrime2);  // coherent prime hash
    return index % m.size + m.offset;
}

template <int F>
__global__ void __launch_bounds__(kBlock) hashgrid_tcnn_forward_kernel(
    NgpHashGridDescriptor d, float log2_b, const uint32_t *__restrict__ offset_table,
    const float *__restrict__ coords_rm, const float *__restrict__ params, float *__restrict__ encoded_rm,
    float *__restrict__ dy_dcoords_rm) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    const uint32_t level = blockIdx.y;
    if (i >= d.n_coords) return;
    const TcnnLevel m = tcnn_level(offset_table, level, d.N_min, log2_b);
    uint32_t base[3];
    float fr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float p = fmaf(m.scale, __ldg(coords_rm + (size_t)k * d.n_coords + i), 0.5f);  // pos_fract()
        float fl = floorf(p);
        base[k] = (uint32_t)(int)fl;
        fr[k] = p - fl;
    }
    float vals[8][F];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t v[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = base[k] + ((c >> k) & 1);  // tcnn: bit k of the corner = axis k
        RowIO<float, F>::load(params, tcnn_index(v, m), vals[c]);
    }
    float acc[F];
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float w = 1.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) w *= ((c >> k) & 1) ? fr[k] : 1.f - fr[k];
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] = fmaf(w, vals[c][f], acc[f]);
    }
#pragma unroll
    for (int f = 0; f < F; ++f) encoded_rm[(size_t)(level * F + f) * d.n_coords + i] = acc[f];

    // dy_dx: one float3 per (level*F+f, point), as tcnn's vector_fullp_t<3> array
#pragma unroll
    for (int f = 0; f < F; ++f) {
        float grad[3];
#pragma unroll
        for (int gd = 0; gd < 3; ++gd) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if ((c >> gd) & 1) continue;  // c has bit gd clear; partner = c | (1 << gd)
                float w = m.scale;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (k != gd) w *= ((c >> k) & 1) ? fr[k] : 1.f - fr[k];
                s += w * (vals[c | (1 << gd)][f] - vals[c][f]);
            }
            grad[gd] = s;
        }
        float *dst = dy_dcoords_rm + ((size_t)(level * F + f) * d.n_coords + i) * 3;
        dst[0] = grad[0]; dst[1] = grad[1]; dst[2] = grad[2];
    }
}

template <int F>
__global__ void __launch_bounds__(kBlock) hashgrid_tcnn_backward_kernel(
    NgpHashGridDescriptor d, float log2_b, const uint32_t *__restrict__ offset_table,
    const float *__restrict__ coords_rm, const float *__restrict__ dL_dy_rm, float *__restrict__ dL_dparams) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    const uint32_t level = blockIdx.y;
    if (i >= d.n_coords) return;
    const TcnnLevel m = tcnn_level(offset_table, level, d.N_min, log2_b);
    uint32_t base[3];
    float fr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float p = fmaf(m.scale, __ldg(coords_rm + (size_t)k * d.n_coords + i), 0.5f);
        float fl = floorf(p);
        base[k] = (uint32_t)(int)fl;
        fr[k] = p - fl;
    }
    float g[F];
#pragma unroll
    for (int f = 0; f < F; ++f) g[f] = __ldg(dL_dy_rm + (size_t)(level * F + f) * d.n_coords + i);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t v[3];
        float w = 1.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[k] = base[k] + ((c >> k) & 1);
            w *= ((c >> k) & 1) ? fr[k] : 1.f - fr[k];
        }
        float *dst = dL_dparams + (size_t)tcnn_index(v, m) * F;
        if (F == 2) red_add_v2(dst, w * g[0], w * g[1]);
        else red_add_v4(dst, w * g[0], w * g[1], w * g[F - 2], w * g[F - 1]);
    }
}

// kernel_grid_backward_input<float,3>: dL/dx[d] = sum_k dL/dy[k] * dy_dx[k][d]
__global__ void __launch_bounds__(kBlock) hashgrid_tcnn_backward_input_kernel(
    uint32_t n, uint32_t LF, const float *__restrict__ dL_dy_rm, const float *__restrict__ dy_dcoords_rm,
    float *__restrict__ dL_dcoords_rm) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (uint32_t k = 0; k < LF; ++k) {
        const float g = __ldg(dL_dy_rm + (size_t)k * n + i);
        const float *j = dy_dcoords_rm + ((size_t)k * n + i) * 3;
        a0 = fmaf(g, __ldg(j + 0), a0);
        a1 = fmaf(g, __ldg(j + 1), a1);
        a2 = fmaf(g, __ldg(j + 2), a2);
    }
    dL_dcoords_rm[i] = a0;
    dL_dcoords_rm[(size_t)n + i] = a1;
    dL_dcoords_rm[2 * (size_t)n + i] = a2;
}

// zero-fill `rows x F` floats where rows = offset_table[L] lives in device memory
__global__ void __launch_bounds__(kBlock) zero_rows_kernel(const uint32_t *__restrict__ offset_table, uint32_t L,
                                                            uint32_t F, float *__restrict__ out) {
    const size_t n = (size_t)__ldg(offset_table + L) * F;
    const size_t n4 = ((uintptr_t)out % 16 == 0) ? n / 4 : 0;
    for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n4; i += (size_t)gridDim.x * kBlock)
        reinterpret_cast<float4 *>(out)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += (size_t)gridDim.x * kBlock) out[i] = 0.f;
}

// Levels per pass for a table that does not fit in L2 (see "level-major passes"): the largest power of two whose level
// groups each stay under `budget` bytes of rows; 0 = the whole table fits, use the single-pass kernels.
// NGP_B200_HG_LPG / NGP_B200_HG_BWD_LPG override it (tuning; 16 = force the single pass).
constexpr size_t kL2TableBudget = 72u << 20;  // of B200's 126 MB L2: rows in flight next to the streamed inputs / outputs
// (measured, tools/hashenc_sweep.py: 64 MB groups -- 4 levels at T = 2^21, 2 at 2^22, 1 at 2^23 -- are the fastest)
unsigned levels_per_pass(const NgpHashGridA1Descriptor *d, size_t row_bytes, const char *env) {
    if (const char *e = getenv(env)) {
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) return (unsigned)v;
        if (v >= 16) return 0;
    }
    if ((size_t)d->offsets[d->L] * row_bytes <= (64u << 20)) return 0;
    for (unsigned lpg = 8; lpg > 1; lpg >>= 1) {
        bool ok = true;
        for (uint32_t l0 = 0; l0 < d->L && ok; l0 += lpg) {
            const uint32_t l1 = l0 + lpg < d->L ? l0 + lpg : d->L;
            ok = (size_t)(d->offsets[l1] - d->offsets[l0]) * row_bytes <= kL2TableBudget;
        }
        if (ok) return lpg;
    }
    return 1;
}

bool a1_validate(const NgpHashGridA1Descriptor *d, const char *op) {
    if (d->L == 0 || d->L > NGP_HG_MAX_LEVELS || (d->dim != 2 && d->dim != 3) || (d->F != 2 && d->F != 4) ||
        d->table_dtype > 1) {
        set_error(NGP_ERR_ARGUMENT, "%s: unsupported L=%u dim=%u F=%u table_dtype=%u", op, d->L, d->dim, d->F,
                  d->table_dtype);
        return false;
    }
    for (uint32_t l = 0; l < d->L; ++l) {
        if (d->offsets[l + 1] <= d->offsets[l]) {
            set_error(NGP_ERR_ARGUMENT, "%s: level %u has no rows (offsets %u..%u)", op, l, d->offsets[l], d->offsets[l + 1]);
            return false;
        }
        // a hashed level indexes `hash mod wrap`: every such row must exist
        const uint32_t wrap = d->wrap_T ? d->wrap_T : d->offsets[l + 1] - d->offsets[l];
        if (((d->hashed_mask >> l) & 1u) && (uint64_t)d->offsets[l] + wrap > d->offsets[d->L]) {
            set_error(NGP_ERR_ARGUMENT, "%s: hashed level %u reaches row %llu of a %u-row table", op, l,
                      (unsigned long long)d->offsets[l] + wrap, d->offsets[d->L]);
            return false;
        }
    }
    return true;
}

}  // namespace
}  // namespace ngp

extern "C" {

void ngp_hashgrid_a1_forward(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpHashGridA1Descriptor>(opaque, opaque_len, "hashgrid_a1_forward");
    if (!d || !a1_validate(d, "hashgrid_a1_forward") || d->n_points == 0) return;
    BufferCursor b{buffers};
    const float *pos = b.next<const float>();
    const void *table = b.next<const void>();
    const uint32_t *group_counts = d->rows_per_group ? b.next<const uint32_t>() : nullptr;
    float *enc = b.next<float>();
    const unsigned blocks = div_up((unsigned long long)d->n_points * d->L, kBlock);
    // one aligned load for both corners of an x-pair needs the table base aligned to two rows
    // (and an even row count: the pair load of the clamped last row must stay inside the table)
    const bool paired = (reinterpret_cast<uintptr_t>(table) % (2 * d->F * (d->table_dtype == 0 ? 4 : 2))) == 0 && d->offsets[d->L] % 2 == 0;
    const bool pow2 = d->wrap_T != 0 && (d->wrap_T & (d->wrap_T - 1u)) == 0;
    // tables beyond L2: level-major passes (dim 3, F 2, power-of-two wrap, pair-aligned table: the NeRF geometry)
    const unsigned lpg = (d->dim == 3 && d->F == 2 && paired && pow2 && !d->rows_per_group)
                             ? levels_per_pass(d, d->F * (d->table_dtype == 0 ? 4 : 2), "NGP_B200_HG_LPG") : 0;
    if (lpg) {
        const dim3 grid(div_up((unsigned long long)d->n_points * lpg, kBlock), div_up(d->L, lpg), 1);
        float *dst = enc;
        const bool scratch = lpg < 4;
        if (scratch) {
            dst = static_cast<float *>(workspace(stream, (size_t)d->n_points * d->L * d->F * sizeof(float)));
            if (!dst) return;
        }
#define NGP_FWD_G(TT, LPG, SCR) \
    hashgrid_a1_forward_group_kernel<3, 2, TT, true, true, LPG, SCR><<<grid, kBlock, 0, stream>>>(*d, pos, static_cast<const TT *>(table), dst)
#define NGP_FWD_GT(TT)                                   \
    do {                                                 \
        if (lpg == 8) NGP_FWD_G(TT, 8, false);           \
        else if (lpg == 4) NGP_FWD_G(TT, 4, false);      \
        else if (lpg == 2) NGP_FWD_G(TT, 2, true);       \
        else NGP_FWD_G(TT, 1, true);                     \
    } while (0)
        if (d->table_dtype == 0) NGP_FWD_GT(float); else NGP_FWD_GT(__half);
#undef NGP_FWD_GT
#undef NGP_FWD_G
        if (!check_launch("hashgrid_a1_forward(level-major)")) return;
        if (scratch) {
            const size_t smem = 64 * (d->L * d->F + 1) * sizeof(float);
            hashgrid_transpose_kernel<2><<<div_up(d->n_points, 64), kBlock, smem, stream>>>(d->n_points, d->L, dst, enc);
            check_launch("hashgrid_a1_forward(transpose)");
        }
        return;
    }
#define NGP_FWD_(DIM, F, TT, P, W) \
    hashgrid_a1_forward_kernel<DIM, F, TT, P, W><<<blocks, kBlock, 0, stream>>>(*d, pos, static_cast<const TT *>(table), group_counts, enc)
#define NGP_FWD(DIM, F, TT)                              \
    do {                                                 \
        if (paired && pow2) NGP_FWD_(DIM, F, TT, true, true);   \
        else if (paired) NGP_FWD_(DIM, F, TT, true, false);     \
        else NGP_FWD_(DIM, F, TT, false, false);                \
    } while (0)
    if (d->table_dtype == 0) {
        if (d->dim == 3 && d->F == 2) NGP_FWD(3, 2, float);
        else if (d->dim == 3) NGP_FWD(3, 4, float);
        else if (d->F == 2) NGP_FWD(2, 2, float);
        else NGP_FWD(2, 4, float);
    } else {
        if (d->dim == 3 && d->F == 2) NGP_FWD(3, 2, __half);
        else if (d->dim == 3) NGP_FWD(3, 4, __half);
        else if (d->F == 2) NGP_FWD(2, 2, __half);
        else NGP_FWD(2, 4, __half);
    }
#undef NGP_FWD
#undef NGP_FWD_
    check_launch("hashgrid_a1_forward");
}

static void launch_hashgrid_a1_backward(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len, bool accumulate);

void ngp_hashgrid_a1_backward(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    launch_hashgrid_a1_backward(stream, buffers, opaque, opaque_len, false);
}

// same, ADDING to d_table instead of zero-filling it first (a batch processed in chunks)
void ngp_hashgrid_a1_backward_acc(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    launch_hashgrid_a1_backward(stream, buffers, opaque, opaque_len, true);
}

static void launch_hashgrid_a1_backward(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len, bool accumulate) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpHashGridA1Descriptor>(opaque, opaque_len, "hashgrid_a1_backward");
    if (!d || !a1_validate(d, "hashgrid_a1_backward")) return;
    BufferCursor b{buffers};
    const float *pos = b.next<const float>();
    const float *d_enc = b.next<const float>();
    float *d_table = b.next<float>();
    if (!accumulate)
        NGP_CUDA_OK(cudaMemsetAsync(d_table, 0, (size_t)d->offsets[d->L] * d->F * sizeof(float), stream), "hashgrid_a1_backward");
    {   // the scatter runs beside kernels that fill the SM's shared memory (the chunked backward of trainer.py pairs it
        // with the 225 KB MLP backward): an SM hosts CTAs of one carve-out at a time, so ask for the same one
        static bool configured = false;  // benign race: idempotent
        if (!configured) {
            cudaFuncSetAttribute(hashgrid_a1_backward_kernel<3, 2, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            configured = true;
        }
    }
    if (d->n_points == 0) return;
    // one CTA per 256 points; all levels in one pass, or level-major passes when the gradient table exceeds L2
    unsigned lpg = levels_per_pass(d, d->F * sizeof(float), "NGP_B200_HG_BWD_LPG");
    if (lpg == 0) lpg = d->L;
    const dim3 blocks(div_up(d->n_points, kBlock), div_up(d->L, lpg), 1);
    const bool paired = reinterpret_cast<uintptr_t>(d_table) % 16 == 0 && d->offsets[d->L] % 2 == 0;
    // one- and two-level passes read d_enc from a level-major copy (see hashgrid_transpose_in_kernel)
    bool lm = lpg <= 2 && lpg < d->L;
    if (lm) {
        auto *scratch = static_cast<float *>(workspace(stream, (size_t)d->n_points * d->L * d->F * sizeof(float)));
        if (!scratch) return;
        const size_t smem = 64 * (d->L * d->F + 1) * sizeof(float);
        if (smem > 48 * 1024) {
            lm = false;  // (L * F > 191: not a geometry of this path) -- read in place
        } else {
            if (d->F == 2) hashgrid_transpose_in_kernel<2><<<div_up(d->n_points, 64), kBlock, smem, stream>>>(d->n_points, d->L, d_enc, scratch);
            else hashgrid_transpose_in_kernel<4><<<div_up(d->n_points, 64), kBlock, smem, stream>>>(d->n_points, d->L, d_enc, scratch);
            if (!check_launch("hashgrid_a1_backward(transpose)")) return;
            d_enc = scratch;
        }
    }
    if (d->dim == 3 && d->F == 2) {
        if (paired) hashgrid_a1_backward_kernel<3, 2, true><<<blocks, kBlock, 0, stream>>>(*d, pos, d_enc, d_table, lpg, lm);
        else hashgrid_a1_backward_kernel<3, 2, false><<<blocks, kBlock, 0, stream>>>(*d, pos, d_enc, d_table, lpg, lm);
    } else if (d->dim == 3) hashgrid_a1_backward_kernel<3, 4, false><<<blocks, kBlock, 0, stream>>>(*d, pos, d_enc, d_table, lpg, lm);
    else if (d->F == 2) {
        if (paired) hashgrid_a1_backward_kernel<2, 2, true><<<blocks, kBlock, 0, stream>>>(*d, pos, d_enc, d_table, lpg, lm);
        else hashgrid_a1_backward_kernel<2, 2, false><<<blocks, kBlock, 0, stream>>>(*d, pos, d_enc, d_table, lpg, lm);
    } else hashgrid_a1_backward_kernel<2, 4, false><<<blocks, kBlock, 0, stream>>>(*d, pos, d_enc, d_table, lpg, lm);
    check_launch("hashgrid_a1_backward");
}

void ngp_hashgrid_encode(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpHashGridDescriptor>(opaque, opaque_len, "hashgrid_encode");
    if (!d) return;
    if (d->F != 2 && d->F != 4) {  // hashgrid.cu:81-85 throws here
        set_error(NGP_ERR_ARGUMENT, "hashgrid_encode: supported values of F (n_features_per_level) are [2, 4], got %u", d->F);
        return;
    }
    if (d->n_coords == 0 || d->L == 0) return;
    BufferCursor b{buffers};
    const uint32_t *offset_table = b.next<const uint32_t>();
    const float *coords_rm = b.next<const float>();
    const float *params = b.next<const float>();
    float *encoded_rm = b.next<float>();
    float *dy_dcoords_rm = b.next<float>();
    const dim3 grid(div_up(d->n_coords, kBlock), d->L, 1);
    const float log2_b = log2f(d->per_level_scale);  // hashgrid.cu:60
    if (d->F == 2)
        hashgrid_tcnn_forward_kernel<2><<<grid, kBlock, 0, stream>>>(*d, log2_b, offset_table, coords_rm, params, encoded_rm, dy_dcoords_rm);
    else
        hashgrid_tcnn_forward_kernel<4><<<grid, kBlock, 0, stream>>>(*d, log2_b, offset_table, coords_rm, params, encoded_rm, dy_dcoords_rm);
    check_launch("hashgrid_encode");
}

void ngp_hashgrid_encode_backward(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpHashGridDescriptor>(opaque, opaque_len, "hashgrid_encode_backward");
    if (!d) return;
    if (d->F != 2 && d->F != 4) {
        set_error(NGP_ERR_ARGUMENT, "hashgrid_encode_backward: supported values of F are 2, 4, got %u", d->F);
        return;
    }
    if (d->L == 0) return;
    BufferCursor b{buffers};
    const uint32_t *offset_table = b.next<const uint32_t>();
    const float *coords_rm = b.next<const float>();
    const float *dL_dy_rm = b.next<const float>();
    const float *dy_dcoords_rm = b.next<const float>();
    float *dL_dparams = b.next<float>();
    float *dL_dcoords_rm = b.next<float>();
    // The reference sizes a host-side memset from a host copy of the offset table that an
    // un-synchronised D2H memcpy may still be filling (hashgrid.cu:115-117).  The table size is not
    // in the descriptor, so the zero-fill runs on the device and reads offset_table[L] there.
    zero_rows_kernel<<<148 * 8, kBlock, 0, stream>>>(offset_table, d->L, d->F, dL_dparams);
    if (!check_launch("hashgrid_encode_backward(zero)")) return;
    if (d->n_coords == 0) return;
    const dim3 grid(div_up(d->n_coords, kBlock), d->L, 1);
    const float log2_b = log2f(d->per_level_scale);
    if (d->F == 2)
        hashgrid_tcnn_backward_kernel<2><<<grid, kBlock, 0, stream>>>(*d, log2_b, offset_table, coords_rm, dL_dy_rm, dL_dparams);
    else
        hashgrid_tcnn_backward_kernel<4><<<grid, kBlock, 0, stream>>>(*d, log2_b, offset_table, coords_rm, dL_dy_rm, dL_dparams);
    if (!check_launch("hashgrid_encode_backward")) return;
    hashgrid_tcnn_backward_input_kernel<<<div_up(d->n_coords, kBlock), kBlock, 0, stream>>>(
        d->n_coords, d->L * d->F, dL_dy_rm, dy_dcoords_rm, dL_dcoords_rm);
    check_launch("hashgrid_encode_backward(input)");
}

}  // extern "C"
