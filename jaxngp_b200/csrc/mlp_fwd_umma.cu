// mlp_fwd_umma.cu -- HashGridEncoder + NeRF MLP forward with the five dense layers on tcgen05 (kind::tf32,
// accumulators in TMEM): the sm_100a-native counterpart of nerf_fused_forward_kernel (mlp.cu), same contract.
//
// Work unit: a tile of 128 consecutive samples, owned by a GROUP of 128 threads (4 warps); thread r of the group
// owns sample row r from the gather to the output, so nothing is exchanged between threads except through the
// tensor core.  A CTA holds kGroups independent groups (their tiles interleave: while one group waits for its
// gathers another runs an epilogue and the tensor core works for a third) and one shared copy of the weights.
//
//   gather : thread r encodes its sample on all 16 levels (a warp = 32 neighbouring samples of a ray on ONE level at
//            a time: coarse-level corners coalesce into few L1 wavefronts), rounds to tf32 and writes row r of the
//            K-major SWIZZLE_128B panel P_A                                     [enc also to HBM for the backward]
//   layer  : one elected thread issues D[128 x N] = A[128 x K] . W[K x N] as K/8 tcgen05.mma (A = panel, K-major;
//            B = weights W[in][out], MN-major SWIZZLE_128B_BASE32B panels staged once per CTA), commits to the
//            group's mbarrier
//   epilog : thread r reads TMEM lane r (tcgen05.ld 32x32b), applies ReLU / exp / SH / sigmoid, rounds to tf32 and
//            writes row r of the next layer's A panel (P_B, or P_A again for [x | SH(dir)])
// Every MMA is awaited before its epilogue, so ONE 32 KB panel pair and 64 TMEM columns per group suffice.
// Instruction budget per 16 samples: ~1300 warp instructions against ~2100 for the mma.sync kernel, whose B-fragment
// shared-memory loads, HMMA issue and fragment conversions the tensor core now does asynchronously.
#include "common.cuh"
#include "hashgrid.cuh"
#include "marching.cuh"
#include "umma.cuh"

namespace ngp {
namespace {

#ifndef NGP_UMMA_GROUPS
#define NGP_UMMA_GROUPS 4
#endif
constexpr int kGroups = NGP_UMMA_GROUPS, kGroupThreads = 128, kThreadsU = kGroups * kGroupThreads;
constexpr uint32_t kTile = 128;
// one 32 KB panel pair per group: every MMA is awaited before the epilogue that overwrites its A operand, so the
// 32-wide inputs (enc, then [x | SH]) live in the first panel of the pair the 64-wide activations use
constexpr uint32_t kPA = kTile * 128u, kPB = 2 * kTile * 128u, kGroupBytes = kPB;
// weight panels (MN-major: row = input feature k, 32 output features per 128-byte row; narrow layers zero-padded to 32)
constexpr uint32_t WB0 = 0, WB1 = WB0 + 2 * 32 * 128, WB2 = WB1 + 64 * 128, WB3 = WB2 + 2 * 32 * 128,
                   WB4 = WB3 + 2 * 64 * 128, kWBytes = WB4 + 64 * 128;  // 49152
constexpr int G_W0 = 0, G_W1 = G_W0 + 32 * 64, G_W2 = G_W1 + 64 * 16, G_W3 = G_W2 + 32 * 64, G_W4 = G_W3 + 64 * 64;
constexpr uint32_t kTmemColsPerGroup = 64, kTmemCols = kGroups <= 4 ? 256 : 512;
constexpr size_t kFwdUmmaSmem = 1024 + kWBytes + kGroups * kGroupBytes;

// round to tf32 (nearest, ties away from zero -- cvt.rna.tf32.f32 without its NaN/Inf special case): 2 instructions
__device__ __forceinline__ uint32_t tf32r(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

__device__ __forceinline__ void group_barrier(uint32_t group) {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1u), "n"(kGroupThreads) : "memory");
}

// stage W[K][N] (global, row-major, N_real columns) as ceil(N/32) MN-major panels of K rows
__device__ __forceinline__ void stage_weight(uint8_t *dst, const float *__restrict__ w, int K, int n_real, int n_padded) {
    for (int i = threadIdx.x; i < K * n_padded; i += kThreadsU) {
        const int k = i / n_padded, nn = i % n_padded;
        const float v = nn < n_real ? __ldg(w + k * n_real + nn) : 0.f;
        *reinterpret_cast<uint32_t *>(dst + (nn >> 5) * (K * 128) + umma::panel_offset_mn32(k, nn & 31)) = tf32r(v);
    }
}

// write 4 consecutive features (16-byte chunk `c4` of the 32-wide panel `p`) of row r into a K-major panel set
__device__ __forceinline__ void store_chunk(uint8_t *panel, uint32_t r, uint32_t j0, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    const uint32_t p = j0 >> 5, c4 = (j0 & 31u) >> 2;
    *reinterpret_cast<uint4 *>(panel + p * kPA + r * 128u + ((c4 ^ (r & 7u)) << 4)) = make_uint4(a, b, c, d);
}

// real spherical harmonics, degree 4 (models/encoders.py:365-406; same expressions as sh4_lane in mlp.cu)
__device__ __forceinline__ void sh16(float x, float y, float z, float (&s)[16]) {
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    s[0] = 0.28209479177387814f;
    s[1] = -0.48860251190291987f * y;
    s[2] = 0.48860251190291987f * z;
    s[3] = -0.48860251190291987f * x;
    s[4] = 1.0925484305920792f * xy;
    s[5] = -1.0925484305920792f * yz;
    s[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    s[7] = -1.0925484305920792f * xz;
    s[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    s[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    s[10] = 2.8906114426405538f * xy * z;
    s[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    s[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    s[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    s[14] = 1.4453057213202769f * z * (x2 - y2);
    s[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

// ReLU + tf32 rounding of 32 accumulator columns -> columns col0 .. col0+31 of row r of a K-major panel set
__device__ __forceinline__ void relu_store32(uint8_t *panel, uint32_t r, uint32_t col0, const uint32_t (&v)[32]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) q[k] = tf32r(fmaxf(__uint_as_float(v[4 * c + k]), 0.f));
        store_chunk(panel, r, col0 + 4 * c, q[0], q[1], q[2], q[3]);
    }
}

// Grouped layout: the list of groups (ray slots) that hold at least one sample, in slot order within a warp and
// nearly so across warps (neighbouring slots hold neighbouring rays: keeps their gathers close), so that a tile is
// made of live slots only -- late iterations of the slot-refill loop have most slots idle.
__global__ void __launch_bounds__(256) compact_live_groups_kernel(uint32_t n_groups, const uint32_t *__restrict__ group_counts,
                                                                  uint32_t *__restrict__ live_list, uint32_t *__restrict__ n_live) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
    const bool live = i < n_groups && __ldg(group_counts + i) != 0u;
    const uint32_t votes = __ballot_sync(0xffffffffu, live);
    if (votes == 0u) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(n_live, __popc(votes));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (live) live_list[base + __popc(votes & ((1u << lane) - 1u))] = i;
}

// One 128-sample tile through the gather and the five tcgen05 layers: thread r of `group` owns row r (position x,
// direction dir; rows with live = false gather nothing and produce garbage nobody reads).  o = (density, rgb).
// Shared by the forward op (tiles of the sample array) and the whole-frame kernel (tiles of a group's eight ray slots).
template <typename TT, bool kWriteEnc>
__device__ __forceinline__ void tile_forward(const TT *__restrict__ table, const hg::LevelMeta *s_meta, float bound, const float (&x)[3],
                                             const float (&dir)[3], bool live, float *__restrict__ enc_row, uint8_t *PA, uint8_t *PB,
                                             uint32_t pa, uint32_t pb, uint32_t wa, uint32_t tmem_d, uint32_t tmem_row, uint64_t *bar,
                                             uint32_t &phase, uint32_t group, uint32_t r, float4 &o) {
    constexpr uint32_t kI64 = umma::make_idesc(128, 64, false, true), kI32 = umma::make_idesc(128, 32, false, true);
    // ---- gather: this thread's sample on all 16 levels
    float p01[3];
    hg::unit_pos<3>(x, bound, p01);
#pragma unroll 2
    for (int lv = 0; lv < 16; lv += 2) {  // partially rolled: the fully unrolled gather overflows the instruction cache
        float e0[2], e1[2];
        hg::encode_point_level_pred<TT>(table, s_meta[lv], p01, live, e0);
        hg::encode_point_level_pred<TT>(table, s_meta[lv + 1], p01, live, e1);
        if (kWriteEnc && live) *reinterpret_cast<float4 *>(enc_row + 2 * lv) = make_float4(e0[0], e0[1], e1[0], e1[1]);
        store_chunk(PA, r, 2 * lv, tf32r(e0[0]), tf32r(e0[1]), tf32r(e1[0]), tf32r(e1[1]));
    }
    const float dx = dir[0], dy = dir[1], dz = dir[2];
    uint32_t v[32];
    // ---- layer 0: enc[32] -> 64, ReLU
    umma::fence_smem_to_async();
    group_barrier(group);
    if (r == 0) {
        umma::fence_after_sync();
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks)
            umma::mma_tf32(tmem_d, umma::desc_k_major(pa, ks), umma::desc_mn_major(wa + WB0, ks, 32 * 128), kI64, ks > 0);
        umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1u;
    umma::fence_after_sync();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        umma::tmem_ld32(tmem_row + 32 * h, v);
        umma::tmem_ld_wait();
        relu_store32(PB, r, 32 * h, v);
    }
    // ---- layer 1: 64 -> 16 (padded to 32), no activation; density = exp(x[0]); hin = [x | SH4(dir)]
    umma::fence_before_sync();
    umma::fence_smem_to_async();
    group_barrier(group);
    if (r == 0) {
        umma::fence_after_sync();
#pragma unroll
        for (uint32_t ks = 0; ks < 8; ++ks)
            umma::mma_tf32(tmem_d, umma::desc_k_major(pb + (ks >> 2) * kPA, ks & 3u), umma::desc_mn_major(wa + WB1, ks, 64 * 128), kI32, ks > 0);
        umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1u;
    umma::fence_after_sync();
    umma::tmem_ld32(tmem_row, v);
    umma::tmem_ld_wait();
    const float density = expf(__uint_as_float(v[0]));
#pragma unroll
    for (int c = 0; c < 4; ++c)
        store_chunk(PA, r, 4 * c, tf32r(__uint_as_float(v[4 * c])), tf32r(__uint_as_float(v[4 * c + 1])),
                    tf32r(__uint_as_float(v[4 * c + 2])), tf32r(__uint_as_float(v[4 * c + 3])));
    {
        float s[16];
        sh16(dx, dy, dz, s);
#pragma unroll
        for (int c = 0; c < 4; ++c)
            store_chunk(PA, r, 16 + 4 * c, tf32r(s[4 * c]), tf32r(s[4 * c + 1]), tf32r(s[4 * c + 2]), tf32r(s[4 * c + 3]));
    }
    // ---- layer 2: hin[32] -> 64, ReLU
    umma::fence_before_sync();
    umma::fence_smem_to_async();
    group_barrier(group);
    if (r == 0) {
        umma::fence_after_sync();
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks)
            umma::mma_tf32(tmem_d, umma::desc_k_major(pa, ks), umma::desc_mn_major(wa + WB2, ks, 32 * 128), kI64, ks > 0);
        umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1u;
    umma::fence_after_sync();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        umma::tmem_ld32(tmem_row + 32 * h, v);
        umma::tmem_ld_wait();
        relu_store32(PB, r, 32 * h, v);
    }
    // ---- layer 3: 64 -> 64, ReLU (the MMA has been awaited, so its A panel can take the result)
    umma::fence_before_sync();
    umma::fence_smem_to_async();
    group_barrier(group);
    if (r == 0) {
        umma::fence_after_sync();
#pragma unroll
        for (uint32_t ks = 0; ks < 8; ++ks)
            umma::mma_tf32(tmem_d, umma::desc_k_major(pb + (ks >> 2) * kPA, ks & 3u), umma::desc_mn_major(wa + WB3, ks, 64 * 128), kI64, ks > 0);
        umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1u;
    umma::fence_after_sync();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        umma::tmem_ld32(tmem_row + 32 * h, v);
        umma::tmem_ld_wait();
        relu_store32(PB, r, 32 * h, v);
    }
    // ---- layer 4: 64 -> 3 (padded to 32), sigmoid
    umma::fence_before_sync();
    umma::fence_smem_to_async();
    group_barrier(group);
    if (r == 0) {
        umma::fence_after_sync();
#pragma unroll
        for (uint32_t ks = 0; ks < 8; ++ks)
            umma::mma_tf32(tmem_d, umma::desc_k_major(pb + (ks >> 2) * kPA, ks & 3u), umma::desc_mn_major(wa + WB4, ks, 64 * 128), kI32, ks > 0);
        umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1u;
    umma::fence_after_sync();
    umma::tmem_ld32(tmem_row, v);
    umma::tmem_ld_wait();
    o.x = density;
    o.y = 1.f / (1.f + expf(-__uint_as_float(v[0])));
    o.z = 1.f / (1.f + expf(-__uint_as_float(v[1])));
    o.w = 1.f / (1.f + expf(-__uint_as_float(v[2])));
    umma::fence_before_sync();  // the next tile's first MMA overwrites the accumulator these loads read
}

template <typename TT, bool kWriteEnc>
__global__ void __launch_bounds__(kThreadsU, 1) nerf_fused_forward_umma_kernel(const __grid_constant__ NgpNerfFusedDescriptor d,
                                                                               const float *__restrict__ pos,
                                                                               const TT *__restrict__ table,
                                                                               const float *__restrict__ dirs,
                                                                               const float *__restrict__ weights,
                                                                               const uint32_t *__restrict__ group_counts,
                                                                               const uint32_t *__restrict__ live_list,
                                                                               const uint32_t *__restrict__ n_live_ptr,
                                                                               float *__restrict__ out, float *__restrict__ enc_out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *wsm = base;
    __shared__ uint64_t bars[kGroups];
    __shared__ uint32_t tmem_slot;
    __shared__ hg::LevelMeta s_meta[16];
    const uint32_t tid = threadIdx.x, group = tid / kGroupThreads, r = tid % kGroupThreads, gw = r >> 5;
    uint8_t *PB = base + kWBytes + group * kGroupBytes, *PA = PB;

    stage_weight(wsm + WB0, weights + G_W0, 32, 64, 64);
    stage_weight(wsm + WB1, weights + G_W1, 64, 16, 32);
    stage_weight(wsm + WB2, weights + G_W2, 32, 64, 64);
    stage_weight(wsm + WB3, weights + G_W3, 64, 64, 64);
    stage_weight(wsm + WB4, weights + G_W4, 64, 3, 32);
    if (tid < 16) s_meta[tid] = hg::a1_level(d.grid, tid);
    if (tid < kGroups) umma::mbar_init(&bars[tid], 1);
    if (tid == 0) umma::fence_mbar_init();
    if (tid < 32) umma::tmem_alloc(&tmem_slot, kTmemCols);
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_d = tmem_slot + group * kTmemColsPerGroup;    // accumulator of this group (MMA operand address)
    const uint32_t tmem_row = tmem_d + ((32u * gw) << 16);            // this warp's 32 lanes of it (tcgen05.ld address)
    const uint32_t pa = umma::smem_u32(PA), pb = umma::smem_u32(PB), wa = umma::smem_u32(wsm);
    uint64_t *bar = &bars[group];

    const uint32_t n = d.grid.n_points, rpg = d.grid.rows_per_group;
    const float bound = d.grid.bound;
    // grouped layout: a tile = the next 128 / rpg live groups of the compacted list, thread r = row r % rpg of the
    // (r / rpg)-th of them
    const uint32_t groups_per_tile = rpg ? kTile / rpg : 0u, my_slot = rpg ? r / rpg : 0u, my_k = rpg ? r % rpg : 0u;
    const uint32_t n_live = rpg ? __ldg(n_live_ptr) : 0u;
    const uint32_t n_tiles = rpg ? (n_live + groups_per_tile - 1) / groups_per_tile : (n + kTile - 1) / kTile;
    uint32_t phase = 0;
    for (uint32_t tile = blockIdx.x * kGroups + group; tile < n_tiles; tile += gridDim.x * kGroups) {
        uint32_t row = tile * kTile + r, dir_row = row;
        bool live = row < n;
        if (rpg) {  // padding rows are neither read nor written
            const uint32_t li = tile * groups_per_tile + my_slot;
            live = my_slot < groups_per_tile && li < n_live;
            if (live) {
                dir_row = __ldg(live_list + li);
                row = dir_row * rpg + my_k;
                live = my_k < __ldg(group_counts + dir_row) && row < n;
            }
        }
        float x[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 1.f};
        if (live) {
            x[0] = __ldg(pos + (size_t)row * 3 + 0);
            x[1] = __ldg(pos + (size_t)row * 3 + 1);
            x[2] = __ldg(pos + (size_t)row * 3 + 2);
            dir[0] = __ldg(dirs + (size_t)dir_row * 3 + 0);
            dir[1] = __ldg(dirs + (size_t)dir_row * 3 + 1);
            dir[2] = __ldg(dirs + (size_t)dir_row * 3 + 2);
        }
        float4 o;
        tile_forward<TT, kWriteEnc>(table, s_meta, bound, x, dir, live, kWriteEnc ? enc_out + (size_t)row * 32 : nullptr, PA, PB, pa, pb, wa,
                                    tmem_d, tmem_row, bar, phase, group, r, o);
        if (live) reinterpret_cast<float4 *>(out)[row] = o;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc(tmem_slot, kTmemCols);
}


// ---------------------------------------------------------------- whole-frame inference kernel
// render_image_inference (models/renderers/cuda.py:244-373) as ONE persistent launch: the reference's slot-refill loop
// -- march_rays_inference -> NeRF -> integrate_rays_inference -> scatter, repeated until every ray has terminated, with
// a host check per iteration (cuda.py:318-361) -- runs inside the kernel.  A tile group (128 threads) owns EIGHT ray
// slots, one per half-warp; per pass
//   refill   : a free slot takes the next ray of the frame (one atomicAdd: arrival order, like the reference's kernel)
//   march    : the half-warp marches up to 16 samples of its ray (march::march_chunk<16>, the code of the drop-in op);
//              sample w lands in lane w through a shared-memory staging row
//   NeRF     : the group's 8 x 16 rows go through the gather and the five tcgen05 layers (tile_forward: the code of
//              the forward op); thread = sample, density / colour stay in registers
//   integrate: the half-warp composites its 16 samples in order (integrating.cu:278-314: same expressions, the
//              transmittance chain replayed through shuffles, exact early stop), then terminates the ray (pixel out,
//              slot free) or keeps (T, colour, t) in registers for the next pass.
// No sample, no slot state and no counter leaves the SM between passes; four groups per CTA interleave, so one group's
// dependent bitfield loads and compositing chain run under the others' gathers.  Rays must arrive pre-advanced through
// the empty space in front of them (ngp_march_rays_skip_empty), as for the loop renderer.
constexpr uint32_t kSlots = 8, kCap = 16;
constexpr float kTThresholdFrame = 1e-4f;  // integrating.cu:12

template <typename TT>
__global__ void __launch_bounds__(kThreadsU, 1) nerf_render_frame_kernel(const __grid_constant__ NgpRenderFrameDescriptor d,
                                                                         const float *__restrict__ rays_o,
                                                                         const float *__restrict__ rays_d,
                                                                         const float *__restrict__ t_starts,
                                                                         const float *__restrict__ t_ends,
                                                                         const uint8_t *__restrict__ bitfield,
                                                                         const float *__restrict__ rays_bg,
                                                                         const TT *__restrict__ table,
                                                                         const float *__restrict__ weights,
                                                                         uint32_t *__restrict__ next_ray,
                                                                         float4 *__restrict__ rays_rgbd,
                                                                         unsigned long long *__restrict__ counters) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *wsm = base;
    __shared__ uint64_t bars[kGroups];
    __shared__ uint32_t tmem_slot;
    __shared__ hg::LevelMeta s_meta[16];
    __shared__ float s_stage[kGroups][kSlots][kCap][5];   // x, y, z, ds, t of the samples a slot marched this pass
    __shared__ uint32_t s_count[kGroups][kSlots], s_idle[kGroups][kSlots];
    const uint32_t tid = threadIdx.x, group = tid / kGroupThreads, r = tid % kGroupThreads, gw = r >> 5;
    const uint32_t slot = r >> 4, lane = r & 15u, shift = tid & 16u, mask = 0xFFFFu << shift;
    uint8_t *PB = base + kWBytes + group * kGroupBytes, *PA = PB;

    stage_weight(wsm + WB0, weights + G_W0, 32, 64, 64);
    stage_weight(wsm + WB1, weights + G_W1, 64, 16, 32);
    stage_weight(wsm + WB2, weights + G_W2, 32, 64, 64);
    stage_weight(wsm + WB3, weights + G_W3, 64, 64, 64);
    stage_weight(wsm + WB4, weights + G_W4, 64, 3, 32);
    if (tid < 16) s_meta[tid] = hg::a1_level(d.grid, tid);
    if (tid < kGroups) umma::mbar_init(&bars[tid], 1);
    if (tid == 0) umma::fence_mbar_init();
    if (tid < 32) umma::tmem_alloc(&tmem_slot, kTmemCols);
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_d = tmem_slot + group * kTmemColsPerGroup, tmem_row = tmem_d + ((32u * gw) << 16);
    const uint32_t pa = umma::smem_u32(PA), pb = umma::smem_u32(PB), wa = umma::smem_u32(wsm);
    uint64_t *bar = &bars[group];

    const NgpMarchingInferenceDescriptor &mp = d.march;
    const march::Grid g = march::make_grid(mp.diagonal_n_steps, mp.K, mp.G, mp.bound, mp.stepsize_portion, bitfield);
    const uint32_t N = mp.n_total_rays;
    const float bound = d.grid.bound;
    // slot state, uniform across the slot's 16 lanes
    bool have_ray = false, exhausted = false, live = false;
    uint32_t ray_idx = 0, phase = 0;
    march::Ray ray = {};
    float t_cur = 0.f, t_end = 0.f, T = 1.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned long long n_samples_total = 0, n_rays_done = 0;
    float *stage = &s_stage[group][slot][0][0];

    for (;;) {
        // ---- refill (cuda.py:326-361's admission, one ray per free slot)
        if (!have_ray && !exhausted) {
            uint32_t next = 0;
            if (lane == 0) next = atomicAdd(next_ray, 1u);
            next = __shfl_sync(mask, next, 0, 16);
            if (next >= N) {
                exhausted = true;
            } else {
                ray_idx = next;
                ray = march::load_ray(rays_o, rays_d, next);
                t_cur = __ldg(t_starts + next);
                t_end = __ldg(t_ends + next);
                T = 1.f;
                acc = make_float4(0.f, 0.f, 0.f, 0.f);
                have_ray = true;
            }
        }
        // ---- march up to kCap samples of this slot's ray.  marching.cu:317 is evaluated on EVERY pass (strict): a ray whose
        // previous pass ended past t_end -- its far-plane sample was that pass's 16th -- must not meet the far-plane rule again
        uint32_t n = 0;
        live = have_ray && !(t_end < t_cur);
        if (live)
            n = march::march_chunk<16>(
                g, ray, t_cur, t_end, kCap, lane, mask, shift,
                [&](uint32_t w, const march::EvalPoint &s, float t) {
                    stage[w * 5 + 0] = s.px;
                    stage[w * 5 + 1] = s.py;
                    stage[w * 5 + 2] = s.pz;
                    stage[w * 5 + 3] = s.ds;
                    stage[w * 5 + 4] = t;
                },
                [&](uint32_t w, float ds) { stage[w * 5 + 3] = ds; });
        if (lane == 0) {
            s_count[group][slot] = n;
            s_idle[group][slot] = (!have_ray && exhausted) ? 1u : 0u;
        }
        group_barrier(group);  // staging rows and counts of the whole group are visible
        uint32_t n_group = 0, idle = 0;
#pragma unroll
        for (uint32_t k = 0; k < kSlots; ++k) {
            n_group += s_count[group][k];
            idle += s_idle[group][k];
        }
        if (idle == kSlots) break;  // group-uniform: every slot is free and the frame has no rays left
        const bool row_live = lane < n;
        float x[3] = {0.f, 0.f, 0.f}, ds = 0.f, z = 0.f;
        if (row_live) {
            x[0] = stage[lane * 5 + 0];
            x[1] = stage[lane * 5 + 1];
            x[2] = stage[lane * 5 + 2];
            ds = stage[lane * 5 + 3];
            z = stage[lane * 5 + 4];
        }
        // ---- NeRF on the group's rows (skipped, group-uniformly, when no slot marched a sample this pass)
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n_group) {
            const float dir[3] = {ray.dx, ray.dy, ray.dz};
            tile_forward<TT, false>(table, s_meta, bound, x, dir, row_live, nullptr, PA, PB, pa, pb, wa, tmem_d, tmem_row, bar, phase, group, r, o);
        }
        // ---- integrate this slot's samples in order (integrating.cu:278-314), all 16 lanes in lockstep
        if (have_ray) {
            const float alpha = row_live ? 1.f - __expf(-o.x * ds) : 0.f;
#pragma unroll
            for (uint32_t s = 0; s < kCap; ++s) {
                const float a = __shfl_sync(mask, alpha, s, 16);
                const float cr = __shfl_sync(mask, o.y, s, 16), cg = __shfl_sync(mask, o.z, s, 16), cb = __shfl_sync(mask, o.w, s, 16);
                const float zz = __shfl_sync(mask, z, s, 16);
                if (T > kTThresholdFrame && s < n) {
                    const float w = T * a;
                    acc.x += w * cr;
                    acc.y += w * cg;
                    acc.z += w * cb;
                    acc.w += w * zz;
                    T *= (1.f - a);
                }
            }
            bool term;
            float4 out;
            if (T <= kTThresholdFrame) {  // integrating.cu:293-301
                const float idenom = 1.f / (1.f - T);
                term = true;
                out = make_float4(acc.x * idenom, acc.y * idenom, acc.z * idenom, acc.w * idenom);
            } else {  // integrating.cu:302-314
                term = n < kCap;
                out = acc;
                if (term) {
                    out.x = acc.x + T * __ldg(rays_bg + 3 * (size_t)ray_idx + 0);
                    out.y = acc.y + T * __ldg(rays_bg + 3 * (size_t)ray_idx + 1);
                    out.z = acc.z + T * __ldg(rays_bg + 3 * (size_t)ray_idx + 2);
                }
            }
            n_samples_total += n;
            if (term) {
                if (lane == 0) rays_rgbd[ray_idx] = out;
                ++n_rays_done;
                have_ray = false;
            }
        }
        group_barrier(group);  // everyone has read this pass's staging rows and counts before the next pass overwrites them
    }
    if (lane == 0) {
        if (n_rays_done) atomicAdd(counters + 0, n_rays_done);
        if (n_samples_total) atomicAdd(counters + 1, n_samples_total);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc(tmem_slot, kTmemCols);
}

}  // namespace
}  // namespace ngp

extern "C" void ngp_render_frame(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpRenderFrameDescriptor>(opaque, opaque_len, "render_frame");
    if (!d) return;
    const NgpHashGridA1Descriptor &gd = d->grid;
    BufferCursor b{buffers};
    const float *rays_o = b.next<const float>();
    const float *rays_d = b.next<const float>();
    const float *t_starts = b.next<const float>();
    const float *t_ends = b.next<const float>();
    const uint8_t *bitfield = b.next<const uint8_t>();
    const float *rays_bg = b.next<const float>();
    const void *table = b.next<const void>();
    const float *weights = b.next<const float>();
    uint32_t *next_ray = b.next<uint32_t>();
    unsigned long long *counters = b.next<unsigned long long>();
    float4 *rays_rgbd = b.next<float4>();
    const size_t pair_bytes = gd.table_dtype == 0 ? 16 : 8;
    if (gd.dim != 3 || gd.L != 16 || gd.F != 2 || gd.table_dtype > 1 || gd.wrap_T == 0 || (gd.wrap_T & (gd.wrap_T - 1u)) != 0 ||
        reinterpret_cast<uintptr_t>(table) % pair_bytes != 0 || gd.offsets[gd.L] % 2 != 0 || gd.rows_per_group != 0) {
        set_error(NGP_ERR_ARGUMENT, "render_frame: needs dim=3 L=16 F=2, power-of-two wrap_T and a table of an even number of rows aligned to "
                  "two rows (got dim=%u L=%u F=%u wrap_T=%u)", gd.dim, gd.L, gd.F, gd.wrap_T);
        return;
    }
    if (d->march.K == 0 || d->march.G == 0 || d->march.G > 1024 || d->march.march_steps_cap != kCap) {
        set_error(NGP_ERR_ARGUMENT, "render_frame: expected K > 0, 0 < G <= 1024 and march_steps_cap = %u, got K=%u G=%u cap=%u", kCap,
                  d->march.K, d->march.G, d->march.march_steps_cap);
        return;
    }
    NGP_CUDA_OK(cudaMemsetAsync(next_ray, 0, sizeof(uint32_t), stream), "render_frame");
    NGP_CUDA_OK(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned long long), stream), "render_frame");
    if (d->march.n_total_rays == 0) return;
    const unsigned blocks = min(div_up(d->march.n_total_rays, kGroups * kSlots), 148u);  // persistent: one CTA per SM
#define NGP_FRAME(TT)                                                                                                  \
    do {                                                                                                               \
        static bool configured = false; /* benign race: idempotent */                                                  \
        if (!configured) {                                                                                             \
            cudaFuncSetAttribute(nerf_render_frame_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdUmmaSmem); \
            configured = true;                                                                                         \
        }                                                                                                              \
        nerf_render_frame_kernel<TT><<<blocks, kThreadsU, kFwdUmmaSmem, stream>>>(                                      \
            *d, rays_o, rays_d, t_starts, t_ends, bitfield, rays_bg, static_cast<const TT *>(table), weights, next_ray, rays_rgbd, counters); \
    } while (0)
    if (gd.table_dtype == 0) NGP_FRAME(float); else NGP_FRAME(__half);
#undef NGP_FRAME
    check_launch("render_frame");
}

extern "C" void ngp_nerf_fused_forward_umma(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpNerfFusedDescriptor>(opaque, opaque_len, "nerf_fused_forward_umma");
    if (!d) return;
    const NgpHashGridA1Descriptor &gd = d->grid;
    BufferCursor b{buffers};
    const float *pos = b.next<const float>();
    const void *table = b.next<const void>();
    const float *dirs = b.next<const float>();
    const float *weights = b.next<const float>();
    const uint32_t *group_counts = gd.rows_per_group ? b.next<const uint32_t>() : nullptr;
    float *out = b.next<float>();
    float *enc_out = d->write_enc ? b.next<float>() : nullptr;
    const size_t pair_bytes = gd.table_dtype == 0 ? 16 : 8;
    if (gd.dim != 3 || gd.L != 16 || gd.F != 2 || gd.table_dtype > 1 || gd.wrap_T == 0 || (gd.wrap_T & (gd.wrap_T - 1u)) != 0 ||
        reinterpret_cast<uintptr_t>(table) % pair_bytes != 0 || gd.offsets[gd.L] % 2 != 0 || d->density_only) {
        set_error(NGP_ERR_ARGUMENT,
                  "nerf_fused_forward_umma: needs dim=3 L=16 F=2, power-of-two wrap_T, a table of an even number of rows aligned to two rows and the "
                  "full (density + colour) output (got dim=%u L=%u F=%u wrap_T=%u density_only=%u); use nerf_fused_forward",
                  gd.dim, gd.L, gd.F, gd.wrap_T, d->density_only);
        return;
    }
    if (gd.n_points == 0) return;
    if (gd.rows_per_group > kTile) {
        set_error(NGP_ERR_ARGUMENT, "nerf_fused_forward_umma: rows_per_group %u exceeds the %u-row tile", gd.rows_per_group, kTile);
        return;
    }
    const unsigned tiles = div_up(gd.n_points, kTile);
    unsigned blocks = min(div_up(tiles, kGroups), 148u);  // persistent: one CTA (three tile groups) per SM
    uint32_t *live_list = nullptr, *n_live = nullptr;
    if (gd.rows_per_group) {
        const unsigned n_groups = gd.n_points / gd.rows_per_group;
        auto *ws = static_cast<uint32_t *>(workspace(stream, ((size_t)n_groups + 1) * sizeof(uint32_t)));
        if (!ws) return;
        n_live = ws;
        live_list = ws + 1;
        NGP_CUDA_OK(cudaMemsetAsync(n_live, 0, sizeof(uint32_t), stream), "nerf_fused_forward_umma");
        if (n_groups == 0) return;
        compact_live_groups_kernel<<<div_up(n_groups, 256), 256, 0, stream>>>(n_groups, group_counts, live_list, n_live);
        blocks = min(div_up(div_up(n_groups, kTile / gd.rows_per_group), kGroups), 148u);
    }
#define NGP_FUSED_U(TT, WE)                                                                                            \
    do {                                                                                                               \
        static bool configured = false; /* benign race: idempotent */                                                  \
        if (!configured) {                                                                                             \
            cudaFuncSetAttribute(nerf_fused_forward_umma_kernel<TT, WE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdUmmaSmem); \
            configured = true;                                                                                         \
        }                                                                                                              \
        nerf_fused_forward_umma_kernel<TT, WE><<<blocks, kThreadsU, kFwdUmmaSmem, stream>>>(                            \
            *d, pos, static_cast<const TT *>(table), dirs, weights, group_counts, live_list, n_live, out, enc_out);    \
    } while (0)
    if (gd.table_dtype == 0) {
        if (d->write_enc) NGP_FUSED_U(float, true); else NGP_FUSED_U(float, false);
    } else {
        if (d->write_enc) NGP_FUSED_U(__half, true); else NGP_FUSED_U(__half, false);
    }
#undef NGP_FUSED_U
    check_launch("nerf_fused_forward_umma");
}
