// mlp_bwd_tc.cu -- backward of the NeRF MLP of make_nerf_ngp (models/nerfs.py:27-128,216-238,422-454) with EVERY
// matrix product on tcgen05: forward recompute, input-gradient chain and weight gradients.  Same contract and operand
// rounding as nerf_mlp_backward_umma_kernel (mlp.cu), whose recompute / delta chain still runs on mma.sync.
//
// What made the chain fit (DESIGN.md 9.1 counted 288 KB of shared memory for it): the chain's A operand never lives in
// shared memory.  tcgen05.mma takes A from TENSOR MEMORY (lane = row, one 32-bit column per K value), and that is the
// layout the previous layer's accumulator already has -- so the epilogue threads read their row of the accumulator
// (tcgen05.ld), apply the activation / ReLU mask, round to tf32 and write it back IN PLACE (tcgen05.st): the next
// chain MMA reads it from there.  Shared memory only holds what the weight gradients need -- activations and deltas
// as MN-major SWIZZLE_128B_BASE32B panels, sample-major rows written once by the thread that owns the sample -- and
// ONE 48 KB weight-operand region, which a bulk async copy (cp.async.bulk, no thread touches the data) re-fills from
// pre-swizzled global images twice per tile: [W as MN-major B] for the recompute, [W as K-major B] for the delta chain.
//
// Work unit: a tile of 128 samples, 256 threads.  Thread (half h, row r) owns columns 32h .. 32h+31 of row r of every
// 64-wide accumulator (warps w and w+4 reach the same TMEM lanes), so a 64-wide epilogue is 32 columns per thread.
//   tile:  enc -> TMEM | L0 | relu -> h0 | L1 | [x | SH] -> hin | L2 | relu -> h1 | L3 | relu -> h2 | L4 | sigmoid, d_a3
//          | B4 | mask h2 -> d_a2 | B3 | mask h1 -> d_a1 | B2 | + density term -> d_x | B1 | mask h0 -> d_a0 | B0 | d_enc out
//   with dW4 = h2^T d_a3, dW3 = h1^T d_a2, dW2^T = d_a1^T hin, dW1 = h0^T d_x, dW0^T = d_a0^T enc issued behind the
//   chain MMA of the same step (the tensor core runs them in issue order, under the next epilogue), accumulating in
//   five TMEM tiles that live for the CTA's lifetime (192 columns) and are flushed once with atomics.
#include "common.cuh"
#include "umma.cuh"

namespace ngp {
namespace {

constexpr uint32_t kTile = 128, kThreadsT = 256;
constexpr uint32_t kPanel = kTile * 128u;  // 16 KB: 128 samples x 32 features, MN-major rows
// shared memory (byte offsets from a 1024-byte aligned base)
constexpr uint32_t P_H0 = 0, P_HIN = P_H0 + 2 * kPanel, P_H1 = P_HIN + kPanel, P_H2 = P_H1 + 2 * kPanel,
                   P_DA = P_H2 + 2 * kPanel, P_S = P_DA + 2 * kPanel, P_W = P_S + kPanel, kWRegion = 49152,
                   kSmemBytes = P_W + kWRegion;  // 212,992
constexpr uint32_t P_ENC = P_H1;  // enc takes over h1's first panel once dW3 has read h1
// forward weight image: W[in][out] as MN-major B (identical to mlp_fwd_umma.cu's staging)
constexpr uint32_t WF0 = 0, WF1 = WF0 + 2 * 32 * 128, WF2 = WF1 + 64 * 128, WF3 = WF2 + 2 * 32 * 128, WF4 = WF3 + 2 * 64 * 128,
                   kWFBytes = WF4 + 64 * 128;  // 49152
// backward weight image: W[in][out] as K-major B (N = in rows, K = out in 32-wide panels)
constexpr uint32_t WB4 = 0, WB3 = WB4 + 64 * 128, WB2 = WB3 + 2 * 64 * 128, WB1 = WB2 + 2 * 16 * 128, WB0 = WB1 + 64 * 128,
                   kWBBytes = WB0 + 2 * 32 * 128;  // 45056
constexpr int G_W0 = 0, G_W1 = G_W0 + 32 * 64, G_W2 = G_W1 + 64 * 16, G_W3 = G_W2 + 32 * 64, G_W4 = G_W3 + 64 * 64;
// tensor memory columns: two chain regions (A of step k = D of step k-1, rewritten in place), then the dW accumulators
constexpr uint32_t T_R0 = 0, T_R1 = 64, T_ACC = 128, T_W3 = T_ACC + 0, T_W0 = T_ACC + 64, T_W2 = T_ACC + 96, T_W1 = T_ACC + 128,
                   T_W4 = T_ACC + 160, kTmemColsT = 512;

__device__ __forceinline__ uint32_t tf32r(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

// ---- pre-swizzled weight images (one tiny launch per backward call: the weights change every step)
__global__ void __launch_bounds__(256) mlp_weight_images_kernel(uint4 *__restrict__ images) {  // zero fill (padding columns / rows)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (kWFBytes + kWBBytes) / 16; i += gridDim.x * blockDim.x)
        images[i] = make_uint4(0u, 0u, 0u, 0u);
}
// second pass (after the zero fill): scatter the real values
struct WSpec { int g_off, K, N; uint32_t f_off, b_off; int b_rows; };  // W[K = in][N = out]
__global__ void __launch_bounds__(256) mlp_weight_scatter_kernel(const float *__restrict__ w, uint8_t *__restrict__ img_f,
                                                                 uint8_t *__restrict__ img_b) {
    const WSpec specs[5] = {{G_W0, 32, 64, WF0, WB0, 32}, {G_W1, 64, 16, WF1, WB1, 64}, {G_W2, 32, 64, WF2, WB2, 16},
                            {G_W3, 64, 64, WF3, WB3, 64}, {G_W4, 64, 3, WF4, WB4, 64}};
    const WSpec s = specs[blockIdx.x];
    for (int i = threadIdx.x; i < s.K * s.N; i += blockDim.x) {
        const int k = i / s.N, nn = i % s.N;  // in-feature k, out-feature nn
        const uint32_t v = tf32r(__ldg(w + s.g_off + i));
        // forward B, MN-major: row = in-feature, 32 out-features per row, panels of K rows per 32 out-features
        *reinterpret_cast<uint32_t *>(img_f + s.f_off + (nn >> 5) * (s.K * 128) + umma::panel_offset_mn32(k, nn & 31)) = v;
        // backward B, K-major: row = in-feature (only the first b_rows are used), 32 out-features per row
        if (k < s.b_rows)
            *reinterpret_cast<uint32_t *>(img_b + s.b_off + (nn >> 5) * (s.b_rows * 128) + umma::panel_offset(k, nn & 31)) = v;
    }
}

// real spherical harmonics, degree 4 (models/encoders.py:365-406; same expressions as mlp_fwd_umma.cu)
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

// row r of an MN-major panel set <- NV consecutive features starting at feature j0 (j0 and NV multiples of 8)
template <int NV>
__device__ __forceinline__ void store_row(uint8_t *panels, uint32_t r, uint32_t j0, const uint32_t (&v)[NV]) {
#pragma unroll
    for (int c = 0; c < NV / 8; ++c) {
        const uint32_t j = j0 + 8 * c;
        uint8_t *dst = panels + (j >> 5) * kPanel + r * 128u + ((((j & 31u) >> 3) ^ (r & 3u)) << 5);
        *reinterpret_cast<uint4 *>(dst) = make_uint4(v[8 * c + 0], v[8 * c + 1], v[8 * c + 2], v[8 * c + 3]);
        *reinterpret_cast<uint4 *>(dst + 16) = make_uint4(v[8 * c + 4], v[8 * c + 5], v[8 * c + 6], v[8 * c + 7]);
    }
}

// D[64 x N] (+)= A^T . B over the tile's 128 samples: A = 64-feature panel pair, B = N-feature panel(s), both MN-major
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool accumulate) {
#pragma unroll 4
    for (uint32_t ks = 0; ks < kTile / 8; ++ks)
        umma::mma_tf32(tmem_d, umma::desc_mn_major(a_addr, ks, kPanel), umma::desc_mn_major(b_addr, ks, kPanel), idesc,
                       accumulate || ks > 0);
}

// descriptors of the MMA operands are address + constant: the issuing thread adds the step to a precomputed base
__device__ __forceinline__ uint64_t desc_add(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

__global__ void __launch_bounds__(kThreadsT + 32, 1) nerf_mlp_backward_tc_kernel(uint32_t n, const float *__restrict__ enc,
                                                                                 const float *__restrict__ dirs,
                                                                                 const float *__restrict__ d_drgbs,
                                                                                 const uint8_t *__restrict__ img_f,
                                                                                 const uint8_t *__restrict__ img_b,
                                                                                 float *__restrict__ d_enc,
                                                                                 float *__restrict__ d_weights) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar_chain, bar_wt, bar_w4, bar_w3, bar_w0;
    __shared__ uint32_t tmem_slot;
    const uint32_t tid = threadIdx.x, half = (tid >> 7) & 1u, r = tid & 127u, warp = tid >> 5;
    // warps 0-7: epilogues (thread = row r, column half).  Warp 8: its first lane issues every MMA and bulk copy, so the
    // serial descriptor / issue work (24 instructions per step) is nobody's epilogue time.
    const bool is_issuer_warp = warp == 8;

    if (tid == 0) {
        umma::mbar_init(&bar_chain, 1);
        umma::mbar_init(&bar_wt, 1);
        umma::mbar_init(&bar_w4, 1);
        umma::mbar_init(&bar_w3, 1);
        umma::mbar_init(&bar_w0, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(&tmem_slot, kTmemColsT);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t n_tiles = (n + kTile - 1) / kTile;
    uint32_t it = 0;

    if (is_issuer_warp) {
        // ================================================================ MMA / copy issue
        // The whole warp walks the steps (bar.sync is warp-wide); lane 0 alone issues, the others fall through to the barrier.
        {
            const bool lead = (tid & 31u) == 0;
            const uint32_t sa = umma::smem_u32(sm), wa = sa + P_W;
            constexpr uint32_t kF64 = umma::make_idesc(128, 64, false, true), kF32 = umma::make_idesc(128, 32, false, true);
            constexpr uint32_t kB64 = umma::make_idesc(128, 64, false, false), kB32 = umma::make_idesc(128, 32, false, false),
                               kB16 = umma::make_idesc(128, 16, false, false);
            constexpr uint32_t kG32 = umma::make_idesc(64, 32, true, true), kG64 = umma::make_idesc(64, 64, true, true);
            const uint64_t dF0 = umma::desc_mn_major(wa + WF0, 0, 32 * 128), dF1 = umma::desc_mn_major(wa + WF1, 0, 64 * 128),
                           dF2 = umma::desc_mn_major(wa + WF2, 0, 32 * 128), dF3 = umma::desc_mn_major(wa + WF3, 0, 64 * 128),
                           dF4 = umma::desc_mn_major(wa + WF4, 0, 64 * 128);
            const uint64_t dB4 = umma::desc_k_major(wa + WB4, 0), dB3 = umma::desc_k_major(wa + WB3, 0), dB2 = umma::desc_k_major(wa + WB2, 0),
                           dB1 = umma::desc_k_major(wa + WB1, 0), dB0 = umma::desc_k_major(wa + WB0, 0);
            const uint64_t pH0 = umma::desc_mn_major(sa + P_H0, 0, kPanel), pHIN = umma::desc_mn_major(sa + P_HIN, 0, kPanel),
                           pH1 = umma::desc_mn_major(sa + P_H1, 0, kPanel), pH2 = umma::desc_mn_major(sa + P_H2, 0, kPanel),
                           pDA = umma::desc_mn_major(sa + P_DA, 0, kPanel), pS = umma::desc_mn_major(sa + P_S, 0, kPanel),
                           pENC = umma::desc_mn_major(sa + P_ENC, 0, kPanel);
            // forward chain step: D[128 x N] = A[128 x 8*KS] (TMEM) . W (MN-major B: 8 K-rows = 1024 bytes per step)
            auto fwd = [&](uint32_t d_col, uint32_t a_col, uint64_t db, uint32_t KS, uint32_t idesc) {
                if (!lead) return;
                for (uint32_t ks = 0; ks < KS; ++ks)
                    umma::mma_tf32_ts(tmem + d_col, tmem + a_col + 8 * ks, desc_add(db, ks * 1024u), idesc, ks > 0);
                umma::commit(&bar_chain);
            };
            // delta chain step: W as K-major B: 8 K-values = 32 bytes per step, 32-wide K panels `panel` bytes apart
            auto bwd = [&](uint32_t d_col, uint32_t a_col, uint64_t db, uint32_t KS, uint32_t panel, uint32_t idesc) {
                if (!lead) return;
                for (uint32_t ks = 0; ks < KS; ++ks)
                    umma::mma_tf32_ts(tmem + d_col, tmem + a_col + 8 * ks, desc_add(db, (ks >> 2) * panel + (ks & 3u) * 32u), idesc, ks > 0);
                umma::commit(&bar_chain);
            };
            // D[64 x N] (+)= A^T . B over the tile's 128 samples, both operands MN-major panels
            auto wgrad = [&](uint32_t d_col, uint64_t da, uint64_t db, uint32_t idesc, bool acc) {
                if (!lead) return;
#pragma unroll 4
                for (uint32_t ks = 0; ks < kTile / 8; ++ks)
                    umma::mma_tf32(tmem + d_col, desc_add(da, ks * 1024u), desc_add(db, ks * 1024u), idesc, acc || ks > 0);
            };
            uint32_t ph_chain = 0, ph_wt = 0;
            auto refill = [&](const uint8_t *image, uint32_t bytes) {  // async copy of a weight image into the operand region
                if (!lead) return;
                umma::mbar_expect_tx(&bar_wt, bytes);
                umma::bulk_copy_g2s(sm + P_W, image, bytes, &bar_wt);
            };
            auto wait_on = [&](uint64_t *bar, uint32_t parity) {
                if (lead) umma::mbar_wait(bar, parity);
            };
            auto commit_to = [&](uint64_t *bar) {
                if (lead) umma::commit(bar);
            };
            if (blockIdx.x < n_tiles) refill(img_f, kWFBytes);  // forward weights of the first tile
            auto step_barrier = [&]() {
                __syncwarp();
                asm volatile("bar.sync 1, %0;" ::"n"(kThreadsT + 32) : "memory");
                umma::fence_after_sync();
            };
            for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const bool acc = it > 0;
                step_barrier();  // enc in R0
                wait_on(&bar_wt, ph_wt); ph_wt ^= 1u;  // forward weights have landed
                fwd(T_R1, T_R0, dF0, 4, kF64); ph_chain ^= 1u;                      // L0
                step_barrier();  // h0 in R1
                fwd(T_R0, T_R1, dF1, 8, kF32); ph_chain ^= 1u;                      // L1
                step_barrier();  // hin in R0
                fwd(T_R1, T_R0, dF2, 4, kF64); ph_chain ^= 1u;                      // L2
                step_barrier();  // h1 in R1
                fwd(T_R0, T_R1, dF3, 8, kF64); ph_chain ^= 1u;                      // L3
                step_barrier();  // h2 in R0
                fwd(T_R1, T_R0, dF4, 8, kF32);                                      // L4
                wait_on(&bar_chain, ph_chain); ph_chain ^= 1u;
                refill(img_b, kWBBytes);  // every forward MMA is done: the region takes the delta chain's operands
                step_barrier();  // d_a3 in R1 (its panel row follows the barrier)
                wait_on(&bar_wt, ph_wt); ph_wt ^= 1u;
                bwd(T_R0, T_R1, dB4, 1, 0, kB64); ph_chain ^= 1u;                   // B4
                step_barrier();  // d_a2 in R0; panels h2, d_a3 complete
                bwd(T_R1, T_R0, dB3, 8, 64 * 128, kB64); ph_chain ^= 1u;            // B3
                wgrad(T_W4, pH2, pS, kG32, acc);                                    // dW4 = h2^T . d_a3
                commit_to(&bar_w4);
                step_barrier();  // d_a1 in R1; panel d_a2 complete
                bwd(T_R0, T_R1, dB2, 8, 16 * 128, kB16); ph_chain ^= 1u;            // B2
                wgrad(T_W3, pH1, pDA, kG64, acc);                                   // dW3 = h1^T . d_a2
                commit_to(&bar_w3);
                step_barrier();  // d_x in R0; panel d_a1 (in h2's place) complete
                bwd(T_R1, T_R0, dB1, 2, 0, kB64); ph_chain ^= 1u;                   // B1
                wgrad(T_W2, pH2, pHIN, kG32, acc);                                  // dW2^T = d_a1^T . hin
                step_barrier();  // d_a0 in R1; panel d_x complete
                bwd(T_R0, T_R1, dB0, 8, 32 * 128, kB32);                            // B0
                wgrad(T_W1, pH0, pS, kG32, acc);                                    // dW1 = h0^T . d_x
                wait_on(&bar_chain, ph_chain); ph_chain ^= 1u;
                if (tile + gridDim.x < n_tiles) refill(img_f, kWFBytes);  // the delta chain is done: forward weights for the next tile
                step_barrier();  // end of tile; panels d_a0, enc complete
                wgrad(T_W0, pDA, pENC, kG32, acc);                                  // dW0^T = d_a0^T . enc
                commit_to(&bar_w0);
            }
        }
    } else {
        // ================================================================ epilogues: thread = (row r, column half)
        const uint32_t lane_addr = tmem + ((32u * (warp & 3u)) << 16);  // this thread's TMEM lane (tcgen05.ld / st address)
        uint32_t ph_chain = 0;
        // everything the epilogue wrote becomes visible to the tensor core, then the issuer launches the next MMAs.
        // Only lane 0 of each epilogue warp and the issuing lane arrive... (bar.sync counts threads: all 256 + the issuer's warp)
        auto publish = [&]() {
            umma::tmem_st_wait();
            umma::fence_smem_to_async();
            umma::fence_before_sync();
            asm volatile("bar.sync 1, %0;" ::"n"(kThreadsT + 32) : "memory");
        };
        auto await_chain = [&]() {
            umma::mbar_wait(&bar_chain, ph_chain);
            ph_chain ^= 1u;
            umma::fence_after_sync();
        };
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t row = tile * kTile + r;
            const bool live = row < n;
            uint32_t v[32];
            // ---- inputs.  half 0: the sample's encoding (kept in registers as tf32 bits until dW0) and dL/d(out);
            //      half 1: the direction
            uint32_t e[32];
            float4 dd = make_float4(0.f, 0.f, 0.f, 0.f);
            float dir[3] = {0.f, 0.f, 1.f};
            if (half == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 q = live ? __ldg(reinterpret_cast<const float4 *>(enc + (size_t)row * 32) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    e[4 * c + 0] = tf32r(q.x); e[4 * c + 1] = tf32r(q.y); e[4 * c + 2] = tf32r(q.z); e[4 * c + 3] = tf32r(q.w);
                }
                umma::tmem_st32(lane_addr + T_R0, e);
                if (live) dd = __ldg(reinterpret_cast<const float4 *>(d_drgbs) + row);
            } else if (live) {
                dir[0] = __ldg(dirs + (size_t)row * 3 + 0);
                dir[1] = __ldg(dirs + (size_t)row * 3 + 1);
                dir[2] = __ldg(dirs + (size_t)row * 3 + 2);
            }
            publish();      // -> L0: enc[32] (R0) . W0 -> R1[64]
            await_chain();
            uint32_t mask0 = 0, mask1 = 0, mask2 = 0;  // ReLU masks of this thread's 32 columns of h0, h1, h2
            auto relu_step = [&](uint32_t region, uint32_t panel, uint32_t &mask) {
                umma::tmem_ld32(lane_addr + region + 32 * half, v);
                umma::tmem_ld_wait();
                mask = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float x = __uint_as_float(v[j]);
                    mask |= (x > 0.f ? 1u : 0u) << j;
                    v[j] = tf32r(fmaxf(x, 0.f));
                }
                umma::tmem_st32(lane_addr + region + 32 * half, v);
            };
            // The MN-major panel row (the weight gradients' operand) is written AFTER the barrier that releases the next
            // chain MMA: the shared-memory stores and their proxy fence run under that MMA instead of ahead of it.  The
            // weight gradient that reads a panel is issued one barrier later (see the issuing warp's schedule).
            relu_step(T_R1, P_H0, mask0);
            publish();      // -> L1: h0[64] (R1) . W1 -> R0[32] (16 real): x; density logit x0; hin = [x | SH4(dir)]
            if (it > 0) umma::mbar_wait(&bar_w0, (it - 1u) & 1u);  // the previous tile's weight gradients have read every panel
            store_row<32>(sm + P_H0, r, 32 * half, v);
            await_chain();
            float x0 = 0.f;
            {
                uint32_t q[16];
                if (half == 0) {
                    umma::tmem_ld16(lane_addr + T_R0, q);
                    umma::tmem_ld_wait();
                    x0 = __uint_as_float(q[0]);
#pragma unroll
                    for (int j = 0; j < 16; ++j) q[j] = tf32r(__uint_as_float(q[j]));
                    umma::tmem_st16(lane_addr + T_R0, q);
                } else {
                    float s[16];
                    sh16(dir[0], dir[1], dir[2], s);
#pragma unroll
                    for (int j = 0; j < 16; ++j) q[j] = tf32r(s[j]);
                    umma::tmem_st16(lane_addr + T_R0 + 16, q);
                }
                publish();  // -> L2: hin[32] (R0) . W2 -> R1[64]
                store_row<16>(sm + P_HIN, r, 16 * half, q);
            }
            await_chain();
            relu_step(T_R1, P_H1, mask1);
            publish();      // -> L3: h1[64] (R1) . W3 -> R0[64]
            store_row<32>(sm + P_H1, r, 32 * half, v);
            await_chain();
            relu_step(T_R0, P_H2, mask2);
            publish();      // -> L4: h2[64] (R0) . W4 -> R1[32] (3 real): rgb = sigmoid; d_a3 = dL/drgb * rgb * (1 - rgb)
            store_row<32>(sm + P_H2, r, 32 * half, v);
            await_chain();
            {
                uint32_t q[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) q[j] = 0u;
                if (half == 0) {
                    uint32_t o[3];
                    {
                        uint32_t t16[16];
                        umma::tmem_ld16(lane_addr + T_R1, t16);
                        umma::tmem_ld_wait();
                        o[0] = t16[0]; o[1] = t16[1]; o[2] = t16[2];
                    }
                    const float dr[3] = {dd.y, dd.z, dd.w};
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const float s = 1.f / (1.f + expf(-__uint_as_float(o[j])));
                        q[j] = tf32r(dr[j] * s * (1.f - s));
                    }
                    const uint32_t a8[8] = {q[0], q[1], q[2], 0u, 0u, 0u, 0u, 0u};
                    umma::tmem_st8(lane_addr + T_R1, a8);
                }
                publish();  // -> B4: d_a3[8] (R1) . W4^T -> R0[64]
                store_row<16>(sm + P_S, r, 16 * half, q);  // the narrow delta panel: [d_a3 (3) | zeros]
            }
            await_chain();
            auto mask_step = [&](uint32_t region, uint32_t mask) {  // delta = accumulator masked by the layer's ReLU, in place
                umma::tmem_ld32(lane_addr + region + 32 * half, v);
                umma::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = ((mask >> j) & 1u) ? tf32r(__uint_as_float(v[j])) : 0u;
                umma::tmem_st32(lane_addr + region + 32 * half, v);
            };
            mask_step(T_R0, mask2);  // d_a2
            publish();      // -> B3: d_a2[64] (R0) . W3^T -> R1[64];   then dW4 = h2^T . d_a3
            store_row<32>(sm + P_DA, r, 32 * half, v);
            await_chain();
            mask_step(T_R1, mask1);  // d_a1 -> takes over the h2 panels once dW4 has read them
            publish();      // -> B2: d_a1[64] (R1) . W2^T[:, :16] -> R0[16];   then dW3 = h1^T . d_a2
            umma::mbar_wait(&bar_w4, it & 1u);
            store_row<32>(sm + P_H2, r, 32 * half, v);
            await_chain();
            {   // d_x = (d_a1 . W2^T)[:, :16] + dL/d(density) * exp(clip(x0, -15, 15)) on column 0 (nerfs.py:231-234)
                uint32_t q[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) q[j] = 0u;
                if (half == 0) {
                    umma::tmem_ld16(lane_addr + T_R0, q);
                    umma::tmem_ld_wait();
                    q[0] = __float_as_uint(__uint_as_float(q[0]) + dd.x * expf(fminf(fmaxf(x0, -15.f), 15.f)));
#pragma unroll
                    for (int j = 0; j < 16; ++j) q[j] = tf32r(__uint_as_float(q[j]));
                    umma::tmem_st16(lane_addr + T_R0, q);
                }
                publish();  // -> B1: d_x[16] (R0) . W1^T -> R1[64];   then dW2^T = d_a1^T . hin
                store_row<16>(sm + P_S, r, 16 * half, q);  // [d_x (16) | zeros]  (dW4 has read the panel: bar_w4 above)
            }
            await_chain();
            mask_step(T_R1, mask0);  // d_a0 -> takes over the d_a2 panels; enc takes over h1's first panel (dW3 has read both)
            publish();      // -> B0: d_a0[64] (R1) . W0^T -> R0[32];   then dW1 = h0^T . d_x
            umma::mbar_wait(&bar_w3, it & 1u);
            store_row<32>(sm + P_DA, r, 32 * half, v);
            if (half == 0) store_row<32>(sm + P_ENC, r, 0, e);
            await_chain();
            {   // d_enc: 16 columns per thread
                uint32_t q[16];
                umma::tmem_ld16(lane_addr + T_R0 + 16 * half, q);
                umma::tmem_ld_wait();
                if (live) {
                    float4 *dst = reinterpret_cast<float4 *>(d_enc + (size_t)row * 32 + 16 * half);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        dst[c] = make_float4(__uint_as_float(q[4 * c]), __uint_as_float(q[4 * c + 1]), __uint_as_float(q[4 * c + 2]),
                                             __uint_as_float(q[4 * c + 3]));
                }
            }
            // end of tile: the last panel rows (d_a0, enc) become visible to the tensor core for dW0, and the next tile's
            // first tcgen05.st (half 0) may overwrite the columns the other half's loads just read from the same lanes
            umma::fence_smem_to_async();
            umma::fence_before_sync();
            asm volatile("bar.sync 1, %0;" ::"n"(kThreadsT + 32) : "memory");
            umma::fence_after_sync();
        }
    }
    // the issuing warp's idle lanes and every epilogue thread meet here; `it` = tiles this CTA processed
    it = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;

    // ---- flush the CTA's weight gradients: row m of an M = 64 accumulator sits in lane (m % 16) + 32 (m / 16)
    if (it > 0) {
        umma::mbar_wait(&bar_w0, (it - 1u) & 1u);
        umma::fence_after_sync();
        const uint32_t lane = tid & 31u, q = warp & 3u, m = 16u * q + lane;
        const uint32_t taddr = tmem + ((32u * q) << 16);
        const bool mine = lane < 16u;
        uint32_t v[32];
        if (warp >= 8) {
            // the issuing warp holds no accumulator lanes of its own
        } else if (warp < 4) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // dW3[m][32h + j]
                umma::tmem_ld32(taddr + T_W3 + 32 * h, v);
                umma::tmem_ld_wait();
                if (mine)
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(d_weights + G_W3 + m * 64 + 32 * h + j, __uint_as_float(v[j]));
            }
            umma::tmem_ld32(taddr + T_W0, v);  // dW0^T[m = out][j = in]
            umma::tmem_ld_wait();
            if (mine)
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(d_weights + G_W0 + j * 64 + m, __uint_as_float(v[j]));
        } else {
            umma::tmem_ld32(taddr + T_W2, v);  // dW2^T[m = out][j = in]
            umma::tmem_ld_wait();
            if (mine)
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(d_weights + G_W2 + j * 64 + m, __uint_as_float(v[j]));
            umma::tmem_ld32(taddr + T_W1, v);  // dW1[m][j < 16]
            umma::tmem_ld_wait();
            if (mine)
#pragma unroll
                for (int j = 0; j < 16; ++j) atomicAdd(d_weights + G_W1 + m * 16 + j, __uint_as_float(v[j]));
            umma::tmem_ld32(taddr + T_W4, v);  // dW4[m][j < 3]
            umma::tmem_ld_wait();
            if (mine)
#pragma unroll
                for (int j = 0; j < 3; ++j) atomicAdd(d_weights + G_W4 + m * 3 + j, __uint_as_float(v[j]));
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, kTmemColsT);
}

}  // namespace
}  // namespace ngp

extern "C" void ngp_nerf_mlp_backward_tc(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpNerfMlpDescriptor>(opaque, opaque_len, "nerf_mlp_backward_tc");
    if (!d) return;
    if (d->density_only || d->rows_per_group) {
        set_error(NGP_ERR_ARGUMENT, "nerf_mlp_backward_tc: the backward takes the plain layout (density_only = rows_per_group = 0)");
        return;
    }
    BufferCursor b{buffers};
    const float *enc = b.next<const float>();
    const float *dirs = b.next<const float>();
    const float *weights = b.next<const float>();
    const float *d_drgbs = b.next<const float>();
    float *d_enc = b.next<float>();
    float *d_weights = b.next<float>();
    NGP_CUDA_OK(cudaMemsetAsync(d_weights, 0, 9408 * sizeof(float), stream), "nerf_mlp_backward_tc");
    if (d->n_samples == 0) return;
    auto *ws = static_cast<uint8_t *>(workspace(stream, kWFBytes + kWBBytes));
    if (!ws) return;
    uint8_t *img_f = ws, *img_b = ws + kWFBytes;
    mlp_weight_images_kernel<<<12, 256, 0, stream>>>(reinterpret_cast<uint4 *>(ws));
    mlp_weight_scatter_kernel<<<5, 256, 0, stream>>>(weights, img_f, img_b);
    if (!check_launch("nerf_mlp_backward_tc(weight images)")) return;
    static bool configured = false;  // benign race: idempotent
    if (!configured) {
        cudaFuncSetAttribute(nerf_mlp_backward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemBytes + 1024));
        configured = true;
    }
    const unsigned tiles = div_up(d->n_samples, kTile);
    nerf_mlp_backward_tc_kernel<<<min(tiles, 148u), kThreadsT + 32, kSmemBytes + 1024, stream>>>(d->n_samples, enc, dirs, d_drgbs, img_f, img_b,
                                                                                           d_enc, d_weights);
    check_launch("nerf_mlp_backward_tc");
}
