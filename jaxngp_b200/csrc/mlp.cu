// mlp.cu -- fully fused NeRF MLP of make_nerf_ngp (models/nerfs.py:27-128,216-238,422-454), forward and
// backward, on the tensor cores (TF32 mma.sync m16n8k8, f32 accumulate -- the precision XLA gives the
// reference's f32 Dense layers on Ampere+ GPUs).
//
//   enc[32] -> W0[32x64] -> ReLU -> W1[64x16] = x ; density = exp(x[0])        (trunc_exp, nerfs.py:222-238)
//   [x(16) | SH4(dir)(16)] -> W2[32x64] -> ReLU -> W3[64x64] -> ReLU -> W4[64x3] -> sigmoid = rgb
//
// Forward: a warp carries 16 samples through all five layers in registers.  The accumulator fragment of
// one layer IS the A fragment of the next (columns 2t,2t+1 of every 8-block feed k = t, t+4), which only
// needs the weight rows permuted the same way when B fragments are read from shared memory -- no
// shuffles, no shared-memory round trip for activations.  Weights (9408 values) sit in shared memory.
//
// Backward: one CTA owns 128 samples.  Phase 1 (per warp, 16 samples): recompute the forward, keep the
// activations of the 128 samples in shared memory, and walk the layers back producing one delta matrix
// at a time.  Phase 2 (whole CTA, after each delta is written): dW += act^T * delta with the CTA's
// 128 samples as the reduction dimension; every warp owns a fixed set of dW tiles that stay in its
// registers for the whole kernel (76 tiles over 8 warps = 40 registers/thread) and are flushed to the
// global gradient with one atomicAdd per element per CTA at the end.  Activations never touch HBM:
// traffic is enc (128 B) + dirs (12 B) + d_drgbs (16 B) in, d_enc (128 B) out per sample.
#include "common.cuh"
#include "hashgrid.cuh"
#include "umma.cuh"

namespace ngp {
namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kBlockSamples = kWarps * 16;  // 128

// shared-memory weight layout: row stride = N + 4 (mod 32 == 4 -> forward B-fragment loads are conflict free)
constexpr int S_W0 = 68, S_W1 = 20, S_W2 = 68, S_W3 = 68, S_W4 = 12;
constexpr int O_W0 = 0, O_W1 = O_W0 + 32 * S_W0, O_W2 = O_W1 + 64 * S_W1, O_W3 = O_W2 + 32 * S_W2,
              O_W4 = O_W3 + 64 * S_W3, kWeightFloats = O_W4 + 64 * S_W4;  // 10752
// global flat weight layout [W0 | W1 | W2 | W3 | W4], row-major [in][out] (flax Dense kernels)
constexpr int G_W0 = 0, G_W1 = G_W0 + 32 * 64, G_W2 = G_W1 + 64 * 16, G_W3 = G_W2 + 32 * 64, G_W4 = G_W3 + 64 * 64,
              kGlobalWeights = G_W4 + 64 * 3;  // 9408
// activation buffers of the backward kernel: row stride = K + 8 (mod 32 == 8 -> fragment stores/loads conflict free)
constexpr int S_A64 = 72, S_A32 = 40, S_D = 72;
constexpr int O_H0 = 0, O_HIN = O_H0 + kBlockSamples * S_A64, O_H1 = O_HIN + kBlockSamples * S_A32,
              O_H2 = O_H1 + kBlockSamples * S_A64, O_D = O_H2 + kBlockSamples * S_A64,
              kActFloats = O_D + kBlockSamples * S_D;

// ---- tcgen05 wgrad path of the backward kernel: activations / deltas of the CTA's 128 samples as MN-major tf32
// panels (umma.cuh: rows = sample, 32 features per 128-byte row, SWIZZLE_128B_BASE32B), byte offsets from a
// 1024-byte aligned base.  PB_S holds the two narrow deltas (d_a3: 3 of 32 columns, d_x: 16 of 32), zero elsewhere.
constexpr uint32_t kPanel = kBlockSamples * 128u;  // 16 KB: 128 samples x 32 features
constexpr uint32_t PB_ENC = 0, PB_H0 = PB_ENC + kPanel, PB_HIN = PB_H0 + 2 * kPanel, PB_H1 = PB_HIN + kPanel,
                   PB_H2 = PB_H1 + 2 * kPanel, PB_S = PB_H2 + 2 * kPanel, PB_DA = PB_S + kPanel,
                   kPanelBytes = PB_DA + 2 * kPanel;  // 180224
// TMEM columns of the weight-gradient accumulators (f32, M = 64 rows each, live for the whole kernel)
constexpr uint32_t T_W3 = 0, T_W0 = 64, T_W2 = 96, T_W1 = 128, T_W4 = 160, kTmemCols = 256;

// byte offset of this lane's float2 (row 16*warp + g, columns col .. col+1 with col = 8*block + 2t) in a panel set;
// row g + 8 is 1024 bytes further (same swizzle phase)
__device__ __forceinline__ uint32_t panel_frag_offset(uint32_t warp, uint32_t g, uint32_t t, int col8) {
    return (uint32_t)(col8 >> 5) * kPanel + (16u * warp + g) * 128u + (((((uint32_t)col8 >> 3) & 3u) ^ (g & 3u)) << 5) + t * 8u;
}
// store A-fragment-ordered tf32 bits (c_to_a layout: [0] = (g, 2t), [1] = (g+8, 2t), [2] = (g, 2t+1), [3] = (g+8, 2t+1))
template <int NT>
__device__ __forceinline__ void store_frag_panel(uint8_t *__restrict__ panel, const uint32_t (&a)[NT][4], int col0,
                                                 uint32_t warp, uint32_t g, uint32_t t) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const uint32_t off = panel_frag_offset(warp, g, t, col0 + 8 * nt);
        *reinterpret_cast<uint2 *>(panel + off) = make_uint2(a[nt][0], a[nt][2]);
        *reinterpret_cast<uint2 *>(panel + off + 1024u) = make_uint2(a[nt][1], a[nt][3]);
    }
}

// round to tf32: nearest, ties away from zero -- cvt.rna.tf32.f32 for every finite input, in two integer
// instructions instead of the four the conversion compiles to (it special-cases NaN/Inf, which never occur here)
__device__ __forceinline__ uint32_t tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// accumulator fragment (rows g, g+8; cols 2t, 2t+1) -> A fragment under the k-permutation t <-> 2t, t+4 <-> 2t+1
__device__ __forceinline__ void c_to_a(const float (&c)[4], uint32_t (&a)[4]) {
    a[0] = tf32(c[0]);
    a[1] = tf32(c[2]);
    a[2] = tf32(c[1]);
    a[3] = tf32(c[3]);
}

// out[16 x 8*NT] = in[16 x 8*KT] * W, W in shared memory [8*KT][STRIDE] (tf32 bits)
template <int KT, int NT, int STRIDE>
__device__ __forceinline__ void layer_forward(const uint32_t (&a)[KT][4], const uint32_t *__restrict__ W,
                                              float (&acc)[NT][4], uint32_t g, uint32_t t) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
        const uint32_t *w0 = W + (8 * kt + 2 * t) * STRIDE + g;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt], a[kt], w0[8 * nt], w0[STRIDE + 8 * nt]);
    }
}

// d_in[16 x 8*NT_IN] = d_out[16 x 8*KT_OUT] * W^T, same shared-memory W [8*NT_IN.. rows][STRIDE]
template <int KT_OUT, int NT_IN, int STRIDE>
__device__ __forceinline__ void layer_backward(const uint32_t (&a)[KT_OUT][4], const uint32_t *__restrict__ W,
                                               float (&acc)[NT_IN][4], uint32_t g, uint32_t t) {
#pragma unroll
    for (int nt = 0; nt < NT_IN; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    // kt outer, nt inner: consecutive MMAs go to different accumulators (no back-to-back dependency)
    const uint32_t *w0 = W + g * STRIDE + 2 * t;
#pragma unroll
    for (int kt = 0; kt < KT_OUT; ++kt) {
#pragma unroll
        for (int nt = 0; nt < NT_IN; ++nt) {
            const uint2 b = *reinterpret_cast<const uint2 *>(w0 + 8 * nt * STRIDE + 8 * kt);
            mma_tf32(acc[nt], a[kt], b.x, b.y);
        }
    }
}

// store an accumulator tile set into a [16 rows of this warp][stride] shared buffer
template <int NT, int STRIDE>
__device__ __forceinline__ void store_frag(float *__restrict__ buf, const float (&acc)[NT][4], int col0, uint32_t g,
                                           uint32_t t) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        *reinterpret_cast<float2 *>(buf + g * STRIDE + col0 + 8 * nt + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<float2 *>(buf + (g + 8) * STRIDE + col0 + 8 * nt + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
    }
}

// SH degree 4, the four coefficients this lane feeds into the second half of hin: indices 2t, 2t+1, 8+2t, 9+2t
// (same basis, order and signs as models/encoders.py:365-406)
__device__ __forceinline__ void sh4_lane(float x, float y, float z, uint32_t t, float (&o)[4]) {
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    const float s0 = 0.28209479177387814f, s1 = -0.48860251190291987f * y, s2 = 0.48860251190291987f * z,
                s3 = -0.48860251190291987f * x, s4 = 1.0925484305920792f * xy, s5 = -1.0925484305920792f * yz,
                s6 = 0.94617469575755997f * z2 - 0.31539156525251999f, s7 = -1.0925484305920792f * xz,
                s8 = 0.54627421529603959f * x2 - 0.54627421529603959f * y2,
                s9 = 0.59004358992664352f * y * (-3.0f * x2 + y2), s10 = 2.8906114426405538f * xy * z,
                s11 = 0.45704579946446572f * y * (1.0f - 5.0f * z2), s12 = 0.3731763325901154f * z * (5.0f * z2 - 3.0f),
                s13 = 0.45704579946446572f * x * (1.0f - 5.0f * z2), s14 = 1.4453057213202769f * z * (x2 - y2),
                s15 = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
    o[0] = t == 0 ? s0 : t == 1 ? s2 : t == 2 ? s4 : s6;
    o[1] = t == 0 ? s1 : t == 1 ? s3 : t == 2 ? s5 : s7;
    o[2] = t == 0 ? s8 : t == 1 ? s10 : t == 2 ? s12 : s14;
    o[3] = t == 0 ? s9 : t == 1 ? s11 : t == 2 ? s13 : s15;
}

// Stage the 9,408 weights as tf32 bits in the padded shared layout.  Every thread issues ALL of its (up to 37) loads before
// the first conversion: as five loops of load -> convert -> store this prologue was 38 dependent trips to L2 (~14 us, 9 % of
// the backward kernel's stall samples).  The matrices start at multiples of kThreads floats, so for a given k every thread
// is in the same matrix and the address arithmetic below resolves at compile time.
__device__ __forceinline__ void load_weights(uint32_t *__restrict__ sw, const float *__restrict__ w) {
    static_assert(G_W1 % kThreads == 0 && G_W2 % kThreads == 0 && G_W3 % kThreads == 0 && G_W4 % kThreads == 0, "matrix boundaries");
    constexpr int kPer = (kGlobalWeights + kThreads - 1) / kThreads;  // 37
    const int tid = threadIdx.x;
    if (tid >= kThreads) return;  // a kernel with extra (non-chain) warps: they do not take part
    float v[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int i = tid + k * kThreads;
        v[k] = i < kGlobalWeights ? __ldg(w + i) : 0.f;
    }
    // W4 is padded from 3 to 8 output columns with zeros
    for (int j = tid; j < 64 * 5; j += kThreads) sw[O_W4 + (j / 5) * S_W4 + 3 + j % 5] = 0u;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int i = tid + k * kThreads;
        const uint32_t x = tf32(v[k]);
        if (k * kThreads < G_W1) sw[O_W0 + (i / 64) * S_W0 + i % 64] = x;
        else if (k * kThreads < G_W2) sw[O_W1 + ((i - G_W1) / 16) * S_W1 + (i - G_W1) % 16] = x;
        else if (k * kThreads < G_W3) sw[O_W2 + ((i - G_W2) / 64) * S_W2 + (i - G_W2) % 64] = x;
        else if (k * kThreads < G_W4) sw[O_W3 + ((i - G_W3) / 64) * S_W3 + (i - G_W3) % 64] = x;
        else if (i < kGlobalWeights) sw[O_W4 + ((i - G_W4) / 3) * S_W4 + (i - G_W4) % 3] = x;
    }
}

struct FwdState {  // what the backward needs from the recomputed forward, in fragment layout
    float x0[2];     // x[:, 0] of rows g, g+8 (valid on lanes t == 0)
    float rgb[2][2]; // sigmoid outputs: cols 2t, 2t+1 of rows g, g+8 (t == 0: r, g; t == 1: b, pad)
};

// enc A fragments of a 16-row tile: cols 8kt+2t, 8kt+2t+1 of rows g, g+8 (float2 loads, every 32 B sector fully used)
__device__ __forceinline__ void load_enc_fragments(const float *__restrict__ enc, uint32_t r_lo, uint32_t r_hi, bool ok_lo,
                                                   bool ok_hi, uint32_t t, uint32_t (&a_in)[4][4]) {
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
        const float2 lo = ok_lo ? __ldg(reinterpret_cast<const float2 *>(enc + (size_t)r_lo * 32 + 8 * kt + 2 * t)) : make_float2(0.f, 0.f);
        const float2 hi = ok_hi ? __ldg(reinterpret_cast<const float2 *>(enc + (size_t)r_hi * 32 + 8 * kt + 2 * t)) : make_float2(0.f, 0.f);
        a_in[kt][0] = tf32(lo.x);
        a_in[kt][1] = tf32(hi.x);
        a_in[kt][2] = tf32(lo.y);
        a_in[kt][3] = tf32(hi.y);
    }
}

// Forward for this warp's 16 rows starting at `row0` (global sample index) from the enc fragments `a_in`.
// kKeep == 1: activations are also written to the warp's rows of the padded shared activation buffers (act, f32);
// kKeep == 2: ... as tf32 bits into the MN-major panels of the tcgen05 wgrad path (act = panel base).
template <int kKeep, bool kDensityOnly>
__device__ __forceinline__ void warp_forward_from(const uint32_t *__restrict__ sw, float *__restrict__ act, uint32_t warp,
                                                  uint32_t row0, uint32_t n, const uint32_t (&a_in)[4][4],
                                                  const float *__restrict__ dirs, uint32_t g, uint32_t t, FwdState &st,
                                                  uint32_t (&a_h2)[8][4], float (&out_rgb)[1][4],
                                                  uint32_t rows_per_group = 0, bool ok_lo_in = true, bool ok_hi_in = true,
                                                  const float *__restrict__ pre_dirs = nullptr) {
    const uint32_t r_lo = row0 + g, r_hi = row0 + g + 8;
    const bool ok_lo = r_lo < n && ok_lo_in, ok_hi = r_hi < n && ok_hi_in;
    // grouped layout: one direction per group of rows (ray), else one per row
    const uint32_t d_lo = rows_per_group ? r_lo / rows_per_group : r_lo, d_hi = rows_per_group ? r_hi / rows_per_group : r_hi;
    float *my_rows_64 = nullptr, *my_rows_32 = nullptr;
    (void)my_rows_64;
    (void)my_rows_32;
    // layer 0: 32 -> 64, ReLU
    uint32_t a_h0[8][4];
    {
        float acc[8][4];
        layer_forward<4, 8, S_W0>(a_in, sw + O_W0, acc, g, t);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[nt][i] = fmaxf(acc[nt][i], 0.f);
            c_to_a(acc[nt], a_h0[nt]);
        }
        if (kKeep == 1) store_frag<8, S_A64>(act + O_H0 + warp * 16 * S_A64, acc, 0, g, t);
        if (kKeep == 2) store_frag_panel<8>(reinterpret_cast<uint8_t *>(act) + PB_H0, a_h0, 0, warp, g, t);
    }
    // layer 1: 64 -> 16 (no activation); density = exp(x[0])
    uint32_t a_hin[4][4];
    {
        float acc[2][4];
        layer_forward<8, 2, S_W1>(a_h0, sw + O_W1, acc, g, t);
        st.x0[0] = acc[0][0];
        st.x0[1] = acc[0][2];
        c_to_a(acc[0], a_hin[0]);
        c_to_a(acc[1], a_hin[1]);
        if (kKeep == 1) store_frag<2, S_A32>(act + O_HIN + warp * 16 * S_A32, acc, 0, g, t);
    }
    if (kDensityOnly) return;
    // direction encoding: SH degree 4 into columns 16..31 of hin
    {
        float sh_lo[4], sh_hi[4];
        // pre_dirs: the six values below, already in registers (the backward kernel fetches them a block ahead)
        const float dx0 = pre_dirs ? pre_dirs[0] : ok_lo ? __ldg(dirs + (size_t)d_lo * 3 + 0) : 0.f,
                    dy0 = pre_dirs ? pre_dirs[1] : ok_lo ? __ldg(dirs + (size_t)d_lo * 3 + 1) : 0.f,
                    dz0 = pre_dirs ? pre_dirs[2] : ok_lo ? __ldg(dirs + (size_t)d_lo * 3 + 2) : 1.f;
        const float dx1 = pre_dirs ? pre_dirs[3] : ok_hi ? __ldg(dirs + (size_t)d_hi * 3 + 0) : 0.f,
                    dy1 = pre_dirs ? pre_dirs[4] : ok_hi ? __ldg(dirs + (size_t)d_hi * 3 + 1) : 0.f,
                    dz1 = pre_dirs ? pre_dirs[5] : ok_hi ? __ldg(dirs + (size_t)d_hi * 3 + 2) : 1.f;
        sh4_lane(dx0, dy0, dz0, t, sh_lo);
        sh4_lane(dx1, dy1, dz1, t, sh_hi);
        a_hin[2][0] = tf32(sh_lo[0]); a_hin[2][1] = tf32(sh_hi[0]); a_hin[2][2] = tf32(sh_lo[1]); a_hin[2][3] = tf32(sh_hi[1]);
        a_hin[3][0] = tf32(sh_lo[2]); a_hin[3][1] = tf32(sh_hi[2]); a_hin[3][2] = tf32(sh_lo[3]); a_hin[3][3] = tf32(sh_hi[3]);
        if (kKeep == 2) store_frag_panel<4>(reinterpret_cast<uint8_t *>(act) + PB_HIN, a_hin, 0, warp, g, t);
        if (kKeep == 1) {
            float *buf = act + O_HIN + warp * 16 * S_A32;
            *reinterpret_cast<float2 *>(buf + g * S_A32 + 16 + 2 * t) = make_float2(sh_lo[0], sh_lo[1]);
            *reinterpret_cast<float2 *>(buf + g * S_A32 + 24 + 2 * t) = make_float2(sh_lo[2], sh_lo[3]);
            *reinterpret_cast<float2 *>(buf + (g + 8) * S_A32 + 16 + 2 * t) = make_float2(sh_hi[0], sh_hi[1]);
            *reinterpret_cast<float2 *>(buf + (g + 8) * S_A32 + 24 + 2 * t) = make_float2(sh_hi[2], sh_hi[3]);
        }
    }
    // layer 2: 32 -> 64, ReLU
    uint32_t a_h1[8][4];
    {
        float acc[8][4];
        layer_forward<4, 8, S_W2>(a_hin, sw + O_W2, acc, g, t);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[nt][i] = fmaxf(acc[nt][i], 0.f);
            c_to_a(acc[nt], a_h1[nt]);
        }
        if (kKeep == 1) store_frag<8, S_A64>(act + O_H1 + warp * 16 * S_A64, acc, 0, g, t);
        if (kKeep == 2) store_frag_panel<8>(reinterpret_cast<uint8_t *>(act) + PB_H1, a_h1, 0, warp, g, t);
    }
    // layer 3: 64 -> 64, ReLU
    {
        float acc[8][4];
        layer_forward<8, 8, S_W3>(a_h1, sw + O_W3, acc, g, t);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[nt][i] = fmaxf(acc[nt][i], 0.f);
            c_to_a(acc[nt], a_h2[nt]);
        }
        if (kKeep == 1) store_frag<8, S_A64>(act + O_H2 + warp * 16 * S_A64, acc, 0, g, t);
        if (kKeep == 2) store_frag_panel<8>(reinterpret_cast<uint8_t *>(act) + PB_H2, a_h2, 0, warp, g, t);
    }
    // layer 4: 64 -> 3 (padded to 8), sigmoid
    layer_forward<8, 1, S_W4>(a_h2, sw + O_W4, out_rgb, g, t);
#pragma unroll
    for (int i = 0; i < 4; ++i) out_rgb[0][i] = 1.f / (1.f + expf(-out_rgb[0][i]));
    st.rgb[0][0] = out_rgb[0][0];
    st.rgb[0][1] = out_rgb[0][1];
    st.rgb[1][0] = out_rgb[0][2];
    st.rgb[1][1] = out_rgb[0][3];
}

template <int kKeep, bool kDensityOnly>
__device__ __forceinline__ void warp_forward(const uint32_t *__restrict__ sw, float *__restrict__ act, uint32_t warp,
                                             uint32_t row0, uint32_t n, const float *__restrict__ enc,
                                             const float *__restrict__ dirs, uint32_t g, uint32_t t, FwdState &st,
                                             uint32_t (&a_h2)[8][4], float (&out_rgb)[1][4],
                                             uint32_t rows_per_group = 0, bool ok_lo_in = true, bool ok_hi_in = true) {
    const uint32_t r_lo = row0 + g, r_hi = row0 + g + 8;
    uint32_t a_in[4][4];
    load_enc_fragments(enc, r_lo, r_hi, r_lo < n && ok_lo_in, r_hi < n && ok_hi_in, t, a_in);
    warp_forward_from<kKeep, kDensityOnly>(sw, act, warp, row0, n, a_in, dirs, g, t, st, a_h2, out_rgb, rows_per_group,
                                           ok_lo_in, ok_hi_in);
}

// ---------------------------------------------------------------- forward kernel
template <bool kDensityOnly>
__global__ void __launch_bounds__(kThreads) nerf_mlp_forward_kernel(uint32_t n, uint32_t rows_per_group,
                                                                    const float *__restrict__ enc,
                                                                    const float *__restrict__ dirs,
                                                                    const float *__restrict__ weights,
                                                                    const uint32_t *__restrict__ group_counts,
                                                                    float *__restrict__ out) {
    extern __shared__ __align__(16) uint32_t smem_u32[];
    uint32_t *sw = smem_u32;
    load_weights(sw, weights);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3u;
    const uint32_t n_tiles = (n + 15u) / 16u;
    for (uint32_t tile = blockIdx.x * kWarps + warp; tile < n_tiles; tile += gridDim.x * kWarps) {
        const uint32_t row0 = tile * 16u;
        bool live_lo = row0 + g < n, live_hi = row0 + g + 8 < n;
        if (rows_per_group) {  // padding rows of the grouped layout are neither read nor written
            const uint32_t r_lo = row0 + g, r_hi = row0 + g + 8;
            live_lo = live_lo && r_lo % rows_per_group < __ldg(group_counts + r_lo / rows_per_group);
            live_hi = live_hi && r_hi % rows_per_group < __ldg(group_counts + r_hi / rows_per_group);
            if (!__any_sync(0xffffffffu, live_lo || live_hi)) continue;
        }
        FwdState st;
        uint32_t a_h2[8][4];
        float rgb[1][4];
        warp_forward<false, kDensityOnly>(sw, nullptr, warp, row0, n, enc, dirs, g, t, st, a_h2, rgb, rows_per_group, live_lo, live_hi);
        if (kDensityOnly) {
            if (t == 0) {
                if (live_lo) out[row0 + g] = expf(st.x0[0]);
                if (live_hi) out[row0 + g + 8] = expf(st.x0[1]);
            }
        } else {
            // lane t == 0 of each row group assembles (density, r, g, b): b comes from lane t == 1
            const float b_lo = __shfl_sync(0xffffffffu, st.rgb[0][0], (lane & ~3u) + 1);
            const float b_hi = __shfl_sync(0xffffffffu, st.rgb[1][0], (lane & ~3u) + 1);
            if (t == 0) {
                if (live_lo)
                    reinterpret_cast<float4 *>(out)[row0 + g] = make_float4(expf(st.x0[0]), st.rgb[0][0], st.rgb[0][1], b_lo);
                if (live_hi)
                    reinterpret_cast<float4 *>(out)[row0 + g + 8] = make_float4(expf(st.x0[1]), st.rgb[1][0], st.rgb[1][1], b_hi);
            }
        }
    }
}

// ---------------------------------------------------------------- encoder fused in front of the MLP (forward)
// The hash-grid gather feeds the MLP's first A fragments directly: lane (g, t) of a warp owns rows g, g+8 of the
// warp's 16-sample tile and needs feature columns 8kt+2t, 8kt+2t+1 = BOTH features of level 4kt+t, kt = 0..3, so it
// gathers 2 points x 4 levels x 8 corners itself (branch-free predicated loads, 16 rows in flight at a time) and the
// [n, 32] encoding never goes through HBM (inference), or is written once for the backward pass (training,
// kWriteEnc) without being read back.  While one warp waits on its gathers (L1/L2-bound) others run their
// tensor-core layers (issue-bound): the two halves of the forward overlap inside one SM.
// Same arithmetic as hashgrid_a1_forward + nerf_mlp_forward => same bits.
template <typename TT, bool kDensityOnly, bool kWriteEnc>
__global__ void __launch_bounds__(kThreads, 3) nerf_fused_forward_kernel(const __grid_constant__ NgpNerfFusedDescriptor d,
                                                                      const float *__restrict__ pos,
                                                                      const TT *__restrict__ table,
                                                                      const float *__restrict__ dirs,
                                                                      const float *__restrict__ weights,
                                                                      const uint32_t *__restrict__ group_counts,
                                                                      float *__restrict__ out, float *__restrict__ enc_out) {
    extern __shared__ __align__(16) uint32_t smem_u32[];
    uint32_t *sw = smem_u32;
    hg::LevelMeta *s_meta = reinterpret_cast<hg::LevelMeta *>(smem_u32 + kWeightFloats);
    load_weights(sw, weights);
    if (threadIdx.x < 16) s_meta[threadIdx.x] = hg::a1_level(d.grid, threadIdx.x);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3u;
    const uint32_t n = d.grid.n_points, rows_per_group = d.grid.rows_per_group;
    const float bound = d.grid.bound;
    const uint32_t n_tiles = (n + 15u) / 16u;
    for (uint32_t tile = blockIdx.x * kWarps + warp; tile < n_tiles; tile += gridDim.x * kWarps) {
        const uint32_t row0 = tile * 16u, r_lo = row0 + g, r_hi = row0 + g + 8;
        bool live_lo = r_lo < n, live_hi = r_hi < n;
        if (rows_per_group) {  // padding rows of the grouped layout are neither read nor written
            live_lo = live_lo && r_lo % rows_per_group < __ldg(group_counts + r_lo / rows_per_group);
            live_hi = live_hi && r_hi % rows_per_group < __ldg(group_counts + r_hi / rows_per_group);
            if (!__any_sync(0xffffffffu, live_lo || live_hi)) continue;
        }
        float x_lo[3] = {0.f, 0.f, 0.f}, x_hi[3] = {0.f, 0.f, 0.f};
        if (live_lo) {
            x_lo[0] = __ldg(pos + (size_t)r_lo * 3 + 0); x_lo[1] = __ldg(pos + (size_t)r_lo * 3 + 1); x_lo[2] = __ldg(pos + (size_t)r_lo * 3 + 2);
        }
        if (live_hi) {
            x_hi[0] = __ldg(pos + (size_t)r_hi * 3 + 0); x_hi[1] = __ldg(pos + (size_t)r_hi * 3 + 1); x_hi[2] = __ldg(pos + (size_t)r_hi * 3 + 2);
        }
        float p_lo[3], p_hi[3];
        hg::unit_pos<3>(x_lo, bound, p_lo);
        hg::unit_pos<3>(x_hi, bound, p_hi);
        uint32_t a_in[4][4];
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            const hg::LevelMeta m = s_meta[4 * kt + t];
            float lo[2], hi[2];
            hg::encode_point_level_pred<TT>(table, m, p_lo, live_lo, lo);
            hg::encode_point_level_pred<TT>(table, m, p_hi, live_hi, hi);
            a_in[kt][0] = tf32(lo[0]);
            a_in[kt][1] = tf32(hi[0]);
            a_in[kt][2] = tf32(lo[1]);
            a_in[kt][3] = tf32(hi[1]);
            if (kWriteEnc) {
                if (live_lo) *reinterpret_cast<float2 *>(enc_out + (size_t)r_lo * 32 + 8 * kt + 2 * t) = make_float2(lo[0], lo[1]);
                if (live_hi) *reinterpret_cast<float2 *>(enc_out + (size_t)r_hi * 32 + 8 * kt + 2 * t) = make_float2(hi[0], hi[1]);
            }
        }
        FwdState st;
        uint32_t a_h2[8][4];
        float rgb[1][4];
        warp_forward_from<false, kDensityOnly>(sw, nullptr, warp, row0, n, a_in, dirs, g, t, st, a_h2, rgb, rows_per_group,
                                               live_lo, live_hi);
        if (kDensityOnly) {
            if (t == 0) {
                if (live_lo) out[r_lo] = expf(st.x0[0]);
                if (live_hi) out[r_hi] = expf(st.x0[1]);
            }
        } else {
            const float b_lo = __shfl_sync(0xffffffffu, st.rgb[0][0], (lane & ~3u) + 1);
            const float b_hi = __shfl_sync(0xffffffffu, st.rgb[1][0], (lane & ~3u) + 1);
            if (t == 0) {
                if (live_lo) reinterpret_cast<float4 *>(out)[r_lo] = make_float4(expf(st.x0[0]), st.rgb[0][0], st.rgb[0][1], b_lo);
                if (live_hi) reinterpret_cast<float4 *>(out)[r_hi] = make_float4(expf(st.x0[1]), st.rgb[1][0], st.rgb[1][1], b_hi);
            }
        }
    }
}

// ---------------------------------------------------------------- backward kernel
// dW tile accumulation over the CTA's 128 samples: acc[mt] += act[:, 16mt' .. ]^T * D[:, 8nt ..]
template <int MT, int S_ACT>
__device__ __forceinline__ void dw_accumulate(float (&acc)[MT][4], const float *__restrict__ act, int feat0,
                                              const float *__restrict__ D, int col0, uint32_t g, uint32_t t) {
#pragma unroll 4
    for (int ks = 0; ks < kBlockSamples / 8; ++ks) {
        const int s = 8 * ks + t;
        const uint32_t b0 = tf32(D[s * S_D + col0 + g]), b1 = tf32(D[(s + 4) * S_D + col0 + g]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            uint32_t a[4];
            const float *p = act + s * S_ACT + feat0 + 16 * mt + g;
            a[0] = tf32(p[0]);
            a[1] = tf32(p[8]);
            a[2] = tf32(p[4 * S_ACT]);
            a[3] = tf32(p[4 * S_ACT + 8]);
            mma_tf32(acc[mt], a, b0, b1);
        }
    }
}

// same with the activation operand (enc) read from global memory, rows base .. base+127
template <int MT>
__device__ __forceinline__ void dw_accumulate_global(float (&acc)[MT][4], const float *__restrict__ enc, uint32_t base,
                                                     uint32_t n, const float *__restrict__ D, int col0, uint32_t g, uint32_t t) {
#pragma unroll 4
    for (int ks = 0; ks < kBlockSamples / 8; ++ks) {
        const int s = 8 * ks + t;
        const uint32_t b0 = tf32(D[s * S_D + col0 + g]), b1 = tf32(D[(s + 4) * S_D + col0 + g]);
        const bool ok0 = base + s < n, ok1 = base + s + 4 < n;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            uint32_t a[4];
            const float *p = enc + (size_t)(base + s) * 32 + 16 * mt + g;
            a[0] = tf32(ok0 ? __ldg(p) : 0.f);
            a[1] = tf32(ok0 ? __ldg(p + 8) : 0.f);
            a[2] = tf32(ok1 ? __ldg(p + 4 * 32) : 0.f);
            a[3] = tf32(ok1 ? __ldg(p + 4 * 32 + 8) : 0.f);
            mma_tf32(acc[mt], a, b0, b1);
        }
    }
}

// flush one 16x8 dW tile: c0 = (row g, col 2t), c1 = (g, 2t+1), c2 = (g+8, 2t), c3 = (g+8, 2t+1)
__device__ __forceinline__ void flush_tile(float *__restrict__ dW, int n_cols, int row0, int col0, const float (&c)[4],
                                           uint32_t g, uint32_t t) {
    const int c0 = col0 + 2 * t;
    if (c0 < n_cols) {
        atomicAdd(dW + (row0 + g) * n_cols + c0, c[0]);
        atomicAdd(dW + (row0 + g + 8) * n_cols + c0, c[2]);
    }
    if (c0 + 1 < n_cols) {
        atomicAdd(dW + (row0 + g) * n_cols + c0 + 1, c[1]);
        atomicAdd(dW + (row0 + g + 8) * n_cols + c0 + 1, c[3]);
    }
}

__global__ void __launch_bounds__(kThreads, 1) nerf_mlp_backward_kernel(uint32_t n, const float *__restrict__ enc,
                                                                        const float *__restrict__ dirs,
                                                                        const float *__restrict__ weights,
                                                                        const float *__restrict__ d_drgbs,
                                                                        float *__restrict__ d_enc,
                                                                        float *__restrict__ d_weights) {
    extern __shared__ __align__(16) uint32_t smem_u32[];
    uint32_t *sw = smem_u32;
    float *act = reinterpret_cast<float *>(smem_u32 + kWeightFloats);
    load_weights(sw, weights);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3u;

    // dW tiles owned by this warp for the whole kernel
    float dW0[2][4] = {}, dW1[1][4] = {}, dW2[2][4] = {}, dW3[4][4] = {}, dW4[1][4] = {};
    float *D = act + O_D;
    float *D_mine = D + warp * 16 * S_D;

    const uint32_t n_blocks = (n + kBlockSamples - 1) / kBlockSamples;
    for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const uint32_t base = blk * kBlockSamples, row0 = base + warp * 16;
        const uint32_t r_lo = row0 + g, r_hi = row0 + g + 8;
        // ---- forward recompute (activations of the CTA's 128 samples -> shared memory)
        FwdState st;
        uint32_t a_tmp[8][4];
        float rgb[1][4];
        warp_forward<true, false>(sw, act, warp, row0, n, enc, dirs, g, t, st, a_tmp, rgb);
        const float4 dd_lo = r_lo < n ? __ldg(reinterpret_cast<const float4 *>(d_drgbs) + r_lo) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 dd_hi = r_hi < n ? __ldg(reinterpret_cast<const float4 *>(d_drgbs) + r_hi) : make_float4(0.f, 0.f, 0.f, 0.f);

        // ---- layer 4 delta: d_a3 = d_rgb * rgb * (1 - rgb), cols (2t, 2t+1) of the padded 8
        float d3[1][4];
        {
            const float dr_lo0 = t == 0 ? dd_lo.y : t == 1 ? dd_lo.w : 0.f, dr_lo1 = t == 0 ? dd_lo.z : 0.f;
            const float dr_hi0 = t == 0 ? dd_hi.y : t == 1 ? dd_hi.w : 0.f, dr_hi1 = t == 0 ? dd_hi.z : 0.f;
            d3[0][0] = dr_lo0 * st.rgb[0][0] * (1.f - st.rgb[0][0]);
            d3[0][1] = dr_lo1 * st.rgb[0][1] * (1.f - st.rgb[0][1]);
            d3[0][2] = dr_hi0 * st.rgb[1][0] * (1.f - st.rgb[1][0]);
            d3[0][3] = dr_hi1 * st.rgb[1][1] * (1.f - st.rgb[1][1]);
        }
        store_frag<1, S_D>(D_mine, d3, 0, g, t);
        __syncthreads();  // D(d_a3) and all activations of the block are visible
        if (warp < 4) dw_accumulate<1, S_A64>(dW4, act + O_H2, 16 * warp, D, 0, g, t);
        // d_h2 = d_a3 * W4^T, masked by ReLU
        uint32_t a_d[8][4];
        float dacc[8][4];
        {
            uint32_t a3[1][4];
            c_to_a(d3[0], a3[0]);
            layer_backward<1, 8, S_W4>(a3, sw + O_W4, dacc, g, t);
            const float *h = act + O_H2 + warp * 16 * S_A64;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float2 lo = *reinterpret_cast<const float2 *>(h + g * S_A64 + 8 * nt + 2 * t);
                const float2 hi = *reinterpret_cast<const float2 *>(h + (g + 8) * S_A64 + 8 * nt + 2 * t);
                dacc[nt][0] = lo.x > 0.f ? dacc[nt][0] : 0.f;
                dacc[nt][1] = lo.y > 0.f ? dacc[nt][1] : 0.f;
                dacc[nt][2] = hi.x > 0.f ? dacc[nt][2] : 0.f;
                dacc[nt][3] = hi.y > 0.f ? dacc[nt][3] : 0.f;
                c_to_a(dacc[nt], a_d[nt]);
            }
        }
        __syncthreads();  // everyone finished reading D(d_a3)
        store_frag<8, S_D>(D_mine, dacc, 0, g, t);  // D = d_a2
        __syncthreads();
        dw_accumulate<4, S_A64>(dW3, act + O_H1, 0, D, 8 * warp, g, t);
        // d_h1 = d_a2 * W3^T, masked
        {
            layer_backward<8, 8, S_W3>(a_d, sw + O_W3, dacc, g, t);
            const float *h = act + O_H1 + warp * 16 * S_A64;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float2 lo = *reinterpret_cast<const float2 *>(h + g * S_A64 + 8 * nt + 2 * t);
                const float2 hi = *reinterpret_cast<const float2 *>(h + (g + 8) * S_A64 + 8 * nt + 2 * t);
                dacc[nt][0] = lo.x > 0.f ? dacc[nt][0] : 0.f;
                dacc[nt][1] = lo.y > 0.f ? dacc[nt][1] : 0.f;
                dacc[nt][2] = hi.x > 0.f ? dacc[nt][2] : 0.f;
                dacc[nt][3] = hi.y > 0.f ? dacc[nt][3] : 0.f;
                c_to_a(dacc[nt], a_d[nt]);
            }
        }
        __syncthreads();
        store_frag<8, S_D>(D_mine, dacc, 0, g, t);  // D = d_a1
        __syncthreads();
        dw_accumulate<2, S_A32>(dW2, act + O_HIN, 0, D, 8 * warp, g, t);
        // d_x = (d_a1 * W2^T)[:, :16] + d_density * exp(clip(x0, -15, 15)) on column 0 (nerfs.py:231-234)
        float dx[2][4];
        uint32_t a_dx[2][4];
        {
            layer_backward<8, 2, S_W2>(a_d, sw + O_W2, dx, g, t);
            if (t == 0) {
                dx[0][0] += dd_lo.x * expf(fminf(fmaxf(st.x0[0], -15.f), 15.f));
                dx[0][2] += dd_hi.x * expf(fminf(fmaxf(st.x0[1], -15.f), 15.f));
            }
            c_to_a(dx[0], a_dx[0]);
            c_to_a(dx[1], a_dx[1]);
        }
        __syncthreads();
        store_frag<2, S_D>(D_mine, dx, 0, g, t);  // D = d_x (16 columns)
        __syncthreads();
        dw_accumulate<1, S_A64>(dW1, act + O_H0, 16 * (warp >> 1), D, 8 * (warp & 1), g, t);
        // d_h0 = d_x * W1^T, masked
        {
            layer_backward<2, 8, S_W1>(a_dx, sw + O_W1, dacc, g, t);
            const float *h = act + O_H0 + warp * 16 * S_A64;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float2 lo = *reinterpret_cast<const float2 *>(h + g * S_A64 + 8 * nt + 2 * t);
                const float2 hi = *reinterpret_cast<const float2 *>(h + (g + 8) * S_A64 + 8 * nt + 2 * t);
                dacc[nt][0] = lo.x > 0.f ? dacc[nt][0] : 0.f;
                dacc[nt][1] = lo.y > 0.f ? dacc[nt][1] : 0.f;
                dacc[nt][2] = hi.x > 0.f ? dacc[nt][2] : 0.f;
                dacc[nt][3] = hi.y > 0.f ? dacc[nt][3] : 0.f;
                c_to_a(dacc[nt], a_d[nt]);
            }
        }
        __syncthreads();
        store_frag<8, S_D>(D_mine, dacc, 0, g, t);  // D = d_a0
        __syncthreads();
        dw_accumulate_global<2>(dW0, enc, base, n, D, 8 * warp, g, t);
        // d_enc = d_a0 * W0^T -> global
        {
            float de[4][4];
            layer_backward<8, 4, S_W0>(a_d, sw + O_W0, de, g, t);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                if (r_lo < n) *reinterpret_cast<float2 *>(d_enc + (size_t)r_lo * 32 + 8 * nt + 2 * t) = make_float2(de[nt][0], de[nt][1]);
                if (r_hi < n) *reinterpret_cast<float2 *>(d_enc + (size_t)r_hi * 32 + 8 * nt + 2 * t) = make_float2(de[nt][2], de[nt][3]);
            }
        }
        __syncthreads();  // before the next block overwrites the activation buffers and D
    }

    // ---- flush this CTA's share of the weight gradient
    flush_tile(d_weights + G_W0, 64, 0, 8 * warp, dW0[0], g, t);
    flush_tile(d_weights + G_W0, 64, 16, 8 * warp, dW0[1], g, t);
    flush_tile(d_weights + G_W1, 16, 16 * (warp >> 1), 8 * (warp & 1), dW1[0], g, t);
    flush_tile(d_weights + G_W2, 64, 0, 8 * warp, dW2[0], g, t);
    flush_tile(d_weights + G_W2, 64, 16, 8 * warp, dW2[1], g, t);
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) flush_tile(d_weights + G_W3, 64, 16 * mt, 8 * warp, dW3[mt], g, t);
    if (warp < 4) flush_tile(d_weights + G_W4, 3, 16 * warp, 0, dW4[0], g, t);
}

// ---------------------------------------------------------------- backward kernel, weight gradients on tcgen05
// Same per-warp register chain as above (forward recompute and the delta of every layer on mma.sync, 16 samples
// per warp), but the weight gradients -- 60 % of the instructions of the kernel above, all of them shared-memory
// fragment loads feeding small MMAs -- move to the 5th-generation tensor core: every activation / delta of the
// CTA's 128 samples is written ONCE, as tf32 bits, into an MN-major shared-memory panel (the bits the next layer's
// A fragment uses anyway), and ONE thread issues dW (+)= act^T . delta as 16 tcgen05.mma (M = 64, K = 8 samples
// each) per layer.  The five accumulators (192 TMEM columns) stay in tensor memory for the CTA's whole lifetime and
// are flushed with one atomicAdd per element at the end.  The MMAs run asynchronously under the next layer's
// register chain; mbarriers (tcgen05.commit) only guard the reuse of a panel:
//   bar_w4: dW4 has read h2 / the narrow panel  -> d_a1 may overwrite h2, d_x the narrow panel
//   bar_w3: dW3 has read d_a2                   -> d_a0 may overwrite it
//   bar_w0: everything of this block is done    -> the next block may overwrite the panels
template <int NT>
__device__ __forceinline__ void relu_mask_to_a(const uint8_t *__restrict__ panel, uint32_t warp, uint32_t g, uint32_t t,
                                               float (&dacc)[NT][4], uint32_t (&a_d)[NT][4]) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const uint32_t off = panel_frag_offset(warp, g, t, 8 * nt);
        const float2 lo = *reinterpret_cast<const float2 *>(panel + off);
        const float2 hi = *reinterpret_cast<const float2 *>(panel + off + 1024u);
        dacc[nt][0] = lo.x > 0.f ? dacc[nt][0] : 0.f;
        dacc[nt][1] = lo.y > 0.f ? dacc[nt][1] : 0.f;
        dacc[nt][2] = hi.x > 0.f ? dacc[nt][2] : 0.f;
        dacc[nt][3] = hi.y > 0.f ? dacc[nt][3] : 0.f;
        c_to_a(dacc[nt], a_d[nt]);
    }
}

// D[64 x N] (+)= A^T . B over the 128 samples: A = 64-feature panel pair, B = N-feature panel(s), both MN-major
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool accumulate) {
#pragma unroll 4
    for (uint32_t ks = 0; ks < kBlockSamples / 8; ++ks)
        umma::mma_tf32(tmem_d, umma::desc_mn_major(a_addr, ks, kPanel), umma::desc_mn_major(b_addr, ks, kPanel), idesc,
                       accumulate || ks > 0);
}

// One (sample, level) of the hash-table scatter, from the MLP backward's own registers: the gradient pair (g0, g1) of
// sample position p01 at level m goes to the 2^3 corner rows with the trilinear weights (hashgrid_a1_backward_kernel's
// arithmetic; x-neighbour rows that are adjacent in memory share one 16-byte reduction).
__device__ __forceinline__ void scatter_sample_level(float *__restrict__ d_table, const hg::LevelMeta &m, const float (&p01)[3],
                                                     float g0, float g1) {
    if (g0 == 0.f && g1 == 0.f) return;  // padded / masked samples carry exact zeros
    uint32_t base[3];
    float fr[3];
    hg::a1_cell<3>(p01, m.scale, base, fr);
#pragma unroll
    for (int c = 0; c < 4; ++c) {  // corner c: x bit clear; its partner c + 4: x bit set
        uint32_t va[3], vb[3];
        float w = 1.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t bit = (c >> (2 - k)) & 1;
            va[k] = base[k] + bit;
            vb[k] = va[k];
            if (k > 0) w *= bit ? fr[k] : 1.f - fr[k];
        }
        vb[0] = base[0] + 1u;
        const float wa = w * (1.f - fr[0]), wb = w * fr[0];
        uint32_t ra = hg::grid_row_unclamped<3, true>(va, m), rb = hg::grid_row_unclamped<3, true>(vb, m);
        float a0 = wa * g0, a1 = wa * g1, b0 = wb * g0, b1 = wb * g1;
        if (ra > m.last_row) { ra = m.last_row; a0 = a1 = 0.f; }  // XLA drops the update of an out-of-range row (hg::grid_row)
        if (rb > m.last_row) { rb = m.last_row; b0 = b1 = 0.f; }
        if ((ra ^ rb) == 1u) {
            const bool a_hi = ra & 1u;
            red_add_v4(d_table + (size_t)(ra & ~1u) * 2, a_hi ? b0 : a0, a_hi ? b1 : a1, a_hi ? a0 : b0, a_hi ? a1 : b1);
        } else {
            red_add_v2(d_table + (size_t)ra * 2, a0, a1);
            red_add_v2(d_table + (size_t)rb * 2, b0, b1);
        }
    }
}

// Two CONSECUTIVE samples of a ray at one level (a thread of the eight-warp backward kernel holds rows 2g and 2g + 1 of its
// warp's tile): when both fall into the same cell -- nearly always at the coarse levels, where a cell spans tens of march
// steps -- their eight corner rows are the same rows, and one set of reductions carries both (the weights differ, the sums
// are formed in registers).  Otherwise two ordinary scatters.  Same sums as scatter_sample_level on each, to rounding order.
__device__ __forceinline__ void scatter_pair_level(float *__restrict__ d_table, const hg::LevelMeta &m, const float (&pa)[3],
                                                   const float (&pb)[3], float a0, float a1, float b0, float b1) {
    const bool has_a = a0 != 0.f || a1 != 0.f, has_b = b0 != 0.f || b1 != 0.f;  // padded / masked samples carry exact zeros
    if (!has_a && !has_b) return;
    uint32_t base[3], base_b[3];
    float fa[3], fb[3];
    hg::a1_cell<3>(pa, m.scale, base, fa);
    hg::a1_cell<3>(pb, m.scale, base_b, fb);
    const bool same = base[0] == base_b[0] && base[1] == base_b[1] && base[2] == base_b[2];
    if (!same || !has_a) {  // b on its own (also when a is empty: then b's cell is the one to scatter into)
        scatter_sample_level(d_table, m, pb, b0, b1);
        if (!has_a) return;
        b0 = b1 = 0.f;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {  // corner c: x bit clear; its partner c + 4: x bit set
        uint32_t va[3], vb[3];
        float wa = 1.f, wb = 1.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t bit = (c >> (2 - k)) & 1;
            va[k] = base[k] + bit;
            vb[k] = va[k];
            if (k > 0) {
                wa *= bit ? fa[k] : 1.f - fa[k];
                wb *= bit ? fb[k] : 1.f - fb[k];
            }
        }
        vb[0] = base[0] + 1u;
        // corner with the x bit clear (rows ra) and set (rows rb): contributions of sample a and sample b
        const float wa0 = wa * (1.f - fa[0]), wa1 = wa * fa[0], wb0 = wb * (1.f - fb[0]), wb1 = wb * fb[0];
        uint32_t ra = hg::grid_row_unclamped<3, true>(va, m), rb = hg::grid_row_unclamped<3, true>(vb, m);
        float x0 = wa0 * a0 + wb0 * b0, x1 = wa0 * a1 + wb0 * b1, y0 = wa1 * a0 + wb1 * b0, y1 = wa1 * a1 + wb1 * b1;
        if (ra > m.last_row) { ra = m.last_row; x0 = x1 = 0.f; }  // XLA drops the update of an out-of-range row (hg::grid_row)
        if (rb > m.last_row) { rb = m.last_row; y0 = y1 = 0.f; }
        if ((ra ^ rb) == 1u) {
            const bool a_hi = ra & 1u;
            red_add_v4(d_table + (size_t)(ra & ~1u) * 2, a_hi ? y0 : x0, a_hi ? y1 : x1, a_hi ? x0 : y0, a_hi ? x1 : y1);
        } else {
            red_add_v2(d_table + (size_t)ra * 2, x0, x1);
            red_add_v2(d_table + (size_t)rb * 2, y0, y1);
        }
    }
}

// Flush a CTA's weight-gradient accumulators (tensor memory, M = 64: row m sits in lane (m % 16) + 32 (m / 16)) into the flat
// gradient with reductions; called by warps 0..7 once the last MMA has completed.
__device__ __forceinline__ void flush_weight_gradients(uint32_t tmem, uint32_t warp, uint32_t lane, float *__restrict__ d_weights) {
    const uint32_t q = warp & 3u, m = 16u * q + lane;
    const uint32_t taddr = tmem + ((32u * q) << 16);
    const bool mine = lane < 16u;
    uint32_t v[32];
    // rows that are contiguous in the flat gradient go out as 16-byte reductions when the buffer allows it
    const bool vec = (reinterpret_cast<uintptr_t>(d_weights) & 15u) == 0;
    auto add_row = [&](float *dst, int count) {  // count: 32 or 16 consecutive floats from v[]
        if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (j < count) red_add_v4(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < count) atomicAdd(dst + j, __uint_as_float(v[j]));
        }
    };
    // warps 0-3: dW3[:, :32] and dW0; warps 4-7: dW3[:, 32:], dW2, dW1, dW4
    umma::tmem_ld32(taddr + T_W3 + (warp < 4 ? 0 : 32), v);
    umma::tmem_ld_wait();
    if (mine) add_row(d_weights + G_W3 + m * 64 + (warp < 4 ? 0 : 32), 32);
    if (warp < 4) {
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
        if (mine) add_row(d_weights + G_W1 + m * 16, 16);
        umma::tmem_ld32(taddr + T_W4, v);  // dW4[m][j < 3]
        umma::tmem_ld_wait();
        if (mine)
#pragma unroll
            for (int j = 0; j < 3; ++j) atomicAdd(d_weights + G_W4 + m * 3 + j, __uint_as_float(v[j]));
    }
}

// kScatter: the input gradient never leaves the SM -- each thread scatters its fragment of d_enc (rows g, g + 8 of the
// warp's 16 samples, levels t, t + 4, t + 8, t + 12) straight into the hash-table gradient (ngp_nerf_mlp_backward_scatter).
//
// kBwdIssuerWarp: a ninth warp issues every weight-gradient MMA.  The chain warps never wait for each other -- each reads
// and writes only its own 16 rows of the panels -- so what used to be five block-wide barriers per block (so that thread 0
// could issue once every row was written) become non-blocking `bar.arrive`s on five named barriers the issuing warp
// `bar.sync`s on in turn; warp 0 no longer carries the 80 MMA issues (and their descriptor arithmetic) of every block
// with seven warps waiting for it at the next barrier.  A chain warp cannot lap the issuer: its next block starts with a
// wait on bar_w0, which the issuer commits after the fifth barrier of this block.
#ifndef NGP_BWD_ISSUER_WARP
#define NGP_BWD_ISSUER_WARP 1
#endif
constexpr bool kBwdIssuerWarp = NGP_BWD_ISSUER_WARP != 0;
constexpr int kBwdThreads = kThreads + (kBwdIssuerWarp ? 32 : 0);

template <bool kScatter>
__global__ void __launch_bounds__(kBwdThreads, 1) nerf_mlp_backward_umma_kernel(uint32_t n, const float *__restrict__ enc,
                                                                             const float *__restrict__ dirs,
                                                                             const float *__restrict__ weights,
                                                                             const float *__restrict__ d_drgbs,
                                                                             float *__restrict__ d_enc,
                                                                             float *__restrict__ d_weights,
                                                                             const __grid_constant__ NgpHashGridA1Descriptor gd,
                                                                             const float *__restrict__ pos,
                                                                             float *__restrict__ d_table) {
    __shared__ hg::LevelMeta s_meta[kScatter ? 16 : 1];
    if (kScatter && threadIdx.x < 16) s_meta[threadIdx.x] = hg::a1_level(gd, threadIdx.x);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // aligned by an offset from the array (not through an integer round trip) so that the compiler keeps the
    // shared address space and emits LDS/STS rather than generic loads and stores
    uint8_t *panels = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint32_t *sw = reinterpret_cast<uint32_t *>(panels + kPanelBytes);
    __shared__ uint64_t bar_w4, bar_w3, bar_w0;
    __shared__ uint32_t tmem_slot;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, g = lane >> 2, t = lane & 3u;

    load_weights(sw, weights);
    for (uint32_t i = tid; i < kPanel / 16; i += kBwdThreads) reinterpret_cast<uint4 *>(panels + PB_S)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        umma::mbar_init(&bar_w4, 1);
        umma::mbar_init(&bar_w3, 1);
        umma::mbar_init(&bar_w0, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(&tmem_slot, kTmemCols);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t pa = umma::smem_u32(panels);
    constexpr uint32_t kI32 = umma::make_idesc(64, 32, true, true), kI64 = umma::make_idesc(64, 64, true, true);
    float *act = reinterpret_cast<float *>(panels);

    const uint32_t n_blocks = (n + kBlockSamples - 1) / kBlockSamples;
    uint32_t it = 0;
    // the five weight-gradient products of a block, in the order their operands become complete
    auto issue = [&](int k, bool acc) {
        if (k == 0) {
            issue_wgrad(tmem + T_W4, pa + PB_H2, pa + PB_S, kI32, acc);  // dW4 = h2^T . d_a3
            umma::commit(&bar_w4);
        } else if (k == 1) {
            issue_wgrad(tmem + T_W3, pa + PB_H1, pa + PB_DA, kI64, acc);  // dW3 = h1^T . d_a2
            umma::commit(&bar_w3);
        } else if (k == 2) {
            issue_wgrad(tmem + T_W2, pa + PB_H2, pa + PB_HIN, kI32, acc);  // dW2^T = d_a1^T . hin
        } else if (k == 3) {
            issue_wgrad(tmem + T_W1, pa + PB_H0, pa + PB_S, kI32, acc);  // dW1 = h0^T . d_x
        } else {
            issue_wgrad(tmem + T_W0, pa + PB_DA, pa + PB_ENC, kI32, acc);  // dW0^T = d_a0^T . enc
            umma::commit(&bar_w0);
        }
    };
    // a chain thread's panel rows of step k are written: make them visible to the tensor core and tell the issuer
    auto publish = [&](int k, bool acc) {
        umma::fence_smem_to_async();
        if (kBwdIssuerWarp) {
            asm volatile("bar.arrive %0, %1;" ::"r"(1 + k), "n"(kBwdThreads) : "memory");
        } else {
            __syncthreads();
            if (tid == 0) {
                umma::fence_after_sync();
                issue(k, acc);
            }
        }
    };
    if (kBwdIssuerWarp && warp == kWarps) {
        // the whole warp walks the barriers (bar.sync is warp-wide); lane 0 alone issues
        for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++it) {
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                asm volatile("bar.sync %0, %1;" ::"r"(1 + k), "n"(kBwdThreads) : "memory");
                umma::fence_after_sync();
                if (lane == 0) issue(k, it > 0);
                __syncwarp();
            }
        }
    } else {
    // A block's global inputs (enc fragments, output cotangents, directions) are fetched a WHOLE block ahead, into a
    // second register set: with two warps per scheduler nothing else hides a trip to L2 or HBM.  (Measured alone this
    // changed nothing -- the long-scoreboard samples turned out to be the weight-staging prologue -- but it costs nothing
    // either: one block per SM leaves the registers free.)
    struct Inputs {
        uint32_t a_in[4][4];
        float4 dd_lo, dd_hi;
        float dir[6];
    };
    Inputs cur, nxt;
    auto fetch = [&](uint32_t blk, Inputs &x) {
        // fragment rows g and g + 8 of the warp's tile hold samples 2g and 2g + 1: any assignment of the tile's 16 samples to
        // its rows gives the same sums, and this one puts two CONSECUTIVE samples of a ray into each thread for the scatter
        const uint32_t r_lo = blk * kBlockSamples + warp * 16 + 2 * g, r_hi = r_lo + 1;
        load_enc_fragments(enc, r_lo, r_hi, r_lo < n, r_hi < n, t, x.a_in);
        x.dd_lo = r_lo < n ? __ldg(reinterpret_cast<const float4 *>(d_drgbs) + r_lo) : make_float4(0.f, 0.f, 0.f, 0.f);
        x.dd_hi = r_hi < n ? __ldg(reinterpret_cast<const float4 *>(d_drgbs) + r_hi) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            x.dir[k] = r_lo < n ? __ldg(dirs + (size_t)r_lo * 3 + k) : (k == 2 ? 1.f : 0.f);
            x.dir[3 + k] = r_hi < n ? __ldg(dirs + (size_t)r_hi * 3 + k) : (k == 2 ? 1.f : 0.f);
        }
    };
    if (blockIdx.x < n_blocks) fetch(blockIdx.x, cur);
    else fetch(n_blocks, cur);  // out of range: zeros, so that no register is read uninitialised
    for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++it) {
        if (blk + gridDim.x < n_blocks) fetch(blk + gridDim.x, nxt);
        uint32_t (&a_in)[4][4] = cur.a_in;
        const float4 dd_lo = cur.dd_lo, dd_hi = cur.dd_hi;
        const uint32_t par = it & 1u;
        const bool acc = it > 0;
        const uint32_t base = blk * kBlockSamples, row0 = base + warp * 16;
        const uint32_t r_lo = row0 + 2 * g, r_hi = r_lo + 1;  // see fetch
        if (it > 0) umma::mbar_wait(&bar_w0, par ^ 1u);  // the previous block's wgrad MMAs have read every panel
        store_frag_panel<4>(panels + PB_ENC, a_in, 0, warp, g, t);
        // ---- forward recompute; h0, hin, h1, h2 of the CTA's 128 samples -> panels
        FwdState st;
        uint32_t a_tmp[8][4];
        float rgb[1][4];
        warp_forward_from<2, false>(sw, act, warp, row0, n, a_in, dirs, g, t, st, a_tmp, rgb, 0, true, true, cur.dir);

        // ---- layer 4 delta: d_a3 = d_rgb * rgb * (1 - rgb), cols (2t, 2t+1) of the padded 8
        float d3[1][4];
        uint32_t a3[2][4];
        {
            const float dr_lo0 = t == 0 ? dd_lo.y : t == 1 ? dd_lo.w : 0.f, dr_lo1 = t == 0 ? dd_lo.z : 0.f;
            const float dr_hi0 = t == 0 ? dd_hi.y : t == 1 ? dd_hi.w : 0.f, dr_hi1 = t == 0 ? dd_hi.z : 0.f;
            d3[0][0] = dr_lo0 * st.rgb[0][0] * (1.f - st.rgb[0][0]);
            d3[0][1] = dr_lo1 * st.rgb[0][1] * (1.f - st.rgb[0][1]);
            d3[0][2] = dr_hi0 * st.rgb[1][0] * (1.f - st.rgb[1][0]);
            d3[0][3] = dr_hi1 * st.rgb[1][1] * (1.f - st.rgb[1][1]);
            c_to_a(d3[0], a3[0]);
            a3[1][0] = a3[1][1] = a3[1][2] = a3[1][3] = 0u;  // columns 8..15 held d_x of the previous block
        }
        store_frag_panel<2>(panels + PB_S, a3, 0, warp, g, t);
        publish(0, acc);  // enc, h0, hin, h1, h2, d_a3 -> dW4
        // d_a2 = (d_a3 . W4^T) masked by h2 > 0
        uint32_t a_d[8][4];
        float dacc[8][4];
        {
            uint32_t a3k[1][4];
            a3k[0][0] = a3[0][0]; a3k[0][1] = a3[0][1]; a3k[0][2] = a3[0][2]; a3k[0][3] = a3[0][3];
            layer_backward<1, 8, S_W4>(a3k, sw + O_W4, dacc, g, t);
            relu_mask_to_a<8>(panels + PB_H2, warp, g, t, dacc, a_d);
        }
        store_frag_panel<8>(panels + PB_DA, a_d, 0, warp, g, t);
        publish(1, acc);  // d_a2 -> dW3
        // d_a1 = (d_a2 . W3^T) masked by h1 > 0 -> takes over the h2 panels
        layer_backward<8, 8, S_W3>(a_d, sw + O_W3, dacc, g, t);
        relu_mask_to_a<8>(panels + PB_H1, warp, g, t, dacc, a_d);
        umma::mbar_wait(&bar_w4, par);
        store_frag_panel<8>(panels + PB_H2, a_d, 0, warp, g, t);
        publish(2, acc);  // d_a1 -> dW2
        // d_x = (d_a1 . W2^T)[:, :16] + d_density * exp(clip(x0, -15, 15)) on column 0 (nerfs.py:231-234)
        float dx[2][4];
        uint32_t a_dx[2][4];
        layer_backward<8, 2, S_W2>(a_d, sw + O_W2, dx, g, t);
        if (t == 0) {
            dx[0][0] += dd_lo.x * expf(fminf(fmaxf(st.x0[0], -15.f), 15.f));
            dx[0][2] += dd_hi.x * expf(fminf(fmaxf(st.x0[1], -15.f), 15.f));
        }
        c_to_a(dx[0], a_dx[0]);
        c_to_a(dx[1], a_dx[1]);
        store_frag_panel<2>(panels + PB_S, a_dx, 0, warp, g, t);
        publish(3, acc);  // d_x -> dW1
        // d_a0 = (d_x . W1^T) masked by h0 > 0 -> takes over the d_a2 panels
        layer_backward<2, 8, S_W1>(a_dx, sw + O_W1, dacc, g, t);
        relu_mask_to_a<8>(panels + PB_H0, warp, g, t, dacc, a_d);
        umma::mbar_wait(&bar_w3, par);
        store_frag_panel<8>(panels + PB_DA, a_d, 0, warp, g, t);
        publish(4, acc);  // d_a0 -> dW0
        // d_enc = d_a0 . W0^T -> global, or straight into the table gradient
        {
            float de[4][4];
            layer_backward<8, 4, S_W0>(a_d, sw + O_W0, de, g, t);
            if (!kScatter) {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    if (r_lo < n) *reinterpret_cast<float2 *>(d_enc + (size_t)r_lo * 32 + 8 * nt + 2 * t) = make_float2(de[nt][0], de[nt][1]);
                    if (r_hi < n) *reinterpret_cast<float2 *>(d_enc + (size_t)r_hi * 32 + 8 * nt + 2 * t) = make_float2(de[nt][2], de[nt][3]);
                }
            } else {
                float xl[3] = {0.f, 0.f, 0.f}, xh[3] = {0.f, 0.f, 0.f}, pl[3], ph[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (r_lo < n) xl[k] = __ldg(pos + (size_t)r_lo * 3 + k);
                    if (r_hi < n) xh[k] = __ldg(pos + (size_t)r_hi * 3 + k);
                }
                hg::unit_pos<3>(xl, gd.bound, pl);
                hg::unit_pos<3>(xh, gd.bound, ph);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const hg::LevelMeta m = s_meta[4 * nt + t];  // column 8 nt + 2 t is feature 0 of level 4 nt + t
                    // rows beyond n carry exact zeros (their cotangents were loaded as zeros): nothing is scattered for them
                    scatter_pair_level(d_table, m, pl, ph, de[nt][0], de[nt][1], de[nt][2], de[nt][3]);
                }
            }
        }
        cur = nxt;
    }
    }  // chain warps

    // ---- flush the CTA's weight gradients
    if (it > 0 && warp < kWarps) {
        umma::mbar_wait(&bar_w0, (it - 1u) & 1u);
        umma::fence_after_sync();
        flush_weight_gradients(tmem, warp, lane, d_weights);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------- backward, sixteen chain warps (column-split pairs)
// The kernel above keeps ONE 225 KB block per SM, i.e. two chain warps per scheduler, and its profile is all latency
// (issue slots 20 % used, tensor pipe 25 %).  The panels cannot shrink, so this kernel doubles the warps over the SAME
// panels instead: warps p and p + 8 share row tile p (16 samples) and each computes HALF the output columns of every
// layer.  A layer's output goes into the MN-major panels anyway (the weight gradients read it there); the partner's half
// comes back from the panel as A fragments after a 64-thread named barrier (eight pair barriers per block), the own
// half never leaves the registers.  Per warp and block: 148 MMAs instead of 288 (layer 4 and the SH basis are done by
// both warps), 26 + 26 extra 8-byte shared loads, and half the accumulator registers.  Barrier ids: 0 block, 1-5 issuer,
// 6-13 pairs.
// MEASURED (C2 batch, L2 flushed): correct on its first run, and SLOWER than the eight-warp kernel -- 0.153 ms against 0.132 ms
// alone, 0.286 against 0.265 ms with the table scatter fused in, training step 0.504 against 0.498 ms: eight exchanges per
// block (store, 64-thread barrier, load) sit on every warp's critical path and cost more than four warps per scheduler
// hide.  It stays as the opt-in arm (NGP_B200_MLP_BWD_SPLIT=1) with its parity tests.
constexpr int kSplitChainWarps = 16;
constexpr int kSplitThreads = (kSplitChainWarps + 1) * 32;  // 544: sixteen chain warps + the issuing warp

template <int NT>
__device__ __forceinline__ void zero_acc(float (&acc)[NT][4]) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
}
// acc[16 x 8 NT] += a[16 x 8 KT] . W[8 KT rows][8 NT cols]; W points at the first of those rows and columns
template <int KT, int NT, int STRIDE>
__device__ __forceinline__ void fwd_acc(float (&acc)[NT][4], const uint32_t (&a)[KT][4], const uint32_t *__restrict__ W, uint32_t g, uint32_t t) {
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
        const uint32_t *w0 = W + (8 * kt + 2 * t) * STRIDE + g;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt], a[kt], w0[8 * nt], w0[STRIDE + 8 * nt]);
    }
}
// acc[16 x 8 NT] += a[16 x 8 KT] . W^T; W points at row (first output column) and column (first k) of the stored [in][out] matrix
template <int KT, int NT, int STRIDE>
__device__ __forceinline__ void bwd_acc(float (&acc)[NT][4], const uint32_t (&a)[KT][4], const uint32_t *__restrict__ W, uint32_t g, uint32_t t) {
    const uint32_t *w0 = W + g * STRIDE + 2 * t;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const uint2 b = *reinterpret_cast<const uint2 *>(w0 + 8 * nt * STRIDE + 8 * kt);
            mma_tf32(acc[nt], a[kt], b.x, b.y);
        }
    }
}
// inverse of store_frag_panel: A fragments of columns col0 .. col0 + 8 NT of row tile `tile`
template <int NT>
__device__ __forceinline__ void load_frag_panel(const uint8_t *__restrict__ panel, uint32_t (&a)[NT][4], int col0, uint32_t tile, uint32_t g, uint32_t t) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const uint32_t off = panel_frag_offset(tile, g, t, col0 + 8 * nt);
        const uint2 lo = *reinterpret_cast<const uint2 *>(panel + off);
        const uint2 hi = *reinterpret_cast<const uint2 *>(panel + off + 1024u);
        a[nt][0] = lo.x; a[nt][2] = lo.y; a[nt][1] = hi.x; a[nt][3] = hi.y;
    }
}
// ReLU mask from the activation's panel columns col0.. (own rows), then accumulator -> A fragments
template <int NT>
__device__ __forceinline__ void mask_to_a(const uint8_t *__restrict__ panel, int col0, uint32_t tile, uint32_t g, uint32_t t,
                                          float (&dacc)[NT][4], uint32_t (&a_d)[NT][4]) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const uint32_t off = panel_frag_offset(tile, g, t, col0 + 8 * nt);
        const float2 lo = *reinterpret_cast<const float2 *>(panel + off);
        const float2 hi = *reinterpret_cast<const float2 *>(panel + off + 1024u);
        dacc[nt][0] = lo.x > 0.f ? dacc[nt][0] : 0.f;
        dacc[nt][1] = lo.y > 0.f ? dacc[nt][1] : 0.f;
        dacc[nt][2] = hi.x > 0.f ? dacc[nt][2] : 0.f;
        dacc[nt][3] = hi.y > 0.f ? dacc[nt][3] : 0.f;
        c_to_a(dacc[nt], a_d[nt]);
    }
}
template <int NT>
__device__ __forceinline__ void relu_to_a(float (&acc)[NT][4], uint32_t (&a)[NT][4]) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = fmaxf(acc[nt][i], 0.f);
        c_to_a(acc[nt], a[nt]);
    }
}

template <bool kScatter>
__global__ void __launch_bounds__(kSplitThreads, 1) nerf_mlp_backward_split_kernel(uint32_t n, const float *__restrict__ enc,
                                                                                    const float *__restrict__ dirs,
                                                                                    const float *__restrict__ weights,
                                                                                    const float *__restrict__ d_drgbs,
                                                                                    float *__restrict__ d_enc,
                                                                                    float *__restrict__ d_weights,
                                                                                    const __grid_constant__ NgpHashGridA1Descriptor gd,
                                                                                    const float *__restrict__ pos,
                                                                                    float *__restrict__ d_table) {
    __shared__ hg::LevelMeta s_meta[kScatter ? 16 : 1];
    if (kScatter && threadIdx.x < 16) s_meta[threadIdx.x] = hg::a1_level(gd, threadIdx.x);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *panels = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint32_t *sw = reinterpret_cast<uint32_t *>(panels + kPanelBytes);
    __shared__ uint64_t bar_w4, bar_w3, bar_w0;
    __shared__ uint32_t tmem_slot;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, g = lane >> 2, t = lane & 3u;

    load_weights(sw, weights);  // threads 0..255
    for (uint32_t i = tid; i < kPanel / 16; i += kSplitThreads) reinterpret_cast<uint4 *>(panels + PB_S)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        umma::mbar_init(&bar_w4, 1);
        umma::mbar_init(&bar_w3, 1);
        umma::mbar_init(&bar_w0, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(&tmem_slot, kTmemCols);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t pa = umma::smem_u32(panels);
    constexpr uint32_t kI32 = umma::make_idesc(64, 32, true, true), kI64 = umma::make_idesc(64, 64, true, true);

    const uint32_t n_blocks = (n + kBlockSamples - 1) / kBlockSamples;
    uint32_t it = 0;
    if (warp == kSplitChainWarps) {
        // ---- issuing warp: the five weight-gradient products of a block, each once every chain warp has published its rows
        for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++it) {
            const bool acc = it > 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                asm volatile("bar.sync %0, %1;" ::"r"(1 + k), "n"(kSplitThreads) : "memory");
                umma::fence_after_sync();
                if (lane == 0) {
                    if (k == 0) {
                        issue_wgrad(tmem + T_W4, pa + PB_H2, pa + PB_S, kI32, acc);  // dW4 = h2^T . d_a3
                        umma::commit(&bar_w4);
                    } else if (k == 1) {
                        issue_wgrad(tmem + T_W3, pa + PB_H1, pa + PB_DA, kI64, acc);  // dW3 = h1^T . d_a2
                        umma::commit(&bar_w3);
                    } else if (k == 2) {
                        issue_wgrad(tmem + T_W2, pa + PB_H2, pa + PB_HIN, kI32, acc);  // dW2^T = d_a1^T . hin
                    } else if (k == 3) {
                        issue_wgrad(tmem + T_W1, pa + PB_H0, pa + PB_S, kI32, acc);  // dW1 = h0^T . d_x
                    } else {
                        issue_wgrad(tmem + T_W0, pa + PB_DA, pa + PB_ENC, kI32, acc);  // dW0^T = d_a0^T . enc
                        umma::commit(&bar_w0);
                    }
                }
                __syncwarp();
            }
        }
    } else {
        // ---- chain warps: row tile p, column half h
        const uint32_t p = warp & 7u, h = warp >> 3, o = 1u - h;
        auto publish = [&](int k) {
            umma::fence_smem_to_async();
            asm volatile("bar.arrive %0, %1;" ::"r"(1 + k), "n"(kSplitThreads) : "memory");
        };
        auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(6u + p) : "memory"); };
        const uint32_t *W0 = sw + O_W0, *W1 = sw + O_W1, *W2 = sw + O_W2, *W3 = sw + O_W3, *W4 = sw + O_W4;

        uint32_t a_in[4][4];
        float4 dd_lo = make_float4(0.f, 0.f, 0.f, 0.f), dd_hi = dd_lo;
        float dir[6] = {0.f, 0.f, 1.f, 0.f, 0.f, 1.f};
        auto fetch = [&](uint32_t blk) {
            const uint32_t r_lo = blk * kBlockSamples + p * 16 + g, r_hi = r_lo + 8;
            load_enc_fragments(enc, r_lo, r_hi, r_lo < n, r_hi < n, t, a_in);
            dd_lo = r_lo < n ? __ldg(reinterpret_cast<const float4 *>(d_drgbs) + r_lo) : make_float4(0.f, 0.f, 0.f, 0.f);
            dd_hi = r_hi < n ? __ldg(reinterpret_cast<const float4 *>(d_drgbs) + r_hi) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dir[k] = r_lo < n ? __ldg(dirs + (size_t)r_lo * 3 + k) : (k == 2 ? 1.f : 0.f);
                dir[3 + k] = r_hi < n ? __ldg(dirs + (size_t)r_hi * 3 + k) : (k == 2 ? 1.f : 0.f);
            }
        };
        if (blockIdx.x < n_blocks) fetch(blockIdx.x);
        for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++it) {
            const uint32_t par = it & 1u;
            const uint32_t r_lo = blk * kBlockSamples + p * 16 + g, r_hi = r_lo + 8;
            if (it > 0) umma::mbar_wait(&bar_w0, par ^ 1u);  // the previous block's weight gradients have read every panel
            // enc columns 16 h .. 16 h + 15 of the tile -> panel (the partner stores the other half)
            if (h == 0) store_frag_panel<2>(panels + PB_ENC, reinterpret_cast<const uint32_t(&)[2][4]>(a_in[0]), 0, p, g, t);
            else store_frag_panel<2>(panels + PB_ENC, reinterpret_cast<const uint32_t(&)[2][4]>(a_in[2]), 16, p, g, t);

            uint32_t a_own[4][4], a_oth[4][4];
            float acc[4][4];
            // layer 0: 32 -> 64, ReLU; own columns 32 h ..
            zero_acc(acc);
            fwd_acc<4, 4, S_W0>(acc, a_in, W0 + 32 * h, g, t);
            relu_to_a(acc, a_own);
            store_frag_panel<4>(panels + PB_H0, a_own, 32 * h, p, g, t);
            pair_sync();
            load_frag_panel<4>(panels + PB_H0, a_oth, 32 * o, p, g, t);
            // layer 1: 64 -> 16; own columns 8 h ..; x0 (column 0) lives with h == 0
            uint32_t a_x[1][4], a_xo[1][4], a_sh[2][4];
            float x0_lo, x0_hi;
            {
                float x[1][4];
                zero_acc(x);
                fwd_acc<4, 1, S_W1>(x, a_own, W1 + 32 * h * S_W1 + 8 * h, g, t);
                fwd_acc<4, 1, S_W1>(x, a_oth, W1 + 32 * o * S_W1 + 8 * h, g, t);
                x0_lo = x[0][0];
                x0_hi = x[0][2];
                c_to_a(x[0], a_x[0]);
                float sh_lo[4], sh_hi[4];
                sh4_lane(dir[0], dir[1], dir[2], t, sh_lo);
                sh4_lane(dir[3], dir[4], dir[5], t, sh_hi);
                a_sh[0][0] = tf32(sh_lo[0]); a_sh[0][1] = tf32(sh_hi[0]); a_sh[0][2] = tf32(sh_lo[1]); a_sh[0][3] = tf32(sh_hi[1]);
                a_sh[1][0] = tf32(sh_lo[2]); a_sh[1][1] = tf32(sh_hi[2]); a_sh[1][2] = tf32(sh_lo[3]); a_sh[1][3] = tf32(sh_hi[3]);
            }
            store_frag_panel<1>(panels + PB_HIN, a_x, 8 * h, p, g, t);
            if (h == 0) store_frag_panel<1>(panels + PB_HIN, reinterpret_cast<const uint32_t(&)[1][4]>(a_sh[0]), 16, p, g, t);
            else store_frag_panel<1>(panels + PB_HIN, reinterpret_cast<const uint32_t(&)[1][4]>(a_sh[1]), 24, p, g, t);
            pair_sync();
            load_frag_panel<1>(panels + PB_HIN, a_xo, 8 * o, p, g, t);
            // layer 2: 32 -> 64, ReLU
            zero_acc(acc);
            fwd_acc<1, 4, S_W2>(acc, a_x, W2 + 8 * h * S_W2 + 32 * h, g, t);
            fwd_acc<1, 4, S_W2>(acc, a_xo, W2 + 8 * o * S_W2 + 32 * h, g, t);
            fwd_acc<2, 4, S_W2>(acc, a_sh, W2 + 16 * S_W2 + 32 * h, g, t);
            relu_to_a(acc, a_own);
            store_frag_panel<4>(panels + PB_H1, a_own, 32 * h, p, g, t);
            pair_sync();
            load_frag_panel<4>(panels + PB_H1, a_oth, 32 * o, p, g, t);
            // layer 3: 64 -> 64, ReLU
            zero_acc(acc);
            fwd_acc<4, 4, S_W3>(acc, a_own, W3 + 32 * h * S_W3 + 32 * h, g, t);
            fwd_acc<4, 4, S_W3>(acc, a_oth, W3 + 32 * o * S_W3 + 32 * h, g, t);
            relu_to_a(acc, a_own);
            store_frag_panel<4>(panels + PB_H2, a_own, 32 * h, p, g, t);
            pair_sync();
            load_frag_panel<4>(panels + PB_H2, a_oth, 32 * o, p, g, t);
            // layer 4: 64 -> 3 (padded to 8), sigmoid -- both warps of the pair; then d_a3 = d_rgb * rgb * (1 - rgb)
            uint32_t a3[1][4];
            {
                float rgb[1][4];
                zero_acc(rgb);
                fwd_acc<4, 1, S_W4>(rgb, a_own, W4 + 32 * h * S_W4, g, t);
                fwd_acc<4, 1, S_W4>(rgb, a_oth, W4 + 32 * o * S_W4, g, t);
#pragma unroll
                for (int i = 0; i < 4; ++i) rgb[0][i] = 1.f / (1.f + expf(-rgb[0][i]));
                const float dr_lo0 = t == 0 ? dd_lo.y : t == 1 ? dd_lo.w : 0.f, dr_lo1 = t == 0 ? dd_lo.z : 0.f;
                const float dr_hi0 = t == 0 ? dd_hi.y : t == 1 ? dd_hi.w : 0.f, dr_hi1 = t == 0 ? dd_hi.z : 0.f;
                float d3[4];
                d3[0] = dr_lo0 * rgb[0][0] * (1.f - rgb[0][0]);
                d3[1] = dr_lo1 * rgb[0][1] * (1.f - rgb[0][1]);
                d3[2] = dr_hi0 * rgb[0][2] * (1.f - rgb[0][2]);
                d3[3] = dr_hi1 * rgb[0][3] * (1.f - rgb[0][3]);
                c_to_a(d3, a3[0]);
            }
            {   // PB_S: columns 0..7 = d_a3 (h == 0), columns 8..15 = zeros (they held d_x of the previous block; h == 1)
                uint32_t z[1][4] = {{0u, 0u, 0u, 0u}};
                if (h == 0) store_frag_panel<1>(panels + PB_S, a3, 0, p, g, t);
                else store_frag_panel<1>(panels + PB_S, z, 8, p, g, t);
            }
            publish(0);  // enc, h0, hin, h1, h2, d_a3 -> dW4
            // d_a2 = (d_a3 . W4^T) masked by h2 > 0; own columns 32 h ..
            zero_acc(acc);
            bwd_acc<1, 4, S_W4>(acc, a3, W4 + 32 * h * S_W4, g, t);
            mask_to_a<4>(panels + PB_H2, 32 * h, p, g, t, acc, a_own);
            store_frag_panel<4>(panels + PB_DA, a_own, 32 * h, p, g, t);
            publish(1);  // d_a2 -> dW3
            pair_sync();
            load_frag_panel<4>(panels + PB_DA, a_oth, 32 * o, p, g, t);
            // d_a1 = (d_a2 . W3^T) masked by h1 > 0 -> takes over the h2 panels
            zero_acc(acc);
            bwd_acc<4, 4, S_W3>(acc, a_own, W3 + 32 * h * S_W3 + 32 * h, g, t);
            bwd_acc<4, 4, S_W3>(acc, a_oth, W3 + 32 * h * S_W3 + 32 * o, g, t);
            mask_to_a<4>(panels + PB_H1, 32 * h, p, g, t, acc, a_own);
            umma::mbar_wait(&bar_w4, par);
            store_frag_panel<4>(panels + PB_H2, a_own, 32 * h, p, g, t);
            publish(2);  // d_a1 -> dW2
            pair_sync();
            load_frag_panel<4>(panels + PB_H2, a_oth, 32 * o, p, g, t);
            // d_x = (d_a1 . W2^T)[:, :16] + d_density * exp(clip(x0, -15, 15)) on column 0 (nerfs.py:231-234); own columns 8 h ..
            {
                float dx[1][4];
                zero_acc(dx);
                bwd_acc<4, 1, S_W2>(dx, a_own, W2 + 8 * h * S_W2 + 32 * h, g, t);
                bwd_acc<4, 1, S_W2>(dx, a_oth, W2 + 8 * h * S_W2 + 32 * o, g, t);
                if (h == 0 && t == 0) {
                    dx[0][0] += dd_lo.x * expf(fminf(fmaxf(x0_lo, -15.f), 15.f));
                    dx[0][2] += dd_hi.x * expf(fminf(fmaxf(x0_hi, -15.f), 15.f));
                }
                c_to_a(dx[0], a_x[0]);
            }
            store_frag_panel<1>(panels + PB_S, a_x, 8 * h, p, g, t);
            publish(3);  // d_x -> dW1
            pair_sync();
            load_frag_panel<1>(panels + PB_S, a_xo, 8 * o, p, g, t);
            // d_a0 = (d_x . W1^T) masked by h0 > 0 -> takes over the d_a2 panels
            zero_acc(acc);
            bwd_acc<1, 4, S_W1>(acc, a_x, W1 + 32 * h * S_W1 + 8 * h, g, t);
            bwd_acc<1, 4, S_W1>(acc, a_xo, W1 + 32 * h * S_W1 + 8 * o, g, t);
            mask_to_a<4>(panels + PB_H0, 32 * h, p, g, t, acc, a_own);
            umma::mbar_wait(&bar_w3, par);
            store_frag_panel<4>(panels + PB_DA, a_own, 32 * h, p, g, t);
            publish(4);  // d_a0 -> dW0
            pair_sync();
            load_frag_panel<4>(panels + PB_DA, a_oth, 32 * o, p, g, t);
            if (blk + gridDim.x < n_blocks) fetch(blk + gridDim.x);
            // d_enc = d_a0 . W0^T; own columns 16 h .. 16 h + 15 -> global, or straight into the table gradient
            {
                float de[2][4];
                zero_acc(de);
                bwd_acc<4, 2, S_W0>(de, a_own, W0 + 16 * h * S_W0 + 32 * h, g, t);
                bwd_acc<4, 2, S_W0>(de, a_oth, W0 + 16 * h * S_W0 + 32 * o, g, t);
                if (!kScatter) {
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        if (r_lo < n) *reinterpret_cast<float2 *>(d_enc + (size_t)r_lo * 32 + 16 * h + 8 * nt + 2 * t) = make_float2(de[nt][0], de[nt][1]);
                        if (r_hi < n) *reinterpret_cast<float2 *>(d_enc + (size_t)r_hi * 32 + 16 * h + 8 * nt + 2 * t) = make_float2(de[nt][2], de[nt][3]);
                    }
                } else {
                    float xl[3] = {0.f, 0.f, 0.f}, xh[3] = {0.f, 0.f, 0.f}, pl[3], ph[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (r_lo < n) xl[k] = __ldg(pos + (size_t)r_lo * 3 + k);
                        if (r_hi < n) xh[k] = __ldg(pos + (size_t)r_hi * 3 + k);
                    }
                    hg::unit_pos<3>(xl, gd.bound, pl);
                    hg::unit_pos<3>(xh, gd.bound, ph);
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        const hg::LevelMeta m = s_meta[8 * h + 4 * nt + t];  // column 16 h + 8 nt + 2 t is feature 0 of that level
                        if (r_lo < n) scatter_sample_level(d_table, m, pl, de[nt][0], de[nt][1]);
                        if (r_hi < n) scatter_sample_level(d_table, m, ph, de[nt][2], de[nt][3]);
                    }
                }
            }
        }
    }

    // ---- flush the CTA's weight gradients
    if (it > 0 && warp < kWarps) {
        umma::mbar_wait(&bar_w0, (it - 1u) & 1u);
        umma::fence_after_sync();
        flush_weight_gradients(tmem, warp, lane, d_weights);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, kTmemCols);
}


// which backward kernel a launch takes: eight chain warps (default) or sixteen in column-split pairs (NGP_B200_MLP_BWD_SPLIT=1);
// read per launch so that a test or an A/B run can switch inside one process
static bool backward_split() {
    const char *e = getenv("NGP_B200_MLP_BWD_SPLIT");
    return e && e[0] == '1';
}

constexpr size_t kFwdSmem = kWeightFloats * sizeof(uint32_t);
constexpr size_t kFusedSmem = kFwdSmem + 16 * sizeof(hg::LevelMeta);
constexpr size_t kBwdSmem = (kWeightFloats + kActFloats) * sizeof(uint32_t);
constexpr size_t kBwdUmmaSmem = 1024 + kPanelBytes + kWeightFloats * sizeof(uint32_t);  // 224,256 B

}  // namespace
}  // namespace ngp

extern "C" {

void ngp_nerf_mlp_forward(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpNerfMlpDescriptor>(opaque, opaque_len, "nerf_mlp_forward");
    if (!d || d->n_samples == 0) return;
    BufferCursor b{buffers};
    const float *enc = b.next<const float>();
    const float *dirs = b.next<const float>();
    const float *weights = b.next<const float>();
    const uint32_t *group_counts = d->rows_per_group ? b.next<const uint32_t>() : nullptr;
    float *out = b.next<float>();
    static bool configured = false;  // benign race: the attribute call is idempotent
    if (!configured) {
        cudaFuncSetAttribute(nerf_mlp_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
        cudaFuncSetAttribute(nerf_mlp_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
        configured = true;
    }
    const unsigned tiles = div_up(d->n_samples, 16);
    const unsigned blocks = min(div_up(tiles, kWarps), 148u * 4u);
    if (d->density_only)
        nerf_mlp_forward_kernel<true><<<blocks, kThreads, kFwdSmem, stream>>>(d->n_samples, d->rows_per_group, enc, dirs, weights, group_counts, out);
    else
        nerf_mlp_forward_kernel<false><<<blocks, kThreads, kFwdSmem, stream>>>(d->n_samples, d->rows_per_group, enc, dirs, weights, group_counts, out);
    check_launch("nerf_mlp_forward");
}

static void launch_nerf_mlp_backward(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len, bool accumulate) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpNerfMlpDescriptor>(opaque, opaque_len, "nerf_mlp_backward");
    if (!d) return;
    BufferCursor b{buffers};
    const float *enc = b.next<const float>();
    const float *dirs = b.next<const float>();
    const float *weights = b.next<const float>();
    const float *d_drgbs = b.next<const float>();
    float *d_enc = b.next<float>();
    float *d_weights = b.next<float>();
    if (!accumulate) NGP_CUDA_OK(cudaMemsetAsync(d_weights, 0, kGlobalWeights * sizeof(float), stream), "nerf_mlp_backward");
    if (d->n_samples == 0) return;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(nerf_mlp_backward_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdUmmaSmem);
        cudaFuncSetAttribute(nerf_mlp_backward_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdUmmaSmem);
        configured = true;
    }
    const unsigned blocks = min(div_up(d->n_samples, kBlockSamples), 148u);
    if (backward_split())
        nerf_mlp_backward_split_kernel<false><<<blocks, kSplitThreads, kBwdUmmaSmem, stream>>>(d->n_samples, enc, dirs, weights, d_drgbs, d_enc, d_weights,
                                                                                            NgpHashGridA1Descriptor{}, nullptr, nullptr);
    else
        nerf_mlp_backward_umma_kernel<false><<<blocks, kBwdThreads, kBwdUmmaSmem, stream>>>(d->n_samples, enc, dirs, weights, d_drgbs, d_enc, d_weights,
                                                                                         NgpHashGridA1Descriptor{}, nullptr, nullptr);
    check_launch("nerf_mlp_backward");
}

// MLP backward with the hash-table scatter fused behind it: d_enc is never written, each thread scatters its fragment
// (d_table is zero-filled here first, like ngp_hashgrid_a1_backward does).
void ngp_nerf_mlp_backward_scatter(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *gd = descriptor<NgpHashGridA1Descriptor>(opaque, opaque_len, "nerf_mlp_backward_scatter");
    if (!gd) return;
    BufferCursor b{buffers};
    const float *enc = b.next<const float>();
    const float *dirs = b.next<const float>();
    const float *weights = b.next<const float>();
    const float *d_drgbs = b.next<const float>();
    const float *pos = b.next<const float>();
    float *d_weights = b.next<float>();
    float *d_table = b.next<float>();
    if (gd->dim != 3 || gd->L != 16 || gd->F != 2 || gd->wrap_T == 0 || (gd->wrap_T & (gd->wrap_T - 1u)) != 0 ||
        reinterpret_cast<uintptr_t>(d_table) % 16 != 0 || gd->offsets[gd->L] % 2 != 0 || gd->rows_per_group != 0) {
        set_error(NGP_ERR_ARGUMENT, "nerf_mlp_backward_scatter: needs dim=3 L=16 F=2, power-of-two wrap_T and a 16-byte aligned gradient table "
                  "of an even number of rows (got dim=%u L=%u F=%u wrap_T=%u); use nerf_mlp_backward + hashgrid_a1_backward", gd->dim, gd->L, gd->F, gd->wrap_T);
        return;
    }
    NGP_CUDA_OK(cudaMemsetAsync(d_weights, 0, kGlobalWeights * sizeof(float), stream), "nerf_mlp_backward_scatter");
    NGP_CUDA_OK(cudaMemsetAsync(d_table, 0, (size_t)gd->offsets[gd->L] * gd->F * sizeof(float), stream), "nerf_mlp_backward_scatter");
    if (gd->n_points == 0) return;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(nerf_mlp_backward_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdUmmaSmem);
        cudaFuncSetAttribute(nerf_mlp_backward_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdUmmaSmem);
        configured = true;
    }
    const unsigned blocks = min(div_up(gd->n_points, kBlockSamples), 148u);
    if (backward_split())
        nerf_mlp_backward_split_kernel<true><<<blocks, kSplitThreads, kBwdUmmaSmem, stream>>>(gd->n_points, enc, dirs, weights, d_drgbs, nullptr, d_weights, *gd,
                                                                                           pos, d_table);
    else
        nerf_mlp_backward_umma_kernel<true><<<blocks, kBwdThreads, kBwdUmmaSmem, stream>>>(gd->n_points, enc, dirs, weights, d_drgbs, nullptr, d_weights, *gd, pos,
                                                                                        d_table);
    check_launch("nerf_mlp_backward_scatter");
}

void ngp_nerf_mlp_backward(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    launch_nerf_mlp_backward(stream, buffers, opaque, opaque_len, false);
}

// same, ADDING to d_weights instead of defining it (a batch processed in chunks: the first chunk defines, the others add)
void ngp_nerf_mlp_backward_acc(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    launch_nerf_mlp_backward(stream, buffers, opaque, opaque_len, true);
}

void ngp_nerf_mlp_backward_mma(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpNerfMlpDescriptor>(opaque, opaque_len, "nerf_mlp_backward_mma");
    if (!d) return;
    BufferCursor b{buffers};
    const float *enc = b.next<const float>();
    const float *dirs = b.next<const float>();
    const float *weights = b.next<const float>();
    const float *d_drgbs = b.next<const float>();
    float *d_enc = b.next<float>();
    float *d_weights = b.next<float>();
    NGP_CUDA_OK(cudaMemsetAsync(d_weights, 0, kGlobalWeights * sizeof(float), stream), "nerf_mlp_backward_mma");
    if (d->n_samples == 0) return;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(nerf_mlp_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);
        configured = true;
    }
    const unsigned blocks = min(div_up(d->n_samples, kBlockSamples), 148u);
    nerf_mlp_backward_kernel<<<blocks, kThreads, kBwdSmem, stream>>>(d->n_samples, enc, dirs, weights, d_drgbs, d_enc, d_weights);
    check_launch("nerf_mlp_backward_mma");
}

void ngp_nerf_fused_forward(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpNerfFusedDescriptor>(opaque, opaque_len, "nerf_fused_forward");
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
        reinterpret_cast<uintptr_t>(table) % pair_bytes != 0 || gd.offsets[gd.L] % 2 != 0) {
        set_error(NGP_ERR_ARGUMENT,
                  "nerf_fused_forward: needs dim=3 L=16 F=2, power-of-two wrap_T and a table of an even number of rows aligned to two rows "
                  "(got dim=%u L=%u F=%u wrap_T=%u); use hashgrid_a1_forward + nerf_mlp_forward", gd.dim, gd.L, gd.F, gd.wrap_T);
        return;
    }
    if (gd.n_points == 0) return;
    const unsigned tiles = div_up(gd.n_points, 16);
    // persistent grid: 3 CTAs of 8 warps are resident per SM (80 registers, 43 KB shared memory); more CTAs than that
    // only repeat the weight staging
    const unsigned blocks = min(div_up(tiles, kWarps), 148u * 3u);
#define NGP_FUSED(TT, DO, WE)                                                                                          \
    do {                                                                                                               \
        static bool configured = false; /* benign race: idempotent */                                                  \
        if (!configured) {                                                                                             \
            cudaFuncSetAttribute(nerf_fused_forward_kernel<TT, DO, WE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem); \
            configured = true;                                                                                         \
        }                                                                                                              \
        nerf_fused_forward_kernel<TT, DO, WE><<<blocks, kThreads, kFusedSmem, stream>>>(                                \
            *d, pos, static_cast<const TT *>(table), dirs, weights, group_counts, out, enc_out);                       \
    } while (0)
    if (gd.table_dtype == 0) {
        if (d->density_only) { if (d->write_enc) NGP_FUSED(float, true, true); else NGP_FUSED(float, true, false); }
        else { if (d->write_enc) NGP_FUSED(float, false, true); else NGP_FUSED(float, false, false); }
    } else {
        if (d->density_only) { if (d->write_enc) NGP_FUSED(__half, true, true); else NGP_FUSED(__half, true, false); }
        else { if (d->write_enc) NGP_FUSED(__half, false, true); else NGP_FUSED(__half, false, false); }
    }
#undef NGP_FUSED
    check_launch("nerf_fused_forward");
}

}  // extern "C"
