// common.cuh -- shared host/device helpers of libngp_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/ngp_b200.h"

namespace ngp {

// ---------------------------------------------------------------- status (never throws)
enum : int {
    NGP_OK = 0,
    NGP_ERR_DESCRIPTOR = -1,  // opaque_len does not match the descriptor struct
    NGP_ERR_ARGUMENT = -2,    // unsupported static argument (F, dim, L ...)
    NGP_ERR_WORKSPACE = -3,   // scratch allocation failed
};

void set_error(int status, const char *fmt, ...);
void clear_error();

// returns nullptr (and records NGP_ERR_DESCRIPTOR) on size mismatch; mirrors serde.h:35-40
template <typename T>
inline const T *descriptor(const char *opaque, size_t opaque_len, const char *op) {
    if (opaque_len != sizeof(T)) {
        set_error(NGP_ERR_DESCRIPTOR, "%s: invalid opaque object size, expected %zu, got %zu", op,
                  sizeof(T), opaque_len);
        return nullptr;
    }
    return reinterpret_cast<const T *>(opaque);
}

inline bool check_launch(const char *op) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error((int)e, "%s: %s", op, cudaGetErrorString(e));
        return false;
    }
    return true;
}

#define NGP_CUDA_OK(expr, op)                                                   \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) {                                                \
            ngp::set_error((int)_e, "%s: %s failed: %s", op, #expr, cudaGetErrorString(_e)); \
            return;                                                             \
        }                                                                       \
    } while (0)

// Per-(device, stream) scratch block, created on first use and grown geometrically.  Work
// enqueued on one stream is ordered, so reusing the block across calls on that stream is safe;
// different streams / host threads get different blocks.
void *workspace(cudaStream_t stream, size_t bytes);

struct BufferCursor {
    void **buffers;
    int i = 0;
    template <typename T>
    T *next() { return static_cast<T *>(buffers[i++]); }
};

inline unsigned div_up(unsigned long long a, unsigned b) { return (unsigned)((a + b - 1) / b); }

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {  // marching.cu:52-58
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t morton3d_encode(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits10(x) | (expand_bits10(y) << 1) | (expand_bits10(z) << 2);
}
__device__ __forceinline__ uint32_t compact_bits10(uint32_t x) {  // marching.cu:70-77
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }

// vectorised no-return reduction: one 8-byte L2 atomic instead of two (sm_90+)
__device__ __forceinline__ void red_add_v2(float *addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
// Huber loss (optax.huber_loss, mean over the three channels) of one ray's colour against its ground-truth pixel
// composited onto the ray's background (utils/data.py:459-463), and its gradient scaled by inv_n = 1 / n_valid
// (app/nerf/_utils.py:151-165).  Shared by huber_loss_grad (trainops.cu) and the fused integrate + loss kernel.
__device__ __forceinline__ float4 huber_ray(float4 pred, uchar4 px, float bg0, float bg1, float bg2, float delta, float inv_n,
                                            float &per_ray) {
    const float a = (float)px.w / 255.f;
    const float gt[3] = {(float)px.x / 255.f, (float)px.y / 255.f, (float)px.z / 255.f};
    const float pr[3] = {pred.x, pred.y, pred.z}, bg[3] = {bg0, bg1, bg2};
    float gr[3];
    per_ray = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float target = gt[c] * a + bg[c] * (1.f - a);
        const float err = pr[c] - target;
        const float ae = fabsf(err), q = fminf(ae, delta);
        per_ray += 0.5f * q * q + delta * (ae - q);
        gr[c] = fminf(fmaxf(err, -delta), delta) * (1.f / 3.f) * inv_n;
    }
    per_ray *= (1.f / 3.f);
    return make_float4(gr[0], gr[1], gr[2], 0.f);
}

// ---------------------------------------------------------------- counter-based random numbers
// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11): the random inputs of the path
// (march perturbations models/renderers/cuda.py:118-122, random backgrounds app/nerf/_utils.py:134-136, the cell draws and
// jitter of utils/types.py:1170-1206) are drawn by the reference with jax.random OUTSIDE its ops; here they are a pure
// function of (seed, stream, call counter, element index), so a captured CUDA graph and an eager replay see the same
// numbers.  Known-answer vectors of the Random123 distribution are checked in tests/test_rng.py.
struct Philox4 { uint32_t x, y, z, w; };
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}
// 23 random mantissa bits under exponent 0 -> [1, 2) -> [0, 1): jax.random.uniform's construction
__host__ __device__ __forceinline__ float bits_to_unit_float(uint32_t bits) {
#ifdef __CUDA_ARCH__
    return __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
#else
    uint32_t u = (bits >> 9) | 0x3F800000u;
    float f;
    memcpy(&f, &u, 4);
    return f - 1.0f;
#endif
}
// the four uniforms of element `i` of call `counter` on `stream_id` (NgpRngDescriptor)
__device__ __forceinline__ float4 philox_uniform4(uint32_t i, uint32_t counter, uint32_t stream_id, uint32_t seed_lo, uint32_t seed_hi) {
    const Philox4 r = philox4x32_10(i, counter, stream_id, 0u, seed_lo, seed_hi);
    return make_float4(bits_to_unit_float(r.x), bits_to_unit_float(r.y), bits_to_unit_float(r.z), bits_to_unit_float(r.w));
}
// A call counter that lives in device memory and is bumped by the LAST block of the launch that consumed it, so that
// replaying a captured graph draws fresh numbers each time: state = {counter, blocks_done}.  Every block reads
// state[0] when it starts; the last one to finish increments it and re-arms the ticket.
__device__ __forceinline__ void rng_state_finish(uint32_t *state) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(state + 1, 1u) == gridDim.x - 1u) {
            state[1] = 0u;
            __threadfence();
            atomicAdd(state, 1u);
        }
    }
}

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
#endif

}  // namespace ngp
