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

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
#endif

}  // namespace ngp
