// gridops.cu -- morton3d / morton3d_invert / packbits for sm_100a.
//
// Replaces deps/volume-rendering-jax/lib/impl/marching.cu:399-433,606-665 (morton) and
// packbits.cu:9-72.  All three are pure streaming kernels; packbits reads 16 B and writes 4 B of
// mask per thread and assembles the bitfield with shuffles so every global access is a full vector.
#include "common.cuh"

namespace ngp {
namespace {

constexpr int kBlock = 256;

__global__ void __launch_bounds__(kBlock) morton3d_kernel(uint32_t length,
                                                           const uint32_t *__restrict__ xyzs,
                                                           uint32_t *__restrict__ idcs) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= length) return;
    idcs[i] = morton3d_encode(__ldg(xyzs + 3 * i), __ldg(xyzs + 3 * i + 1), __ldg(xyzs + 3 * i + 2));
}

__global__ void __launch_bounds__(kBlock) morton3d_invert_kernel(uint32_t length,
                                                                  const uint32_t *__restrict__ idcs,
                                                                  uint32_t *__restrict__ xyzs) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= length) return;
    uint32_t m = __ldg(idcs + i);
    xyzs[3 * i + 0] = compact_bits10(m);
    xyzs[3 * i + 1] = compact_bits10(m >> 1);
    xyzs[3 * i + 2] = compact_bits10(m >> 2);
}

// One thread = 4 cells (float4 in, uchar4 mask out); 8 threads = one 32-bit word of the bitfield.
// bit k of byte i = density[8i+k] > threshold[8i+k]  (packbits.cu:28-33, LSB first).
template <bool kScalarThreshold>
__global__ void __launch_bounds__(kBlock) packbits_vec_kernel(uint32_t n_bytes,
                                                               const float *__restrict__ threshold,
                                                               const float *__restrict__ density,
                                                               uint8_t *__restrict__ mask,
                                                               uint8_t *__restrict__ bitfield) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;  // index of this thread's 4-cell group
    uint32_t n_quads = n_bytes * 2;
    uint32_t nibble = 0;
    if (q < n_quads) {
        float4 d = __ldg(reinterpret_cast<const float4 *>(density) + q);
        float4 t;
        if (kScalarThreshold) {
            float s = __ldg(threshold);
            t = make_float4(s, s, s, s);
        } else {
            t = __ldg(reinterpret_cast<const float4 *>(threshold) + q);
        }
        uint32_t b0 = d.x > t.x, b1 = d.y > t.y, b2 = d.z > t.z, b3 = d.w > t.w;
        reinterpret_cast<uint32_t *>(mask)[q] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        nibble = b0 | (b1 << 1) | (b2 << 2) | (b3 << 3);
    }
    uint32_t lane = threadIdx.x & 31u;
    uint32_t word = nibble << (4 * (lane & 7u));
    word |= __shfl_xor_sync(0xffffffffu, word, 1);
    word |= __shfl_xor_sync(0xffffffffu, word, 2);
    word |= __shfl_xor_sync(0xffffffffu, word, 4);
    if ((lane & 7u) == 0 && q < n_quads) {
        uint32_t byte0 = q >> 1;
        if (byte0 + 4 <= n_bytes) {
            reinterpret_cast<uint32_t *>(bitfield)[byte0 >> 2] = word;
        } else {
            for (uint32_t k = 0; byte0 + k < n_bytes; ++k) bitfield[byte0 + k] = (uint8_t)(word >> (8 * k));
        }
    }
}

// generic (any alignment) fallback: one thread per output byte
template <bool kScalarThreshold>
__global__ void __launch_bounds__(kBlock) packbits_byte_kernel(uint32_t n_bytes,
                                                                const float *__restrict__ threshold,
                                                                const float *__restrict__ density,
                                                                uint8_t *__restrict__ mask,
                                                                uint8_t *__restrict__ bitfield) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bytes) return;
    float s = kScalarThreshold ? __ldg(threshold) : 0.f;
    uint32_t byte = 0;
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) {
        float t = kScalarThreshold ? s : __ldg(threshold + i * 8 + k);
        uint32_t p = __ldg(density + i * 8 + k) > t;
        mask[i * 8 + k] = (uint8_t)p;
        byte |= p << k;
    }
    bitfield[i] = (uint8_t)byte;
}

template <bool kScalarThreshold>
void packbits_launch(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len, const char *op) {
    clear_error();
    auto *desc = descriptor<NgpPackbitsDescriptor>(opaque, opaque_len, op);
    if (!desc) return;
    if (desc->n_bytes == 0) return;
    BufferCursor b{buffers};
    const float *thr = b.next<const float>();
    const float *den = b.next<const float>();
    uint8_t *mask = b.next<uint8_t>();
    uint8_t *bits = b.next<uint8_t>();
    bool aligned = ((uintptr_t)den % 16 == 0) && ((uintptr_t)mask % 4 == 0) && ((uintptr_t)bits % 4 == 0) &&
                   (kScalarThreshold || (uintptr_t)thr % 16 == 0);
    if (aligned) {
        packbits_vec_kernel<kScalarThreshold><<<div_up((unsigned long long)desc->n_bytes * 2, kBlock), kBlock, 0, stream>>>(
            desc->n_bytes, thr, den, mask, bits);
    } else {
        packbits_byte_kernel<kScalarThreshold><<<div_up(desc->n_bytes, kBlock), kBlock, 0, stream>>>(
            desc->n_bytes, thr, den, mask, bits);
    }
    check_launch(op);
}

}  // namespace
}  // namespace ngp

extern "C" {

void ngp_pack_density_into_bits(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    ngp::packbits_launch<false>(stream, buffers, opaque, opaque_len, "pack_density_into_bits");
}

void ngp_packbits_scalar(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    ngp::packbits_launch<true>(stream, buffers, opaque, opaque_len, "packbits_scalar");
}

void ngp_morton3d(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *desc = descriptor<NgpMorton3DDescriptor>(opaque, opaque_len, "morton3d");
    if (!desc || desc->length == 0) return;
    BufferCursor b{buffers};
    const uint32_t *xyzs = b.next<const uint32_t>();
    uint32_t *idcs = b.next<uint32_t>();
    morton3d_kernel<<<div_up(desc->length, kBlock), kBlock, 0, stream>>>(desc->length, xyzs, idcs);
    check_launch("morton3d");
}

void ngp_morton3d_invert(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *desc = descriptor<NgpMorton3DDescriptor>(opaque, opaque_len, "morton3d_invert");
    if (!desc || desc->length == 0) return;
    BufferCursor b{buffers};
    const uint32_t *idcs = b.next<const uint32_t>();
    uint32_t *xyzs = b.next<uint32_t>();
    morton3d_invert_kernel<<<div_up(desc->length, kBlock), kBlock, 0, stream>>>(desc->length, idcs, xyzs);
    check_launch("morton3d_invert");  // the reference omits this check (marching.cu:656-665)
}

}  // extern "C"
