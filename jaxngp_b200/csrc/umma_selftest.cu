// umma_selftest.cu -- known-answer test of the tcgen05 operand formats used by the tensor-core MLP (umma.cuh):
// one CTA computes, with kind::tf32 MMAs on swizzled shared-memory panels and accumulators in TMEM,
//   D1 = A  . W        [128 x 64]   forward shape : A K-major,  B = W[in][out] MN-major      (M=128, N=64, K=32)
//   D2 = G  . W^T      [128 x 32]   dgrad shape   : A K-major,  B = W[in][out] K-major       (M=128, N=32, K=64)
//   D3 = G^T . A       [ 64 x 32]   wgrad shape   : both operands MN-major, K = 128 samples  (M= 64, N=32, K=128)
// (K-major operands in SWIZZLE_128B panels, MN-major ones in SWIZZLE_128B_BASE32B panels)
// and writes them out through tcgen05.ld.  tests/ compares with torch.  Diagnostic entry point, not on the hot path.
#include "common.cuh"
#include "umma.cuh"

namespace ngp {
namespace {

__global__ void __launch_bounds__(128) umma_selftest_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                            const float *__restrict__ G, float *__restrict__ D1,
                                                            float *__restrict__ D2, float *__restrict__ D3) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte aligned panels
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sA = base;                   // K-major  [128 rows][32]          16 KB
    uint8_t *sW = sA + 16384;             // K-major  2 panels [32 rows][32]   8 KB   (panel p = out columns 32p .. 32p+31)
    uint8_t *sG = sW + 8192;              // K-major  2 panels [128 rows][32] 32 KB
    uint8_t *mA = sG + 32768;             // MN-major copies of the same three matrices (rows = K index)
    uint8_t *mW = mA + 16384;
    uint8_t *mG = mW + 8192;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_slot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;

    for (uint32_t i = tid; i < 128 * 32; i += 128) {
        const uint32_t r = i / 32, j = i % 32;
        *reinterpret_cast<uint32_t *>(sA + umma::panel_offset(r, j)) = umma::to_tf32(A[i]);
        *reinterpret_cast<uint32_t *>(mA + umma::panel_offset_mn32(r, j)) = umma::to_tf32(A[i]);
    }
    for (uint32_t i = tid; i < 32 * 64; i += 128) {
        const uint32_t r = i / 64, c = i % 64;
        *reinterpret_cast<uint32_t *>(sW + (c / 32) * 4096 + umma::panel_offset(r, c % 32)) = umma::to_tf32(W[i]);
        *reinterpret_cast<uint32_t *>(mW + (c / 32) * 4096 + umma::panel_offset_mn32(r, c % 32)) = umma::to_tf32(W[i]);
    }
    for (uint32_t i = tid; i < 128 * 64; i += 128) {
        const uint32_t r = i / 64, c = i % 64;
        *reinterpret_cast<uint32_t *>(sG + (c / 32) * 16384 + umma::panel_offset(r, c % 32)) = umma::to_tf32(G[i]);
        *reinterpret_cast<uint32_t *>(mG + (c / 32) * 16384 + umma::panel_offset_mn32(r, c % 32)) = umma::to_tf32(G[i]);
    }
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(&tmem_base_slot, 128);
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_base_slot;
    const uint32_t t_d1 = tmem, t_d2 = tmem + 64, t_d3 = tmem + 96;

    if (tid == 0) {
        const uint32_t a = umma::smem_u32(sA), w = umma::smem_u32(sW), g = umma::smem_u32(sG);
        const uint32_t ma = umma::smem_u32(mA), mw = umma::smem_u32(mW), mg = umma::smem_u32(mG);
        constexpr uint32_t i1 = umma::make_idesc(128, 64, false, true);
        constexpr uint32_t i2 = umma::make_idesc(128, 32, false, false);
        constexpr uint32_t i3 = umma::make_idesc(64, 32, true, true);
        for (uint32_t ks = 0; ks < 4; ++ks)  // K = 32 input features
            umma::mma_tf32(t_d1, umma::desc_k_major(a, ks), umma::desc_mn_major(mw, ks, 4096), i1, ks > 0);
        for (uint32_t ks = 0; ks < 8; ++ks)  // K = 64 output features, two panels
            umma::mma_tf32(t_d2, umma::desc_k_major(g + (ks / 4) * 16384, ks % 4), umma::desc_k_major(w + (ks / 4) * 4096, ks % 4),
                           i2, ks > 0);
        for (uint32_t ks = 0; ks < 16; ++ks)  // K = 128 samples
            umma::mma_tf32(t_d3, umma::desc_mn_major(mg, ks, 16384), umma::desc_mn_major(ma, ks, 16384), i3, ks > 0);
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();

    const uint32_t lane_base = (warp * 32u) << 16;
    uint32_t v[32];
    umma::tmem_ld32(t_d1 + lane_base, v);
    umma::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D1[tid * 64 + j] = __uint_as_float(v[j]);
    umma::tmem_ld32(t_d1 + lane_base + 32, v);
    umma::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D1[tid * 64 + 32 + j] = __uint_as_float(v[j]);
    umma::tmem_ld32(t_d2 + lane_base, v);
    umma::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D2[tid * 32 + j] = __uint_as_float(v[j]);
    // M = 64 accumulators live in lanes 0-15 of each 32-lane quarter: row m -> lane (m % 16) + 32 * (m / 16)
    umma::tmem_ld32(t_d3 + lane_base, v);
    umma::tmem_ld_wait();
    if ((tid & 31u) < 16u)
        for (int j = 0; j < 32; ++j) D3[(warp * 16 + (tid & 31u)) * 32 + j] = __uint_as_float(v[j]);

    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 128);
}

}  // namespace
}  // namespace ngp

extern "C" void ngp_umma_selftest(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    (void)opaque;
    if (opaque_len != 0) {
        set_error(NGP_ERR_DESCRIPTOR, "umma_selftest: takes no descriptor");
        return;
    }
    BufferCursor b{buffers};
    const float *A = b.next<const float>();
    const float *W = b.next<const float>();
    const float *G = b.next<const float>();
    float *D1 = b.next<float>();
    float *D2 = b.next<float>();
    float *D3 = b.next<float>();
    constexpr int smem = 2 * (16384 + 8192 + 32768) + 1024;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    umma_selftest_kernel<<<1, 128, smem, stream>>>(A, W, G, D1, D2, D3);
    check_launch("umma_selftest");
}
