// umma.cuh -- thin inline-PTX layer over the sm_100a tensor-core path: tcgen05.mma (kind::tf32) with operands in
// shared memory and accumulators in tensor memory (TMEM), tcgen05.ld for the epilogues, mbarrier completion.
//
// Operand format ("panel"): rows of 32 floats = 128 bytes, panel base 1024-byte aligned.
//   * K-major  (rows = M or N index, the 32 floats of a row = 32 consecutive K): SWIZZLE_128B -- inside each row the
//              16-byte chunk c of row r sits at chunk position c ^ (r & 7) (address bits [4,7) ^= bits [7,10)).
//              Forward / dgrad A operands, dgrad B operands (weights W[in][out] with N = in, K = out).
//   * MN-major (rows = K index, the 32 floats of a row = 32 consecutive M or N): for tf32 only
//              SWIZZLE_128B_BASE32B exists (32-byte chunks, c ^ (r & 3)).  Forward B operands (weights with N = out,
//              K = in) and both wgrad operands (activations / deltas with K = the sample index).
// The two swizzles differ, so a tf32 tensor read in both orientations needs two shared-memory copies (16-bit
// operands use SWIZZLE_128B for both).  That shapes the kernels: the forward (mlp_fwd_umma.cu) only needs K-major
// activations + MN-major weights; the backward (mlp.cu) keeps its per-warp input-gradient chain on mma.sync and gives
// the tensor core the weight gradients, whose operands (activations, deltas) are both MN-major -- activations +
// deltas + weights in BOTH formats would exceed 227 KB at 128-sample tiles.
#pragma once
#include <cstdint>

namespace ngp {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r, column j < 32) inside a panel
__device__ __forceinline__ uint32_t panel_offset(uint32_t r, uint32_t j) {
    return r * 128u + ((((j >> 2) ^ (r & 7u)) << 4) | ((j & 3u) << 2));
}

// byte offset of element (K-row r, column j < 32) inside an MN-major tf32 panel.  For 32-bit operands the MN-major
// orientation exists in ONE shared-memory format only, SWIZZLE_128B_BASE32B: 128-byte rows, 4-row atoms, and the
// 32-byte chunk c of row r stored at chunk position c ^ (r & 3) (address bits [5,7) ^= bits [7,9)) -- NOT the
// K-major SWIZZLE_128B pattern above, so a tf32 tensor used in both orientations needs two copies.
__device__ __forceinline__ uint32_t panel_offset_mn32(uint32_t r, uint32_t j) {
    return r * 128u + ((((j >> 3) ^ (r & 3u)) << 5) | ((j & 7u) << 2));
}

// ---- shared-memory matrix descriptors (64 bit): start address, leading / stride byte offsets (16-byte units),
//      version = 1 (Blackwell), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
// K-major operand: `rows` x 8 K-values per instruction, taken at K offset 8*kstep (< 32) of the panel at `panel_addr`
__device__ __forceinline__ uint64_t desc_k_major(uint32_t panel_addr, uint32_t kstep) {
    return make_desc(panel_addr + kstep * 32u, 16u, 1024u);
}
// MN-major operand: 8 K-rows per instruction = the 8-row group `kstep` of the panel(s); consecutive blocks of 32
// M/N values are `mn_block_stride` bytes apart (the panel stride)
__device__ __forceinline__ uint64_t desc_mn_major(uint32_t panel_addr, uint32_t kstep, uint32_t mn_block_stride) {
    return make_desc(panel_addr + kstep * 1024u, mn_block_stride, 512u, 1 /* SWIZZLE_128B_BASE32B: 4-row atoms */);
}

// ---- instruction descriptor (32 bit) for kind::tf32 with f32 accumulation
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) /* D = f32 */ | (2u << 7) /* A = tf32 */ | (2u << 10) /* B = tf32 */ | ((a_mn_major ? 1u : 0u) << 15) |
           ((b_mn_major ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// Same with the A operand in tensor memory: A[M x 8] = lanes 0 .. M-1, columns tmem_a .. tmem_a + 7 (one 32-bit column
// per tf32 value; K-major only).  The accumulator of one layer, rewritten in place by the epilogue threads
// (tcgen05.ld -> activation -> tcgen05.st), is the next layer's A operand: the activation chain never touches shared memory.
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// all MMAs issued so far by this thread arrive on the mbarrier when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// transaction-counting use of an mbarrier: one arrival that also announces `bytes` of asynchronous copies
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bulk copy global -> shared through the async proxy (no thread touches the data); completes `bytes` on the mbarrier.
// Addresses and size are multiples of 16 bytes.
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_global, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_global), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- fences
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (the tensor core reads operands through it)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tensor memory: allocation by one full warp; the base address (lane << 16 | column) lands in shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t n_cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(n_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t n_cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(n_cols) : "memory");
}

// ---- TMEM -> registers: warp w of a warpgroup reads lanes 32*(w%4) .. +31, thread i its lane's consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// ---- registers -> TMEM: thread i of warp w writes consecutive columns of lane 32*(w%4) + i
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

}  // namespace umma
}  // namespace ngp
