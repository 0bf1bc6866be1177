// exchange.cu -- the per-step gradient exchange of ray-sharded training fused with the optimizer:
// reduce-scatter + Adam + all-gather as ONE kernel per rank over NVLink peer / multicast memory.
//
//   every rank r owns elements [shard_begin, shard_begin + n) of the flat buffers [hash table | MLP weights]
//   1. start barrier: block b of every rank tells block b of every peer "my gradient buffer is complete"
//   2. owner reads the SUM of its shard:   multimem.ld_reduce.add.v4.f32 through the NVSwitch (one 16-byte request,
//      reduced in the switch), or `world` peer loads added in rank order when no multicast mapping exists
//   3. Adam on the shard (m, v never leave the owner; same arithmetic as ngp_adam_step, adam.cuh)
//   4. owner stores the new parameters into every replica: multimem.st.v4.f32, or one st per peer
//   5. end barrier (release/acquire at system scope): every replica is complete and nobody still reads the
//      gradients when any rank's kernel ends, so the next step may overwrite them
//
// Link traffic per rank and step: (world-1)/world of the buffer in, the same out -- what NCCL's reduce-scatter and
// all-gather move -- but in one launch, without the two collective launches and the HBM round trip of the reduced
// shard and of the updated shard in between.  The kernel occupies `n_blocks` SMs only (NVLink saturates long before
// the SMs do); the next batch's march runs on the others.
//
// The barrier is the usual signal-pad handshake: word [signal_base + block * world + sender] of the RECEIVER's pad is
// flipped 0 -> 1 by the sender (system-scope CAS, spinning while it is still 1 from the previous round) and 1 -> 0 by
// the receiver, so the pads need zeroing once and the two barriers of a launch -- and consecutive launches -- can share
// their words.  Every rank launches the same grid, so block b exists on every rank; blocks only wait on flags in
// memory, never on co-residency.
#include "adam.cuh"

namespace ngp {
namespace {

static_assert(sizeof(NgpAdamDescriptor) == 64 && sizeof(NgpAdamExchangeDescriptor) == 96, "descriptor wire format");
constexpr int kMaxWorld = 8;
constexpr int kThreads = 256;  // x 128 registers = half an SM: the next batch's march (or a second CTA) fits beside it.  (512 threads
                               // per CTA: the per-peer-load flavour alone drops from 0.156 to 0.108 ms at N = 2, but the march prefetched under
                               // the exchange no longer fits on the SM and the training step goes from 0.586 to 0.662 ms.)
// 16-byte remote requests per thread in flight.  Per-peer loads want 8 (0.240 -> 0.156 ms at N = 2); the switch-reduced
// flavour is no faster for it (0.158 / 0.157 ms) and the extra registers cost the overlapped march its place on the SM
// (training step 0.572 -> 0.586 ms at N = 2), so it keeps 4.
constexpr int kUnrollMultimem = 4, kUnrollPeer = 8;

__device__ __forceinline__ uint32_t cas_release_sys(uint32_t *addr, uint32_t expect, uint32_t desired) {
    uint32_t old;
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(desired) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t *addr, uint32_t expect, uint32_t desired) {
    uint32_t old;
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(desired) : "memory");
    return old;
}

__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// A peer that never arrives (crashed rank, mismatched launch) must not hang the GPU: after `timeout_ns` the waiting
// thread reports and traps, which fails this rank's launch with an error the host sees at its next synchronisation.
__device__ __noinline__ void peer_timeout(uint32_t rank, uint32_t peer, uint32_t block, int phase) {
    printf("[ngp_b200] adam_step_exchange: rank %u block %u gave up waiting for rank %u (%s)\n", rank, block, peer,
           phase == 0 ? "its pad word is still set from the previous round" : "it never signalled");
    __trap();
}

// Block b of this rank meets block b of every peer.  Writes of the whole block before the call are visible to every
// peer thread after its matching call returns (bar.sync orders them before the releasing CAS, which is cumulative).
__device__ __forceinline__ void meet_peers(uint32_t *const *signal_ptrs, uint32_t slot0, uint32_t rank, uint32_t world,
                                           uint64_t timeout_ns) {
    __syncthreads();
    if (threadIdx.x < world) {
        const uint32_t peer = threadIdx.x;
        uint32_t *theirs = signal_ptrs[peer] + slot0 + rank;  // my word in the peer's pad
        uint32_t *mine = signal_ptrs[rank] + slot0 + peer;    // the peer's word in my pad
        const uint64_t t0 = global_ns();
        for (uint32_t spins = 0; cas_release_sys(theirs, 0u, 1u) != 0u; ++spins)
            if ((spins & 1023u) == 1023u && global_ns() - t0 > timeout_ns) peer_timeout(rank, peer, blockIdx.x, 0);
        for (uint32_t spins = 0; cas_acquire_sys(mine, 1u, 0u) != 1u; ++spins)
            if ((spins & 1023u) == 1023u && global_ns() - t0 > timeout_ns) peer_timeout(rank, peer, blockIdx.x, 1);
    }
    __syncthreads();
}

__device__ __forceinline__ float4 multimem_sum4(const float4 *mc_addr) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc_addr) : "memory");
    return r;
}
__device__ __forceinline__ void multimem_store4(float4 *mc_addr, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// peer memory is not cached in the local L2 (only L1, which a relaxed.sys access bypasses): a plain 16-byte request
__device__ __forceinline__ float4 peer_load4(const float4 *addr) {
    float4 r;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void peer_store4(float4 *addr, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool kMultimem>
__global__ void __launch_bounds__(kThreads, 2) adam_exchange_kernel(
    NgpAdamExchangeDescriptor d, const uint32_t *__restrict__ step_ptr, float *__restrict__ m, float *__restrict__ v,
    const uint64_t *__restrict__ grads_ptrs, const uint64_t *__restrict__ params_ptrs,
    const uint64_t *__restrict__ signal_ptrs, const float *grads_mc, float *params_mc) {
    __shared__ uint32_t *s_signal[kMaxWorld];
    __shared__ const float4 *s_grads[kMaxWorld];
    __shared__ float4 *s_params[kMaxWorld];
    if (threadIdx.x < d.world) {
        s_signal[threadIdx.x] = reinterpret_cast<uint32_t *>(signal_ptrs[threadIdx.x]);
        s_grads[threadIdx.x] = reinterpret_cast<const float4 *>(grads_ptrs[threadIdx.x]) + d.shard_begin / 4;
        s_params[threadIdx.x] = reinterpret_cast<float4 *>(params_ptrs[threadIdx.x]) + d.shard_begin / 4;
    }
    const uint32_t slot0 = d.signal_base + blockIdx.x * d.world;
    const uint64_t timeout_ns = (uint64_t)(d.timeout_ms ? d.timeout_ms : 20000u) * 1000000ull;
    meet_peers(s_signal, slot0, d.rank, d.world, timeout_ns);  // (1) every peer's backward has written its gradients

    const AdamStepConstants c = adam_step_constants(d.adam, __ldg(step_ptr));
    const float4 *g_mc = reinterpret_cast<const float4 *>(grads_mc) + d.shard_begin / 4;
    float4 *p_mc = reinterpret_cast<float4 *>(params_mc) + d.shard_begin / 4;
    const float4 *p_own = s_params[d.rank];
    float4 *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
    const size_t n4 = d.adam.n / 4;
    const size_t stride = (size_t)gridDim.x * kThreads;
    constexpr int kUnroll = kMultimem ? kUnrollMultimem : kUnrollPeer;
    constexpr int kLocal = 4;  // local state per pass
    for (size_t base = (size_t)blockIdx.x * kThreads + threadIdx.x; base < n4; base += stride * kUnroll) {
        float4 g[kUnroll], p[kLocal], mm[kLocal], vv[kLocal];
        // (2) all the remote requests of this round first: kUnroll x 16 B per thread in flight over the link
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const size_t i = base + u * stride;
            if (i < n4) {
                if constexpr (kMultimem) {
                    g[u] = multimem_sum4(g_mc + i);
                } else {
                    g[u] = peer_load4(s_grads[0] + i);
                    for (uint32_t r = 1; r < d.world; ++r) {  // rank order: the same sum on every launch
                        const float4 t = peer_load4(s_grads[r] + i);
                        g[u].x += t.x, g[u].y += t.y, g[u].z += t.z, g[u].w += t.w;
                    }
                }
            }
        }
        // local state and update kLocal at a time: the registers of one pass are free before the next pass's loads
#pragma unroll
        for (int h = 0; h < kUnroll; h += kLocal) {
#pragma unroll
            for (int u = h; u < h + kLocal; ++u) {
                const size_t i = base + u * stride;
                if (i < n4) {
                    p[u - h] = p_own[i];
                    mm[u - h] = __ldcs(m4 + i);
                    vv[u - h] = __ldcs(v4 + i);
                }
            }
#pragma unroll
            for (int u = h; u < h + kLocal; ++u) {
                const size_t i = base + u * stride;
                if (i < n4) {
                    const float wd = (i * 4 >= d.adam.decay_begin) ? d.adam.weight_decay : 0.f;
                    adam_update4(d.adam, c, wd, p[u - h], g[u], mm[u - h], vv[u - h]);  // (3)
                    __stcs(m4 + i, mm[u - h]);
                    __stcs(v4 + i, vv[u - h]);
                    if constexpr (kMultimem) {  // (4) one store, replicated by the switch into every rank's buffer
                        multimem_store4(p_mc + i, p[u - h]);
                    } else {
                        for (uint32_t r = 0; r < d.world; ++r) peer_store4(s_params[r] + i, p[u - h]);
                    }
                }
            }
        }
    }
    meet_peers(s_signal, slot0, d.rank, d.world, timeout_ns);  // (5) replicas complete, gradients no longer read
}

}  // namespace
}  // namespace ngp

extern "C" void ngp_adam_step_exchange(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpAdamExchangeDescriptor>(opaque, opaque_len, "adam_step_exchange");
    if (!d) return;
    if (d->adam.n % 4 != 0 || d->adam.decay_begin % 4 != 0 || d->shard_begin % 4 != 0) {
        set_error(NGP_ERR_ARGUMENT, "adam_step_exchange: n, decay_begin and shard_begin must be multiples of 4, got %llu, %llu, %llu",
                  (unsigned long long)d->adam.n, (unsigned long long)d->adam.decay_begin, (unsigned long long)d->shard_begin);
        return;
    }
    if (d->world < 1 || d->world > (uint32_t)kMaxWorld || d->rank >= d->world) {
        set_error(NGP_ERR_ARGUMENT, "adam_step_exchange: rank %u of world %u (1..%d ranks supported)", d->rank, d->world, kMaxWorld);
        return;
    }
    if (d->n_blocks < 1 || d->n_blocks > 148u) {  // every block must become resident while its peers spin on it
        set_error(NGP_ERR_ARGUMENT, "adam_step_exchange: n_blocks must be in 1..148, got %u", d->n_blocks);
        return;
    }
    BufferCursor b{buffers};
    const uint32_t *step = b.next<const uint32_t>();
    float *m = b.next<float>();
    float *v = b.next<float>();
    const uint64_t *grads_ptrs = b.next<const uint64_t>();
    const uint64_t *params_ptrs = b.next<const uint64_t>();
    const uint64_t *signal_ptrs = b.next<const uint64_t>();
    const float *grads_mc = b.next<const float>();
    float *params_mc = b.next<float>();
    if (d->use_multimem && (!grads_mc || !params_mc)) {
        set_error(NGP_ERR_ARGUMENT, "adam_step_exchange: use_multimem needs the multicast bases of both buffers");
        return;
    }
    // the launch happens even for an empty shard: the peers wait for this rank at both barriers
    if (d->use_multimem)
        adam_exchange_kernel<true><<<d->n_blocks, kThreads, 0, stream>>>(*d, step, m, v, grads_ptrs, params_ptrs, signal_ptrs, grads_mc, params_mc);
    else
        adam_exchange_kernel<false><<<d->n_blocks, kThreads, 0, stream>>>(*d, step, m, v, grads_ptrs, params_ptrs, signal_ptrs, grads_mc, params_mc);
    check_launch("adam_step_exchange");
}
