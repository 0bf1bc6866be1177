// ogrid.cu -- density-grid update of NeRFState.update_ogrid_density / threshold_ogrid
// (utils/types.py:1149-1239) as device-side steps with no host round trips:
//
//   sample positions : cell index -> morton invert -> [-1,1] -> scale to the cascade -> jitter   (:1193-1206)
//   decay + max      : density[alive] *= 0.95 ; density[idx] = max(density[idx], new)             (:1162-1164,1219-1221)
//   threshold        : thr = min(thr_max, mean over alive cells of cascade 0)                      (:1229-1230,143-144)
//
// followed by ngp_packbits_scalar (gridops.cu) reading `thr` from device memory.  The reference's
// `.at[idx].set(max(...))` with duplicate indices is last-writer-wins; here it is a true
// atomic max (exact when indices are unique, >= the reference otherwise).
#include "common.cuh"

namespace ngp {
namespace {

constexpr int kBlock = 256;
constexpr int kScanBlock = 1024;
constexpr uint32_t kChunkWords = 32, kChunkCells = 1024;  // cell draws: the bitfield in chunks of 1024 cells

// occupied cells per chunk, one thread per chunk (grids up to 256^3: the draw kernel scans these counts in shared memory)
__global__ void __launch_bounds__(kBlock) ogrid_chunk_count_kernel(uint32_t n_chunks, uint32_t n_words, const uint32_t *__restrict__ bits,
                                                                   uint32_t *__restrict__ counts) {
    const uint32_t c = blockIdx.x * kBlock + threadIdx.x;
    if (c >= n_chunks) return;
    const uint32_t w0 = c * kChunkWords;
    uint32_t cnt = 0;
    if (w0 + kChunkWords <= n_words) {
        const uint4 *p = reinterpret_cast<const uint4 *>(bits + w0);
#pragma unroll
        for (int k = 0; k < (int)kChunkWords / 4; ++k) {
            const uint4 v = __ldg(p + k);
            cnt += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
        }
    } else {
        for (uint32_t w = w0; w < n_words; ++w) cnt += __popc(__ldg(bits + w));
    }
    counts[c] = cnt;
}

// sample position of cell `m` (Morton index inside the cascade) with jitter draws u (:1193-1206)
__device__ __forceinline__ void ogrid_cell_position(uint32_t m, uint32_t G, float mip_bound, const float (&u)[3], float *out) {
    const float half_cell = __fdiv_rn(mip_bound, (float)G);             // :1196
    const float span = __fadd_rn(mip_bound, -half_cell);                // :1197
    const float inv = (float)(G - 1u);
    const uint32_t c[3] = {compact_bits10(m), compact_bits10(m >> 1), compact_bits10(m >> 2)};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float x = __fadd_rn(__fmul_rn(__fdiv_rn((float)c[k], inv), 2.f), -1.f);  // :1194
        x = __fmul_rn(x, span);
        // jax.random.uniform(minval=-h, maxval=h): max(minval, u * (maxval - minval) + minval)
        const float jitter = fmaxf(-half_cell, __fadd_rn(__fmul_rn(u[k], __fadd_rn(half_cell, half_cell)), -half_cell));
        out[k] = __fadd_rn(x, jitter);
    }
}

__global__ void __launch_bounds__(kBlock) ogrid_sample_positions_kernel(NgpOgridSampleDescriptor d,
                                                                         const uint32_t *__restrict__ idx,
                                                                         const float *__restrict__ uniforms,
                                                                         float *__restrict__ coords) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= d.n_points) return;
    const float u[3] = {__ldg(uniforms + 3 * (size_t)i), __ldg(uniforms + 3 * (size_t)i + 1), __ldg(uniforms + 3 * (size_t)i + 2)};
    float xyz[3];
    ogrid_cell_position(__ldg(idx + i), d.G, d.mip_bound, u, xyz);
    coords[3 * (size_t)i + 0] = xyz[0];
    coords[3 * (size_t)i + 1] = xyz[1];
    coords[3 * (size_t)i + 2] = xyz[2];
}

__global__ void __launch_bounds__(kBlock) ogrid_decay_kernel(uint32_t n, float decay, const float *__restrict__ in,
                                                              float *__restrict__ out) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const float v = in[i];
    out[i] = v >= 0.f ? v * decay : v;  // dead cells carry -1 and are never touched (utils/types.py:1339-1343)
}

__global__ void __launch_bounds__(kBlock) ogrid_scatter_max_kernel(uint32_t m, uint32_t n_cells,
                                                                    const uint32_t *__restrict__ idx,
                                                                    const float *__restrict__ vals,
                                                                    float *__restrict__ grid) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= m) return;
    const uint32_t c = __ldg(idx + i);
    const float v = __ldg(vals + i);
    if (c >= n_cells || !(v >= 0.f)) return;
    // a culled cell carries -1 for good (utils/types.py:1339-1343): an integer max would overwrite its negative bit
    // pattern with any density >= 0.  Nothing else writes the cell in this launch besides other maxima.
    if (grid[c] < 0.f) return;
    // non-negative floats order like their bit patterns; the cell holds a non-negative value too
    atomicMax(reinterpret_cast<int *>(grid) + c, __float_as_int(v));
}

// mean over alive (>= 0) cells + threshold; two-level deterministic reduction, last block finalises.  Each thread sums
// at most kThrPerThread values in f32 (float4 loads) before widening: f64 adds are scarce on this part, and a short f32
// run loses nothing against the f32 mean the reference takes (utils/types.py:143-144).
constexpr int kThrPerThread = 16;
__global__ void __launch_bounds__(kBlock) ogrid_threshold_kernel(uint32_t n, float thr_max, const float *__restrict__ grid,
                                                                  double *__restrict__ partial_sum,
                                                                  unsigned long long *__restrict__ partial_cnt,
                                                                  uint32_t *__restrict__ done, float *__restrict__ thr_out) {
    __shared__ double s_sum[kBlock / 32];
    __shared__ unsigned long long s_cnt[kBlock / 32];
    __shared__ bool s_last;
    double sum = 0.0;
    unsigned long long cnt = 0;
    const uint32_t n4 = (reinterpret_cast<uintptr_t>(grid) % 16 == 0) ? n / 4 : 0;
    for (uint32_t base = blockIdx.x * kBlock + threadIdx.x; base < n4; base += gridDim.x * kBlock * (kThrPerThread / 4)) {
        float part = 0.f;
        uint32_t c = 0;
#pragma unroll
        for (int u = 0; u < kThrPerThread / 4; ++u) {
            const uint32_t i = base + u * gridDim.x * kBlock;
            if (i < n4) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(grid) + i);
                if (v.x >= 0.f) { part += v.x; ++c; }
                if (v.y >= 0.f) { part += v.y; ++c; }
                if (v.z >= 0.f) { part += v.z; ++c; }
                if (v.w >= 0.f) { part += v.w; ++c; }
            }
        }
        sum += (double)part;
        cnt += c;
    }
    for (uint32_t i = n4 * 4 + blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
        const float v = __ldg(grid + i);
        if (v >= 0.f) { sum += (double)v; ++cnt; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (lane == 0) { s_sum[warp] = sum; s_cnt[warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bs = 0.0;
        unsigned long long bc = 0;
        for (int w = 0; w < kBlock / 32; ++w) { bs += s_sum[w]; bc += s_cnt[w]; }
        partial_sum[blockIdx.x] = bs;
        partial_cnt[blockIdx.x] = bc;
        __threadfence();
        s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && warp == 0) {  // fixed tree over the per-block partials: deterministic
        double ts = 0.0;
        unsigned long long tc = 0;
        for (uint32_t b = lane; b < gridDim.x; b += 32u) {
            ts += reinterpret_cast<volatile double *>(partial_sum)[b];
            tc += reinterpret_cast<volatile unsigned long long *>(partial_cnt)[b];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ts += __shfl_xor_sync(0xffffffffu, ts, o);
            tc += __shfl_xor_sync(0xffffffffu, tc, o);
        }
        if (lane == 0) {
            const float mean = tc ? (float)(ts / (double)tc) : 0.f;
            *thr_out = fminf(thr_max, mean);
            *done = 0u;
        }
    }
}

// ---------------------------------------------------------------- cell draws of update_ogrid_density (:1166-1206)
// The bitfield of a cascade (Morton order, LSB first: bit i = cell i) is cut into chunks of 1024 cells (32 words);
// `prefix[c]` = occupied cells in chunks [0, c).  One CTA scans them: 2048 chunks at G = 128.

__global__ void __launch_bounds__(kScanBlock) ogrid_chunk_prefix_kernel(uint32_t n_chunks, uint32_t n_words,
                                                                         const uint32_t *__restrict__ bits,
                                                                         uint32_t *__restrict__ prefix) {
    __shared__ uint32_t s_warp[kScanBlock / 32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_chunks; base += kScanBlock) {
        const uint32_t c = base + threadIdx.x;
        uint32_t cnt = 0;
        if (c < n_chunks) {
            const uint32_t w0 = c * kChunkWords;
            if (w0 + kChunkWords <= n_words) {
                const uint4 *p = reinterpret_cast<const uint4 *>(bits + w0);
#pragma unroll
                for (int k = 0; k < (int)kChunkWords / 4; ++k) {
                    const uint4 v = __ldg(p + k);
                    cnt += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
                }
            } else {
                for (uint32_t w = w0; w < n_words; ++w) cnt += __popc(__ldg(bits + w));
            }
        }
        uint32_t incl = cnt;  // inclusive scan inside the warp, then across the warps
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31u) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= (uint32_t)o) w += t;
            }
            s_warp[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t before = carry + (warp ? s_warp[warp - 1] : 0u) + incl - cnt;
        if (c < n_chunks) prefix[c] = before;
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1) s_carry = carry + s_warp[kScanBlock / 32 - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) prefix[n_chunks] = s_carry;
}

// One thread per drawn cell.  mode 1 (update_all, :1166-1169): cell i of the trainable cells.  mode 0 (:1170-1191):
// the first n_first draws are uniform among the trainable cells (jran.choice, replace=True), the next n_second uniform
// among the OCCUPIED cells (jran.choice with p = occ_mask: the k-th occupied cell, k = ceil(total * (1 - u)) as
// jax's inverse-CDF search does; an empty grid yields the first trainable cell like searchsorted on an all-zero CDF).
// The fourth..sixth uniforms of the same Philox block are the jitter inside the cell.
// kSmemPrefix: `prefix` holds per-chunk COUNTS (ogrid_chunk_count_kernel) and every block scans them into its own
// shared-memory prefix first (<= kMaxSmemChunks chunks: grids up to 256^3); otherwise `prefix` is the global exclusive
// prefix of ogrid_chunk_prefix_kernel.
constexpr uint32_t kMaxSmemChunks = 16384;
template <bool kSmemPrefix>
__global__ void __launch_bounds__(kBlock) ogrid_draw_cells_kernel(NgpOgridDrawDescriptor d, const uint32_t *__restrict__ bits,
                                                                   const uint32_t *__restrict__ alive,
                                                                   const uint32_t *__restrict__ prefix,
                                                                   uint32_t *__restrict__ rng_state,
                                                                   uint32_t *__restrict__ idx_out, float *__restrict__ coords) {
    extern __shared__ uint32_t s_prefix[];  // [n_chunks + 1] when kSmemPrefix
    const uint32_t n_chunks = (d.n_cells + kChunkCells - 1u) / kChunkCells;
    const uint32_t *pre = prefix;
    if (kSmemPrefix) {  // block-wide exclusive scan of the chunk counts: kBlock threads x `per` consecutive chunks each
        __shared__ uint32_t s_warp[kBlock / 32];
        const uint32_t per = (n_chunks + kBlock - 1u) / kBlock, c0 = threadIdx.x * per;
        uint32_t sum = 0;
        for (uint32_t k = 0; k < per; ++k)
            if (c0 + k < n_chunks) sum += __ldg(prefix + c0 + k);
        uint32_t incl = sum;
        const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        if (lane == 31u) s_warp[warp] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (uint32_t w = 0; w < warp; ++w) base += s_warp[w];
        uint32_t run = base + incl - sum;
        for (uint32_t k = 0; k < per; ++k)
            if (c0 + k < n_chunks) {
                s_prefix[c0 + k] = run;
                run += __ldg(prefix + c0 + k);
            }
        if (threadIdx.x == kBlock - 1u) s_prefix[n_chunks] = base + incl;
        __syncthreads();
        pre = s_prefix;
    }
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    const uint32_t counter = rng_state[0];
    const uint32_t n_draws = d.mode ? d.n_alive : d.n_first + d.n_second;
    if (i < n_draws) {
        const Philox4 r = philox4x32_10(i, counter, d.rng.stream_id, 0u, d.rng.seed_lo, d.rng.seed_hi);
        uint32_t cell;
        if (d.mode || i < d.n_first) {
            const uint32_t k = d.mode ? i : __umulhi(r.x, d.n_alive);  // uniform in [0, n_alive)
            cell = alive ? __ldg(alive + k) : k;
        } else {
            const uint32_t total = pre[n_chunks];
            if (total == 0u) {
                cell = alive ? __ldg(alive) : 0u;
            } else {
                const float u = bits_to_unit_float(r.x);
                uint32_t k = (uint32_t)ceilf(__fmul_rn((float)total, __fadd_rn(1.f, -u)));  // in [1, total] up to rounding
                k = min(max(k, 1u), total) - 1u;                                              // 0-based rank
                uint32_t lo = 0, hi = n_chunks;  // last chunk with prefix[c] <= k
                while (hi - lo > 1u) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (pre[mid] <= k) lo = mid; else hi = mid;
                }
                uint32_t rest = k - pre[lo];
                const uint32_t w0 = lo * kChunkWords, n_words = (d.n_cells + 31u) / 32u;
                cell = 0u;
                if (w0 + kChunkWords <= n_words) {  // the chunk's 32 words as 8 independent 16-byte loads, then registers only
                    uint4 q[kChunkWords / 4];
#pragma unroll
                    for (int j = 0; j < (int)kChunkWords / 4; ++j) q[j] = __ldg(reinterpret_cast<const uint4 *>(bits + w0) + j);
                    bool found = false;
#pragma unroll
                    for (int j = 0; j < (int)kChunkWords / 4; ++j) {
                        const uint32_t ws[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const uint32_t pc = __popc(ws[e]);
                            if (!found && rest < pc) {
                                cell = (w0 + 4u * j + e) * 32u + __fns(ws[e], 0u, (int)rest + 1);
                                found = true;
                            }
                            if (!found) rest -= pc;
                        }
                    }
                } else {
                    for (uint32_t w = w0; w < n_words; ++w) {
                        const uint32_t word = __ldg(bits + w), pc = __popc(word);
                        if (rest < pc) {
                            cell = w * 32u + __fns(word, 0u, (int)rest + 1);
                            break;
                        }
                        rest -= pc;
                    }
                }
            }
        }
        const float u3[3] = {bits_to_unit_float(r.y), bits_to_unit_float(r.z), bits_to_unit_float(r.w)};
        float xyz[3];
        ogrid_cell_position(cell, d.G, d.mip_bound, u3, xyz);
        idx_out[i] = cell;
        coords[3 * (size_t)i + 0] = xyz[0];
        coords[3 * (size_t)i + 1] = xyz[1];
        coords[3 * (size_t)i + 2] = xyz[2];
    }
    rng_state_finish(rng_state);
}

}  // namespace
}  // namespace ngp

extern "C" {

void ngp_ogrid_sample_positions(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpOgridSampleDescriptor>(opaque, opaque_len, "ogrid_sample_positions");
    if (!d || d->n_points == 0) return;
    if (d->G < 2 || d->G > 1024) {
        set_error(NGP_ERR_ARGUMENT, "ogrid_sample_positions: expected 2 <= G <= 1024, got %u", d->G);
        return;
    }
    BufferCursor b{buffers};
    const uint32_t *idx = b.next<const uint32_t>();
    const float *uniforms = b.next<const float>();
    float *coords = b.next<float>();
    ogrid_sample_positions_kernel<<<div_up(d->n_points, kBlock), kBlock, 0, stream>>>(*d, idx, uniforms, coords);
    check_launch("ogrid_sample_positions");
}

void ngp_ogrid_decay_max(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpOgridUpdateDescriptor>(opaque, opaque_len, "ogrid_decay_max");
    if (!d) return;
    BufferCursor b{buffers};
    const float *grid_in = b.next<const float>();
    const uint32_t *idx = b.next<const uint32_t>();
    const float *vals = b.next<const float>();
    float *grid_out = b.next<float>();
    if (d->n_cells) {
        ogrid_decay_kernel<<<div_up(d->n_cells, kBlock), kBlock, 0, stream>>>(d->n_cells, d->decay, grid_in, grid_out);
        if (!check_launch("ogrid_decay_max(decay)")) return;
    }
    if (d->n_updates) {
        ogrid_scatter_max_kernel<<<div_up(d->n_updates, kBlock), kBlock, 0, stream>>>(d->n_updates, d->n_cells, idx, vals, grid_out);
        check_launch("ogrid_decay_max(scatter)");
    }
}

void ngp_ogrid_draw_cells(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpOgridDrawDescriptor>(opaque, opaque_len, "ogrid_draw_cells");
    if (!d) return;
    if (d->G < 2 || d->G > 1024 || d->n_cells == 0 || d->n_cells % 32 != 0 || d->n_alive == 0 || d->n_alive > d->n_cells) {
        set_error(NGP_ERR_ARGUMENT, "ogrid_draw_cells: expected 2 <= G <= 1024, n_cells a positive multiple of 32 and 0 < n_alive <= n_cells, "
                  "got G=%u n_cells=%u n_alive=%u", d->G, d->n_cells, d->n_alive);
        return;
    }
    BufferCursor b{buffers};
    const uint32_t *bits = b.next<const uint32_t>();
    const uint32_t *alive = b.next<const uint32_t>();  // may be null: every cell is trainable
    uint32_t *rng_state = b.next<uint32_t>();
    uint32_t *idx = b.next<uint32_t>();
    float *coords = b.next<float>();
    if (d->has_alive != (alive != nullptr)) {
        set_error(NGP_ERR_ARGUMENT, "ogrid_draw_cells: has_alive=%u but the alive buffer is %s", d->has_alive, alive ? "given" : "null");
        return;
    }
    const uint32_t n_draws = d->mode ? d->n_alive : d->n_first + d->n_second;
    if (n_draws == 0) return;
    uint32_t *prefix = nullptr;
    bool smem_prefix = false;
    const uint32_t n_chunks = div_up(d->n_cells, kChunkCells);
    if (!d->mode && d->n_second) {
        if (reinterpret_cast<uintptr_t>(bits) % 16 != 0) {
            set_error(NGP_ERR_ARGUMENT, "ogrid_draw_cells: the bitfield must be 16-byte aligned");
            return;
        }
        prefix = static_cast<uint32_t *>(workspace(stream, (size_t)(n_chunks + 1) * sizeof(uint32_t)));
        if (!prefix) return;
        smem_prefix = n_chunks <= kMaxSmemChunks;
        if (smem_prefix) ogrid_chunk_count_kernel<<<div_up(n_chunks, kBlock), kBlock, 0, stream>>>(n_chunks, d->n_cells / 32u, bits, prefix);
        else ogrid_chunk_prefix_kernel<<<1, kScanBlock, 0, stream>>>(n_chunks, d->n_cells / 32u, bits, prefix);
        if (!check_launch("ogrid_draw_cells(prefix)")) return;
    }
    if (smem_prefix) {
        const size_t smem = (size_t)(n_chunks + 1) * sizeof(uint32_t);
        static bool configured = false;  // benign race: idempotent
        if (!configured) {
            cudaFuncSetAttribute(ogrid_draw_cells_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((kMaxSmemChunks + 1) * sizeof(uint32_t)));
            configured = true;
        }
        ogrid_draw_cells_kernel<true><<<div_up(n_draws, kBlock), kBlock, smem, stream>>>(*d, bits, alive, prefix, rng_state, idx, coords);
    } else {
        ogrid_draw_cells_kernel<false><<<div_up(n_draws, kBlock), kBlock, 0, stream>>>(*d, bits, alive, prefix, rng_state, idx, coords);
    }
    check_launch("ogrid_draw_cells");
}

void ngp_ogrid_threshold(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpOgridThresholdDescriptor>(opaque, opaque_len, "ogrid_threshold");
    if (!d) return;
    BufferCursor b{buffers};
    const float *grid = b.next<const float>();
    float *thr = b.next<float>();
    const unsigned blocks = max(1u, min(div_up(d->n_cells, kBlock * kThrPerThread), 148u * 8u));
    const size_t ws_bytes = 16 + (size_t)blocks * 16;
    auto *ws = static_cast<char *>(workspace(stream, ws_bytes));
    if (!ws) return;
    // the `done` counter re-arms itself; zero it once per call anyway (cheap, keeps first use simple)
    NGP_CUDA_OK(cudaMemsetAsync(ws, 0, 16, stream), "ogrid_threshold");
    ogrid_threshold_kernel<<<blocks, kBlock, 0, stream>>>(d->n_cells, d->thr_max, grid, reinterpret_cast<double *>(ws + 16),
                                                           reinterpret_cast<unsigned long long *>(ws + 16 + (size_t)blocks * 8),
                                                           reinterpret_cast<uint32_t *>(ws), thr);
    check_launch("ogrid_threshold");
}

}  // extern "C"
