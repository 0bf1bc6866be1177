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

__global__ void __launch_bounds__(kBlock) ogrid_sample_positions_kernel(NgpOgridSampleDescriptor d,
                                                                         const uint32_t *__restrict__ idx,
                                                                         const float *__restrict__ uniforms,
                                                                         float *__restrict__ coords) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= d.n_points) return;
    const uint32_t m = __ldg(idx + i);
    const float half_cell = __fdiv_rn(d.mip_bound, (float)d.G);         // :1196
    const float span = __fadd_rn(d.mip_bound, -half_cell);              // :1197
    const float inv = (float)(d.G - 1u);
    const uint32_t c[3] = {compact_bits10(m), compact_bits10(m >> 1), compact_bits10(m >> 2)};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float x = __fadd_rn(__fmul_rn(__fdiv_rn((float)c[k], inv), 2.f), -1.f);  // :1194
        x = __fmul_rn(x, span);
        // jax.random.uniform(minval=-h, maxval=h): max(minval, u * (maxval - minval) + minval)
        const float jitter = fmaxf(-half_cell, __fadd_rn(__fmul_rn(__ldg(uniforms + 3 * (size_t)i + k), __fadd_rn(half_cell, half_cell)), -half_cell));
        coords[3 * (size_t)i + k] = __fadd_rn(x, jitter);
    }
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
    // non-negative floats order like their bit patterns; the cell holds a non-negative value too
    atomicMax(reinterpret_cast<int *>(grid) + c, __float_as_int(v));
}

// mean over alive (>= 0) cells + threshold; two-level deterministic reduction, last block finalises
__global__ void __launch_bounds__(kBlock) ogrid_threshold_kernel(uint32_t n, float thr_max, const float *__restrict__ grid,
                                                                  double *__restrict__ partial_sum,
                                                                  unsigned long long *__restrict__ partial_cnt,
                                                                  uint32_t *__restrict__ done, float *__restrict__ thr_out) {
    __shared__ double s_sum[kBlock / 32];
    __shared__ unsigned long long s_cnt[kBlock / 32];
    __shared__ bool s_last;
    double sum = 0.0;
    unsigned long long cnt = 0;
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
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
    if (s_last && threadIdx.x == 0) {
        double ts = 0.0;
        unsigned long long tc = 0;
        for (uint32_t b = 0; b < gridDim.x; ++b) {  // fixed order: deterministic
            ts += reinterpret_cast<volatile double *>(partial_sum)[b];
            tc += reinterpret_cast<volatile unsigned long long *>(partial_cnt)[b];
        }
        const float mean = tc ? (float)(ts / (double)tc) : 0.f;
        *thr_out = fminf(thr_max, mean);
        *done = 0u;
    }
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

void ngp_ogrid_threshold(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpOgridThresholdDescriptor>(opaque, opaque_len, "ogrid_threshold");
    if (!d) return;
    BufferCursor b{buffers};
    const float *grid = b.next<const float>();
    float *thr = b.next<float>();
    const unsigned blocks = max(1u, min(div_up(d->n_cells, kBlock * 8), 148u * 2u));
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
