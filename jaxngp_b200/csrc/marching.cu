// marching.cu -- occupancy-grid ray marching for sm_100a.
//
// Replaces deps/volume-rendering-jax/lib/impl/marching.cu:101-397,435-604.
//
// Numerics: every floating-point operation that decides where a sample lands is pinned with
// round-to-nearest intrinsics in exactly the shape nvcc gives the reference source (FFMA for
// o + t*d, for t_start + ds*noise and for the two contractions inside the next-voxel distance;
// IEEE division for 1/d, 1/G, pos/mip_bound and the ds lower bound) -- read off the reference SASS,
// see DESIGN.md section 4 -- so sample counts, grid indices, positions and z values are bit-equal
// to the reference kernels'.
//
// Structure (training march): one persistent kernel, one warp per ray.  A CTA takes a tile of 8
// consecutive rays (dynamic ticket), counts each ray's occupied steps, and obtains the tile's
// exclusive prefix by decoupled look-back, so sample ranges are handed out in ray order without a
// second launch or a host round trip.  The first tile whose inclusive prefix reaches `total_samples`
// publishes a cut; tiles behind the cut stop marching (they poll the cut flag) -- this is the
// reference's "budget already full" early-out (marching.cu:135) made deterministic.  Rays that got a
// range then re-march and write their samples; a tail kernel zero-fills the unused sample slots
// (observable API, marching/__init__.py:60-68).  The reference zero-fills everything with 10
// memsets before its kernel (marching.cu:474-483).
#include "common.cuh"
#include "marching.cuh"

namespace ngp {
namespace {

using namespace march;

// ---------------------------------------------------------------- training march
//
// One WARP per ray.  The parameter sequence a ray can visit is fixed by its start:
// t_{k+1} = t_k + ds(t_k) (the same rounded additions whether the reference takes a sample step or
// walks to the next voxel, marching.cu:178-189), and the reference visits a subsequence of it:
// after an occupied point the next one, after an empty point the first t_j >= next_t.  So a warp
// evaluates 32 consecutive chain points at once (32 independent bitfield lookups in flight instead
// of a 600-deep dependent chain per thread), then replays the visit rule over the 32 results with
// ballots -- runs of occupied points are consumed in one go, skips with one shuffle + one ballot.
// Same visited set, same samples, bit for bit; ~30x shorter critical path per ray.
constexpr int kRaysPerTile = 8;   // warps per CTA, one ray each
constexpr int kMarchBlock = kRaysPerTile * 32;
constexpr uint32_t kDrainBatch = 64;  // tiles claimed per ticket once the budget is known to be full
constexpr uint64_t kFlagAgg = 1ull << 62, kFlagIncl = 2ull << 62, kValueMask = (1ull << 62) - 1;

struct MarchScratch {
    uint32_t ticket;
    uint32_t cut_inv;  // 0 = no cut yet; else 0xFFFFFFFF - (index of a tile whose inclusive prefix >= total_samples)
    unsigned long long status[1];  // [num_tiles]
};

__device__ __forceinline__ bool behind_cut(const MarchScratch *ws, uint32_t tile) {
    uint32_t v = *reinterpret_cast<const volatile uint32_t *>(&ws->cut_inv);
    return v != 0 && (0xFFFFFFFFu - v) < tile;
}

struct SampleSink {
    uint32_t *idcs;
    float *xyzs, *dirs, *dss, *z_vals;
};

// Marches one ray with the whole warp.  `limit` = most samples this ray may emit (count pass: the
// per-ray cap diagonal_n_steps*bound, marching.cu:165; write pass: the count found in pass 1).
// Returns the number of samples; in the count pass returns 0xFFFFFFFF if the tile fell behind the cut.
template <bool kWrite>
__device__ __forceinline__ uint32_t march_ray_warp(const Grid &g, const Ray &r, float t0, float t_end, uint32_t limit,
                                                   uint32_t ray_index, const SampleSink &out,
                                                   const MarchScratch *ws, uint32_t tile) {
    const uint32_t lane = threadIdx.x & 31u;
    const bool const_ds = g.portion == 0.f;  // ds(t) == ds_lo clamped: no dependence on t
    const float ds0 = calc_ds(g, 0.f);
    uint32_t n = 0;
    float t_base = t0;
    bool pending = false;
    float pend_t = 0.f;
    while (n < limit) {
        // chain: lane j holds t_base advanced j times
        float t_next_base;
        const float t = chain_points(g, t_base, lane, const_ds, ds0, t_next_base);
        const bool in = t < t_end;
        const uint32_t V = __ballot_sync(0xffffffffu, in);
        if (V == 0u) break;
        uint32_t v = 0;
        bool skip_eval = false;
        if (pending) {  // still walking towards a voxel boundary found in an earlier chunk
            const uint32_t m = __ballot_sync(0xffffffffu, t >= pend_t);
            if (m == 0u) { v = 32; skip_eval = true; }
            else { v = __ffs(m) - 1; pending = false; }
        }
        if (!skip_eval) {
            const EvalPoint s = eval_point(g, r, t);
            const uint32_t O = __ballot_sync(0xffffffffu, s.occupied && in);
            bool done = false;
            while (v < 32u) {
                if (!((V >> v) & 1u) || n >= limit) { done = true; break; }
                if ((O >> v) & 1u) {
                    const uint32_t rest = O >> v;  // consecutive occupied points starting at v
                    uint32_t run = (rest == (0xFFFFFFFFu >> v)) ? 32u - v : (uint32_t)__ffs(~rest) - 1u;
                    run = min(run, limit - n);
                    if (kWrite && lane >= v && lane < v + run) {
                        const uint32_t w = n + (lane - v);
                        out.idcs[w] = ray_index;
                        out.xyzs[w * 3 + 0] = s.px;
                        out.xyzs[w * 3 + 1] = s.py;
                        out.xyzs[w * 3 + 2] = s.pz;
                        out.dirs[w * 3 + 0] = r.dx;
                        out.dirs[w * 3 + 1] = r.dy;
                        out.dirs[w * 3 + 2] = r.dz;
                        out.dss[w] = s.ds;
                        out.z_vals[w] = t;
                    }
                    n += run;
                    v += run;
                } else {
                    const float nt = __shfl_sync(0xffffffffu, s.next_t, v);
                    const uint32_t above = (v == 31u) ? 0u : (0xFFFFFFFFu << (v + 1u));
                    const uint32_t m = __ballot_sync(0xffffffffu, t >= nt) & above;
                    if (m == 0u) { pending = true; pend_t = nt; v = 32; }
                    else v = __ffs(m) - 1;
                }
            }
            if (done) break;
        }
        t_base = t_next_base;
        if (!kWrite && behind_cut(ws, tile)) return 0xFFFFFFFFu;
    }
    return n;
}

__global__ void __launch_bounds__(kMarchBlock) march_rays_kernel(
    NgpMarchingDescriptor p, MarchScratch *__restrict__ ws, uint32_t num_tiles,
    const float *__restrict__ rays_o, const float *__restrict__ rays_d, const float *__restrict__ t_starts,
    const float *__restrict__ t_ends, const float *__restrict__ noises, const uint8_t *__restrict__ bitfield,
    uint32_t *__restrict__ next_sample_write_location, uint32_t *__restrict__ number_of_exceeded_samples,
    uint8_t *__restrict__ ray_is_valid, uint32_t *__restrict__ rays_n_samples,
    uint32_t *__restrict__ rays_sample_startidx, uint32_t *__restrict__ idcs, float *__restrict__ xyzs,
    float *__restrict__ dirs, float *__restrict__ dss, float *__restrict__ z_vals) {
    __shared__ uint32_t s_tile, s_drain_end;
    __shared__ uint32_t s_count[kRaysPerTile];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_abandon, s_drain;

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const Grid g = make_grid(p.diagonal_n_steps, p.K, p.G, p.bound, p.stepsize_portion, bitfield);
    // per-ray cap: `(float)n < diagonal_n_steps * bound` (marching.cu:165) <=> n < ceil(that product)
    const float max_steps_f = __fmul_rn((float)p.diagonal_n_steps, p.bound);
    const uint32_t cap = max_steps_f > 0.f ? (uint32_t)fminf(ceilf(max_steps_f), 4294967040.f) : 0u;

    for (;;) {  // persistent: tiles are taken in ticket order, so earlier rays are never starved by later ones
        __syncthreads();
        if (threadIdx.x == 0) {
            s_tile = atomicAdd(&ws->ticket, 1u);
            s_drain_end = s_tile + 1u;
            s_abandon = 0;
            s_drain = behind_cut(ws, s_tile) ? 1 : 0;  // decided once per CTA: the branch below must be uniform
        }
        __syncthreads();
        uint32_t tile = s_tile;
        if (tile >= num_tiles) return;
        if (s_drain) {
            // Drain: every remaining tile lies behind the cut (tickets only grow), so its rays early-out
            // (marching.cu:135).  Claim tiles 64 at a time and clear their per-ray outputs with full-width stores.
            for (;;) {
                const uint32_t first_ray = tile * kRaysPerTile;
                const uint32_t end_ray = min(p.n_rays, (s_drain_end) * kRaysPerTile);
                for (uint32_t r = first_ray + threadIdx.x; r < end_ray; r += kMarchBlock) {
                    ray_is_valid[r] = 0;
                    rays_n_samples[r] = 0u;
                    rays_sample_startidx[r] = 0u;
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    s_tile = atomicAdd(&ws->ticket, kDrainBatch);
                    s_drain_end = s_tile + kDrainBatch;
                }
                __syncthreads();
                tile = s_tile;
                if (tile >= num_tiles) return;
            }
        }
        const uint32_t i = tile * kRaysPerTile + warp;
        const bool in_range = i < p.n_rays;

        // ---- pass 1: count occupied steps (warp-cooperative)
        Ray ray = {};
        float t_end = 0.f, t0 = 0.f;
        uint32_t n = 0;
        bool hit_box = false;
        if (in_range) {
            const float t_start = __ldg(t_starts + i);
            t_end = __ldg(t_ends + i);
            hit_box = t_end > t_start;  // marching.cu:151
            if (hit_box) {
                ray = load_ray(rays_o, rays_d, i);
                t0 = __fmaf_rn(calc_ds(g, t_start), __ldg(noises + i), t_start);  // marching.cu:164
            }
        }
        bool abandoned = behind_cut(ws, tile);
        if (hit_box && !abandoned) {
            n = march_ray_warp<false>(g, ray, t0, t_end, cap, i, SampleSink{}, ws, tile);
            if (n == 0xFFFFFFFFu) { abandoned = true; n = 0; }
        }
        if (lane == 0) {
            s_count[warp] = n;
            if (abandoned) s_abandon = 1;
        }
        __syncthreads();
        uint32_t excl_in_tile = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < kRaysPerTile; ++w) {
            const uint32_t c = s_count[w];
            if (w < (int)warp) excl_in_tile += c;
            tile_total += c;
        }

        // ---- decoupled look-back for the tile's exclusive prefix: warp 0 inspects 32 predecessors per round
        if (warp == 0) {
            volatile unsigned long long *status = ws->status;
            unsigned long long prefix = 0;
            bool give_up = s_abandon != 0;
            if (!give_up) {
                if (lane == 0) status[tile] = (tile == 0 ? kFlagIncl : kFlagAgg) | tile_total;
                __threadfence();
                int64_t hi = (int64_t)tile - 1;  // newest predecessor not yet accounted for
                while (hi >= 0) {
                    const int64_t j = hi - (int64_t)lane;
                    unsigned long long st = 0;
                    bool ready;
                    do {  // wait until all 32 (existing) predecessors of this round have published something
                        st = j >= 0 ? status[j] : kFlagIncl;
                        ready = __all_sync(0xffffffffu, st != 0);
                        if (!ready && behind_cut(ws, tile)) { give_up = true; break; }
                    } while (!ready);
                    if (give_up) break;
                    // nearest predecessor holding an inclusive prefix ends the walk
                    const uint32_t incl_mask = __ballot_sync(0xffffffffu, (st & kFlagIncl) != 0 && j >= 0);
                    const uint32_t stop_lane = incl_mask ? (uint32_t)__ffs(incl_mask) - 1u : 32u;
                    unsigned long long v = (j >= 0 && lane <= stop_lane) ? (st & kValueMask) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    prefix += v;
                    if (incl_mask) break;
                    hi -= 32;
                }
                if (!give_up && lane == 0) {
                    if (tile != 0) status[tile] = kFlagIncl | (prefix + tile_total);
                    __threadfence();
                    if (prefix + tile_total >= p.total_samples) atomicMax(&ws->cut_inv, 0xFFFFFFFFu - tile);
                    if (tile == num_tiles - 1 && prefix + tile_total < p.total_samples) {
                        *next_sample_write_location = (uint32_t)(prefix + tile_total);  // no cut anywhere
                        *number_of_exceeded_samples = 0u;
                    }
                }
            }
            if (lane == 0) {
                if (give_up) s_abandon = 1;
                s_prefix = prefix;
            }
        }
        __syncthreads();
        if (!in_range) continue;

        const bool tile_abandoned = s_abandon != 0;
        const unsigned long long start64 = s_prefix + excl_in_tile;
        bool valid = false;
        uint32_t n_out = 0, start_out = 0;
        if (!tile_abandoned && start64 < p.total_samples && hit_box) {  // else: marching.cu:135 / :151
            const uint32_t start = (uint32_t)start64;
            if (n == 0) {
                valid = true;  // marching.cu:196-199
            } else {
                if (start64 + n >= p.total_samples && lane == 0) {  // the ray at which the budget fills
                    *next_sample_write_location = start + n;
                    *number_of_exceeded_samples = (start64 + n > p.total_samples) ? n : 0u;
                }
                if (start64 + n <= p.total_samples) {
                    valid = true;
                    n_out = n;
                    start_out = start;
                }
            }
        }
        if (lane == 0) {
            ray_is_valid[i] = valid ? 1 : 0;
            rays_n_samples[i] = n_out;
            rays_sample_startidx[i] = start_out;
        }
        // ---- pass 2: march again and write (marching.cu:224-267)
        if (n_out) {
            SampleSink sink{idcs + start_out, xyzs + (size_t)start_out * 3, dirs + (size_t)start_out * 3,
                            dss + start_out, z_vals + start_out};
            march_ray_warp<true>(g, ray, t0, t_end, n_out, i, sink, ws, tile);
        }
    }
}

// zero the sample slots nobody wrote: [next - exceeded, total_samples)
__global__ void __launch_bounds__(256) march_rays_tail_kernel(
    uint32_t total_samples, const uint32_t *__restrict__ next_loc, const uint32_t *__restrict__ exceeded,
    uint32_t *__restrict__ idcs, float *__restrict__ xyzs, float *__restrict__ dirs, float *__restrict__ dss,
    float *__restrict__ z_vals) {
    const uint32_t used = min(*next_loc - *exceeded, total_samples);
    for (uint32_t s = used + blockIdx.x * blockDim.x + threadIdx.x; s < total_samples; s += gridDim.x * blockDim.x) {
        idcs[s] = 0u;
        dss[s] = 0.f;
        z_vals[s] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            xyzs[(size_t)s * 3 + k] = 0.f;
            dirs[(size_t)s * 3 + k] = 0.f;
        }
    }
}

// ---------------------------------------------------------------- inference march
//
// Two launches.  (1) rank kernel: fresh rays are handed to terminated slots in SLOT ORDER (the
// reference uses atomicAdd arrival order, marching.cu:300): every 1024-slot block scans its flags
// and writes each slot's rank inside the block plus the block total.  (2) march kernel: one WARP
// per slot.  Like the training march it evaluates 32 consecutive points of the fixed parameter
// chain at once and replays the reference's visit rule over the results (bit-equal sample
// positions, counts and resume parameter t), instead of one thread walking a dependent chain of
// bitfield loads per slot: with 10^5 slots in flight the kernel is issue-bound, not latency-bound.
constexpr int kRankBlock = 256, kRankSlots = 1024;  // 4 flags per thread
constexpr int kInferWarps = 8;

__global__ void __launch_bounds__(kRankBlock) march_rays_inference_rank_kernel(
    uint32_t n_rays, const uint8_t *__restrict__ terminated, uint32_t *__restrict__ block_total,
    uint32_t *__restrict__ rank_in_block, const uint32_t *__restrict__ counter_src, uint32_t *__restrict__ counter_snapshot) {
    // in-place variant: the march kernel overwrites the ray counter it also reads, so it reads this copy instead
    if (counter_snapshot && blockIdx.x == 0 && threadIdx.x == 0) *counter_snapshot = *counter_src;
    __shared__ uint32_t s_warp[kRankBlock / 32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t first = blockIdx.x * kRankSlots + threadIdx.x * 4u;
    uint32_t f[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) f[k] = (first + k < n_rays && terminated[first + k]) ? 1u : 0u;
    const uint32_t mine = f[0] + f[1] + f[2] + f[3];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kRankBlock / 32; ++w) {
        if (w < (int)warp) base += s_warp[w];
        total += s_warp[w];
    }
    uint32_t r = base + incl - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (first + k < n_rays) rank_in_block[first + k] = r;
        r += f[k];
    }
    if (threadIdx.x == 0) block_total[blockIdx.x] = total;
}

// kInPlace (renderer fast path, ngp_march_rays_inference_inplace): the scatter the reference does after the op
// (t_starts.at[indices].set(t_starts_out), marching/__init__.py:156) happens here -- every ray belongs to exactly one
// slot -- and the slot's ray direction is copied out for the MLP (cuda.py:222-228 gathers it with rays_d[indices]).
// (Latency-bound: five dependent loads per slot.  Capped at 32 registers for full occupancy -- measured 11 % faster
// than the 57-register build despite 48 bytes of spills.)
// W = lanes per slot: a whole warp (32) or half a warp (16: two slots per warp, each half with its own lane mask --
// at march_steps_cap <= 16 one 16-point chunk of the chain already covers a full pass through occupied space).
template <bool kInPlace, int W>
__global__ void __launch_bounds__(kInferWarps * 32, 8) march_rays_inference_kernel(
    NgpMarchingInferenceDescriptor p, const float *__restrict__ rays_o, const float *__restrict__ rays_d,
    const float *t_starts, const float *__restrict__ t_ends, const uint8_t *__restrict__ bitfield,
    const uint32_t *__restrict__ next_ray_index_in, const uint8_t *__restrict__ terminated,
    const uint32_t *indices_in, const uint32_t *__restrict__ block_total,
    const uint32_t *__restrict__ rank_in_block, uint32_t *__restrict__ next_ray_index, uint32_t *indices_out,
    uint32_t *__restrict__ n_samples, float *t_starts_out, float *__restrict__ xyzs,
    float *__restrict__ dss, float *__restrict__ z_vals, float *__restrict__ ray_dirs,
    const float *__restrict__ t_admitted) {
    constexpr uint32_t kW = W, kWMask = W == 32 ? 0xFFFFFFFFu : 0xFFFFu;
    const uint32_t lane = threadIdx.x & (kW - 1u);
    const uint32_t shift = W == 32 ? 0u : (threadIdx.x & 16u);  // position of this slot's lanes in the warp
    const uint32_t mask = kWMask << shift;
    const uint32_t i = blockIdx.x * (kInferWarps * 32u / kW) + threadIdx.x / kW;  // slot of this lane group
    if (i >= p.n_rays) return;
    auto group_sum = [&](uint32_t v) {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, W);
        return v;
    };
    auto ballot = [&](bool pred) { return (__ballot_sync(mask, pred) >> shift) & kWMask; };
    const uint32_t counter_in = __ldg(next_ray_index_in);
    const bool mine = terminated[i] != 0;
    uint32_t ray_idx;
    if (t_admitted) {  // admission and the empty-space prefix were done by march_rays_inference_admit_kernel
        ray_idx = indices_out[i];
    } else if (mine || i == p.n_rays - 1) {  // terminated slots before this slot's 1024-block
        const uint32_t nb = mine ? i / kRankSlots : 0u, nb_all = div_up_dev(p.n_rays, kRankSlots);
        uint32_t before = 0, all = 0;
        for (uint32_t j = lane; j < nb_all; j += kW) {
            const uint32_t c = (j < nb || i == p.n_rays - 1) ? __ldg(block_total + j) : 0u;
            if (j < nb) before += c;
            all += c;
        }
        before = group_sum(before);
        if (i == p.n_rays - 1) {
            all = group_sum(all);
            if (lane == 0) *next_ray_index = counter_in + all;
        }
        ray_idx = mine ? counter_in + before + __ldg(rank_in_block + i) : indices_in[i];
    } else {
        ray_idx = indices_in[i];
    }
    if (lane == 0 && !t_admitted) indices_out[i] = ray_idx;

    const uint32_t cap = p.march_steps_cap;
    float *__restrict__ o_xyzs = xyzs + (size_t)i * cap * 3;
    float *__restrict__ o_dss = dss + (size_t)i * cap;
    float *__restrict__ o_z = z_vals + (size_t)i * cap;

    uint32_t steps = 0;
    float t_cur = 0.f, t_end = 0.f;
    bool live = ray_idx < p.n_total_rays;
    if (live) {
        t_cur = t_admitted ? __ldg(t_admitted + i) : t_starts[ray_idx];  // plain load: the in-place variant writes this element back below
        t_end = __ldg(t_ends + ray_idx);
        live = !(t_end < t_cur);  // marching.cu:317 (strict)
    }
    if (live) {
        const Grid g = make_grid(p.diagonal_n_steps, p.K, p.G, p.bound, p.stepsize_portion, bitfield);
        const Ray ray = load_ray(rays_o, rays_d, ray_idx);
        steps = march_chunk<W>(
            g, ray, t_cur, t_end, cap, lane, mask, shift,
            [&](uint32_t w, const EvalPoint &s, float t) {
                o_xyzs[w * 3 + 0] = s.px;
                o_xyzs[w * 3 + 1] = s.py;
                o_xyzs[w * 3 + 2] = s.pz;
                o_dss[w] = s.ds;
                o_z[w] = t;
            },
            [&](uint32_t w, float ds) { o_dss[w] = ds; });
    }
    if (lane == 0) {
        n_samples[i] = steps;
        const float t_store = live ? t_cur : 0.f;  // the reference leaves the memset zeros for rays it skips
        if (!kInPlace) t_starts_out[i] = t_store;
        else if (ray_idx < p.n_total_rays) t_starts_out[ray_idx] = t_store;  // out-of-range indices are dropped
    }
    if (kInPlace && lane < 3) ray_dirs[3 * (size_t)i + lane] = ray_idx < p.n_total_rays ? __ldg(rays_d + 3 * (size_t)ray_idx + lane) : 0.f;
    // zero the unused tail (reference: memsets, marching.cu:565-570).  The in-place variant feeds consumers that only
    // read the first n_samples rows of a slot (grouped encoder/MLP, integrate_rays_inference), so it skips the fill:
    // 320 bytes per idle slot per pass, most of this kernel's traffic late in a frame.
    if (!kInPlace) {
        __syncwarp(mask);
        for (uint32_t k = steps * 3 + lane; k < cap * 3; k += kW) o_xyzs[k] = 0.f;
        for (uint32_t k = steps + lane; k < cap; k += kW) {
            o_dss[k] = 0.f;
            o_z[k] = 0.f;
        }
    }
}

// ---------------------------------------------------------------- admission + empty-space prefix, one thread per slot
// Front end of the drop-in march_rays_inference: hands fresh rays to terminated slots in slot order (same rule as the
// cooperative kernel) and runs the empty-space prefix of the slot's ray (march_rays_skip_empty_kernel's rule: the
// visit sequence touches ~1 chain point in 5 while nothing is occupied, which a thread hopping voxel to voxel walks
// far cheaper than 16 lanes evaluating consecutive chain points).  A frame's first pass starts every ray at the near
// plane: measured 1.55 ms -> see DESIGN for the cooperative kernel alone on 262,144 fresh slots against 0.49 ms for
// the reference's thread-per-slot loop.  The cooperative kernel then starts at t_admitted[slot]: same samples, same final t.
__global__ void __launch_bounds__(128) march_rays_inference_admit_kernel(
    NgpMarchingInferenceDescriptor p, const float *__restrict__ rays_o, const float *__restrict__ rays_d,
    const float *__restrict__ t_starts, const float *__restrict__ t_ends, const uint8_t *__restrict__ bitfield,
    const uint32_t *__restrict__ next_ray_index_in, const uint8_t *__restrict__ terminated,
    const uint32_t *__restrict__ indices_in, const uint32_t *__restrict__ block_total,
    const uint32_t *__restrict__ rank_in_block, uint32_t *__restrict__ next_ray_index, uint32_t *__restrict__ indices_out,
    float *__restrict__ t_admitted) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t counter_in = __ldg(next_ray_index_in);
    // a warp's 32 slots lie in one 1024-slot rank block: the terminated slots before that block, summed cooperatively
    const uint32_t warp_first = i - lane;
    if (warp_first >= p.n_rays) return;
    const uint32_t nb = warp_first / kRankSlots, nb_all = div_up_dev(p.n_rays, kRankSlots);
    const bool last_warp = warp_first + 32u >= p.n_rays;
    uint32_t before = 0, all = 0;
    for (uint32_t j = lane; j < (last_warp ? nb_all : nb); j += 32u) {
        const uint32_t c = __ldg(block_total + j);
        if (j < nb) before += c;
        all += c;
    }
    before = __reduce_add_sync(0xffffffffu, before);
    if (last_warp) {
        all = __reduce_add_sync(0xffffffffu, all);
        if (lane == 0) *next_ray_index = counter_in + all;
    }
    if (i >= p.n_rays) return;
    const uint32_t ray_idx = terminated[i] ? counter_in + before + __ldg(rank_in_block + i) : __ldg(indices_in + i);
    indices_out[i] = ray_idx;
    float t = 0.f;
    if (ray_idx < p.n_total_rays) {
        t = __ldg(t_starts + ray_idx);
        const float t_end = __ldg(t_ends + ray_idx);
        if (!(t_end < t)) {  // marching.cu:317
            const Grid g = make_grid(p.diagonal_n_steps, p.K, p.G, p.bound, p.stepsize_portion, bitfield);
            const Ray ray = load_ray(rays_o, rays_d, ray_idx);
            float t_prev = t;
            while (t < t_end) {
                const Step s = march_step<true>(g, ray, t);
                if (s.occupied) break;
                t_prev = t;
                t = s.t_next;
            }
            if (!(t < t_end)) t = t_prev;  // see march_rays_skip_empty_kernel
        }
    }
    t_admitted[i] = t;
}

// ---------------------------------------------------------------- empty-space pre-advance (inference)
// The state a ray carries between march_rays_inference calls is its parameter t alone
// (marching/__init__.py:156), and until the reference's loop (marching.cu:323-365) meets its first
// occupied point it only moves t along the visit sequence without emitting anything.  This kernel runs
// that prefix for every ray of a frame once, one THREAD per ray (the visit sequence touches ~1 chain
// point in 5, so a thread hopping voxel to voxel does a fraction of the warp-cooperative march's work):
// t_out = the first visited point that is occupied, or the last visited point < t_end if there is none.  Feeding
// t_out to march_rays_inference as t_starts yields bit-identical samples and final t.
__global__ void __launch_bounds__(128) march_rays_skip_empty_kernel(
    NgpMarchingInferenceDescriptor p, const float *__restrict__ rays_o, const float *__restrict__ rays_d,
    const float *t_starts, const float *__restrict__ t_ends, const uint8_t *__restrict__ bitfield, float *t_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_total_rays) return;
    float t = t_starts[i];
    const float t_end = __ldg(t_ends + i);
    if (!(t_end < t)) {  // marching.cu:317
        const Grid g = make_grid(p.diagonal_n_steps, p.K, p.G, p.bound, p.stepsize_portion, bitfield);
        const Ray ray = load_ray(rays_o, rays_d, i);
        float t_prev = t;
        while (t < t_end) {
            const Step s = march_step<true>(g, ray, t);
            if (s.occupied) break;
            t_prev = t;
            t = s.t_next;
        }
        // nothing occupied before t_end: hand back the LAST visited point still inside the ray, so the march
        // re-visits it, steps past t_end to the same final t and applies the far-plane rule (marching.cu:367-394)
        if (!(t < t_end)) t = t_prev;
    }
    t_out[i] = t;
}

}  // namespace
}  // namespace ngp

namespace ngp {
thread_local int g_march_ctas_per_sm = 0;  // 0 = default (4)
}

extern "C" {

void ngp_b200_set_march_ctas_per_sm(int ctas_per_sm) { ngp::g_march_ctas_per_sm = ctas_per_sm; }

void ngp_march_rays(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *desc = descriptor<NgpMarchingDescriptor>(opaque, opaque_len, "march_rays");
    if (!desc) return;
    if (desc->K == 0 || desc->G == 0 || desc->G > 1024) {
        set_error(NGP_ERR_ARGUMENT, "march_rays: expected K > 0 and 0 < G <= 1024, got K=%u G=%u", desc->K, desc->G);
        return;
    }
    BufferCursor b{buffers};
    const float *rays_o = b.next<const float>();
    const float *rays_d = b.next<const float>();
    const float *t_starts = b.next<const float>();
    const float *t_ends = b.next<const float>();
    const float *noises = b.next<const float>();
    const uint8_t *bitfield = b.next<const uint8_t>();
    uint32_t *next_loc = b.next<uint32_t>();
    uint32_t *exceeded = b.next<uint32_t>();
    uint8_t *valid = b.next<uint8_t>();
    uint32_t *rays_n = b.next<uint32_t>();
    uint32_t *rays_start = b.next<uint32_t>();
    uint32_t *idcs = b.next<uint32_t>();
    float *xyzs = b.next<float>();
    float *dirs = b.next<float>();
    float *dss = b.next<float>();
    float *z_vals = b.next<float>();

    const uint32_t num_tiles = div_up(desc->n_rays, kRaysPerTile);
    if (num_tiles == 0) {
        NGP_CUDA_OK(cudaMemsetAsync(next_loc, 0, sizeof(uint32_t), stream), "march_rays");
        NGP_CUDA_OK(cudaMemsetAsync(exceeded, 0, sizeof(uint32_t), stream), "march_rays");
    } else {
        const size_t ws_bytes = sizeof(MarchScratch) + (size_t)num_tiles * sizeof(unsigned long long);
        auto *ws = static_cast<MarchScratch *>(workspace(stream, ws_bytes));
        if (!ws) return;
        NGP_CUDA_OK(cudaMemsetAsync(ws, 0, ws_bytes, stream), "march_rays");
        // persistent grid: 4 CTAs of 8 warps per SM keep ~4.7k rays in flight, in ray order; a caller that runs the
        // march underneath other kernels (trainer.py prefetches the next batch) asks for fewer so that it does not
        // crowd them out (ngp_b200_set_march_ctas_per_sm)
        const unsigned per_sm = g_march_ctas_per_sm > 0 ? (unsigned)g_march_ctas_per_sm : 4u;
        {   // An SM only hosts CTAs of one shared-memory carve-out at a time.  The fused MLP backward needs the
            // largest one (225 KB), so a march running underneath it must ask for the same configuration or the
            // backward's CTAs wait until the persistent march CTAs have left their SMs.  Costs the march nothing
            // measurable: its working set (the 256 KB bitfield) is an L2 hit either way.
            static bool configured = false;  // benign race: idempotent
            if (!configured) {
                cudaFuncSetAttribute(march_rays_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                configured = true;
            }
        }
        const unsigned grid = min(num_tiles, 148u * per_sm);
        march_rays_kernel<<<grid, kMarchBlock, 0, stream>>>(*desc, ws, num_tiles, rays_o, rays_d, t_starts, t_ends,
                                                           noises, bitfield, next_loc, exceeded, valid, rays_n,
                                                           rays_start, idcs, xyzs, dirs, dss, z_vals);
        if (!check_launch("march_rays")) return;
    }
    if (desc->total_samples) {
        const unsigned blocks = min(div_up(desc->total_samples, 256u), 148u * 8u);
        march_rays_tail_kernel<<<blocks, 256, 0, stream>>>(desc->total_samples, next_loc, exceeded, idcs, xyzs, dirs,
                                                           dss, z_vals);
        check_launch("march_rays(tail)");
    }
}

static void launch_march_rays_inference(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len, bool in_place) {
    using namespace ngp;
    clear_error();
    const char *op = in_place ? "march_rays_inference_inplace" : "march_rays_inference";
    auto *desc = descriptor<NgpMarchingInferenceDescriptor>(opaque, opaque_len, op);
    if (!desc) return;
    if (desc->K == 0 || desc->G == 0 || desc->G > 1024) {
        set_error(NGP_ERR_ARGUMENT, "%s: expected K > 0 and 0 < G <= 1024, got K=%u G=%u", op, desc->K, desc->G);
        return;
    }
    BufferCursor b{buffers};
    const float *rays_o = b.next<const float>();
    const float *rays_d = b.next<const float>();
    float *t_starts = b.next<float>();
    const float *t_ends = b.next<const float>();
    const uint8_t *bitfield = b.next<const uint8_t>();
    uint32_t *next_in = b.next<uint32_t>();
    const uint8_t *terminated = b.next<const uint8_t>();
    uint32_t *indices_in = b.next<uint32_t>();
    uint32_t *next_out = in_place ? next_in : b.next<uint32_t>();
    uint32_t *indices_out = in_place ? indices_in : b.next<uint32_t>();
    uint32_t *n_samples = b.next<uint32_t>();
    float *t_starts_out = in_place ? t_starts : b.next<float>();
    float *xyzs = b.next<float>();
    float *dss = b.next<float>();
    float *z_vals = b.next<float>();
    float *ray_dirs = in_place ? b.next<float>() : nullptr;
    if (desc->n_rays == 0) {
        if (!in_place)
            NGP_CUDA_OK(cudaMemcpyAsync(next_out, next_in, sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream), op);
        return;
    }
    const unsigned rank_blocks = div_up(desc->n_rays, kRankSlots);
    auto *ws = static_cast<uint32_t *>(workspace(stream, ((size_t)rank_blocks + 2 * (size_t)desc->n_rays + 1) * sizeof(uint32_t)));
    if (!ws) return;
    uint32_t *block_total = ws, *rank_in_block = ws + rank_blocks, *snapshot = ws + rank_blocks + desc->n_rays;
    float *t_admitted = in_place ? nullptr : reinterpret_cast<float *>(snapshot + 1);
    march_rays_inference_rank_kernel<<<rank_blocks, kRankBlock, 0, stream>>>(desc->n_rays, terminated, block_total, rank_in_block,
                                                                             next_in, in_place ? snapshot : nullptr);
    if (!check_launch(op)) return;
    if (!in_place) {  // the renderer's in-place path pre-advances every ray once per frame (ngp_march_rays_skip_empty)
        march_rays_inference_admit_kernel<<<div_up(desc->n_rays, 128), 128, 0, stream>>>(
            *desc, rays_o, rays_d, t_starts, t_ends, bitfield, next_in, terminated, indices_in, block_total, rank_in_block,
            next_out, indices_out, t_admitted);
        if (!check_launch(op)) return;
    }
    // half a warp per slot when one 16-point chunk of the chain covers the cap, else a whole warp
    const bool half = desc->march_steps_cap <= 16;
    const unsigned grid = div_up(desc->n_rays, half ? 2 * kInferWarps : kInferWarps);
#define NGP_MARCH_INF(IP, W, counter, dirs_ptr)                                                                        \
    march_rays_inference_kernel<IP, W><<<grid, kInferWarps * 32, 0, stream>>>(                                         \
        *desc, rays_o, rays_d, t_starts, t_ends, bitfield, counter, terminated, indices_in, block_total, rank_in_block, \
        next_out, indices_out, n_samples, t_starts_out, xyzs, dss, z_vals, dirs_ptr, t_admitted)
    if (in_place) {
        if (half) NGP_MARCH_INF(true, 16, snapshot, ray_dirs); else NGP_MARCH_INF(true, 32, snapshot, ray_dirs);
    } else {
        if (half) NGP_MARCH_INF(false, 16, next_in, nullptr); else NGP_MARCH_INF(false, 32, next_in, nullptr);
    }
#undef NGP_MARCH_INF
    check_launch(op);
}

void ngp_march_rays_inference(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    launch_march_rays_inference(stream, buffers, opaque, opaque_len, false);
}

void ngp_march_rays_inference_inplace(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    launch_march_rays_inference(stream, buffers, opaque, opaque_len, true);
}

void ngp_march_rays_skip_empty(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *desc = descriptor<NgpMarchingInferenceDescriptor>(opaque, opaque_len, "march_rays_skip_empty");
    if (!desc) return;
    if (desc->K == 0 || desc->G == 0 || desc->G > 1024) {
        set_error(NGP_ERR_ARGUMENT, "march_rays_skip_empty: expected K > 0 and 0 < G <= 1024, got K=%u G=%u", desc->K, desc->G);
        return;
    }
    if (desc->n_total_rays == 0) return;
    BufferCursor b{buffers};
    const float *rays_o = b.next<const float>();
    const float *rays_d = b.next<const float>();
    const float *t_starts = b.next<const float>();
    const float *t_ends = b.next<const float>();
    const uint8_t *bitfield = b.next<const uint8_t>();
    float *t_out = b.next<float>();
    march_rays_skip_empty_kernel<<<div_up(desc->n_total_rays, 128), 128, 0, stream>>>(*desc, rays_o, rays_d, t_starts, t_ends,
                                                                                      bitfield, t_out);
    check_launch("march_rays_skip_empty");
}

}  // extern "C"
