// integrating.cu -- volume-rendering integrator (forward, backward, resumable inference) for sm_100a.
//
// Replaces deps/volume-rendering-jax/lib/impl/integrating.cu:24-322,325-507.
//
// The reference walks each ray with one thread (scalar, uncoalesced loads, a latency chain of
// n_samples global loads).  Here a warp walks a ray (rays dealt round-robin to warps, next chunk's
// samples prefetched while the current one is composited): each lane loads one sample with full-width
// coalesced accesses (float4 for drgbs) and evaluates its alpha in parallel; only the
// transmittance recurrence T <- T * (1 - alpha) is replayed in sample order through shuffles, with
// the same two rounded operations per sample as the reference, so the early-stop decision
// (T <= 1e-4) and therefore `measured_batch_size` are integer-exact.  Colour/depth sums are
// warp-tree reductions (abs error ~1e-7, tolerance 1e-4).
#include "common.cuh"

namespace ngp {
namespace {

constexpr float kTThreshold = 1e-4f;  // integrating.cu:12
constexpr int kBlock = 128;

// Replays T_{k+1} = T_k * (1 - alpha_k) over the (up to 32) samples held by the lanes, in sample order, with
// exactly the reference's two rounded operations per sample, and applies its stopping rule
// `for (; T > T_THRESHOLD && idx < n; ++idx)` (integrating.cu:61): sample k is composited iff every
// transmittance before it, including its own T_before, exceeded the threshold.  The 32 products are
// computed unconditionally (lanes >= count contribute the exact factor 1), so the only dependent chain is
// 32 FMULs; the shuffles feeding it are independent and pipeline.
// Returns the number of composited samples; T becomes the transmittance after the last composited one.
__device__ __forceinline__ uint32_t transmittance_chain(float one_minus, uint32_t count, uint32_t lane, float &T,
                                                        float &Tb, bool &active) {
    float run = T, mine = T;
#pragma unroll
    for (uint32_t k = 0; k < 32; ++k) {
        const float a = __shfl_sync(0xffffffffu, one_minus, k);
        if (lane == k) mine = run;  // transmittance before sample k
        run = __fmul_rn(run, a);
    }
    Tb = mine;
    // first sample whose T_before is already at or below the threshold stops the ray
    const uint32_t stop = __ballot_sync(0xffffffffu, !(mine > kTThreshold) || lane >= count);
    const uint32_t processed = stop ? (uint32_t)__ffs(stop) - 1u : 32u;
    active = lane < processed;
    // transmittance after the last composited sample = T_before of the first one that was not
    T = processed < 32u ? __shfl_sync(0xffffffffu, mine, processed & 31u) : run;
    return processed;
}

struct SampleRegs {
    float4 v;
    float z, dt;
};

__device__ __forceinline__ SampleRegs load_sample(const float4 *__restrict__ drgbs, const float *__restrict__ z_vals,
                                                  const float *__restrict__ dss, uint32_t s, bool ok) {
    SampleRegs r;
    r.v = ok ? __ldg(drgbs + s) : make_float4(0.f, 0.f, 0.f, 0.f);
    r.z = ok ? __ldg(z_vals + s) : 0.f;
    r.dt = ok ? __ldg(dss + s) : 0.f;
    return r;
}

// One warp per ray, rays dealt round-robin to warps: the rays that own samples sit at the front of the batch
// (march_rays hands out the budget in ray order), so striding spreads them over all SMs.
__global__ void __launch_bounds__(kBlock) integrate_rays_kernel(
    uint32_t n_rays, const uint32_t *__restrict__ rays_sample_startidx, const uint32_t *__restrict__ rays_n_samples,
    const float *__restrict__ bgs, const float *__restrict__ dss, const float *__restrict__ z_vals,
    const float4 *__restrict__ drgbs, uint32_t *__restrict__ measured_batch_size, float4 *__restrict__ final_rgbds,
    float *__restrict__ final_opacities) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = gridDim.x * (kBlock / 32);
    const uint32_t warp_global = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    uint32_t composited = 0;

    // rays without samples: T = 1, opacity 0, colour = background (integrating.cu:91-96); thread per ray
    for (uint32_t ray = blockIdx.x * kBlock + threadIdx.x; ray < n_rays; ray += gridDim.x * kBlock) {
        if (__ldg(rays_n_samples + ray) == 0) {
            final_opacities[ray] = 0.f;
            final_rgbds[ray] = make_float4(__ldg(bgs + 3 * (size_t)ray + 0), __ldg(bgs + 3 * (size_t)ray + 1),
                                           __ldg(bgs + 3 * (size_t)ray + 2), 0.f);
        }
    }
    for (uint32_t base = 0; base < n_rays; base += warps_total * 32u) {
        // lane l looks at ray (base + l * warps_total + warp_global): 32 candidate rays per warp per round
        const uint32_t cand = base + lane * warps_total + warp_global;
        const uint32_t cand_n = cand < n_rays ? __ldg(rays_n_samples + cand) : 0u;
        uint32_t todo = __ballot_sync(0xffffffffu, cand_n != 0u);
        while (todo) {
            const uint32_t q = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t ray = base + q * warps_total + warp_global;
            const uint32_t n = __shfl_sync(0xffffffffu, cand_n, q);
            const uint32_t start = __ldg(rays_sample_startidx + ray);
            float T = 1.f, r = 0.f, g = 0.f, b = 0.f, depth = 0.f;
            SampleRegs cur = load_sample(drgbs, z_vals, dss, start + lane, lane < n);
            for (uint32_t c = 0; c < n && T > kTThreshold; c += 32) {
                const uint32_t count = min(32u, n - c);
                const SampleRegs nxt = load_sample(drgbs, z_vals, dss, start + c + 32u + lane, c + 32u + lane < n);  // prefetch
                float one_minus = 1.f, alpha = 0.f;
                if (lane < count) {
                    alpha = 1.f - __expf(-cur.v.x * cur.dt);  // integrating.cu:64
                    one_minus = 1.f - alpha;
                }
                float Tb;
                bool active;
                composited += transmittance_chain(one_minus, count, lane, T, Tb, active);
                const float w = active ? Tb * alpha : 0.f;
                r += w * cur.v.y;
                g += w * cur.v.z;
                b += w * cur.v.w;
                depth += w * cur.z;
                cur = nxt;
            }
            r = warp_sum(r);
            g = warp_sum(g);
            b = warp_sum(b);
            depth = warp_sum(depth);
            if (lane == 0) {
                const float opacity = 1.f - T;
                final_opacities[ray] = opacity;
                float4 out;
                if (T <= kTThreshold) {  // integrating.cu:85-90
                    const float idenom = 1.f / opacity;
                    out = make_float4(r * idenom, g * idenom, b * idenom, depth * idenom);
                } else {  // integrating.cu:91-96
                    out = make_float4(r + T * __ldg(bgs + 3 * (size_t)ray + 0), g + T * __ldg(bgs + 3 * (size_t)ray + 1),
                                      b + T * __ldg(bgs + 3 * (size_t)ray + 2), depth);
                }
                final_rgbds[ray] = out;
            }
        }
    }
    // `composited` is identical on all lanes; one atomic per warp
    if (lane == 0 && composited) atomicAdd(measured_batch_size, composited);
}

__device__ __forceinline__ float warp_incl_scan(float v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

__global__ void __launch_bounds__(kBlock) integrate_rays_backward_kernel(
    uint32_t n_rays, float near_distance, const uint32_t *__restrict__ rays_sample_startidx,
    const uint32_t *__restrict__ rays_n_samples, const float *__restrict__ bgs, const float *__restrict__ dss,
    const float *__restrict__ z_vals, const float4 *__restrict__ drgbs, const float4 *__restrict__ final_rgbds,
    const float *__restrict__ final_opacities, const float4 *__restrict__ dL_dfinal_rgbds,
    float *__restrict__ dL_dbgs, float *__restrict__ dL_dz_vals, float4 *__restrict__ dL_ddrgbs) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = gridDim.x * (kBlock / 32);
    const uint32_t warp_global = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);

    // rays without samples: T stays 1, the whole colour gradient flows to the background (integrating.cu:235-239)
    for (uint32_t ray = blockIdx.x * kBlock + threadIdx.x; ray < n_rays; ray += gridDim.x * kBlock) {
        if (__ldg(rays_n_samples + ray) == 0) {
            const float4 dfin = __ldg(dL_dfinal_rgbds + ray);
            dL_dbgs[3 * (size_t)ray + 0] = dfin.x;
            dL_dbgs[3 * (size_t)ray + 1] = dfin.y;
            dL_dbgs[3 * (size_t)ray + 2] = dfin.z;
        }
    }
    for (uint32_t base = 0; base < n_rays; base += warps_total * 32u) {
        const uint32_t cand = base + lane * warps_total + warp_global;
        const uint32_t cand_n = cand < n_rays ? __ldg(rays_n_samples + cand) : 0u;
        uint32_t todo = __ballot_sync(0xffffffffu, cand_n != 0u);
        while (todo) {
            const uint32_t q = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t ray = base + q * warps_total + warp_global;
            const uint32_t n = __shfl_sync(0xffffffffu, cand_n, q);
            const uint32_t start = __ldg(rays_sample_startidx + ray);
            const float4 dfin = __ldg(dL_dfinal_rgbds + ray);
            const float4 fin = __ldg(final_rgbds + ray);
            const float opac = __ldg(final_opacities + ray);
            const bool terminated = opac >= 1.f - kTThreshold;  // integrating.cu:158
            const float bgw = terminated ? 0.f : 1.f - opac;
            const float bg0 = __ldg(bgs + 3 * (size_t)ray + 0) * bgw, bg1 = __ldg(bgs + 3 * (size_t)ray + 1) * bgw,
                        bg2 = __ldg(bgs + 3 * (size_t)ray + 2) * bgw;
            float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, cd = 0.f;  // running composited colour / depth
            SampleRegs cur = load_sample(drgbs, z_vals, dss, start + lane, lane < n);
            for (uint32_t c = 0; c < n && T > kTThreshold; c += 32) {
                const uint32_t count = min(32u, n - c);
                const uint32_t s = start + c + lane;
                const SampleRegs nxt = load_sample(drgbs, z_vals, dss, s + 32u, c + 32u + lane < n);  // prefetch
                float one_minus = 1.f, alpha = 0.f;
                if (lane < count) {
                    alpha = 1.f - __expf(-cur.v.x * cur.dt);  // integrating.cu:181
                    one_minus = 1.f - alpha;
                }
                float Tb;
                bool active;
                transmittance_chain(one_minus, count, lane, T, Tb, active);
                const float w = active ? Tb * alpha : 0.f;
                const float Ta = Tb * one_minus;  // transmittance after this sample, integrating.cu:193
                const float ir = cr + warp_incl_scan(w * cur.v.y, lane);
                const float ig = cg + warp_incl_scan(w * cur.v.z, lane);
                const float ib = cb + warp_incl_scan(w * cur.v.w, lane);
                const float id = cd + warp_incl_scan(w * cur.z, lane);
                cr = __shfl_sync(0xffffffffu, ir, 31);
                cg = __shfl_sync(0xffffffffu, ig, 31);
                cb = __shfl_sync(0xffffffffu, ib, 31);
                cd = __shfl_sync(0xffffffffu, id, 31);
                if (active) {
                    const float z = cur.z;
                    dL_dz_vals[s] = w * dfin.w;  // integrating.cu:196
                    const float acc = dfin.x * (Ta * cur.v.y - (fin.x - ir) - bg0) + dfin.y * (Ta * cur.v.z - (fin.y - ig) - bg1) +
                                      dfin.z * (Ta * cur.v.w - (fin.z - ib) - bg2) + dfin.w * (Ta * z - (fin.w - id));
                    const float dsig = cur.dt * acc;                                          // :199-215
                    const float reg = (cur.v.x > 4e-5f && z < near_distance) ? 1e-4f : 0.f;   // :221
                    const float scal = fminf(z * z, 1.f);                                     // :225
                    dL_ddrgbs[s] = make_float4(scal * dsig + reg, w * dfin.x, w * dfin.y, w * dfin.z);
                }
                cur = nxt;
            }
            if (lane == 0 && T > kTThreshold) {  // integrating.cu:235-239
                dL_dbgs[3 * (size_t)ray + 0] = T * dfin.x;
                dL_dbgs[3 * (size_t)ray + 1] = T * dfin.y;
                dL_dbgs[3 * (size_t)ray + 2] = T * dfin.z;
            }
        }
    }
}

__global__ void __launch_bounds__(kBlock) count_valid_rays_kernel(uint32_t n, const uint8_t *__restrict__ valid,
                                                                   uint32_t *__restrict__ n_valid) {
    uint32_t c = 0;
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) c += valid[i] ? 1u : 0u;
    c = warp_sum_u32(c);
    if ((threadIdx.x & 31u) == 0 && c) atomicAdd(n_valid, c);
}

// Training fast path (ngp_integrate_loss_fused): integrate_rays, the Huber loss against the ground-truth pixels and
// integrate_rays_backward for the colour/density gradient in ONE pass per ray -- the warp that composited a ray
// keeps its result in registers, forms dL/dfinal from it and walks the ray's samples (still in L1/L2) a second time
// for dL/ddrgbs.  Same arithmetic in the same order as the three ops (final_rgbds, opacities and dL_ddrgbs come out
// bit-identical); the background / depth gradients nothing consumes in training are not produced.
__global__ void __launch_bounds__(kBlock) integrate_loss_fused_kernel(
    NgpIntegrateLossDescriptor p, const uint32_t *__restrict__ rays_sample_startidx, const uint32_t *__restrict__ rays_n_samples,
    const float *__restrict__ bgs, const float *__restrict__ dss, const float *__restrict__ z_vals,
    const float4 *__restrict__ drgbs, const uint8_t *__restrict__ valid, const int32_t *__restrict__ perm,
    const uchar4 *__restrict__ rgbas, const uint32_t *__restrict__ n_valid_ptr, uint32_t *__restrict__ measured_batch_size,
    float4 *__restrict__ final_rgbds, float *__restrict__ final_opacities, float4 *__restrict__ dL_ddrgbs,
    float *__restrict__ loss) {
    const uint32_t n_rays = p.n_rays;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = gridDim.x * (kBlock / 32);
    const uint32_t warp_global = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    const float inv_n = 1.f / (float)__ldg(n_valid_ptr);
    uint32_t composited = 0;
    float loss_acc = 0.f;

    // rays without samples: colour = background, no sample gradients; they still count in the loss
    for (uint32_t ray = blockIdx.x * kBlock + threadIdx.x; ray < n_rays; ray += gridDim.x * kBlock) {
        if (__ldg(rays_n_samples + ray) == 0) {
            const float b0 = __ldg(bgs + 3 * (size_t)ray + 0), b1 = __ldg(bgs + 3 * (size_t)ray + 1), b2 = __ldg(bgs + 3 * (size_t)ray + 2);
            const float4 out = make_float4(b0, b1, b2, 0.f);
            final_opacities[ray] = 0.f;
            final_rgbds[ray] = out;
            if (valid[ray]) {
                float per_ray;
                huber_ray(out, __ldg(rgbas + (uint32_t)__ldg(perm + ray)), b0, b1, b2, p.delta, inv_n, per_ray);
                loss_acc += per_ray;
            }
        }
    }
    for (uint32_t base = 0; base < n_rays; base += warps_total * 32u) {
        const uint32_t cand = base + lane * warps_total + warp_global;
        const uint32_t cand_n = cand < n_rays ? __ldg(rays_n_samples + cand) : 0u;
        uint32_t todo = __ballot_sync(0xffffffffu, cand_n != 0u);
        while (todo) {
            const uint32_t q = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t ray = base + q * warps_total + warp_global;
            const uint32_t n = __shfl_sync(0xffffffffu, cand_n, q);
            const uint32_t start = __ldg(rays_sample_startidx + ray);
            const float b0 = __ldg(bgs + 3 * (size_t)ray + 0), b1 = __ldg(bgs + 3 * (size_t)ray + 1), b2 = __ldg(bgs + 3 * (size_t)ray + 2);
            // ---- forward (integrate_rays_kernel)
            float T = 1.f, r = 0.f, g = 0.f, b = 0.f, depth = 0.f;
            {
                SampleRegs cur = load_sample(drgbs, z_vals, dss, start + lane, lane < n);
                for (uint32_t c = 0; c < n && T > kTThreshold; c += 32) {
                    const uint32_t count = min(32u, n - c);
                    const SampleRegs nxt = load_sample(drgbs, z_vals, dss, start + c + 32u + lane, c + 32u + lane < n);
                    float one_minus = 1.f, alpha = 0.f;
                    if (lane < count) {
                        alpha = 1.f - __expf(-cur.v.x * cur.dt);
                        one_minus = 1.f - alpha;
                    }
                    float Tb;
                    bool active;
                    composited += transmittance_chain(one_minus, count, lane, T, Tb, active);
                    const float w = active ? Tb * alpha : 0.f;
                    r += w * cur.v.y;
                    g += w * cur.v.z;
                    b += w * cur.v.w;
                    depth += w * cur.z;
                    cur = nxt;
                }
            }
            r = warp_sum(r);
            g = warp_sum(g);
            b = warp_sum(b);
            depth = warp_sum(depth);
            const float opac = 1.f - T;
            float4 fin;
            if (T <= kTThreshold) {
                const float idenom = 1.f / opac;
                fin = make_float4(r * idenom, g * idenom, b * idenom, depth * idenom);
            } else {
                fin = make_float4(r + T * b0, g + T * b1, b + T * b2, depth);
            }
            if (lane == 0) {
                final_opacities[ray] = opac;
                final_rgbds[ray] = fin;
            }
            // ---- loss and its gradient (huber_loss_grad_kernel)
            float4 dfin = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid[ray]) {
                float per_ray;
                dfin = huber_ray(fin, __ldg(rgbas + (uint32_t)__ldg(perm + ray)), b0, b1, b2, p.delta, inv_n, per_ray);
                if (lane == 0) loss_acc += per_ray;
            }
            // ---- backward (integrate_rays_backward_kernel), colour / density gradient only
            const bool terminated = opac >= 1.f - kTThreshold;
            const float bgw = terminated ? 0.f : 1.f - opac;
            const float bg0 = b0 * bgw, bg1 = b1 * bgw, bg2 = b2 * bgw;
            float cr = 0.f, cg = 0.f, cb = 0.f, cd = 0.f;
            T = 1.f;
            SampleRegs cur = load_sample(drgbs, z_vals, dss, start + lane, lane < n);
            for (uint32_t c = 0; c < n && T > kTThreshold; c += 32) {
                const uint32_t count = min(32u, n - c);
                const uint32_t s = start + c + lane;
                const SampleRegs nxt = load_sample(drgbs, z_vals, dss, s + 32u, c + 32u + lane < n);
                float one_minus = 1.f, alpha = 0.f;
                if (lane < count) {
                    alpha = 1.f - __expf(-cur.v.x * cur.dt);
                    one_minus = 1.f - alpha;
                }
                float Tb;
                bool active;
                transmittance_chain(one_minus, count, lane, T, Tb, active);
                const float w = active ? Tb * alpha : 0.f;
                const float Ta = Tb * one_minus;
                const float ir = cr + warp_incl_scan(w * cur.v.y, lane);
                const float ig = cg + warp_incl_scan(w * cur.v.z, lane);
                const float ib = cb + warp_incl_scan(w * cur.v.w, lane);
                const float id = cd + warp_incl_scan(w * cur.z, lane);
                cr = __shfl_sync(0xffffffffu, ir, 31);
                cg = __shfl_sync(0xffffffffu, ig, 31);
                cb = __shfl_sync(0xffffffffu, ib, 31);
                cd = __shfl_sync(0xffffffffu, id, 31);
                if (active) {
                    const float z = cur.z;
                    const float acc = dfin.x * (Ta * cur.v.y - (fin.x - ir) - bg0) + dfin.y * (Ta * cur.v.z - (fin.y - ig) - bg1) +
                                      dfin.z * (Ta * cur.v.w - (fin.z - ib) - bg2) + dfin.w * (Ta * z - (fin.w - id));
                    const float dsig = cur.dt * acc;
                    const float reg = (cur.v.x > 4e-5f && z < p.near_distance) ? 1e-4f : 0.f;
                    const float scal = fminf(z * z, 1.f);
                    dL_ddrgbs[s] = make_float4(scal * dsig + reg, w * dfin.x, w * dfin.y, w * dfin.z);
                }
                cur = nxt;
            }
        }
    }
    if (lane == 0 && composited) atomicAdd(measured_batch_size, composited);
    loss_acc = warp_sum(loss_acc);
    if (lane == 0 && loss_acc != 0.f) atomicAdd(loss, loss_acc * inv_n);
}

// kInPlace (renderer fast path, ngp_integrate_rays_inference_inplace): results go straight to row `ray` of the frame
// state (the reference scatters them afterwards, integrating/__init__.py:108-109; every ray belongs to one slot) and
// the terminated-ray and sample counts are ACCUMULATED into 64-bit counters instead of being returned per call.
template <bool kInPlace>
__global__ void __launch_bounds__(kBlock) integrate_rays_inference_kernel(
    NgpIntegratingInferenceDescriptor p, const float *__restrict__ rays_bg, const float4 *rays_rgbd,
    const float *rays_T, const uint32_t *__restrict__ n_samples, const uint32_t *__restrict__ indices,
    const float *__restrict__ dss, const float *__restrict__ z_vals, const float4 *__restrict__ drgbs,
    uint32_t *__restrict__ terminate_cnt, uint8_t *__restrict__ terminated, float4 *rays_rgbd_out,
    float *rays_T_out, unsigned long long *__restrict__ counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool term = false;
    uint32_t my_samples = 0;
    if (i < p.n_rays) {
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
        float T_out = 0.f;
        const uint32_t ns = __ldg(n_samples + i), ray = __ldg(indices + i);
        my_samples = ns;
        if (ray < p.n_total_rays) {
            const uint32_t cap = p.march_steps_cap;
            const float *__restrict__ rds = dss + (size_t)i * cap;
            const float *__restrict__ rz = z_vals + (size_t)i * cap;
            const float4 *__restrict__ rc = drgbs + (size_t)i * cap;
            float T = rays_T[ray];
            float4 acc = rays_rgbd[ray];
            for (uint32_t s = 0; T > kTThreshold && s < ns; ++s) {  // integrating.cu:278-291
                const float4 v = __ldg(rc + s);
                const float alpha = 1.f - __expf(-v.x * __ldg(rds + s));
                const float w = T * alpha;
                acc.x += w * v.y;
                acc.y += w * v.z;
                acc.z += w * v.w;
                acc.w += w * __ldg(rz + s);
                T *= (1.f - alpha);
            }
            if (T <= kTThreshold) {  // integrating.cu:293-301
                const float idenom = 1.f / (1.f - T);
                term = true;
                T_out = 0.f;
                out = make_float4(acc.x * idenom, acc.y * idenom, acc.z * idenom, acc.w * idenom);
            } else {  // integrating.cu:302-314
                term = ns < cap;
                T_out = T;
                out = acc;
                if (term) {
                    out.x = acc.x + T * __ldg(rays_bg + 3 * (size_t)ray + 0);
                    out.y = acc.y + T * __ldg(rays_bg + 3 * (size_t)ray + 1);
                    out.z = acc.z + T * __ldg(rays_bg + 3 * (size_t)ray + 2);
                }
            }
        }
        terminated[i] = term ? 1 : 0;
        if (!kInPlace) {
            rays_rgbd_out[i] = out;
            rays_T_out[i] = T_out;
        } else if (ray < p.n_total_rays) {
            rays_rgbd_out[ray] = out;
            rays_T_out[ray] = T_out;
        }
    }
    const uint32_t votes = __popc(__ballot_sync(0xffffffffu, term));
    if (!kInPlace) {
        if ((threadIdx.x & 31u) == 0 && votes) atomicAdd(terminate_cnt, votes);
    } else {
        const uint32_t smp = warp_sum_u32(my_samples);
        if ((threadIdx.x & 31u) == 0) {
            if (votes) atomicAdd(counters + 0, (unsigned long long)votes);
            if (smp) atomicAdd(counters + 1, (unsigned long long)smp);
        }
    }
}

}  // namespace
}  // namespace ngp

extern "C" {

void ngp_integrate_rays(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *desc = descriptor<NgpIntegratingDescriptor>(opaque, opaque_len, "integrate_rays");
    if (!desc) return;
    BufferCursor b{buffers};
    const uint32_t *start = b.next<const uint32_t>();
    const uint32_t *ns = b.next<const uint32_t>();
    const float *bgs = b.next<const float>();
    const float *dss = b.next<const float>();
    const float *z_vals = b.next<const float>();
    const float4 *drgbs = b.next<const float4>();
    uint32_t *mbs = b.next<uint32_t>();
    float4 *final_rgbds = b.next<float4>();
    float *final_opacities = b.next<float>();
    NGP_CUDA_OK(cudaMemsetAsync(mbs, 0, sizeof(uint32_t), stream), "integrate_rays");
    if (desc->n_rays == 0) return;
    const unsigned blocks = min(div_up(desc->n_rays, kBlock), 148u * 8u);
    integrate_rays_kernel<<<blocks, kBlock, 0, stream>>>(desc->n_rays, start, ns, bgs, dss, z_vals, drgbs, mbs,
                                                         final_rgbds, final_opacities);
    check_launch("integrate_rays");
}

void ngp_integrate_rays_backward(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *desc = descriptor<NgpIntegratingBackwardDescriptor>(opaque, opaque_len, "integrate_rays_backward");
    if (!desc) return;
    BufferCursor b{buffers};
    const uint32_t *start = b.next<const uint32_t>();
    const uint32_t *ns = b.next<const uint32_t>();
    const float *bgs = b.next<const float>();
    const float *dss = b.next<const float>();
    const float *z_vals = b.next<const float>();
    const float4 *drgbs = b.next<const float4>();
    const float4 *final_rgbds = b.next<const float4>();
    const float *final_opacities = b.next<const float>();
    const float4 *dL_dfinal = b.next<const float4>();
    float *dL_dbgs = b.next<float>();
    float *dL_dz = b.next<float>();
    float4 *dL_dd = b.next<float4>();
    // samples that are not composited (after early stop, or unused slots) get zero gradient
    NGP_CUDA_OK(cudaMemsetAsync(dL_dbgs, 0, (size_t)desc->n_rays * 3 * sizeof(float), stream), "integrate_rays_backward");
    NGP_CUDA_OK(cudaMemsetAsync(dL_dz, 0, (size_t)desc->total_samples * sizeof(float), stream), "integrate_rays_backward");
    NGP_CUDA_OK(cudaMemsetAsync(dL_dd, 0, (size_t)desc->total_samples * 4 * sizeof(float), stream), "integrate_rays_backward");
    if (desc->n_rays == 0) return;
    const unsigned blocks = min(div_up(desc->n_rays, kBlock), 148u * 8u);
    integrate_rays_backward_kernel<<<blocks, kBlock, 0, stream>>>(desc->n_rays, desc->near_distance, start, ns, bgs,
                                                                  dss, z_vals, drgbs, final_rgbds, final_opacities,
                                                                  dL_dfinal, dL_dbgs, dL_dz, dL_dd);
    check_launch("integrate_rays_backward");
}

void ngp_integrate_rays_inference(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *desc = descriptor<NgpIntegratingInferenceDescriptor>(opaque, opaque_len, "integrate_rays_inference");
    if (!desc) return;
    BufferCursor b{buffers};
    const float *rays_bg = b.next<const float>();
    const float4 *rays_rgbd = b.next<const float4>();
    const float *rays_T = b.next<const float>();
    const uint32_t *ns = b.next<const uint32_t>();
    const uint32_t *indices = b.next<const uint32_t>();
    const float *dss = b.next<const float>();
    const float *z_vals = b.next<const float>();
    const float4 *drgbs = b.next<const float4>();
    uint32_t *terminate_cnt = b.next<uint32_t>();
    uint8_t *terminated = b.next<uint8_t>();
    float4 *rgbd_out = b.next<float4>();
    float *T_out = b.next<float>();
    NGP_CUDA_OK(cudaMemsetAsync(terminate_cnt, 0, sizeof(uint32_t), stream), "integrate_rays_inference");
    if (desc->n_rays == 0) return;
    integrate_rays_inference_kernel<false><<<div_up(desc->n_rays, kBlock), kBlock, 0, stream>>>(
        *desc, rays_bg, rays_rgbd, rays_T, ns, indices, dss, z_vals, drgbs, terminate_cnt, terminated, rgbd_out, T_out, nullptr);
    check_launch("integrate_rays_inference");
}

void ngp_integrate_rays_inference_inplace(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *desc = descriptor<NgpIntegratingInferenceDescriptor>(opaque, opaque_len, "integrate_rays_inference_inplace");
    if (!desc || desc->n_rays == 0) return;
    BufferCursor b{buffers};
    const float *rays_bg = b.next<const float>();
    float4 *rays_rgbd = b.next<float4>();
    float *rays_T = b.next<float>();
    const uint32_t *ns = b.next<const uint32_t>();
    const uint32_t *indices = b.next<const uint32_t>();
    const float *dss = b.next<const float>();
    const float *z_vals = b.next<const float>();
    const float4 *drgbs = b.next<const float4>();
    uint8_t *terminated = b.next<uint8_t>();
    auto *counters = b.next<unsigned long long>();
    integrate_rays_inference_kernel<true><<<div_up(desc->n_rays, kBlock), kBlock, 0, stream>>>(
        *desc, rays_bg, rays_rgbd, rays_T, ns, indices, dss, z_vals, drgbs, nullptr, terminated, rays_rgbd, rays_T, counters);
    check_launch("integrate_rays_inference_inplace");
}

void ngp_integrate_loss_fused(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *desc = descriptor<NgpIntegrateLossDescriptor>(opaque, opaque_len, "integrate_loss_fused");
    if (!desc) return;
    BufferCursor b{buffers};
    const uint32_t *startidx = b.next<const uint32_t>();
    const uint32_t *ns = b.next<const uint32_t>();
    const float *bgs = b.next<const float>();
    const float *dss = b.next<const float>();
    const float *z_vals = b.next<const float>();
    const float4 *drgbs = b.next<const float4>();
    const uint8_t *valid = b.next<const uint8_t>();
    const int32_t *perm = b.next<const int32_t>();
    const uchar4 *rgbas = b.next<const uchar4>();
    uint32_t *mbs = b.next<uint32_t>();
    float4 *final_rgbds = b.next<float4>();
    float *final_opacities = b.next<float>();
    float4 *dL_dd = b.next<float4>();
    float *loss = b.next<float>();
    uint32_t *n_valid = b.next<uint32_t>();
    NGP_CUDA_OK(cudaMemsetAsync(dL_dd, 0, (size_t)desc->total_samples * 4 * sizeof(float), stream), "integrate_loss_fused");
    // the three scalar outputs are zero-filled with one memset node when the caller laid them out back to back
    // (trainops.integrate_loss_fused does)
    if (n_valid == mbs + 1 && reinterpret_cast<uint32_t *>(loss) == mbs + 2) {
        NGP_CUDA_OK(cudaMemsetAsync(mbs, 0, 3 * sizeof(uint32_t), stream), "integrate_loss_fused");
    } else {
        NGP_CUDA_OK(cudaMemsetAsync(mbs, 0, sizeof(uint32_t), stream), "integrate_loss_fused");
        NGP_CUDA_OK(cudaMemsetAsync(n_valid, 0, sizeof(uint32_t), stream), "integrate_loss_fused");
        NGP_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float), stream), "integrate_loss_fused");
    }
    if (desc->n_rays == 0) return;
    count_valid_rays_kernel<<<min(div_up(desc->n_rays, kBlock), 148u * 4u), kBlock, 0, stream>>>(desc->n_rays, valid, n_valid);
    integrate_loss_fused_kernel<<<min(div_up(desc->n_rays, kBlock), 148u * 8u), kBlock, 0, stream>>>(
        *desc, startidx, ns, bgs, dss, z_vals, drgbs, valid, perm, rgbas, n_valid, mbs, final_rgbds, final_opacities, dL_dd, loss);
    check_launch("integrate_loss_fused");
}

}  // extern "C"
