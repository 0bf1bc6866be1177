// marching.cuh -- device code of the occupancy-grid march shared by marching.cu (the drop-in ops) and render_persistent.cu
// (the whole-frame inference kernel): grid / ray set-up, one reference step (marching.cu:166-190), the evaluation of a
// chain point and the chain t_{k+1} = t_k + ds(t_k) itself.  Every floating-point operation that decides where a sample
// lands is pinned to the shape nvcc gives the reference source.
#pragma once
#include "common.cuh"

namespace ngp {
namespace march {

constexpr float kTwoSqrt3 = 3.4641015529632568359f;  // 2 * (float)SQRT3, volrend.h:18

__device__ __forceinline__ uint32_t div_up_dev(uint32_t a, uint32_t b) { return (a + b - 1u) / b; }

struct Grid {
    uint32_t K, G, G3;
    float Gf, inv_G, bound, portion, ds_lo, ds_hi;
    const uint32_t *__restrict__ bits;  // bitfield viewed as little-endian 32-bit words
};

__device__ __forceinline__ Grid make_grid(uint32_t steps, uint32_t K, uint32_t G, float bound, float portion,
                                          const uint8_t *bitfield) {
    Grid g;
    g.K = K;
    g.G = G;
    g.G3 = G * G * G;
    g.Gf = (float)G;
    g.inv_G = __fdiv_rn(1.f, g.Gf);                                                    // marching.cu:158
    g.bound = bound;
    g.portion = portion;
    g.ds_lo = __fdiv_rn(__fmul_rn(fminf(bound, 1.f), kTwoSqrt3), (float)steps);        // marching.cu:20
    g.ds_hi = __fmul_rn(__fmul_rn(bound, kTwoSqrt3), g.inv_G);                         // marching.cu:21
    g.bits = reinterpret_cast<const uint32_t *>(bitfield);
    return g;
}

__device__ __forceinline__ float calc_ds(const Grid &g, float t) {  // marching.cu:15-23
    return fminf(fmaxf(__fmul_rn(t, g.portion), g.ds_lo), g.ds_hi);
}

__device__ __forceinline__ uint32_t mip_of(float v, uint32_t K) {  // marching.cu:25-39
    int e;
    frexpf(v, &e);
    return (uint32_t)min(max(e, 0), (int)K - 1);
}

struct Ray {
    float ox, oy, oz, dx, dy, dz, ix, iy, iz;
};

__device__ __forceinline__ Ray load_ray(const float *__restrict__ rays_o, const float *__restrict__ rays_d, uint32_t r) {
    Ray ray;
    ray.ox = __ldg(rays_o + 3 * (size_t)r + 0);
    ray.oy = __ldg(rays_o + 3 * (size_t)r + 1);
    ray.oz = __ldg(rays_o + 3 * (size_t)r + 2);
    ray.dx = __ldg(rays_d + 3 * (size_t)r + 0);
    ray.dy = __ldg(rays_d + 3 * (size_t)r + 1);
    ray.dz = __ldg(rays_d + 3 * (size_t)r + 2);
    ray.ix = __fdiv_rn(1.f, ray.dx);  // marching.cu:157
    ray.iy = __fdiv_rn(1.f, ray.dy);
    ray.iz = __fdiv_rn(1.f, ray.dz);
    return ray;
}

struct Step {
    float px, py, pz, ds, t_next;
    bool occupied;
};

// distance along one axis to the next voxel boundary (marching.cu:182-183)
__device__ __forceinline__ float axis_delta(float gp, float d, float inv_d, float pos, float inv_G, float mip_bound) {
    float ng = floorf(__fmaf_rn(copysignf(1.f, d), .5f, __fadd_rn(gp, .5f)));
    float a = __fmaf_rn(ng, inv_G, -.5f);
    a = __fadd_rn(a, a);
    return __fmul_rn(__fmaf_rn(mip_bound, a, -pos), inv_d);
}

// One marching step at parameter t (marching.cu:166-190)
template <bool kWantSkip>
__device__ __forceinline__ Step march_step(const Grid &g, const Ray &r, float t) {
    Step s;
    s.px = __fmaf_rn(t, r.dx, r.ox);
    s.py = __fmaf_rn(t, r.dy, r.oy);
    s.pz = __fmaf_rn(t, r.dz, r.oz);
    s.ds = calc_ds(g, t);
    uint32_t cascade = 0;
    if (g.K > 1) {  // marching.cu:79-98
        float linf = fmaxf(fabsf(s.px), fmaxf(fabsf(s.py), fabsf(s.pz)));
        cascade = max(mip_of(linf, g.K), mip_of(__fmul_rn(s.ds, g.Gf), g.K));
    }
    float mip_bound = fminf((float)(1u << cascade), g.bound);
    float gx = __fmul_rn(__fmul_rn(__fadd_rn(__fdiv_rn(s.px, mip_bound), 1.f), .5f), g.Gf);
    float gy = __fmul_rn(__fmul_rn(__fadd_rn(__fdiv_rn(s.py, mip_bound), 1.f), .5f), g.Gf);
    float gz = __fmul_rn(__fmul_rn(__fadd_rn(__fdiv_rn(s.pz, mip_bound), 1.f), .5f), g.Gf);
    int gmax = (int)g.G - 1;
    uint32_t ux = (uint32_t)min(max(__float2int_rd(gx), 0), gmax);  // marching.cu:41-50
    uint32_t uy = (uint32_t)min(max(__float2int_rd(gy), 0), gmax);
    uint32_t uz = (uint32_t)min(max(__float2int_rd(gz), 0), gmax);
    uint32_t idx = cascade * g.G3 + morton3d_encode(ux, uy, uz);
    s.occupied = (__ldg(g.bits + (idx >> 5)) >> (idx & 31u)) & 1u;  // == byte[idx>>3] & (1 << (idx&7))
    s.t_next = __fadd_rn(t, s.ds);
    if (kWantSkip && !s.occupied) {
        float ax = axis_delta(gx, r.dx, r.ix, s.px, g.inv_G, mip_bound);
        float ay = axis_delta(gy, r.dy, r.iy, s.py, g.inv_G, mip_bound);
        float az = axis_delta(gz, r.dz, r.iz, s.pz, g.inv_G, mip_bound);
        float next_t = __fadd_rn(t, fmaxf(0.f, fminf(ax, fminf(ay, az))));
        // (bounded: a non-finite boundary distance would spin forever here, as it does in the reference)
        for (uint32_t guard = 0; s.t_next < next_t && guard < (1u << 20); ++guard) s.t_next = __fadd_rn(s.t_next, calc_ds(g, s.t_next));
    }
    return s;
}

struct EvalPoint {
    float px, py, pz, ds, next_t;
    bool occupied;
};

// occupancy at chain point t, and (if empty) the parameter of the next voxel boundary
__device__ __forceinline__ EvalPoint eval_point(const Grid &g, const Ray &r, float t) {
    EvalPoint s;
    s.px = __fmaf_rn(t, r.dx, r.ox);
    s.py = __fmaf_rn(t, r.dy, r.oy);
    s.pz = __fmaf_rn(t, r.dz, r.oz);
    s.ds = calc_ds(g, t);
    uint32_t cascade = 0;
    if (g.K > 1) {
        float linf = fmaxf(fabsf(s.px), fmaxf(fabsf(s.py), fabsf(s.pz)));
        cascade = max(mip_of(linf, g.K), mip_of(__fmul_rn(s.ds, g.Gf), g.K));
    }
    const float mip_bound = fminf((float)(1u << cascade), g.bound);
    // pos / mip_bound: when mip_bound is a power of two the quotient is an exact scaling, so the
    // multiplication by its (exact) reciprocal rounds identically to the IEEE division
    const uint32_t mb_bits = __float_as_uint(mip_bound);
    float qx, qy, qz;
    if ((mb_bits & 0x007FFFFFu) == 0u) {
        const float inv = __uint_as_float(0x7F000000u - mb_bits);  // 2^-e for 2^e
        qx = __fmul_rn(s.px, inv);
        qy = __fmul_rn(s.py, inv);
        qz = __fmul_rn(s.pz, inv);
    } else {
        qx = __fdiv_rn(s.px, mip_bound);
        qy = __fdiv_rn(s.py, mip_bound);
        qz = __fdiv_rn(s.pz, mip_bound);
    }
    const float gx = __fmul_rn(__fmul_rn(__fadd_rn(qx, 1.f), .5f), g.Gf);
    const float gy = __fmul_rn(__fmul_rn(__fadd_rn(qy, 1.f), .5f), g.Gf);
    const float gz = __fmul_rn(__fmul_rn(__fadd_rn(qz, 1.f), .5f), g.Gf);
    const int gmax = (int)g.G - 1;
    const uint32_t ux = (uint32_t)min(max(__float2int_rd(gx), 0), gmax);
    const uint32_t uy = (uint32_t)min(max(__float2int_rd(gy), 0), gmax);
    const uint32_t uz = (uint32_t)min(max(__float2int_rd(gz), 0), gmax);
    const uint32_t idx = cascade * g.G3 + morton3d_encode(ux, uy, uz);
    s.occupied = (__ldg(g.bits + (idx >> 5)) >> (idx & 31u)) & 1u;
    s.next_t = 0.f;
    if (!s.occupied) {
        const float ax = axis_delta(gx, r.dx, r.ix, s.px, g.inv_G, mip_bound);
        const float ay = axis_delta(gy, r.dy, r.iy, s.py, g.inv_G, mip_bound);
        const float az = axis_delta(gz, r.dz, r.iz, s.pz, g.inv_G, mip_bound);
        s.next_t = __fadd_rn(t, fmaxf(0.f, fminf(ax, fminf(ay, az))));
    }
    return s;
}

// The chain t_{k+1} = fl(t_k + ds(t_k)) from `t_base`: lane j gets t_j, `t_next_base` = t_32 (warp-uniform).
// Constant ds (stepsize_portion == 0, the NeRF-synthetic configuration): while t stays inside one binade
// every t_k is a multiple of u = ulp(t_base) and fl(t_k + ds) = t_k + d with the SAME d = fl(t_base + ds) -
// t_base (ds rounded to the u grid), unless ds sits exactly on a rounding tie of that grid (then the
// result depends on the parity of t_k).  So when t_base + 32 d is still in t_base's binade and there is no
// tie, t_j = t_base + j d exactly -- one FMA per lane whose exact result is representable -- bit-equal to
// the 31 dependent additions.  Otherwise (3 binade crossings per ray, ties, t_base == 0) the additions are
// replayed one by one.
// W = lanes that share the chain (32: the whole warp; 16: half a warp, `mask` = that half's lanes, `lane` < 16)
template <int W = 32>
__device__ __forceinline__ float chain_points(const Grid &g, float t_base, uint32_t lane, bool const_ds, float ds0,
                                              float &t_next_base, uint32_t mask = 0xffffffffu) {
    if (const_ds) {
        const float t1 = __fadd_rn(t_base, ds0);
        const float d = __fadd_rn(t1, -t_base);            // exact (Sterbenz-like: both on the u grid, same binade checked below)
        const float err = __fadd_rn(ds0, -d);              // exact rounding error of t_base + ds0
        const float t32 = __fmaf_rn((float)W, d, t_base);  // t_W
        const uint32_t e0 = __float_as_uint(t_base) >> 23, e32 = __float_as_uint(t32) >> 23;  // sign 0: biased exponents
        const float half_u = __uint_as_float(((e0 > 24u ? e0 : 24u) - 24u) << 23);             // ulp(t_base) / 2
        const bool fast = t_base > 0.f && e0 == e32 && e0 > 24u && fabsf(err) != half_u && d > 0.f;
        if (fast) {
            t_next_base = t32;
            return __fmaf_rn((float)lane, d, t_base);
        }
        float t = t_base;
#pragma unroll
        for (int i = 0; i < W - 1; ++i) {
            const float tn = __fadd_rn(t, ds0);
            if (i < (int)lane) t = tn;
        }
        const float t_last = __shfl_sync(mask, t, W - 1, W);
        t_next_base = __fadd_rn(t_last, ds0);
        return t;
    }
    float t = t_base;
#pragma unroll 4
    for (int i = 0; i < W - 1; ++i) {
        const float tn = __fadd_rn(t, calc_ds(g, t));
        if (i < (int)lane) t = tn;
    }
    const float t_last = __shfl_sync(mask, t, W - 1, W);
    t_next_base = __fadd_rn(t_last, calc_ds(g, t_last));
    return t;
}


// One pass of the reference's inference loop (marching.cu:323-394) for one ray with W lanes (32: a warp; 16: half a warp,
// `mask` = that half's lanes, `shift` = its position, `lane` < W): marches from `t_cur` (always a point of the chain)
// until `cap` samples are out or the ray ends, applies the far-plane rule, advances `t_cur` and returns the number of
// samples.  `emit(w, point, t)` is called by the lane that holds sample w; `fix_ds(w, ds)` by lane 0 when the far-plane
// rule shortens the last sample.  Shared by march_rays_inference (emit = stores into the op's outputs) and the
// persistent frame kernel (emit = shared-memory staging): same code, same samples.
template <int W, class Emit, class FixDs>
__device__ __forceinline__ uint32_t march_chunk(const Grid &g, const Ray &ray, float &t_cur, float t_end, uint32_t cap, uint32_t lane,
                                                uint32_t mask, uint32_t shift, Emit emit, FixDs fix_ds) {
    constexpr uint32_t kW = W, kWMask = W == 32 ? 0xFFFFFFFFu : 0xFFFFu;
    auto ballot = [&](bool pred) { return (__ballot_sync(mask, pred) >> shift) & kWMask; };
    uint32_t steps = 0;
    const bool const_ds = g.portion == 0.f;
    const float ds0 = calc_ds(g, 0.f);
    float last_ds = 0.f, last_z = 0.f;  // most recent sample (for the far-plane clip below)
    // marching.cu:323-365: `t_cur` is the reference's loop variable t, always a point of the chain
    while (steps < cap && t_cur < t_end) {
        float t_next_base;
        const float t = chain_points<W>(g, t_cur, lane, const_ds, ds0, t_next_base, mask);
        const EvalPoint s = eval_point(g, ray, t);
        const uint32_t V = ballot(t < t_end);
        const uint32_t O = ballot(s.occupied);
        uint32_t v = 0;  // chain index of the point the reference visits next
        for (;;) {
            if (v >= kW) { t_cur = t_next_base; break; }
            if (steps >= cap || !((V >> v) & 1u)) { t_cur = __shfl_sync(mask, t, v, W); break; }
            if ((O >> v) & 1u) {
                const uint32_t rest = (O & V) >> v;  // consecutive occupied in-range points starting at v
                uint32_t run = (rest == (kWMask >> v)) ? kW - v : (uint32_t)__ffs(~rest) - 1u;
                run = min(run, cap - steps);
                if (lane >= v && lane < v + run) emit(steps + (lane - v), s, t);
                steps += run;
                v += run;
                last_ds = __shfl_sync(mask, s.ds, v - 1u, W);
                last_z = __shfl_sync(mask, t, v - 1u, W);
            } else {  // empty: the next visited point is the first chain point at or past the voxel boundary
                const float nt = __shfl_sync(mask, s.next_t, v, W);
                const uint32_t above = (v == kW - 1u) ? 0u : ((kWMask << (v + 1u)) & kWMask);
                const uint32_t m = ballot(t >= nt) & above;
                if (m) {
                    v = __ffs(m) - 1;
                } else {  // boundary beyond this chunk: walk the chain like the reference (marching.cu:186-188)
                    float tc = t_next_base;
                    for (uint32_t guard = 0; tc < nt && guard < (1u << 20); ++guard) tc = __fadd_rn(tc, calc_ds(g, tc));
                    t_cur = tc;
                    break;
                }
            }
        }
    }
    if (t_cur >= t_end) {  // far-plane sample, marching.cu:367-394
        const EvalPoint s = eval_point(g, ray, t_end);  // same value on every lane
        if (s.occupied) {
            if (steps > 0 && __fadd_rn(last_ds, last_z) >= t_end) {
                if (lane == 0) fix_ds(steps - 1, __fadd_rn(t_end, -last_z));
            }
            if (steps < cap) {
                if (lane == 0) emit(steps, s, t_end);
                ++steps;
            } else {
                t_cur = t_end;
            }
        }
    }
    return steps;
}

}  // namespace march
}  // namespace ngp
