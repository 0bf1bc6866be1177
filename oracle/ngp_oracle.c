/*
 * ngp_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the jaxngp hot path, in plain C.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (jaxngp_b200/) never does and has no CPU fallback.
 *
 * Parity status:
 *   - morton / packbits / march / integrate: PINNED against the reference's own CUDA ops
 *     (oracle/_ref/libvolrend_ref.so, built unmodified from /root/reference by oracle/build_ref.sh)
 *     run on a B200; the captured outputs are committed under tests/golden/ (made by
 *     oracle/make_golden.py) and tests/test_oracle_golden.py checks this file against them.
 *   - hash-grid encoder (pure-JAX HashGridEncoder, models/encoders.py): PINNED against the reference's own
 *     code -- jax is absent, so oracle/ref_shim.py executes the unmodified HashGridEncoder.__call__ on numpy
 *     stand-ins; oracle/make_golden_encoder.py committed its outputs (tests/golden/encoder_reference.npz) and
 *     tests/test_oracle_golden.py requires this file to reproduce them exactly.  The tiny-cuda-nn variant
 *     (jaxtcnn.hashgrid_encode) stays "parity unpinned": tiny-cuda-nn v1.6 is not on disk.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC  (oracle/Makefile).
 * -ffp-contract=off matters: every fused multiply-add below is an explicit fmaf() placed where
 * nvcc contracts the reference source (verified in the reference SASS).
 *
 * All citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define T_THRESHOLD 1e-4f /* deps/volume-rendering-jax/lib/impl/integrating.cu:12 */
#define TWO_SQRT3 3.4641015529632568359f /* 2 * (float)SQRT3, volrend.h:18, marching.cu:20-21 */

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------- morton (marching.cu:52-77) */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t compact_bits(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

/* marching.cu:399-414 */
void orc_morton3d(uint32_t length, const uint32_t *xyzs, uint32_t *idcs) {
    for (uint32_t i = 0; i < length; ++i)
        idcs[i] = morton3(xyzs[i * 3 + 0], xyzs[i * 3 + 1], xyzs[i * 3 + 2]);
}

/* marching.cu:416-433 */
void orc_morton3d_invert(uint32_t length, const uint32_t *idcs, uint32_t *xyzs) {
    for (uint32_t i = 0; i < length; ++i) {
        xyzs[i * 3 + 0] = compact_bits(idcs[i] >> 0);
        xyzs[i * 3 + 1] = compact_bits(idcs[i] >> 1);
        xyzs[i * 3 + 2] = compact_bits(idcs[i] >> 2);
    }
}

/* ---------------------------------------------------------------- packbits (packbits.cu:9-34) */
void orc_packbits(uint32_t n_bytes, const float *threshold, const float *density,
                  uint8_t *occupied_mask, uint8_t *bitfield) {
    for (uint32_t i = 0; i < n_bytes; ++i) {
        uint8_t byte = 0;
        for (uint32_t k = 0; k < 8; ++k) {
            int p = density[i * 8 + k] > threshold[i * 8 + k];
            occupied_mask[i * 8 + k] = (uint8_t)p;
            byte |= (uint8_t)(p << k);
        }
        bitfield[i] = byte;
    }
}

/* ---------------------------------------------------------------- marching */
typedef struct {
    uint32_t K, G, G3;
    float Gf, inv_G, bound, portion, ds_lo, ds_hi;
    const uint8_t *bits;
} grid_t;

static void grid_init(grid_t *g, uint32_t steps, uint32_t K, uint32_t G, float bound,
                      float portion, const uint8_t *bits) {
    g->K = K;
    g->G = G;
    g->G3 = G * G * G;
    g->Gf = (float)G;
    g->inv_G = 1.f / (float)G; /* marching.cu:158 */
    g->bound = bound;
    g->portion = portion;
    /* calc_ds bounds, marching.cu:15-23 */
    g->ds_lo = (TWO_SQRT3 * fminf(bound, 1.f)) / (float)steps;
    g->ds_hi = (TWO_SQRT3 * bound) * g->inv_G;
    g->bits = bits;
}

static inline float calc_ds(const grid_t *g, float t) {
    return fminf(fmaxf(t * g->portion, g->ds_lo), g->ds_hi);
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* float -> int with round-toward-minus-infinity and saturation, like F2I.FLOOR */
static inline int floor_to_int(float x) {
    float f = floorf(x);
    if (!(f == f)) return 0;
    if (f >= 2147483648.f) return 2147483647;
    if (f <= -2147483648.f) return (-2147483647 - 1);
    return (int)f;
}

static inline uint32_t mip_of(float v, uint32_t K) { /* marching.cu:25-39 */
    int e;
    frexpf(v, &e);
    return (uint32_t)clampi(e, 0, (int)K - 1);
}

/* One marching step at parameter t (marching.cu:166-190).  Returns occupancy, writes the sample
 * position, its ds, and the next t to visit. */
static inline int march_step(const grid_t *g, const float *o, const float *d, const float *inv_d,
                             float t, float *pos, float *ds_out, float *t_next) {
    pos[0] = fmaf(t, d[0], o[0]);
    pos[1] = fmaf(t, d[1], o[1]);
    pos[2] = fmaf(t, d[2], o[2]);
    float ds = calc_ds(g, t);
    uint32_t cascade = 0;
    if (g->K > 1) { /* marching.cu:79-98 */
        float linf = fmaxf(fabsf(pos[0]), fmaxf(fabsf(pos[1]), fabsf(pos[2])));
        uint32_t a = mip_of(linf, g->K), b = mip_of(ds * g->Gf, g->K);
        cascade = a > b ? a : b;
    }
    float mip_bound = fminf((float)(1u << cascade), g->bound);
    float gp[3];
    int gi[3];
    for (int k = 0; k < 3; ++k) {
        gp[k] = ((pos[k] / mip_bound + 1.f) * .5f) * g->Gf;
        gi[k] = clampi(floor_to_int(gp[k]), 0, (int)g->G - 1);
    }
    uint32_t idx = cascade * g->G3 + morton3((uint32_t)gi[0], (uint32_t)gi[1], (uint32_t)gi[2]);
    int occupied = (g->bits[idx >> 3] >> (idx & 7u)) & 1;
    float tn = t + ds;
    if (!occupied) {
        float del[3];
        for (int k = 0; k < 3; ++k) {
            float ng = floorf(fmaf(copysignf(1.f, d[k]), .5f, gp[k] + .5f));
            float a = fmaf(ng, g->inv_G, -.5f);
            a = a + a;
            del[k] = fmaf(mip_bound, a, -pos[k]) * inv_d[k];
        }
        float next_t = t + fmaxf(0.f, fminf(del[0], fminf(del[1], del[2])));
        while (tn < next_t) tn += calc_ds(g, tn);
    }
    *ds_out = ds;
    *t_next = tn;
    return occupied;
}

/*
 * march_rays (marching.cu:101-268), with the sample compaction done in RAY ORDER: this is the
 * reference kernel executed with thread i's atomicAdd arriving i-th (the reference leaves the
 * arrival order to the hardware).  Rays after the ray at which the running counter
 * reaches total_samples early-out (marching.cu:135) and stay invalid.
 */
void orc_march_rays(uint32_t n_rays, uint32_t total_samples, uint32_t steps, uint32_t K, uint32_t G,
                    float bound, float portion, const float *rays_o, const float *rays_d,
                    const float *t_starts, const float *t_ends, const float *noises,
                    const uint8_t *bits, uint32_t *next_loc, uint32_t *exceeded, uint8_t *valid,
                    uint32_t *rays_n, uint32_t *rays_start, uint32_t *idcs, float *xyzs,
                    float *dirs, float *dss, float *z_vals) {
    grid_t g;
    grid_init(&g, steps, K, G, bound, portion, bits);
    memset(valid, 0, n_rays);
    memset(rays_n, 0, n_rays * sizeof(uint32_t));
    memset(rays_start, 0, n_rays * sizeof(uint32_t));
    memset(idcs, 0, total_samples * sizeof(uint32_t));
    memset(xyzs, 0, (size_t)total_samples * 3 * sizeof(float));
    memset(dirs, 0, (size_t)total_samples * 3 * sizeof(float));
    memset(dss, 0, total_samples * sizeof(float));
    memset(z_vals, 0, total_samples * sizeof(float));
    uint32_t *cnt = (uint32_t *)calloc(n_rays ? n_rays : 1, sizeof(uint32_t));
    uint8_t *hit_box = (uint8_t *)calloc(n_rays ? n_rays : 1, 1);
    float max_steps = (float)steps * bound; /* marching.cu:165 */

    /* pass 1: count (independent per ray) */
#pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t i = 0; i < n_rays; ++i) {
        const float *o = rays_o + 3 * i, *d = rays_d + 3 * i;
        float ts = t_starts[i], te = t_ends[i];
        if (te <= ts) continue; /* marching.cu:151 */
        hit_box[i] = 1;
        float inv_d[3] = {1.f / d[0], 1.f / d[1], 1.f / d[2]};
        float t = fmaf(calc_ds(&g, ts), noises[i], ts);
        uint32_t n = 0;
        float pos[3], ds, tn;
        while ((float)n < max_steps && t < te) {
            if (march_step(&g, o, d, inv_d, t, pos, &ds, &tn)) ++n;
            t = tn;
        }
        cnt[i] = n;
    }
    /* compaction in ray order */
    uint32_t counter = 0, exc = 0;
    for (uint32_t i = 0; i < n_rays; ++i) {
        if (counter >= total_samples) break; /* marching.cu:135 */
        if (!hit_box[i]) continue;
        if (cnt[i] == 0) { valid[i] = 1; continue; } /* marching.cu:196-199 */
        uint32_t start = counter;
        counter += cnt[i]; /* marching.cu:205 */
        if (start + cnt[i] > total_samples) { exc += cnt[i]; cnt[i] = 0; continue; }
        rays_n[i] = cnt[i];
        rays_start[i] = start;
        valid[i] = 1;
    }
    *next_loc = counter;
    *exceeded = exc;
    /* pass 2: write */
#pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t i = 0; i < n_rays; ++i) {
        uint32_t n = rays_n[i];
        if (!n) continue;
        const float *o = rays_o + 3 * i, *d = rays_d + 3 * i;
        float ts = t_starts[i], te = t_ends[i];
        float inv_d[3] = {1.f / d[0], 1.f / d[1], 1.f / d[2]};
        float t = fmaf(calc_ds(&g, ts), noises[i], ts);
        uint32_t s = 0, base = rays_start[i];
        float pos[3], ds, tn;
        while (s < n && t < te) {
            if (march_step(&g, o, d, inv_d, t, pos, &ds, &tn)) {
                uint32_t w = base + s;
                idcs[w] = i;
                xyzs[w * 3 + 0] = pos[0]; xyzs[w * 3 + 1] = pos[1]; xyzs[w * 3 + 2] = pos[2];
                dirs[w * 3 + 0] = d[0]; dirs[w * 3 + 1] = d[1]; dirs[w * 3 + 2] = d[2];
                dss[w] = ds;
                z_vals[w] = t;
                ++s;
            }
            t = tn;
        }
    }
    free(cnt);
    free(hit_box);
}

/*
 * march_rays_inference (marching.cu:271-397).  Slots whose `terminated` flag is set take fresh
 * rays in SLOT ORDER (the reference hands them out in atomic arrival order, marching.cu:300).
 */
void orc_march_rays_inference(uint32_t n_total_rays, uint32_t n_rays, uint32_t steps, uint32_t K,
                              uint32_t G, uint32_t cap, float bound, float portion,
                              const float *rays_o, const float *rays_d, const float *t_starts,
                              const float *t_ends, const uint8_t *bits,
                              const uint32_t *next_ray_index_in, const uint8_t *terminated,
                              const uint32_t *indices_in, uint32_t *next_ray_index,
                              uint32_t *indices_out, uint32_t *n_samples, float *t_starts_out,
                              float *xyzs, float *dss, float *z_vals) {
    grid_t g;
    grid_init(&g, steps, K, G, bound, portion, bits);
    memset(n_samples, 0, n_rays * sizeof(uint32_t));
    memset(t_starts_out, 0, n_rays * sizeof(float));
    memset(xyzs, 0, (size_t)n_rays * cap * 3 * sizeof(float));
    memset(dss, 0, (size_t)n_rays * cap * sizeof(float));
    memset(z_vals, 0, (size_t)n_rays * cap * sizeof(float));
    uint32_t counter = *next_ray_index_in;
    for (uint32_t i = 0; i < n_rays; ++i) indices_out[i] = terminated[i] ? counter++ : indices_in[i];
    *next_ray_index = counter;
#pragma omp parallel for schedule(dynamic, 64)
    for (uint32_t i = 0; i < n_rays; ++i) {
        uint32_t r = indices_out[i];
        if (r >= n_total_rays) continue;
        const float *o = rays_o + 3 * (size_t)r, *d = rays_d + 3 * (size_t)r;
        float ts = t_starts[r], te = t_ends[r];
        if (te < ts) continue; /* marching.cu:317, strict */
        float inv_d[3] = {1.f / d[0], 1.f / d[1], 1.f / d[2]};
        float *rx = xyzs + (size_t)i * cap * 3, *rds = dss + (size_t)i * cap, *rz = z_vals + (size_t)i * cap;
        uint32_t s = 0;
        float t = ts, pos[3], ds, tn;
        while (s < cap && t < te) {
            if (march_step(&g, o, d, inv_d, t, pos, &ds, &tn)) {
                rx[s * 3 + 0] = pos[0]; rx[s * 3 + 1] = pos[1]; rx[s * 3 + 2] = pos[2];
                rds[s] = ds;
                rz[s] = t;
                ++s;
            }
            t = tn;
        }
        if (t >= te) { /* far-plane sample, marching.cu:367-394 */
            float tn2;
            /* occupancy at t_end; the skip part of march_step is irrelevant here */
            int occ = march_step(&g, o, d, inv_d, te, pos, &ds, &tn2);
            if (occ) {
                if (s > 0 && rds[s - 1] + rz[s - 1] >= te) rds[s - 1] = te - rz[s - 1];
                if (s < cap) {
                    rx[s * 3 + 0] = pos[0]; rx[s * 3 + 1] = pos[1]; rx[s * 3 + 2] = pos[2];
                    rds[s] = ds;
                    rz[s] = te;
                    ++s;
                } else {
                    t = te;
                }
            }
        }
        n_samples[i] = s;
        t_starts_out[i] = t;
    }
}

/* ---------------------------------------------------------------- integrating */
/* integrating.cu:24-104 */
void orc_integrate_rays(uint32_t n_rays, const uint32_t *start, const uint32_t *nsamp,
                        const float *bgs, const float *dss, const float *z_vals, const float *drgbs,
                        uint32_t *measured_batch_size, float *final_rgbds, float *final_opacities) {
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : total)
    for (uint32_t i = 0; i < n_rays; ++i) {
        const float *rds = dss + start[i], *rz = z_vals + start[i], *rc = drgbs + (size_t)start[i] * 4;
        uint32_t n = nsamp[i], s = 0;
        float depth = 0.f, T = 1.f, r = 0.f, g = 0.f, b = 0.f;
        for (; T > T_THRESHOLD && s < n; ++s) {
            float alpha = 1.f - expf(-rc[s * 4] * rds[s]);
            float w = T * alpha;
            r = fmaf(w, rc[s * 4 + 1], r);
            g = fmaf(w, rc[s * 4 + 2], g);
            b = fmaf(w, rc[s * 4 + 3], b);
            depth = fmaf(w, rz[s], depth);
            T *= 1.f - alpha;
        }
        float opacity = 1.f - T;
        final_opacities[i] = opacity;
        if (T <= T_THRESHOLD) { /* integrating.cu:85-90 */
            float id = 1.f / opacity;
            final_rgbds[i * 4 + 0] = r * id; final_rgbds[i * 4 + 1] = g * id;
            final_rgbds[i * 4 + 2] = b * id; final_rgbds[i * 4 + 3] = depth * id;
        } else { /* integrating.cu:91-96 */
            final_rgbds[i * 4 + 0] = fmaf(T, bgs[i * 3 + 0], r);
            final_rgbds[i * 4 + 1] = fmaf(T, bgs[i * 3 + 1], g);
            final_rgbds[i * 4 + 2] = fmaf(T, bgs[i * 3 + 2], b);
            final_rgbds[i * 4 + 3] = depth;
        }
        total += s;
    }
    *measured_batch_size = (uint32_t)total;
}

/* integrating.cu:106-240 */
void orc_integrate_rays_backward(uint32_t n_rays, uint32_t total_samples, float near_distance,
                                 const uint32_t *start, const uint32_t *nsamp, const float *bgs,
                                 const float *dss, const float *z_vals, const float *drgbs,
                                 const float *final_rgbds, const float *final_opacities,
                                 const float *dL_dfinal_rgbds, float *dL_dbgs, float *dL_dz_vals,
                                 float *dL_ddrgbs) {
    memset(dL_dbgs, 0, (size_t)n_rays * 3 * sizeof(float));
    memset(dL_dz_vals, 0, (size_t)total_samples * sizeof(float));
    memset(dL_ddrgbs, 0, (size_t)total_samples * 4 * sizeof(float));
#pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t i = 0; i < n_rays; ++i) {
        const float *rds = dss + start[i], *rz = z_vals + start[i], *rc = drgbs + (size_t)start[i] * 4;
        const float *fin = final_rgbds + i * 4, *dfin = dL_dfinal_rgbds + i * 4, *bg = bgs + i * 3;
        float *gz = dL_dz_vals + start[i], *gc = dL_ddrgbs + (size_t)start[i] * 4;
        float opac = final_opacities[i];
        int terminated = opac >= 1.f - T_THRESHOLD; /* integrating.cu:158 */
        float bgw = terminated ? 0.f : 1.f - opac;
        uint32_t n = nsamp[i];
        float T = 1.f, cur[3] = {0.f, 0.f, 0.f}, cur_depth = 0.f;
        for (uint32_t s = 0; T > T_THRESHOLD && s < n; ++s) {
            float z = rz[s], dt = rds[s], density = rc[s * 4];
            float alpha = 1.f - expf(-density * dt);
            float w = T * alpha;
            cur[0] += w * rc[s * 4 + 1]; cur[1] += w * rc[s * 4 + 2]; cur[2] += w * rc[s * 4 + 3];
            cur_depth += w * z;
            T *= 1.f - alpha;
            gz[s] = w * dfin[3]; /* integrating.cu:196 */
            float acc = 0.f;
            for (int k = 0; k < 3; ++k)
                acc += dfin[k] * (T * rc[s * 4 + 1 + k] - (fin[k] - cur[k]) - bg[k] * bgw);
            acc += dfin[3] * (T * z - (fin[3] - cur_depth));
            float dsig = dt * acc;
            float reg = (density > 4e-5 && z < near_distance) ? 1e-4f : 0.f; /* :221 */
            float scal = fminf(z * z, 1.f);                                     /* :225 */
            gc[s * 4 + 0] = scal * dsig + reg;
            gc[s * 4 + 1] = w * dfin[0]; gc[s * 4 + 2] = w * dfin[1]; gc[s * 4 + 3] = w * dfin[2];
        }
        if (T > T_THRESHOLD) { /* integrating.cu:235-239 */
            dL_dbgs[i * 3 + 0] = T * dfin[0]; dL_dbgs[i * 3 + 1] = T * dfin[1];
            dL_dbgs[i * 3 + 2] = T * dfin[2];
        }
    }
}

/* integrating.cu:242-322 */
void orc_integrate_rays_inference(uint32_t n_total_rays, uint32_t n_rays, uint32_t cap,
                                  const float *rays_bg, const float *rays_rgbd, const float *rays_T,
                                  const uint32_t *n_samples, const uint32_t *indices,
                                  const float *dss, const float *z_vals, const float *drgbs,
                                  uint32_t *terminate_cnt, uint8_t *terminated, float *rgbd_out,
                                  float *T_out) {
    memset(terminated, 0, n_rays);
    memset(rgbd_out, 0, (size_t)n_rays * 4 * sizeof(float));
    memset(T_out, 0, (size_t)n_rays * sizeof(float));
    uint32_t cnt = 0;
    for (uint32_t i = 0; i < n_rays; ++i) {
        uint32_t ns = n_samples[i], r = indices[i];
        if (r >= n_total_rays) continue;
        const float *rds = dss + (size_t)i * cap, *rz = z_vals + (size_t)i * cap, *rc = drgbs + (size_t)i * cap * 4;
        float T = rays_T[r], cr = rays_rgbd[r * 4 + 0], cg = rays_rgbd[r * 4 + 1],
              cb = rays_rgbd[r * 4 + 2], depth = rays_rgbd[r * 4 + 3];
        for (uint32_t s = 0; T > T_THRESHOLD && s < ns; ++s) {
            float alpha = 1.f - expf(-rc[s * 4] * rds[s]);
            float w = T * alpha;
            cr = fmaf(w, rc[s * 4 + 1], cr); cg = fmaf(w, rc[s * 4 + 2], cg);
            cb = fmaf(w, rc[s * 4 + 3], cb); depth = fmaf(w, rz[s], depth);
            T *= 1.f - alpha;
        }
        if (T <= T_THRESHOLD) {
            float id = 1.f / (1.f - T);
            terminated[i] = 1;
            T_out[i] = 0.f;
            rgbd_out[i * 4 + 0] = cr * id; rgbd_out[i * 4 + 1] = cg * id;
            rgbd_out[i * 4 + 2] = cb * id; rgbd_out[i * 4 + 3] = depth * id;
        } else {
            terminated[i] = ns < cap;
            rgbd_out[i * 4 + 3] = depth;
            T_out[i] = T;
            if (terminated[i]) {
                rgbd_out[i * 4 + 0] = fmaf(T, rays_bg[r * 3 + 0], cr);
                rgbd_out[i * 4 + 1] = fmaf(T, rays_bg[r * 3 + 1], cg);
                rgbd_out[i * 4 + 2] = fmaf(T, rays_bg[r * 3 + 2], cb);
            } else {
                rgbd_out[i * 4 + 0] = cr; rgbd_out[i * 4 + 1] = cg; rgbd_out[i * 4 + 2] = cb;
            }
        }
        cnt += terminated[i];
    }
    *terminate_cnt = cnt;
}

/* ---------------------------------------------------------------- hash-grid encoder
 * models/encoders.py:82-256 (HashGridEncoder.__call__), dim in {2,3}.
 * Level metadata (scale f32, res u32, offset u32, hashed flag) is computed by the caller exactly
 * as encoders.py:89-103 does (double precision on the host, then cast).  `wrap` = T reproduces the
 * reference's `mod T` on every level (encoders.py:187); tcnn-style wrapping passes
 * wrap = 0 meaning "level size".
 */
static const uint32_t PRIMES[3] = {1u, 2654435761u, 805459861u}; /* encoders.py:169 */

static inline uint32_t hg_index(uint32_t dim, const uint32_t *v, uint32_t res, int hashed,
                                uint32_t wrap, uint32_t offset) {
    uint32_t idx;
    if (hashed) { /* encoders.py:157-177 */
        idx = v[0] ^ (v[1] * PRIMES[1]);
        if (dim == 3) idx ^= v[2] * PRIMES[2];
    } else { /* encoders.py:134-155, uint32 wrap-around arithmetic */
        idx = v[0] + v[1] * res;
        if (dim == 3) idx += v[2] * res * res;
    }
    return idx % wrap + offset; /* encoders.py:187-188 */
}

/* forward: enc[n, L*F] */
void orc_hashgrid_encode(uint32_t n, uint32_t dim, uint32_t L, uint32_t F, const float *scales,
                         const uint32_t *res, const uint32_t *offsets, const uint8_t *hashed,
                         uint32_t wrap_T, float bound, const float *pos, const float *table,
                         float *enc) {
    uint32_t nc = 1u << dim;
#pragma omp parallel for schedule(static)
    for (uint32_t p = 0; p < n; ++p) {
        float p01[3];
        for (uint32_t k = 0; k < dim; ++k) p01[k] = (pos[p * dim + k] + bound) / (2 * bound); /* :87 */
        for (uint32_t l = 0; l < L; ++l) {
            uint32_t wrap = wrap_T ? wrap_T : (offsets[l + 1] - offsets[l]);
            float fr[3];
            uint32_t base[3];
            for (uint32_t k = 0; k < dim; ++k) {
                float ps = p01[k] * scales[l] + 0.5f; /* :218, mul then add (contraction off) */
                float fl = floorf(ps);
                base[k] = (uint32_t)(int32_t)fl; /* :116-123 */
                fr[k] = ps - fl;                 /* jnp.modf, :204 */
            }
            float acc[8] = {0};
            for (uint32_t c = 0; c < nc; ++c) {
                uint32_t v[3] = {0, 0, 0};
                float w = 1.f;
                for (uint32_t k = 0; k < dim; ++k) {
                    /* corner order: last axis fastest (encoders.py:16-33) */
                    uint32_t bit = (c >> (dim - 1 - k)) & 1u;
                    v[k] = base[k] + bit;
                    float wk = bit ? fr[k] : 1.f - fr[k]; /* :204-213 */
                    wk = fminf(fmaxf(wk, 0.f), 1.f);
                    w *= wk;
                }
                uint32_t row = hg_index(dim, v, res[l], hashed[l], wrap, offsets[l]);
                /* latents[indices] (encoders.py:226): XLA clamps an out-of-range gather index (last level dense) */
                if (row >= offsets[L]) row = offsets[L] - 1;
                for (uint32_t f = 0; f < F; ++f) acc[f] += w * table[(size_t)row * F + f];
            }
            for (uint32_t f = 0; f < F; ++f) enc[(size_t)p * L * F + l * F + f] = acc[f]; /* :231-233 */
        }
    }
}

/* backward: d_table[rows, F] (double accumulation) = sum over points/corners of w * d_enc */
void orc_hashgrid_backward(uint32_t n, uint32_t dim, uint32_t L, uint32_t F, const float *scales,
                           const uint32_t *res, const uint32_t *offsets, const uint8_t *hashed,
                           uint32_t wrap_T, float bound, const float *pos, const float *d_enc,
                           double *d_table) {
    uint32_t nc = 1u << dim;
    memset(d_table, 0, (size_t)offsets[L] * F * sizeof(double));
    /* parallel over points with atomic adds (rows are shared between points, and dense levels
     * spill into their successor's rows under wrap_T, Q1) */
    for (uint32_t l = 0; l < L; ++l) {
        uint32_t wrap = wrap_T ? wrap_T : (offsets[l + 1] - offsets[l]);
#pragma omp parallel for schedule(static)
        for (uint32_t p = 0; p < n; ++p) {
            float fr[3];
            uint32_t base[3];
            for (uint32_t k = 0; k < dim; ++k) {
                float p01 = (pos[p * dim + k] + bound) / (2 * bound);
                float ps = p01 * scales[l] + 0.5f;
                float fl = floorf(ps);
                base[k] = (uint32_t)(int32_t)fl;
                fr[k] = ps - fl;
            }
            for (uint32_t c = 0; c < nc; ++c) {
                uint32_t v[3] = {0, 0, 0};
                float w = 1.f;
                for (uint32_t k = 0; k < dim; ++k) {
                    uint32_t bit = (c >> (dim - 1 - k)) & 1u;
                    v[k] = base[k] + bit;
                    float wk = bit ? fr[k] : 1.f - fr[k];
                    wk = fminf(fmaxf(wk, 0.f), 1.f);
                    w *= wk;
                }
                uint32_t row = hg_index(dim, v, res[l], hashed[l], wrap, offsets[l]);
                if (row >= offsets[L]) continue; /* ... and its transposed scatter-add drops the update */
                for (uint32_t f = 0; f < F; ++f) {
                    double upd = (double)w * (double)d_enc[(size_t)p * L * F + l * F + f];
#pragma omp atomic
                    d_table[(size_t)row * F + f] += upd;
                }
            }
        }
    }
}
