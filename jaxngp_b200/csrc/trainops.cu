// trainops.cu -- the glue of the reference's train_step around the four ops, as two kernels:
//   * ngp_make_training_rays: pixel index -> camera ray -> world ray -> AABB near/far
//       (app/nerf/_utils.py:93-115, utils/types.py:398-439 for an undistorted PERSPECTIVE camera,
//        models/renderers/cuda.py:57-97)
//   * ngp_huber_loss_grad: ground-truth fetch + alpha blend + Huber(delta) over valid rays and its
//       gradient w.r.t. the rendered colours (app/nerf/_utils.py:151-165, utils/data.py:443-464)
// XLA fuses these elementwise chains for the reference; a torch host would spend ~40 launches on them.
#include "common.cuh"

namespace ngp {
namespace {

constexpr int kBlock = 256;

// kRng: the step's random inputs leave the same kernel -- noises[i] (march perturbation, cuda.py:118-122) and
// bgs[i] (random background of the loss, _utils.py:134-136) = the four Philox uniforms of element i of this call.
template <bool kRng>
__global__ void __launch_bounds__(kBlock) make_training_rays_kernel(NgpTrainingRaysDescriptor d, NgpRngDescriptor rng,
                                                                     const int32_t *__restrict__ perm,
                                                                     const float *__restrict__ transforms,
                                                                     uint32_t *__restrict__ rng_state,
                                                                     float *__restrict__ rays_o, float *__restrict__ rays_d,
                                                                     float *__restrict__ t_starts, float *__restrict__ t_ends,
                                                                     float *__restrict__ noises, float *__restrict__ bgs) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (kRng) {
        const uint32_t counter = rng_state[0];
        if (i < d.n_rays) {
            const float4 u = philox_uniform4(i, counter, rng.stream_id, rng.seed_lo, rng.seed_hi);
            noises[i] = u.x;
            bgs[3 * (size_t)i + 0] = u.y;
            bgs[3 * (size_t)i + 1] = u.z;
            bgs[3 * (size_t)i + 2] = u.w;
        }
        rng_state_finish(rng_state);  // every thread of the block reaches this (no early return above)
    }
    if (i >= d.n_rays) return;
    const uint32_t hw = d.width * d.height;
    const uint32_t p = (uint32_t)__ldg(perm + i);
    const uint32_t view = min(p / hw, d.n_views - 1u), pix = p % hw;
    const uint32_t x = pix % d.width, y = pix / d.width;
    // utils/types.py:413-417,434-439: pixel centre, CV -> CG flip, normalise
    const float cx = __fdiv_rn(__fadd_rn(__fadd_rn((float)x, .5f), -d.cx), d.fx);
    const float cy = -__fdiv_rn(__fadd_rn(__fadd_rn((float)y, .5f), -d.cy), d.fy);
    const float cz = -1.f;
    const float inv_norm = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), 1.f)));
    const float dc[3] = {cx * inv_norm, cy * inv_norm, cz * inv_norm};
    const float *tf = transforms + (size_t)view * 12;
    float o[3], dw[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {  // _utils.py:107-113: d_world[r] = sum_c d_cam[c] * R[r][c]
        dw[r] = __fadd_rn(__fadd_rn(__fmul_rn(dc[0], __ldg(tf + 3 * r + 0)), __fmul_rn(dc[1], __ldg(tf + 3 * r + 1))),
                          __fmul_rn(dc[2], __ldg(tf + 3 * r + 2)));
        o[r] = __ldg(tf + 9 + r);
    }
    // models/renderers/cuda.py:57-97
    float t_start = -INFINITY, t_end = INFINITY;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float eps = 1e-15f;
        const float dd = signbit(dw[r]) ? fminf(dw[r], -eps) : fmaxf(dw[r], eps);
        const float t0 = __fdiv_rn(__fadd_rn(-d.bound, -o[r]), dd), t1 = __fdiv_rn(__fadd_rn(d.bound, -o[r]), dd);
        t_start = fmaxf(t_start, fminf(t0, t1));
        t_end = fminf(t_end, fmaxf(t0, t1));
        rays_o[3 * (size_t)i + r] = o[r];
        rays_d[3 * (size_t)i + r] = dw[r];
    }
    t_starts[i] = fmaxf(t_start, 0.f);
    t_ends[i] = t_end;
}

__global__ void __launch_bounds__(kBlock) philox_uniform_kernel(NgpPhiloxDescriptor d, float4 *__restrict__ out) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i < d.n) out[i] = philox_uniform4(i, d.counter, d.rng.stream_id, d.rng.seed_lo, d.rng.seed_hi);
}

__global__ void __launch_bounds__(kBlock) count_valid_kernel(uint32_t n, const uint8_t *__restrict__ valid,
                                                              uint32_t *__restrict__ n_valid) {
    uint32_t c = 0;
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) c += valid[i] ? 1u : 0u;
    c = warp_sum_u32(c);
    if ((threadIdx.x & 31u) == 0 && c) atomicAdd(n_valid, c);
}

__global__ void __launch_bounds__(kBlock) huber_loss_grad_kernel(NgpHuberLossDescriptor d, const float4 *__restrict__ final_rgbds,
                                                                  const uint8_t *__restrict__ valid,
                                                                  const int32_t *__restrict__ perm,
                                                                  const uchar4 *__restrict__ rgbas,
                                                                  const float *__restrict__ bgs,
                                                                  const uint32_t *__restrict__ n_valid_ptr,
                                                                  float4 *__restrict__ dL_dfinal, float *__restrict__ loss) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    const uint32_t n_valid = __ldg(n_valid_ptr);
    const float inv_n = 1.f / (float)n_valid;  // n_valid == 0 -> the reference's loss is 0/0 as well
    float per_ray = 0.f;
    if (i < d.n_rays) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid[i]) {
            const float4 pred = __ldg(final_rgbds + i);
            const uchar4 px = __ldg(rgbas + (uint32_t)__ldg(perm + i));
            g = huber_ray(pred, px, __ldg(bgs + 3 * (size_t)i + 0), __ldg(bgs + 3 * (size_t)i + 1), __ldg(bgs + 3 * (size_t)i + 2),
                          d.delta, inv_n, per_ray);
        }
        dL_dfinal[i] = g;
    }
    per_ray = warp_sum(per_ray);
    if ((threadIdx.x & 31u) == 0 && per_ray != 0.f) atomicAdd(loss, per_ray * inv_n);
}

}  // namespace
}  // namespace ngp

extern "C" {

void ngp_make_training_rays(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpTrainingRaysDescriptor>(opaque, opaque_len, "make_training_rays");
    if (!d || d->n_rays == 0) return;
    if (d->width == 0 || d->height == 0 || d->n_views == 0) {
        set_error(NGP_ERR_ARGUMENT, "make_training_rays: empty image or view set");
        return;
    }
    BufferCursor b{buffers};
    const int32_t *perm = b.next<const int32_t>();
    const float *transforms = b.next<const float>();
    float *rays_o = b.next<float>();
    float *rays_d = b.next<float>();
    float *t_starts = b.next<float>();
    float *t_ends = b.next<float>();
    make_training_rays_kernel<false><<<div_up(d->n_rays, kBlock), kBlock, 0, stream>>>(*d, NgpRngDescriptor{}, perm, transforms, nullptr, rays_o,
                                                                                       rays_d, t_starts, t_ends, nullptr, nullptr);
    check_launch("make_training_rays");
}

void ngp_make_training_rays_rng(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpTrainingRaysRngDescriptor>(opaque, opaque_len, "make_training_rays_rng");
    if (!d || d->rays.n_rays == 0) return;
    if (d->rays.width == 0 || d->rays.height == 0 || d->rays.n_views == 0) {
        set_error(NGP_ERR_ARGUMENT, "make_training_rays_rng: empty image or view set");
        return;
    }
    BufferCursor b{buffers};
    const int32_t *perm = b.next<const int32_t>();
    const float *transforms = b.next<const float>();
    uint32_t *rng_state = b.next<uint32_t>();
    float *rays_o = b.next<float>();
    float *rays_d = b.next<float>();
    float *t_starts = b.next<float>();
    float *t_ends = b.next<float>();
    float *noises = b.next<float>();
    float *bgs = b.next<float>();
    make_training_rays_kernel<true><<<div_up(d->rays.n_rays, kBlock), kBlock, 0, stream>>>(d->rays, d->rng, perm, transforms, rng_state, rays_o,
                                                                                            rays_d, t_starts, t_ends, noises, bgs);
    check_launch("make_training_rays_rng");
}

void ngp_philox_uniform(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpPhiloxDescriptor>(opaque, opaque_len, "philox_uniform");
    if (!d || d->n == 0) return;
    BufferCursor b{buffers};
    float4 *out = b.next<float4>();
    philox_uniform_kernel<<<div_up(d->n, kBlock), kBlock, 0, stream>>>(*d, out);
    check_launch("philox_uniform");
}

void ngp_huber_loss_grad(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpHuberLossDescriptor>(opaque, opaque_len, "huber_loss_grad");
    if (!d) return;
    BufferCursor b{buffers};
    const float4 *final_rgbds = b.next<const float4>();
    const uint8_t *valid = b.next<const uint8_t>();
    const int32_t *perm = b.next<const int32_t>();
    const uchar4 *rgbas = b.next<const uchar4>();
    const float *bgs = b.next<const float>();
    float4 *dL = b.next<float4>();
    float *loss = b.next<float>();
    uint32_t *n_valid = b.next<uint32_t>();
    NGP_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float), stream), "huber_loss_grad");
    NGP_CUDA_OK(cudaMemsetAsync(n_valid, 0, sizeof(uint32_t), stream), "huber_loss_grad");
    if (d->n_rays == 0) return;
    count_valid_kernel<<<min(div_up(d->n_rays, kBlock), 148u * 4u), kBlock, 0, stream>>>(d->n_rays, valid, n_valid);
    huber_loss_grad_kernel<<<div_up(d->n_rays, kBlock), kBlock, 0, stream>>>(*d, final_rgbds, valid, perm, rgbas, bgs, n_valid, dL, loss);
    check_launch("huber_loss_grad");
}

}  // extern "C"
