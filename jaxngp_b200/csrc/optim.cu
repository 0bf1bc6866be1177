// optim.cu -- Adam step of the reference's optimizer (app/nerf/_utils.py:19-77) as one streaming
// kernel over the flat parameter buffer [hash table | MLP weights].
//
//   lr(t)   : optax.exponential_decay(init, transition_steps, decay_rate, transition_begin, staircase, end_value)
//   update  : -lr * m_hat / (sqrt(v_hat + eps_root) + eps)              (optax.adam, b1=.9 b2=.99 eps=eps_root=1e-15)
//   decay   : + weight_decay * p for the MLP weights only, added AFTER the lr scaling with a plus sign,
//             exactly as the reference chains optax.add_decayed_weights behind adam (_utils.py:45-77)
//
// The step counter lives in device memory so the whole training step can sit in one CUDA graph.
#include "common.cuh"

namespace ngp {
namespace {

__global__ void __launch_bounds__(256) adam_kernel(NgpAdamDescriptor d, const uint32_t *__restrict__ step_ptr,
                                                   float *__restrict__ params, const float *__restrict__ grads,
                                                   float *__restrict__ m, float *__restrict__ v) {
    const uint32_t t = __ldg(step_ptr) + 1u;  // optax counts from 1 for the bias correction
    // learning-rate schedule evaluated at count = t - 1 (optax schedules see the pre-increment count)
    float lr = d.lr_init;
    const float count = (float)(t - 1u);
    if (d.transition_steps > 0) {
        float p = fmaxf(count - (float)d.transition_begin, 0.f) / (float)d.transition_steps;
        if (d.staircase) p = floorf(p);
        lr = (count <= (float)d.transition_begin) ? d.lr_init : d.lr_init * powf(d.decay_rate, p);
        lr = d.decay_rate < 1.f ? fmaxf(lr, d.lr_end) : fminf(lr, d.lr_end);
    }
    const float bc1 = 1.f - powf(d.b1, (float)t), bc2 = 1.f - powf(d.b2, (float)t);
    const size_t n4 = d.n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 p = reinterpret_cast<float4 *>(params)[i];
        float4 g = __ldg(reinterpret_cast<const float4 *>(grads) + i);
        float4 mm = reinterpret_cast<float4 *>(m)[i];
        float4 vv = reinterpret_cast<float4 *>(v)[i];
        const float wd = (i * 4 >= d.decay_begin) ? d.weight_decay : 0.f;
        float *pp = &p.x, *gg = &g.x, *pm = &mm.x, *pv = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = gg[k] * d.grad_scale;
            pm[k] = d.b1 * pm[k] + (1.f - d.b1) * gk;
            pv[k] = d.b2 * pv[k] + (1.f - d.b2) * gk * gk;
            const float upd = -lr * (pm[k] / bc1) / (sqrtf(pv[k] / bc2 + d.eps_root) + d.eps);
            pp[k] = pp[k] + (upd + wd * pp[k]);
        }
        reinterpret_cast<float4 *>(params)[i] = p;
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
    }
}

// out = sa * a + sb * b + c on device-resident u32 scalars: the step counter after an optimizer step (a = out = step,
// c = 1) and train_step's `next - exceeded` (marching/__init__.py:91) without a host round trip or a library launch
__global__ void u32_axpy_kernel(NgpU32AxpyDescriptor d, const uint32_t *a, const uint32_t *b, uint32_t *out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *out = (uint32_t)(d.sa * (int32_t)*a + d.sb * (int32_t)*b + d.c);
}

}  // namespace
}  // namespace ngp

extern "C" void ngp_u32_axpy(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpU32AxpyDescriptor>(opaque, opaque_len, "u32_axpy");
    if (!d) return;
    BufferCursor b{buffers};
    const uint32_t *a = b.next<const uint32_t>();
    const uint32_t *bb = b.next<const uint32_t>();
    uint32_t *out = b.next<uint32_t>();
    u32_axpy_kernel<<<1, 32, 0, stream>>>(*d, a, bb, out);
    check_launch("u32_axpy");
}

extern "C" void ngp_adam_step(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len) {
    using namespace ngp;
    clear_error();
    auto *d = descriptor<NgpAdamDescriptor>(opaque, opaque_len, "adam_step");
    if (!d) return;
    if (d->n % 4 != 0 || d->decay_begin % 4 != 0) {
        set_error(NGP_ERR_ARGUMENT, "adam_step: n and decay_begin must be multiples of 4, got %llu, %llu",
                  (unsigned long long)d->n, (unsigned long long)d->decay_begin);
        return;
    }
    BufferCursor b{buffers};
    const uint32_t *step = b.next<const uint32_t>();
    float *params = b.next<float>();
    const float *grads = b.next<const float>();
    float *m = b.next<float>();
    float *v = b.next<float>();
    if (d->n == 0) return;
    const unsigned blocks = min(div_up(d->n / 4, 256u), 148u * 8u);
    adam_kernel<<<blocks, 256, 0, stream>>>(*d, step, params, grads, m, v);
    check_launch("adam_step");
}
