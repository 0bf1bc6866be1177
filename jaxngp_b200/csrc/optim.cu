// optim.cu -- Adam step of the reference's optimizer (app/nerf/_utils.py:19-77) as one streaming
// kernel over the flat parameter buffer [hash table | MLP weights].
//
//   lr(t)   : optax.exponential_decay(init, transition_steps, decay_rate, transition_begin, staircase, end_value)
//   update  : -lr * m_hat / (sqrt(v_hat + eps_root) + eps)              (optax.adam, b1=.9 b2=.99 eps=eps_root=1e-15)
//   decay   : + weight_decay * p for the MLP weights only, added AFTER the lr scaling with a plus sign,
//             exactly as the reference chains optax.add_decayed_weights behind adam (_utils.py:45-77)
//
// The step counter lives in device memory so the whole training step can sit in one CUDA graph.
#include "adam.cuh"

namespace ngp {
namespace {

// two float4 per thread and iteration: 8 independent 16-byte loads in flight before the first dependent instruction
__global__ void __launch_bounds__(256) adam_kernel(NgpAdamDescriptor d, const uint32_t *__restrict__ step_ptr,
                                                   float *__restrict__ params, const float *__restrict__ grads,
                                                   float *__restrict__ m, float *__restrict__ v) {
    const AdamStepConstants c = adam_step_constants(d, __ldg(step_ptr));
    const size_t n4 = d.n / 4, stride = (size_t)gridDim.x * blockDim.x;
    float4 *p4 = reinterpret_cast<float4 *>(params), *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
    const float4 *g4 = reinterpret_cast<const float4 *>(grads);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
        const size_t j = i + stride;
        const bool two = j < n4;
        // The moments are touched by this kernel only: they stream through L2 evict-first (ld.cs / st.cs), so that the
        // gradients the backward just left in L2 are read from there and the parameters written here stay resident for
        // the next step's gathers instead of being pushed out by 195 MB of moments.
        float4 p0 = p4[i], g0 = __ldg(g4 + i), m0 = __ldcs(m4 + i), v0 = __ldcs(v4 + i);
        float4 p1 = p0, g1 = g0, m1 = m0, v1 = v0;
        if (two) { p1 = p4[j]; g1 = __ldg(g4 + j); m1 = __ldcs(m4 + j); v1 = __ldcs(v4 + j); }
        adam_update4(d, c, (i * 4 >= d.decay_begin) ? d.weight_decay : 0.f, p0, g0, m0, v0);
        p4[i] = p0; __stcs(m4 + i, m0); __stcs(v4 + i, v0);
        if (two) {
            adam_update4(d, c, (j * 4 >= d.decay_begin) ? d.weight_decay : 0.f, p1, g1, m1, v1);
            p4[j] = p1; __stcs(m4 + j, m1); __stcs(v4 + j, v1);
        }
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
