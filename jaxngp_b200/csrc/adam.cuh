// adam.cuh -- the arithmetic of the reference's optimizer (app/nerf/_utils.py:19-77), shared by the local step
// (optim.cu) and the step fused with the gradient exchange (exchange.cu) so that both produce the same bits.
//
//   lr(t)   : optax.exponential_decay(init, transition_steps, decay_rate, transition_begin, staircase, end_value)
//   update  : -lr * m_hat / (sqrt(v_hat + eps_root) + eps)              (optax.adam, b1=.9 b2=.99 eps=eps_root=1e-15)
//   decay   : + weight_decay * p for the MLP weights only, added AFTER the lr scaling with a plus sign,
//             exactly as the reference chains optax.add_decayed_weights behind adam (_utils.py:45-77)
#pragma once
#include "common.cuh"

namespace ngp {

struct AdamStepConstants {
    float lr, inv_bc1, inv_bc2;
};

// `completed` = number of optimizer steps already applied (the device-resident counter)
__device__ __forceinline__ AdamStepConstants adam_step_constants(const NgpAdamDescriptor &d, uint32_t completed) {
    const uint32_t t = completed + 1u;  // optax counts from 1 for the bias correction
    // learning-rate schedule evaluated at count = t - 1 (optax schedules see the pre-increment count)
    float lr = d.lr_init;
    const float count = (float)(t - 1u);
    if (d.transition_steps > 0) {
        float p = fmaxf(count - (float)d.transition_begin, 0.f) / (float)d.transition_steps;
        if (d.staircase) p = floorf(p);
        lr = (count <= (float)d.transition_begin) ? d.lr_init : d.lr_init * powf(d.decay_rate, p);
        lr = d.decay_rate < 1.f ? fmaxf(lr, d.lr_end) : fminf(lr, d.lr_end);
    }
    AdamStepConstants c;
    c.lr = lr;
    // the two bias corrections are per-step constants: their reciprocals are taken once (correctly rounded) and
    // multiplied in, instead of two IEEE divisions per parameter -- the kernel moves 28 B per parameter and was
    // spending as many issue slots on its three divisions and the square root as on the memory traffic
    c.inv_bc1 = __fdiv_rn(1.f, 1.f - powf(d.b1, (float)t));
    c.inv_bc2 = __fdiv_rn(1.f, 1.f - powf(d.b2, (float)t));
    return c;
}

// four consecutive parameters; wd = weight_decay or 0 for this float4
__device__ __forceinline__ void adam_update4(const NgpAdamDescriptor &d, const AdamStepConstants &c, float wd, float4 &p,
                                             float4 g, float4 &mm, float4 &vv) {
    float *pp = &p.x, *gg = &g.x, *pm = &mm.x, *pv = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float gk = gg[k] * d.grad_scale;
        pm[k] = d.b1 * pm[k] + (1.f - d.b1) * gk;
        pv[k] = d.b2 * pv[k] + (1.f - d.b2) * gk * gk;
        const float upd = __fdividef(-c.lr * (pm[k] * c.inv_bc1), __fsqrt_rn(pv[k] * c.inv_bc2 + d.eps_root) + d.eps);
        pp[k] = pp[k] + (upd + wd * pp[k]);
    }
}

}  // namespace ngp
