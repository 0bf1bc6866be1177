/*
 * ngp_b200.h -- C ABI of libngp_b200.so: hand-written sm_100a kernels for jaxngp's NeRF hot path.
 *
 * Every op is an XLA *legacy* GPU custom call, the exact ABI the reference registers
 * (deps/volume-rendering-jax/lib/ffi.cc:17-51, deps/jax-tcnn/lib/ffi.cc:25-30):
 *
 *     void op(cudaStream_t stream, void **buffers, const char *opaque, size_t opaque_len);
 *
 * `buffers` = operands in lowering order followed by results, all device pointers owned by the
 * caller; results arrive uninitialised and every promised byte is defined by the op; `opaque` is
 * the raw little-endian bytes of the descriptor struct (deps/serde-helper/serde.h:30-40).  Ops only
 * enqueue work on `stream`: no host synchronisation, no allocation on the data path (one cached
 * scratch block per stream is created on first use), re-entrant across host threads.
 *
 * Errors never cross the boundary as C++ exceptions (the reference throws through XLA's C
 * callback, volrend.h:9-16): a failed call records a thread-local status, readable with
 * ngp_b200_last_status()/ngp_b200_last_error(), and enqueues nothing.
 *
 * Section A are drop-ins for the reference's registered targets (same names in the registry, same
 * descriptors byte for byte, same buffer order).  Section B are additions of this library (the
 * pure-JAX HashGridEncoder of models/encoders.py as one kernel pair, the density-grid update,
 * fused training helpers); they use the same calling convention so that one binding serves all.
 *
 * All `file:line` citations are relative to the reference checkout (blurgyy/jaxngp).
 */
#ifndef NGP_B200_H_
#define NGP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __DRIVER_TYPES_H__
typedef struct CUstream_st *cudaStream_t;
#endif

#define NGP_B200_ABI_VERSION 1

/* ------------------------------------------------------------------ descriptors (section A)
 * Byte-identical to deps/volume-rendering-jax/lib/impl/volrend.h:23-118 and
 * deps/jax-tcnn/lib/impl/tcnnutils.h:11-29. */
typedef struct { uint32_t n_bytes; } NgpPackbitsDescriptor;                       /* volrend.h:116-118 */
typedef struct { uint32_t length; } NgpMorton3DDescriptor;                        /* volrend.h:110-113 */
typedef struct {                                                                   /* volrend.h:57-85  */
    uint32_t n_rays, total_samples, diagonal_n_steps, K, G;
    float bound, stepsize_portion;
} NgpMarchingDescriptor;
typedef struct {                                                                   /* volrend.h:88-108 */
    uint32_t n_total_rays, n_rays, diagonal_n_steps, K, G, march_steps_cap;
    float bound, stepsize_portion;
} NgpMarchingInferenceDescriptor;
typedef struct { uint32_t n_rays, total_samples; } NgpIntegratingDescriptor;       /* volrend.h:23-29  */
typedef struct {                                                                   /* volrend.h:32-43  */
    uint32_t n_rays, total_samples;
    float near_distance;
} NgpIntegratingBackwardDescriptor;
typedef struct { uint32_t n_total_rays, n_rays, march_steps_cap; } NgpIntegratingInferenceDescriptor; /* volrend.h:46-55 */
typedef struct {                                                                   /* tcnnutils.h:11-29 */
    uint32_t n_coords, L, F, N_min;
    float per_level_scale;
} NgpHashGridDescriptor;

/* ------------------------------------------------------------------ section A: drop-in targets */

/* replaces volrendjax::pack_density_into_bits (volrend.h:121-126, packbits.cu:36-72)
 * in : density_threshold f32[N], density_grid f32[N]      out: occupied_mask bool[N], bitfield u8[N/8] */
void ngp_pack_density_into_bits(cudaStream_t, void **, const char *, size_t);

/* replaces volrendjax::march_rays (volrend.h:128-133, marching.cu:435-522)
 * in : rays_o f32[n,3], rays_d f32[n,3], t_starts f32[n], t_ends f32[n], noises f32[n], bitfield u8[K*G^3/8]
 * out: next_sample_write_location u32[1], number_of_exceeded_samples u32[1], ray_is_valid bool[n],
 *      rays_n_samples u32[n], rays_sample_startidx u32[n], idcs u32[S], xyzs f32[S,3], dirs f32[S,3],
 *      dss f32[S], z_vals f32[S]
 * Sample ranges are handed out in RAY ORDER (deterministic); the reference hands them out in
 * atomic arrival order (marching.cu:205).  See DESIGN.md "march_rays compaction". */
void ngp_march_rays(cudaStream_t, void **, const char *, size_t);
/* Launch hint for the calling host thread: persistent CTAs per SM of the following ngp_march_rays launches
 * (0 = default, 4).  A caller that overlaps the march with other kernels uses 1 so that it runs underneath them. */
void ngp_b200_set_march_ctas_per_sm(int ctas_per_sm);

/* replaces volrendjax::march_rays_inference (volrend.h:135-140, marching.cu:524-604)
 * in : rays_o f32[N,3], rays_d f32[N,3], t_starts f32[N], t_ends f32[N], bitfield u8[..],
 *      next_ray_index_in u32[1], terminated bool[n], indices_in u32[n]
 * out: next_ray_index u32[1], indices_out u32[n], n_samples u32[n], t_starts_out f32[n],
 *      xyzs f32[n,cap,3], dss f32[n,cap], z_vals f32[n,cap]
 * No stream synchronisation (the reference has one, marching.cu:562); fresh rays are handed to
 * terminated slots in slot order. */
void ngp_march_rays_inference(cudaStream_t, void **, const char *, size_t);

/* Renderer fast path (no reference counterpart): advances every ray of a frame through the
 * empty space in front of its first occupied point, following the reference loop's visit rule exactly
 * (marching.cu:323-365), so that march_rays_inference started from t_out emits bit-identical samples.
 * descriptor: NgpMarchingInferenceDescriptor (n_rays and march_steps_cap unused)
 * in : rays_o f32[N,3], rays_d f32[N,3], t_starts f32[N], t_ends f32[N], bitfield u8[..]
 * out: t_out f32[N] (may alias t_starts) */
void ngp_march_rays_skip_empty(cudaStream_t, void **, const char *, size_t);

/* Renderer fast paths: the two inference ops with the scatters the reference's Python wrappers do afterwards
 * (marching/__init__.py:156, integrating/__init__.py:108-109) folded into the kernels; same descriptors.
 * march_rays_inference_inplace
 *   in/out: t_starts f32[N] (advanced in place), next_ray_index u32[1], indices u32[n]
 *   in    : rays_o, rays_d, t_ends, bitfield, terminated bool[n]
 *   out   : n_samples u32[n], xyzs f32[n,cap,3], dss f32[n,cap], z_vals f32[n,cap], ray_dirs f32[n,3]
 *           (only the first n_samples[i] rows of slot i are defined: unlike march_rays_inference, which zero-fills
 *            the tail as the reference does, this variant feeds consumers that never read past n_samples)
 *   buffers: rays_o, rays_d, t_starts, t_ends, bitfield, next_ray_index, terminated, indices,
 *            n_samples, xyzs, dss, z_vals, ray_dirs
 * integrate_rays_inference_inplace
 *   in/out: rays_rgbd f32[N,4], rays_T f32[N], counters u64[2] (+= terminated rays, += samples marched)
 *   in    : rays_bg, n_samples, indices, dss, z_vals, drgbs        out: terminated bool[n]
 *   buffers: rays_bg, rays_rgbd, rays_T, n_samples, indices, dss, z_vals, drgbs, terminated, counters */
void ngp_march_rays_inference_inplace(cudaStream_t, void **, const char *, size_t);
void ngp_integrate_rays_inference_inplace(cudaStream_t, void **, const char *, size_t);

/* replace volrendjax::morton3d / morton3d_invert (volrend.h:143-154, marching.cu:606-665)
 * morton3d: in xyzs u32[len,3], out idcs u32[len];  invert: in idcs u32[len], out xyzs u32[len,3] */
void ngp_morton3d(cudaStream_t, void **, const char *, size_t);
void ngp_morton3d_invert(cudaStream_t, void **, const char *, size_t);

/* replaces volrendjax::integrate_rays (volrend.h:156-161, integrating.cu:325-377)
 * in : rays_sample_startidx u32[n], rays_n_samples u32[n], bgs f32[n,3], dss f32[S], z_vals f32[S], drgbs f32[S,4]
 * out: measured_batch_size u32[1], final_rgbds f32[n,4], final_opacities f32[n] */
void ngp_integrate_rays(cudaStream_t, void **, const char *, size_t);

/* replaces volrendjax::integrate_rays_backward (volrend.h:163-168, integrating.cu:379-446)
 * in : startidx, n_samples, bgs, dss, z_vals, drgbs, final_rgbds f32[n,4], final_opacities f32[n], dL_dfinal_rgbds f32[n,4]
 * out: dL_dbgs f32[n,3], dL_dz_vals f32[S], dL_ddrgbs f32[S,4] */
void ngp_integrate_rays_backward(cudaStream_t, void **, const char *, size_t);

/* replaces volrendjax::integrate_rays_inference (volrend.h:170-175, integrating.cu:448-507)
 * in : rays_bg f32[N,3], rays_rgbd f32[N,4], rays_T f32[N], n_samples u32[n], indices u32[n],
 *      dss f32[n,cap], z_vals f32[n,cap], drgbs f32[n,cap,4]
 * out: terminate_cnt u32[1], terminated bool[n], rays_rgbd_out f32[n,4], rays_T_out f32[n] */
void ngp_integrate_rays_inference(cudaStream_t, void **, const char *, size_t);

/* replaces jaxtcnn::hashgrid_encode (tcnnutils.h:31-36, hashgrid.cu:20-88; tiny-cuda-nn v1.6
 * kernel_grid<float,3,F,CoherentPrime> semantics: f32-on-device level scale, index % level size)
 * in : offset_table u32[L+1], coords_rm f32[3,n], params f32[rows,F]
 * out: encoded_rm f32[L*F,n], dy_dcoords_rm f32[3*L*F,n] (stored as float3 per (feature,point)) */
void ngp_hashgrid_encode(cudaStream_t, void **, const char *, size_t);

/* replaces jaxtcnn::hashgrid_encode_backward (tcnnutils.h:38-43, hashgrid.cu:90-174)
 * in : offset_table u32[L+1], coords_rm f32[3,n], dL_dy_rm f32[L*F,n], dy_dcoords_rm f32[3*L*F,n]
 * out: dL_dparams f32[rows,F], dL_dcoords_rm f32[3,n] */
void ngp_hashgrid_encode_backward(cudaStream_t, void **, const char *, size_t);

/* ------------------------------------------------------------------ section B: additions */

#define NGP_HG_MAX_LEVELS 32

/* One kernel pair for the pure-JAX HashGridEncoder.__call__ (models/encoders.py:82-256): the level
 * table is computed by the host exactly as encoders.py:89-103 does and travels in the descriptor.
 * wrap_T != 0 reproduces `indices mod T` on every level (encoders.py:187); 0 = modulo level size. */
typedef struct {
    uint32_t n_points, dim, L, F;
    uint32_t wrap_T;
    uint32_t table_dtype; /* 0 = f32 rows (API dtype), 1 = f16 rows (storage variant) */
    float bound;
    uint32_t hashed_mask; /* bit l set = level l uses the spatial hash */
    float scales[NGP_HG_MAX_LEVELS];
    uint32_t res[NGP_HG_MAX_LEVELS];
    uint32_t offsets[NGP_HG_MAX_LEVELS + 1];
    /* forward only: if non-zero, points come in groups of `rows_per_group` (the [n_rays, march_steps_cap]
     * layout of march_rays_inference) and a 4th input buffer group_counts u32[n_points / rows_per_group]
     * says how many leading rows of each group are real samples; the other rows are neither read nor written. */
    uint32_t rows_per_group;
} NgpHashGridA1Descriptor;

/* in : pos f32[n,dim], table (f32|f16)[rows,F] [, group_counts]   out: enc f32[n,L*F] */
void ngp_hashgrid_a1_forward(cudaStream_t, void **, const char *, size_t);
/* in : pos f32[n,dim], d_enc f32[n,L*F]                 out: d_table f32[rows,F] (zero-filled, then scatter-added) */
void ngp_hashgrid_a1_backward(cudaStream_t, void **, const char *, size_t);
/* same, without the zero-fill: scatter-ADDS into d_table (a batch processed in chunks) */
void ngp_hashgrid_a1_backward_acc(cudaStream_t, void **, const char *, size_t);

/* packbits with a scalar threshold read from device memory (drops the broadcast array the
 * reference materialises, packbits/__init__.py:25-28).
 * in : threshold f32[1], density f32[N]                  out: occupied_mask bool[N], bitfield u8[N/8] */
void ngp_packbits_scalar(cudaStream_t, void **, const char *, size_t);

/* Density-grid update, utils/types.py:1149-1239 (NeRFState.update_ogrid_density / threshold_ogrid).
 * sample_positions: in idx u32[M] (Morton cell indices inside the cascade), uniforms f32[M,3] in [0,1)
 *                   out coords f32[M,3]                                              (:1193-1206)
 * decay_max       : in density f32[N], idx u32[M], new_density f32[M]; out density f32[N] (may alias the
 *                   input): alive cells * decay, then max with the new values        (:1162-1164,1219-1221)
 * threshold       : in density f32[n_cells] (alive cells of the cascades the mean runs over)
 *                   out thr f32[1] = min(thr_max, mean)                              (:1229-1230,143-144) */
typedef struct { uint32_t n_points, G; float mip_bound; } NgpOgridSampleDescriptor;
typedef struct { uint32_t n_cells, n_updates; float decay; } NgpOgridUpdateDescriptor;
typedef struct { uint32_t n_cells; float thr_max; } NgpOgridThresholdDescriptor;
void ngp_ogrid_sample_positions(cudaStream_t, void **, const char *, size_t);
void ngp_ogrid_decay_max(cudaStream_t, void **, const char *, size_t);
void ngp_ogrid_threshold(cudaStream_t, void **, const char *, size_t);

/* Random inputs of the path.  The reference draws them with jax.random outside its ops (march perturbations
 * models/renderers/cuda.py:118-122, random backgrounds app/nerf/_utils.py:134-136, the cell draws and the jitter of
 * utils/types.py:1170-1206); here they are Philox4x32-10 blocks, a pure function of (seed, stream_id, call counter,
 * element index): element i of call c is philox(counter = {i, c, stream_id, 0}, key = {seed_lo, seed_hi}), each word
 * mapped to [0, 1) by jax.random.uniform's construction (23 mantissa bits under exponent 0, minus 1).
 * Ops that consume a call counter take `rng_state u32[2]` = {counter, ticket} in device memory: every block reads the
 * counter, the last block of the launch increments it -- a replayed CUDA graph draws fresh numbers on every replay. */
typedef struct { uint32_t seed_lo, seed_hi, stream_id; } NgpRngDescriptor;
/* philox_uniform: out f32[n,4] = the four uniforms of every element of call `counter` (tests, eager replays) */
typedef struct { uint32_t n, counter; NgpRngDescriptor rng; } NgpPhiloxDescriptor;
void ngp_philox_uniform(cudaStream_t, void **, const char *, size_t);

/* Cell draws + sample positions of update_ogrid_density (utils/types.py:1166-1206) for one cascade in one op.
 * mode 1 (update_all): every trainable cell once, in order.  mode 0: n_first cells uniform among the trainable cells,
 * then n_second cells uniform among the currently OCCUPIED ones (jran.choice with p = occ_mask: inverse CDF over a
 * prefix count of the cascade's bitfield), with replacement.  Each cell gets a jittered point inside it (:1193-1206).
 * in : bitfield u32[n_cells/32] (the cascade's slice of the occupancy bitfield, 16-byte aligned),
 *      alive u32[n_alive] (Morton indices of the trainable cells inside the cascade) or NULL if has_alive == 0,
 *      rng_state u32[2] (in/out)
 * out: idx u32[M], coords f32[M,3]        M = mode ? n_alive : n_first + n_second */
typedef struct {
    uint32_t n_cells, G, n_alive, has_alive, mode, n_first, n_second;
    float mip_bound;
    NgpRngDescriptor rng;
} NgpOgridDrawDescriptor;
void ngp_ogrid_draw_cells(cudaStream_t, void **, const char *, size_t);

/* Fully fused NeRF MLP of make_nerf_ngp (models/nerfs.py:27-128,216-238,422-454) on the tensor cores.
 * weights = flat f32[9408] = [density W0 32x64 | density W1 64x16 | rgb W0 32x64 | rgb W1 64x64 | rgb W2 64x3],
 * each row-major [in][out] like the flax Dense kernels.
 * forward : in enc f32[n,32], dirs f32[n,3] (unit), weights [, group_counts];  out drgbs f32[n,4]  (density_only: out f32[n], dirs unused)
 * backward: in enc, dirs, weights, d_drgbs f32[n,4];          out d_enc f32[n,32], d_weights f32[9408] */
typedef struct {
    uint32_t n_samples, density_only;
    /* forward only: grouped layout of march_rays_inference.  If rows_per_group != 0: dirs is f32[n_groups,3]
     * (one direction per ray) and a 5th input buffer group_counts u32[n_groups] masks the rows as above. */
    uint32_t rows_per_group;
} NgpNerfMlpDescriptor;
void ngp_nerf_mlp_forward(cudaStream_t, void **, const char *, size_t);
void ngp_nerf_mlp_backward(cudaStream_t, void **, const char *, size_t);
/* MLP backward with the hash-table scatter of ngp_hashgrid_a1_backward fused behind it: the input gradient d_enc never
 * leaves the SM (each thread scatters the fragment it holds).  descriptor: NgpHashGridA1Descriptor (n_points = samples;
 * dim 3, L 16, F 2, power-of-two wrap_T).
 * in : enc f32[n,32], dirs f32[n,3], weights f32[9408], d_drgbs f32[n,4], pos f32[n,3]
 * out: d_weights f32[9408], d_table f32[rows,2] (zero-filled, then scatter-added) */
void ngp_nerf_mlp_backward_scatter(cudaStream_t, void **, const char *, size_t);
/* Same contract, but d_weights is ADDED to instead of defined (a batch processed in chunks: first chunk = the op above). */
void ngp_nerf_mlp_backward_acc(cudaStream_t, void **, const char *, size_t);
/* Same contract, weight gradients on mma.sync instead of tcgen05/TMEM (the cross-check arm of the tests). */
void ngp_nerf_mlp_backward_mma(cudaStream_t, void **, const char *, size_t);
/* Same contract, EVERY matrix product on tcgen05 (csrc/mlp_bwd_tc.cu): forward recompute and delta chain with the A
 * operand in tensor memory (each accumulator rewritten in place by its epilogue), weight gradients from shared-memory
 * panels.  Same tf32 operand rounding; results agree with the two kernels above to f32 accumulation-order error. */
void ngp_nerf_mlp_backward_tc(cudaStream_t, void **, const char *, size_t);

/* HashGridEncoder fused in front of the MLP forward: the [n,32] encoding feeds the first layer's
 * tensor-core fragments directly.  Bit-identical to hashgrid_a1_forward followed by nerf_mlp_forward.
 * Needs dim=3, L=16, F=2, power-of-two wrap_T, table aligned to two rows; otherwise NGP_ERR_ARGUMENT (status -2).
 * in : pos f32[n,3], table (f32|f16)[rows,2], dirs f32[n,3] (grouped: f32[n_groups,3]; density_only: unused),
 *      weights f32[9408] [, group_counts u32[n_groups] if grid.rows_per_group != 0]
 * out: drgbs f32[n,4] (density_only: f32[n]) [, enc f32[n,32] if write_enc] */
typedef struct { NgpHashGridA1Descriptor grid; uint32_t density_only, write_enc; } NgpNerfFusedDescriptor;
void ngp_nerf_fused_forward(cudaStream_t, void **, const char *, size_t);
/* Same contract (density_only = 0 only) with the five dense layers on tcgen05.mma kind::tf32 and TMEM accumulators:
 * 128-sample tiles, one thread per sample from the gather to the output.  Same tf32 operand rounding as the mma.sync
 * kernel; results agree to f32 accumulation-order error (rel 1e-5), not bit for bit. */
void ngp_nerf_fused_forward_umma(cudaStream_t, void **, const char *, size_t);

/* Whole-frame inference kernel: render_image_inference's slot-refill loop (models/renderers/cuda.py:244-373:
 * march_rays_inference -> NeRF -> integrate_rays_inference -> scatter, until every ray has terminated) as ONE persistent
 * launch -- ray slots, samples and compositing state stay on the SM; a free slot takes the next ray with one atomicAdd.
 * Same march code, same encoder + tcgen05 MLP code and the same compositing expressions as the ops it replaces.
 * `march.march_steps_cap` must be 16 and `march.n_rays` is unused; rays arrive pre-advanced by ngp_march_rays_skip_empty.
 * in : rays_o f32[N,3], rays_d f32[N,3], t_starts f32[N], t_ends f32[N], bitfield u8[K*G^3/8], rays_bg f32[N,3],
 *      table (f32|f16)[rows,2], weights f32[9408]
 * scratch (zeroed by the op): next_ray u32[1], counters u64[2] (rays terminated, samples marched)
 * out: rays_rgbd f32[N,4] (colour composited onto the background, depth) */
typedef struct { NgpHashGridA1Descriptor grid; NgpMarchingInferenceDescriptor march; } NgpRenderFrameDescriptor;
void ngp_render_frame(cudaStream_t, void **, const char *, size_t);

/* Diagnostic: known-answer test of the tcgen05 (kind::tf32) operand formats of csrc/umma.cuh -- forward, dgrad and
 * wgrad GEMM shapes on swizzled shared-memory panels with TMEM accumulators.  No descriptor (opaque_len = 0).
 * in : A f32[128,32], W f32[32,64], G f32[128,64]     out: A.W f32[128,64], G.W^T f32[128,32], G^T.A f32[64,32] */
void ngp_umma_selftest(cudaStream_t, void **, const char *, size_t);

/* Training glue around the four ops (XLA fuses these elementwise chains for the reference):
 * make_training_rays: app/nerf/_utils.py:93-115 + utils/types.py:398-439 (undistorted PERSPECTIVE camera)
 *                     + models/renderers/cuda.py:57-97
 *   in : perm i32[n] (index into views*H*W), transforms f32[V,12] (R row-major, t)
 *   out: rays_o f32[n,3], rays_d f32[n,3], t_starts f32[n], t_ends f32[n]
 * huber_loss_grad   : app/nerf/_utils.py:151-165 + utils/data.py:443-464
 *   in : final_rgbds f32[n,4], ray_is_valid bool[n], perm i32[n], rgbas u8[P,4], bgs f32[n,3]
 *   out: dL_dfinal_rgbds f32[n,4], loss f32[1], n_valid_rays u32[1] */
typedef struct { uint32_t n_rays, width, height, n_views; float fx, fy, cx, cy, bound; } NgpTrainingRaysDescriptor;
typedef struct { uint32_t n_rays; float delta; } NgpHuberLossDescriptor;
void ngp_make_training_rays(cudaStream_t, void **, const char *, size_t);
void ngp_huber_loss_grad(cudaStream_t, void **, const char *, size_t);
/* make_training_rays plus the step's random inputs from the same kernel (see NgpRngDescriptor above):
 *   in : perm i32[n], transforms f32[V,12], rng_state u32[2] (in/out)
 *   out: rays_o, rays_d, t_starts, t_ends as above, noises f32[n] (uniform x), bgs f32[n,3] (uniforms y, z, w) */
typedef struct { NgpTrainingRaysDescriptor rays; NgpRngDescriptor rng; } NgpTrainingRaysRngDescriptor;
void ngp_make_training_rays_rng(cudaStream_t, void **, const char *, size_t);

/* integrate_rays (integrating.cu:24-99) + the Huber loss of train_step (app/nerf/_utils.py:151-165, target pixel
 * composited onto the random background, utils/data.py:459-463) + integrate_rays_backward (integrating.cu:101-240,
 * colour / density gradient only) in one pass per ray: bit-identical final_rgbds, opacities and dL_ddrgbs to the
 * three ops in sequence.
 *   in : rays_sample_startidx u32[n], rays_n_samples u32[n], bgs f32[n,3], dss f32[S], z_vals f32[S], drgbs f32[S,4],
 *        ray_is_valid bool[n], perm i32[n], rgbas u8[P,4]
 *   out: measured_batch_size u32[1], final_rgbds f32[n,4], final_opacities f32[n], dL_ddrgbs f32[S,4], loss f32[1],
 *        n_valid_rays u32[1] */
typedef struct { uint32_t n_rays, total_samples; float near_distance, delta; } NgpIntegrateLossDescriptor;
void ngp_integrate_loss_fused(cudaStream_t, void **, const char *, size_t);

/* Adam step of app/nerf/_utils.py:19-77 over a flat f32 buffer [hash table | MLP weights]; elements
 * at index >= decay_begin also receive the reference's (additive) decayed-weights term.
 * in : step u32[1] (device-resident count of completed steps), params f32[n] (updated in place),
 *      grads f32[n], m f32[n], v f32[n] (updated in place) */
typedef struct {
    uint64_t n, decay_begin;
    float lr_init, lr_end, decay_rate;
    uint32_t transition_steps, transition_begin, staircase;
    float b1, b2, eps, eps_root, weight_decay, grad_scale;
} NgpAdamDescriptor;
void ngp_adam_step(cudaStream_t, void **, const char *, size_t);

/* out[0] = sa * a[0] + sb * b[0] + c on device-resident 32-bit scalars (a, b, out may alias): the step counter after an
 * optimizer step, and train_step's `next_sample_write_location - number_of_exceeded_samples` (marching/__init__.py:91).
 * in : a u32[1], b u32[1]        out: out u32[1] */
typedef struct { int32_t sa, sb, c; } NgpU32AxpyDescriptor;
void ngp_u32_axpy(cudaStream_t, void **, const char *, size_t);

/* Gradient exchange fused with the optimizer, for ray-sharded data parallelism (the reference trains on one GPU;
 * this is app/nerf/_utils.py:19-77 applied to the SUM of every rank's gradient).  ONE kernel per rank does
 * what reduce-scatter -> ngp_adam_step -> all-gather do: it waits until every peer's gradient buffer is complete,
 * reads this rank's shard of the summed gradient straight from the peers over NVLink (multimem.ld_reduce through the
 * NVSwitch when multicast pointers are given, otherwise one load per peer in rank order), applies Adam to the shard
 * (m, v are shard-local), and stores the updated parameters into EVERY rank's parameter buffer (multimem.st, or one
 * store per peer).  Each element is reduced exactly once, by its owner, so replicas stay bit-identical.  All ranks
 * must launch it with the same descriptor apart from rank / shard_begin.
 * in : step u32[1], m f32[n], v f32[n] (n = adam.n, this rank's shard; updated in place),
 *      grads_ptrs u64[world] (device array: base of rank r's flat gradient buffer, mapped in this process),
 *      params_ptrs u64[world] (same for the flat parameter buffers; element shard_begin + i of every one is written),
 *      signal_ptrs u64[world] (rank r's signal pad: zero-initialised u32 words, >= signal_base + n_blocks*world),
 *      grads_mc, params_mc (multicast bases of the two buffers, read only if use_multimem) */
typedef struct {
    NgpAdamDescriptor adam;            /* n = shard length, decay_begin relative to the shard */
    uint64_t shard_begin;              /* first element of this rank's shard in the flat buffers (multiple of 4) */
    uint32_t rank, world;              /* world <= 8 */
    uint32_t use_multimem, n_blocks;   /* n_blocks: same on every rank, <= resident CTAs of the device */
    uint32_t signal_base, timeout_ms;  /* first u32 word of the pads this op may use; how long a block waits for a
                                          peer before it traps (0 = 20 s): a missing rank fails the launch, not the GPU */
} NgpAdamExchangeDescriptor;
void ngp_adam_step_exchange(cudaStream_t, void **, const char *, size_t);

/* ------------------------------------------------------------------ section A, status-returning form
 * XLA's API_VERSION_STATUS_RETURNING custom-call signature (the api_version jax's `custom_call` lowering emits by
 * default): the four arguments above plus an XlaCustomCallStatus*.  Each `_status` entry point runs the op of the same
 * name and, if it failed and `status` is non-null, reports the failure through XLA's XlaCustomCallStatusSetFailure
 * (resolved in the process at first use, or supplied with ngp_b200_set_status_failure_fn), so the error surfaces as an
 * XlaRuntimeError instead of the C++ exception the reference throws through XLA's C frames (volrend.h:9-16).
 * Register these under the reference's target names (jaxngp_b200/jax_ffi.py); the four-argument symbols stay for
 * hosts that call them directly. */
typedef struct XlaCustomCallStatus_ XlaCustomCallStatus;
typedef void (*ngp_set_failure_fn)(XlaCustomCallStatus *, const char *message, size_t message_len);
void ngp_b200_set_status_failure_fn(ngp_set_failure_fn fn);
void ngp_pack_density_into_bits_status(cudaStream_t, void **, const char *, size_t, XlaCustomCallStatus *);
void ngp_march_rays_status(cudaStream_t, void **, const char *, size_t, XlaCustomCallStatus *);
void ngp_march_rays_inference_status(cudaStream_t, void **, const char *, size_t, XlaCustomCallStatus *);
void ngp_morton3d_status(cudaStream_t, void **, const char *, size_t, XlaCustomCallStatus *);
void ngp_morton3d_invert_status(cudaStream_t, void **, const char *, size_t, XlaCustomCallStatus *);
void ngp_integrate_rays_status(cudaStream_t, void **, const char *, size_t, XlaCustomCallStatus *);
void ngp_integrate_rays_backward_status(cudaStream_t, void **, const char *, size_t, XlaCustomCallStatus *);
void ngp_integrate_rays_inference_status(cudaStream_t, void **, const char *, size_t, XlaCustomCallStatus *);
void ngp_hashgrid_encode_status(cudaStream_t, void **, const char *, size_t, XlaCustomCallStatus *);
void ngp_hashgrid_encode_backward_status(cudaStream_t, void **, const char *, size_t, XlaCustomCallStatus *);

/* ------------------------------------------------------------------ status */
int ngp_b200_abi_version(void);
/* 0 = last call on this host thread succeeded; otherwise a cudaError_t or a negative ngp code */
int ngp_b200_last_status(void);
const char *ngp_b200_last_error(void);
void ngp_b200_clear_error(void);

#ifdef __cplusplus
}
#endif
#endif /* NGP_B200_H_ */
