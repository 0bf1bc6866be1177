"""Opaque descriptor factories: the wire format is the raw little-endian bytes of the POD structs in
include/ngp_b200.h, byte-identical to the reference's (deps/volume-rendering-jax/lib/ffi.cc:55-207,
deps/jax-tcnn/lib/ffi.cc:31-56 via serde-helper/serde.h:30-33).  Same names, arguments, validation.
"""
import struct

HG_MAX_LEVELS = 32


def _u32(x, name):
    x = int(x)
    if not 0 <= x < 2 ** 32:
        raise ValueError(f"{name} must fit in uint32, got {x}")
    return x


def make_packbits_descriptor(n_bytes):
    if n_bytes == 0:  # ffi.cc:57-59
        raise RuntimeError("expected n_bytes to be a positive integer, got 0")
    return struct.pack("<I", _u32(n_bytes, "n_bytes"))


def make_morton3d_descriptor(length):
    return struct.pack("<I", _u32(length, "length"))


def make_marching_descriptor(n_rays, total_samples, diagonal_n_steps, K, G, bound, stepsize_portion):
    if K == 0:  # ffi.cc:79-81
        raise RuntimeError("expected K to be a positive integer, got 0")
    return struct.pack("<5I2f", _u32(n_rays, "n_rays"), _u32(total_samples, "total_samples"),
                       _u32(diagonal_n_steps, "diagonal_n_steps"), _u32(K, "K"), _u32(G, "G"),
                       float(bound), float(stepsize_portion))


def make_marching_inference_descriptor(n_total_rays, n_rays, diagonal_n_steps, K, G, march_steps_cap,
                                       bound, stepsize_portion):
    if K == 0:  # ffi.cc:114-116
        raise RuntimeError("expected K to be a positive integer, got 0")
    return struct.pack("<6I2f", _u32(n_total_rays, "n_total_rays"), _u32(n_rays, "n_rays"),
                       _u32(diagonal_n_steps, "diagonal_n_steps"), _u32(K, "K"), _u32(G, "G"),
                       _u32(march_steps_cap, "march_steps_cap"), float(bound), float(stepsize_portion))


def make_integrating_descriptor(n_rays, total_samples):
    return struct.pack("<2I", _u32(n_rays, "n_rays"), _u32(total_samples, "total_samples"))


def make_integrating_backward_descriptor(n_rays, total_samples, near_distance):
    return struct.pack("<2If", _u32(n_rays, "n_rays"), _u32(total_samples, "total_samples"), float(near_distance))


def make_integrating_inference_descriptor(n_total_rays, n_rays, march_steps_cap):
    return struct.pack("<3I", _u32(n_total_rays, "n_total_rays"), _u32(n_rays, "n_rays"),
                       _u32(march_steps_cap, "march_steps_cap"))


def make_hashgrid_descriptor(n_coords, L, F, N_min, per_level_scale):
    return struct.pack("<4If", _u32(n_coords, "n_coords"), _u32(L, "L"), _u32(F, "F"), _u32(N_min, "N_min"),
                       float(per_level_scale))


def make_hashgrid_a1_descriptor(n_points, dim, L, F, wrap_T, table_dtype, bound, hashed, scales, res, offsets,
                                rows_per_group=0):
    """NgpHashGridA1Descriptor (include/ngp_b200.h): level table of models/encoders.py:89-103."""
    if not 0 < L <= HG_MAX_LEVELS:
        raise ValueError(f"L must be in (0, {HG_MAX_LEVELS}], got {L}")
    mask = 0
    for l, h in enumerate(hashed):
        mask |= (1 << l) if h else 0
    pad = HG_MAX_LEVELS - L
    return struct.pack(
        f"<6IfI{HG_MAX_LEVELS}f{HG_MAX_LEVELS}I{HG_MAX_LEVELS + 1}II",
        _u32(n_points, "n_points"), _u32(dim, "dim"), _u32(L, "L"), _u32(F, "F"), _u32(wrap_T, "wrap_T"),
        _u32(table_dtype, "table_dtype"), float(bound), mask,
        *([float(s) for s in scales] + [0.0] * pad),
        *([int(r) for r in res] + [0] * pad),
        *([int(o) for o in offsets] + [0] * pad), _u32(rows_per_group, "rows_per_group"))


def make_adam_descriptor(n, decay_begin, lr_init, lr_end, decay_rate, transition_steps, transition_begin, staircase,
                         b1, b2, eps, eps_root, weight_decay, grad_scale=1.0):
    """NgpAdamDescriptor (include/ngp_b200.h): optimizer constants of app/nerf/_utils.py:19-77."""
    return struct.pack("<2Q3f3I6f", int(n), int(decay_begin), float(lr_init), float(lr_end), float(decay_rate),
                       int(transition_steps), int(transition_begin), int(bool(staircase)), float(b1), float(b2),
                       float(eps), float(eps_root), float(weight_decay), float(grad_scale))


def make_adam_exchange_descriptor(adam_descriptor, shard_begin, rank, world, use_multimem, n_blocks, signal_base,
                                  timeout_ms=0):
    """NgpAdamExchangeDescriptor (include/ngp_b200.h): the Adam constants of this rank's shard followed by where the
    shard sits in the flat buffers and who takes part in the exchange."""
    assert len(adam_descriptor) == 64
    return adam_descriptor + struct.pack("<Q6I", int(shard_begin), _u32(rank, "rank"), _u32(world, "world"),
                                         int(bool(use_multimem)), _u32(n_blocks, "n_blocks"),
                                         _u32(signal_base, "signal_base"), _u32(timeout_ms, "timeout_ms"))


def make_ogrid_sample_descriptor(n_points, G, mip_bound):
    return struct.pack("<2If", _u32(n_points, "n_points"), _u32(G, "G"), float(mip_bound))


def make_ogrid_update_descriptor(n_cells, n_updates, decay):
    return struct.pack("<2If", _u32(n_cells, "n_cells"), _u32(n_updates, "n_updates"), float(decay))


def make_ogrid_threshold_descriptor(n_cells, thr_max):
    return struct.pack("<If", _u32(n_cells, "n_cells"), float(thr_max))


def make_nerf_mlp_descriptor(n_samples, density_only=False, rows_per_group=0):
    return struct.pack("<3I", _u32(n_samples, "n_samples"), int(bool(density_only)), _u32(rows_per_group, "rows_per_group"))


def make_training_rays_descriptor(n_rays, width, height, n_views, fx, fy, cx, cy, bound):
    return struct.pack("<4I5f", _u32(n_rays, "n_rays"), _u32(width, "width"), _u32(height, "height"),
                       _u32(n_views, "n_views"), float(fx), float(fy), float(cx), float(cy), float(bound))


def make_integrate_loss_descriptor(n_rays, total_samples, near_distance, delta):
    return struct.pack("<2I2f", _u32(n_rays, "n_rays"), _u32(total_samples, "total_samples"), float(near_distance), float(delta))


def make_huber_loss_descriptor(n_rays, delta):
    return struct.pack("<If", _u32(n_rays, "n_rays"), float(delta))


def make_rng_descriptor(seed, stream_id):
    """NgpRngDescriptor: 64-bit seed as the Philox key, `stream_id` separates independent consumers."""
    seed = int(seed) & (2 ** 64 - 1)
    return struct.pack("<3I", seed & 0xFFFFFFFF, seed >> 32, _u32(stream_id, "stream_id"))


def make_philox_descriptor(n, counter, seed, stream_id):
    return struct.pack("<2I", _u32(n, "n"), _u32(counter, "counter")) + make_rng_descriptor(seed, stream_id)


def make_training_rays_rng_descriptor(n_rays, width, height, n_views, fx, fy, cx, cy, bound, seed, stream_id):
    return make_training_rays_descriptor(n_rays, width, height, n_views, fx, fy, cx, cy, bound) + make_rng_descriptor(seed, stream_id)


def make_ogrid_draw_descriptor(n_cells, G, n_alive, has_alive, update_all, n_first, n_second, mip_bound, seed, stream_id):
    return struct.pack("<7If", _u32(n_cells, "n_cells"), _u32(G, "G"), _u32(n_alive, "n_alive"), int(bool(has_alive)),
                       int(bool(update_all)), _u32(n_first, "n_first"), _u32(n_second, "n_second"),
                       float(mip_bound)) + make_rng_descriptor(seed, stream_id)


def make_u32_axpy_descriptor(sa, sb, c):
    return struct.pack("<3i", int(sa), int(sb), int(c))
