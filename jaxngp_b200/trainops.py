"""Fused training glue (csrc/trainops.cu): ray generation and the Huber loss with its gradient."""
import torch

from . import _lib, descriptors


def make_training_rays(perm: torch.Tensor, transforms: torch.Tensor, cam: dict, bound: float):
    """perm int32[n] -> (rays_o [n,3], rays_d [n,3], t_starts [n], t_ends [n]); app/nerf/_utils.py:93-115 +
    models/renderers/cuda.py:57-97 for an undistorted perspective camera."""
    n, dev = perm.shape[0], perm.device
    o = torch.empty(n, 3, dtype=torch.float32, device=dev)
    d = torch.empty(n, 3, dtype=torch.float32, device=dev)
    ts = torch.empty(n, dtype=torch.float32, device=dev)
    te = torch.empty(n, dtype=torch.float32, device=dev)
    if n:
        _lib.call("ngp_make_training_rays", [perm, transforms, o, d, ts, te],
                  descriptors.make_training_rays_descriptor(n, cam["width"], cam["height"], transforms.shape[0],
                                                            cam["fx"], cam["fy"], cam["cx"], cam["cy"], bound))
    return o, d, ts, te


#: Philox stream ids (NgpRngDescriptor.stream_id): independent consumers of one seed
STREAM_TRAIN_RAYS, STREAM_OGRID = 1, 2


def new_rng_state(device, counter=0):
    """Device-resident call counter of the ops that draw random numbers: int32[2] = {counter, ticket}."""
    return torch.tensor([counter, 0], dtype=torch.int32, device=device)


def philox_uniform(n: int, counter: int, seed: int, stream_id: int, device) -> torch.Tensor:
    """f32 [n, 4]: the four uniforms every element of call `counter` draws (csrc/common.cuh philox_uniform4)."""
    out = torch.empty(n, 4, dtype=torch.float32, device=device)
    if n:
        _lib.call("ngp_philox_uniform", [out], descriptors.make_philox_descriptor(n, counter, seed, stream_id))
    return out


def make_training_rays_rng(perm: torch.Tensor, transforms: torch.Tensor, cam: dict, bound: float, rng_state: torch.Tensor,
                           seed: int, stream_id: int = STREAM_TRAIN_RAYS):
    """``make_training_rays`` plus the step's random inputs from the same launch: noises [n] (march perturbations,
    models/renderers/cuda.py:118-122) and bgs [n, 3] (random backgrounds, app/nerf/_utils.py:134-136) = the Philox
    uniforms of call ``rng_state[0]``, which the launch increments."""
    n, dev = perm.shape[0], perm.device
    o = torch.empty(n, 3, dtype=torch.float32, device=dev)
    d = torch.empty(n, 3, dtype=torch.float32, device=dev)
    ts = torch.empty(n, dtype=torch.float32, device=dev)
    te = torch.empty(n, dtype=torch.float32, device=dev)
    noises = torch.empty(n, dtype=torch.float32, device=dev)
    bgs = torch.empty(n, 3, dtype=torch.float32, device=dev)
    if n:
        _lib.call("ngp_make_training_rays_rng", [perm, transforms, rng_state, o, d, ts, te, noises, bgs],
                  descriptors.make_training_rays_rng_descriptor(n, cam["width"], cam["height"], transforms.shape[0], cam["fx"],
                                                                cam["fy"], cam["cx"], cam["cy"], bound, seed, stream_id))
    return o, d, ts, te, noises, bgs


def u32_axpy(a: torch.Tensor, b: torch.Tensor, sa: int, sb: int, c: int, out: torch.Tensor = None) -> torch.Tensor:
    """out[0] = sa * a[0] + sb * b[0] + c on device scalars (int32 tensors carrying uint32 bits), one 1-thread launch."""
    if out is None:
        out = torch.empty(1, dtype=torch.int32, device=a.device)
    _lib.call("ngp_u32_axpy", [a, b, out], descriptors.make_u32_axpy_descriptor(sa, sb, c))
    return out


def huber_loss_grad(final_rgbds, ray_is_valid, perm, rgbas_u8, bgs, delta=0.1):
    """Returns (dL_dfinal_rgbds [n,4], loss [1], n_valid_rays int32[1]); app/nerf/_utils.py:151-165."""
    n, dev = final_rgbds.shape[0], final_rgbds.device
    dL = torch.empty(n, 4, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    n_valid = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.call("ngp_huber_loss_grad", [final_rgbds, ray_is_valid, perm, rgbas_u8, bgs, dL, loss, n_valid],
              descriptors.make_huber_loss_descriptor(n, delta))
    return dL, loss, n_valid


def integrate_loss_fused(near_distance, rays_sample_startidx, rays_n_samples, bgs, dss, z_vals, drgbs, ray_is_valid, perm,
                         rgbas_u8, delta=0.1):
    """integrate_rays -> Huber loss -> integrate_rays_backward in one launch (csrc/integrating.cu).  Returns
    (measured_batch_size [1], final_rgbds [n,4], final_opacities [n], dL_ddrgbs [S,4], loss [1], n_valid_rays [1])."""
    n, S, dev = bgs.shape[0], drgbs.shape[0], drgbs.device
    scalars = torch.empty(4, dtype=torch.int32, device=dev)  # [mbs | n_valid | loss]: one zero-fill for the three
    mbs, n_valid, loss = scalars[0:1], scalars[1:2], scalars[2:3].view(torch.float32)
    fin = torch.empty(n, 4, dtype=torch.float32, device=dev)
    opac = torch.empty(n, dtype=torch.float32, device=dev)
    d_drgbs = torch.empty(S, 4, dtype=torch.float32, device=dev)
    _lib.call("ngp_integrate_loss_fused",
              [rays_sample_startidx, rays_n_samples, bgs, dss, z_vals, drgbs, ray_is_valid, perm, rgbas_u8,
               mbs, fin, opac, d_drgbs, loss, n_valid],
              descriptors.make_integrate_loss_descriptor(n, S, near_distance, delta))
    return mbs, fin, opac, d_drgbs, loss, n_valid
