"""Pins the CPU oracle, in two layers, against golden vectors of the reference itself.  Runs without a GPU.

1. The compiled ops (morton, packbits, march, integrate, the inference loop): outputs of the reference's own CUDA
   kernels, built unmodified and run on a B200 (oracle/make_golden.py).
2. Everything the reference keeps in Python/JAX (hash-grid encoder, NeRF model, loss, optimizer, density-grid update,
   ray generation, renderer loop and op wrappers, the whole training-step forward and its gradients by finite
   differences): outputs of the reference's UNMODIFIED source executed on numpy stand-ins for jax/flax/optax
   (oracle/ref_shim.py, oracle/make_golden_{encoder,nerf,ogrid,rays,train,train_forward,train_grad,render}.py)."""
import hashlib
import os

import numpy as np
import pytest

from tests import inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated yet (run oracle/make_golden.py on a GPU box)")
    return np.load(path)


def test_morton_packbits_golden(oracle):
    g = load("morton_packbits.npz")
    rng = np.random.Generator(np.random.PCG64(0))
    xyz = rng.integers(0, 1024, (4096, 3), dtype=np.uint32)
    idx = rng.integers(0, 2 ** 30, 4096, dtype=np.uint32)
    den = rng.normal(size=4096 * 8).astype(np.float32)
    assert np.array_equal(oracle.morton3d(xyz), g["morton"])
    assert np.array_equal(oracle.morton3d_invert(idx), g["invert"])
    mask, bits = oracle.packbits(0.25, den)
    assert np.array_equal(mask, g["mask"]) and np.array_equal(bits, g["bits"])


def _canonical(out):
    nxt, exc, valid, rn, rs, idcs, xyzs, dirs, dss, zs = out
    rays = np.nonzero(rn > 0)[0]
    sel = np.concatenate([np.arange(rs[r], rs[r] + rn[r]) for r in rays]) if len(rays) else np.zeros(0, np.int64)
    return dict(next=nxt, exceeded=exc, valid=valid, n_samples=rn, xyzs=xyzs[sel], dss=dss[sel], z_vals=zs[sel],
                idcs=idcs[sel])


@pytest.mark.parametrize("case", ["scene", "cascades", "dense", "miss"])
def test_march_rays_golden_bit_exact(oracle, case):
    g = load(f"march_{case}.npz")
    st, arrays = inputs.march_case(case)
    can = _canonical(oracle.march_rays(**st, **arrays, raw=True))
    for k in ("next", "exceeded", "valid", "n_samples"):
        assert np.array_equal(can[k], g[k]), (case, k)
    for k in ("xyzs", "dss", "z_vals", "idcs"):
        full = np.ascontiguousarray(can[k])
        assert np.array_equal(full[:4096].view(np.uint8), g[k + "_head"].view(np.uint8)), (case, k)
        digest = np.frombuffer(hashlib.sha256(full.tobytes()).digest(), np.uint8)
        assert np.array_equal(digest, g[k + "_sha256"]), (case, k)  # every bit of every sample


@pytest.mark.parametrize("case", ["scene", "cascades"])
@pytest.mark.parametrize("scale", [1.0, 0.02])
def test_integrate_golden(oracle, case, scale):
    g = load(f"integrate_{case}_{scale}.npz")
    st, arrays = inputs.march_case(case)
    nxt, exc, valid, rn, rs, idcs, xyzs, dirs, dss, zs = oracle.march_rays(**st, **arrays, raw=True)
    drgbs = inputs.drgbs_for(xyzs, 21, scale)  # oracle layout == the golden file's canonical ray order
    bgs = np.random.Generator(np.random.PCG64(22)).random((rn.shape[0], 3), dtype=np.float32)
    dfin = np.random.Generator(np.random.PCG64(99)).normal(size=(rn.shape[0], 4)).astype(np.float32)
    mbs, rgbd, opac = oracle.integrate_rays(0.3, rs, rn, bgs, dss, zs, drgbs)
    # expf (oracle) vs ex2.approx (reference kernel): abs 1e-4 on colours; the composited-sample count
    # may differ for the rare ray whose transmittance sits within rounding of the 1e-4 threshold
    assert np.allclose(rgbd, g["rgbd"], atol=1e-4, rtol=0) and np.allclose(opac, g["opac"], atol=1e-4, rtol=0)
    assert abs(mbs - int(g["mbs"])) <= 4
    dbg, dz, dd = oracle.integrate_rays_backward(0.3, rs, rn, bgs, dss, zs, drgbs, g["rgbd"], g["opac"], dfin)
    k = g["dz"].shape[0]
    scale_d = max(1.0, float(np.abs(g["dd"]).max()))
    assert np.allclose(dbg, g["dbg"], atol=1e-5, rtol=1e-4)
    assert np.allclose(dz[:k], g["dz"], atol=1e-5, rtol=1e-3)
    assert np.allclose(dd[:k], g["dd"], atol=2e-4 * scale_d, rtol=1e-3)
    assert np.isclose(dz.astype(np.float64).sum(), float(g["dz_sum"]), rtol=1e-3, atol=1e-2)


def test_inference_loop_golden(oracle):
    from jaxngp_b200 import synthetic as S
    g = load("inference_loop.npz")
    st, fr, bits, n_slots = inputs.inference_case()
    N = fr["rays_o"].shape[0]
    ts = fr["t_starts"].copy()
    bg, rgbd, T = np.ones((N, 3), np.float32), np.zeros((N, 4), np.float32), np.ones(N, np.float32)
    term, idx, nri = np.ones(n_slots, np.bool_), np.zeros(n_slots, np.uint32), np.zeros(1, np.uint32)
    rendered, it, ns_total = 0, 0, 0
    while rendered < N and it < 400:
        nri, idx, ns, ts, xyzs, dss, zs, _ = oracle.march_rays_inference(
            **st, rays_o=fr["rays_o"], rays_d=fr["rays_d"], t_starts=ts, t_ends=fr["t_ends"], occupancy_bitfield=bits,
            next_ray_index_in=nri, terminated=term, indices=idx)
        x = xyzs.reshape(-1, 3)
        drgbs = np.concatenate([S.density(x)[:, None] * 0.5, S.colour(x)], -1).reshape(n_slots, -1, 4).astype(np.float32)
        cnt, term, rgbd, T = oracle.integrate_rays_inference(bg, rgbd, T, ns, idx, dss, zs, drgbs)
        rendered += cnt
        ns_total += int(ns.sum())
        it += 1
    assert it == int(g["iterations"]) and ns_total == int(g["ns_total"])
    assert np.allclose(rgbd, g["rgbd"], atol=1e-4) and np.allclose(T, g["T"], atol=1e-4)


ENCODER_CONFIGS = [(3, 2 ** 19, 2048), (2, 2 ** 19, 2 ** 19), (3, 2 ** 14, 512), (2, 2 ** 20, 2 ** 19)]  # last: models/imagefit.py:28-37


@pytest.mark.parametrize("dim,T,N_max", ENCODER_CONFIGS)
def test_hashgrid_encoder_golden_from_reference_code(oracle, dim, T, N_max):
    """The encoder oracle against the reference's OWN ``HashGridEncoder.__call__`` (models/encoders.py:82-256, run
    unmodified on numpy by oracle/make_golden_encoder.py through oracle/ref_shim.py): table geometry, level growth
    factor, and every encoded feature, for the C2/C4 shape, the 2-D imagefit shape (C1) and a small table.  Both CPU
    restatements (C and numpy) must reproduce the reference's float32 result exactly -- same operations in the same
    order -- including the points on the cube faces (Q1 spill rows)."""
    from oracle import hashgrid_np as H
    g = load("encoder_reference.npz")
    key = f"d{dim}_T{T}_N{N_max}"
    lv = H.level_table(16, T, 2, 16, N_max, dim)
    rows = int(lv["offsets"][-1])
    assert rows == int(g[key + "_rows"])  # the reference's self.param(..., (offsets[-1], F)) accepted this shape
    assert abs(float(lv["b"]) - float(g[key + "_b"])) < 1e-12
    pts, ref_enc = g[key + "_pts"], g[key + "_enc"]
    table = inputs.encoder_table(rows, 2, amp=1.0)
    enc_c = oracle.hashgrid_encode(lv, pts, 1.0, table)
    assert enc_c.dtype == np.float32 and np.array_equal(enc_c, ref_enc)
    assert np.abs(np.asarray(H.encode(lv, pts, 1.0, table), np.float32) - ref_enc).max() <= 2e-6


def test_density_grid_update_golden_from_reference_code(oracle):
    """oracle/ogrid_np.py (the checker of the CUDA grid-update kernels) against the reference's OWN
    ``OccupancyDensityGrid`` / ``NeRFState.update_ogrid_density`` / ``threshold_ogrid`` (utils/types.py:93-144,
    1149-1239; executed unmodified on numpy by oracle/make_golden_ogrid.py): two cascades, a full update followed by
    a sampled one (scripted draws, no repeated cells), thresholding through the mean-density branch after each."""
    from oracle import ogrid_np
    from oracle.make_golden_ogrid import density_fn
    g = load("ogrid_reference.npz")
    G, K, bound, steps = int(g["G"]), int(g["K"]), float(g["bound"]), int(g["steps"])
    G3 = G ** 3
    density = np.zeros(K * G3, np.float32)
    thr_max = 0.01 * steps / (2 * min(bound, 1) * 3 ** 0.5)  # utils/types.py:1367-1369
    for tag, update_all in (("all", True), ("sampled", False)):
        for cas in range(K):
            sl = slice(cas * G3, (cas + 1) * G3)
            idx = np.arange(G3, dtype=np.uint32) if update_all else np.concatenate([g[f"{tag}_c{cas}_first"], g[f"{tag}_c{cas}_second"]])
            coords = ogrid_np.sample_positions(idx, g[f"{tag}_c{cas}_jitter"], G, cas, bound)
            density[sl] = ogrid_np.decay_and_max(density[sl], idx, density_fn(coords))
        assert np.array_equal(density, g[f"{tag}_density"]), tag
        thr = ogrid_np.threshold(density[:G3], thr_max)
        assert abs(float(thr) - float(g[f"{tag}_threshold"])) <= 1e-6 * float(thr) and float(thr) < thr_max
        mask, bits = oracle.packbits(float(thr), density)
        assert np.array_equal(np.asarray(mask).astype(bool), g[f"{tag}_occ_mask"]), tag
        assert np.array_equal(np.asarray(bits), g[f"{tag}_occupancy"]), tag


def test_ray_generation_golden_from_reference_code():
    """Ray generation and the AABB near/far test against the reference's OWN code (Camera.make_ray_directions_from_
    pixel_coordinates utils/types.py:398-439, the ray construction nested in train_step app/nerf/_utils.py:96-115,
    make_rays_worldspace / make_near_far_from_bound models/renderers/cuda.py:22-97; run unmodified on numpy by
    oracle/make_golden_rays.py).  The numpy restatement that feeds every march test and smoke() is bit-exact; the torch
    host mirror (renderers.py, the checker of the fused make_training_rays kernel) agrees to one ulp."""
    import torch
    from jaxngp_b200 import renderers as R, synthetic as S
    g = load("rays_reference.npz")
    cam, tf = S.camera(), S.poses(100)
    perm = g["perm"].astype(np.int64)
    hw = cam["width"] * cam["height"]
    view, pix = perm // hw, perm % hw
    o, d = S.pixel_rays(tf, view, pix)
    ts, te = S.near_far(o, d)
    for got, key in ((o, "train_o"), (d, "train_d"), (ts, "train_t_starts"), (te, "train_t_ends")):
        assert got.dtype == np.float32 and np.array_equal(got, g[key]), key
    pt, vt = torch.from_numpy(pix), torch.from_numpy(view)
    d_cam = R.make_ray_directions(pt % cam["width"], pt // cam["width"], cam)
    tft = torch.from_numpy(tf)[vt]
    d_world = (d_cam[:, None, :] * tft[:, :9].reshape(-1, 3, 3)).sum(-1)
    ts_t, te_t = R.make_near_far_from_bound(1.0, tft[:, 9:], d_world)
    assert np.abs(d_world.numpy() - g["train_d"]).max() <= 2.4e-7
    assert np.allclose(ts_t.numpy(), g["train_t_starts"], rtol=1e-6, atol=1e-7)
    assert np.allclose(te_t.numpy(), g["train_t_ends"], rtol=1e-6, atol=1e-7)
    small = dict(width=48, height=32, fx=cam["fx"] * 48 / cam["width"], fy=cam["fy"] * 48 / cam["width"], cx=24.0, cy=16.0)
    fo, fd = R.make_rays_worldspace(small, torch.from_numpy(tf[7]))
    fts, fte = R.make_near_far_from_bound(1.0, fo, fd)
    assert np.array_equal(fo.numpy(), g["frame_o"]) and np.abs(fd.numpy() - g["frame_d"]).max() <= 2.4e-7
    assert np.allclose(fts.numpy(), g["frame_t_starts"], rtol=1e-6, atol=1e-7)
    assert np.allclose(fte.numpy(), g["frame_t_ends"], rtol=1e-6, atol=1e-7)


def test_nerf_model_golden_from_reference_code(oracle):
    """The MLP oracle (oracle/train_np.py: the checker of the fused tensor-core MLP kernels) against the reference's OWN
    ``make_nerf_ngp`` model (models/nerfs.py NeRF / CoordinateBasedMLP / trunc_exp, models/encoders.py
    SphericalHarmonicsEncoder; run unmodified on numpy by oracle/make_golden_nerf.py): layer order and widths, the
    density split, the [x | SH(dir)] concatenation, both activations, the density-only branch, the SH basis, and
    trunc_exp's backward rule.  The torch host mirror (jaxngp_b200/nerf.py sh4, trunc_exp) is checked too."""
    import torch
    from jaxngp_b200 import nerf as nerf_mod
    from oracle import hashgrid_np as H
    from oracle import train_np as T
    g = load("nerf_reference.npz")
    lv = H.level_table(16, 2 ** 19, 2, 16, 2048, 3)
    table = inputs.encoder_table(int(lv["offsets"][-1]), 2, amp=1.0)
    w = {k: g[k] for k in ("density_w0", "density_w1", "rgb_w0", "rgb_w1", "rgb_w2")}
    enc = oracle.hashgrid_encode(lv, g["xyz"], 1.0, table)
    drgbs, cache = T.mlp_forward(w, enc, g["dirs"])
    assert np.allclose(drgbs, g["drgbs"], rtol=2e-6, atol=1e-7), np.abs(drgbs - g["drgbs"]).max()
    assert np.allclose(drgbs[:, :1], g["density_only"], rtol=2e-6, atol=1e-7)
    assert np.array_equal(T.sh4(g["dirs"]), g["sh"])  # same expressions, float32 throughout
    # trunc_exp: forward exp(x), backward exp(clip(x, -15, 15)) * g (models/nerfs.py:222-238)
    x, gy = g["trunc_exp_x"], g["trunc_exp_g"]
    assert np.array_equal(np.exp(x), g["trunc_exp_fwd"])
    assert np.array_equal((np.exp(np.clip(x, -15, 15)) * gy).astype(np.float32), g["trunc_exp_grad"])
    d_drgbs = np.zeros_like(drgbs)
    d_drgbs[:, 0] = 1.0  # the oracle's backward applies the same rule to the density channel
    xt = torch.from_numpy(x.copy()).requires_grad_(True)
    nerf_mod.trunc_exp(xt).backward(torch.from_numpy(gy.copy()))
    assert np.allclose(xt.grad.numpy(), g["trunc_exp_grad"], rtol=1e-6, atol=0)
    sh_t = nerf_mod.sh4(torch.from_numpy(g["dirs"].copy())).numpy()
    assert np.abs(sh_t - g["sh"]).max() <= 2.4e-7


def test_loss_and_optimizer_golden_from_reference_code():
    """oracle/train_np.py's loss and Adam -- the checkers of huber_loss_grad / integrate_loss_fused and of the Adam kernel
    -- against the reference's OWN loss_fn (nested in train_step, app/nerf/_utils.py:117-162, with
    blend_rgba_image_array utils/data.py:443-464) and make_optimizer (app/nerf/_utils.py:19-77), run unmodified on numpy
    by oracle/make_golden_train.py (optax's primitives restated from their definitions): loss value, its gradient
    (central differences of the reference's loss), background compositing, the staircase learning-rate schedule, and
    four optimizer steps including the additive decayed-weights term on the MLP weights only."""
    from oracle import train_np as T
    g = load("train_reference.npz")
    gt, bg, pred, valid = g["gt_rgba"], g["bg"], g["pred"], g["valid"]
    target = gt[:, :3] * gt[:, 3:] + bg * (1 - gt[:, 3:])
    assert np.array_equal(target.astype(np.float32), g["blend"])
    loss, grad = T.huber_grad(pred[:, :3], target, valid)
    assert abs(loss - float(g["loss"])) <= 1e-7 * abs(loss)
    fd = g["loss_fd_grad"]
    assert np.abs(grad[:32] - fd).max() <= 2e-3 * np.abs(fd).max() + 1e-7
    assert (np.abs(fd).sum(-1) == 0).sum() == int((~valid[:32]).sum())  # masked rays carry no gradient
    opt = T.AdamNp(lr=1e-2)
    assert np.allclose([opt.lr_at(int(c)) for c in g["lr_counts"]], g["lr_values"], rtol=1e-12)
    keys = ("position_encoder", "density_mlp", "rgb_mlp")
    params = {k: g[f"opt_p0_{k}"].copy() for k in keys}
    for step in range(4):
        opt.step(params, {k: g[f"opt_g{step}_{k}"] for k in keys}, decay_keys=("density_mlp", "rgb_mlp"))
        for k in keys:
            ref = g[f"opt_p{step + 1}_{k}"]
            assert np.abs(params[k] - ref).max() <= 2e-7 * max(1.0, np.abs(ref).max()) + 1e-9, (step, k)
    # the weight-decay term is additive (optax.add_decayed_weights chained AFTER the lr-scaled Adam update) and only
    # on the MLPs: with zero gradients from the start the table would not move while the MLP weights grow by 1e-6 p
    w0 = g["opt_p0_density_mlp"]
    probe = T.AdamNp(lr=1e-2)
    p = {"density_mlp": w0.copy(), "position_encoder": g["opt_p0_position_encoder"].copy()}
    probe.step(p, {k: np.zeros_like(v) for k, v in p.items()}, decay_keys=("density_mlp",))
    assert np.array_equal(p["position_encoder"], g["opt_p0_position_encoder"])
    assert np.allclose(p["density_mlp"], w0 * np.float32(1 + 1e-6), rtol=1e-7)
    # the CUDA kernel's descriptor carries the same hyper-parameters (trainer.py)
    import struct
    from jaxngp_b200 import descriptors as D
    d = D.make_adam_descriptor(n=8, decay_begin=4, lr_init=1e-2, lr_end=1e-4, decay_rate=1 / 3, transition_steps=10_000,
                               transition_begin=10_000, staircase=True, b1=0.9, b2=0.99, eps=1e-15, eps_root=1e-15,
                               weight_decay=1e-6, grad_scale=1.0)
    assert isinstance(d, bytes) and len(d) > 0


def test_rendered_frame_golden_from_reference_code(oracle):
    """The inference path against the reference's OWN renderer: ``render_image_inference`` /
    ``march_and_integrate_inference`` (models/renderers/cuda.py:165-373: slot-refill loop with power-of-two batching,
    8192 slots x 8 steps) over the volume-rendering-jax Python wrappers (their ``.at[indices].set`` scatters, dropped
    out-of-range slots), run unmodified on numpy by oracle/make_golden_render.py with the C oracle under the two
    primitives.  Replaying the frame here with the oracle's own wrappers and a plain one-check-per-pass loop must give
    the same u8 image and depth map: that pins the wrapper semantics the oracle restates (and which the CUDA
    ``*_inplace`` ops fold into their kernels) and shows the result does not depend on how passes are batched."""
    from jaxngp_b200 import synthetic as S
    g = load("render_reference.npz")
    Wd, Hd, view = int(g["width"]), int(g["height"]), int(g["view"])
    fr = S.frame_rays(view=view, width=Wd, height=Hd)
    bits = S.occupancy_bitfield()
    N, n_slots, cap = Wd * Hd, min(8192, Wd * Hd), 8
    st = dict(diagonal_n_steps=1024, K=1, G=128, march_steps_cap=cap, bound=1.0, stepsize_portion=0.0)
    ts = fr["t_starts"].copy()
    bg, rgbd, T = np.ones((N, 3), np.float32), np.zeros((N, 4), np.float32), np.ones(N, np.float32)
    term, idx, nri = np.ones(n_slots, np.bool_), np.zeros(n_slots, np.uint32), np.zeros(1, np.uint32)
    rendered, passes = 0, 0
    while rendered < N and passes < 2000:
        nri, idx, ns, ts, xyzs, dss, zs, _ = oracle.march_rays_inference(
            **st, rays_o=fr["rays_o"], rays_d=fr["rays_d"], t_starts=ts, t_ends=fr["t_ends"], occupancy_bitfield=bits,
            next_ray_index_in=nri, terminated=term, indices=idx)
        x = xyzs.reshape(-1, 3)
        drgbs = np.concatenate([S.density(x)[:, None] * 0.5, S.colour(x)], -1).reshape(n_slots, cap, 4).astype(np.float32)
        cnt, term, rgbd, T = oracle.integrate_rays_inference(bg, rgbd, T, ns, idx, dss, zs, drgbs)
        rendered += cnt
        passes += 1
    assert rendered == N
    f32_to_u8 = lambda img: np.clip(np.round(img * 255), 0, 255).astype(np.uint8)  # utils/data.py:42-43
    assert np.array_equal(f32_to_u8(rgbd[:, :3]).reshape(Hd, Wd, 3), g["image"])
    depth = rgbd[:, 3:]
    assert np.array_equal(f32_to_u8((depth - depth.min()) / (depth.max() - depth.min() + 1e-15)).reshape(Hd, Wd), g["distance"])


def test_training_step_forward_golden_from_reference_code(oracle):
    """oracle/train_np.train_step -- the checker of the GPU training step and the CPU baseline of bench.py -- against the
    forward of one whole training step executed from the reference's OWN source (perm -> rays -> march_rays ->
    HashGridEncoder -> NeRF MLP -> integrate_rays -> Huber loss; app/nerf/_utils.py:93-162, models/renderers/cuda.py,
    the volume-rendering-jax wrappers, models/nerfs.py, models/encoders.py; oracle/make_golden_train_forward.py): the
    same rays are valid, the same number of samples is marched and composited (with early termination), same loss."""
    from jaxngp_b200 import synthetic as S
    from oracle import train_np as T
    from oracle.make_golden_train_forward import N_VIEWS, TOTAL_SAMPLES, make_inputs
    g = load("train_forward_reference.npz")
    d = make_inputs()
    cam = d["cam"]
    perm = d["perm"].astype(np.int64)
    hw = cam["width"] * cam["height"]
    o, dd = S.pixel_rays(S.poses(N_VIEWS), perm // hw, perm % hw)
    ts, te = S.near_far(o, dd)
    rays = dict(rays_o=o, rays_d=dd, t_starts=ts, t_ends=te, noises=d["noises"])
    params = dict(table=d["table"], **d["w"])
    gt = d["rgba_rows"].astype(np.float32) / np.float32(255)
    m, _ = T.train_step(params, None, d["lv"], S.occupancy_bitfield(), rays, gt, d["bg"], TOTAL_SAMPLES, apply=False)
    assert m["n_valid_rays"] == int(g["n_valid_rays"]) and 0 < m["n_valid_rays"] < perm.shape[0]
    assert m["measured_batch_size_before_compaction"] == int(g["measured_batch_size_before_compaction"])
    assert m["measured_batch_size"] == int(g["measured_batch_size"]) < m["measured_batch_size_before_compaction"]
    assert abs(m["loss"] - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))


def test_training_step_gradients_against_differences_of_the_reference_forward(oracle):
    """Analytic gradients of oracle/train_np.train_step (the checker of the CUDA backward pass) against central
    differences of the loss that the reference's OWN source computes for one training step
    (oracle/make_golden_train_grad.py).  Colour-path parameters (rgb MLP weights, density-MLP output columns 1..15):
    the reference's hand-written VJP is the true derivative there and the two agree.  Density-path parameters: the
    reference's integrate_rays_backward is NOT the derivative of its forward (integrating.cu:199-225 subtracts the
    background term of a non-terminated ray twice and rescales by min(z^2, 1)) -- the oracle and the CUDA kernels follow
    the reference's kernel (checked against its outputs on the GPU), so they must differ from the differences."""
    from jaxngp_b200 import synthetic as S
    from oracle import train_np as T
    from oracle.make_golden_train_grad import N_VIEWS, TOTAL_SAMPLES, make_inputs
    g = load("train_grad_reference.npz")
    d = make_inputs()
    cam, perm = d["cam"], d["perm"].astype(np.int64)
    hw = cam["width"] * cam["height"]
    o, dd = S.pixel_rays(S.poses(N_VIEWS), perm // hw, perm % hw)
    ts, te = S.near_far(o, dd)
    rays = dict(rays_o=o, rays_d=dd, t_starts=ts, t_ends=te, noises=d["noises"])
    gt = d["rgba_rows"].astype(np.float32) / np.float32(255)
    m, grads = T.train_step(dict(table=d["table"], **d["w"]), None, d["lv"], S.occupancy_bitfield(), rays, gt, d["bg"],
                            TOTAL_SAMPLES, apply=False)
    assert abs(m["loss"] - float(g["loss"])) <= 1e-5 * float(g["loss"])
    colour, density = 0, 0
    for name, k, fd in zip(g["names"], g["flat"], g["fd"]):
        name, k = str(name), int(k)
        an = float(np.asarray(grads[name]).reshape(-1)[k])
        colour_path = name.startswith("rgb_") or (name == "density_w1" and k % 16 != 0)
        if colour_path:
            assert abs(an - fd) <= 0.06 * abs(fd) + 2e-7, (name, k, an, fd)
            colour += 1
        else:
            assert abs(an - fd) > 0.3 * abs(fd), (name, k, an, fd)  # the reference's density VJP is not the derivative
            density += 1
    assert colour >= 7 and density >= 3


@pytest.mark.parametrize("T,N_max", [(2 ** 19, 2048), (2 ** 14, 512)])
def test_tcnn_encoder_python_half_golden_from_reference_code(T, N_max):
    """What the reference's ``TCNNHashGridEncoder.__call__`` (models/encoders.py:259-305, run unmodified by
    oracle/make_golden_encoder.py) hands to ``jaxtcnn.hashgrid_encode``: level offsets without the 8-alignment of the
    pure-JAX encoder (6,098,108 rows at C2), the per-level scale, the transposed unit-cube coordinates.  The host mirror
    (jaxngp_b200/encoders.py TCNNHashGridEncoder, the oracle's level table with align=1) must pass the same.  The CUDA
    half of that path is tiny-cuda-nn v1.6, which is not on disk: "parity unpinned" (DESIGN.md section 2)."""
    from jaxngp_b200 import encoders as E
    from oracle import hashgrid_np as H
    g = load("encoder_reference.npz")
    key = f"tcnn_T{T}_N{N_max}"
    offs, desc, coords = g[key + "_offsets"], g[key + "_desc"], g[key + "_coords_rm"]
    lt = E.make_level_table(16, T, 2, 16, N_max, 3, E.TCNNHashGridEncoder.align)
    assert list(lt.offsets) == offs.tolist() == H.level_table(16, T, 2, 16, N_max, 3, align=1)["offsets"].tolist()
    assert (int(desc[0]), int(desc[1]), int(desc[2])) == (16, 2, 16) and abs(float(desc[3]) - lt.b) < 1e-12
    pts = inputs.encoder_points(16, 3)
    assert np.array_equal(((pts + np.float32(1.0)) / np.float32(2.0)).T, coords)


def test_update_cadence_golden_from_reference_code():
    """The density-grid update cadence of the training loop (utils/types.py:1380-1396 via oracle/make_golden_ogrid.py):
    interval min(16, step // 16 + 1), full updates for the first 256 steps.  trainer.Trainer mirrors the properties."""
    from jaxngp_b200.trainer import Trainer
    g = load("ogrid_reference.npz")
    probe = Trainer.__new__(Trainer)  # the properties only read .step
    for step, interval, call, all_cells in zip(g["cadence_steps"], g["cadence_interval"], g["cadence_call"], g["cadence_all"]):
        probe.step = int(step)
        assert probe.update_ogrid_interval == int(interval)
        assert probe.should_call_update_ogrid == bool(call) and probe.should_update_all_ogrid_cells == bool(all_cells)
    assert int(g["cadence_interval"][240]) == 16 and bool(g["cadence_call"][256])


def test_op_contracts_golden_from_reference_code():
    """Abstract-evaluation contracts: the host mirror (jaxngp_b200.volrendjax.*) answers a table of well-formed and
    malformed operand signatures the way the reference's OWN ``*_abstract`` rules do (volume-rendering-jax
    {marching,integrating,packbits,morton3d}/abstract.py and jax-tcnn hashgrid_tcnn/abstract.py, run unmodified by
    oracle/make_golden_contracts.py; jaxngp_b200.jaxtcnn is the mirror of the latter): the same
    exception class for every malformed case; for the well-formed ones the checks pass and the call reaches the launch,
    which refuses CPU tensors (there is no CPU path)."""
    import json
    import torch
    from jaxngp_b200 import _lib, volrendjax as V
    table = json.load(open(os.path.join(GOLDEN, "contracts_reference.json")))
    tdt = {"float32": torch.float32, "float16": torch.float16, "uint32": torch.int32, "int32": torch.int32,
           "uint8": torch.uint8, "int8": torch.int8, "bool": torch.bool}

    def tensor(shape, dtype):
        if dtype == "int32" and False:
            pass
        return torch.zeros(tuple(shape), dtype=tdt[dtype])

    def call(op, ops, st):
        a = [tensor(s, d) for s, d in ops]
        if op == "march_rays_abstract":
            return V.march_rays(st["total_samples"], st["diagonal_n_steps"], st["K"], st["G"], st["bound"], st["stepsize_portion"], *a)
        if op == "march_rays_inference_abstract":
            return V.march_rays_inference(st["diagonal_n_steps"], st["K"], st["G"], st["march_steps_cap"], st["bound"],
                                          st["stepsize_portion"], *a)
        if op == "integrate_rays_abstract":
            return V.integrate_rays(0.3, *a)
        if op == "integrate_rays_inference_abstract":
            return V.integrate_rays_inference(*a)
        if op == "pack_density_into_bits_abstract":
            return V.packbits(*a)
        if op == "morton3d_abstract":
            return V.morton3d(*a)
        if op == "morton3d_invert_abstract":
            return V.morton3d_invert(*a)
        if op == "hashgrid_encode_abstract":
            from jaxngp_b200 import jaxtcnn
            return jaxtcnn.hashgrid_encode(jaxtcnn.HashGridMetadata(**st), *a)
        raise KeyError(op)

    seen = set()
    for row in table:
        # an int32 bitfield is the reference's "wrong dtype" case; uint32 operands travel as int32 in the mirror
        expected = row["answer"].get("raises", "NgpError")
        if row["case"] == "integrate bgs shape":
            # the abstract rule never sees this operand: the reference's wrapper broadcasts bgs to (n_rays, 3) first
            # (integrating/impl.py:60) and jnp.broadcast_to raises ValueError; the mirror does the same
            expected = "ValueError"
        try:
            call(row["op"], row["operands"], row["static"])
            got = "no error"
        except Exception as exc:  # noqa: BLE001 -- the class is what is compared
            got = type(exc).__name__
        assert got == expected, (row["case"], got, expected)
        seen.add(expected)
    assert {"AssertionError", "NotImplementedError", "ValueError", "RuntimeError", "NgpError"} <= seen
    assert _lib.NgpError.__name__ == "NgpError"


def test_mark_untrained_density_grid_golden_from_reference_code(oracle):
    """The one-time camera-visibility culling (utils/types.py:1241-1362) against the reference's OWN method, run
    unmodified by oracle/make_golden_mark_untrained.py: the numpy restatement and the torch host code that the product
    runs (device-agnostic integer / float32 ops, here on the CPU) reproduce the alive marker, the culled densities, the
    alive index table and -- through the restated thresholds -- the occupancy of both a fresh (step 0) and a trained state."""
    import torch
    from jaxngp_b200 import ogrid as product_ogrid
    from jaxngp_b200 import synthetic as S
    from oracle import ogrid_np
    g = load("mark_untrained_reference.npz")
    G, K, bound = int(g["G"]), int(g["K"]), float(g["bound"])
    G3 = G ** 3
    cam = S.camera()
    alive_ref = np.zeros(K * G3, bool)
    alive_ref[g["step0_alive_indices"]] = True
    assert np.array_equal(g["step0_alive_indices"], g["step300_alive_indices"])
    assert np.array_equal(np.diff(g["step0_alive_indices_offset"]), alive_ref.reshape(K, G3).sum(1))
    alive_np = ogrid_np.visible_cells(K, G, bound, g["poses"], cam)
    assert np.array_equal(alive_np, alive_ref)
    alive_t = product_ogrid.visible_cells(K, G, bound, torch.from_numpy(g["poses"]), cam).numpy()
    assert np.array_equal(alive_t, alive_ref)
    thr_max = 0.01 * 1024 / (2 * min(bound, 1) * 3 ** 0.5)
    for step in (0, 300):
        marked, mask, bits, alive_idx = ogrid_np.mark_untrained(g["density_in"], alive_np, step, thr_max, G3)
        assert np.array_equal(marked, g[f"step{step}_density"]) and np.array_equal(alive_idx, g[f"step{step}_alive_indices"])
        assert np.array_equal(mask, g[f"step{step}_occ_mask"]) and np.array_equal(bits, g[f"step{step}_occupancy"])
    assert np.array_equal(g["step0_occ_mask"], alive_ref)  # threshold -0.5: every trainable cell starts occupied
    # Morton inversion of the host code = the compiled op's
    idx = np.arange(0, G3, 7, dtype=np.uint32)
    assert np.array_equal(product_ogrid._morton3d_invert_host(torch.from_numpy(idx.astype(np.int64))).numpy(),
                          oracle.morton3d_invert(idx).astype(np.int64))


@pytest.mark.parametrize("dim,T,N_max", [(3, 2 ** 14, 512), (2, 2 ** 12, 256)])
def test_total_variation_branch_golden_from_reference_code(dim, T, N_max):
    """The encoder's optional regulariser (models/encoders.py:234-254, ``tv_scale > 0``; off in make_nerf_ngp) from the
    reference's unmodified ``HashGridEncoder.__call__``: the host module's torch restatement (index ops, not a kernel:
    it is off the hot path) gives the same value."""
    import torch
    from jaxngp_b200 import encoders as E
    from oracle import hashgrid_np as H
    g = load("encoder_reference.npz")
    lv = H.level_table(16, T, 2, 16, N_max, dim)
    table = inputs.encoder_table(int(lv["offsets"][-1]), 2, amp=1.0)
    module = E.HashGridEncoder(16, T, 2, 16, N_max, tv_scale=0.25, dim=dim, device="cpu")
    with torch.no_grad():
        module.latents.copy_(torch.from_numpy(table))
        tv = module.tv_scale * module._total_variation(torch.from_numpy(inputs.encoder_points(64, dim)), 1.0)
    ref = float(g[f"tv_d{dim}_T{T}_N{N_max}"])
    assert ref > 0 and abs(float(tv) - ref) <= 1e-6 * ref
