"""BASELINE-size runs checked through size-independent properties (the oracle would take minutes at these sizes):
C2 march (2^18 rays, 2^18 sample budget), C4 encoder (2^22 points), C2 integrate, C3 frame."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_march_rays_c2_invariants():
    """march_rays at the C2 size: the compaction is an exclusive prefix sum in ray order, the budget rule of
    marching.cu:135,205-209 holds, every sample lies on its ray inside an occupied cell of the bitfield, unused
    slots are zero (marching/__init__.py:60-68), and a second run returns the same bits (deterministic)."""
    from jaxngp_b200 import synthetic as S, volrendjax as V
    n = 1 << 18
    r = S.training_rays(n, seed=1000000007)
    bits_np = S.occupancy_bitfield()
    args = dict(total_samples=1 << 18, diagonal_n_steps=1024, K=1, G=128, bound=1.0, stepsize_portion=0.0,
                rays_o=torch.from_numpy(r["rays_o"]).to(DEV), rays_d=torch.from_numpy(r["rays_d"]).to(DEV),
                t_starts=torch.from_numpy(r["t_starts"]).to(DEV), t_ends=torch.from_numpy(r["t_ends"]).to(DEV),
                noises=torch.from_numpy(r["noises"]).to(DEV), occupancy_bitfield=torch.from_numpy(bits_np).to(DEV), raw=True)
    out = V.march_rays(**args)
    nxt, exc, valid, rn, rs, idcs, xyzs, dirs, dss, zs = out
    nxt, exc = int(nxt[0]) & 0xFFFFFFFF, int(exc[0]) & 0xFFFFFFFF
    used = nxt - exc
    rn64, rs64 = rn.long() & 0xFFFFFFFF, rs.long() & 0xFFFFFFFF
    assert 0 < used <= 1 << 18 and int(rn64.sum()) == used
    got = rn64 > 0
    assert torch.equal(rs64[got], (torch.cumsum(rn64, 0) - rn64)[got])          # ray-order exclusive prefix
    assert bool(valid[got].all())                                                 # rays with samples are valid
    last = int(torch.nonzero(got).max())
    if exc:  # the first ray that did not fit: invalid, and every later ray early-outs (marching.cu:135)
        assert not bool(valid[last + 1:].any()) and int(rn64[last + 1:].sum()) == 0
    # samples: on the ray, inside an occupied cell
    i = idcs[:used].long()
    o, d = args["rays_o"][i], args["rays_d"][i]
    assert torch.allclose(xyzs[:used], o + zs[:used, None] * d, atol=2e-6)
    assert torch.equal(dirs[:used], d)
    assert bool((zs[:used] >= args["t_starts"][i]).all()) and bool((zs[:used] < args["t_ends"][i]).all())
    g = ((xyzs[:used] + 1) * 0.5 * 128).floor().clamp(0, 127).int()
    cell = V.morton3d(g.contiguous()).long() & 0xFFFFFFFF
    occ = (torch.from_numpy(bits_np).to(DEV)[cell >> 3].int() >> (cell & 7).int()) & 1
    assert bool(occ.all())
    assert float(dss[:used].min()) == float(dss[:used].max()) == pytest.approx(2 * 3 ** 0.5 / 1024, rel=1e-6)
    for t in (idcs, xyzs, dirs, dss, zs):
        assert not bool(t[used:].any())                                           # zero-filled tail
    again = V.march_rays(**args)
    for a, b in zip(out, again):
        assert torch.equal(a, b)


def test_hashgrid_c4_linearity_and_adjointness():
    """Encoder at the C4 size (2^22 points, T = 2^19): forward is linear in the table, and the backward is its
    adjoint: <enc(table), d_enc> == <table, backward(d_enc)> (a checksum of every gather against every scatter)."""
    from jaxngp_b200 import encoders as E
    n = 1 << 22
    g = torch.Generator(device=DEV)
    lt = E.make_level_table(16, 2 ** 19, 2, 16, 2048, 3)
    pos = torch.rand(n, 3, device=DEV, generator=g.manual_seed(42)) * 2 - 1
    ta = torch.rand(lt.rows, 2, device=DEV, generator=g.manual_seed(43)) * 2 - 1
    tb = torch.rand(lt.rows, 2, device=DEV, generator=g.manual_seed(45)) * 2 - 1
    ea, eb = E.hashgrid_forward(lt, pos, 1.0, ta), E.hashgrid_forward(lt, pos, 1.0, tb)
    eab = E.hashgrid_forward(lt, pos, 1.0, 0.5 * ta - 2.0 * tb)
    assert torch.allclose(eab, 0.5 * ea - 2.0 * eb, atol=2e-5)
    assert float(ea.abs().max()) <= 1.0 + 1e-5                                    # convex combination of |rows| <= 1
    d_enc = torch.randn(n, 32, device=DEV, generator=g.manual_seed(44))
    grad = E.hashgrid_backward(lt, pos, 1.0, d_enc)
    lhs = (ea.double() * d_enc.double()).sum()
    rhs = (ta.double() * grad.double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-4 * float(lhs.abs() + rhs.abs() + 1)
    # constant table -> every feature equals the constant (the 8 weights sum to one)
    ones = E.hashgrid_forward(lt, pos[: 1 << 16], 1.0, torch.full((lt.rows, 2), 0.25, device=DEV))
    assert torch.allclose(ones, torch.full_like(ones, 0.25), atol=1e-6)


def test_integrate_rays_c2_bounds_and_gradient_check():
    """integrate_rays on a C2-sized march: opacity in [0, 1], colour within the hull of the samples' colours and
    the background, and the backward kernel's colour gradients agree with a finite difference of the forward along a
    random direction (a checksum over all 2^18 samples)."""
    from jaxngp_b200 import synthetic as S, volrendjax as V
    from jaxngp_b200.volrendjax.integrating import _integrate_bwd, _integrate_fwd
    n = 1 << 18
    r = S.training_rays(n, seed=11)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)  # noqa: E731
    out = V.march_rays(1 << 18, 1024, 1, 128, 1.0, 0.0, t(r["rays_o"]), t(r["rays_d"]), t(r["t_starts"]), t(r["t_ends"]),
                       t(r["noises"]), t(S.occupancy_bitfield()), raw=True)
    _, _, _, rn, rs, _, xyzs, _, dss, zs = out
    g = torch.Generator(device=DEV).manual_seed(1)
    drgbs = torch.cat([torch.rand(xyzs.shape[0], 1, device=DEV, generator=g) * 8, torch.rand(xyzs.shape[0], 3, device=DEV, generator=g)], -1).contiguous()
    bg = torch.rand(n, 3, device=DEV, generator=g)
    mbs, rgbd, opac = _integrate_fwd(rs, rn, bg, dss, zs, drgbs)
    assert 0 < int(mbs[0]) <= 1 << 18
    assert float(opac.min()) >= 0.0 and float(opac.max()) <= 1.0 + 1e-6
    assert float(rgbd[:, :3].min()) >= -1e-6 and float(rgbd[:, :3].max()) <= 1.0 + 1e-4
    none = (rn.long() & 0xFFFFFFFF) == 0
    assert torch.equal(rgbd[none, :3], bg[none]) and not bool(opac[none].any())   # empty rays show the background
    w = torch.randn(n, 4, device=DEV, generator=g)
    _, _, d_drgbs = _integrate_bwd(0.0, rs, rn, bg, dss, zs, drgbs, rgbd, opac, w.contiguous())
    # the colour path is linear (weight * dL/dfinal, integrating.cu:225-232); the density gradient carries the
    # reference's min(z^2, 1) scaling and is checked against the oracle in test_gpu_parity.py instead
    v = torch.randn_like(drgbs)
    v[:, 0] = 0
    eps = 1e-2
    _, rp, _ = _integrate_fwd(rs, rn, bg, dss, zs, (drgbs + eps * v).contiguous())
    _, rm, _ = _integrate_fwd(rs, rn, bg, dss, zs, (drgbs - eps * v).contiguous())
    fd = ((rp.double() - rm.double()) * w.double()).sum() / (2 * eps)
    an = (d_drgbs.double() * v.double()).sum()
    assert abs(float(fd - an)) <= 2e-2 * float(fd.abs() + an.abs() + 1)


def test_frame_c3_fast_renderer_equals_reference_loop():
    """C3 (800x800, 640,000 rays): the graph renderer (skip-empty pre-pass, in-place ops, fused encoder+MLP, 262144
    slots x 16 steps) against the reference's slot-refill loop at its own defaults (8192 slots x 8 steps, op by op,
    unfused) on the same rays: every pixel of the u8 image identical, every ray rendered exactly once."""
    from jaxngp_b200 import renderers
    from jaxngp_b200.trainer import Scene, Trainer
    scene = Scene(DEV, n_views=4)
    tr = Trainer(device=DEV, scene=scene)
    gen = torch.Generator(device=DEV).manual_seed(0)
    for it in range(64):
        tr.train_step(torch.randint(0, scene.n_pixels, (tr.n_rays,), device=DEV, generator=gen, dtype=torch.int32))
        if (it + 1) % 16 == 0:
            tr.update_ogrid()
    pose = scene.transforms[2]
    ref_rgb, ref_depth = renderers.render_image_inference(tr.nerf, scene.cam, pose, tr.occupancy, grouped=False)
    o, d = renderers.make_rays_worldspace(scene.cam, pose)
    ts, te = renderers.make_near_far_from_bound(1.0, o, d)
    tr.nerf.grouped_impl = "mma"  # same arithmetic as the unfused ops: same bits
    R = renderers.InferenceRenderer(tr.nerf, scene.cam, tr.occupancy)
    rgb, depth = R.render_rays(o, d, ts, te)
    assert int(R.counters[0]) == 640000
    assert torch.equal(rgb.reshape(ref_rgb.shape), ref_rgb)
    assert torch.allclose(depth.reshape(ref_depth.shape), ref_depth, atol=1e-5)
    assert float((ref_rgb.float().mean())) < 250  # the object is visible (not an all-background frame)
    # default: dense layers on tcgen05 (f32 accumulation order differs): at most one grey level on a few pixels
    tr.nerf.grouped_impl = "umma"
    R2 = renderers.InferenceRenderer(tr.nerf, scene.cam, tr.occupancy, persistent=False)
    rgb2, depth2 = R2.render_rays(o, d, ts, te)
    assert int(R2.counters[0]) == 640000
    diff = (rgb2.reshape(ref_rgb.shape).int() - ref_rgb.int()).abs()
    assert int(diff.max()) <= 1 and float((diff > 0).float().mean()) < 1e-2, (int(diff.max()), float((diff > 0).float().mean()))
    assert torch.allclose(depth2.reshape(ref_depth.shape), ref_depth, atol=1e-3)
    # the whole loop as ONE persistent kernel (the default renderer): same march code, same encoder + MLP code, same
    # compositing expressions as the tcgen05 loop above -- every ray once, the same samples, the same image
    R3 = renderers.InferenceRenderer(tr.nerf, scene.cam, tr.occupancy)
    assert R3.persistent
    for _ in range(2):  # second frame: counters and the ray ticket start over
        rgb3, depth3 = R3.render_rays(o, d, ts, te)
        assert int(R3.counters[0]) == 640000
        assert int(R3.counters[1]) == int(R2.counters[1])  # same samples marched: same termination decisions
        assert torch.equal(rgb3, rgb2)
        assert torch.allclose(depth3, depth2, atol=1e-6)
    # rays it owns only (tile sharding): a strided subset renders the same pixels
    pix = torch.arange(0, 640000, 7, device=DEV, dtype=torch.int32)
    R4 = renderers.InferenceRenderer(tr.nerf, scene.cam, tr.occupancy, pixel_indices=pix)
    rgb4, _ = R4.render(pose)
    R5 = renderers.InferenceRenderer(tr.nerf, scene.cam, tr.occupancy, pixel_indices=pix, persistent=False)
    rgb5, _ = R5.render(pose)
    assert torch.equal(rgb4, rgb5) and int(R4.counters[0]) == pix.numel()
