"""End-to-end parity of the GPU training step (trainer.py) against the CPU oracle's training step
(oracle/train_np.py), and convergence of a short training run on the procedural scene."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def small_scene():
    from jaxngp_b200.trainer import Scene
    return Scene(DEV, n_views=4, width=100, height=100)


def _params_np(tr):
    n = tr.nerf
    d = dict(table=tr.table.detach().cpu().numpy().copy())
    for k in ("density_w0", "density_w1", "rgb_w0", "rgb_w1", "rgb_w2"):
        d[k] = getattr(n, k).detach().cpu().numpy().copy()
    return d


def test_fused_ray_generation_matches_host_restatement(small_scene):
    from jaxngp_b200 import synthetic as S, trainops
    rng = np.random.Generator(np.random.PCG64(4))
    perm = rng.integers(0, small_scene.n_pixels, 50000).astype(np.int32)
    o, d, ts, te = trainops.make_training_rays(torch.from_numpy(perm).to(DEV), small_scene.transforms, small_scene.cam, 1.0)
    hw = small_scene.width * small_scene.height
    ro, rd = S.pixel_rays(small_scene.transforms.cpu().numpy(), perm // hw, perm % hw, small_scene.cam)
    rts, rte = S.near_far(ro, rd)
    assert np.array_equal(o.cpu().numpy(), ro)
    assert np.allclose(d.cpu().numpy(), rd, atol=2e-7, rtol=0)
    assert np.allclose(ts.cpu().numpy(), rts, rtol=1e-5, atol=1e-6) and np.allclose(te.cpu().numpy(), rte, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("fused_mlp", [True, False])
def test_train_step_gradients_match_oracle(small_scene, fused_mlp):
    from jaxngp_b200 import synthetic as S
    from jaxngp_b200.trainer import Trainer
    from oracle import hashgrid_np as H, train_np as T
    n_rays, total = 4096, 1 << 16
    tr = Trainer(device=DEV, n_rays=n_rays, total_samples=total, scene=small_scene, use_graph=False, fused_mlp=fused_mlp)
    tr.occupancy.copy_(small_scene.bitfield_gt)
    # a table with visible features so the comparison is not dominated by zeros
    tr.table.uniform_(-0.5, 0.5, generator=torch.Generator(device=DEV).manual_seed(3))
    rng = np.random.Generator(np.random.PCG64(12))
    perm = rng.integers(0, small_scene.n_pixels, n_rays).astype(np.int32)
    noises = rng.random(n_rays, dtype=np.float32)
    bg = rng.random((n_rays, 3), dtype=np.float32)
    params = _params_np(tr)
    out = tr._step_body(torch.from_numpy(perm).to(DEV), torch.from_numpy(noises).to(DEV), torch.from_numpy(bg).to(DEV),
                        apply=False)
    o, d = small_scene.rays(torch.from_numpy(perm).to(DEV).long())
    o, d = o.cpu().numpy(), d.cpu().numpy()
    ts, te = S.near_far(o, d)
    rays = dict(rays_o=o, rays_d=d, t_starts=ts, t_ends=te, noises=noises)
    gt = small_scene.rgbas_u8[torch.from_numpy(perm).to(DEV).long()].cpu().numpy().astype(np.float32) / 255
    lv = H.level_table(16, 2 ** 19, 2, 16, 2048, 3)
    m, grads = T.train_step(params, T.AdamNp(), lv, small_scene.bitfield_gt.cpu().numpy(), rays, gt, bg, total, apply=False)
    # integer outputs: exact
    assert int(out["measured_batch_size_before_compaction"]) == m["measured_batch_size_before_compaction"]
    assert int(out["n_valid_rays"]) == m["n_valid_rays"]
    assert abs(int(out["measured_batch_size"]) - m["measured_batch_size"]) <= 4
    assert np.isclose(float(out["loss"]), m["loss"], rtol=2e-3)
    # gradients: table rel 1e-2 (north star), MLP weights rel 2e-2 of the largest entry (TF32 matmuls)
    g_table = tr.table_grad.cpu().numpy()
    assert np.abs(g_table - grads["table"]).max() <= 1e-2 * np.abs(grads["table"]).max()
    off = 0
    from jaxngp_b200 import nerf as nerf_mod
    for k, i, o in nerf_mod.MLP_SHAPES:
        got = tr.mlp_grad[off:off + i * o].view(i, o).cpu().numpy()
        off += i * o
        ref = grads[k]
        assert np.abs(got - ref).max() <= 2e-2 * np.abs(ref).max(), k


def test_adam_kernel_matches_reference_optimizer():
    from jaxngp_b200 import _lib, descriptors
    from oracle import train_np as T
    rng = np.random.Generator(np.random.PCG64(0))
    n, split = 4096, 1024
    p0 = rng.normal(size=n).astype(np.float32)
    params = {"a": p0[:split].copy(), "b": p0[split:].copy()}
    opt = T.AdamNp(lr=1e-2)
    p = torch.from_numpy(p0).to(DEV)
    m, v, step = torch.zeros_like(p), torch.zeros_like(p), torch.zeros(1, dtype=torch.int32, device=DEV)
    desc = descriptors.make_adam_descriptor(n, split, 1e-2, 1e-4, 1 / 3, 10_000, 10_000, True, 0.9, 0.99, 1e-15, 1e-15, 1e-6)
    for it in range(5):
        g = rng.normal(size=n).astype(np.float32) * (it != 2)  # one all-zero gradient step: eps=1e-15 path
        opt.step(params, {"a": g[:split], "b": g[split:]}, decay_keys=("b",))
        _lib.call("ngp_adam_step", [step, p, torch.from_numpy(g).to(DEV), m, v], desc)
        step += 1
    ref = np.concatenate([params["a"], params["b"]])
    assert np.allclose(p.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)


def test_short_training_run_converges(small_scene):
    from jaxngp_b200.trainer import Trainer
    n_rays = 1 << 14
    tr = Trainer(device=DEV, n_rays=n_rays, total_samples=1 << 17, scene=small_scene, use_graph=True)
    gen = torch.Generator(device=DEV).manual_seed(1)
    losses = []
    for it in range(200):
        perm = torch.randint(0, small_scene.n_pixels, (n_rays,), device=DEV, generator=gen, dtype=torch.int32)
        out = tr.train_step(perm)
        if (it + 1) % 16 == 0:
            tr.update_ogrid()
        if it % 20 == 0 or it == 199:
            losses.append(float(out["loss"]))
    assert np.isfinite(losses).all()
    assert losses[-1] < 0.35 * losses[0], losses  # Huber loss drops by > 3x in 200 steps
    occ = float(tr.occ_mask.float().mean())
    assert 0.0 < occ < 0.6, occ  # the learned occupancy grid has pruned most of the empty space


def test_fused_mlp_matches_fp32_reference():
    """csrc/mlp.cu (TF32 tensor cores) against a plain PyTorch fp32 reference of the same network
    (models/nerfs.py:27-128): forward abs/rel 2e-3 (TF32 has 10 mantissa bits), gradients 1e-2 of scale."""
    from jaxngp_b200 import nerf as nerf_mod
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator(device=DEV).manual_seed(5)
    for n in (1, 15, 16, 129, 5000):
        model = nerf_mod.NeRF(bound=1.0, device=DEV, generator=gen, T=2 ** 14)
        w = model.mlp_flat.detach().clone().requires_grad_(True)
        enc = (torch.randn(n, 32, device=DEV, generator=gen) * 0.5).requires_grad_(True)
        dirs = torch.nn.functional.normalize(torch.randn(n, 3, device=DEV, generator=gen), dim=-1)
        views, off = {}, 0
        for k, i, o in nerf_mod.MLP_SHAPES:
            views[k] = w[off:off + i * o].view(i, o)
            off += i * o
        x = torch.relu(enc @ views["density_w0"]) @ views["density_w1"]
        density = nerf_mod.trunc_exp(x[:, :1])
        h = torch.cat([x, nerf_mod.sh4(dirs)], dim=-1)
        rgb = torch.sigmoid(torch.relu(torch.relu(h @ views["rgb_w0"]) @ views["rgb_w1"]) @ views["rgb_w2"])
        ref = torch.cat([density, rgb], dim=-1)
        got = nerf_mod.mlp_forward(enc.detach().contiguous(), dirs, w.detach())
        assert torch.allclose(got, ref, rtol=3e-3, atol=3e-3), (n, (got - ref).abs().max())
        dens = nerf_mod.mlp_forward(enc.detach().contiguous(), None, w.detach())
        assert torch.allclose(dens, ref[:, 0], rtol=3e-3, atol=3e-3)
        d_out = torch.randn(n, 4, device=DEV, generator=gen)
        g_enc_ref, g_w_ref = torch.autograd.grad(ref, [enc, w], d_out)
        g_enc, g_w = nerf_mod.mlp_backward(enc.detach().contiguous(), dirs, w.detach(), d_out)
        # a TF32-rounded pre-activation within ~1e-3 of zero can land on the other side of a ReLU, which changes
        # that sample's input gradient by O(1): judge d_enc by its relative Frobenius error and by how
        # many entries deviate, not by the single worst entry
        err = (g_enc - g_enc_ref).abs()
        assert torch.linalg.norm(g_enc - g_enc_ref) <= 3e-2 * torch.linalg.norm(g_enc_ref) + 1e-6, n
        assert (err > 1e-2 * g_enc_ref.abs().max()).float().mean() <= 5e-3, n
        assert (g_w - g_w_ref).abs().max() <= 1e-2 * g_w_ref.abs().max() + 1e-6, n
    torch.backends.cuda.matmul.allow_tf32 = True


def test_fused_integrate_loss_matches_the_three_ops(small_scene):
    """ngp_integrate_loss_fused against integrate_rays -> huber_loss_grad -> integrate_rays_backward on a marched
    batch (rays with and without samples, early-terminated rays, an overflowing budget): same bits for the composited
    colours, opacities and sample gradients, same counts, loss to summation-order error."""
    from jaxngp_b200 import synthetic, trainops
    from jaxngp_b200.trainer import Trainer
    from jaxngp_b200.volrendjax.integrating import _integrate_bwd, _integrate_fwd
    for n_rays, budget in ((1 << 14, 1 << 17), (1 << 14, 1 << 13)):
        tr = Trainer(device=DEV, n_rays=n_rays, total_samples=budget, scene=small_scene, use_graph=False)
        tr.grid.occupancy.copy_(small_scene.bitfield_gt)
        gen = torch.Generator(device=DEV).manual_seed(2)
        perm = torch.randint(0, small_scene.n_pixels, (n_rays,), device=DEV, generator=gen, dtype=torch.int32)
        _, _, valid, rn, rs, _, xyzs, dirs, dss, zs, bg = tr._march_body(perm)
        S = xyzs.shape[0]
        drgbs = torch.rand(S, 4, device=DEV, generator=gen) * torch.tensor([40.0, 1.0, 1.0, 1.0], device=DEV)
        drgbs[: S // 3, 0] *= 0.01  # thin media: rays that never saturate
        eff, fin, opac = _integrate_fwd(rs, rn, bg, dss, zs, drgbs)
        d_fin, loss, nv = trainops.huber_loss_grad(fin, valid, perm, small_scene.rgbas_u8, bg)
        _, _, d_ref = _integrate_bwd(synthetic.NEAR, rs, rn, bg, dss, zs, drgbs, fin, opac, d_fin)
        mbs, fin2, opac2, d2, loss2, nv2 = trainops.integrate_loss_fused(synthetic.NEAR, rs, rn, bg, dss, zs, drgbs, valid, perm,
                                                                         small_scene.rgbas_u8)
        assert int(mbs) == int(eff) and int(nv2) == int(nv) and int(nv) > 0
        assert torch.equal(fin2, fin) and torch.equal(opac2, opac)
        assert torch.equal(d2, d_ref)
        assert abs(float(loss2) - float(loss)) <= 1e-5 * abs(float(loss)) + 1e-9
        assert float(d_ref.abs().max()) > 0


@pytest.mark.parametrize("split", ["1", "0"])
@pytest.mark.parametrize("n", [100, 128 * 148 * 3 + 77, (1 << 18) + 5])
def test_mlp_backward_tcgen05_wgrad_matches_mma_sync_arm(n, split, monkeypatch):
    """The backward kernels against the all-mma.sync arm.  split = 0: eight chain warps, the same per-warp register chain
    (d_enc: same bits; the default); they differ in where the weight gradients are reduced: tcgen05.mma into TMEM accumulators
    that live across the CTA's blocks vs mma.sync into registers.  split = 1 (opt-in): sixteen chain warps in column-split pairs
    that exchange half of every layer through the panels -- the k halves are summed in a different order, so d_enc agrees to
    tf32 rounding, not bit for bit.  Several blocks per CTA exercise the accumulate flag, the panel reuse barriers, the pair
    barriers and both parities."""
    from jaxngp_b200 import nerf as nerf_mod
    monkeypatch.setenv("NGP_B200_MLP_BWD_SPLIT", split)
    gen = torch.Generator(device=DEV).manual_seed(11)
    model = nerf_mod.NeRF(bound=1.0, device=DEV, generator=gen, T=2 ** 14)
    w = model.mlp_flat.detach().clone()
    enc = torch.randn(n, 32, device=DEV, generator=gen) * 0.5
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, device=DEV, generator=gen), dim=-1)
    d_out = torch.randn(n, 4, device=DEV, generator=gen)
    for rep in range(2):  # second call: TMEM re-allocated, accumulators start from zero again
        g_enc_a, g_w_a = nerf_mod.mlp_backward(enc, dirs, w, d_out, impl="umma")
        g_enc_b, g_w_b = nerf_mod.mlp_backward(enc, dirs, w, d_out, impl="mma")
        torch.cuda.synchronize()
        if split == "0":
            assert (g_enc_a != g_enc_b).float().mean() <= 1e-5, rep
        else:
            assert (g_enc_a - g_enc_b).abs().max() <= 5e-3 * g_enc_b.abs().max(), (rep, (g_enc_a - g_enc_b).abs().max(), g_enc_b.abs().max())
            assert (g_enc_a - g_enc_b).abs().mean() <= 1e-4 * g_enc_b.abs().max(), rep
        assert (g_w_a - g_w_b).abs().max() <= 2e-4 * g_w_b.abs().max(), (rep, (g_w_a - g_w_b).abs().max(), g_w_b.abs().max())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_fused_encoder_mlp_forward_is_bit_identical_to_the_two_ops(dtype):
    """ngp_nerf_fused_forward (encoder gather feeding the MLP fragments) against hashgrid_a1_forward +
    nerf_mlp_forward: same arithmetic in the same order, so every output bit must agree -- plain, density-only
    (density-grid update) and grouped (march_rays_inference layout with padding rows) variants."""
    from jaxngp_b200 import encoders as E, nerf as nerf_mod
    gen = torch.Generator(device=DEV).manual_seed(3)
    lt = E.make_level_table(16, 2 ** 19, 2, 16, 2048, 3)
    table = ((torch.rand(lt.rows, 2, device=DEV, generator=gen) - 0.5) * 2).to(dtype)
    w = (torch.rand(nerf_mod.MLP_NUMEL, device=DEV, generator=gen) - 0.5) * 0.6
    n = 50021  # not a multiple of 16: the last tile is partial
    pos = torch.rand(n, 3, device=DEV, generator=gen) * 2 - 1
    pos[:64] = torch.tensor([1.0, -1.0, 0.999999], device=DEV)  # cube faces: Q1 spill rows on the dense levels
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, device=DEV, generator=gen), dim=-1)
    assert nerf_mod.fused_supported(lt, table)
    enc = E.hashgrid_forward(lt, pos, 1.0, table)
    ref = nerf_mod.mlp_forward(enc, dirs, w)
    got, got_enc = nerf_mod.fused_forward(lt, pos, 1.0, table, dirs, w, want_enc=True)
    assert torch.equal(got_enc, enc) and torch.equal(got, ref)
    assert torch.equal(nerf_mod.fused_forward(lt, pos, 1.0, table, dirs, w), ref)
    assert torch.equal(nerf_mod.fused_forward(lt, pos, 1.0, table, None, w), nerf_mod.mlp_forward(enc, None, w))
    # grouped: 3000 rays x 12 slots, random fill counts (0 .. 12)
    n_rays, cap = 3000, 12
    counts = torch.randint(0, cap + 1, (n_rays,), device=DEV, generator=gen, dtype=torch.int32)
    gpos = torch.rand(n_rays * cap, 3, device=DEV, generator=gen) * 2 - 1
    gdirs = torch.nn.functional.normalize(torch.randn(n_rays, 3, device=DEV, generator=gen), dim=-1)
    genc = E.hashgrid_forward(lt, gpos, 1.0, table, group_counts=counts, rows_per_group=cap)
    gref = nerf_mod.mlp_forward(genc, gdirs, w, group_counts=counts, rows_per_group=cap).reshape(n_rays, cap, 4)
    ggot = nerf_mod.fused_forward(lt, gpos, 1.0, table, gdirs, w, group_counts=counts, rows_per_group=cap).reshape(n_rays, cap, 4)
    live = torch.arange(cap, device=DEV)[None, :] < counts[:, None]
    assert torch.equal(ggot[live], gref[live])
    # and the ungrouped evaluation of the same rows
    full = nerf_mod.mlp_forward(E.hashgrid_forward(lt, gpos, 1.0, table), gdirs[:, None, :].expand(-1, cap, -1).reshape(-1, 3).contiguous(), w)
    assert torch.equal(ggot[live], full.reshape(n_rays, cap, 4)[live])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_fused_forward_tcgen05_matches_mma_sync_kernel(dtype):
    """ngp_nerf_fused_forward_umma (dense layers on tcgen05.mma, accumulators in TMEM, one thread per sample) against
    the mma.sync kernel: the encoding it keeps for the backward must be the same bits (same gather), the outputs agree
    to accumulation-order error.  Plain layout with a partial last tile, several tiles per group, and the grouped
    (march_rays_inference) layout with padding rows and fully padded tiles."""
    from jaxngp_b200 import encoders as E, nerf as nerf_mod
    gen = torch.Generator(device=DEV).manual_seed(4)
    lt = E.make_level_table(16, 2 ** 19, 2, 16, 2048, 3)
    table = ((torch.rand(lt.rows, 2, device=DEV, generator=gen) - 0.5) * 2).to(dtype)
    w = (torch.rand(nerf_mod.MLP_NUMEL, device=DEV, generator=gen) - 0.5) * 0.6
    for n in (1, 127, 128 * 148 * 3 * 2 + 77):
        pos = torch.rand(n, 3, device=DEV, generator=gen) * 2 - 1
        dirs = torch.nn.functional.normalize(torch.randn(n, 3, device=DEV, generator=gen), dim=-1)
        ref, ref_enc = nerf_mod.fused_forward(lt, pos, 1.0, table, dirs, w, want_enc=True)
        got, got_enc = nerf_mod.fused_forward(lt, pos, 1.0, table, dirs, w, want_enc=True, impl="umma")
        assert torch.equal(got_enc, ref_enc), n
        assert torch.allclose(got, ref, rtol=2e-5, atol=2e-6), (n, (got - ref).abs().max())
        got2 = nerf_mod.fused_forward(lt, pos, 1.0, table, dirs, w, impl="umma")
        assert torch.equal(got2, got), n
    n_rays, cap = 40000, 16
    counts = torch.randint(0, cap + 1, (n_rays,), device=DEV, generator=gen, dtype=torch.int32)
    counts[1000:3000] = 0  # fully padded tiles
    gpos = torch.rand(n_rays * cap, 3, device=DEV, generator=gen) * 2 - 1
    gdirs = torch.nn.functional.normalize(torch.randn(n_rays, 3, device=DEV, generator=gen), dim=-1)
    live = torch.arange(cap, device=DEV)[None, :] < counts[:, None]
    for rows in (cap, 12):
        c = counts.clamp(max=rows)
        lv = torch.arange(rows, device=DEV)[None, :] < c[:, None]
        gp = gpos[: n_rays * rows]
        gref = nerf_mod.fused_forward(lt, gp, 1.0, table, gdirs, w, group_counts=c, rows_per_group=rows).reshape(n_rays, rows, 4)
        ggot = nerf_mod.fused_forward(lt, gp, 1.0, table, gdirs, w, group_counts=c, rows_per_group=rows, impl="umma").reshape(n_rays, rows, 4)
        assert torch.allclose(ggot[lv], gref[lv], rtol=2e-5, atol=2e-6), rows


def test_graph_renderer_is_bit_identical_to_reference_loop(small_scene):
    """InferenceRenderer (one CUDA graph per loop iteration, grouped encoder/MLP) against
    render_image_inference (the reference's host loop, op for op) on the same rays."""
    from jaxngp_b200 import nerf as nerf_mod, renderers
    gen = torch.Generator(device=DEV).manual_seed(9)
    model = nerf_mod.NeRF(bound=1.0, device=DEV, generator=gen)
    with torch.no_grad():  # make the field non-trivial: visible densities inside the occupied region
        model.position_encoder.latents.uniform_(-1.0, 1.0, generator=gen)
    cam, pose, bits = small_scene.cam, small_scene.transforms[1], small_scene.bitfield_gt
    model.grouped_impl = "mma"  # the bit-identical arm; the tcgen05 arm is compared in test_gpu_fullsize_properties.py
    ref_rgb, ref_depth = renderers.render_image_inference(model, cam, pose, bits, n_rays=1024, march_steps_cap=8, grouped=False)
    o, d = renderers.make_rays_worldspace(cam, pose)
    ts, te = renderers.make_near_far_from_bound(1.0, o, d)
    # (1024, 8) twice: a second renderer of the same shape finds everything warm, so its capture-time warm-up
    # iteration starts at once (regression: the state snapshot used to race with it)
    for n_slots, cap in ((1024, 8), (1024, 8), (4096, 16), (100000, 32)):
        R = renderers.InferenceRenderer(model, cam, bits, n_rays=n_slots, march_steps_cap=cap)
        for _ in range(2):  # second call replays the captured graph on fresh state
            rgb, depth = R.render_rays(o, d, ts, te)
            assert torch.equal(rgb.reshape(ref_rgb.shape), ref_rgb), (n_slots, cap)
            assert torch.allclose(depth.reshape(ref_depth.shape), ref_depth, atol=1e-5)
    # own ray generator: rays differ from the torch restatement by an ulp (test_fused_ray_generation...), which moves
    # DDA skips by a fraction of a step; in this high-frequency random field that shows up as +-1 grey level on some
    # pixels (measured: mean 0.16 levels) and more on the few silhouette pixels
    rgb2, _ = R.render(pose)
    diff = (rgb2.reshape(ref_rgb.shape).int() - ref_rgb.int()).abs().float()
    assert diff.mean() < 0.5 and (diff > 2).float().mean() < 0.02


def test_graph_prefetch_train_step_matches_eager_steps(small_scene):
    """The path bench.py times -- ``train_step(perm, next_perm)``: two march graphs + two compute graphs over
    double-buffered slots, the next batch's march prefetched on a side stream -- against the same K steps run eagerly
    (``use_graph=False``) with SUPPLIED perturbations and backgrounds: the Philox blocks the captured ray-generation kernel
    draws from its device-resident call counter, regenerated here call by call.  A committed density-grid update sits in
    the middle (it drops the prefetch and changes the bitfield the next march reads).

    Integer outputs that depend on the march only (sample count before compaction, valid rays) must be identical at every
    step; the composited sample count and the loss to float tolerance; parameters after K steps to atomic-order
    tolerance (the hash-table scatter and the loss reduction are float atomics)."""
    from jaxngp_b200 import trainops
    from jaxngp_b200.trainer import Trainer
    n_rays, total, K, update_at = 1 << 14, 1 << 17, 10, 5
    gen = torch.Generator(device=DEV).manual_seed(21)
    perms = [torch.randint(0, small_scene.n_pixels, (n_rays,), device=DEV, generator=gen, dtype=torch.int32) for _ in range(K)]
    runs = {}
    for arm in ("graph", "eager"):
        tr = Trainer(device=DEV, n_rays=n_rays, total_samples=total, scene=small_scene, use_graph=arm == "graph", seed=77)
        init = tr.flat_params.clone()
        outs = []
        for k in range(K):
            if arm == "graph":
                out = tr.train_step(perms[k], perms[k + 1] if k + 1 < K else None)
            else:
                u = trainops.philox_uniform(n_rays, k, tr.rng_seed, trainops.STREAM_TRAIN_RAYS, DEV)
                tr.step += 1
                out = tr._step_body(perms[k], u[:, 0].contiguous(), u[:, 1:].contiguous())
            outs.append({name: v.clone() for name, v in out.items()})
            if k + 1 == update_at:
                tr.update_ogrid(update_all=True, commit=True)
        torch.cuda.synchronize()
        runs[arm] = dict(outs=outs, params=tr.flat_params.clone(), init=init, occ=tr.occupancy.clone(),
                         step_dev=int(tr.step_dev), step=tr.step, rng=tr.rng_state.tolist(), m=tr.adam_m.clone())
    g, e = runs["graph"], runs["eager"]
    assert torch.equal(g["init"], e["init"])
    assert g["step_dev"] == e["step_dev"] == K and g["step"] == e["step"] == K  # the graph warm-up leaves no trace
    assert g["rng"] == [K, 0]  # one Philox call per captured march, none for the warm-up
    for k in range(K):
        a, b = g["outs"][k], e["outs"][k]
        if k < update_at:  # same bitfield on both arms: the march is bit-identical
            assert int(a["measured_batch_size_before_compaction"]) == int(b["measured_batch_size_before_compaction"]), k
            assert int(a["n_valid_rays"]) == int(b["n_valid_rays"]), k
        else:  # the updated bitfield thresholds densities that differ in their last bits between the arms
            assert abs(int(a["measured_batch_size_before_compaction"]) - int(b["measured_batch_size_before_compaction"])) \
                <= 2e-3 * int(b["measured_batch_size_before_compaction"]), k
        assert abs(int(a["measured_batch_size"]) - int(b["measured_batch_size"])) <= 2e-3 * int(b["measured_batch_size"]) + 4, k
        assert abs(float(a["loss"]) - float(b["loss"])) <= 2e-3 * abs(float(b["loss"])), (k, float(a["loss"]), float(b["loss"]))
    assert (g["occ"] != e["occ"]).float().mean() < 2e-3
    moved = (e["params"] - e["init"]).abs()
    diff = (g["params"] - e["params"]).abs()
    assert float(moved.max()) > 1e-3
    # Adam normalises the gradient: an entry whose gradient sum is pure rounding noise can take a full-size step of
    # either sign, so a handful of entries may differ by O(lr); everything else agrees closely
    assert float((diff > 1e-4 + 1e-2 * moved).float().mean()) < 2e-3, float((diff > 1e-4 + 1e-2 * moved).float().mean())
    assert float(torch.linalg.norm(diff)) <= 3e-2 * float(torch.linalg.norm(moved))


def test_prefetch_key_sees_in_place_refills(small_scene):
    """A staging buffer refilled in place is a different batch: the prefetched march must not be reused for it."""
    from jaxngp_b200.trainer import Trainer
    n_rays = 1 << 12
    tr = Trainer(device=DEV, n_rays=n_rays, total_samples=1 << 15, scene=small_scene, use_graph=True, seed=5)
    gen = torch.Generator(device=DEV).manual_seed(3)
    a = torch.randint(0, small_scene.n_pixels, (n_rays,), device=DEV, generator=gen, dtype=torch.int32)
    b = torch.randint(0, small_scene.n_pixels, (n_rays,), device=DEV, generator=gen, dtype=torch.int32)
    stage = a.clone()
    tr.train_step(a, stage)  # prefetches `stage` (= a's values)
    stage.copy_(b)           # refilled in place while prefetched
    out = tr.train_step(stage)
    torch.cuda.synchronize()
    assert torch.equal(tr._static_perm[1 - tr._slot], b)  # the step marched the refilled batch, not the stale one
    assert torch.isfinite(out["loss"])


@pytest.mark.parametrize("n", [1, 100, 128 * 148 * 2 + 77, (1 << 18) + 5])
def test_mlp_backward_all_tcgen05_matches_the_other_arms(n):
    """csrc/mlp_bwd_tc.cu (recompute, delta chain and weight gradients all on tcgen05, chain operands in tensor memory,
    weight operands re-staged by bulk copies) against the mma.sync arm and a plain fp32 PyTorch reference of the same
    network: same tf32 operand rounding, so the arms agree to f32 accumulation order except where a pre-activation within
    rounding of zero lands on the other side of a ReLU (counted, not bounded entrywise).  Several tiles per CTA exercise
    the accumulate flags, every panel-reuse barrier in both parities and both weight images; the second call re-allocates
    tensor memory."""
    from jaxngp_b200 import nerf as nerf_mod
    gen = torch.Generator(device=DEV).manual_seed(13)
    model = nerf_mod.NeRF(bound=1.0, device=DEV, generator=gen, T=2 ** 14)
    w = model.mlp_flat.detach().clone()
    enc = torch.randn(n, 32, device=DEV, generator=gen) * 0.5
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, device=DEV, generator=gen), dim=-1)
    d_out = torch.randn(n, 4, device=DEV, generator=gen)
    g_enc_b, g_w_b = nerf_mod.mlp_backward(enc, dirs, w, d_out, impl="mma")
    for rep in range(2):
        g_enc_a, g_w_a = nerf_mod.mlp_backward(enc, dirs, w, d_out, impl="tc")
        torch.cuda.synchronize()
        assert torch.isfinite(g_enc_a).all() and torch.isfinite(g_w_a).all()
        assert torch.linalg.norm(g_enc_a - g_enc_b) <= 2e-3 * torch.linalg.norm(g_enc_b) + 1e-7, (rep, n)
        assert ((g_enc_a - g_enc_b).abs() > 1e-3 * g_enc_b.abs().max()).float().mean() <= 1e-3, (rep, n)
        assert (g_w_a - g_w_b).abs().max() <= 1e-3 * g_w_b.abs().max() + 1e-7, (rep, n, (g_w_a - g_w_b).abs().max(), g_w_b.abs().max())
    # fp32 reference (autograd)
    torch.backends.cuda.matmul.allow_tf32 = False
    wr = w.clone().requires_grad_(True)
    er = enc.clone().requires_grad_(True)
    views, off = {}, 0
    for k, i, o in nerf_mod.MLP_SHAPES:
        views[k] = wr[off:off + i * o].view(i, o)
        off += i * o
    x = torch.relu(er @ views["density_w0"]) @ views["density_w1"]
    h = torch.cat([x, nerf_mod.sh4(dirs)], dim=-1)
    rgb = torch.sigmoid(torch.relu(torch.relu(h @ views["rgb_w0"]) @ views["rgb_w1"]) @ views["rgb_w2"])
    ref = torch.cat([nerf_mod.trunc_exp(x[:, :1]), rgb], dim=-1)
    g_enc_ref, g_w_ref = torch.autograd.grad(ref, [er, wr], d_out)
    torch.backends.cuda.matmul.allow_tf32 = True
    assert torch.linalg.norm(g_enc_a - g_enc_ref) <= 3e-2 * torch.linalg.norm(g_enc_ref) + 1e-6
    assert (g_w_a - g_w_ref).abs().max() <= 1e-2 * g_w_ref.abs().max() + 1e-6


def test_persistent_frame_kernel_matches_the_loop_renderer(small_scene):
    """ngp_render_frame (the slot-refill loop inside one persistent kernel) against the graph-replayed loop on small
    frames: random high-frequency field, frames with fewer rays than one CTA's slots, empty frames, a ray budget that is
    not a multiple of the slot count."""
    from jaxngp_b200 import nerf as nerf_mod, renderers
    gen = torch.Generator(device=DEV).manual_seed(19)
    model = nerf_mod.NeRF(bound=1.0, device=DEV, generator=gen)
    with torch.no_grad():
        model.position_encoder.latents.uniform_(-1.0, 1.0, generator=gen)
    cam, bits = small_scene.cam, small_scene.bitfield_gt
    n_pix = cam["width"] * cam["height"]
    for pixels in (None, torch.arange(0, n_pix, 3, device=DEV, dtype=torch.int32), torch.arange(5, device=DEV, dtype=torch.int32)):
        loop = renderers.InferenceRenderer(model, cam, bits, pixel_indices=pixels, persistent=False)
        one = renderers.InferenceRenderer(model, cam, bits, pixel_indices=pixels)
        assert one.persistent and not loop.persistent
        for view in (1, 2):
            a, da = loop.render(small_scene.transforms[view])
            b, db = one.render(small_scene.transforms[view])
            assert int(one.counters[0]) == one.N and int(one.counters[1]) == int(loop.counters[1])
            assert torch.equal(a, b) and torch.allclose(da, db, atol=1e-6)
    empty = renderers.InferenceRenderer(model, cam, torch.zeros_like(bits))  # nothing occupied: every pixel is background
    rgb, _ = empty.render(small_scene.transforms[0])
    assert int(empty.counters[0]) == empty.N and int(empty.counters[1]) == 0 and bool((rgb == 255).all())


def test_chunked_two_stream_backward_equals_single_pass(small_scene):
    """Trainer._backward in its three arrangements: the default (one kernel, table scatter fused behind the MLP backward), and
    a pipeline of 4 / 3 sample slices on two streams (MLP backward of slice c+1 beside the table scatter of slice c,
    accumulating entry points), against the single pass of the two ops: the same gradient sums, to atomic order."""
    from jaxngp_b200 import nerf as nerf_mod, synthetic
    from jaxngp_b200.trainer import Trainer
    n_rays, total = 1 << 15, 1 << 18
    tr = Trainer(device=DEV, n_rays=n_rays, total_samples=total, scene=small_scene, use_graph=False)
    tr.grid.occupancy.copy_(small_scene.bitfield_gt)
    tr.table.uniform_(-0.5, 0.5, generator=torch.Generator(device=DEV).manual_seed(3))
    gen = torch.Generator(device=DEV).manual_seed(8)
    perm = torch.randint(0, small_scene.n_pixels, (n_rays,), device=DEV, generator=gen, dtype=torch.int32)
    _, _, _, _, _, _, xyzs, dirs, _, _, _ = tr._march_body(perm)
    drgbs, enc = nerf_mod.fused_forward(tr.levels, xyzs, synthetic.BOUND, tr.table, dirs, tr.mlp_flat, want_enc=True)
    d_drgbs = torch.randn(total, 4, device=DEV, generator=gen)
    grads = {}
    assert tr.bwd_fused_scatter and tr.bwd_chunks == 1  # the defaults: one kernel for MLP backward + table scatter
    for chunks in ("fused", 1, 4, 3):
        tr.bwd_fused_scatter = chunks == "fused"
        tr.bwd_chunks = 1 if chunks == "fused" else chunks
        tr.flat_grads.fill_(7.0)  # must be overwritten, not added to
        tr._backward(enc, dirs, xyzs, d_drgbs)
        torch.cuda.synchronize()
        grads[chunks] = tr.flat_grads.clone()
    ref = grads[1]
    assert float(ref[: tr.table_numel].abs().max()) > 0 and float(ref[tr.table_numel:tr.n_params].abs().max()) > 0
    for chunks in ("fused", 4, 3):
        g = grads[chunks]
        assert (g[: tr.table_numel] - ref[: tr.table_numel]).abs().max() <= 1e-4 * ref[: tr.table_numel].abs().max()
        assert (g[tr.table_numel:tr.n_params] - ref[tr.table_numel:tr.n_params]).abs().max() <= 1e-4 * ref[tr.table_numel:tr.n_params].abs().max()


@pytest.mark.parametrize("split", ["1", "0"])
def test_mlp_backward_with_fused_table_scatter_equals_the_two_ops(split, monkeypatch):
    """ngp_nerf_mlp_backward_scatter (the table scatter issued from the MLP backward's own d_enc fragments) against
    nerf_mlp_backward + hashgrid_a1_backward: the same weight gradients (same kernel code) and the same table sums to
    atomic order; cube-face points exercise the spill rows, a partial last block the row guards."""
    from jaxngp_b200 import encoders as E, nerf as nerf_mod
    monkeypatch.setenv("NGP_B200_MLP_BWD_SPLIT", split)  # both ops below take the same kernel family
    gen = torch.Generator(device=DEV).manual_seed(23)
    lt = E.make_level_table(16, 2 ** 19, 2, 16, 2048, 3)
    w = (torch.rand(nerf_mod.MLP_NUMEL, device=DEV, generator=gen) - 0.5) * 0.6
    table = (torch.rand(lt.rows, 2, device=DEV, generator=gen) - 0.5) * 2
    for n in (77, 128 * 148 * 2 + 5):
        pos = torch.rand(n, 3, device=DEV, generator=gen) * 2 - 1
        # the second half as march samples: runs of 64 points 0.0034 apart along random directions, so that consecutive
        # samples share cells at the coarse levels (the merged path of the pair scatter) and split at the fine ones
        k = torch.arange(n - n // 2, device=DEV)
        run = k // 64
        o = (torch.rand(int(run.max()) + 1, 3, device=DEV, generator=gen) - 0.5)
        d = torch.nn.functional.normalize(torch.randn(int(run.max()) + 1, 3, device=DEV, generator=gen), dim=-1)
        pos[n // 2:] = (o[run] + d[run] * (k % 64).to(torch.float32)[:, None] * 0.0034).clamp(-1, 1)
        pos[:32] = torch.tensor([1.0, -1.0, 0.999999], device=DEV)
        dirs = torch.nn.functional.normalize(torch.randn(n, 3, device=DEV, generator=gen), dim=-1)
        enc = E.hashgrid_forward(lt, pos, 1.0, table)
        d_out = torch.randn(n, 4, device=DEV, generator=gen)
        d_out[n // 2: n // 2 + 10] = 0  # masked samples: exact zeros, nothing scattered
        d_enc, d_w_ref = nerf_mod.mlp_backward(enc, dirs, w, d_out)
        d_t_ref = E.hashgrid_backward(lt, pos, 1.0, d_enc)
        d_w = torch.full((nerf_mod.MLP_NUMEL,), 3.0, device=DEV)
        guard = torch.full((lt.rows + 1024, 2), 5.0, device=DEV)
        d_t = guard[: lt.rows]
        nerf_mod.mlp_backward_scatter(lt, pos, 1.0, enc, dirs, w, d_out, d_w, d_t)
        torch.cuda.synchronize()
        assert bool((guard[lt.rows:] == 5.0).all())
        assert (d_w - d_w_ref).abs().max() <= 2e-4 * d_w_ref.abs().max()
        assert (d_t - d_t_ref).abs().max() <= 1e-4 * d_t_ref.abs().max(), n
