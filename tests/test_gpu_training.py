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


def test_train_step_gradients_match_oracle(small_scene):
    from jaxngp_b200 import synthetic as S
    from jaxngp_b200.trainer import Trainer
    from oracle import hashgrid_np as H, train_np as T
    n_rays, total = 4096, 1 << 16
    tr = Trainer(device=DEV, n_rays=n_rays, total_samples=total, scene=small_scene, use_graph=False)
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
    for view, k in zip(tr.mlp_grad_views, ("density_w0", "density_w1", "rgb_w0", "rgb_w1", "rgb_w2")):
        ref = grads[k]
        assert np.abs(view.cpu().numpy() - ref).max() <= 2e-2 * np.abs(ref).max(), k


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
