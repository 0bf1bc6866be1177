"""BASELINE config C1 (2-D image fit, models/imagefit.py): the CPU restatement against central differences of its own
forward (CPU), and on the GPU the harness around the 2-D hash-grid kernels
against that restatement for a few optimizer steps."""
import os

import numpy as np
import pytest


def _setup(seed=0, T=2 ** 12, N_max=256, n=96):
    from oracle import hashgrid_np as H
    rng = np.random.default_rng(seed)
    lv = H.level_table(16, T, 2, 16, N_max, 2)
    rows = int(lv["offsets"][-1])
    table = rng.uniform(-1, 1, (rows, 2)).astype(np.float32) * 0.5
    params = {}
    for name, i, o in (("linear1", 32, 128), ("linear2", 128, 128), ("color_predictor", 128, 3)):
        params[name] = (rng.normal(size=(i, o)) / np.sqrt(i), rng.normal(size=o) * 0.1)
    uv = rng.uniform(0, 1, (n, 2)).astype(np.float32)
    target = rng.uniform(0, 1, (n, 3))
    return lv, table, params, uv, target


def test_imagefit_oracle_gradients_match_central_differences():
    from oracle import imagefit_np as I
    lv, table, params, uv, target = _setup()
    loss, g, g_table, d_enc = I.loss_and_grads(lv, table, params, uv, target)
    assert 0 < loss < 1
    rng = np.random.default_rng(1)
    h = 1e-6  # float64 MLP: small enough that no ReLU changes side
    enc = I.encode(lv, table, uv)

    def loss_with(name=None, which=0, idx=None, delta=0.0, enc_idx=None):
        p = {k: (v[0].copy(), v[1].copy()) for k, v in params.items()}
        e = np.asarray(enc, np.float64).copy()
        if name is not None:
            p[name][which][idx] += delta
        if enc_idx is not None:
            e[enc_idx] += delta
        rgb, _ = I.mlp_forward(p, e)
        return float(np.mean((rgb - target) ** 2))

    for name in I.LAYERS:
        for which in (0, 1):
            arr = g[name][which]
            for _ in range(3):
                idx = tuple(rng.integers(0, s) for s in arr.shape)
                fd = (loss_with(name, which, idx, h) - loss_with(name, which, idx, -h)) / (2 * h)
                assert abs(fd - arr[idx]) <= 1e-4 * abs(fd) + 1e-10, (name, which, idx, fd, arr[idx])
    for _ in range(6):  # gradient at the encoder output ...
        idx = (rng.integers(0, enc.shape[0]), rng.integers(0, 32))
        fd = (loss_with(enc_idx=idx, delta=h) - loss_with(enc_idx=idx, delta=-h)) / (2 * h)
        assert abs(fd - d_enc[idx]) <= 1e-4 * abs(fd) + 1e-10, (idx, fd, d_enc[idx])
    # ... carried into the table by the encoder's adjoint (it is linear in the table): <d_enc, enc(table)> = <g_table, table>
    lhs = float((d_enc * np.asarray(enc, np.float64)).sum())
    rhs = float((g_table * table.astype(np.float64)).sum())
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs) + 1e-9 and np.count_nonzero(g_table) > 0
    # optax.adam known answer: first step moves every weight by lr * sign(g) (bias-corrected m / sqrt(v) = g / |g|)
    p, m, v = I.adam_update(np.array([1.0, -2.0]), np.array([0.5, -3.0]), np.zeros(2), np.zeros(2), 1, lr=1e-3)
    assert np.allclose(p, [1.0 - 1e-3, -2.0 + 1e-3], atol=1e-12)


def test_imagefit_host_model_shapes_and_names():
    import torch
    from jaxngp_b200 import imagefit
    gen = torch.Generator().manual_seed(0)
    model = imagefit.ImageFitter(T=2 ** 12, N_max=256, device="cpu", generator=gen)
    assert model.encoder.dim == 2 and model.encoder.levels.rows % 8 == 0
    assert {k: tuple(v.shape) for k, v in model.kernels.items()} == {"linear1": (32, 128), "linear2": (128, 128), "color_predictor": (128, 3)}
    assert all(float(b.detach().abs().max()) == 0 for b in model.biases.values())  # bias_init zeros (imagefit.py:52)
    k = model.kernels["linear2"].detach()
    assert abs(float(k.std()) - (1 / 128) ** 0.5) < 0.1 * (1 / 128) ** 0.5 and float(k.abs().max()) <= 2 * (1 / 128) ** 0.5 / 0.8796 + 1e-6
    uv = imagefit.pixel_uv(torch.tensor([0, 5, 1024 * 3 + 7]), 1024, 1024)
    assert torch.allclose(uv, torch.tensor([[0.0, 0.0], [5 / 1024, 0.0], [7 / 1024, 3 / 1024]]))
    out = model.mlp(torch.zeros(4, 32))
    assert out.shape == (4, 3) and torch.allclose(out, torch.full((4, 3), 0.5))  # zero encoding, zero biases: sigmoid(0)
    with pytest.raises(AssertionError):
        model(torch.zeros(4, 3))


@pytest.mark.gpu
def test_imagefit_harness_matches_oracle_steps():
    import torch
    from jaxngp_b200 import imagefit
    from oracle import imagefit_np as I
    lv, table, params, uv, target = _setup(seed=3, n=512)
    dev = "cuda:0"
    model = imagefit.ImageFitter(T=2 ** 12, N_max=256, device=dev)
    with torch.no_grad():
        model.encoder.latents.copy_(torch.from_numpy(table).to(dev))
        for name in I.LAYERS:
            model.kernels[name].copy_(torch.from_numpy(params[name][0].astype(np.float32)).to(dev))
            model.biases[name].copy_(torch.from_numpy(params[name][1].astype(np.float32)).to(dev))
    opt = imagefit.make_optimizer(model, lr=1e-3)
    torch.backends.cuda.matmul.allow_tf32 = False
    t_uv, t_rgb = torch.from_numpy(uv).to(dev), torch.from_numpy(target.astype(np.float32)).to(dev)
    tab = table.astype(np.float64)
    state = {k: [np.zeros_like(v[0]), np.zeros_like(v[0]), np.zeros_like(v[1]), np.zeros_like(v[1])] for k, v in params.items()}
    mt, vt = np.zeros_like(tab), np.zeros_like(tab)
    for step in range(1, 4):
        loss_o, g, g_table, _ = I.loss_and_grads(lv, tab.astype(np.float32), params, uv, target)
        loss_g = float(imagefit.train_step(model, opt, t_uv, t_rgb))
        assert abs(loss_g - loss_o) <= 1e-4 * loss_o
        tab, mt, vt = I.adam_update(tab, g_table, mt, vt, step)
        for k in I.LAYERS:
            w, mw, vw = I.adam_update(params[k][0], g[k][0], state[k][0], state[k][1], step)
            b, mb, vb = I.adam_update(params[k][1], g[k][1], state[k][2], state[k][3], step)
            params[k], state[k] = (w, b), [mw, vw, mb, vb]
    # Adam's first steps move a weight by ~lr whatever the size of its gradient, so an entry whose gradient cancels to
    # rounding noise may step the other way in float32: allow a handful of those, none elsewhere
    diff = np.abs(model.kernels["linear2"].detach().cpu().numpy() - params["linear2"][0])
    assert np.mean(diff > 2e-4) <= 1e-3 and np.median(diff) <= 1e-5, (float(diff.max()), float(np.mean(diff > 2e-4)))
