"""EXACT parity at BASELINE's full sizes (C2: 2^18 rays): march_rays against the C oracle bit for bit (the oracle marches
2^18 rays in seconds on the host cores) and against the reference's own CUDA kernel (oracle/_ref) per ray; integrate_rays
and its backward against the reference kernels on the same GPU and against the C oracle.  Two budgets: the C2 budget of
2^18 samples (the budget rule decides which rays get samples) and one that nothing overflows (every ray of the batch,
~41 M samples)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
N_RAYS = 1 << 18
NAMES = ["next", "exceeded", "valid", "n_samples", "startidx", "idcs", "xyzs", "dirs", "dss", "z_vals"]


def _t(a):
    a = np.ascontiguousarray(a)
    return torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).to(DEV)


def _rays(seed):
    from jaxngp_b200 import synthetic as S
    r = S.training_rays(N_RAYS, seed=seed)
    st = dict(diagonal_n_steps=1024, K=1, G=128, bound=1.0, stepsize_portion=0.0)
    arrays = {k: r[k] for k in ("rays_o", "rays_d", "t_starts", "t_ends", "noises")}
    arrays["occupancy_bitfield"] = S.occupancy_bitfield()
    return st, arrays


@pytest.mark.parametrize("budget", [1 << 18, 1 << 26])
def test_march_rays_c2_bit_exact_vs_oracle(oracle, budget):
    from jaxngp_b200 import volrendjax as V
    st, arrays = _rays(1000000007)
    got = V.march_rays(total_samples=budget, **st, **{k: _t(v) for k, v in arrays.items()}, raw=True)
    exp = oracle.march_rays(total_samples=budget, **st, **arrays, raw=True)
    for name, g, e in zip(NAMES, got, exp):
        e = torch.from_numpy(np.ascontiguousarray(np.asarray(e)).view(np.uint8).reshape(-1))
        g = g.contiguous().view(torch.uint8).reshape(-1).cpu()
        assert g.shape == e.shape and torch.equal(g, e), f"budget 2^{budget.bit_length() - 1}: {name} differs from the oracle"
    used = (int(got[0][0]) & 0xFFFFFFFF) - (int(got[1][0]) & 0xFFFFFFFF)
    assert used > (1 << 17)
    if budget == 1 << 18:
        assert int(got[1][0]) > 0  # the C2 budget overflows: the budget rule was exercised


def _canonical(rn, rs, payload, order_rn, order_rs):
    """Gather a per-sample payload laid out by (rn, rs) into the layout of (order_rn, order_rs) -- ray by ray."""
    has = order_rn > 0
    rays = torch.nonzero(has).reshape(-1)
    counts = order_rn[rays]
    ray_of = torch.repeat_interleave(rays, counts)
    first = torch.cumsum(counts, 0) - counts
    within = torch.arange(int(counts.sum()), device=rn.device) - torch.repeat_interleave(first, counts)
    return payload[rs[ray_of] + within], order_rs[ray_of] + within


@pytest.mark.parametrize("budget", [1 << 18, 1 << 26])
def test_march_rays_c2_vs_reference_kernel(ref, budget):
    """Every ray marched here carries the reference kernel's payload bit for bit.  The reference hands out sample ranges
    in atomic arrival order, so payloads are compared after gathering by each side's own start index, and the reference
    runs with a budget nothing overflows (which rays a full budget serves is arrival-order luck there, ray order here;
    what a served ray receives does not depend on the budget).  Without overflow on our side: every counter and flag too."""
    from jaxngp_b200 import volrendjax as V
    st, arrays = _rays(7)
    a = {k: _t(v) for k, v in arrays.items()}
    got = V.march_rays(total_samples=budget, **st, **a, raw=True)
    exp = ref.march_rays(total_samples=1 << 26, **st, **a, raw=True)
    assert int(exp[1][0]) == 0
    g_rn, g_rs, e_rn, e_rs = (x.long() & 0xFFFFFFFF for x in (got[3], got[4], exp[3], exp[4]))
    if int(got[1][0]) == 0:
        assert int(got[0][0]) == int(exp[0][0])
        assert torch.equal(got[2], exp[2]) and torch.equal(g_rn, e_rn)
    else:
        assert budget == 1 << 18
    served = got[2] & (g_rn > 0)
    assert int(served.sum()) > (1000 if budget == 1 << 18 else 100000), int(served.sum())
    assert bool(exp[2][served].all()) and torch.equal(g_rn[served], e_rn[served])
    order_rn = torch.where(served, g_rn, torch.zeros_like(g_rn))
    for k in (5, 6, 7, 8, 9):  # idcs, xyzs, dirs, dss, z_vals
        ours, _ = _canonical(g_rn, g_rs, got[k], order_rn, g_rs)
        theirs, _ = _canonical(e_rn, e_rs, exp[k], order_rn, g_rs)
        assert torch.equal(ours.contiguous().view(torch.uint8), theirs.contiguous().view(torch.uint8)), NAMES[k]


def test_integrate_rays_c2_vs_reference_kernels_and_oracle(oracle, ref):
    """integrate_rays + backward on a C2 march (2^18 rays, 2^18 samples): count integer-equal to the reference kernel,
    colours / opacities abs 1e-4, gradients rtol 1e-4 (north star's tolerances); the C oracle agrees likewise."""
    from jaxngp_b200 import volrendjax as V
    from jaxngp_b200.volrendjax.integrating import _integrate_bwd, _integrate_fwd
    from tests import inputs
    st, arrays = _rays(11)
    a = {k: _t(v) for k, v in arrays.items()}
    m = V.march_rays(total_samples=1 << 18, **st, **a, raw=True)
    _, _, _, rn, rs, _, xyzs, _, dss, zs = m
    drgbs_np = inputs.drgbs_for(np.zeros((1 << 18, 3), np.float32), 5, 0.3)
    rng = np.random.Generator(np.random.PCG64(6))
    bgs_np = rng.random((N_RAYS, 3), dtype=np.float32)
    dfin_np = rng.normal(size=(N_RAYS, 4)).astype(np.float32)
    drgbs, bgs, dfin = _t(drgbs_np), _t(bgs_np), _t(dfin_np)
    mbs, rgbd, opac = _integrate_fwd(rs, rn, bgs, dss, zs, drgbs)
    rmbs, rrgbd, ropac = ref.integrate_rays(0.3, rs, rn, bgs, dss, zs, drgbs)
    assert int(mbs[0]) == int(rmbs) and int(mbs[0]) > 1000
    assert torch.allclose(rgbd, rrgbd, atol=1e-4, rtol=0) and torch.allclose(opac, ropac, atol=1e-4, rtol=0)
    dbg, dz, dd = _integrate_bwd(0.3, rs, rn, bgs, dss, zs, drgbs, rgbd, opac, dfin)
    rdbg, rdz, rdd = ref.integrate_rays_backward(0.3, rs, rn, bgs, dss, zs, drgbs, rgbd, opac, dfin)
    scale = max(1.0, float(rdd.abs().max()))
    assert torch.allclose(dd, rdd, atol=2e-4 * scale, rtol=1e-4)
    assert torch.allclose(dz, rdz, atol=1e-5, rtol=1e-4) and torch.allclose(dbg, rdbg, atol=1e-6, rtol=1e-5)
    n = lambda x: x.cpu().numpy().view(np.uint32) if x.dtype == torch.int32 else x.cpu().numpy()  # noqa: E731
    ombs, orgbd, oopac = oracle.integrate_rays(0.3, n(rs), n(rn), bgs_np, n(dss), n(zs), drgbs_np)
    assert np.allclose(n(rgbd), orgbd, atol=1e-4, rtol=0) and np.allclose(n(opac), oopac, atol=1e-4, rtol=0)
    assert abs(int(mbs[0]) - ombs) <= 4
    odbg, odz, odd = oracle.integrate_rays_backward(0.3, n(rs), n(rn), bgs_np, n(dss), n(zs), drgbs_np, n(rgbd), n(opac), dfin_np)
    assert np.allclose(n(dd), odd, atol=2e-4 * scale, rtol=1e-3)


def test_hashgrid_c2_table_all_dense_levels_clamp_like_xla():
    """A table whose LAST level is dense (no hashed level behind it): `mod T` lets the outer half-cell's vertex index run
    past the last row; jnp indexing clamps the gather to the last row and its transpose scatters there.  Points on the
    +bound faces exercise it; the oracle's numpy restatement (clamping like XLA) is the checker."""
    from jaxngp_b200 import encoders as E
    from oracle import hashgrid_np as H
    for dim, L, T, n_min, n_max in ((3, 4, 2 ** 19, 16, 64), (2, 6, 2 ** 14, 4, 100)):
        lt = E.make_level_table(L, T, 2, n_min, n_max, dim)
        lv = H.level_table(L, T, 2, n_min, n_max, dim)
        assert not any(lt.hashed)
        rng = np.random.Generator(np.random.PCG64(dim))
        pts = rng.uniform(-1, 1, (4096, dim)).astype(np.float32)
        pts[:512] = np.float32(1.0)                       # the far corner: every vertex coordinate = res on every level
        pts[512:1024, 0] = np.float32(1.0)
        pts[1024:1536, dim - 1] = np.float32(0.99999994)
        tab = rng.uniform(-1, 1, (lt.rows, 2)).astype(np.float32)
        idx, _ = H.indices_and_weights(lv, pts, 1.0)
        assert dim == 2 or int(idx.max()) >= lt.rows  # 3-D case: rows past the table are asked for (scale 63 -> vertex 64 = res)
        enc = E.hashgrid_forward(lt, _t(pts), 1.0, _t(tab)).cpu().numpy()
        assert np.allclose(enc, H.encode(lv, pts, 1.0, tab), atol=2e-6)
        d_enc = rng.normal(size=(4096, 2 * L)).astype(np.float32)
        guard = torch.full((lt.rows + 4096, 2), 7.0, device=DEV)
        out = guard[: lt.rows]
        E.hashgrid_backward(lt, _t(pts), 1.0, _t(d_enc), out=out)
        assert bool((guard[lt.rows:] == 7.0).all())      # nothing written past the table
        ref_g = H.backward(lv, pts, 1.0, d_enc, 2)
        assert np.abs(out.cpu().numpy() - ref_g).max() <= 1e-4 * np.abs(ref_g).max()
