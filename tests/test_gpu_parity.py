"""Parity tests proper: the CUDA path (through the C ABI of libngp_b200.so) against
  (1) the CPU oracle (oracle/ngp_oracle.c, ray-order semantics) -- bit-exact for integer work, and
  (2) the reference's own CUDA kernels (oracle/_ref) on the same B200, on the same seeded inputs.
"""
import numpy as np
import pytest
import torch

from tests import inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def t(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a).to(DEV)


def n(x):
    x = x.detach().cpu().numpy()
    return x.view(np.uint32) if x.dtype == np.int32 else x


# ------------------------------------------------------------------ morton / packbits
def test_morton_kat_and_roundtrip(oracle):
    from jaxngp_b200 import volrendjax as V
    xyz = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1023, 1023, 1023], [5, 9, 1000]], np.uint32)
    got = n(V.morton3d(t(xyz)))
    assert got[:4].tolist() == [1, 2, 4, 0x3FFFFFFF]  # marching.cu:52-68
    rng = np.random.Generator(np.random.PCG64(0))
    xyz = rng.integers(0, 1024, (100003, 3), dtype=np.uint32)
    m = V.morton3d(t(xyz))
    assert np.array_equal(n(m), oracle.morton3d(xyz))
    assert np.array_equal(n(V.morton3d_invert(m)), xyz)
    idx = rng.integers(0, 2 ** 30, 70001, dtype=np.uint32)
    assert np.array_equal(n(V.morton3d_invert(t(idx))), oracle.morton3d_invert(idx))


def test_morton_vs_reference(ref):
    from jaxngp_b200 import volrendjax as V
    rng = np.random.Generator(np.random.PCG64(1))
    xyz = t(rng.integers(0, 1024, (50000, 3), dtype=np.uint32))
    assert torch.equal(V.morton3d(xyz), ref.morton3d(xyz))
    idx = t(rng.integers(0, 2 ** 30, 50000, dtype=np.uint32))
    assert torch.equal(V.morton3d_invert(idx), ref.morton3d_invert(idx))


@pytest.mark.parametrize("nbits", [8, 24, 1000 * 8, 128 ** 3])
def test_packbits(oracle, nbits):
    from jaxngp_b200 import volrendjax as V
    rng = np.random.Generator(np.random.PCG64(nbits))
    den = rng.normal(size=nbits).astype(np.float32)
    den[::7] = -1.0
    for thr in (0.25, rng.normal(size=nbits).astype(np.float32)):
        mask, bits = V.packbits(t(np.asarray(thr, np.float32)) if isinstance(thr, np.ndarray) else thr, t(den))
        omask, obits = oracle.packbits(thr, den)
        assert np.array_equal(n(mask), omask) and np.array_equal(n(bits), obits)
    # LSB-first known answer (packbits.cu:28-33)
    _, b = V.packbits(0.5, t(np.array([1, 0, 0, 0, 0, 0, 0, 1], np.float32)))
    assert n(b).tolist() == [0x81]
    with pytest.raises(ValueError):
        V.packbits(0.5, torch.zeros(12, device=DEV))
    with pytest.raises(NotImplementedError):
        V.packbits(0.5, torch.zeros(16, device=DEV, dtype=torch.float64))


def test_packbits_vs_reference(ref):
    from jaxngp_b200 import volrendjax as V
    rng = np.random.Generator(np.random.PCG64(5))
    den = t(rng.normal(size=128 ** 3).astype(np.float32))
    m1, b1 = V.packbits(0.1, den)
    m2, b2 = ref.packbits(0.1, den)
    assert torch.equal(m1, m2) and torch.equal(b1, b2)


# ------------------------------------------------------------------ march_rays
def run_march(mod, st, arrays, conv):
    return mod.march_rays(**st, **{k: conv(v) for k, v in arrays.items()}, raw=True)


@pytest.mark.parametrize("case", inputs.MARCH_CASES)
def test_march_rays_vs_oracle_bit_exact(oracle, case):
    from jaxngp_b200 import volrendjax as V
    st, arrays = inputs.march_case(case)
    got = [n(x) for x in run_march(V, st, arrays, t)]
    exp = run_march(oracle, st, arrays, lambda a: a)
    names = ["next", "exceeded", "valid", "n_samples", "startidx", "idcs", "xyzs", "dirs", "dss", "z_vals"]
    for name, g, e in zip(names, got, exp):
        e = np.asarray(e)
        assert g.shape == e.shape, name
        assert np.array_equal(g.view(np.uint8), e.view(np.uint8)), f"{case}: {name} differs from the oracle"
    if case == "overflow":
        assert int(got[1][0]) > 0  # the overflow path was exercised
    if case == "dense":
        assert got[3].max() == 1024  # per-ray cap diagonal_n_steps * bound (marching.cu:165)


@pytest.mark.parametrize("case", inputs.MARCH_CASES)
def test_march_rays_vs_reference(ref, case):
    """Parity with the reference CUDA kernel, as defined in SURVEY 7 hard part 2: the reference hands
    out sample ranges in atomic arrival order, so compare per ray after gathering by each side's own
    start index; counters and validity are compared when nothing overflows."""
    from jaxngp_b200 import volrendjax as V
    st, arrays = inputs.march_case(case)
    got = [n(x) for x in run_march(V, st, arrays, t)]
    exp = [n(x) for x in run_march(ref, st, arrays, t)]
    S = st["total_samples"]
    g_next, g_exc, g_valid, g_n, g_start = got[:5]
    e_next, e_exc, e_valid, e_n, e_start = exp[:5]
    overflow = int(e_exc[0]) > 0 or int(g_exc[0]) > 0 or int(e_next[0]) >= S
    if not overflow:
        assert int(g_next[0]) == int(e_next[0]) and int(g_exc[0]) == int(e_exc[0]) == 0
        assert np.array_equal(g_valid, e_valid)
        assert np.array_equal(g_n, e_n)
    both = g_valid & e_valid
    assert np.array_equal(g_n[both], e_n[both])
    # ranges disjoint and inside [0, S)
    for nn, ss in ((g_n, g_start), (e_n, e_start)):
        order = np.argsort(ss[nn > 0], kind="stable")
        s_sorted, n_sorted = ss[nn > 0][order].astype(np.int64), nn[nn > 0][order].astype(np.int64)
        assert np.all(s_sorted[1:] >= s_sorted[:-1] + n_sorted[:-1]) and (len(s_sorted) == 0 or s_sorted[-1] + n_sorted[-1] <= S)
    rays = np.nonzero(both & (g_n > 0))[0]
    for k in (5, 6, 7, 8, 9):  # idcs, xyzs, dirs, dss, z_vals: bit-equal payloads
        for r in rays[:: max(1, len(rays) // 400)]:
            a = got[k][g_start[r]: g_start[r] + g_n[r]]
            b = exp[k][e_start[r]: e_start[r] + e_n[r]]
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), (case, k, r)
    if not overflow:  # whole-array check through a canonical (ray-ordered) layout
        for k in (6, 8, 9):
            ga = np.concatenate([got[k][g_start[r]: g_start[r] + g_n[r]] for r in rays]) if len(rays) else np.zeros(0)
            ea = np.concatenate([exp[k][e_start[r]: e_start[r] + e_n[r]] for r in rays]) if len(rays) else np.zeros(0)
            assert np.array_equal(ga, ea)


def test_march_rays_contract_errors():
    from jaxngp_b200 import volrendjax as V
    st, arrays = inputs.march_case("miss")
    a = {k: t(v) for k, v in arrays.items()}
    with pytest.raises(AssertionError):
        V.march_rays(**st, **{**a, "occupancy_bitfield": a["occupancy_bitfield"][:-1]})
    with pytest.raises(NotImplementedError):
        V.march_rays(**st, **{**a, "rays_o": a["rays_o"].double()})
    with pytest.raises((AssertionError, RuntimeError)):  # chex.assert_scalar_positive(K), then ffi.cc:79-81
        V.march_rays(**{**st, "K": 0, "G": 0}, **{**a, "occupancy_bitfield": a["occupancy_bitfield"][:0]})


# ------------------------------------------------------------------ integrate_rays fwd / bwd
def _integrate_inputs(oracle, case, seed, scale=1.0):
    st, arrays = inputs.march_case(case)
    m = run_march(oracle, st, arrays, lambda a: a)
    _, _, valid, rn, rs, idcs, xyzs, dirs, dss, zs = m
    drgbs = inputs.drgbs_for(xyzs, seed, scale)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    bgs = rng.random((rn.shape[0], 3), dtype=np.float32)
    return rs, rn, bgs, dss, zs, drgbs


@pytest.mark.parametrize("case,scale", [("scene", 1.0), ("scene", 0.02), ("cascades", 1.0), ("dense", 0.05), ("miss", 1.0)])
def test_integrate_rays_forward(oracle, ref, case, scale):
    from jaxngp_b200 import volrendjax as V
    rs, rn, bgs, dss, zs, drgbs = _integrate_inputs(oracle, case, 21, scale)
    mbs, rgbd, opac = V.integrate_rays(0.3, t(rs), t(rn), t(bgs), t(dss), t(zs), t(drgbs))
    rmbs, rrgbd, ropac = ref.integrate_rays(0.3, t(rs), t(rn), t(bgs), t(dss), t(zs), t(drgbs))
    # vs the reference kernel on the same GPU: count integer-exact, colours abs 1e-4
    assert int(mbs) == int(rmbs)
    assert torch.allclose(rgbd, rrgbd, atol=1e-4, rtol=0) and torch.allclose(opac, ropac, atol=1e-4, rtol=0)
    # vs the CPU oracle (expf instead of ex2.approx): abs 1e-4; count within the few samples whose
    # early-stop decision sits within an ulp of the threshold
    ombs, orgbd, oopac = oracle.integrate_rays(0.3, rs, rn, bgs, dss, zs, drgbs)
    assert np.allclose(n(rgbd), orgbd, atol=1e-4, rtol=0) and np.allclose(n(opac), oopac, atol=1e-4, rtol=0)
    assert abs(int(mbs) - ombs) <= max(4, ombs // 100000)


@pytest.mark.parametrize("case,scale", [("scene", 1.0), ("scene", 0.02), ("cascades", 1.0)])
def test_integrate_rays_backward(oracle, ref, case, scale):
    from jaxngp_b200 import volrendjax as V
    rs, rn, bgs, dss, zs, drgbs = _integrate_inputs(oracle, case, 33, scale)
    rng = np.random.Generator(np.random.PCG64(99))
    dfin = rng.normal(size=(rn.shape[0], 4)).astype(np.float32)
    drgbs_t = t(drgbs).requires_grad_(True)
    zs_t = t(zs).requires_grad_(True)
    bgs_t = t(bgs).requires_grad_(True)
    mbs, rgbd, opac = V.integrate_rays(0.3, t(rs), t(rn), bgs_t, t(dss), zs_t, drgbs_t)
    rgbd.backward(t(dfin))
    rdbg, rdz, rdd = ref.integrate_rays_backward(0.3, t(rs), t(rn), t(bgs), t(dss), t(zs), t(drgbs), rgbd.detach(),
                                                 opac.detach(), t(dfin))
    scale_d = max(1.0, float(rdd.abs().max()))
    assert torch.allclose(drgbs_t.grad, rdd, atol=2e-4 * scale_d, rtol=1e-4)
    assert torch.allclose(zs_t.grad, rdz, atol=1e-5, rtol=1e-4)
    assert torch.allclose(bgs_t.grad, rdbg, atol=1e-6, rtol=1e-5)
    odbg, odz, odd = oracle.integrate_rays_backward(0.3, rs, rn, bgs, dss, zs, drgbs, n(rgbd.detach()), n(opac.detach()), dfin)
    assert np.allclose(n(drgbs_t.grad), odd, atol=2e-4 * scale_d, rtol=1e-3)
    assert np.allclose(n(zs_t.grad), odz, atol=1e-5, rtol=1e-3)
    assert np.allclose(n(bgs_t.grad), odbg, atol=1e-5, rtol=1e-4)


# ------------------------------------------------------------------ inference loop (march + integrate)
def test_inference_loop_matches_oracle_and_reference(oracle, ref):
    """Drives models/renderers/cuda.py:318-361's slot-refill loop with an analytic radiance field on
    all three implementations and compares every intermediate."""
    from jaxngp_b200 import synthetic as S
    from jaxngp_b200 import volrendjax as V
    st, fr, bits, n_slots = inputs.inference_case()
    N = fr["rays_o"].shape[0]
    cap = st["march_steps_cap"]

    def field(xyzs):
        xyz = xyzs.reshape(-1, 3)
        return np.concatenate([S.density(xyz)[:, None] * 0.5, S.colour(xyz)], -1).reshape(*xyzs.shape[:-1], 4).astype(np.float32)

    def loop(mod, conv, back):
        o, d, ts, te, b = tuple(conv(fr[k]) for k in ("rays_o", "rays_d", "t_starts", "t_ends")) + (conv(bits),)
        bg = conv(np.ones((N, 3), np.float32))
        rgbd, T = conv(np.zeros((N, 4), np.float32)), conv(np.ones(N, np.float32))
        term, idx, nri = conv(np.ones(n_slots, np.bool_)), conv(np.zeros(n_slots, np.uint32)), conv(np.zeros(1, np.uint32))
        trace, rendered, it = [], 0, 0
        while rendered < N and it < 400:
            out = mod.march_rays_inference(**st, rays_o=o, rays_d=d, t_starts=ts, t_ends=te, occupancy_bitfield=b,
                                           next_ray_index_in=nri, terminated=term, indices=idx)
            nri, idx, ns, ts, xyzs, dss, zs = out[:7]
            drgbs = conv(field(back(xyzs)))
            cnt, term, rgbd, T = mod.integrate_rays_inference(bg, rgbd, T, ns, idx, dss, zs, drgbs)
            rendered += int(cnt)
            trace.append((back(idx).copy(), back(ns).copy(), back(dss).copy(), back(zs).copy(), back(xyzs).copy(), int(cnt)))
            it += 1
        return back(rgbd), back(T), trace, it

    g_rgbd, g_T, g_trace, g_it = loop(V, t, n)
    o_rgbd, o_T, o_trace, o_it = loop(oracle, lambda a: a, lambda a: np.asarray(a))
    assert g_it == o_it and g_it < 400
    for (gi, gn, gd, gz, gx, gc), (oi, on, od, oz, ox, oc) in zip(g_trace, o_trace):
        assert np.array_equal(gi, oi) and np.array_equal(gn, on) and gc == oc  # slot-order admission: exact
        assert np.array_equal(gd, od) and np.array_equal(gz, oz) and np.array_equal(gx, ox)
    assert np.allclose(g_rgbd, o_rgbd, atol=1e-4) and np.allclose(g_T, o_T, atol=1e-4)
    # the reference assigns fresh rays to slots in atomic arrival order; the finished image must agree
    r_rgbd, r_T, r_trace, r_it = loop(ref, t, n)
    assert np.allclose(g_rgbd, r_rgbd, atol=1e-4) and np.allclose(g_T, r_T, atol=1e-4)
    # set of admitted rays per iteration is identical
    for (gi, gn, *_), (ri, rn_, *_) in zip(g_trace, r_trace):
        assert np.array_equal(np.sort(gi), np.sort(ri))
        assert np.array_equal(gn[np.argsort(gi, kind="stable")], rn_[np.argsort(ri, kind="stable")])


@pytest.mark.parametrize("case,cap,n_slots", [("scene", 8, 5000), ("scene", 16, 3000), ("cascades", 5, 2500),
                                              ("cascades", 40, 1500), ("dense", 1, 1100), ("dense", 33, 700), ("miss", 8, 300)])
def test_march_rays_inference_slot_states_bit_exact(oracle, ref, case, cap, n_slots):
    """march_rays_inference alone, three consecutive calls over arbitrary slot states (some slots terminated,
    some resuming mid-ray, slot count spanning several 1024-slot rank blocks, more slots than rays left):
    every output bit-equal to the oracle; per-ray payloads bit-equal to the reference kernel.  'dense'
    (all-ones grid) exercises the far-plane rule (marching.cu:367-394), 'cascades' K=3 with exponential steps."""
    from jaxngp_b200 import volrendjax as V
    st, arr = inputs.march_case(case)
    st = dict(diagonal_n_steps=st["diagonal_n_steps"], K=st["K"], G=st["G"], march_steps_cap=cap, bound=st["bound"],
              stepsize_portion=st["stepsize_portion"])
    N = arr["rays_o"].shape[0]
    rng = np.random.Generator(np.random.PCG64(cap * 1000 + n_slots))
    term = np.ones(n_slots, np.bool_)
    idx = np.zeros(n_slots, np.uint32)
    nri = np.zeros(1, np.uint32)
    ts_o = arr["t_starts"].copy()
    ts_g = t(ts_o)
    fixed = {k: arr[k] for k in ("rays_o", "rays_d", "t_ends")}
    fixed_g = {k: t(v) for k, v in fixed.items()}
    bits_g = t(arr["occupancy_bitfield"])
    idx_g, nri_g = t(idx), t(nri)
    for call in range(3):
        o = oracle.march_rays_inference(**st, **fixed, t_starts=ts_o, occupancy_bitfield=arr["occupancy_bitfield"],
                                        next_ray_index_in=nri, terminated=term, indices=idx)
        g = V.march_rays_inference(**st, **fixed_g, t_starts=ts_g, occupancy_bitfield=bits_g, next_ray_index_in=nri_g,
                                   terminated=t(term), indices=idx_g)
        names = ("next_ray_index", "indices", "n_samples", "t_starts", "xyzs", "dss", "z_vals")
        for name, a, b in zip(names, g[:7], o[:7]):
            a, b = n(a), np.asarray(b)
            assert np.array_equal(a.view(np.uint32).ravel(), b.view(np.uint32).ravel()), (call, name)
        if call == 0:  # reference kernel: arrival-order admission, so compare per admitted ray
            r = ref.march_rays_inference(**st, **fixed_g, t_starts=t(arr["t_starts"]), occupancy_bitfield=bits_g,
                                         next_ray_index_in=t(np.zeros(1, np.uint32)), terminated=t(term), indices=t(idx))
            gi, ri = n(g[1]).astype(np.int64), n(r[1]).astype(np.int64)
            assert np.array_equal(np.sort(gi), np.sort(ri))
            go, ro = np.argsort(gi, kind="stable"), np.argsort(ri, kind="stable")
            for k in (2, 4, 5, 6):
                a, b = n(g[k])[go], n(r[k])[ro]
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), k
            live = np.sort(gi)[np.sort(gi) < N]
            assert np.array_equal(n(g[3]).view(np.uint32)[live], n(r[3]).view(np.uint32)[live])
        nri, idx, ns, ts_o = o[0], o[1], o[2], o[3]
        nri_g, idx_g, ts_g = g[0], g[1], g[3]
        # next state: slots that ran out of samples terminate, plus a random third (early stop by transmittance)
        term = (ns < cap) | (rng.random(n_slots) < 0.33)


@pytest.mark.parametrize("case", ["scene", "cascades", "dense", "miss"])
def test_march_rays_skip_empty_preserves_inference_march(case):
    """ngp_march_rays_skip_empty (renderer fast path) moves t_starts to the first occupied visited point; the
    samples march_rays_inference then emits, and the t it hands back, must not change by a single bit."""
    from jaxngp_b200 import _lib, descriptors, volrendjax as V
    st, arr = inputs.march_case(case)
    st = dict(diagonal_n_steps=st["diagonal_n_steps"], K=st["K"], G=st["G"], march_steps_cap=12, bound=st["bound"],
              stepsize_portion=st["stepsize_portion"])
    N = arr["rays_o"].shape[0]
    o, d, ts, te, bits = (t(arr[k]) for k in ("rays_o", "rays_d", "t_starts", "t_ends", "occupancy_bitfield"))
    ts2 = torch.empty_like(ts)
    desc = descriptors.make_marching_inference_descriptor(N, N, st["diagonal_n_steps"], st["K"], st["G"], 12, st["bound"],
                                                          st["stepsize_portion"])
    _lib.call("ngp_march_rays_skip_empty", [o, d, ts, te, bits, ts2], desc)
    assert (ts2 >= ts).all()
    if case != "dense":
        assert (ts2 > ts).any() or case == "miss"
    term, idx, nri = t(np.ones(N, np.bool_)), t(np.zeros(N, np.uint32)), t(np.zeros(1, np.uint32))
    a = V.march_rays_inference(**st, rays_o=o, rays_d=d, t_starts=ts, t_ends=te, occupancy_bitfield=bits,
                               next_ray_index_in=nri, terminated=term, indices=idx)
    b = V.march_rays_inference(**st, rays_o=o, rays_d=d, t_starts=ts2, t_ends=te, occupancy_bitfield=bits,
                               next_ray_index_in=nri, terminated=term, indices=idx)
    for k, (x, y) in enumerate(zip(a[:7], b[:7])):
        assert torch.equal(x.view(torch.int32) if x.dtype == torch.float32 else x,
                           y.view(torch.int32) if y.dtype == torch.float32 else y), k


# ------------------------------------------------------------------ hash-grid encoder
@pytest.mark.parametrize("dim,T,N_max", [(3, 2 ** 19, 2048), (3, 2 ** 14, 512), (2, 2 ** 19, 2 ** 19), (2, 2 ** 12, 4096)])
def test_hashgrid_forward_backward_vs_oracle(oracle, dim, T, N_max):
    from jaxngp_b200 import encoders as E
    from oracle import hashgrid_np as H
    lt = E.make_level_table(16, T, 2, 16, N_max, dim)
    lv = H.level_table(16, T, 2, 16, N_max, dim)
    assert list(lt.offsets) == lv["offsets"].tolist() and list(lt.res) == lv["res"].tolist()
    assert np.array_equal(np.asarray(lt.scales, np.float32), lv["scales"])
    pts = inputs.encoder_points(20011, dim)
    pts[:7] = np.array([[-1.0] * dim, [1.0] * dim, [0.0] * dim, [0.999999] * dim, [0.96] * dim, [-0.5] * dim, [0.933334] * dim], np.float32)
    table = inputs.encoder_table(lt.rows, 2, amp=1.0)
    enc = E.hashgrid_forward(lt, t(pts), 1.0, t(table))
    ref_enc = oracle.hashgrid_encode(lv, pts, 1.0, table)
    assert np.allclose(n(enc), ref_enc, rtol=1e-3, atol=1e-6)  # north star: features rel 1e-3
    assert np.abs(n(enc) - ref_enc).max() < 2e-6                # in fact a few ulp (sum order only)
    rng = np.random.Generator(np.random.PCG64(44))
    d_enc = rng.normal(size=enc.shape).astype(np.float32)
    g = E.hashgrid_backward(lt, t(pts), 1.0, t(d_enc))
    ref_g = oracle.hashgrid_backward(lv, pts, 1.0, d_enc, 2)
    denom = np.abs(ref_g).max()
    assert np.abs(n(g) - ref_g).max() <= 1e-2 * denom          # north star: table grads rel 1e-2
    assert np.abs(n(g) - ref_g).max() <= 1e-4 * denom          # measured: atomic reordering only
    assert np.array_equal(n(g) != 0, ref_g != 0)                # exactly the same rows are touched


@pytest.mark.parametrize("dim,T,N_max", [(3, 2 ** 19, 2048), (2, 2 ** 19, 2 ** 19), (3, 2 ** 14, 512)])
def test_hashgrid_vs_reference_code_golden(dim, T, N_max):
    """The CUDA encoder against golden vectors from the reference's OWN HashGridEncoder.__call__ (models/encoders.py,
    executed unmodified by oracle/make_golden_encoder.py; tests/golden/encoder_reference.npz): encoded features within
    the north star's rel 1e-3 (in fact a few ulp), and the backward through the adjoint identity
    <d_enc, enc_ref(table)> = <d_table, table> -- the encoder is linear in the table, so the table gradient of the
    reference's autodiff is J^T d_enc for the same J whose action the golden vectors record."""
    import os
    from jaxngp_b200 import encoders as E
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encoder_reference.npz"))
    key = f"d{dim}_T{T}_N{N_max}"
    lt = E.make_level_table(16, T, 2, 16, N_max, dim)
    assert lt.rows == int(g[key + "_rows"]) and abs(lt.b - float(g[key + "_b"])) < 1e-12
    pts, ref_enc = g[key + "_pts"], g[key + "_enc"]
    table = inputs.encoder_table(lt.rows, 2, amp=1.0)
    enc = n(E.hashgrid_forward(lt, t(pts), 1.0, t(table)))
    assert np.allclose(enc, ref_enc, rtol=1e-3, atol=1e-6)
    assert np.abs(enc - ref_enc).max() < 2e-6
    rng = np.random.Generator(np.random.PCG64(45))
    d_enc = rng.normal(size=ref_enc.shape).astype(np.float32)
    grad = n(E.hashgrid_backward(lt, t(pts), 1.0, t(d_enc)))
    lhs = float((d_enc.astype(np.float64) * ref_enc.astype(np.float64)).sum())
    rhs = float((grad.astype(np.float64) * table.astype(np.float64)).sum())
    assert abs(lhs - rhs) <= 1e-4 * (abs(lhs) + np.abs(d_enc).sum() * 1e-3)


def test_hashgrid_backward_ray_ordered_runs(oracle):
    """Samples in ray order (neighbouring rows share grid cells on the coarse levels) with zero-gradient
    rows sprinkled in: exercises the run aggregation of the scatter kernel."""
    from jaxngp_b200 import encoders as E
    from oracle import hashgrid_np as H
    lt = E.make_level_table(16, 2 ** 19, 2, 16, 2048, 3)
    lv = H.level_table(16, 2 ** 19, 2, 16, 2048, 3)
    rng = np.random.Generator(np.random.PCG64(77))
    n_rays, per_ray = 300, 37
    o = rng.uniform(-0.9, 0.9, (n_rays, 1, 3))
    d = rng.normal(size=(n_rays, 1, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    tt = (np.arange(per_ray) * (2 * np.sqrt(3) / 1024))[None, :, None]
    pts = np.clip(o + tt * d, -1, 1).reshape(-1, 3).astype(np.float32)
    d_enc = rng.normal(size=(pts.shape[0], 32)).astype(np.float32)
    d_enc[rng.random(pts.shape[0]) < 0.15] = 0.0          # masked samples
    d_enc[5000:5100, ::2] = 0.0                            # partially zero rows stay active
    g = E.hashgrid_backward(lt, t(pts), 1.0, t(d_enc))
    ref_g = oracle.hashgrid_backward(lv, pts, 1.0, d_enc, 2)
    assert np.abs(n(g) - ref_g).max() <= 1e-4 * np.abs(ref_g).max()
    assert np.array_equal(n(g) != 0, ref_g != 0)


def test_hashgrid_module_autograd_and_fp16(oracle):
    from jaxngp_b200 import encoders as E
    from oracle import hashgrid_np as H
    gen = torch.Generator(device=DEV).manual_seed(0)
    enc_mod = E.HashGridEncoder(L=16, T=2 ** 19, F=2, N_min=16, N_max=2048, tv_scale=0.0, device=DEV, generator=gen)
    assert tuple(enc_mod.latents.shape) == (6098120, 2)  # rows at C2
    assert float(enc_mod.latents.detach().abs().max()) <= 1e-4
    pts = inputs.encoder_points(4099)
    out, tv = enc_mod(t(pts), 1.0)
    assert tv == 0 and tuple(out.shape) == (4099, 32)
    lv = H.level_table(16, 2 ** 19, 2, 16, 2048, 3)
    ref_out = oracle.hashgrid_encode(lv, pts, 1.0, n(enc_mod.latents))
    assert np.allclose(n(out), ref_out, rtol=1e-3, atol=1e-9)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    ref_g = oracle.hashgrid_backward(lv, pts, 1.0, n(w), 2)
    assert np.abs(n(enc_mod.latents.grad) - ref_g).max() <= 1e-4 * np.abs(ref_g).max()
    # fp16 storage variant: features within rel 1e-3 of the fp32-table oracle
    half = E.hashgrid_forward(enc_mod.levels, t(pts), 1.0, enc_mod.latents.detach().half())
    assert np.allclose(n(half), ref_out, rtol=2e-3, atol=1e-7)


def test_hashgrid_tcnn_path(oracle):
    """jaxtcnn.hashgrid_encode (tiny-cuda-nn indexing, SoA layout).  Parity unpinned by the reference;
    pinned here to the a1 oracle with `% level size` wrapping and un-aligned rows, tolerance rel 1e-3
    (the level scale is f32-on-device instead of f64-on-host)."""
    from jaxngp_b200 import encoders as E
    from jaxngp_b200 import jaxtcnn as J
    from oracle import hashgrid_np as H
    lt = E.make_level_table(16, 2 ** 19, 2, 16, 2048, 3, align=1)
    assert lt.rows == 6098108  # encoders.py:275
    lv = H.level_table(16, 2 ** 19, 2, 16, 2048, 3, align=1)
    pts = inputs.encoder_points(8191)
    table = inputs.encoder_table(lt.rows, 2, amp=1.0)
    desc = J.HashGridMetadata(L=16, F=2, N_min=16, per_level_scale=lt.b)
    params = t(table).requires_grad_(True)
    coords = t(((pts + 1) / 2).T.copy()).requires_grad_(True)
    out = J.hashgrid_encode(desc, t(np.asarray(lt.offsets, np.uint32)), coords, params)
    assert tuple(out.shape) == (32, 8191)
    # tiny-cuda-nn derives the level scale in f32 on the device: exp2f(l * log2f(b)) * N_min - 1
    # (SURVEY Q3); feed the oracle the same f32 recipe.  A 1-ulp difference in a scale of ~2000 moves a
    # point by 1e-4 cells, i.e. ~1e-3 of the O(1) random table used here, hence the two-level check.
    l = np.arange(16, dtype=np.float32)
    lv = dict(lv, scales=(np.exp2(l * np.log2(np.float32(lt.b))) * np.float32(16) - np.float32(1)).astype(np.float32))
    ref_out = oracle.hashgrid_encode(lv, pts, 1.0, table, wrap="tcnn")
    err = np.abs(n(out).T - ref_out)
    assert np.quantile(err, 0.99) < 1e-3 and err.max() < 2e-2, (np.quantile(err, [0.5, 0.99, 0.999]), err.max())
    assert np.allclose(n(out).T[:, :20], ref_out[:, :20], rtol=1e-3, atol=1e-4)  # levels 0-9: scales < 300
    w = torch.randn_like(out)
    (out * w).sum().backward()
    ref_g = oracle.hashgrid_backward(lv, pts, 1.0, n(w).T.copy(), 2, wrap="tcnn")
    assert np.abs(n(params.grad) - ref_g).max() <= 2e-2 * np.abs(ref_g).max()
    # d/dcoords against central differences of the oracle, on a coarse all-dense table where a step
    # of 1e-3 rarely crosses a cell face (multilinear interpolation is exactly linear inside a cell)
    lt2 = E.make_level_table(4, 2 ** 19, 2, 4, 32, 3, align=1)
    lv2 = H.level_table(4, 2 ** 19, 2, 4, 32, 3, align=1)
    table2 = inputs.encoder_table(lt2.rows, 2, seed=5, amp=1.0)
    desc2 = J.HashGridMetadata(L=4, F=2, N_min=4, per_level_scale=lt2.b)
    pts2 = inputs.encoder_points(2048, seed=6) * 0.95
    coords2 = t(((pts2 + 1) / 2).T.copy()).requires_grad_(True)
    out2 = J.hashgrid_encode(desc2, t(np.asarray(lt2.offsets, np.uint32)), coords2, t(table2))
    w2 = torch.randn_like(out2)
    (out2 * w2).sum().backward()
    eps = 1e-3
    gnum = np.zeros_like(pts2)
    for k in range(3):
        dp = np.zeros(3, np.float32)
        dp[k] = eps
        hi = oracle.hashgrid_encode(lv2, pts2 + dp, 1.0, table2, wrap="tcnn").astype(np.float64)
        lo = oracle.hashgrid_encode(lv2, pts2 - dp, 1.0, table2, wrap="tcnn").astype(np.float64)
        gnum[:, k] = ((hi - lo) * n(w2).T).sum(-1) / (2 * eps) * 2  # d pos01 / d pos = 1/2
    err = np.abs(n(coords2.grad).T - gnum).max(-1) / np.abs(gnum).max()
    assert np.quantile(err, 0.8) < 2e-2, np.quantile(err, [0.5, 0.8, 0.95])
    with pytest.raises(NotImplementedError):
        J.hashgrid_encode(desc, t(np.asarray(lt.offsets, np.uint32)), coords[:2], params)


# ------------------------------------------------------------------ density-grid update (a8)
def test_ogrid_update_vs_oracle():
    from jaxngp_b200 import ogrid as OG
    from oracle import ogrid_np as ON
    G, bound = 32, 4.0
    rng = np.random.Generator(np.random.PCG64(8))
    for cas in (0, 2):
        M = 5000
        idx = rng.integers(0, G ** 3, M, dtype=np.uint32)
        u = rng.random((M, 3), dtype=np.float32)
        got = n(OG.sample_positions(t(idx), t(u), G, cas, bound))
        exp = ON.sample_positions(idx, u, G, cas, bound)
        assert np.allclose(got, exp, rtol=0, atol=1e-6)
        mip = min(bound, 2.0 ** cas)
        assert np.abs(got).max() <= mip  # jittered points stay inside the cascade
    # decay + max with duplicates and dead (-1) cells
    den = rng.random(G ** 3, dtype=np.float32) * 3
    den[rng.integers(0, G ** 3, 2000)] = -1.0
    idx = rng.integers(0, G ** 3, 20000, dtype=np.uint32)
    idx = idx[den[idx] >= 0]  # only alive cells are ever selected (utils/types.py:1154-1156)
    new = (rng.random(idx.shape[0], dtype=np.float32) * 4).astype(np.float32)
    got = n(OG.decay_and_max(t(den), t(idx), t(new)))
    exp = ON.decay_and_max(den, idx, new)
    assert np.array_equal(got, exp)  # multiplication and max are exact: bit-equal
    # in place
    den_t = t(den)
    OG.decay_and_max(den_t, t(idx), t(new), out=den_t)
    assert np.array_equal(n(den_t), exp)
    # threshold = min(thr_max, mean over alive cells)
    for thr_max in (0.5, 100.0):
        got_thr = float(OG.threshold(t(exp), thr_max))
        assert np.isclose(got_thr, float(ON.threshold(exp, thr_max)), rtol=1e-6)


def test_ogrid_full_update_matches_oracle_bitfield(oracle):
    """update_ogrid_density + threshold_ogrid with supplied draws and an analytic density function:
    density grid bit-equal, bitfield bit-equal."""
    from jaxngp_b200 import ogrid as OG, synthetic as S
    from oracle import ogrid_np as ON
    G = 64
    grid = OG.OccupancyDensityGrid(1, G, device=DEV)
    rng = np.random.Generator(np.random.PCG64(3))
    grid.density.copy_(t(rng.random(G ** 3, dtype=np.float32)))
    den0 = n(grid.density).copy()
    M = G ** 3 // 2
    draws_np = dict(first=rng.integers(0, G ** 3, M // 2, dtype=np.uint32), second=rng.integers(0, G ** 3, M // 2, dtype=np.uint32),
                    jitter=rng.random((M, 3), dtype=np.float32))
    draws = {k: t(v) for k, v in draws_np.items()}

    def density_fn(xyz):
        inside = (xyz ** 2).sum(-1) < 0.45 ** 2
        return torch.where(inside, 64.0, 0.0) + 0.001

    OG.update_ogrid_density(grid, density_fn, 0, False, 1.0, 1 << 16, draws=draws)
    idx_np = np.concatenate([draws_np["first"], draws_np["second"]])
    coords = ON.sample_positions(idx_np, draws_np["jitter"], G, 0, 1.0)
    new = (np.where((coords.astype(np.float32) ** 2).sum(-1) < np.float32(0.45 ** 2), 64.0, 0.0) + 0.001).astype(np.float32)
    exp = ON.decay_and_max(den0, idx_np, new)
    got = n(grid.density)
    # points within an ulp of the sphere surface may fall on the other side: allow a handful of cells
    assert (got != exp).sum() <= 3
    thr, mask, bits = OG.threshold_ogrid(grid, 1024, 1.0)
    thr_exp = ON.threshold(got, OG.density_threshold_from_min_step_size(1024, 1.0))
    assert np.isclose(float(thr), float(thr_exp), rtol=1e-6)
    omask, obits = oracle.packbits(float(thr), got)
    assert np.array_equal(n(mask), omask) and np.array_equal(n(bits), obits)
    assert np.array_equal(n(grid.occupancy), obits)
