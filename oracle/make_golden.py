"""TEST INFRASTRUCTURE: generates the golden vectors under tests/golden/ by running the reference's
own CUDA ops (oracle/_ref/libvolrend_ref.so, compiled unmodified from /root/reference) on a B200:

    gpurun -- python oracle/make_golden.py gpurun_out/golden      # then copy into tests/golden/

The vectors pin the CPU oracle (tests/test_oracle_golden.py, runs without a GPU).  Inputs are
regenerated from seeds by tests/inputs.py, so only outputs (and digests of large ones) are stored.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import inputs, refops  # noqa: E402

DEV = "cuda:0"


def t(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a).to(DEV)


def n(x):
    x = x.detach().cpu().numpy()
    return x.view(np.uint32) if x.dtype == np.int32 else x


def canonical_march(out, S):
    """Re-lay the reference's arrival-ordered samples in ray order so the result is deterministic."""
    nxt, exc, valid, rn, rs, idcs, xyzs, dirs, dss, zs = out
    rays = np.nonzero(rn > 0)[0]
    sel = np.concatenate([np.arange(rs[r], rs[r] + rn[r]) for r in rays]) if len(rays) else np.zeros(0, np.int64)
    return dict(next=nxt, exceeded=exc, valid=valid, n_samples=rn, xyzs=xyzs[sel], dss=dss[sel], z_vals=zs[sel],
                idcs=idcs[sel])


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    # morton + packbits
    rng = np.random.Generator(np.random.PCG64(0))
    xyz = rng.integers(0, 1024, (4096, 3), dtype=np.uint32)
    idx = rng.integers(0, 2 ** 30, 4096, dtype=np.uint32)
    den = rng.normal(size=4096 * 8).astype(np.float32)
    mask, bits = refops.packbits(0.25, t(den))
    np.savez_compressed(os.path.join(outdir, "morton_packbits.npz"), morton=n(refops.morton3d(t(xyz))),
                        invert=n(refops.morton3d_invert(t(idx))), mask=n(mask), bits=n(bits))
    # march_rays (no-overflow cases only: the reference's overflow behaviour is order dependent)
    for case in ("scene", "cascades", "dense", "miss"):
        st, arrays = inputs.march_case(case)
        out = [n(x) for x in refops.march_rays(**st, **{k: t(v) for k, v in arrays.items()}, raw=True)]
        can = canonical_march(out, st["total_samples"])
        assert int(out[1][0]) == 0 and int(out[0][0]) < st["total_samples"], "golden cases must not overflow"
        for k in ("xyzs", "dss", "z_vals", "idcs"):  # digest of the full payload + a 4096-sample head
            full = np.ascontiguousarray(can.pop(k))
            can[k + "_sha256"] = np.frombuffer(hashlib.sha256(full.tobytes()).digest(), np.uint8)
            can[k + "_head"] = full[:4096]
        np.savez_compressed(os.path.join(outdir, f"march_{case}.npz"), **can)
        # integrate fwd/bwd on the reference's own march output (kept in its own layout)
        if case in ("scene", "cascades"):
            nxt, exc, valid, rn, rs, idcs, xyzs, dirs, dss, zs = out
            # canonical layout inputs: recompute through the oracle-independent ray order
            rays = np.nonzero(rn > 0)[0]
            sel = np.concatenate([np.arange(rs[r], rs[r] + rn[r]) for r in rays])
            cn = rn.copy()
            cs = np.zeros_like(rs)
            cs[rays] = np.concatenate([[0], np.cumsum(rn[rays])[:-1]]).astype(np.uint32)
            S = st["total_samples"]
            cd, cz, cx = np.zeros(S, np.float32), np.zeros(S, np.float32), np.zeros((S, 3), np.float32)
            cd[: len(sel)], cz[: len(sel)], cx[: len(sel)] = dss[sel], zs[sel], xyzs[sel]
            for scale in (1.0, 0.02):
                drgbs = inputs.drgbs_for(cx, 21, scale)
                bgs = np.random.Generator(np.random.PCG64(22)).random((rn.shape[0], 3), dtype=np.float32)
                dfin = np.random.Generator(np.random.PCG64(99)).normal(size=(rn.shape[0], 4)).astype(np.float32)
                mbs, rgbd, opac = refops.integrate_rays(0.3, t(cs), t(cn), t(bgs), t(cd), t(cz), t(drgbs))
                dbg, dz, dd = refops.integrate_rays_backward(0.3, t(cs), t(cn), t(bgs), t(cd), t(cz), t(drgbs), rgbd, opac, t(dfin))
                keep = slice(0, min(len(sel), 4096))
                np.savez_compressed(os.path.join(outdir, f"integrate_{case}_{scale}.npz"), mbs=int(mbs), rgbd=n(rgbd),
                                    opac=n(opac), dbg=n(dbg), dz=n(dz)[keep], dd=n(dd)[keep],
                                    dz_sum=float(n(dz).astype(np.float64).sum()), dd_sum=n(dd).astype(np.float64).sum(0))
    # inference loop: final image of the slot-refill loop on an analytic field
    from jaxngp_b200 import synthetic as S
    st, fr, bits_, n_slots = inputs.inference_case()
    N = fr["rays_o"].shape[0]
    o, d, ts, te, b = tuple(t(fr[k]) for k in ("rays_o", "rays_d", "t_starts", "t_ends")) + (t(bits_),)
    bg, rgbd, T = torch.ones(N, 3, device=DEV), torch.zeros(N, 4, device=DEV), torch.ones(N, device=DEV)
    term, idx_, nri = torch.ones(n_slots, dtype=torch.bool, device=DEV), torch.zeros(n_slots, dtype=torch.int32, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    rendered, it, ns_total = 0, 0, 0
    while rendered < N and it < 400:
        nri, idx_, ns, ts, xyzs, dss, zs, _ = refops.march_rays_inference(**st, rays_o=o, rays_d=d, t_starts=ts, t_ends=te,
                                                                         occupancy_bitfield=b, next_ray_index_in=nri,
                                                                         terminated=term, indices=idx_)
        x = n(xyzs).reshape(-1, 3)
        drgbs = t(np.concatenate([S.density(x)[:, None] * 0.5, S.colour(x)], -1).reshape(n_slots, -1, 4).astype(np.float32))
        cnt, term, rgbd, T = refops.integrate_rays_inference(bg, rgbd, T, ns, idx_, dss, zs, drgbs)
        rendered += int(cnt)
        ns_total += int(ns.sum())
        it += 1
    np.savez_compressed(os.path.join(outdir, "inference_loop.npz"), rgbd=n(rgbd), T=n(T), iterations=it, ns_total=ns_total)
    print("golden vectors written to", outdir, sorted(os.listdir(outdir)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
