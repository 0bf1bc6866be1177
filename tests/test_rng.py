"""The random inputs of the path (csrc/common.cuh Philox4x32-10): known-answer vectors of the published generator on the
CPU restatement, the kernels against that restatement bit for bit, the device-resident call counter, and the cell draws
of the density-grid update (utils/types.py:1166-1206) as one op."""
import numpy as np
import pytest
import torch

DEV = "cuda:0"


def test_philox_known_answer_vectors():
    from oracle import philox_np as P
    for ctr, key, expected in P.KAT:
        got = P.philox4x32_10(*[np.uint32(c) for c in ctr], *key)
        assert [int(x) for x in got] == list(expected)
    u = P.uniform4(1000, 3, 0x123456789ABCDEF, 2)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    assert abs(float(u.mean()) - 0.5) < 0.02


@pytest.mark.gpu
def test_philox_kernel_matches_restatement_bit_for_bit():
    from jaxngp_b200 import trainops
    from oracle import philox_np as P
    for n, counter, seed, stream in ((1, 0, 0, 0), (1000, 7, 0xDEADBEEF12345678, 1), (70001, 2 ** 32 - 1, 5, 2)):
        got = trainops.philox_uniform(n, counter, seed, stream, DEV).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), P.uniform4(n, counter, seed, stream).view(np.uint32))


@pytest.mark.gpu
def test_training_rays_rng_draws_call_counter_blocks_and_advances():
    """make_training_rays_rng = make_training_rays + the Philox block of the current call; the counter moves by one per
    launch, eagerly and under CUDA-graph replay."""
    from jaxngp_b200 import synthetic as S, trainops
    n = 5000
    tf = torch.from_numpy(S.poses(4)).to(DEV)
    cam = S.camera()
    perm = torch.randint(0, 4 * cam["width"] * cam["height"], (n,), device=DEV, dtype=torch.int32)
    state = trainops.new_rng_state(DEV, counter=11)
    ref = trainops.make_training_rays(perm, tf, cam, 1.0)
    for call in range(3):
        o, d, ts, te, noises, bgs = trainops.make_training_rays_rng(perm, tf, cam, 1.0, state, seed=99)
        for a, b in zip((o, d, ts, te), ref):
            assert torch.equal(a, b)
        u = trainops.philox_uniform(n, 11 + call, 99, trainops.STREAM_TRAIN_RAYS, DEV)
        assert torch.equal(noises, u[:, 0]) and torch.equal(bgs, u[:, 1:].contiguous())
        assert state.tolist() == [12 + call, 0]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = trainops.make_training_rays_rng(perm, tf, cam, 1.0, state, seed=99)
    before = int(state[0])  # capture does not execute
    for replay in range(3):
        g.replay()
        u = trainops.philox_uniform(n, before + replay, 99, trainops.STREAM_TRAIN_RAYS, DEV)
        assert torch.equal(out[4], u[:, 0])
    assert int(state[0]) == before + 3


@pytest.mark.gpu
@pytest.mark.parametrize("G,culled", [(128, False), (32, True)])
def test_ogrid_draw_cells(G, culled):
    """ngp_ogrid_draw_cells against a host restatement of the same draws: first half uniform among the trainable cells,
    second half the k-th occupied cell by inverse CDF, jittered positions inside the drawn cells."""
    from jaxngp_b200 import ogrid
    from oracle import philox_np as P
    rng = np.random.Generator(np.random.PCG64(G))
    G3 = G ** 3
    K = 2
    grid = ogrid.OccupancyDensityGrid(K, G, device=DEV)
    mask = rng.random(K * G3) < 0.07
    mask[G3:G3 + 5000] = False
    alive_np = None
    if culled:
        alive_flag = rng.random(K * G3) < 0.6
        mask &= alive_flag
        grid.alive_indices = torch.from_numpy(np.nonzero(alive_flag)[0].astype(np.int32)).to(DEV)
        grid.alive_indices_offset = [0, int(alive_flag[:G3].sum()), int(alive_flag.sum())]
    grid.occ_mask.copy_(torch.from_numpy(mask).to(DEV))
    grid.occupancy.copy_(torch.from_numpy(np.packbits(mask, bitorder="little")).to(DEV))
    grid.seed = 424242
    for cas in range(K):
        if culled:
            alive_np = np.nonzero(alive_flag[cas * G3:(cas + 1) * G3])[0]
        n_alive = G3 if alive_np is None else len(alive_np)
        mip_bound = min(4.0, 2.0 ** cas)
        for update_all in (True, False):
            counter = int(grid.rng_state[0])
            idx, coords = ogrid.draw_cells(grid, cas, update_all, 4.0)
            assert int(grid.rng_state[0]) == counter + 1
            idx_np, coords_np = idx.cpu().numpy(), coords.cpu().numpy()
            m = idx_np.shape[0]
            u = P.uniform4(m, counter, grid.seed, ogrid.STREAM_OGRID)
            bits = P.philox4x32_10(np.arange(m, dtype=np.uint32), np.uint32(counter), np.uint32(ogrid.STREAM_OGRID), np.uint32(0),
                                   grid.seed & 0xFFFFFFFF, grid.seed >> 32)[0]
            if update_all:
                expect = np.arange(G3) if alive_np is None else alive_np
            else:
                half = max(1, max(1, n_alive // 2) // 2)
                assert m == 2 * half
                k = ((bits[:half].astype(np.uint64) * np.uint64(n_alive)) >> np.uint64(32)).astype(np.int64)
                first = k if alive_np is None else alive_np[k]
                occ = np.nonzero(mask[cas * G3:(cas + 1) * G3])[0]
                total = np.float32(len(occ))
                kk = np.ceil(total * (np.float32(1) - u[half:, 0])).astype(np.int64)
                second = occ[np.clip(kk, 1, len(occ)) - 1]
                expect = np.concatenate([first, second])
                assert mask[cas * G3:(cas + 1) * G3][idx_np[half:]].all()  # second half: occupied cells only
            assert np.array_equal(idx_np, expect)
            # positions (utils/types.py:1193-1206): cell centre line + jitter within half a cell
            def compact(x):
                x = x & 0x49249249
                x = (x | (x >> 2)) & 0xC30C30C3
                x = (x | (x >> 4)) & 0x0F00F00F
                x = (x | (x >> 8)) & 0xFF0000FF
                x = (x | (x >> 16)) & 0x0000FFFF
                return x
            cells = np.stack([compact(idx_np.astype(np.int64) >> s) for s in range(3)], -1).astype(np.float32)
            half_cell = np.float32(mip_bound) / np.float32(G)
            base = (cells / np.float32(G - 1) * np.float32(2) - np.float32(1)) * (np.float32(mip_bound) - half_cell)
            jit = np.maximum(-half_cell, u[:, 1:] * (half_cell + half_cell) - half_cell)
            assert np.allclose(coords_np, base + jit, rtol=0, atol=1e-6)
            assert (np.abs(coords_np - base) <= half_cell * (1 + 1e-6)).all()


@pytest.mark.gpu
def test_ogrid_draw_cells_empty_grid_and_argument_checks():
    from jaxngp_b200 import _lib, descriptors, ogrid
    grid = ogrid.OccupancyDensityGrid(1, 32, device=DEV)
    grid.occupancy.zero_()
    idx, _ = ogrid.draw_cells(grid, 0, False, 1.0)
    half = idx.shape[0] // 2
    assert (idx[half:] == 0).all()  # all-zero CDF: searchsorted lands on the first cell
    with pytest.raises(_lib.NgpError):
        _lib.call("ngp_ogrid_draw_cells", [grid.occupancy, 0, grid.rng_state, idx, torch.empty(idx.shape[0], 3, device=DEV)],
                  descriptors.make_ogrid_draw_descriptor(32 ** 3, 1, 32 ** 3, False, False, 4, 4, 1.0, 0, 2))
