"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: flat-gradient all-reduce, density-grid
max-reduce, disjoint per-rank ray streams, and tile-sharded image gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jaxngp_b200 import dp
    import bench
    flat = torch.full((1000,), float(rank + 1))
    dp.allreduce_flat_gradients(flat)
    # ZeRO-1 exchange: reduce-scatter the gradient, update the own shard, all-gather the parameters
    n = 64
    grads = torch.arange(n, dtype=torch.float32) * (rank + 1)
    params = torch.zeros(n)
    lo, hi = dp.shard_bounds(n, rank, world)
    shard = dp.reduce_scatter_flat_gradients(grads, rank, world)
    params[lo:hi] = -shard  # "optimizer" on the shard
    dp.all_gather_flat_parameters(params, rank, world)
    assert torch.equal(params, -3.0 * torch.arange(n, dtype=torch.float32)), params
    # a replica (world_size = 1 trainer inside the 2-rank job, bench.py's render model) exchanges nothing
    own = torch.arange(n, dtype=torch.float32) * (rank + 1)
    assert dp.reduce_scatter_flat_gradients(own, 0, 1).numel() == n
    assert torch.equal(own, torch.arange(n, dtype=torch.float32) * (rank + 1))
    assert torch.equal(dp.all_gather_flat_parameters(own, 0, 1), torch.arange(n, dtype=torch.float32) * (rank + 1))
    grid = torch.arange(16, dtype=torch.float32) * (1 if rank == 0 else -1) + rank
    dp.allreduce_density_grid(grid)
    H, Wd = 100, 7
    rows = dp.tile_rows(H, rank, world, tile=8)
    local = (rows[:, None] * Wd + torch.arange(Wd)[None]).to(torch.float32)[..., None]
    img = dp.gather_image(rows, local, H)
    fast = dp.ImageGather(H, Wd, 1, rank, world, "cpu", dtype=torch.float32, tile=8)(local)
    assert torch.equal(fast, img)
    perms = bench.host_perms(2, rank)
    q.put((rank, flat.sum().item(), grid.tolist(), img[..., 0].tolist(), rows.tolist(), perms[:, :64].tolist()))
    dist.destroy_process_group()


def test_dp_host_logic_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, g0, img0, rows0, p0), (r1, s1, g1, img1, rows1, p1) = res
    assert s0 == s1 == 3000.0                      # SUM over ranks, identical on both
    assert g0 == g1 == [max(i, -i + 1) for i in range(16)]
    assert sorted(rows0 + rows1) == list(range(100)) and not set(rows0) & set(rows1)
    expect = (np.arange(100)[:, None] * 7 + np.arange(7)[None]).tolist()
    assert img0 == expect and img1 == expect      # every rank ends with the whole image
    assert p0 != p1                                # disjoint ray streams per rank


def test_shard_layout_and_image_bands_for_1_2_4_8_ranks():
    """Host-side partition math of the multi-GPU paths, no process group needed: ZeRO-1 shards of the padded flat
    buffer are equal, float4-aligned and cover it; the interleaved row bands of ImageGather partition the image and
    its precomputed scatter puts every band row back where it belongs."""
    from jaxngp_b200 import dp, nerf as nerf_mod
    from jaxngp_b200 import encoders as E
    lt = E.make_level_table(16, 2 ** 19, 2, 16, 2048, 3)
    n_params = lt.rows * 2 + nerf_mod.MLP_NUMEL
    total = -(-n_params // 32) * 32  # trainer.py: padded so that every rank's shard (world <= 8) is float4-aligned
    H, W = 800, 5
    full = (torch.arange(H)[:, None] * W + torch.arange(W)[None]).to(torch.float32)[..., None]
    for world in (1, 2, 4, 8):
        bounds = [dp.shard_bounds(total, r, world) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == total
        assert all(b[1] == bounds[i + 1][0] for i, b in enumerate(bounds[:-1]))
        assert len({hi - lo for lo, hi in bounds}) == 1 and all(lo % 4 == 0 for lo, _ in bounds)
        gathers = [dp.ImageGather(H, W, 1, r, world, "cpu", dtype=torch.float32) for r in range(world)]
        rows = torch.cat([g.local_rows for g in gathers])
        assert sorted(rows.tolist()) == list(range(H))
        # emulate the all-gather: every rank's padded send buffer, stacked in rank order
        recv = torch.zeros(world, gathers[0].max_rows, W, 1)
        for r, g in enumerate(gathers):
            recv[r, : g.n_local] = full[g.local_rows]
        g0 = gathers[0]
        out = torch.empty(H, W, 1)
        out[g0.dst] = recv.reshape(world * g0.max_rows, W, 1)[g0.src]
        assert torch.equal(out, full)
