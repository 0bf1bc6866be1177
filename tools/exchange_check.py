"""Check and time the fused gradient exchange (csrc/exchange.cu) against the NCCL path it replaces.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/exchange_check.py [--mode peer|peer-p2p] [--numel 12200000] [--steps 3] [--time 50]

Every rank fills its gradient buffer with its own seeded values, then both arms run the same optimizer steps from the
same parameters:
  NCCL arm : reduce-scatter -> ngp_adam_step on the shard -> all-gather          (jaxngp_b200/dp.py)
  peer arm : ngp_adam_step_exchange                                              (jaxngp_b200/exchange.py)
Checks: (1) every replica of the peer arm holds the SAME bits (one owner reduces each element), (2) the peer arm agrees
with the NCCL arm to float32 summation-order tolerance, (3) moments agree likewise.  Prints one JSON line from rank 0
and exits non-zero on a mismatch.  Needs >= 2 GPUs of one NVLink domain; there is no CPU path.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxngp_b200 import _lib, descriptors, dp, exchange  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="peer", choices=["peer", "peer-p2p"])
    ap.add_argument("--numel", type=int, default=12_196_240 + 10_240)  # C2: table 6,098,120 x 2 + MLP weights
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--time", type=int, default=50, help="timed launches per arm (0 = skip)")
    ap.add_argument("--blocks", type=int, default=None)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    numel = -(-args.numel // (4 * world)) * 4 * world
    lo, hi = dp.shard_bounds(numel, rank, world)
    px = exchange.PeerExchange(numel, rank, world, dev, mode=args.mode, n_blocks=args.blocks)

    gen = torch.Generator(device=dev).manual_seed(7)  # same initial parameters everywhere
    init = torch.randn(numel, device=dev, generator=gen) * 0.1
    decay_begin = min(max(numel - 10_240 - lo, 0), hi - lo) // 4 * 4
    adam = descriptors.make_adam_descriptor(n=hi - lo, decay_begin=decay_begin, lr_init=1e-2, lr_end=1e-4, decay_rate=1 / 3,
                                            transition_steps=10_000, transition_begin=10_000, staircase=True, b1=0.9, b2=0.99,
                                            eps=1e-15, eps_root=1e-15, weight_decay=1e-6, grad_scale=1.0 / world)

    def grads_of(step):
        g = torch.Generator(device=dev).manual_seed(1000 * step + rank)
        return torch.randn(numel, device=dev, generator=g) * (torch.rand(numel, device=dev, generator=g) < 0.3)  # sparse like a table gradient

    # NCCL arm
    p_ref, m_ref, v_ref = init.clone(), torch.zeros(hi - lo, device=dev), torch.zeros(hi - lo, device=dev)
    step_ref = torch.zeros(1, dtype=torch.int32, device=dev)
    # peer arm
    px.params.copy_(init)
    m, v = torch.zeros(hi - lo, device=dev), torch.zeros(hi - lo, device=dev)
    step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    for s in range(args.steps):
        g = grads_of(s)
        g_ref = g.clone()
        shard = dp.reduce_scatter_flat_gradients(g_ref, rank, world)
        _lib.call("ngp_adam_step", [step_ref, p_ref[lo:hi], shard, m_ref, v_ref], adam)
        dp.all_gather_flat_parameters(p_ref, rank, world)
        step_ref += 1
        px.grads.copy_(g)
        px.step(step_dev, m, v, adam, lo)
        step_dev += 1
    torch.cuda.synchronize()

    digest = torch.stack([px.params.view(torch.int32).sum(dtype=torch.int64), px.params.view(torch.int32)[::7].sum(dtype=torch.int64)])
    digests = [torch.empty_like(digest) for _ in range(world)]
    dist.all_gather(digests, digest)
    replicas_identical = all(bool((d == digests[0]).all()) for d in digests)
    err_p = float((px.params - p_ref).abs().max())
    err_m = float((m - m_ref).abs().max())
    err_v = float((v - v_ref).abs().max())
    scale = float(p_ref.abs().max())
    # replicas must hold the SAME bits.  Against the NCCL arm only the summation order of the per-rank gradients differs
    # (ring order there, the switch's order here): moments agree to f32 rounding; Adam divides by sqrt(v), so an element whose
    # summed gradient is nearly cancelling moves by a visible fraction of lr = 1e-2 either way (measured 4e-5 at 8 ranks)
    ok = replicas_identical and err_p <= 2e-4 * max(scale, 1.0) and err_m <= 1e-6 and err_v <= 1e-6

    timing = {}
    if args.time > 0:
        def timed(fn):
            for _ in range(5):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.time):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / args.time], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)

        def nccl_arm():
            shard = dp.reduce_scatter_flat_gradients(g_ref, rank, world)
            _lib.call("ngp_adam_step", [step_ref, p_ref[lo:hi], shard, m_ref, v_ref], adam)
            dp.all_gather_flat_parameters(p_ref, rank, world)

        timing = {"nccl_ms": timed(nccl_arm), "peer_ms": timed(lambda: px.step(step_dev, m, v, adam, lo))}
        link_bytes = 2 * (world - 1) / world * numel * 4  # in + out per rank
        timing["peer_link_gbs_per_rank"] = link_bytes / timing["peer_ms"] / 1e6

    if rank == 0:
        print(json.dumps({"mode": args.mode, "multimem": px.use_multimem, "world": world, "numel": numel, "n_blocks": px.n_blocks,
                          "steps": args.steps, "replicas_identical": replicas_identical, "max_abs_err_params": err_p,
                          "max_abs_err_m": err_m, "max_abs_err_v": err_v, "ok": ok, **timing}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
