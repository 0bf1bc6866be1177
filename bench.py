#!/usr/bin/env python
"""bench.py -- training samples/s of the NeRF hot path on BASELINE.json's configs[1] (C2:
NeRF-synthetic-shaped, 100 synthetic 800x800 posed views, 2^18 rays = 2^18 sample slots per step,
128^3 density grid, diagonal_n_steps=1024), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA kernels via the C ABI)
    python bench.py --impl reference ...                     # the reference path on the host cores (CPU oracle)
    python bench.py --impl reference-gpu ...                 # the reference's own CUDA ops + restated torch encoder

Prints ONE JSON line on rank 0.  A "step" = train_step (ray generation, march, encode, MLP, integrate,
loss, backward, gradient scatter, [all-reduce], Adam) + 1/16 of a density-grid update, i.e. the
reference's training loop at its own cadence (app/nerf/train.py:46-88).  `value` has the step's inputs
(pixel indices) resident in HBM; `e2e` copies them from pinned host memory and reads the loss back
every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


if "--impl" in sys.argv and "reference" in sys.argv[sys.argv.index("--impl") + 1:][:1]:
    # the CPU arm uses every host core it can: torchrun exports OMP_NUM_THREADS=1 to its workers, which would pin
    # numpy's BLAS and the C oracle's OpenMP loops to one thread -- undo that before numpy loads its thread pools
    os.environ["OMP_NUM_THREADS"] = str(host_cores())
    os.environ.pop("MKL_NUM_THREADS", None)
    os.environ.pop("OPENBLAS_NUM_THREADS", None)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS = 1 << 18
TOTAL_SAMPLES = 1 << 18
OGRID_EVERY = 16
METRIC = "train samples/s (NeRF-synthetic shape, 2^18 samples/step)"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def tensor_peak_tf32():
    """Dense tf32 tensor peak for a kernel timed alone: half the MEASURED bf16 burst rate (tf32 runs at half the
    bf16 rate on this part: 1.1 vs 2.25 PFLOP/s nominal); fallback = B200_PROFILING.md's 1.59 PFLOP/s bf16."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["bf16_tflops"]) / 2, "measured bf16_tflops / 2 (MEASURED_PEAKS.json; tf32 = half the bf16 rate)"
    except Exception:
        return 1590.0 / 2, "fallback bf16 1.59 PFLOP/s / 2 (B200_PROFILING.md)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            for k in ("hbm_gbs", "hbm_gbps", "hbm_GBps"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.window = [None, None]

    def mark(self, which):
        self.window[which] = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        allrows = [r for r in self.rows if len(r) >= 8 and r[1].isdigit()]
        t0, t1 = self.window
        rows = [r[1:] for r in allrows if t0 is not None and t1 is not None and t0 <= r[0] <= t1 + 0.1]
        in_window = len(rows)
        if not rows:  # timed region shorter than the sampling period: take the samples nearest to it
            rows = [r[1:] for r in sorted(allrows, key=lambda r: abs(r[0] - (t1 or 0)))[:3]]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median([float(r[0]) for r in rows])), "sm_max_mhz": float(rows[0][1]),
                "power_w": float(np.median([float(r[2]) for r in rows if r[2].replace(".", "").isdigit()] or [0])),
                "samples": len(rows), "samples_inside_timed_region": in_window, "reasons": reasons}


def host_perms(n_steps, rank, seed=1000000007):
    """Per-step pixel indices, disjoint PCG64 streams per rank."""
    rng = np.random.Generator(np.random.PCG64(seed + rank))
    return rng.integers(0, 100 * 800 * 800, size=(n_steps, N_RAYS), dtype=np.int64).astype(np.int32)


# ------------------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist

    from jaxngp_b200 import _lib, encoders, synthetic
    from jaxngp_b200.trainer import Trainer

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started early: nvidia-smi needs a moment before its first sample
    tr = Trainer(device=dev, n_rays=N_RAYS, total_samples=TOTAL_SAMPLES, rank=rank, world_size=world,
                 use_graph=not args.no_graph, exchange=args.exchange)
    # converged occupancy of the analytic scene (see DESIGN.md "bench state"): bitfield and its unpacked mask
    tr.grid.occupancy.copy_(tr.scene.bitfield_gt)
    tr.grid.occ_mask.copy_(torch.from_numpy(np.unpackbits(tr.scene.bitfield_gt.cpu().numpy(), bitorder="little").astype(bool)).to(dev))
    perms_host = torch.from_numpy(host_perms(W + K, rank)).pin_memory()
    perms_dev = perms_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    host_loss = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_events = [torch.cuda.Event() for _ in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, first, e2e):
        samples = torch.zeros((), dtype=torch.int64, device=dev)
        loss_host = 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        launches0 = _lib.launch_count
        e0.record()
        for k in range(n_steps):
            src = perms_host if e2e else perms_dev  # e2e: H2D of every step's inputs from pinned memory
            nxt = src[first + k + 1] if k + 1 < n_steps else None  # next batch: its march overlaps this step's optimizer
            out = tr.train_step(src[first + k], nxt)
            samples += out["measured_batch_size_before_compaction"]
            if (k + 1) % OGRID_EVERY == 0:
                tr.update_ogrid(update_all=False, commit=False)
            if e2e:  # D2H read of every step's result, one step behind so that the host never stalls the device:
                # the loss goes to pinned memory asynchronously and is consumed while the next step runs
                host_loss[k & 1].copy_(out["loss"].reshape(1), non_blocking=True)
                loss_events[k & 1].record()
                if k > 0:
                    loss_events[(k - 1) & 1].synchronize()
                    loss_host = float(host_loss[(k - 1) & 1])
        if e2e:
            loss_events[(n_steps - 1) & 1].synchronize()
            loss_host = float(host_loss[(n_steps - 1) & 1])
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(samples)
        return float(ms), int(samples), loss_host, _lib.launch_count - launches0

    for k in range(W):  # warm-up (captures the step's CUDA graphs), same call pattern as the timed loop
        tr.train_step(perms_dev[k], perms_dev[k + 1] if k + 1 < W else None)
    tr.update_ogrid(update_all=False, commit=False)
    if args.profile:  # ncu --profile-from-start off: only these steps are captured
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for k in range(args.profile):
            tr.train_step(perms_dev[W + k])
        tr.update_ogrid(update_all=False, commit=False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    # K steps per timed region, as asked; the region is repeated R times (same batches, fresh random draws) and the
    # MEDIAN region is reported with the spread beside it: 20 steps are an 11 ms sample otherwise
    R = args.repeats if args.repeats > 0 else (25 if K * 25 <= 2000 else max(5, 2000 // K))
    sampler.mark(0)
    regions = [timed(K, W, e2e=False) for _ in range(R)]
    regions_e2e = [timed(K, W, e2e=True) for _ in range(max(3, R // 3))]
    sampler.mark(1)
    med = sorted(regions, key=lambda r: r[0])[len(regions) // 2]
    ms, samples = med[0], med[1]
    med_e = sorted(regions_e2e, key=lambda r: r[0])[len(regions_e2e) // 2]
    ms_e2e, samples_e2e, loss = med_e[0], med_e[1], med_e[2]
    spread = {"repeats": R, "ms_per_step_min": min(r[0] for r in regions) / K, "ms_per_step_median": ms / K,
              "ms_per_step_max": max(r[0] for r in regions) / K, "e2e_repeats": len(regions_e2e),
              "e2e_ms_per_step_min": min(r[0] for r in regions_e2e) / K, "e2e_ms_per_step_max": max(r[0] for r in regions_e2e) / K,
              "timed_s_total": (sum(r[0] for r in regions) + sum(r[0] for r in regions_e2e)) / 1e3}
    clocks = sampler.stop() if rank == 0 else None

    # launches per step: count the kernels of one eager (un-graphed) step with the torch profiler's CUPTI view
    tr_counts = count_launches(tr, perms_dev[W])

    # dominant kernel in isolation, on the step's own sample positions: CUDA events on the launching
    # stream, L2 flushed between iterations
    roof = dominant_kernel_roofline(tr, perms_dev[W], flush) if rank == 0 else None
    per_op = per_op_vs_reference(tr, perms_dev[W], flush) if rank == 0 and world == 1 and not args.no_extras else None

    render = hashenc = None
    if not args.no_extras:
        del flush
        tr.release_graph()
        render = render_bench(dev, rank, world, scene=tr.scene, cap=args.render_cap)
        if rank == 0:
            hashenc = hashenc_bench(dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = cpu_baseline_sample() if world == 1 and not args.no_cpu_baseline else None
    value = samples / (ms / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (tf32 tensor-core matmuls in the MLP)", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1] (C2): 100 synthetic 800x800 posed views of a procedural scene, "
                               "2^18 rays = 2^18 sample slots per step per GPU, 128^3 density grid, diagonal_n_steps=1024, "
                               "hash grid L=16 T=2^19 F=2, random-init weights, analytic occupancy",
                   "n_rays_per_gpu": N_RAYS, "total_samples_per_gpu": TOTAL_SAMPLES, "ogrid_update_every": OGRID_EVERY,
                   "parallelism": f"ray-sharded dp{world}; " + (
                       f"flat gradient reduce-scatter -> Adam on 1/{world} of the parameters -> parameter all-gather (NCCL)"
                       if tr.peer_exchange is None else
                       f"one fused kernel per rank: summed gradient shard read over NVLink "
                       f"({'multimem.ld_reduce' if tr.peer_exchange.use_multimem else 'peer loads'}) -> Adam on 1/{world} of the "
                       f"parameters -> stored into every replica ({tr.peer_exchange.n_blocks} CTAs)"),
                   "l2": "per-step working set (table+grads+moments+activations ~0.5 GB) exceeds the 126 MB L2; no flush",
                   "cuda_graph": not args.no_graph, "exchange_in_graph": tr._exchange_graph is not None},
        "samples_per_step": samples / K, "timing": spread,
        "e2e": {"value": samples_e2e / (ms_e2e / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": N_RAYS * 4, "d2h_bytes_per_step": 4, "loss": loss,
                "note": "every step: pixel indices copied from pinned host memory, loss copied back to pinned host memory "
                        "(read by the host one step later, while the next step runs)"},
        "gpu_launches": tr_counts["ours"] * K,
        "launches_per_step": tr_counts,
        "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "per_op_vs_ref": per_op, "render": render, "hashenc": hashenc,
    }
    if render is not None:  # BASELINE's second metric, where the driver's per-N records can see it
        line["render_fps"], line["render_rays_per_s"] = render["fps"], render["rays_per_s"]
        line["render_e2e_fps"] = render["e2e"]["fps"]
    if hashenc is not None:
        line["hashenc_fwd_bwd_gbs_T19"] = hashenc["sweep"][0]["fwd_bwd_gbs"]
    if world == 1 and not args.no_ref_gpu and not args.no_extras:
        # the comparator the north star names: the reference's own CUDA ops on this B200 (whole training step)
        torch.cuda.empty_cache()
        ref = run_subprocess_json(["--impl", "reference-gpu", "--steps", str(min(K, 10)), "--warmup", "3"])
        line["ref_gpu"] = ref
        if "value" in ref:
            line["vs_ref_gpu"] = {"ratio": value / ref["value"], "ours_samples_per_s": value, "reference_samples_per_s": ref["value"],
                                  "comparator": ref.get("config", {}).get("workload")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def csrc_sha256():
    """Digest of the kernel sources (what a committed ncu capture must have been taken from to be quoted)."""
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "jaxngp_b200", "csrc", "*.cu*"))):
        h.update(open(f, "rb").read())
    return h.hexdigest()


def count_launches(tr, perm):
    """Kernel launches of one training step, split into this repo's kernels and library/eager ones."""
    import torch
    from torch.profiler import ProfilerActivity, profile
    names = []
    try:
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            tr._step_body(perm)
            torch.cuda.synchronize()
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in ev.name.lower() and "memset" not in ev.name.lower():
                names.append(ev.name)
    except Exception as exc:  # CUPTI not available: fall back to the binding's own call counter
        return {"ours": 6, "library": None, "note": f"profiler unavailable: {exc}"}
    ours = [n for n in names if "ngp::" in n or n.startswith("ngp")]
    return {"ours": len(ours), "library": len(names) - len(ours), "ours_names": sorted(set(n.split("(")[0][-60:] for n in ours))}


def time_once(fn, flush):
    """One launch timed alone: L2 flushed (a 256 MB fill), then a ~150 us spin kernel so that the host has enqueued
    `fn`'s launches before the start event completes (otherwise Python's launch latency sits inside the interval)."""
    import torch
    flush.fill_(1)
    torch.cuda._sleep(300_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def pick_roof(flops, nbytes, ms, tensor_peak_tflops, hbm_peak_gbs):
    """Roofline object of a kernel that both contracts (`flops` algorithmic FLOP per launch) and streams (`nbytes`
    algorithmic bytes per launch) and takes `ms`: the roof that bounds it is the lower one at its arithmetic intensity,
    i.e. the tensor roof iff flops / nbytes >= tensor peak / HBM peak.  Both fractions are reported either way."""
    tflops, gbs = flops / (ms * 1e-3) / 1e12, nbytes / (ms * 1e-3) / 1e9
    intensity, balance = flops / nbytes, tensor_peak_tflops * 1e12 / (hbm_peak_gbs * 1e9)
    both = {"frac_of_tensor_peak": round(tflops / tensor_peak_tflops, 4), "frac_of_hbm_peak": round(gbs / hbm_peak_gbs, 4),
            "arithmetic_intensity_flop_per_byte": round(intensity, 1), "machine_balance_flop_per_byte": round(balance, 1),
            "tensor_peak_tflops": tensor_peak_tflops, "hbm_peak_gbs": hbm_peak_gbs}
    if intensity >= balance:
        return {"bound": "tensor", "achieved": round(tflops, 2), "peak": tensor_peak_tflops, "unit": "TFLOP/s",
                "frac": both["frac_of_tensor_peak"], **both}
    return {"bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak_gbs, "unit": "GB/s", "frac": both["frac_of_hbm_peak"], **both}


def dominant_kernel_roofline(tr, perm, flush):
    """Times every kernel of the step alone, on the step's own tensors, with CUDA events on the launching
    stream and an L2 flush before each launch, then reports the roofline of the slowest one.
    Algorithmic bytes per unit are SURVEY 8(d)'s figures."""
    import torch
    from jaxngp_b200 import _lib, encoders, nerf as nerf_mod, synthetic, trainops
    from jaxngp_b200.volrendjax import march_rays
    from jaxngp_b200.volrendjax.integrating import _integrate_bwd, _integrate_fwd
    hbm, src = peaks()
    dev = perm.device
    sc = tr.scene
    noises, bg = torch.rand(N_RAYS, device=dev), torch.rand(N_RAYS, 3, device=dev)
    o, d, ts, te = trainops.make_training_rays(perm, sc.transforms, sc.cam, synthetic.BOUND)
    march = lambda: march_rays(TOTAL_SAMPLES, synthetic.DIAGONAL_N_STEPS, synthetic.K, synthetic.G, synthetic.BOUND, 0.0,
                               o, d, ts, te, noises, tr.occupancy, raw=True)
    nxt, exc, valid, rn, rs, idcs, xyzs, dirs, dss, zs = march()
    n = xyzs.shape[0]
    used = int((nxt - exc)[0])
    n_hit = int((rn > 0).sum())
    enc = encoders.hashgrid_forward(tr.levels, xyzs, 1.0, tr.table)
    drgbs = nerf_mod.mlp_forward(enc, dirs, tr.mlp_flat)
    eff, fin, opac = _integrate_fwd(rs, rn, bg, dss, zs, drgbs)
    d_fin, _, _ = trainops.huber_loss_grad(fin, valid, perm, sc.rgbas_u8, bg)
    _, _, d_drgbs = _integrate_bwd(synthetic.NEAR, rs, rn, bg, dss, zs, drgbs, fin, opac, d_fin)
    d_enc, _ = nerf_mod.mlp_backward(enc, dirs, tr.mlp_flat, d_drgbs)
    P = tr.shard_hi - tr.shard_lo  # this rank's optimizer shard (all parameters on one GPU)
    # the step runs the backward as a pipeline of `bwd_chunks` slices (trainer._backward): the two backward kernels are
    # timed on the launch the step makes -- the first slice -- so that `achieved`, the ncu launch list and `traffic` agree
    fused_bwd = bool(tr.fused_mlp and tr.bwd_fused_scatter and tr.fused_encoder)  # trainer._backward's first branch
    nb = -(-n // tr.bwd_chunks // 128) * 128 if (not fused_bwd and tr.bwd_chunks > 1 and n >= 128 * 148 * tr.bwd_chunks) else n
    kernels = (
        # name, launch, algorithmic bytes per launch, note
        ("march_rays", march, N_RAYS * 45 + used * 36 + (synthetic.K * synthetic.G ** 3) // 8, "36 B/ray in, 9 B/ray + 36 B/sample out, bitfield"),
        ("nerf_fused_forward", lambda: nerf_mod.fused_forward(tr.levels, xyzs, 1.0, tr.table, dirs, tr.mlp_flat, want_enc=True),
         n * (1164 + 12 + 16), "encoder 1164 B/point (xyz, 128 corner rows, enc kept for the backward) + dir 12 + drgb 16; 18.8 kFLOP/sample"),
        ("integrate_rays", lambda: _integrate_fwd(rs, rn, bg, dss, zs, drgbs), used * 24 + N_RAYS * 40, "24 B/sample + 40 B/ray"),
        ("huber_loss_grad", lambda: trainops.huber_loss_grad(fin, valid, perm, sc.rgbas_u8, bg), N_RAYS * (16 + 1 + 4 + 4 + 12 + 16), "53 B/ray"),
        ("integrate_rays_backward", lambda: _integrate_bwd(synthetic.NEAR, rs, rn, bg, dss, zs, drgbs, fin, opac, d_fin), used * 44 + N_RAYS * 68 + n * 20, "44 B/sample + 68 B/ray + zero-fill 20 B/slot"),
        ("integrate_loss_fused (replaces the three above in the step)",
         lambda: trainops.integrate_loss_fused(synthetic.NEAR, rs, rn, bg, dss, zs, drgbs, valid, perm, sc.rgbas_u8),
         used * 44 + N_RAYS * 70 + n * 16, "44 B/sample + 70 B/ray + zero-fill 16 B/slot"),
        ("nerf_mlp_backward", lambda: nerf_mod.mlp_backward(enc[:nb], dirs[:nb], tr.mlp_flat, d_drgbs[:nb]), nb * (128 + 12 + 16 + 128),
         f"284 B/sample; 56 kFLOP/sample; one launch = {nb} sample slots "
         + ("(stand-alone op: the step runs the fused kernel below instead)" if fused_bwd else f"(the step makes {-(-n // nb)} such launches)")),
        ("hashgrid_a1_backward", lambda: encoders.hashgrid_backward(tr.levels, xyzs[:nb], 1.0, d_enc[:nb], out=tr.table_grad), nb * 1164 + tr.table_numel * 4,
         f"1164 B/point + table zero-fill; one launch = {nb} points (first of {-(-n // nb)} slices)"),
        ("nerf_mlp_backward_scatter (replaces the two above in the step)" if fused_bwd else "nerf_mlp_backward_scatter",
         lambda: nerf_mod.mlp_backward_scatter(tr.levels, xyzs, 1.0, enc, dirs, tr.mlp_flat, d_drgbs, tr.mlp_grad, tr.table_grad),
         n * (128 + 12 + 16 + 12 + 1024) + tr.table_numel * 4,
         "1192 B/sample (enc, dir, cotangent, position, 128 gradient rows of 8 B) + table zero-fill; 56 kFLOP/sample: MLP backward with "
         "the table scatter issued from its d_enc fragments"),
        ("adam_step", lambda: _lib.call("ngp_adam_step", [tr.step_dev, tr.flat_params[tr.shard_lo:tr.shard_hi], tr.flat_grads[tr.shard_lo:tr.shard_hi], tr.adam_m, tr.adam_v], tr.adam_desc), P * 28, "28 B/param"),
    )
    res = {}
    for name, fn, nbytes, note in kernels:
        times = []
        for _ in range(3 + 20):
            times.append(time_once(fn, flush))
        t_ms = float(np.median(times[3:]))
        res[name] = {"ms": round(t_ms, 4), "algorithmic_bytes": int(nbytes), "achieved_gbs": round(nbytes / (t_ms * 1e-3) / 1e9, 1),
                     "frac_of_hbm": round(nbytes / (t_ms * 1e-3) / 1e9 / hbm, 3), "per_unit": note}
    res["nerf_fused_forward"]["achieved_tflops"] = round(n * 18816 / (res["nerf_fused_forward"]["ms"] * 1e-3) / 1e12, 2)
    res["nerf_mlp_backward"]["achieved_tflops"] = round(nb * 56448 / (res["nerf_mlp_backward"]["ms"] * 1e-3) / 1e12, 2)
    fused_key = [k for k in res if k.startswith("nerf_mlp_backward_scatter")][0]
    res[fused_key]["achieved_tflops"] = round(n * 56448 / (res[fused_key]["ms"] * 1e-3) / 1e12, 2)
    per_step = 0 if fused_bwd else -(-n // nb)  # launches per step of the two separate backward kernels; every other kernel runs once
    for name in ("nerf_mlp_backward", "hashgrid_a1_backward"):
        res[name]["launches_per_step"] = per_step
        res[name]["ms_per_step"] = round(res[name]["ms"] * per_step, 4)
    res[fused_key]["launches_per_step"] = 1 if fused_bwd else 0
    res[fused_key]["ms_per_step"] = res[fused_key]["ms"] if fused_bwd else 0.0
    candidates = [k for k in res if k not in ("integrate_rays", "huber_loss_grad", "integrate_rays_backward")]  # replaced by the fused kernel in the step
    top = max(candidates, key=lambda k: res[k].get("ms_per_step", res[k]["ms"]))  # dominant = most time per step
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel: an ncu counter, so it cannot be taken
    # inside this run.  It is reported only from a committed capture of THIS build -- profiles/dram_traffic_r*.json carries
    # the sha256 of the csrc/ sources it was taken from -- and is null the moment a kernel source changes.
    traffic = traffic_src = None
    try:
        import glob
        cap = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "dram_traffic_r*.json")))[-1]))
        if cap.get("csrc_sha256") == csrc_sha256():
            traffic, traffic_src = cap.get(top), "ncu --set full capture of this build: profiles/" + cap.get("file", "dram_traffic")
        else:
            traffic_src = "no ncu capture of this build (kernel sources changed since the last one): null"
    except Exception:
        traffic_src = "no committed ncu capture"
    common = {"kernel": top, "traffic": traffic, "traffic_source": traffic_src, "sample_slots": n, "samples": used, "rays_with_samples": n_hit,
              "note": "each kernel timed alone with an L2 flush before every launch (isolated times sum to more than the "
                      "graph-replayed step, whose kernels find their inputs in L2); limiter per kernel in DESIGN.md section 5",
              "kernels": res}
    if "achieved_tflops" in res[top]:
        # A kernel that both contracts and streams: the roof that bounds it is the lower of the two at its arithmetic
        # intensity (algorithmic FLOP / algorithmic byte against the machine balance, tensor peak / HBM peak).  The MLP
        # backward alone (199 FLOP/B) sits under the tensor roof; with the table scatter fused in (41 FLOP/B: 1192 B per
        # sample) it sits under the HBM roof.  Both fractions are reported either way.
        tpeak, tsrc = tensor_peak_tf32()
        slots = nb if top == "nerf_mlp_backward" else n
        per_slot = 18816 if top == "nerf_fused_forward" else 56448
        roof = pick_roof(slots * per_slot, res[top]["algorithmic_bytes"], res[top]["ms"], tpeak, hbm)
        roof["peak_source"] = tsrc if roof["bound"] == "tensor" else src
        roof.update({"tensor_peak_source": tsrc, "hbm_peak_source": src,
                     "flops_per_unit": ("18,816 FLOP per sample slot (five dense layers), " if per_slot == 18816 else
                                        "56,448 FLOP per sample slot (forward recompute 18,816 + input gradients + weight gradients), ")
                                       + f"x {slots} slots per launch"})
        return {**roof, **common}
    return {"bound": "hbm", "achieved": res[top]["achieved_gbs"], "peak": hbm, "unit": "GB/s",
            "frac": res[top]["frac_of_hbm"], "peak_source": src, **common}



# ------------------------------------------------------------------------------------------- per op, ours vs theirs
def per_op_vs_reference(tr, perm, flush, iters=20):
    """Every compiled op of the reference (deps/volume-rendering-jax/lib/impl/{marching,integrating,packbits}.cu,
    compiled unmodified into oracle/_ref) against this library's drop-in for it, on IDENTICAL inputs at the BASELINE
    sizes: C2 (2^18 rays, 2^18 sample budget) for the training ops, C3 (640,000 rays of one 800x800 frame, 262,144
    slots x 16 steps, first loop iteration) for the inference ops, the 128^3 grid for packbits / morton.  Each launch
    timed alone with CUDA events on the launching stream, L2 flushed before it; median of `iters`.  ratio = theirs / ours."""
    import torch
    from tests import refops
    if not refops.available():
        return {"unavailable": "oracle/_ref/libvolrend_ref.so not built (needs /root/reference at build time)"}
    from jaxngp_b200 import nerf as nerf_mod, renderers, synthetic, trainops, volrendjax as V
    from jaxngp_b200.volrendjax.integrating import _integrate_bwd, _integrate_fwd
    dev, sc = perm.device, tr.scene
    o, d, ts, te, noises, bg = trainops.make_training_rays_rng(perm, sc.transforms, sc.cam, synthetic.BOUND,
                                                               trainops.new_rng_state(dev), seed=1)
    st = (TOTAL_SAMPLES, synthetic.DIAGONAL_N_STEPS, synthetic.K, synthetic.G, synthetic.BOUND, 0.0)
    bits = tr.occupancy
    nxt, exc, valid, rn, rs, idcs, xyzs, dirs, dss, zs = V.march_rays(*st, o, d, ts, te, noises, bits, raw=True)
    drgbs = nerf_mod.fused_forward(tr.levels, xyzs, 1.0, tr.table, dirs, tr.mlp_flat)
    drgbs = torch.cat([drgbs[:, :1] + 4.0 * (xyzs.norm(dim=-1, keepdim=True) < 0.45), drgbs[:, 1:]], -1).contiguous()  # visible medium
    _, fin, opac = _integrate_fwd(rs, rn, bg, dss, zs, drgbs)
    d_fin = torch.randn(N_RAYS, 4, device=dev)
    # C3: one frame's rays, every slot free
    fo, fd = renderers.make_rays_worldspace(sc.cam, sc.transforms[3])
    fts, fte = renderers.make_near_far_from_bound(synthetic.BOUND, fo, fd)
    N, n_slots, cap = fo.shape[0], 262144, 16
    ist = (synthetic.DIAGONAL_N_STEPS, synthetic.K, synthetic.G, cap, synthetic.BOUND, 0.0)
    nri = torch.zeros(1, dtype=torch.int32, device=dev)
    term = torch.ones(n_slots, dtype=torch.bool, device=dev)
    idx0 = torch.zeros(n_slots, dtype=torch.int32, device=dev)
    m_inf = V.march_rays_inference(*ist, fo, fd, fts, fte, bits, nri, term, idx0)
    _, idx1, ns1, _, ixyz, idss, izs = m_inf[:7]
    idrgbs = torch.rand(n_slots, cap, 4, device=dev) * torch.tensor([6.0, 1.0, 1.0, 1.0], device=dev)
    fbg, frgbd, fT = torch.ones(N, 3, device=dev), torch.zeros(N, 4, device=dev), torch.ones(N, device=dev)
    G3 = synthetic.G ** 3
    density = torch.rand(G3, device=dev)
    thr = torch.full((G3,), 0.4, device=dev)
    cells = torch.randint(0, synthetic.G, (G3, 3), device=dev, dtype=torch.int32)
    codes = torch.randint(0, G3, (G3,), device=dev, dtype=torch.int32)
    used = int((nxt - exc)[0])
    ops = (
        ("march_rays (C2)", lambda m: m.march_rays(*st, o, d, ts, te, noises, bits, raw=True), f"{N_RAYS} rays, {used} samples"),
        ("integrate_rays (C2)", lambda m: m.integrate_rays(synthetic.NEAR, rs, rn, bg, dss, zs, drgbs), f"{used} samples"),
        ("integrate_rays_backward (C2)", lambda m: m.integrate_rays_backward(synthetic.NEAR, rs, rn, bg, dss, zs, drgbs, fin, opac, d_fin)
         if m is refops else _integrate_bwd(synthetic.NEAR, rs, rn, bg, dss, zs, drgbs, fin, opac, d_fin), f"{used} samples"),
        ("march_rays_inference (C3, first iteration)", lambda m: m.march_rays_inference(*ist, fo, fd, fts, fte, bits, nri, term, idx0),
         f"{N} rays, {n_slots} slots x {cap}"),
        ("integrate_rays_inference (C3)", lambda m: m.integrate_rays_inference(fbg, frgbd, fT, ns1, idx1, idss, izs, idrgbs),
         f"{n_slots} slots x {cap}"),
        ("pack_density_into_bits (128^3)", lambda m: m.packbits(thr, density), f"{G3} cells"),
        ("morton3d (128^3)", lambda m: m.morton3d(cells), f"{G3} points"),
        ("morton3d_invert (128^3)", lambda m: m.morton3d_invert(codes), f"{G3} codes"),
    )
    table = {}
    for name, fn, size in ops:
        row = {"inputs": size}
        for arm, mod in (("ours_ms", V), ("reference_ms", refops)):
            for _ in range(3):
                fn(mod)
            row[arm] = round(float(np.median([time_once(lambda: fn(mod), flush) for _ in range(iters)])), 4)
        row["ratio"] = round(row["reference_ms"] / row["ours_ms"], 2)
        table[name] = row
    return {"method": "CUDA events around one launch, L2 flushed, median of %d; ratio = reference_ms / ours_ms (> 1: ours faster). "
                      "Both arms go through the same torch/ctypes wrappers (output allocation included); the reference's "
                      "march_rays_inference also synchronises its stream and copies a counter (marching.cu:561-562)" % iters,
            "reference": "deps/volume-rendering-jax/lib/impl/*.cu compiled unmodified for sm_100a (oracle/build_ref.sh)", "ops": table}

# ------------------------------------------------------------------------------------------- extras (C3, C4)
def render_bench(dev, rank, world, scene=None, frames=10, train_steps=(300, 1024), n_slots=262144, cap=16):
    """BASELINE configs[2] (C3): 800x800 frames through march_rays_inference / integrate_rays_inference
    (InferenceRenderer: the reference's slot-refill loop, one CUDA graph per iteration), image rows dealt to the
    ranks in interleaved 32-row bands, one all-gather of the u8 image per frame.  The model is trained here first
    (identical replicas: same seeds on every rank) and measured at two points: after 300 steps -- round 1's
    configuration, a fuzzy field that still marches ~78 samples per ray -- and after 1024 steps (0.6 s of training), when
    the occupancy grid has pruned the empty space the way a model someone would render has; that one is the line's figure.
    Frame time is proportional to the samples marched, which is reported next to every figure."""
    import torch
    import torch.distributed as dist
    from jaxngp_b200 import dp, renderers
    from jaxngp_b200.trainer import Scene, Trainer
    scene = scene if scene is not None else Scene(dev)
    tr = Trainer(device=dev, scene=scene)  # world_size=1: every rank trains the same replica
    gen = torch.Generator(device=dev).manual_seed(0)
    trained = 0

    def train_to(n_steps):
        nonlocal trained
        for it in range(trained, n_steps):
            perm = torch.randint(0, scene.n_pixels, (tr.n_rays,), device=dev, generator=gen, dtype=torch.int32)
            tr.train_step(perm)
            if (it + 1) % 16 == 0:
                tr.update_ogrid()
        trained = max(trained, n_steps)

    train_to(train_steps[0])
    H, Wd = scene.cam["height"], scene.cam["width"]
    rows = dp.tile_rows(H, rank, world).to(dev)
    pixels = (rows[:, None] * Wd + torch.arange(Wd, device=dev)[None, :]).reshape(-1).to(torch.int32)
    # the default renderer (the whole slot-refill loop as one persistent kernel) and, beside it, round 1's renderer (one
    # CUDA-graph replay per loop iteration, host check a batch behind)
    R = renderers.InferenceRenderer(tr.nerf, scene.cam, tr.occupancy, n_rays=min(n_slots, pixels.numel()), march_steps_cap=cap,
                                    pixel_indices=pixels)
    R_loop = renderers.InferenceRenderer(tr.nerf, scene.cam, tr.occupancy, n_rays=min(n_slots, pixels.numel()), march_steps_cap=cap,
                                         pixel_indices=pixels, persistent=False)
    active = [R]
    views = [(7 * k + 3) % scene.n_views for k in range(frames + 2)]
    gather = dp.ImageGather(H, Wd, 3, rank, world, dev) if world > 1 else None

    def frame(v):
        rgb, _ = active[0].render(scene.transforms[v])
        return gather(rgb.reshape(rows.numel(), Wd, 3)) if world > 1 else rgb

    host_frame = torch.empty((H if world > 1 else rows.numel()) * Wd * 3, dtype=torch.uint8).pin_memory()
    poses_host = scene.transforms.cpu().pin_memory()

    def run(e2e):
        """frames rendered back to back; e2e: the pose comes from pinned host memory and the finished u8 frame is copied
        back to pinned host memory every frame (the copy of frame k overlaps nothing: it is waited for before frame k+1,
        as a viewer would)."""
        for v in views[:2]:
            frame(v)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        smp = torch.zeros((), dtype=torch.int64, device=dev)
        e0.record()
        for v in views[2:]:
            if e2e:
                pose = poses_host[v].to(dev, non_blocking=True)
                rgb, _ = active[0].render(pose)
                img = gather(rgb.reshape(rows.numel(), Wd, 3)) if world > 1 else rgb
                host_frame.copy_(img.reshape(-1), non_blocking=True)
                torch.cuda.current_stream().synchronize()
            else:
                img = frame(v)
            smp += active[0].samples_done
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(smp)
        return float(ms) / frames, int(smp), img

    def measure():
        active[0] = R_loop
        ms_loop, smp_loop, img_loop = run(False)
        active[0] = R
        ms_frame, smp, img = run(False)
        ms_frame_e2e, _, _ = run(True)
        # quality of the frame just rendered against the analytic ground truth of that view (white background)
        v = views[-1]
        gt = scene.rgbas_u8[v * H * Wd:(v + 1) * H * Wd].float() / 255
        gt_rgb = gt[:, :3] * gt[:, 3:] + (1 - gt[:, 3:])
        full = img.reshape(-1, 3).float() / 255 if world > 1 else None
        if full is None:
            full = torch.zeros(H * Wd, 3, device=dev)
            full[pixels.long()] = img.reshape(-1, 3).float() / 255
        psnr = float(-10 * torch.log10(((full - gt_rgb) ** 2).mean()))
        return {"rays_per_s": 640000 / (ms_frame * 1e-3), "fps": 1e3 / ms_frame, "ms_per_frame": ms_frame,
                "samples_per_frame": smp / frames, "samples_per_s": smp / frames / (ms_frame * 1e-3), "psnr_last_frame": psnr,
                "model": f"trained {trained} steps in this run (C2 step)",
                "renderer": "persistent whole-frame kernel (ngp_render_frame)" if R.persistent else "loop (graph replay per iteration)",
                "loop_renderer": {"fps": 1e3 / ms_loop, "ms_per_frame": ms_loop, "samples_per_frame": smp_loop / frames,
                                  "same_image": bool(torch.equal(img, img_loop)),
                                  "note": "round 1's renderer on the same model and views: one CUDA-graph replay per slot-refill "
                                          "iteration (march, compaction, tcgen05 encoder + MLP, integrate), host check a batch behind"},
                "e2e": {"fps": 1e3 / ms_frame_e2e, "rays_per_s": 640000 / (ms_frame_e2e * 1e-3), "ms_per_frame": ms_frame_e2e,
                        "h2d_bytes_per_frame": 48, "d2h_bytes_per_frame": int(host_frame.numel()),
                        "note": "pose copied from pinned host memory, finished u8 frame copied to pinned host memory and waited for, every frame"}}

    early = measure()
    train_to(train_steps[1])
    out = measure()
    out.update({"metric": "800x800 inference render rays/s", "frames": frames, "n_gpus": world, "slots_per_gpu": R.n, "march_steps_cap": cap,
                "sharding": "interleaved 32-row bands, all-gather of the u8 image per frame",
                "after_300_steps_round1_configuration": early})
    return out


def hashenc_bench(dev, n=1 << 22, log2_T=(19, 20, 21, 22, 23, 24), iters=10):
    """BASELINE configs[3] (C4): HashGridEncoder forward + backward on 2^22 uniform points, L=16 F=2, T swept across
    the L2/HBM boundary.  GB/s = SURVEY 8(d)'s algorithmic bytes (1164 B/point each way + the gradient-table
    zero-fill) / CUDA-event time.  Also measures an L2-resident copy as the L2 denominator 8(d) asks for."""
    import torch
    from jaxngp_b200 import encoders as E
    hbm, src = peaks()
    g = torch.Generator(device=dev)
    pos = torch.rand(n, 3, device=dev, generator=g.manual_seed(42)) * 2 - 1
    d_enc = torch.randn(n, 32, device=dev, generator=g.manual_seed(44))

    def timed(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    a, b = torch.empty(8 << 20, device=dev), torch.empty(8 << 20, device=dev)  # 32 MB + 32 MB: L2 resident
    l2_ms = timed(lambda: b.copy_(a))
    sweep = []
    for lt2 in log2_T:
        lt = E.make_level_table(16, 2 ** lt2, 2, 16, 2048, 3)
        table = (torch.rand(lt.rows, 2, device=dev, generator=g.manual_seed(43)) * 2 - 1) * 1e-4
        grad = torch.empty_like(table)
        f_ms = timed(lambda: E.hashgrid_forward(lt, pos, 1.0, table))
        b_ms = timed(lambda: E.hashgrid_backward(lt, pos, 1.0, d_enc, out=grad))
        table16 = table.half()
        f16_ms = timed(lambda: E.hashgrid_forward(lt, pos, 1.0, table16))  # fp16 storage variant: 652 B/point
        del table16
        fb = n * 1164, n * 1164 + table.numel() * 4
        sweep.append({"log2_T": lt2, "table_mb": round(table.numel() * 4 / 2 ** 20, 1), "fwd_ms": round(f_ms, 3),
                      "bwd_ms": round(b_ms, 3), "fwd_gbs": round(fb[0] / f_ms / 1e6, 1), "bwd_gbs": round(fb[1] / b_ms / 1e6, 1),
                      "fwd_f16_table_ms": round(f16_ms, 3), "fwd_f16_table_gbs": round(n * 652 / f16_ms / 1e6, 1),
                      "fwd_bwd_gbs": round((fb[0] + fb[1]) / (f_ms + b_ms) / 1e6, 1),
                      "frac_of_hbm": round((fb[0] + fb[1]) / (f_ms + b_ms) / 1e6 / hbm, 3)})
        del table, grad
    return {"metric": "hash-enc fwd+bwd GB/s (algorithmic bytes, 2^22 points, L=16 F=2, f32 table)", "points": n,
            "hbm_peak_gbs": hbm, "peak_source": src, "l2_resident_copy_gbs": round(2 * a.numel() * 4 / l2_ms / 1e6, 1),
            "l2_note": "64 MB working set copied in place of HBM traffic (torch copy kernel, read+write bytes): the L2 "
                       "denominator SURVEY 8(d) asks for; the gather itself is bound by L1 tag lookups (DESIGN.md 5)",
            "inputs": "points within L2-exceeding arrays (pos 50 MB, enc/d_enc 537 MB): no flush needed", "sweep": sweep}



# ------------------------------------------------------------------------------------------- CPU arms
def cpu_step_setup(n_rays):
    from jaxngp_b200 import synthetic as S
    from oracle import hashgrid_np as H
    from oracle import train_np as T
    lv = H.level_table(16, 2 ** 19, 2, 16, 2048, 3)
    rng = np.random.default_rng(0)

    def glorot(i, o):
        lim = np.sqrt(6 / (i + o))
        return rng.uniform(-lim, lim, (i, o)).astype(np.float32)

    params = dict(table=rng.uniform(-1e-4, 1e-4, (int(lv["offsets"][-1]), 2)).astype(np.float32),
                  density_w0=glorot(32, 64), density_w1=glorot(64, 16), rgb_w0=glorot(32, 64), rgb_w1=glorot(64, 64),
                  rgb_w2=glorot(64, 3))
    bits = S.occupancy_bitfield()
    return T, lv, params, bits, T.AdamNp(), rng


def cpu_step(ctx, n_rays, seed):
    from jaxngp_b200 import synthetic as S
    T, lv, params, bits, opt, rng = ctx
    rays = S.training_rays(n_rays, seed=seed)
    xyz_gt = rng.random((n_rays, 4)).astype(np.float32)  # pixel colours: values do not change the work done
    bg = rng.random((n_rays, 3)).astype(np.float32)
    t0 = time.perf_counter()
    m, _ = T.train_step(params, opt, lv, bits, rays, xyz_gt, bg, n_rays)
    return time.perf_counter() - t0, m["measured_batch_size_before_compaction"]


def cpu_baseline_sample(n_rays=1 << 16, min_steps=8, max_steps=64, cpu_seconds=10.0):
    """A bounded sample of the C2 workload on the host cores: whole steps until ~10 s of CPU work are in."""
    from oracle import oracle as O
    O.build()
    O.set_num_threads(host_cores())
    ctx = cpu_step_setup(n_rays)
    cpu_step(ctx, n_rays, 1)  # warm-up
    tot_t = tot_s = 0.0
    steps = 0
    while steps < min_steps or (tot_t < cpu_seconds and steps < max_steps):
        t, s = cpu_step(ctx, n_rays, 100 + steps)
        tot_t += t
        tot_s += s
        steps += 1
    return {"value": tot_s / tot_t, "unit": "samples/s", "cores": O.num_threads(), "kind": "port",
            "sample": f"{steps} full training steps of the CPU oracle (C march/encode/integrate + numpy MLP/Adam) at "
                      f"n_rays = total_samples = 2^16 (1/4 of the C2 step), {tot_t:.1f} s of CPU work"}


def run_reference_cpu(args):
    """--impl reference: the reference path on the host cores (oracle port; jax is absent, oracle/_ref is CUDA)."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    from oracle import oracle as O
    O.build()
    O.set_num_threads(host_cores())  # torchrun exports OMP_NUM_THREADS=1: this arm uses every core it may run on
    n_rays = 1 << 16
    ctx = cpu_step_setup(n_rays)
    K, W = args.steps, max(1, min(args.warmup, 2))
    K = min(K, 60)  # bounded: ~1 s per step
    for k in range(W):
        cpu_step(ctx, n_rays, k)
    tot_t = tot_s = 0.0
    for k in range(K):
        t, s = cpu_step(ctx, n_rays, 100 + k)
        tot_t += t
        tot_s += s
    v = tot_s / tot_t
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": tot_t / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1] (C2) training step; each step a bounded sample: n_rays = "
                               "total_samples = 2^16 (1/4 of the C2 step), same scene, occupancy and model shape"},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": O.num_threads(), "kind": "port",
                         "sample": f"{K} steps at 2^16 rays/samples per step; reference CPU path restated (jax absent; "
                                   "oracle/_ref holds the reference's CUDA ops, not a CPU build)"},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------- reference GPU arm
def run_reference_gpu(args):
    """B-ref-GPU: the reference's own march/integrate CUDA kernels (oracle/_ref, compiled
    unmodified) + a torch restatement of models/encoders.py (index gather / index_add_ scatter) + the
    same torch MLP, loss and torch Adam, driven eagerly like XLA would dispatch them."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    import torch
    from tests import refops
    if not refops.available():
        print(json.dumps({"impl": "reference-gpu", "unavailable": "oracle/_ref/libvolrend_ref.so not built"}))
        return
    from jaxngp_b200 import nerf as nerf_mod, renderers, synthetic
    from jaxngp_b200.trainer import Scene, huber
    from oracle import hashgrid_np as H
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = True
    K, W = args.steps, max(args.warmup, 3)
    scene = Scene(dev)
    lv = H.level_table(16, 2 ** 19, 2, 16, 2048, 3)
    scales = torch.from_numpy(lv["scales"]).to(dev)[:, None, None]
    res = torch.from_numpy(lv["res"].astype(np.int64)).to(dev)[:, None, None]
    offs = torch.from_numpy(lv["offsets"][:-1].astype(np.int64)).to(dev)[:, None, None]
    hashed = torch.from_numpy(lv["hashed"].astype(bool)).to(dev)[:, None, None]
    verts = torch.tensor([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]],
                         dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1000000007)
    model = nerf_mod.NeRF(bound=1.0, device=dev, generator=gen)
    table = model.position_encoder.latents
    params = [table] + model.mlp_parameters()
    opt = torch.optim.Adam(params, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, fused=True)
    Mask = 0xFFFFFFFF

    def encode(xyz):  # models/encoders.py:82-233 restated op for op (materialises [L,n,8,*] like XLA)
        p01 = (xyz + 1.0) / 2.0
        ps = p01[None] * scales + 0.5
        fl = torch.floor(ps)
        v = (fl[:, :, None, :] + verts[None, None]).to(torch.int64)
        x, y, z = v[..., 0], v[..., 1], v[..., 2]
        dense = (x + y * res + z * res * res) & Mask
        hsh = (x ^ ((y * 2654435761) & Mask) ^ ((z * 805459861) & Mask))
        idx = torch.where(hashed, hsh, dense) % (2 ** 19) + offs
        fr = ps - fl
        w = torch.clamp((1 - verts)[None, None] + (2 * verts - 1)[None, None] * fr[:, :, None, :], 0, 1).prod(-1)
        lat = table[idx]
        enc = (lat * w[..., None]).sum(-2)
        return enc.permute(1, 0, 2).reshape(xyz.shape[0], -1)

    bits = scene.bitfield_gt
    perms = torch.from_numpy(host_perms(W + K, 0)).to(dev)

    def step(perm):
        perm = perm.to(torch.int64)
        o, d = scene.rays(perm)
        ts, te = renderers.make_near_far_from_bound(1.0, o, d)
        noises = torch.rand(N_RAYS, device=dev)
        bg = torch.rand(N_RAYS, 3, device=dev)
        mb, valid, rn, rs, idcs, xyzs, dirs, dss, zs = refops.march_rays(TOTAL_SAMPLES, 1024, 1, 128, 1.0, 0.0, o, d, ts,
                                                                         te, noises, bits)
        enc = encode(xyzs)
        x = torch.relu(enc @ model.density_w0) @ model.density_w1
        density = nerf_mod.trunc_exp(x[:, :1])
        h = torch.cat([x, nerf_mod.sh4(dirs)], dim=-1)
        rgb = torch.sigmoid(torch.relu(torch.relu(h @ model.rgb_w0) @ model.rgb_w1) @ model.rgb_w2)
        drgbs = torch.cat([density, rgb], dim=-1).contiguous()
        drgbs_leaf = drgbs.detach().requires_grad_(False)
        _, rgbd, opac = refops.integrate_rays(0.3, rs, rn, bg, dss, zs, drgbs_leaf)
        gt = scene.rgbas_u8[perm].to(torch.float32) / 255
        gt_rgb = gt[:, :3] * gt[:, 3:] + bg * (1 - gt[:, 3:])
        rgbd_leaf = rgbd.detach().requires_grad_(True)
        n_valid = valid.sum()
        loss = torch.where(valid, huber(rgbd_leaf[:, :3], gt_rgb).mean(-1), 0.0).sum() / n_valid
        (d_rgbd,) = torch.autograd.grad(loss, [rgbd_leaf])
        _, _, d_drgbs = refops.integrate_rays_backward(0.3, rs, rn, bg, dss, zs, drgbs_leaf, rgbd, opac, d_rgbd.contiguous())
        opt.zero_grad(set_to_none=True)
        drgbs.backward(d_drgbs)
        opt.step()
        return mb

    for k in range(W):
        step(perms[k])
    torch.cuda.synchronize()
    samples = torch.zeros((), dtype=torch.int64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        samples += step(perms[W + k]).to(torch.int64)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"impl": "reference-gpu", "metric": METRIC, "value": int(samples) / (ms / 1e3), "unit": "samples/s",
                      "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
                      "data": "synthetic", "dtype": "f32 (tf32 matmuls)",
                      "config": {"workload": "C2 training step: the reference's march_rays / integrate_rays / integrate_rays_backward "
                                             "CUDA kernels compiled unmodified for sm_100a + an EAGER torch restatement of its "
                                             "pure-JAX HashGridEncoder, MLP, loss and Adam (jax is absent; XLA would fuse the "
                                             "encoder's [L,n,8,*] intermediates, so this is a soft comparator for the JAX half "
                                             "and an exact one for the CUDA half -- see per_op_vs_ref); no density-grid update"}}))


def run_subprocess_json(extra):
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + extra, capture_output=True, text=True, timeout=900)
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"unavailable": (r.stderr or r.stdout)[-300:]}
    except Exception as exc:
        return {"unavailable": str(exc)}


def main():
    if os.environ.get("NGP_FAULT_DUMP"):  # debugging aid: print every thread's stack and exit if we are still running
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["NGP_FAULT_DUMP"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)  # ~0.15 s per timed region: a few nvidia-smi samples land inside
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C3 render and C4 hash-encoder measurements")
    ap.add_argument("--no-graph", action="store_true", help="eager launches (for ncu launch lists)")
    ap.add_argument("--render-cap", type=int, default=16,
                    help="march_steps_cap of the C3 renderer (samples per slot and iteration; the frame does not depend on "
                         "it, the number of loop iterations does -- worth raising when a rank's rays all fit its slots)")
    ap.add_argument("--exchange", default=None, choices=["auto", "nccl", "peer", "peer-p2p"],
                    help="gradient exchange at N>1: NCCL reduce-scatter/all-gather around Adam (default, or "
                         "NGP_B200_EXCHANGE) or the fused NVLink kernel of csrc/exchange.cu")
    ap.add_argument("--repeats", type=int, default=0, help="timed K-step regions (0 = 25, fewer when K is large); the median is reported")
    ap.add_argument("--profile", type=int, default=0, help="run N steps between cudaProfilerStart/Stop and exit")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-CUDA-ops arm (a subprocess, N=1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_cpu(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
