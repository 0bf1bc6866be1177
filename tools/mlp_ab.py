"""A/B of the fused encoder + MLP forward (mma.sync register chain vs tcgen05) and of the MLP backward on the sample
positions of one C2 training batch (ray-coherent, 2^18 slots); CUDA events, L2 flushed.  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from jaxngp_b200 import nerf as nerf_mod, synthetic, trainops
from jaxngp_b200.trainer import Trainer
from jaxngp_b200.volrendjax import march_rays


def main():
    dev = torch.device("cuda", 0)
    tr = Trainer(device=dev)
    tr.grid.occupancy.copy_(tr.scene.bitfield_gt)
    perm = torch.randint(0, tr.scene.n_pixels, (tr.n_rays,), device=dev, dtype=torch.int32)
    o, d, ts, te, noises, bg = trainops.make_training_rays_rng(perm, tr.scene.transforms, tr.scene.cam, 1.0, tr.rng_state, 1)
    out = march_rays(1 << 18, 1024, 1, 128, 1.0, 0.0, o, d, ts, te, noises, tr.occupancy, raw=True)
    xyzs, dirs = out[6], out[7]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, iters=20):
        ts_ = []
        for _ in range(3 + iters):
            flush.fill_(1)
            torch.cuda._sleep(300_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts_.append(e0.elapsed_time(e1))
        return round(float(np.median(ts_[3:])), 4)

    res = {}
    for impl in ("mma", "umma"):
        res[f"fused_forward_{impl}_ms"] = timed(lambda: nerf_mod.fused_forward(tr.levels, xyzs, 1.0, tr.table, dirs, tr.mlp_flat, want_enc=True, impl=impl))
    drgbs, enc = nerf_mod.fused_forward(tr.levels, xyzs, 1.0, tr.table, dirs, tr.mlp_flat, want_enc=True)
    d_out = torch.randn_like(drgbs)
    for impl in sys.argv[1:] or ("umma", "mma"):
        res[f"mlp_backward_{impl}_ms"] = timed(lambda: nerf_mod.mlp_backward(enc, dirs, tr.mlp_flat, d_out, impl=impl))
    # the two kernels behind "umma" (sixteen chain warps in column-split pairs / eight), alone and with the table scatter fused in
    for split in ("0", "1"):
        os.environ["NGP_B200_MLP_BWD_SPLIT"] = split
        res[f"mlp_backward_split{split}_ms"] = timed(lambda: nerf_mod.mlp_backward(enc, dirs, tr.mlp_flat, d_out, impl="umma"))
        res[f"mlp_backward_scatter_split{split}_ms"] = timed(lambda: nerf_mod.mlp_backward_scatter(
            tr.levels, xyzs, 1.0, enc, dirs, tr.mlp_flat, d_out, tr.mlp_grad, tr.table_grad))
    del os.environ["NGP_B200_MLP_BWD_SPLIT"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
