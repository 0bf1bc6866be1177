"""Kernel-level breakdown of the C2 training step (eager launches, warm L2 like the graph replay).
Usage: python tools/profile_step.py [n_steps] [fused_encoder 0/1]"""
import collections
import os
import sys


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from jaxngp_b200.trainer import Trainer

dev = "cuda:0"
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
fused = bool(int(sys.argv[2])) if len(sys.argv) > 2 else None
tr = Trainer(device=dev, use_graph=False, fused_encoder=fused)
tr.grid.occupancy.copy_(tr.scene.bitfield_gt)
tr.grid.occ_mask.copy_(torch.from_numpy(np.unpackbits(tr.scene.bitfield_gt.cpu().numpy(), bitorder="little").astype(bool)).to(dev))
gen = torch.Generator(device=dev).manual_seed(0)
perms = [torch.randint(0, tr.scene.n_pixels, (tr.n_rays,), device=dev, generator=gen, dtype=torch.int32) for _ in range(steps + 3)]
for k in range(3):
    tr.train_step(perms[k])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for k in range(steps):
        tr.train_step(perms[3 + k])
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[ev.name[:100]]
        a[0] += ev.device_time
        a[1] += 1
tot = sum(v[0] for v in agg.values())
print(f"fused_encoder={tr.fused_encoder}  GPU busy per step: {tot / steps:.1f} us")
for name, (us, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
    print(f"{us / steps:9.1f} us/step {cnt / steps:5.1f}x {100 * us / tot:5.1f}%  {name}")
