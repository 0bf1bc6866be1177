"""Kernel-level breakdown of one 800x800 inference frame (InferenceRenderer) after a short training run.
Usage: python tools/profile_render.py [train_steps] [n_slots] [cap]"""
import collections
import json
import os
import sys
import time


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
from torch.profiler import ProfilerActivity, profile

from jaxngp_b200 import renderers
from jaxngp_b200.trainer import Scene, Trainer

dev = "cuda:0"
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n_slots = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 16
scene = Scene(dev)
tr = Trainer(device=dev, scene=scene)
gen = torch.Generator(device=dev).manual_seed(0)
t0 = time.perf_counter()
for it in range(steps):
    perm = torch.randint(0, scene.n_pixels, (tr.n_rays,), device=dev, generator=gen, dtype=torch.int32)
    out = tr.train_step(perm)
    if (it + 1) % 16 == 0:
        tr.update_ogrid()
torch.cuda.synchronize()
train_s = time.perf_counter() - t0
R = renderers.InferenceRenderer(tr.nerf, scene.cam, tr.occupancy, n_rays=n_slots, march_steps_cap=cap)
ts = []
for view in (7, 33, 33, 33):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rgb, _ = R.render(scene.transforms[view])
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
samples = int(R.samples_done)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    R.render(scene.transforms[33])
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[ev.name[:90]]
        a[0] += ev.device_time
        a[1] += 1
tot = sum(v[0] for v in agg.values())
print(json.dumps({"train_steps": steps, "train_s": round(train_s, 2), "loss": float(out["loss"]),
                  "occupancy": float(tr.occ_mask.float().mean()), "slots": n_slots, "cap": cap,
                  "frame_ms": [round(t * 1e3, 2) for t in ts], "samples": samples, "gpu_busy_ms": round(tot / 1e3, 2)}))
for name, (us, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"{us / 1e3:9.3f} ms {cnt:5d}x {100 * us / tot:5.1f}%  {name}")
