"""C3: 800x800 inference rendering through the reference's slot-refill loop (renderers.render_image_inference)
with a briefly trained model.  Prints rays/s and fps.  Usage: python tools/render_bench.py [train_steps]"""
import json
import os
import sys
import time


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from jaxngp_b200 import renderers
from jaxngp_b200.trainer import Scene, Trainer

dev = "cuda:0"
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
scene = Scene(dev)
tr = Trainer(device=dev, scene=scene)
gen = torch.Generator(device=dev).manual_seed(0)
for it in range(steps):
    perm = torch.randint(0, scene.n_pixels, (tr.n_rays,), device=dev, generator=gen, dtype=torch.int32)
    out = tr.train_step(perm)
    if (it + 1) % 16 == 0:
        tr.update_ogrid()
torch.cuda.synchronize()
loss = float(out["loss"])
occ = float(tr.occ_mask.float().mean())
cam = scene.cam
res = []
for view in (0, 7, 33):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rgb, depth = renderers.render_image_inference(tr.nerf, cam, scene.transforms[view], tr.occupancy)
    torch.cuda.synchronize()
    res.append(time.perf_counter() - t0)
rgb_ref, _ = renderers.render_image_inference(tr.nerf, cam, scene.transforms[33], tr.occupancy, grouped=False)
sweep = {"grouped_vs_plain_max_abs_diff": int((rgb_ref.int() - rgb.int()).abs().max())}
for n_slots, cap in ((8192, 8), (65536, 8), (131072, 8), (131072, 16), (262144, 16), (262144, 32), (640000, 16), (640000, 32)):
    ts = []
    for view in (7, 33, 33):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rgb2, _ = renderers.render_image_inference(tr.nerf, cam, scene.transforms[view], tr.occupancy, n_rays=n_slots, march_steps_cap=cap)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    sweep[f"{n_slots}x{cap}"] = {"ms": round(min(ts) * 1e3, 2), "max_abs_diff_vs_8192x8": int((rgb2.int() - rgb.int()).abs().max())}
print(json.dumps(sweep))
fast = {}
for n_slots, cap in ((131072, 16), (262144, 16), (262144, 32), (640000, 32), (640000, 64)):
    R = renderers.InferenceRenderer(tr.nerf, cam, tr.occupancy, n_rays=n_slots, march_steps_cap=cap)
    ts = []
    for view in (7, 33, 33, 33):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rgb3, _ = R.render(scene.transforms[view])
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    fast[f"{n_slots}x{cap}"] = {"ms": round(min(ts) * 1e3, 2), "fps": round(1 / min(ts), 1), "samples": int(R.samples_done),
                                "max_abs_diff_vs_reference_loop": int((rgb3.reshape(800, 800, 3).int() - rgb.int()).abs().max())}
print(json.dumps({"graph_renderer": fast}))

gt = scene.rgbas_u8[33 * 640000:34 * 640000].float() / 255
gt_rgb = (gt[:, :3] * gt[:, 3:] + (1 - gt[:, 3:])).reshape(800, 800, 3)
mse = float(((rgb.float() / 255 - gt_rgb) ** 2).mean())
psnr = -10 * torch.log10(torch.tensor(mse)).item()
print(json.dumps({"train_steps": steps, "final_loss": loss, "occupancy": occ, "frame_s": res, "fps": 1 / min(res),
                  "rays_per_s": 640000 / min(res), "psnr_view33": psnr}))
