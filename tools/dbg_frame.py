import torch, sys, os
sys.path.insert(0, os.getcwd())
from jaxngp_b200 import renderers
from jaxngp_b200.trainer import Scene, Trainer
DEV="cuda:0"
scene = Scene(DEV, n_views=4)
tr = Trainer(device=DEV, scene=scene)
gen = torch.Generator(device=DEV).manual_seed(0)
for it in range(64):
    tr.train_step(torch.randint(0, scene.n_pixels, (tr.n_rays,), device=DEV, generator=gen, dtype=torch.int32))
    if (it + 1) % 16 == 0:
        tr.update_ogrid()
pose = scene.transforms[2]
R2 = renderers.InferenceRenderer(tr.nerf, scene.cam, tr.occupancy, persistent=False)
R3 = renderers.InferenceRenderer(tr.nerf, scene.cam, tr.occupancy)
a,da = R2.render(pose); b,db = R3.render(pose)
x = R2.rays_rgbd[:640000].clone(); y = R3.rays_rgbd[:640000].clone()
print("counters", R2.counters.tolist(), R3.counters.tolist())
d = (x-y).abs()
print("rays differing", int((d.max(-1).values>0).sum()), "max abs", float(d.max()), "u8 differing px", int((a!=b).any(-1).sum()))
bad = torch.nonzero(d.max(-1).values>0).reshape(-1)[:10]
for i in bad.tolist(): print(i, x[i].tolist(), y[i].tolist())
# second render of persistent: determinism
c,dc = R3.render(pose); print("persistent deterministic", bool(torch.equal(R3.rays_rgbd[:640000], y)), R3.counters.tolist())
e,de = R2.render(pose); print("loop deterministic", bool(torch.equal(R2.rays_rgbd[:640000], x)), R2.counters.tolist())
