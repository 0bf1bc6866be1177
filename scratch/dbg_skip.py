import numpy as np, torch, sys
sys.path.insert(0, '.')
from tests import inputs
from tests.test_gpu_parity import t, n
from jaxngp_b200 import _lib, descriptors, volrendjax as V
from oracle import oracle as O
O.build()
case = "scene"
st, arr = inputs.march_case(case)
st = dict(diagonal_n_steps=st["diagonal_n_steps"], K=st["K"], G=st["G"], march_steps_cap=12, bound=st["bound"], stepsize_portion=st["stepsize_portion"])
N = arr["rays_o"].shape[0]
o, d, ts, te, bits = (t(arr[k]) for k in ("rays_o", "rays_d", "t_starts", "t_ends", "occupancy_bitfield"))
ts2 = torch.empty_like(ts)
desc = descriptors.make_marching_inference_descriptor(N, N, st["diagonal_n_steps"], st["K"], st["G"], 12, st["bound"], st["stepsize_portion"])
_lib.call("ngp_march_rays_skip_empty", [o, d, ts, te, bits, ts2], desc)
term, idx, nri = t(np.ones(N, np.bool_)), t(np.zeros(N, np.uint32)), t(np.zeros(1, np.uint32))
a = V.march_rays_inference(**st, rays_o=o, rays_d=d, t_starts=ts, t_ends=te, occupancy_bitfield=bits, next_ray_index_in=nri, terminated=term, indices=idx)
b = V.march_rays_inference(**st, rays_o=o, rays_d=d, t_starts=ts2, t_ends=te, occupancy_bitfield=bits, next_ray_index_in=nri, terminated=term, indices=idx)
oa = O.march_rays_inference(**st, rays_o=arr["rays_o"], rays_d=arr["rays_d"], t_starts=arr["t_starts"], t_ends=arr["t_ends"], occupancy_bitfield=arr["occupancy_bitfield"], next_ray_index_in=np.zeros(1,np.uint32), terminated=np.ones(N,np.bool_), indices=np.zeros(N,np.uint32))
ob = O.march_rays_inference(**st, rays_o=arr["rays_o"], rays_d=arr["rays_d"], t_starts=n(ts2), t_ends=arr["t_ends"], occupancy_bitfield=arr["occupancy_bitfield"], next_ray_index_in=np.zeros(1,np.uint32), terminated=np.ones(N,np.bool_), indices=np.zeros(N,np.uint32))
names = ("nri","idx","ns","t","xyz","ds","z")
for k in range(7):
    x, y = n(a[k]), n(b[k])
    xa, yb = np.asarray(oa[k]), np.asarray(ob[k])
    print(names[k], "a==b", np.array_equal(x.view(np.uint32), y.view(np.uint32)), "a==oa", np.array_equal(x.view(np.uint32).ravel(), xa.view(np.uint32).ravel()), "b==ob", np.array_equal(y.view(np.uint32).ravel(), yb.view(np.uint32).ravel()), "oa==ob", np.array_equal(xa.view(np.uint32), yb.view(np.uint32)))
x, y = n(a[3]), n(b[3])
bad = np.nonzero(x.view(np.uint32) != y.view(np.uint32))[0]
print("bad t:", bad[:10], len(bad))
for r in bad[:5]:
    print(r, "ts", arr["t_starts"][r], "ts2", n(ts2)[r], "te", arr["t_ends"][r], "a.t", x[r], "b.t", y[r], "ns", n(a[2])[r], n(b[2])[r], "oa.t", oa[3][r], "ob.t", ob[3][r])
