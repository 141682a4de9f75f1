#!/usr/bin/env python
"""Dev probe: bench workload (120 k triangles, 1 M rays, f32) walked by the plain persistent kernel and by walk_top_kernel
(option traverse_top); prints the walk-kernel and whole-step times of both and checks the outputs are identical."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bvh_b200 import api, capi, scenes
from bvh_b200.dtypes import RAY3F
L = capi.lib()
ctx = api.Context.default()
dev = torch.device("cuda:0")
N_RAYS = int(os.environ.get("N_RAYS", 1_000_000)); N_CUBES = int(os.environ.get("N_CUBES", 10_000))
if os.environ.get("SCENE") == "sponza":
    from bvh_b200.dtypes import BY_PREC
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sponza_tris.npz"))
    tris = z["vertices"][z["triangles"].astype(np.int64)]
    aabbs = np.zeros(len(tris), dtype=BY_PREC["f32"]["aabb"]); aabbs["min"] = tris.min(axis=1); aabbs["max"] = tris.max(axis=1)
    BOUNDS = (aabbs["min"].min(axis=0), aabbs["max"].max(axis=0))
else:
    aabbs = scenes.create_n_cubes_aabbs(N_CUBES)
    BOUNDS = None
d_aabbs = torch.from_numpy(aabbs.view(np.uint8).reshape(-1)).to(dev)
o, d = scenes.ray_endpoints(N_RAYS, bounds=BOUNDS) if BOUNDS is not None else scenes.ray_endpoints(N_RAYS)
d_o, d_d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
d_rays = torch.empty(N_RAYS * RAY3F.itemsize, dtype=torch.uint8, device=dev)
capi.check(L.bvhgpu_rays_new_dev_f32x3(ctx._h, d_o.data_ptr(), d_d.data_ptr(), N_RAYS, d_rays.data_ptr()))
cap = (12 if BOUNDS is not None else 8) * N_RAYS
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
bvh = api.Bvh.build_dev(d_aabbs.data_ptr(), len(aabbs), ctx=ctx)
ctx.synchronize()
res = {}
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
for top in [int(x) for x in os.environ.get('TOPS', '0,1,3500,1700,800,0').split(',')]:
    ctx.set_option("traverse_top", max(top, 0))
    ctx.set_option("walk_grid", 148 * (-top) if top < 0 else 0)          # top = -k: plain persistent kernel with k CTAs (256 threads) per SM
    d_off = torch.empty(N_RAYS + 1, dtype=torch.int32, device=dev); d_hits = torch.zeros(cap, dtype=torch.int32, device=dev)
    for _ in range(5):
        bvh.traverse_dev(d_rays.data_ptr(), N_RAYS, d_off.data_ptr(), d_hits.data_ptr(), cap)
    ctx.synchronize()
    ctx.set_option("profile", 1)
    walk, stepms = [], []
    for _ in range(15):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        bvh.traverse_dev(d_rays.data_ptr(), N_RAYS, d_off.data_ptr(), d_hits.data_ptr(), cap)
        e1.record(stream)
        walk.append(ctx.get_metric("walk_ms"))
        torch.cuda.synchronize()
        stepms.append(e0.elapsed_time(e1))
    ctx.set_option("profile", 0)
    print(f"traverse_top={top}: walk median {sorted(walk)[7]:.4f} ms  min {min(walk):.4f}   step median {sorted(stepms)[7]:.4f} ms   "
          f"visits/ray {bvh.traverse_dev(d_rays.data_ptr(), N_RAYS, d_off.data_ptr(), d_hits.data_ptr(), cap, want_total=True) and 0 or bvh.traverse_stats()[0] / N_RAYS:.1f}", flush=True)
    res[top] = (d_off.cpu().numpy().copy(), d_hits.cpu().numpy().copy())
print("identical:", all(np.array_equal(res[0][0], v[0]) and np.array_equal(res[0][1], v[1]) for v in res.values()))
