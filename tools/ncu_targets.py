#!/usr/bin/env python
"""Workloads for single-kernel ncu captures (run each under `ncu --set full -k regex:<kernel> -c 1 ...`):
    walk          configs[1]: 120 k scene, 1 M create_ray rays, device-resident traversal (walk_persistent_kernel)
    hbm           10 M-triangle f32 scene (640 MB of records), 4 M create_ray rays (walk_persistent_kernel, HBM-bound)
    build10m_f64  configs[4]: 10 M triangles f64, exact SAH build (build_kernel<double>)
    sponza        configs[3] on one GPU: Sponza, 16 M incoherent rays (walk_persistent_kernel, emit_kernel)
Each runs the call `reps` times (default 2: the first warms the pool and the caches, profile the LAST launch with --launch-skip)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bvh_b200 import api, capi, scenes
from bvh_b200.dtypes import BY_PREC

what = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
ctx = api.Context(0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
L = capi.lib()

def dev_rays(o, d):
    n = len(o)
    d_o, d_d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    out = torch.empty(n * 36, dtype=torch.uint8, device=dev)
    capi.check(L.bvhgpu_rays_new_dev_f32x3(ctx._h, d_o.data_ptr(), d_d.data_ptr(), n, out.data_ptr()))
    return out

def traverse(bvh, rays, n, cap, reps):
    off = torch.empty(n + 1, dtype=torch.int32, device=dev); hits = torch.empty(cap, dtype=torch.int32, device=dev)
    for _ in range(reps):
        tot = bvh.traverse_dev(rays.data_ptr(), n, off.data_ptr(), hits.data_ptr(), cap, want_total=True)
    print(what, "rays", n, "hits", tot, "stats", bvh.traverse_stats())

if what == "walk":
    bvh = api.Bvh.build(scenes.create_n_cubes_aabbs(10_000), ctx=ctx)
    o, d = scenes.ray_endpoints(1_000_000)
    traverse(bvh, dev_rays(o, d), 1_000_000, 8_000_000, reps)
elif what == "hbm":
    a = scenes.create_n_cubes_aabbs(833_334)[:10_000_000]
    bvh = api.Bvh.build(a, ctx=ctx)
    o, d = scenes.ray_endpoints(4_000_000)
    traverse(bvh, dev_rays(o, d), 4_000_000, 32_000_000, reps)
elif what == "build10m_f64":
    a = scenes.create_n_cubes_aabbs(833_334, prec="f64")[:10_000_000]
    d_a = torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
    for _ in range(reps):
        b = api.Bvh.build_dev(d_a.data_ptr(), len(a), prec="f64", ctx=ctx)
        ctx.synchronize()
        print("sah", b.sah_cost())
        b.free()
elif what == "sponza":
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sponza_tris.npz"))
    tris = z["vertices"][z["triangles"].astype(np.int64)]
    sp = np.zeros(len(tris), dtype=BY_PREC["f32"]["aabb"]); sp["min"] = tris.min(axis=1); sp["max"] = tris.max(axis=1)
    bvh = api.Bvh.build(sp, ctx=ctx)
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 16_000_000
    o, d = scenes.ray_endpoints(n, bounds=(sp["min"].min(axis=0), sp["max"].max(axis=0)))
    traverse(bvh, dev_rays(o, d), n, 12 * n, reps)
else:
    raise SystemExit("unknown target")
