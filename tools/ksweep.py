import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from bvh_b200 import api, capi, scenes
from bvh_b200.dtypes import BY_PREC
dev = torch.device("cuda", 0); ctx = api.Context(0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
z = np.load("/root/repo/tests/golden/sponza_tris.npz"); tris = z["vertices"][z["triangles"].astype(np.int64)]
sp = np.zeros(len(tris), dtype=BY_PREC["f32"]["aabb"]); sp["min"] = tris.min(axis=1); sp["max"] = tris.max(axis=1)
bvh = api.Bvh.build(sp, ctx=ctx)
def run(rays, label):
    dr = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(dev); n = len(rays)
    off = torch.empty(n + 1, dtype=torch.int32, device=dev); cap = 64 * n; hits = torch.empty(cap, dtype=torch.int32, device=dev)
    for K in (-1,):
        for pers in (2, 1, 0):
            ctx.set_option("traverse_slots", K); ctx.set_option("traverse_persistent", pers)
            ts = []
            for k in range(8):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); bvh.traverse_dev(dr.data_ptr(), n, off.data_ptr(), hits.data_ptr(), cap); b.record(stream); torch.cuda.synchronize()
                if k >= 3: ts.append(a.elapsed_time(b))
            print(f"{label} K={K:2d} persistent={pers}: {sorted(ts)[len(ts)//2]:.3f} ms")
o, d = scenes.pinhole_rays(2048, 2048); run(api.Ray.new(o, d, ctx=ctx), "sponza coherent 4M")
bmin, bmax = sp["min"].min(axis=0), sp["max"].max(axis=0)
o, d = scenes.ray_endpoints(2_000_000, bounds=(bmin, bmax)); run(api.Ray.new(o, d, ctx=ctx), "sponza incoherent 2M")
a = scenes.create_n_cubes_aabbs(10000); b2 = api.Bvh.build(a, ctx=ctx); bvh = b2
o, d = scenes.ray_endpoints(1_000_000); run(api.Ray.new(o, d, ctx=ctx), "cubes 1M")
