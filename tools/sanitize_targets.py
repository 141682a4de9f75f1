#!/usr/bin/env python
"""Small workload touching every kernel family once (builders in all strategy combinations, flatten, traversal, queries,
refit / optimize), meant to be run under `compute-sanitizer --tool memcheck|racecheck|initcheck`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bvh_b200 import api, capi, scenes
ctx = api.Context.default()
rng = np.random.default_rng(0)
for prec, ncubes in (("f32", 60), ("f32", 700), ("f64", 300)):
    a = scenes.create_n_cubes_aabbs(ncubes, prec)
    for small, sub, gang in ((-1, -1, -1), (0, 0, 0), (1, 0, 1), (0, 1, 1)):
        ctx.set_option("build_small", small); ctx.set_option("build_subtree", sub); ctx.set_option("build_gang", gang)
        for mode in (capi.BUILD_EXACT_SAH, capi.BUILD_LBVH, capi.BUILD_LBVH_TREELET):
            b = api.Bvh.build(a, prec=prec, mode=mode)
            b.flatten()
            b.free()
    ctx.set_option("build_small", -1); ctx.set_option("build_subtree", -1); ctx.set_option("build_gang", -1)
    b = api.Bvh.build(a, prec=prec)
    o, d = scenes.ray_endpoints(4096, prec=prec)
    rays = api.Ray.new(o, d, prec=prec)
    for mode in (capi.TRAVERSE_BVH, capi.TRAVERSE_FLAT):
        off, hits = b.traverse_batch(rays, mode=mode)
    pts = rng.uniform(-120000, 120000, (512, 3))
    for mode in (capi.TRAVERSE_BVH, capi.TRAVERSE_FLAT):
        ns, nd = b.nearest_to_batch(pts, mode=mode)
    coff, cand = b.nearest_candidates(pts)
    qoff, qh = b.query_batch(capi.QUERY_BALL, np.concatenate([pts[:64], np.full((64, 1), 5000.0)], axis=1))
    m = a.copy()
    mv = rng.choice(len(a), len(a) // 10, replace=False)
    dl = rng.uniform(-10, 10, (len(mv), 3)).astype(a["min"].dtype)
    m["min"][mv] += dl; m["max"][mv] += dl
    b.refit(m)
    m["min"][mv] += dl; m["max"][mv] += dl
    print(prec, ncubes, "optimize rebuilt", b.optimize(m, 1.5), "hits", len(hits), "candidates", len(cand))
    # round 2: compact ray layout, update_shapes (incremental), closest hit (both modes), ordered traversal, triangle nearest_to
    off2, hits2 = b.traverse_batch(rays, compact=True)
    mv2 = rng.choice(len(a), max(1, len(a) // 50), replace=False)
    m["min"][mv2] += 3; m["max"][mv2] += 3
    print("  update rebuilt", b.update_shapes(mv2, m, 1.5), b.update_shapes(mv2, m, 0.0))
    from oracle import oracle as O
    _, tris = O.create_n_cubes(ncubes, prec=prec, want_tris=True)
    b.set_triangles(tris)
    cs, cd, _ = b.closest_hit(rays, triangles=True)
    cs2, cd2, _ = b.closest_hit(rays, triangles=False)
    ts_, td_ = b.nearest_triangles_batch(pts[:256])
    b.traverse_ordered(rays[:512], True)
    b.free()
# D = 2
from bvh_b200.dtypes import BY_PREC_2D
for prec in ("f32", "f64"):
    a2 = np.zeros(500, dtype=BY_PREC_2D[prec]["aabb"]); mn = rng.uniform(-100, 100, (500, 2)); a2["min"] = mn; a2["max"] = mn + rng.uniform(0, 5, (500, 2))
    b2 = api.Bvh2.build(a2, prec=prec)
    b2.nodes_and_index(); b2.flatten()
    r2 = np.zeros(300, dtype=BY_PREC_2D[prec]["ray"])
    o2 = rng.uniform(-120, 120, (300, 2)); d2 = rng.normal(0, 1, (300, 2)); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    r2["origin"], r2["direction"], r2["inv_direction"] = o2, d2, 1.0 / d2
    for mode in (capi.TRAVERSE_BVH, capi.TRAVERSE_FLAT):
        b2.traverse_batch(r2, mode=mode)
    b2.free()
# host path on a batch large enough to be chunked (under the sanitizer the library takes the copy-then-walk form; forced streaming too)
a = scenes.create_n_cubes_aabbs(300)
b = api.Bvh.build(a)
o, d = scenes.ray_endpoints(300_000)
rays = api.Ray.new(o, d)
for opt in (-1, 1):
    ctx.set_option("traverse_stream", opt)
    b.traverse_batch(rays, compact=True)
ctx.set_option("traverse_stream", -1)
# the sharded step with one rank (every exchange kernel; gloo only swaps the handles)
import torch.distributed as dist, torch
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29671")
dist.init_process_group("gloo", rank=0, world_size=1)
from bvh_b200.dist import ShardedTraversal
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to("cuda:0")
sh = ShardedTraversal(b, len(rays), 4 * len(rays))
for _ in range(3):
    sh.step(d_rays.data_ptr(), len(rays))
goff, ghits = sh.fetch()
print("  sharded hits", len(ghits))
sh.close()
dist.destroy_process_group()
b.free()
print("sanitize targets done")
