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
    b.free()
print("sanitize targets done")
