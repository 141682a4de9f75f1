#!/usr/bin/env python
"""compute-sanitizer workload for the shared-memory top-of-tree walk (walk_top_kernel, all four instantiations) and its record
builder: device path and the forced-streamed host path, BVH and FLAT modes, full and tiny top budgets, after a refit too."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bvh_b200 import api, capi, scenes
ctx = api.Context.default()
a = scenes.create_n_cubes_aabbs(400)
b = api.Bvh.build(a)
o, d = scenes.ray_endpoints(300_000)
rays = api.Ray.new(o, d)
ctx.set_option("traverse_persistent", 1)
ref = None
for top in (0, 1, 64):
    ctx.set_option("traverse_top", top)
    for stream in (0, 1):
        ctx.set_option("traverse_stream", stream)
        for mode in (capi.TRAVERSE_BVH, capi.TRAVERSE_FLAT):
            off, hits = b.traverse_batch(rays, mode=mode, compact=True)
            if ref is None:
                ref = (off.copy(), hits.copy())
            assert np.array_equal(off, ref[0]) and np.array_equal(hits, ref[1]), (top, stream, mode)
m = a.copy(); m["min"] += 1.0; m["max"] += 1.0
b.refit(m)
ctx.set_option("traverse_top", 1); ctx.set_option("traverse_stream", 0)
off, hits = b.traverse_batch(rays[:50_000])
# a very unbalanced tree whose highest histogram bin alone exceeds a 64-entry budget: no top records, everything walks "below"
x = np.cumsum(np.random.default_rng(3).uniform(0, 1, 3000) ** 8 * 1e4)
sk = np.zeros(3000, dtype=a.dtype); sk["min"] = np.stack([x, np.zeros(3000), np.zeros(3000)], axis=1); sk["max"] = sk["min"] + 0.5
b2 = api.Bvh.build(sk)
o2 = np.stack([x[:2000] + 0.25, np.full(2000, -5.0), np.full(2000, 0.25)], axis=1).astype(np.float32); d2 = np.tile(np.array([[0, 1, 0]], np.float32), (2000, 1))
r2 = api.Ray.new(o2, d2)
res = []
for top in (0, 64, 1):
    ctx.set_option("traverse_top", top)
    res.append(b2.traverse_batch(r2))
assert all(np.array_equal(res[0][0], r[0]) and np.array_equal(res[0][1], r[1]) for r in res)
print("sanitize top done: hits", len(ref[1]), len(hits), len(res[0][1]))
