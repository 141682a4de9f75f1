#!/usr/bin/env python
"""10 M f64 shapes, 1 % moved: wall time of bvhgpu_update_f64x3 (C call only, inputs pre-gathered), repeated on ONE tree (moves there and
back), for refit-only and rebuild modes.  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel device times."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bvh_b200 import api, capi, scenes
n_cubes = int(sys.argv[1]) if len(sys.argv) > 1 else 833_334
ctx = api.Context(0)
a = scenes.create_n_cubes_aabbs(n_cubes, "f64")[:10_000_000]
rng = np.random.default_rng(11)
mv = rng.choice(len(a), len(a) // 100, replace=False).astype(np.uint32)
dl = rng.uniform(-10.0, 10.0, (len(mv), 3))
moved = a[mv].copy(); moved["min"] += dl; moved["max"] += dl
home = a[mv].copy()
bvh = api.Bvh.build(a, prec="f64", ctx=ctx)
fn = capi.lib().bvhgpu_update_f64x3
p = lambda x: x.ctypes.data_as(C.c_void_p)
out = {}
for growth, name in ((0.0, "refit_only"), (1.5, "rebuild")):
    ts, rbs, rb = [], [], C.c_size_t(0)
    for k in range(6):                                  # every call moves a FRESH random 1 % of the shapes by <= 10 from their home position
        idx = rng.choice(len(a), len(a) // 100, replace=False).astype(np.uint32)
        src = a[idx].copy(); dl = rng.uniform(-10.0, 10.0, (len(idx), 3)); src["min"] += dl; src["max"] += dl
        ctx.synchronize()
        t0 = time.perf_counter()
        capi.check(fn(bvh._h, p(idx), p(src), len(idx), C.c_double(growth), C.byref(rb)))
        ts.append((time.perf_counter() - t0) * 1e3)
        rbs.append(rb.value)
    out[name] = {"ms_per_call": [round(t, 3) for t in ts], "rebuilt": rbs}
out["sah_cost_after"] = bvh.sah_cost()[0]
print(json.dumps(out))
