#!/usr/bin/env python
"""Device-resident timings of every BASELINE.json configuration on one B200 (CUDA events, median of N).
Writes gpurun_out/configs.json; the headline bench line comes from bench.py, this is the supporting table."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bvh_b200 import api, capi, scenes
from bvh_b200.dtypes import BY_PREC

dev = torch.device("cuda", 0)
ctx = api.Context(0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
out = {}

def timed(fn, reps=10, warm=3):
    ts = []
    for k in range(warm + reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); torch.cuda.synchronize(dev)
        if k >= warm: ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

def to_dev(a):
    return torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)

def build_time(aabbs, prec, mode):
    d = to_dev(aabbs)
    def f():
        b = api.Bvh.build_dev(d.data_ptr(), len(aabbs), prec=prec, ctx=ctx, mode=mode); b.flatten_dev(); f.last = b
    def g():
        f(); 
    ms = []
    for k in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); f(); b.record(stream); torch.cuda.synchronize(dev)
        if k >= 3: ms.append(a.elapsed_time(b))
        f.last.free()
    ms.sort()
    return ms[len(ms) // 2]

def trav_time(bvh, rays, prec="f32"):
    d = BY_PREC[prec]
    dr = to_dev(rays)
    n = len(rays)
    off = torch.empty(n + 1, dtype=torch.int32, device=dev)
    cap = 64 * n
    hits = torch.empty(cap, dtype=torch.int32, device=dev)
    tot = bvh.traverse_dev(dr.data_ptr(), n, off.data_ptr(), hits.data_ptr(), cap, want_total=True)
    ms = timed(lambda: bvh.traverse_dev(dr.data_ptr(), n, off.data_ptr(), hits.data_ptr(), cap))
    visits, _ = bvh.traverse_stats()
    return ms, tot, visits

# configs[0]/[1]: cube scenes, f32
for n_cubes in (100, 1000, 10000, 100000):
    a = scenes.create_n_cubes_aabbs(n_cubes)
    for mode, name in ((capi.BUILD_EXACT_SAH, "exact_sah"), (capi.BUILD_LBVH, "lbvh"), (capi.BUILD_LBVH_TREELET, "lbvh_treelet")):
        ms = build_time(a, "f32", mode)
        out[f"build_{name}_{12*n_cubes}_f32"] = {"ms": ms, "Mprims_per_s": len(a) / ms / 1e3}
a = scenes.create_n_cubes_aabbs(10000)
bvh = api.Bvh.build(a, ctx=ctx)
o, d = scenes.ray_endpoints(1_000_000)
rays = api.Ray.new(o, d, ctx=ctx)
ms, tot, v = trav_time(bvh, rays)
out["traverse_120k_1M_create_ray"] = {"ms": ms, "Mrays_per_s": 1e3 / ms, "hits": tot, "visits_per_ray": v / 1e6}
lb = api.Bvh.build(a, ctx=ctx, mode=capi.BUILD_LBVH)
ms, tot, v = trav_time(lb, rays)
out["traverse_120k_1M_create_ray_on_lbvh_tree"] = {"ms": ms, "Mrays_per_s": 1e3 / ms, "hits": tot, "visits_per_ray": v / 1e6,
                                                    "sah_cost_ratio_vs_exact": lb.sah_cost()[0] / bvh.sah_cost()[0]}
# configs[2]/[3]: Sponza
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sponza_tris.npz"))
tris = z["vertices"][z["triangles"].astype(np.int64)]
sp = np.zeros(len(tris), dtype=BY_PREC["f32"]["aabb"]); sp["min"] = tris.min(axis=1); sp["max"] = tris.max(axis=1)
out["build_exact_sah_sponza_66450_f32"] = {"ms": (m := build_time(sp, "f32", capi.BUILD_EXACT_SAH)), "Mprims_per_s": len(sp) / m / 1e3}
sbvh = api.Bvh.build(sp, ctx=ctx)
for mode, name in ((capi.BUILD_LBVH, "lbvh"), (capi.BUILD_LBVH_TREELET, "lbvh_treelet")):
    m = build_time(sp, "f32", mode)
    tb = api.Bvh.build(sp, ctx=ctx, mode=mode)
    out[f"build_{name}_sponza_66450_f32"] = {"ms": m, "Mprims_per_s": len(sp) / m / 1e3, "sah_cost_ratio_vs_exact": tb.sah_cost()[0] / sbvh.sah_cost()[0],
                                            "sah_geometric_ratio_vs_exact": tb.sah_cost()[1] / sbvh.sah_cost()[1]}
o, d = scenes.pinhole_rays(2048, 2048)
ms, tot, v = trav_time(sbvh, api.Ray.new(o, d, ctx=ctx))
out["traverse_sponza_4M_coherent"] = {"ms": ms, "Mrays_per_s": 4.194304e3 / ms, "hits": tot, "visits_per_ray": v / 4194304}
bmin, bmax = sp["min"].min(axis=0), sp["max"].max(axis=0)
o, d = scenes.ray_endpoints(2_000_000, bounds=(bmin, bmax))
ms, tot, v = trav_time(sbvh, api.Ray.new(o, d, ctx=ctx))
out["traverse_sponza_2M_incoherent_shard"] = {"ms": ms, "Mrays_per_s": 2e3 / ms, "hits": tot, "visits_per_ray": v / 2e6}
# configs[4]: 10 M triangles, f64
a64 = scenes.create_n_cubes_aabbs(833_334, "f64")[:10_000_000]
ms = build_time(a64, "f64", capi.BUILD_EXACT_SAH)
out["build_exact_sah_10M_f64"] = {"ms": ms, "Mprims_per_s": len(a64) / ms / 1e3}
ms = build_time(a64, "f64", capi.BUILD_LBVH)
out["build_lbvh_10M_f64"] = {"ms": ms, "Mprims_per_s": len(a64) / ms / 1e3}
ms = build_time(a64, "f64", capi.BUILD_LBVH_TREELET)
out["build_lbvh_treelet_10M_f64"] = {"ms": ms, "Mprims_per_s": len(a64) / ms / 1e3}
b64 = api.Bvh.build(a64, prec="f64", ctx=ctx)
t0 = time.perf_counter(); b64.refit(a64); out["refit_10M_f64_host_call_ms"] = (time.perf_counter() - t0) * 1e3
# optimize (Bvh::update_shapes counterpart): move 1 % of the 10 M shapes by <= 10.0 (optimization.rs:702), host call incl. the 480 MB H2D
rng = np.random.default_rng(11)
mv = rng.choice(len(a64), len(a64) // 100, replace=False)
a64m = a64.copy(); dl = rng.uniform(-10.0, 10.0, (len(mv), 3)); a64m["min"][mv] += dl; a64m["max"][mv] += dl
c0 = b64.sah_cost()[0]
t0 = time.perf_counter(); rb = b64.optimize(a64m, 1.5); t_opt = (time.perf_counter() - t0) * 1e3
c_opt = b64.sah_cost()[0]
b64.free()
b64r = api.Bvh.build(a64, prec="f64", ctx=ctx); b64r.refit(a64m); c_refit = b64r.sah_cost()[0]; b64r.free()
b64f = api.Bvh.build(a64m, prec="f64", ctx=ctx); c_fresh = b64f.sah_cost()[0]; b64f.free()
out["optimize_10M_f64_1pct"] = {"host_call_ms": t_opt, "rebuilt_shapes": rb, "sah_cost_before_motion": c0, "sah_cost_refit_only": c_refit,
                                "sah_cost_optimize": c_opt, "sah_cost_fresh_build": c_fresh}
del a64, a64m
# the reference's own update benchmark shape (optimization.rs:693-725): 120 k triangles, p % moved by <= 10.0.  (The oracle's update_shapes
# on the same motion is timed and costed in tests/test_gpu_parity.py::test_optimize_vs_reference_update_120k -> gpurun_out/optimize_vs_reference.json.)
a = scenes.create_n_cubes_aabbs(10000)
for pct in (1, 10, 50):
    rng = np.random.default_rng(pct)
    mv = rng.choice(len(a), len(a) * pct // 100, replace=False)
    am = a.copy(); dl = rng.uniform(-10.0, 10.0, (len(mv), 3)).astype(np.float32); am["min"][mv] += dl; am["max"][mv] += dl
    g = api.Bvh.build(a, ctx=ctx)
    ts = []
    for k in range(5):
        gg = api.Bvh.build(a, ctx=ctx); ctx.synchronize()
        t0 = time.perf_counter(); rb = gg.optimize(am, 1.5); ts.append((time.perf_counter() - t0) * 1e3)
        c_opt = gg.sah_cost()[0]; gg.free()
    g.refit(am); c_refit = g.sah_cost()[0]; g.free()
    f = api.Bvh.build(am, ctx=ctx); c_fresh = f.sah_cost()[0]; f.free()
    out[f"optimize_120k_f32_{pct}pct"] = {"host_call_ms": sorted(ts)[2], "rebuilt_shapes": rb, "sah_cost_refit_only": c_refit,
                                         "sah_cost_optimize": c_opt, "sah_cost_fresh_build": c_fresh}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
