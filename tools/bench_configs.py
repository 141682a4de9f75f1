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
# Bvh::update_shapes' own signature: only the changed shapes cross the boundary (5.2 MB instead of 480 MB) and only their root paths are
# touched.  ONE tree, every call moves a fresh random 1 % of the shapes by <= 10 from their home position (steady state of a moving scene;
# the first rebuilding call also grows the memory pool by the builder's scratch and is reported separately).
bu = api.Bvh.build(a64, prec="f64", ctx=ctx)
fnu = capi.lib().bvhgpu_update_f64x3
import ctypes as C
def one_update(growth):
    idx = rng.choice(len(a64), len(a64) // 100, replace=False).astype(np.uint32)
    src = a64[idx].copy(); dlt = rng.uniform(-10.0, 10.0, (len(idx), 3)); src["min"] += dlt; src["max"] += dlt
    rbc = C.c_size_t(0)
    ctx.synchronize()
    t0 = time.perf_counter()
    capi.check(fnu(bu._h, idx.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), len(idx), C.c_double(growth), C.byref(rbc)))
    return (time.perf_counter() - t0) * 1e3, rbc.value
refit_ms = [one_update(0.0)[0] for _ in range(6)]
first_ms, first_rb = one_update(1.5)
steady = [one_update(1.5) for _ in range(6)]
out["update_shapes_10M_f64_1pct"] = {"refit_only_host_call_ms_median": sorted(refit_ms)[3], "rebuild_first_call_ms": first_ms, "rebuild_first_call_rebuilt_shapes": first_rb,
                                     "rebuild_host_call_ms_median": sorted(t for t, _ in steady)[3], "rebuild_host_call_ms_all": [t for t, _ in steady],
                                     "rebuilt_shapes_per_call": [r for _, r in steady], "sah_cost_after": bu.sah_cost()[0], "bytes_h2d_per_call": int(len(a64) // 100 * (4 + 48))}
bu.free()
del a64, a64m
# closest hit on Sponza: 4 M primary rays, triangle mode (Moeller-Trumbore fused, front to back, pruned) and AABB mode
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sponza_tris.npz"))
tri9 = z["vertices"][z["triangles"].astype(np.int64)].astype(np.float32).reshape(-1, 9)
sbvh.set_triangles(tri9)
o, d = scenes.pinhole_rays(2048, 2048)
prim = api.Ray.new(o, d, ctx=ctx)
d_prim = to_dev(prim)
nq = len(prim)
d_s = torch.empty(nq, dtype=torch.int32, device=dev); d_d = torch.empty(nq, dtype=torch.float32, device=dev)
for tri, name in ((1, "triangles"), (0, "aabb")):
    fn = lambda: capi.check(capi.lib().bvhgpu_closest_hit_dev_f32x3(sbvh._h, C.c_void_p(d_prim.data_ptr()), 0, nq, tri, C.c_void_p(d_s.data_ptr()), C.c_void_p(d_d.data_ptr()), None))
    ms = timed(fn)
    hit = int((d_s != -1).sum().item())
    out[f"closest_hit_sponza_4M_primary_{name}"] = {"ms": ms, "Mrays_per_s": nq / ms / 1e3, "rays_with_a_hit": hit}
o, d = scenes.ray_endpoints(2_000_000, bounds=(bmin, bmax))
inc = api.Ray.new(o, d, ctx=ctx); d_inc = to_dev(inc)
d_s2 = torch.empty(len(inc), dtype=torch.int32, device=dev); d_d2 = torch.empty(len(inc), dtype=torch.float32, device=dev)
ms = timed(lambda: capi.check(capi.lib().bvhgpu_closest_hit_dev_f32x3(sbvh._h, C.c_void_p(d_inc.data_ptr()), 0, len(inc), 1, C.c_void_p(d_s2.data_ptr()), C.c_void_p(d_d2.data_ptr()), None)))
out["closest_hit_sponza_2M_incoherent_triangles"] = {"ms": ms, "Mrays_per_s": len(inc) / ms / 1e3, "rays_with_a_hit": int((d_s2 != -1).sum().item())}
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
