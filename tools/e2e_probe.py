#!/usr/bin/env python
"""Where the host-pointer traversal's time goes (one B200): per-call wall time distribution, host-side phase stamps and the device
event marks of bvhgpu_traverse_od_f32x3 / bvhgpu_traverse_f32x3 on the 120 k scene, 1 M create_ray rays."""
import ctypes as C, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bvh_b200 import api, capi, scenes
N = 1_000_000
ctx = api.Context(0)
L = capi.lib()
bvh = api.Bvh.build(scenes.create_n_cubes_aabbs(10_000), ctx=ctx)
o, d = scenes.ray_endpoints(N)
rays = api.Ray.new(o, d, ctx=ctx)
out = {}
QUICK = os.environ.get("E2E_QUICK") == "1"        # the compact layout, automatic streaming only; one line of output
for name, fn, stride in (("od", L.bvhgpu_traverse_od_f32x3, 6), ("full", L.bvhgpu_traverse_f32x3, 9))[:1 if QUICK else 2]:
    h_r = ctx.host_alloc(N * stride * 4, np.float32)
    full = rays.view(np.float32).reshape(-1, 9)
    h_r.reshape(-1, stride)[:] = full[:, :stride]
    h_off = ctx.host_alloc(4 * (N + 1), np.uint32); h_hits = ctx.host_alloc(4 * 8 * N, np.uint32)
    tot = C.c_size_t(0)
    call = lambda: capi.check(fn(bvh._h, 0, h_r.ctypes.data_as(C.c_void_p), N, h_off.ctypes.data_as(C.c_void_p), h_hits.ctypes.data_as(C.c_void_p), 8 * N, C.byref(tot)))
    for stream_opt in ((-1,) if QUICK else (-1, 0)):
        ctx.set_option("traverse_stream", stream_opt)
        for _ in range(5): call()
        ts = []
        for _ in range(200):
            t0 = time.perf_counter(); call(); ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        ctx.set_option("profile", 1)
        ph = []
        for _ in range(20):
            call()
            ph.append([ctx.get_metric(f"e2e_host_us_{k}") for k in range(6)] + [ctx.get_metric(m) * 1e3 for m in ("e2e_h2d_ms", "e2e_walk_ms", "e2e_emit_ms", "e2e_d2h_ms")])
        ctx.set_option("profile", 0)
        med = np.median(np.array(ph), axis=0).round(1).tolist()
        out[f"{name}_stream{stream_opt}"] = {"ms_p10_p50_p90_max": [round(ts[20], 3), round(ts[100], 3), round(ts[180], 3), round(ts[-1], 3)], "streamed": ctx.get_metric("host_streamed"), "stream_write_value": ctx.get_metric("stream_write_value"),
                                             "host_us[alloc, copies enq, walk launched, all enq, st drained, done]": med[:6], "device_us_since_start[h2d done, walk done, emit done, d2h done]": med[6:]}
    ctx.set_option("traverse_stream", -1)
if QUICK:
    v = out["od_stream-1"]
    print("chunks", os.environ.get("BVHGPU_CHUNKS"), "sched", os.environ.get("BVHGPU_CHUNK_SCHEDULE"), "ms p10/p50/p90/max", v["ms_p10_p50_p90_max"],
          "device us [h2d, walk, emit, d2h]", v["device_us_since_start[h2d done, walk done, emit done, d2h done]"])
    sys.exit(0)
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/e2e_probe.json", "w"), indent=1)
