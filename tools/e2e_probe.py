import os, sys, time, numpy as np, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bvh_b200 import api, capi, scenes
from bvh_b200.dtypes import RAY3F
dev = torch.device("cuda", 0)
ctx = api.Context(0)
N = 1_000_000
aabbs = scenes.create_n_cubes_aabbs(10000)
bvh = api.Bvh.build(aabbs, ctx=ctx)
o, d = scenes.ray_endpoints(N)
rays = api.Ray.new(o, d, ctx=ctx)
h_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1).copy()).pin_memory()
h_off = torch.empty(N + 1, dtype=torch.int32).pin_memory()
h_hits = torch.empty(8 * N, dtype=torch.int32).pin_memory()
tot = C.c_size_t(0)
fn = capi.lib().bvhgpu_traverse_f32x3
ctx.set_option("profile", 1)
# plain H2D bandwidth
dbuf = torch.empty(N * 36, dtype=torch.uint8, device=dev)
for _ in range(3): dbuf.copy_(h_rays, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): dbuf.copy_(h_rays, non_blocking=True)
torch.cuda.synchronize(); print("plain H2D 36MB ms", (time.perf_counter() - t0) / 10 * 1e3)
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    capi.check(fn(bvh._h, 0, h_rays.data_ptr(), N, h_off.data_ptr(), h_hits.data_ptr(), 8 * N, C.byref(tot)))
    dt = (time.perf_counter() - t0) * 1e3
    print(f"call {dt:.3f} ms  walk_end {ctx.get_metric('e2e_walk_ms'):.3f}  h2d_end {ctx.get_metric('e2e_h2d_ms'):.3f}  emit_end {ctx.get_metric('e2e_emit_ms'):.3f}  d2h_end {ctx.get_metric('e2e_d2h_ms'):.3f}")
