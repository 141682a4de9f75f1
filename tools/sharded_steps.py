#!/usr/bin/env python
"""K sharded steps of configs[1] on however many ranks torchrun starts (also 1): for launch lists (ncu) and step timing.
usage: torchrun --nproc-per-node N tools/sharded_steps.py [steps] [rays_per_rank] [virtual_world]
virtual_world > world (single process only): this rank pretends to be rank 0 of a larger job whose other shards are empty --
not runnable; use it only to size kernels (gscan over N x rays)."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29655")
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev, rank=rank, world_size=world)
from bvh_b200 import api, capi, scenes
from bvh_b200.dist import ShardedTraversal
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
ctx = api.Context(local)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
bvh = api.Bvh.build(scenes.create_n_cubes_aabbs(10_000), ctx=ctx)
o, d = scenes.ray_endpoints(n, first_ray=rank * n)
rays = api.Ray.new(o, d, ctx=ctx)
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(dev)
sh = ShardedTraversal(bvh, n, 2 * n * world)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = []
for k in range(steps + 3):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); sh.step(d_rays.data_ptr(), n); b.record(stream)
    ev.append((a, b))
torch.cuda.synchronize(dev)
ts = sorted(a.elapsed_time(b) for a, b in ev[3:])
print(f"rank {rank}/{world}: step ms median {ts[len(ts)//2]:.4f} min {ts[0]:.4f} max {ts[-1]:.4f}", flush=True)
sh.close()
dist.destroy_process_group()
