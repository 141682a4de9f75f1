#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_sharded.py : fused P2P sharded traversal == NCCL all-gather path == oracle."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
from bvh_b200 import api, capi, scenes
from bvh_b200.dist import ShardedTraversal, allgather_csr, shard_range
from bvh_b200.dtypes import RAY3F
from oracle import oracle as O
ctx = api.Context(local)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
N = 200_003
shapes = scenes.create_n_cubes_aabbs(2000)
lo, hi = shard_range(N, rank, world)
o, d = scenes.ray_endpoints(hi - lo, first_ray=lo)
rays = api.Ray.new(o, d, ctx=ctx)
d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(dev)
bvh = api.Bvh.build(shapes, ctx=ctx)
cap = 4 * N
sh = ShardedTraversal(bvh, hi - lo, cap)
for it in range(3):
    sh.step(d_rays.data_ptr(), hi - lo)
off, hits = sh.fetch()
# NCCL path
d_off = torch.empty(hi - lo + 1, dtype=torch.int32, device=dev); d_hits = torch.empty(cap, dtype=torch.int32, device=dev)
bvh.traverse_dev(d_rays.data_ptr(), hi - lo, d_off.data_ptr(), d_hits.data_ptr(), cap)
g_off, g_hits = allgather_csr(d_off, d_hits)
ok_nccl = np.array_equal(g_off.cpu().numpy().astype(np.uint32), off) and np.array_equal(g_hits.cpu().numpy().view(np.uint32), hits)
# oracle on the whole batch
oo, dd = scenes.ray_endpoints(N, 0)
allrays = O.ray_new(oo, dd)
res = O.build(shapes)
r = O.traverse(res.nodes, shapes, allrays, O.MODE_RECURSIVE, threads=8)
ok_oracle = np.array_equal(off.astype(np.uint64), r.offsets) and np.array_equal(hits, r.hits)
print(f"rank {rank}/{world}: fused==nccl {ok_nccl}  fused==oracle {ok_oracle}  total hits {len(hits)}", flush=True)
sh.close()
dist.destroy_process_group()
sys.exit(0 if (ok_nccl and ok_oracle) else 1)
