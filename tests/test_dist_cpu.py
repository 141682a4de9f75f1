"""world_size-2 gloo test of the N>1 path's host logic (bvh_b200/dist.py): shard the ray batch,
all-gather the per-rank CSR hit lists, compare with the single-process result.  The per-rank
traversal result is produced by the CPU oracle here (no GPU in this container)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bvh_b200.dist import allgather_csr, shard_range
        from oracle import oracle as O

        shapes = O.create_n_cubes(50)
        res = O.build(shapes)
        rays, _ = O.create_rays(1001)
        lo, hi = shard_range(len(rays), rank, world)
        local = O.traverse(res.nodes, shapes, rays[lo:hi], O.MODE_RECURSIVE)
        off = torch.from_numpy(local.offsets.astype(np.int64))
        hits = torch.from_numpy(np.concatenate([local.hits, np.zeros(7, np.uint32)]).astype(np.int64))   # over-allocated buffer
        g_off, g_hits = allgather_csr(off, hits, len(local.hits) if rank == 0 else None)   # rank 1 lets it read offsets[-1]
        full = O.traverse(res.nodes, shapes, rays, O.MODE_RECURSIVE)
        ok = np.array_equal(g_off.numpy().astype(np.uint64), full.offsets) and np.array_equal(g_hits.numpy().astype(np.uint32), full.hits)
        q.put((rank, bool(ok), lo, hi))
    finally:
        dist.destroy_process_group()


def test_allgather_csr_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    spans = sorted((lo, hi) for _, _, lo, hi in res)
    assert spans == [(0, 501), (501, 1001)]


def test_shard_range_covers_everything():
    from bvh_b200.dist import shard_range

    for n in (0, 1, 7, 16, 1001):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
