#!/usr/bin/env python
"""Exact-SAH build time vs the builder's options (build_small, build_gang) at several sizes; CUDA events, median of 5."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bvh_b200 import api, capi, scenes
dev = torch.device("cuda", 0)
ctx = api.Context(0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def t_build(d, n, prec, mode):
    ms = []
    for k in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); bv = api.Bvh.build_dev(d.data_ptr(), n, prec=prec, ctx=ctx, mode=mode); b.record(stream); torch.cuda.synchronize(dev)
        if k >= 2: ms.append(a.elapsed_time(b))
        bv.free()
    return sorted(ms)[len(ms) // 2]
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1000, 10000, 100000]
precs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["f32"]
for prec in precs:
    for nc in sizes:
        a = scenes.create_n_cubes_aabbs(nc, prec=prec) if prec != "f32" else scenes.create_n_cubes_aabbs(nc)
        d = torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
        for small, sub, gang in ((-1, -1, -1), (0, 0, -1), (0, 1, -1), (1, 0, -1), (1, 1, -1), (0, 1, 0), (0, 1, 1)):
            ctx.set_option("build_small", small); ctx.set_option("build_subtree", sub); ctx.set_option("build_gang", gang)
            e = t_build(d, len(a), prec, capi.BUILD_EXACT_SAH)
            t = t_build(d, len(a), prec, capi.BUILD_LBVH_TREELET)
            print(f"{prec} n={len(a):9d} small={small:2d} subtree={sub:2d} gang={gang:2d}: exact {e:8.3f} ms  treelet {t:8.3f} ms", flush=True)
        del d
