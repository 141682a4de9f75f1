#!/usr/bin/env python
"""Debug aid: run one exact-SAH build with BVHGPU_TRACE and print a timeline of the persistent kernel's tasks."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n_cubes = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
path = "/tmp/bvh_trace.bin"
from bvh_b200 import api, scenes
aabbs = scenes.create_n_cubes_aabbs(n_cubes)
api.Bvh.build(aabbs).free()              # warm-up (module load, pool)
os.environ["BVHGPU_TRACE"] = path
b = api.Bvh.build(aabbs)
del os.environ["BVHGPU_TRACE"]
t = np.fromfile(path, dtype=np.uint32).reshape(-1, 4)
t = t[(t[:, 3] != 0)]
kind, cnt = t[:, 0] >> 28, t[:, 0] & 0x0FFFFFFF
t0, t1 = t[:, 2] / 1e3, t[:, 3] / 1e3
print(f"tasks {len(t)}  end {t1.max():.1f} us")
for k, name in enumerate(["SEG", "BIN", "SCATTER"]):
    m = kind == k
    if m.any():
        d = t1[m] - t0[m]
        print(f"{name:8s} n={m.sum():7d} dur mean {d.mean():7.2f} med {np.median(d):7.2f} max {d.max():7.2f} us | first start {t0[m].min():7.1f} last end {t1[m].max():7.1f}")
# big segments, keyed by their shape count (one generation each): BIN / SCATTER phase windows
segs = {}
for k, c, a, b_ in zip(kind, t[:, 1], t0, t1):
    if k in (1, 2):
        e = segs.setdefault((int(c), int(k)), [a, b_, 0, 0.0]); e[0] = min(e[0], a); e[1] = max(e[1], b_); e[2] += 1; e[3] = max(e[3], b_ - a)
print("count   tiles  BIN first-start .. last-end (longest tile) | SCATTER first-start .. last-end (longest)")
for c in sorted(set(c for c, _ in segs), reverse=True)[:28]:
    bi, sc = segs.get((c, 1)), segs.get((c, 2))
    line = f"{c:7d} {bi[2]:5d}  BIN {bi[0]:7.1f} .. {bi[1]:7.1f} ({bi[3]:5.1f})"
    if sc: line += f" | SCAT {sc[0]:7.1f} .. {sc[1]:7.1f} ({sc[3]:5.1f})"
    print(line)
m = kind == 0
print("SEG tasks by size: ", end="")
for lo, hi in [(2, 8), (8, 32), (32, 64), (64, 128), (128, 257)]:
    mm = m & (cnt >= lo) & (cnt < hi)
    if mm.any():
        print(f"[{lo},{hi}) n={mm.sum()} dur {np.mean(t1[mm]-t0[mm]):.1f}us  ", end="")
print()
