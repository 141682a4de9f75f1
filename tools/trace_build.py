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
m = kind == 4
if m.any():
    print("gang levels (warp 0 of each segment): count tiles start | bin | wait1 | split+scatter | wait2 | = dur")
    i4 = np.nonzero(m)[0]
    order = i4[np.argsort(t0[i4])]
    for i in order[:40]:
        ph = t[i - 1]                      # the kind-5 companion entry precedes it in the buffer
        a, e = t0[i], t1[i]
        b1, w1, s2 = ph[1] / 1e3, ph[2] / 1e3, ph[3] / 1e3
        print(f"   {t[i,1]:8d} {cnt[i]:5d} {a:8.1f} | {b1-a:5.1f} | {w1-b1:5.1f} | {s2-w1:5.1f} | {e-s2:5.1f} | {e-a:6.1f}")
m = kind == 3
if m.any():
    print(f"GANG memberships n={m.sum()} first start {t0[m].min():.1f} last end {t1[m].max():.1f}")
# big segments, keyed by their shape count (one generation each): BIN / SCATTER phase windows
segs = {}
for k, c, a, b_ in zip(kind, t[:, 1], t0, t1):
    if k in (1, 2):
        e = segs.setdefault((int(c), int(k)), [a, b_, 0, 0.0]); e[0] = min(e[0], a); e[1] = max(e[1], b_); e[2] += 1; e[3] = max(e[3], b_ - a)
print("count   tiles  BIN first-start .. last-end (longest tile) | SCATTER first-start .. last-end (longest)")
for c in sorted(set(c for c, _ in segs), reverse=True)[:28]:
    bi, sc = segs.get((c, 1)), segs.get((c, 2))
    if bi is None: continue
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

# concurrency over time: tasks in flight (= busy warps) per 20 us bucket
mt = (kind <= 3)
edges = np.arange(0, t1[mt].max() + 20, 20.0)
busy = np.zeros(len(edges))
for a, b_ in zip(t0[mt], t1[mt]):
    i0, i1 = int(a // 20), int(b_ // 20)
    for i in range(i0, i1 + 1):
        lo, hi = max(a, edges[i]), min(b_, edges[i] + 20)
        busy[i] += max(0.0, hi - lo) / 20.0
print("busy warps per 20us:", " ".join(f"{int(x)}" for x in busy))
ms = kind == 0
starts = np.histogram(t0[ms], bins=edges)[0]
print("SEG starts per 20us:", " ".join(str(x) for x in starts))
