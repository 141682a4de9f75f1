#!/usr/bin/env python
"""Generates tests/golden/sponza_tris.npz from the reference's media/sponza.obj with the reference loader's
fan triangulation (src/testbase.rs:445-487: positions only, polygon (a,b,c,d,..) -> (a,b,c),(a,c,d),..).
Run in the container that has /root/reference; the .npz (float32 vertices, lossless) is committed because
/root/reference does not exist on the GPU box.  f32 parsing: Rust's `str::parse::<f32>` and numpy's float32
conversion are both correctly rounded, so the vertices are bit-identical to what the reference loads."""
import os, sys
import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/media/sponza.obj"
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "sponza_tris.npz")
verts, tris = [], []
for line in open(src):
    if line.startswith("v "):
        p = line.split()
        verts.append((np.float32(p[1]), np.float32(p[2]), np.float32(p[3])))
    elif line.startswith("f "):
        idx = []
        for tok in line.split()[1:]:
            i = int(tok.split("/")[0])
            idx.append(i - 1 if i > 0 else len(verts) + i)
        for k in range(1, len(idx) - 1):                 # anchor, second, third (testbase.rs:461-470)
            tris.append((idx[0], idx[k], idx[k + 1]))
verts = np.array(verts, dtype=np.float32)
tris = np.array(tris, dtype=np.uint32)
np.savez_compressed(dst, vertices=verts, triangles=tris)
print(len(verts), "vertices", len(tris), "triangles ->", dst, os.path.getsize(dst), "bytes")
