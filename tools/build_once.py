#!/usr/bin/env python
"""Profiling helper: two exact-SAH builds of the bench scene (the second is the one to capture with `ncu -s 1 -c 1`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bvh_b200 import api, scenes
n_cubes = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
aabbs = scenes.create_n_cubes_aabbs(n_cubes)
for _ in range(2):
    api.Bvh.build(aabbs).free()
print("ok")
