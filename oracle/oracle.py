"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes/numpy front-end over oracle/build/liboracle.so (the C++ CPU restatement of
the reference's build / flatten / traverse path, see oracle/bvh_oracle.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing under bvh_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "build", "liboracle.so")

U32_MAX = 0xFFFFFFFF


def _dtypes(f):
    aabb = np.dtype([("min", f, (3,)), ("max", f, (3,))])
    ray = np.dtype([("origin", f, (3,)), ("direction", f, (3,)), ("inv_direction", f, (3,))])
    node = np.dtype(
        [("parent", "<u4"), ("child_l", "<u4"), ("child_r", "<u4"), ("shape", "<u4"), ("l_aabb", aabb), ("r_aabb", aabb)]
    )
    flat = np.dtype(
        {
            "names": ["aabb", "entry_index", "exit_index", "shape_index"],
            "formats": [aabb, "<u4", "<u4", "<u4"],
            "offsets": [0, aabb.itemsize, aabb.itemsize + 4, aabb.itemsize + 8],
            "itemsize": 36 if f == "<f4" else 64,
        }
    )
    return aabb, ray, node, flat


AABB3F, RAY3F, NODE3F, FLAT3F = _dtypes("<f4")
AABB3D, RAY3D, NODE3D, FLAT3D = _dtypes("<f8")

_DT = {
    "f32": dict(f=np.float32, aabb=AABB3F, ray=RAY3F, node=NODE3F, flat=FLAT3F),
    "f64": dict(f=np.float64, aabb=AABB3D, ray=RAY3D, node=NODE3D, flat=FLAT3D),
}


def build_library(force: bool = False) -> str:
    """Compile oracle/build/liboracle.so with the committed Makefile (gcc only)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, n)) for n in ("oracle_capi.cpp", "bvh_oracle.hpp", "Makefile"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < src_m:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_library()
        _lib = C.CDLL(_SO)
        _lib.orc_flatten_f32.restype = C.c_uint64
        _lib.orc_flatten_f64.restype = C.c_uint64
        _lib.orc_traverse_batch_f32.restype = C.c_uint64
        _lib.orc_traverse_batch_f64.restype = C.c_uint64
        _lib.orc_splitmix64.restype = C.c_uint64
        _lib.orc_query_batch_f32.restype = C.c_uint64
        _lib.orc_query_batch_f64.restype = C.c_uint64
        _lib.orc_last_build_ns.restype = C.c_uint64
        _lib.orc_hardware_threads.restype = C.c_uint32
        _lib.orc_sizeof.restype = C.c_uint32
        _lib.orc_update_shapes_f32.restype = C.c_uint32
        _lib.orc_update_shapes_f64.restype = C.c_uint32
        # struct layout must agree with the numpy dtypes
        sizes = [_lib.orc_sizeof(i) for i in range(8)]
        want = [d.itemsize for d in (AABB3F, RAY3F, NODE3F, FLAT3F, AABB3D, RAY3D, NODE3D, FLAT3D)]
        assert sizes == want, (sizes, want)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def hardware_threads() -> int:
    return int(lib().orc_hardware_threads())


@dataclass
class BuildResult:
    nodes: np.ndarray          # NODE3F / NODE3D, 2n-1 entries
    node_index: np.ndarray     # u32[n], leaf node index of every shape (BHShape::set_bh_node_index)
    prim_visits: int
    degenerate_splits: int
    max_depth: int
    nosplit_fallthrough: int
    seconds: float = 0.0       # wall time inside the C++ build (no numpy allocation)


def build(aabbs: np.ndarray, prec: str = "f32", threads: int = 1) -> BuildResult:
    """Bvh::build (threads == 1) / Bvh::build_par analogue (threads > 1)."""
    d = _DT[prec]
    aabbs = np.ascontiguousarray(aabbs, dtype=d["aabb"])
    n = len(aabbs)
    nodes = np.zeros(max(2 * n - 1, 0), dtype=d["node"])
    node_index = np.zeros(n, dtype=np.uint32)
    stats = np.zeros(4, dtype=np.uint64)
    if threads <= 1:
        getattr(lib(), f"orc_build_{prec}")(_p(aabbs), C.c_uint32(n), _p(nodes), _p(node_index), _p(stats))
    else:
        getattr(lib(), f"orc_build_par_{prec}")(_p(aabbs), C.c_uint32(n), _p(nodes), _p(node_index), _p(stats), C.c_uint32(threads))
    return BuildResult(nodes, node_index, int(stats[0]), int(stats[1]), int(stats[2]), int(stats[3]), lib().orc_last_build_ns() * 1e-9)


def flatten(nodes: np.ndarray, prec: str = "f32") -> np.ndarray:
    """Bvh::flatten (literal recursion order)."""
    d = _DT[prec]
    nodes = np.ascontiguousarray(nodes, dtype=d["node"])
    n_nodes = len(nodes)
    n = (n_nodes + 1) // 2
    cap = max(3 * n - 2, 1) if n_nodes else 0
    out = np.zeros(cap, dtype=d["flat"])
    ln = getattr(lib(), f"orc_flatten_{prec}")(_p(nodes), C.c_uint32(n_nodes), _p(out), C.c_uint64(cap))
    assert ln <= cap
    return out[:ln]


MODE_RECURSIVE, MODE_FLAT, MODE_ITERATOR = 0, 1, 2


@dataclass
class TraverseResult:
    offsets: np.ndarray    # u64[nrays+1]
    hits: np.ndarray       # u32[total], reference DFS order within each ray
    node_visits: int
    slab_tests: int
    leaf_visits: int
    iter_overflow: bool
    seconds: float = 0.0       # wall time of the traversal section inside C++ (thread create .. join)


def traverse(tree: np.ndarray, shapes: np.ndarray, rays: np.ndarray, mode: int = MODE_FLAT, prec: str = "f32",
             threads: int = 1, count_stats: bool = True) -> TraverseResult:
    """threads > 1: persistent pinned worker pool, rays handed out in chunks (dynamic).  count_stats=False keeps the visit
    counters out of the hot loop (the timed CPU legs of bench.py); `.seconds` is the wall time of the parallel section."""
    d = _DT[prec]
    tdt = d["flat"] if mode == MODE_FLAT else d["node"]
    tree = np.ascontiguousarray(tree, dtype=tdt)
    shapes = np.ascontiguousarray(shapes, dtype=d["aabb"])
    rays = np.ascontiguousarray(rays, dtype=d["ray"])
    nrays = len(rays)
    offsets = np.zeros(nrays + 1, dtype=np.uint64)
    stats = np.zeros(6, dtype=np.uint64)
    ovf = C.c_int(0)
    cap = max(4 * nrays, 1024)
    fn = getattr(lib(), f"orc_traverse_batch_{prec}")
    while True:
        hits = np.empty(cap, dtype=np.uint32)
        stats[5] = 0 if count_stats else 1
        total = fn(C.c_int(mode), _p(tree), C.c_uint32(len(tree)), _p(shapes), _p(rays), C.c_uint64(nrays),
                   _p(offsets), _p(hits), C.c_uint64(cap), _p(stats), C.c_uint32(threads), C.byref(ovf))
        if total <= cap:
            break
        cap = int(total)
    return TraverseResult(offsets, hits[:total].copy(), int(stats[0]), int(stats[1]), int(stats[2]), bool(ovf.value), int(stats[4]) * 1e-9)


QUERY_AABB, QUERY_POINT, QUERY_BALL = 1, 2, 3
_QSTRIDE = {1: 6, 2: 3, 3: 4}


def query(kind: int, queries, nodes, shapes, flat=None, prec="f32"):
    """Bvh::traverse (flat=None) / FlatBvh::traverse for Aabb / Point / Ball queries (src/aabb/intersection.rs:35-45,
    src/ball.rs:85-106).  queries: (n, 6) aabb min+max | (n, 3) points | (n, 4) centre+radius.  Returns (offsets u64, hits u32)."""
    d = _DT[prec]
    q = np.ascontiguousarray(queries, dtype=d["f"]).reshape(-1, _QSTRIDE[kind])
    nodes = np.ascontiguousarray(nodes, dtype=d["node"])
    shapes = np.ascontiguousarray(shapes, dtype=d["aabb"])
    fl = np.ascontiguousarray(flat, dtype=d["flat"]) if flat is not None else np.zeros(0, dtype=d["flat"])
    offsets = np.zeros(len(q) + 1, dtype=np.uint64)
    cap = max(64 * len(q), 1024)
    while True:
        hits = np.empty(cap, dtype=np.uint32)
        total = getattr(lib(), f"orc_query_batch_{prec}")(C.c_int(kind), C.c_int(1 if flat is not None else 0), _p(nodes), C.c_uint32(len(nodes)),
                                                         _p(fl), C.c_uint32(len(fl)), _p(shapes), _p(q), C.c_uint64(len(q)), _p(offsets), _p(hits), C.c_uint64(cap))
        if total <= cap:
            return offsets, hits[:total].copy()
        cap = int(total)


def is_consistent(nodes, shapes, prec="f32") -> bool:
    d = _DT[prec]
    nodes = np.ascontiguousarray(nodes, dtype=d["node"])
    shapes = np.ascontiguousarray(shapes, dtype=d["aabb"])
    return bool(getattr(lib(), f"orc_is_consistent_{prec}")(_p(nodes), C.c_uint32(len(nodes)), _p(shapes)))


def is_tight(nodes, prec="f32") -> bool:
    d = _DT[prec]
    nodes = np.ascontiguousarray(nodes, dtype=d["node"])
    return bool(getattr(lib(), f"orc_is_tight_{prec}")(_p(nodes), C.c_uint32(len(nodes))))


def sah_cost(nodes, prec="f32"):
    d = _DT[prec]
    nodes = np.ascontiguousarray(nodes, dtype=d["node"])
    out = np.zeros(2, dtype=np.float64)
    getattr(lib(), f"orc_sah_cost_{prec}")(_p(nodes), C.c_uint32(len(nodes)), _p(out))
    return float(out[0]), float(out[1])


DIST_AABB, DIST_TRIANGLE = 0, 1


def nearest_to(tree, shapes, points, prec="f32", flat=False, kind=DIST_AABB, tris=None):
    """Bvh::nearest_to (bvh_impl.rs:221-238) / FlatBvh::nearest_to (flat_bvh.rs:513-562) for a batch of points.
    Returns (shape index per point, 0xFFFFFFFF for an empty tree; distance per point)."""
    d = _DT[prec]
    tree = np.ascontiguousarray(tree, dtype=d["flat"] if flat else d["node"])
    shapes = np.ascontiguousarray(shapes, dtype=d["aabb"])
    pts = np.ascontiguousarray(points, dtype=d["f"]).reshape(-1, 3)
    tr = None if tris is None else np.ascontiguousarray(tris, dtype=d["f"]).reshape(-1, 9)
    out_s = np.zeros(len(pts), dtype=np.uint32)
    out_d = np.zeros(len(pts), dtype=d["f"])
    getattr(lib(), f"orc_nearest_batch_{prec}")(C.c_int(1 if flat else 0), C.c_int(kind), _p(tree), C.c_uint32(len(tree)), _p(shapes), _p(tr),
                                                _p(pts), C.c_uint64(len(pts)), _p(out_s), _p(out_d))
    return out_s, out_d


def shape_distances_squared(shapes, point, prec="f32", kind=DIST_AABB, tris=None):
    """PointDistance::distance_squared of every shape to one point (brute-force side of nearest_to_and_verify, testbase.rs:290-312)."""
    d = _DT[prec]
    shapes = np.ascontiguousarray(shapes, dtype=d["aabb"])
    tr = None if tris is None else np.ascontiguousarray(tris, dtype=d["f"]).reshape(-1, 9)
    pt = np.ascontiguousarray(point, dtype=d["f"]).reshape(3)
    out = np.zeros(len(shapes), dtype=d["f"])
    getattr(lib(), f"orc_shape_distance_{prec}")(C.c_int(kind), _p(shapes), _p(tr), C.c_uint32(len(shapes)), _p(pt), _p(out))
    return out


def update_shapes(nodes, node_index, shapes, changed, prec="f32"):
    """Bvh::update_shapes (src/bvh/optimization.rs:290-302): remove then re-insert `changed` (in this order) given the shapes'
    CURRENT AABBs.  Returns new (nodes, node_index); the node array is no longer in build's preorder layout."""
    d = _DT[prec]
    nodes = np.array(nodes, dtype=d["node"], copy=True)
    node_index = np.array(node_index, dtype=np.uint32, copy=True)
    shapes = np.ascontiguousarray(shapes, dtype=d["aabb"])
    changed = np.ascontiguousarray(changed, dtype=np.uint32)
    n = getattr(lib(), f"orc_update_shapes_{prec}")(_p(nodes), C.c_uint32(len(nodes)), _p(node_index), C.c_uint32(len(shapes)),
                                                   _p(shapes), _p(changed), C.c_uint32(len(changed)))
    if n == 0xFFFFFFFF:
        raise RuntimeError("update_shapes: the reference would have panicked on this input")
    return nodes[:n], node_index


def connect_nodes(nodes, shapes, child, parent, left_child, prec="f32"):
    """Bvh::connect_nodes (src/bvh/optimization.rs:34-65) on a copy of `nodes`."""
    d = _DT[prec]
    nodes = np.array(nodes, dtype=d["node"], copy=True)
    shapes = np.ascontiguousarray(shapes, dtype=d["aabb"])
    ok = getattr(lib(), f"orc_connect_nodes_{prec}")(_p(nodes), C.c_uint32(len(nodes)), _p(shapes), C.c_uint32(child), C.c_uint32(parent),
                                                   C.c_int(1 if left_child else 0))
    assert ok
    return nodes


# ---------------------------------------------------------------- fixtures (src/testbase.rs)
def default_bounds(prec="f32") -> np.ndarray:
    b = np.zeros(1, dtype=_DT[prec]["aabb"])
    b["min"] = -100000.0
    b["max"] = 100000.0
    return b


def create_n_cubes(n_cubes: int, bounds=None, prec="f32", want_tris: bool = False):
    """create_n_cubes (testbase.rs:608-615): returns the 12*n_cubes triangle AABBs (and vertices)."""
    d = _DT[prec]
    bounds = default_bounds(prec) if bounds is None else np.ascontiguousarray(bounds, dtype=d["aabb"])
    aabbs = np.zeros(12 * n_cubes, dtype=d["aabb"])
    tris = np.zeros((12 * n_cubes, 3, 3), dtype=d["f"]) if want_tris else None
    getattr(lib(), f"orc_create_n_cubes_{prec}")(C.c_uint32(n_cubes), _p(bounds), _p(tris), _p(aabbs))
    return (aabbs, tris) if want_tris else aabbs


def tri_aabbs(tris: np.ndarray, prec="f32") -> np.ndarray:
    d = _DT[prec]
    tris = np.ascontiguousarray(tris, dtype=d["f"]).reshape(-1, 9)
    out = np.zeros(len(tris), dtype=d["aabb"])
    getattr(lib(), f"orc_tri_aabbs_{prec}")(_p(tris), C.c_uint64(len(tris)), _p(out))
    return out


def create_rays(n: int, bounds=None, seed: int = 0, prec="f32"):
    """n x create_ray (testbase.rs:687-691) chained from `seed`; returns (rays, next_seed)."""
    d = _DT[prec]
    bounds = default_bounds(prec) if bounds is None else np.ascontiguousarray(bounds, dtype=d["aabb"])
    rays = np.zeros(n, dtype=d["ray"])
    s = C.c_uint64(seed)
    getattr(lib(), f"orc_create_rays_{prec}")(C.byref(s), _p(bounds), C.c_uint64(n), _p(rays))
    return rays, int(s.value)


def next_points(n: int, bounds=None, seed: int = 0, prec="f32"):
    d = _DT[prec]
    bounds = default_bounds(prec) if bounds is None else np.ascontiguousarray(bounds, dtype=d["aabb"])
    out = np.zeros((n, 3), dtype=d["f"])
    s = C.c_uint64(seed)
    getattr(lib(), f"orc_next_points_{prec}")(C.byref(s), _p(bounds), C.c_uint64(n), _p(out))
    return out, int(s.value)


def ray_new(origins, dirs, prec="f32") -> np.ndarray:
    """Ray::new (ray_impl.rs:70-80) for arrays of origins / directions."""
    d = _DT[prec]
    origins = np.ascontiguousarray(origins, dtype=d["f"]).reshape(-1, 3)
    dirs = np.ascontiguousarray(dirs, dtype=d["f"]).reshape(-1, 3)
    out = np.zeros(len(origins), dtype=d["ray"])
    getattr(lib(), f"orc_ray_new_{prec}")(_p(origins), _p(dirs), C.c_uint64(len(origins)), _p(out))
    return out


def ray_intersects_aabb(ray: np.ndarray, aabb: np.ndarray, prec="f32") -> bool:
    d = _DT[prec]
    ray = np.ascontiguousarray(ray, dtype=d["ray"]).reshape(1)
    aabb = np.ascontiguousarray(aabb, dtype=d["aabb"]).reshape(1)
    return bool(getattr(lib(), f"orc_ray_intersects_aabb_{prec}")(_p(ray), _p(aabb)))


def ray_slice(ray, aabb, prec="f32"):
    """Ray::intersection_slice_for_aabb (ray_impl.rs:118-145): (tmin, tmax) or None."""
    d = _DT[prec]
    ray = np.ascontiguousarray(ray, dtype=d["ray"]).reshape(1)
    aabb = np.ascontiguousarray(aabb, dtype=d["aabb"]).reshape(1)
    out = np.zeros(2, dtype=d["f"])
    ok = getattr(lib(), f"orc_ray_slice_{prec}")(_p(ray), _p(aabb), _p(out))
    return (out[0], out[1]) if ok else None


def ray_triangle(ray, tri9, prec="f32"):
    """Ray::intersects_triangle (ray_impl.rs:154-213): (distance, u, v); distance = inf for a miss / back face."""
    d = _DT[prec]
    ray = np.ascontiguousarray(ray, dtype=d["ray"]).reshape(1)
    tri = np.ascontiguousarray(tri9, dtype=d["f"]).reshape(9)
    uv = np.zeros(2, dtype=d["f"])
    fn = getattr(lib(), f"orc_ray_triangle_{prec}")
    fn.restype = C.c_float if prec == "f32" else C.c_double
    dist = fn(_p(ray), _p(tri), _p(uv))
    return d["f"](dist), uv[0], uv[1]


CLOSEST_AABB, CLOSEST_TRIANGLE = 0, 1


def closest_hit(nodes, shapes, rays, tris=None, prec="f32"):
    """What a caller of the reference computes per ray: candidates = Bvh::traverse; tris is None: the shape whose AABB is entered first
    (key: entry distance, then DFS order); else Ray::intersects_triangle on every candidate, minimum distance (ties: lower index).
    Returns (shape u32 [U32_MAX = none], distance [inf = none], uv (n, 2))."""
    d = _DT[prec]
    nodes = np.ascontiguousarray(nodes, dtype=d["node"])
    shapes = np.ascontiguousarray(shapes, dtype=d["aabb"])
    rays = np.ascontiguousarray(rays, dtype=d["ray"])
    tr = None if tris is None else np.ascontiguousarray(tris, dtype=d["f"]).reshape(-1, 9)
    out_s = np.zeros(len(rays), dtype=np.uint32)
    out_d = np.zeros(len(rays), dtype=d["f"])
    out_uv = np.zeros((len(rays), 2), dtype=d["f"])
    getattr(lib(), f"orc_closest_hit_batch_{prec}")(C.c_int(CLOSEST_AABB if tris is None else CLOSEST_TRIANGLE), _p(nodes), C.c_uint32(len(nodes)), _p(shapes), _p(tr),
                                                    _p(rays), C.c_uint64(len(rays)), _p(out_s), _p(out_d), _p(out_uv))
    return out_s, out_d, out_uv


def aligned_boxes(prec="f32") -> np.ndarray:
    out = np.zeros(21, dtype=_DT[prec]["aabb"])
    getattr(lib(), f"orc_aligned_boxes_{prec}")(_p(out))
    return out


def aabb_ops(aabb, prec="f32"):
    d = _DT[prec]
    aabb = np.ascontiguousarray(aabb, dtype=d["aabb"]).reshape(1)
    c = np.zeros(3, dtype=d["f"])
    sa = np.zeros(1, dtype=d["f"])
    ax = C.c_int(0)
    getattr(lib(), f"orc_aabb_ops_{prec}")(_p(aabb), _p(c), _p(sa), C.byref(ax))
    return c, sa[0], int(ax.value)


def make_aabbs(mins, maxs, prec="f32") -> np.ndarray:
    d = _DT[prec]
    mins = np.asarray(mins, dtype=d["f"]).reshape(-1, 3)
    maxs = np.asarray(maxs, dtype=d["f"]).reshape(-1, 3)
    out = np.zeros(len(mins), dtype=d["aabb"])
    out["min"] = mins
    out["max"] = maxs
    return out


def unit_boxes(centers, prec="f32") -> np.ndarray:
    """UnitBox::aabb (testbase.rs:84-90): pos + (-0.5) .. pos + 0.5 in T."""
    d = _DT[prec]
    c = np.asarray(centers, dtype=d["f"]).reshape(-1, 3)
    return make_aabbs(c + d["f"](-0.5), c + d["f"](0.5), prec)


def per_ray_lists(offsets, hits):
    return [hits[int(offsets[i]):int(offsets[i + 1])] for i in range(len(offsets) - 1)]
