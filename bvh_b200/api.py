"""Host-side mirror of the reference crate's interface for the hot path, on top of the C ABI.

Names follow the reference (svenstaro/bvh 0.12.0):

    Bvh.build(shapes) / Bvh.build_par(shapes)     src/bvh/bvh_impl.rs:40-96, src/bounding_hierarchy.rs:158-177
    bvh.nodes, shape node indices                  src/bvh/bvh_impl.rs:27-33, src/bounding_hierarchy.rs:53-65
    bvh.flatten() -> FlatBvh                       src/flat_bvh.rs:312-319
    bvh.traverse(ray, shapes)                      src/bvh/bvh_impl.rs:104-119
    bvh.traverse_iterator(ray, shapes)             src/bvh/bvh_impl.rs:128-134
    flat_bvh.traverse(ray, shapes)                 src/flat_bvh.rs:396-431
    Ray.new(origin, direction)                     src/ray/ray_impl.rs:70-80

plus the batched form the GPU exists for: bvh.traverse_batch(rays) -> CSR (offsets, hits).
"Shapes" are anything with an `.aabb()` method (Bounded, src/aabb/aabb_impl.rs:28-56) and
optionally `set_bh_node_index` (BHShape, src/bounding_hierarchy.rs:53-65); numpy AABB arrays are
accepted directly.  Everything below runs on the GPU through libbvh_b200.so; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np

from . import capi
from .dtypes import BY_PREC, BY_PREC_2D, U32_MAX


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """One per device (stream + scratch pool)."""

    _default: dict[int, "Context"] = {}

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        capi.check(capi.lib().bvhgpu_create(device, C.byref(self._h)))
        self.device = device
        self._host = {}

    @classmethod
    def default(cls, device: int = 0) -> "Context":
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def set_stream(self, cuda_stream_handle: int | None):
        """Enqueue on the given cudaStream_t handle (0 = CUDA's legacy default stream); None = the context's own stream."""
        if cuda_stream_handle is None:
            capi.check(capi.lib().bvhgpu_reset_stream(self._h))
        else:
            capi.check(capi.lib().bvhgpu_set_stream(self._h, C.c_void_p(cuda_stream_handle)))

    def synchronize(self):
        capi.check(capi.lib().bvhgpu_synchronize(self._h))

    def launch_count(self) -> int:
        return int(capi.lib().bvhgpu_launch_count(self._h))

    def set_option(self, name: str, value: int):
        capi.check(capi.lib().bvhgpu_set_option(self._h, name.encode(), int(value)))

    def get_metric(self, name: str) -> float:
        out = C.c_double(0.0)
        capi.check(capi.lib().bvhgpu_get_metric(self._h, name.encode(), C.byref(out)))
        return out.value

    def host_alloc(self, nbytes: int, dtype=np.uint8) -> np.ndarray:
        """Pinned host memory on the device's NUMA node (bvhgpu_host_alloc) as a numpy array; release with host_free(arr)."""
        p = C.c_void_p()
        capi.check(capi.lib().bvhgpu_host_alloc(self._h, nbytes, C.byref(p)))
        arr = np.ctypeslib.as_array((C.c_ubyte * nbytes).from_address(p.value)).view(dtype)
        self._host.setdefault(arr.ctypes.data, p)
        return arr

    def host_free(self, arr: np.ndarray):
        p = self._host.pop(arr.ctypes.data)
        capi.check(capi.lib().bvhgpu_host_free(self._h, p))

    def close(self):
        if self._h:
            capi.lib().bvhgpu_destroy(self._h)
            self._h = C.c_void_p()


def _gather_aabbs(shapes, prec):
    d = BY_PREC[prec]
    if isinstance(shapes, np.ndarray):
        return np.ascontiguousarray(shapes, dtype=d["aabb"])
    out = np.zeros(len(shapes), dtype=d["aabb"])
    for i, s in enumerate(shapes):                      # Bounded::aabb(), once per shape
        a = s.aabb()
        out[i]["min"] = a[0]
        out[i]["max"] = a[1]
    return out


class Ray:
    """Ray<T,3> (src/ray/ray_impl.rs:17-29).  Ray.new normalises on the device (bvhgpu_rays_new_dev_*)."""

    @staticmethod
    def new(origins, directions, prec: str = "f32", ctx: Context | None = None) -> np.ndarray:
        import torch

        d = BY_PREC[prec]
        ctx = ctx or Context.default()
        o = np.ascontiguousarray(origins, dtype=d["scalar"]).reshape(-1, 3)
        v = np.ascontiguousarray(directions, dtype=d["scalar"]).reshape(-1, 3)
        n = len(o)
        dev = torch.device("cuda", ctx.device)
        to, tv = torch.from_numpy(o).to(dev), torch.from_numpy(v).to(dev)
        out = torch.empty(n * d["ray"].itemsize, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)
        capi.check(getattr(capi.lib(), f"bvhgpu_rays_new_dev_{d['suffix']}")(ctx._h, to.data_ptr(), tv.data_ptr(), n, out.data_ptr()))
        ctx.synchronize()
        return out.cpu().numpy().view(d["ray"]).copy()


class FlatBvh:
    """FlatBvh = Vec<FlatNode> (src/flat_bvh.rs:153) produced by Bvh.flatten(); traversal runs on the device tree."""

    def __init__(self, bvh: "Bvh", nodes: np.ndarray):
        self._bvh = bvh
        self.nodes = nodes

    def __len__(self):
        return len(self.nodes)

    def traverse(self, ray, shapes=None):
        return self._bvh._traverse_one(ray, shapes, capi.TRAVERSE_FLAT)

    def traverse_batch(self, rays):
        return self._bvh.traverse_batch(rays, mode=capi.TRAVERSE_FLAT)


class Bvh:
    """Device-resident Bvh<T,3>."""

    def __init__(self, handle, prec: str, ctx: Context):
        self._h = handle
        self.prec = prec
        self.ctx = ctx
        self._d = BY_PREC[prec]
        self._nodes = None
        self._node_index = None

    # ---- construction ----------------------------------------------------------------------------
    @classmethod
    def build(cls, shapes, prec: str = "f32", ctx: Context | None = None, mode: int = capi.BUILD_EXACT_SAH) -> "Bvh":
        ctx = ctx or Context.default()
        d = BY_PREC[prec]
        aabbs = _gather_aabbs(shapes, prec)
        h = C.c_void_p()
        capi.check(getattr(capi.lib(), f"bvhgpu_build_{d['suffix']}")(ctx._h, _ptr(aabbs), len(aabbs), mode, C.byref(h)))
        bvh = cls(h, prec, ctx)
        if not isinstance(shapes, np.ndarray) and len(shapes) and hasattr(shapes[0], "set_bh_node_index"):
            for s, ni in zip(shapes, bvh.node_index):    # BHShape::set_bh_node_index
                s.set_bh_node_index(int(ni))
        return bvh

    build_par = build          # rayon is off the hot path: same builder (bounding_hierarchy.rs:170-177)

    @classmethod
    def build_dev(cls, dev_ptr: int, n: int, prec: str = "f32", ctx: Context | None = None, mode: int = capi.BUILD_EXACT_SAH) -> "Bvh":
        """AABBs already on the device (C-ABI layout); asynchronous on the context's stream."""
        ctx = ctx or Context.default()
        d = BY_PREC[prec]
        h = C.c_void_p()
        capi.check(getattr(capi.lib(), f"bvhgpu_build_dev_{d['suffix']}")(ctx._h, C.c_void_p(dev_ptr), n, mode, C.byref(h)))
        return cls(h, prec, ctx)

    @classmethod
    def from_nodes(cls, nodes: np.ndarray, shapes, prec: str = "f32", ctx: Context | None = None) -> "Bvh":
        ctx = ctx or Context.default()
        d = BY_PREC[prec]
        nodes = np.ascontiguousarray(nodes, dtype=d["node"])
        aabbs = _gather_aabbs(shapes, prec)
        h = C.c_void_p()
        capi.check(getattr(capi.lib(), f"bvhgpu_tree_from_nodes_{d['suffix']}")(ctx._h, _ptr(nodes), len(nodes), _ptr(aabbs), len(aabbs), C.byref(h)))
        return cls(h, prec, ctx)

    def free(self):
        if self._h:
            getattr(capi.lib(), f"bvhgpu_tree_free_{self._d['suffix']}")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # ---- Bvh.nodes / node indices ------------------------------------------------------------------
    @property
    def num_shapes(self) -> int:
        return int(getattr(capi.lib(), f"bvhgpu_tree_num_shapes_{self._d['suffix']}")(self._h))

    def _materialise(self):
        if self._nodes is None:
            n = self.num_shapes
            nodes = np.zeros(max(2 * n - 1, 0), dtype=self._d["node"])
            idx = np.zeros(n, dtype=np.uint32)
            capi.check(getattr(capi.lib(), f"bvhgpu_tree_nodes_{self._d['suffix']}")(self._h, _ptr(nodes), _ptr(idx)))
            self._nodes, self._node_index = nodes, idx

    @property
    def nodes(self) -> np.ndarray:
        self._materialise()
        return self._nodes

    @property
    def node_index(self) -> np.ndarray:
        self._materialise()
        return self._node_index

    # ---- flatten -----------------------------------------------------------------------------------
    def flatten(self) -> FlatBvh:
        n = self.num_shapes
        cap = 0 if n == 0 else (1 if n == 1 else 3 * n - 2)
        out = np.zeros(cap, dtype=self._d["flat"])
        ln = C.c_size_t(0)
        capi.check(getattr(capi.lib(), f"bvhgpu_flatten_{self._d['suffix']}")(self._h, _ptr(out), cap, C.byref(ln)))
        return FlatBvh(self, out[: ln.value])

    def flatten_custom(self, constructor):
        """Bvh::flatten_custom (src/flat_bvh.rs:240-251): `constructor(aabb, entry, exit, shape)` applied to every FlatNode in the
        reference's emission order; aabb = (min[3], max[3])."""
        f = self.flatten().nodes
        return [constructor((f["aabb"]["min"][i], f["aabb"]["max"][i]), int(f["entry_index"][i]), int(f["exit_index"][i]), int(f["shape_index"][i])) for i in range(len(f))]

    def flatten_dev(self) -> int:
        """Build the FlatBvh on the device only (no host copy, asynchronous); returns its length."""
        ln = C.c_size_t(0)
        capi.check(getattr(capi.lib(), f"bvhgpu_flatten_{self._d['suffix']}")(self._h, None, 0, C.byref(ln)))
        return ln.value

    # ---- traversal ---------------------------------------------------------------------------------
    def traverse_batch(self, rays: np.ndarray, mode: int = capi.TRAVERSE_BVH, cap: int | None = None, compact: bool = False):
        """CSR (offsets u32[nrays+1], hits u32[total]); hits of a ray are in the reference's DFS order.
        compact=True ships only origin + direction (BVHGPU_RAYS_OD, 6 scalars per ray); the device recomputes inv_direction with
        the division Ray::new uses, so the result is bit-identical."""
        rays = np.ascontiguousarray(rays, dtype=self._d["ray"])
        nrays = len(rays)
        offsets = np.zeros(nrays + 1, dtype=np.uint32)
        cap = max(4 * nrays, 1024) if cap is None else cap
        hits = np.zeros(cap, dtype=np.uint32)
        total = C.c_size_t(0)
        fn = getattr(capi.lib(), f"bvhgpu_traverse_{self._d['suffix']}")
        if compact:
            od = np.empty((nrays, 6), dtype=self._d["scalar"])
            od[:, :3], od[:, 3:] = rays["origin"], rays["direction"]
            rays, fn = od, getattr(capi.lib(), f"bvhgpu_traverse_od_{self._d['suffix']}")
        st = fn(self._h, mode, _ptr(rays), nrays, _ptr(offsets), _ptr(hits), cap, C.byref(total))
        if st == capi.ERR_CAPACITY and total.value <= U32_MAX:
            hits = np.zeros(total.value, dtype=np.uint32)
            capi.check(getattr(capi.lib(), f"bvhgpu_traverse_fetch_{self._d['suffix']}")(self._h, _ptr(hits), total.value))
        else:
            capi.check(st)
        return offsets, hits[: total.value]

    def traverse_ordered(self, rays: np.ndarray, ascending: bool = True):
        """Batched nearest_traverse_iterator / farthest_traverse_iterator: CSR + distances, perfectly sorted per ray."""
        rays = np.ascontiguousarray(rays, dtype=self._d["ray"])
        n = len(rays)
        offsets = np.zeros(n + 1, dtype=np.uint32)
        cap = max(8 * n, 1024)
        fn = getattr(capi.lib(), f"bvhgpu_traverse_ordered_{self._d['suffix']}")
        while True:
            hits = np.zeros(cap, dtype=np.uint32)
            dists = np.zeros(cap, dtype=self._d["scalar"])
            total = C.c_size_t(0)
            st = fn(self._h, _ptr(rays), n, 1 if ascending else 0, _ptr(offsets), _ptr(hits), _ptr(dists), cap, C.byref(total))
            if st == capi.ERR_CAPACITY and total.value <= U32_MAX and total.value > cap:
                cap = total.value
                continue
            capi.check(st)
            return offsets, hits[: total.value], dists[: total.value]

    def set_triangles(self, triangles):
        """Triangle vertices of the shapes (n, 3, 3) or (n, 9): enables closest_hit(..., triangles=True).  Triangle i must lie inside shape i's AABB."""
        t = np.ascontiguousarray(triangles, dtype=self._d["scalar"]).reshape(-1, 9)
        capi.check(getattr(capi.lib(), f"bvhgpu_tree_set_triangles_{self._d['suffix']}")(self._h, _ptr(t), len(t)))

    def closest_hit(self, rays: np.ndarray, triangles: bool = False):
        """Per ray: (shape index or U32_MAX, distance or inf, uv (n, 2)).  triangles=False: the shape whose AABB is entered first (exact);
        triangles=True: Ray::intersects_triangle minimum over the triangles set with set_triangles (front-to-back, distance-pruned)."""
        rays = np.ascontiguousarray(rays, dtype=self._d["ray"])
        n = len(rays)
        shape = np.zeros(n, dtype=np.uint32)
        dist = np.zeros(n, dtype=self._d["scalar"])
        uv = np.zeros((n, 2), dtype=self._d["scalar"])
        capi.check(getattr(capi.lib(), f"bvhgpu_closest_hit_{self._d['suffix']}")(self._h, _ptr(rays), n, 1 if triangles else 0, _ptr(shape), _ptr(dist), _ptr(uv)))
        return shape, dist, uv

    def query_batch(self, kind: int, queries, mode: int = capi.TRAVERSE_BVH):
        """Bvh::traverse with Aabb / Point / Ball queries (IntersectsAabb implementors other than Ray).
        queries: (n, 6) {min,max} for capi.QUERY_AABB, (n, 3) for QUERY_POINT, (n, 4) {center, radius} for QUERY_BALL."""
        stride = {capi.QUERY_AABB: 6, capi.QUERY_POINT: 3, capi.QUERY_BALL: 4}[kind]
        q = np.ascontiguousarray(queries, dtype=self._d["scalar"]).reshape(-1, stride)
        n = len(q)
        offsets = np.zeros(n + 1, dtype=np.uint32)
        cap = max(16 * n, 1024)
        hits = np.zeros(cap, dtype=np.uint32)
        total = C.c_size_t(0)
        st = getattr(capi.lib(), f"bvhgpu_query_{self._d['suffix']}")(self._h, mode, kind, _ptr(q), n, _ptr(offsets), _ptr(hits), cap, C.byref(total))
        if st == capi.ERR_CAPACITY and total.value <= U32_MAX:
            hits = np.zeros(total.value, dtype=np.uint32)
            capi.check(getattr(capi.lib(), f"bvhgpu_traverse_fetch_{self._d['suffix']}")(self._h, _ptr(hits), total.value))
        else:
            capi.check(st)
        return offsets, hits[: total.value]

    def nearest_to_batch(self, points, mode: int = capi.TRAVERSE_BVH):
        """Bvh::nearest_to / FlatBvh::nearest_to (bvh_impl.rs:221-238, flat_bvh.rs:513-562) for shapes whose PointDistance is their AABB
        distance (the reference's UnitBox): (shape index per point, U32_MAX for an empty tree; distance per point)."""
        p = np.ascontiguousarray(points, dtype=self._d["scalar"]).reshape(-1, 3)
        shape = np.zeros(len(p), dtype=np.uint32)
        dist = np.zeros(len(p), dtype=self._d["scalar"])
        capi.check(getattr(capi.lib(), f"bvhgpu_nearest_{self._d['suffix']}")(self._h, mode, _ptr(p), len(p), _ptr(shape), _ptr(dist)))
        return shape, dist

    def nearest_triangles_batch(self, points, mode: int = capi.TRAVERSE_BVH):
        """Bvh::nearest_to / FlatBvh::nearest_to for triangle shapes (set_triangles first): the reference's walk with
        Triangle::distance_squared (testbase.rs:353-443) at the leaves, on the device: (shape index, distance) per point."""
        p = np.ascontiguousarray(points, dtype=self._d["scalar"]).reshape(-1, 3)
        shape = np.zeros(len(p), dtype=np.uint32)
        dist = np.zeros(len(p), dtype=self._d["scalar"])
        capi.check(getattr(capi.lib(), f"bvhgpu_nearest_triangles_{self._d['suffix']}")(self._h, mode, _ptr(p), len(p), _ptr(shape), _ptr(dist)))
        return shape, dist

    def nearest_candidates(self, points):
        """For shapes with their own PointDistance: CSR (offsets, shape indices) of candidate lists that contain the nearest shape of
        every point; evaluate distance_squared on each list and keep the minimum (see `nearest_to`)."""
        p = np.ascontiguousarray(points, dtype=self._d["scalar"]).reshape(-1, 3)
        n = len(p)
        offsets = np.zeros(n + 1, dtype=np.uint32)
        cap = max(64 * n, 1024)
        cand = np.zeros(cap, dtype=np.uint32)
        total = C.c_size_t(0)
        st = getattr(capi.lib(), f"bvhgpu_nearest_candidates_{self._d['suffix']}")(self._h, _ptr(p), n, _ptr(offsets), _ptr(cand), cap, C.byref(total))
        if st == capi.ERR_CAPACITY and total.value <= U32_MAX:
            cand = np.zeros(total.value, dtype=np.uint32)
            capi.check(getattr(capi.lib(), f"bvhgpu_traverse_fetch_{self._d['suffix']}")(self._h, _ptr(cand), total.value))
        else:
            capi.check(st)
        return offsets, cand[: total.value]

    def nearest_to(self, point, shapes, distance_squared):
        """BoundingHierarchy::nearest_to for one point and an arbitrary shape distance: `distance_squared(shape, point)` is the shape's
        PointDistance::distance_squared.  Returns (shape, distance) or None for an empty tree."""
        off, cand = self.nearest_candidates([point])
        best = None
        for s in cand[off[0]:off[1]]:
            d = distance_squared(shapes[int(s)], point)
            if best is None or d < best[1]:
                best = (shapes[int(s)], d)
        return None if best is None else (best[0], float(np.sqrt(best[1])))

    def traverse_dev(self, rays_ptr: int, nrays: int, offsets_ptr: int, hits_ptr: int, cap: int, mode: int = capi.TRAVERSE_BVH,
                     want_total: bool = False):
        total = C.c_size_t(0)
        fn = getattr(capi.lib(), f"bvhgpu_traverse_dev_{self._d['suffix']}")
        capi.check(fn(self._h, mode, C.c_void_p(rays_ptr), nrays, C.c_void_p(offsets_ptr), C.c_void_p(hits_ptr), cap,
                      C.byref(total) if want_total else None))
        return total.value if want_total else None

    def traverse_stats(self):
        out = (C.c_uint64 * 2)()
        capi.check(getattr(capi.lib(), f"bvhgpu_traverse_stats_{self._d['suffix']}")(self._h, out))
        return int(out[0]), int(out[1])

    def _traverse_one(self, ray, shapes, mode):
        rays = np.ascontiguousarray(ray, dtype=self._d["ray"]).reshape(1)
        _, hits = self.traverse_batch(rays, mode)
        return [shapes[int(h)] for h in hits] if shapes is not None else hits.tolist()

    def traverse(self, ray, shapes: Sequence | None = None):
        """Bvh::traverse: the shapes (or shape indices) whose AABB the ray hits, reference order."""
        return self._traverse_one(ray, shapes, capi.TRAVERSE_BVH)

    def traverse_iterator(self, ray, shapes: Sequence | None = None) -> Iterable:
        """BvhTraverseIterator (src/bvh/iter.rs): same sequence as traverse, lazily yielded on the host."""
        return iter(self._traverse_one(ray, shapes, capi.TRAVERSE_BVH))

    # ---- extras ------------------------------------------------------------------------------------
    def sah_cost(self):
        out = (C.c_double * 2)()
        capi.check(getattr(capi.lib(), f"bvhgpu_sah_cost_{self._d['suffix']}")(self._h, out))
        return float(out[0]), float(out[1])

    def refit(self, shapes):
        aabbs = _gather_aabbs(shapes, self.prec)
        capi.check(getattr(capi.lib(), f"bvhgpu_refit_{self._d['suffix']}")(self._h, _ptr(aabbs), len(aabbs)))
        self._nodes = self._node_index = None

    def optimize(self, shapes, max_growth: float = 1.5) -> int:
        """Bvh::update_shapes counterpart (src/bvh/optimization.rs:290-302): refit, then rebuild in place the subtrees whose
        surface area grew by more than `max_growth`.  `shapes` = all shapes with their current AABBs.  Returns the number
        of shapes in the rebuilt subtrees; `node_index` must be re-read (set_bh_node_index) afterwards."""
        aabbs = _gather_aabbs(shapes, self.prec)
        rebuilt = C.c_size_t(0)
        capi.check(getattr(capi.lib(), f"bvhgpu_optimize_{self._d['suffix']}")(self._h, _ptr(aabbs), len(aabbs), C.c_double(max_growth),
                                                                              C.byref(rebuilt)))
        self._nodes = self._node_index = None
        return int(rebuilt.value)

    def update_shapes(self, changed, shapes, max_growth: float = 1.5) -> int:
        """Bvh::update_shapes(changed_shape_indices, shapes) (src/bvh/optimization.rs:304-315): only the changed shapes' AABBs are sent.
        `shapes` = all shapes (objects with .aabb(), or an AABB array); max_growth <= 0: refit only.  Returns the number of shapes in
        rebuilt subtrees; node indices must be re-read (set_bh_node_index) when it is non-zero."""
        idx = np.ascontiguousarray(changed, dtype=np.uint32).reshape(-1)
        if isinstance(shapes, np.ndarray):
            fresh = np.ascontiguousarray(shapes[idx], dtype=self._d["aabb"])
        else:
            fresh = _gather_aabbs([shapes[int(i)] for i in idx], self.prec)
        rebuilt = C.c_size_t(0)
        capi.check(getattr(capi.lib(), f"bvhgpu_update_{self._d['suffix']}")(self._h, _ptr(idx), _ptr(fresh), len(idx), C.c_double(max_growth), C.byref(rebuilt)))
        self._nodes = self._node_index = None
        return int(rebuilt.value)


class Bvh2:
    """Device-resident Bvh<T,2> (the reference is generic in the dimension): build / nodes / flatten / traverse for 2-D AABBs and rays
    (bvhgpu_*_f32x2 / _f64x2).  Rays: structured array with 2-component origin, direction (normalised), inv_direction."""

    def __init__(self, handle, prec: str, ctx: Context, n: int):
        self._h, self.prec, self.ctx, self._d, self.n = handle, prec, ctx, BY_PREC_2D[prec], n

    @classmethod
    def build(cls, aabbs, prec: str = "f32", ctx: Context | None = None, mode: int = capi.BUILD_EXACT_SAH) -> "Bvh2":
        ctx = ctx or Context.default()
        d = BY_PREC_2D[prec]
        a = np.ascontiguousarray(aabbs, dtype=d["aabb"])
        h = C.c_void_p()
        capi.check(getattr(capi.lib(), f"bvhgpu_build_{d['suffix']}")(ctx._h, _ptr(a), len(a), mode, C.byref(h)))
        return cls(h, prec, ctx, len(a))

    def free(self):
        if self._h:
            getattr(capi.lib(), f"bvhgpu_tree_free_{self._d['suffix']}")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def nodes_and_index(self):
        nodes = np.zeros(max(2 * self.n - 1, 0), dtype=self._d["node"])
        idx = np.zeros(self.n, dtype=np.uint32)
        capi.check(getattr(capi.lib(), f"bvhgpu_tree_nodes_{self._d['suffix']}")(self._h, _ptr(nodes), _ptr(idx)))
        return nodes, idx

    def flatten(self) -> np.ndarray:
        cap = 0 if self.n == 0 else (1 if self.n == 1 else 3 * self.n - 2)
        out = np.zeros(cap, dtype=self._d["flat"])
        ln = C.c_size_t(0)
        capi.check(getattr(capi.lib(), f"bvhgpu_flatten_{self._d['suffix']}")(self._h, _ptr(out), cap, C.byref(ln)))
        return out[: ln.value]

    def traverse_batch(self, rays, mode: int = capi.TRAVERSE_BVH):
        rays = np.ascontiguousarray(rays, dtype=self._d["ray"])
        n = len(rays)
        offsets = np.zeros(n + 1, dtype=np.uint32)
        cap = max(16 * n, 1024)
        fn = getattr(capi.lib(), f"bvhgpu_traverse_{self._d['suffix']}")
        while True:
            hits = np.zeros(cap, dtype=np.uint32)
            total = C.c_size_t(0)
            st = fn(self._h, mode, _ptr(rays), n, _ptr(offsets), _ptr(hits), cap, C.byref(total))
            if st == capi.ERR_CAPACITY and total.value > cap and total.value <= U32_MAX:
                cap = total.value
                continue
            capi.check(st)
            return offsets, hits[: total.value]
