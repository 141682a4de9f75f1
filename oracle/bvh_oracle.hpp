// ============================================================================
// oracle/bvh_oracle.hpp -- TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT.
//
// CPU restatement (C++17, host only) of the reference crate's hot path
// (svenstaro/bvh 0.12.0): SAH-binned Bvh::build, Bvh::flatten, and ray
// traversal (recursive, flat, iterator), plus the reference's deterministic
// test fixtures.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, load or call this code; the
// product path (bvh_b200/) never links or falls back to it.
//
// PARITY STATUS: the reference is Rust and cannot be compiled in this image
// (no rustc/cargo, nalgebra/num-traits not vendored).  The restatement is
// pinned against every golden / known-answer test the reference holds for
// this path (tests/test_oracle_goldens.py) and against an
// independent numpy restatement (tests/pyref.py).  Exact node arrays for
// non-trivial scenes are NOT published by the reference, so beyond those
// goldens "parity" means GPU == this oracle, bit for bit.
//
// Arithmetic rules: everything in T, round to nearest,
// NO fused multiply-add (compile with -ffp-contract=off, never -ffast-math).
// All file:line citations are relative to /root/reference/.
// ============================================================================
#pragma once
#include <cstdint>
#include <cstddef>
#include <cmath>
#include <limits>
#include <vector>
#include <stdexcept>
#include <array>
#include <string>
#include <atomic>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <algorithm>

namespace orc {

static constexpr uint32_t U32_MAX = 0xFFFFFFFFu;

// ----------------------------------------------------------------------------
// POD mirrors (host layout shared with the tests' numpy dtypes).
// ----------------------------------------------------------------------------
template <class T> struct Aabb3 {       // src/aabb/aabb_impl.rs:10-16
    T min[3];
    T max[3];
};
template <class T> struct Ray3 {        // src/ray/ray_impl.rs:17-29
    T origin[3];
    T direction[3];
    T inv_direction[3];
};
// BvhNode enum flattened to a POD (src/bvh/bvh_node.rs:21-47).
//   Leaf : child_l == U32_MAX, shape = shape_index, child_r = U32_MAX, AABBs = empty
//   Node : child_l/child_r indices, shape = number of shapes below the node
//          (extra information, not in the reference enum), AABBs of children.
template <class T> struct Node {
    uint32_t parent;
    uint32_t child_l;
    uint32_t child_r;
    uint32_t shape;
    Aabb3<T> l_aabb;
    Aabb3<T> r_aabb;
    bool is_leaf() const { return child_l == U32_MAX; }
};
template <class T> struct FlatNode {    // src/flat_bvh.rs:17-46
    Aabb3<T> aabb;
    uint32_t entry_index;
    uint32_t exit_index;
    uint32_t shape_index;
};

// ----------------------------------------------------------------------------
// L0 geometry.
// ----------------------------------------------------------------------------
// simba simd_min/simd_max for primitive floats [3p nalgebra ^0.34 / simba]:
//   simd_min(a,b) = if a <= b {a} else {b};  simd_max(a,b) = if a >= b {a} else {b}
template <class T> inline T smin(T a, T b) { return a <= b ? a : b; }
template <class T> inline T smax(T a, T b) { return a >= b ? a : b; }
// src/utils.rs:29,52
template <class T> inline T fast_min(T x, T y) { return x < y ? x : y; }
template <class T> inline T fast_max(T x, T y) { return x > y ? x : y; }

template <class T> inline Aabb3<T> aabb_empty() {   // src/aabb/aabb_impl.rs:119-124
    const T inf = std::numeric_limits<T>::infinity();
    return Aabb3<T>{{inf, inf, inf}, {-inf, -inf, -inf}};
}
template <class T> inline Aabb3<T> aabb_join(const Aabb3<T>& a, const Aabb3<T>& b) {  // :303-308
    Aabb3<T> r;
    for (int k = 0; k < 3; ++k) { r.min[k] = smin(a.min[k], b.min[k]); r.max[k] = smax(a.max[k], b.max[k]); }
    return r;
}
template <class T> inline Aabb3<T> aabb_grow(const Aabb3<T>& a, const T p[3]) {      // :375-380
    Aabb3<T> r;
    for (int k = 0; k < 3; ++k) { r.min[k] = smin(a.min[k], p[k]); r.max[k] = smax(a.max[k], p[k]); }
    return r;
}
template <class T> inline void aabb_center(const Aabb3<T>& a, T c[3]) {              // :501-504
    for (int k = 0; k < 3; ++k) c[k] = a.min[k] * T(0.5) + a.max[k] * T(0.5);
}
template <class T> inline T aabb_surface_area(const Aabb3<T>& a) {                   // :459-461, :551-554
    const T sx = a.max[0] - a.min[0], sy = a.max[1] - a.min[1], sz = a.max[2] - a.min[2];
    return T(2) * ((sx * sx + sy * sy) + sz * sz);      // [3p] nalgebra dot for R=3: a + b + c
}
template <class T> inline int aabb_largest_axis(const Aabb3<T>& a) {                 // :594-596, [3p] imax
    T best = a.max[0] - a.min[0];
    int axis = 0;
    for (int k = 1; k < 3; ++k) { const T s = a.max[k] - a.min[k]; if (s > best) { best = s; axis = k; } }
    return axis;
}
// src/aabb/aabb_impl.rs:198-202, 221-224 (used by the invariants only)
template <class T> inline bool aabb_approx_contains_point(const Aabb3<T>& a, const T p[3], T eps) {
    for (int k = 0; k < 3; ++k) { if (!((p[k] - a.min[k]) > -eps)) return false; if (!((p[k] - a.max[k]) < eps)) return false; }
    return true;
}
template <class T> inline bool aabb_approx_contains_aabb(const Aabb3<T>& a, const Aabb3<T>& o, T eps) {
    return aabb_approx_contains_point(a, o.min, eps) && aabb_approx_contains_point(a, o.max, eps);
}

// ----------------------------------------------------------------------------
// L1 ray.
// ----------------------------------------------------------------------------
// src/ray/ray_impl.rs:70-80.  [3p] normalize = v / sqrt((x*x + y*y) + z*z).
template <class T> inline Ray3<T> ray_new(const T o[3], const T d[3]) {
    Ray3<T> r;
    const T n2 = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
    const T n = std::sqrt(n2);
    for (int k = 0; k < 3; ++k) {
        r.origin[k] = o[k];
        r.direction[k] = d[k] / n;
        r.inv_direction[k] = T(1) / r.direction[k];
    }
    return r;
}
// src/ray/intersect_default.rs:16-37 (scalar path; the SIMD specialisations in
// intersect_simd.rs compute the same predicate).
template <class T> inline bool ray_intersects_aabb(const Ray3<T>& ray, const Aabb3<T>& b) {
    T l[3], r[3];
    for (int k = 0; k < 3; ++k) {
        l[k] = (b.min[k] - ray.origin[k]) * ray.inv_direction[k];
        r[k] = (b.max[k] - ray.origin[k]) * ray.inv_direction[k];
    }
    for (int k = 0; k < 3; ++k) if (std::isnan(l[k]) || std::isnan(r[k])) return false;   // utils.rs:113-115
    T tmin = smin(l[0], r[0]), tmax = smax(l[0], r[0]);                                  // inf_sup
    for (int k = 1; k < 3; ++k) {
        tmin = smax(tmin, smin(l[k], r[k]));                                             // inf.max()
        tmax = smin(tmax, smax(l[k], r[k]));                                             // sup.min()
    }
    return tmax >= fast_max(tmin, T(0));
}

// Ray::intersection_slice_for_aabb (src/ray/ray_impl.rs:118-145)
template <class T> inline bool ray_slice_for_aabb(const Ray3<T>& ray, const Aabb3<T>& b, T& tmin_out, T& tmax_out) {
    T l[3], r[3];
    for (int k = 0; k < 3; ++k) {
        l[k] = (b.min[k] - ray.origin[k]) * ray.inv_direction[k];
        r[k] = (b.max[k] - ray.origin[k]) * ray.inv_direction[k];
    }
    for (int k = 0; k < 3; ++k) if (std::isnan(l[k]) || std::isnan(r[k])) return false;
    T tmin = smin(l[0], r[0]), tmax = smax(l[0], r[0]);
    for (int k = 1; k < 3; ++k) { tmin = smax(tmin, smin(l[k], r[k])); tmax = smin(tmax, smax(l[k], r[k])); }
    tmin_out = fast_max(tmin, T(0));
    tmax_out = tmax;
    return !(tmin_out > tmax) ;
}

// Ray::intersects_triangle (src/ray/ray_impl.rs:154-213): Moeller-Trumbore with backface culling.  [3p] nalgebra: cross =
// (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x), 3-term dot (a.x*b.x + a.y*b.y) + a.z*b.z.  Returns the distance
// (+inf = no hit); u, v as the reference leaves them.
template <class T> inline void cross3(const T a[3], const T b[3], T o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
template <class T> inline T ray_intersects_triangle(const Ray3<T>& ray, const T a[3], const T b[3], const T c[3], T& u_out, T& v_out) {
    const T INF = std::numeric_limits<T>::infinity(), EPS = std::numeric_limits<T>::epsilon();
    T ab[3], ac[3], uvec[3], ao[3], vvec[3];
    for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; }
    cross3(ray.direction, ac, uvec);
    const T det = (ab[0] * uvec[0] + ab[1] * uvec[1]) + ab[2] * uvec[2];
    u_out = T(0); v_out = T(0);
    if (det < EPS) return INF;                                   // :178-180 (NaN det falls through, as in the reference)
    const T inv_det = T(1) / det;
    for (int k = 0; k < 3; ++k) ao[k] = ray.origin[k] - a[k];
    const T u = ((ao[0] * uvec[0] + ao[1] * uvec[1]) + ao[2] * uvec[2]) * inv_det;
    u_out = u;
    if (!(u >= T(0) && u <= T(1))) return INF;                   // :191-193 RangeInclusive::contains
    cross3(ao, ab, vvec);
    const T v = ((ray.direction[0] * vvec[0] + ray.direction[1] * vvec[1]) + ray.direction[2] * vvec[2]) * inv_det;
    v_out = v;
    if (v < T(0) || u + v > T(1)) return INF;                    // :201-203
    const T dist = ((ac[0] * vvec[0] + ac[1] * vvec[1]) + ac[2] * vvec[2]) * inv_det;
    return dist > EPS ? dist : INF;                              // :207-211
}

// Other IntersectsAabb implementors (src/aabb/intersection.rs:35-45, src/ball.rs:85-106).
// Query records: kind 1 = Aabb {min,max} (6 T), kind 2 = Point (3 T), kind 3 = Ball {center, radius} (4 T).
template <class T> inline bool aabb_intersects_aabb(const Aabb3<T>& q, const Aabb3<T>& b) {          // aabb_impl.rs:240-248
    for (int i = 0; i < 3; ++i) if (q.max[i] < b.min[i] || b.max[i] < q.min[i]) return false;
    return true;
}
template <class T> inline bool aabb_contains_point(const Aabb3<T>& b, const T p[3]) {                  // aabb_impl.rs:175-177 ([3p] nalgebra: all components)
    for (int i = 0; i < 3; ++i) if (!(p[i] >= b.min[i])) return false;
    for (int i = 0; i < 3; ++i) if (!(p[i] <= b.max[i])) return false;
    return true;
}
template <class T> inline bool ball_intersects_aabb(const T c[3], T radius, const Aabb3<T>& b) {      // ball.rs:85-99
    T d2 = T(0);
    for (int i = 0; i < 3; ++i) {
        T x = c[i];                                    // f32::clamp: if x < min {min} ; if x > max {max}
        if (x < b.min[i]) x = b.min[i];
        if (x > b.max[i]) x = b.max[i];
        const T d = x - c[i];
        d2 = d2 + d * d;                               // powi(2)
    }
    return d2 <= radius * radius;
}
template <class T> inline bool query_intersects(int kind, const T* q, const Aabb3<T>& b) {
    if (kind == 1) return aabb_intersects_aabb(*reinterpret_cast<const Aabb3<T>*>(q), b);
    if (kind == 2) return aabb_contains_point(b, q);
    return ball_intersects_aabb(q, q[3], b);
}
// ----------------------------------------------------------------------------
// L2a build.
// ----------------------------------------------------------------------------
struct BuildStats {
    uint64_t prim_visits = 0;       // P: sum over internal nodes of |indices|
    uint64_t degenerate_splits = 0; // nodes taking bvh_node.rs:114-124
    uint64_t nosplit_fallthrough = 0; // nodes where no cost < +inf (bvh_node.rs:239 never true)
    uint32_t max_depth = 0;
};

template <class T> struct BuildArgs {   // src/bvh/bvh_node.rs:437-446
    uint32_t start, count;              // slice of the (shared) index array
    uint32_t parent, node;
    uint32_t depth;
    Aabb3<T> aabb_bounds, centroid_bounds;
};

// src/utils.rs:97-109
template <class T>
inline void joint_aabb_of_shapes(const uint32_t* idx, uint32_t count, const Aabb3<T>* shapes,
                                 Aabb3<T>& aabb, Aabb3<T>& centroid) {
    aabb = aabb_empty<T>();
    centroid = aabb_empty<T>();
    for (uint32_t i = 0; i < count; ++i) {
        const Aabb3<T>& s = shapes[idx[i]];
        aabb = aabb_join(aabb, s);
        T c[3];
        aabb_center(s, c);
        centroid = aabb_grow(centroid, c);
    }
}

// One call of BvhNode::prep_build (src/bvh/bvh_node.rs:81-180) including
// build_buckets (:183-279).  Returns false for a leaf; otherwise fills l/r.
// `scratch` plays the role of the thread-local bucket vectors (bucket.rs:14-24).
template <class T>
inline bool prep_build(const BuildArgs<T>& a, const Aabb3<T>* shapes, uint32_t* indices,
                       Node<T>* nodes, uint32_t* node_index_of_shape,
                       std::array<std::vector<uint32_t>, 6>& scratch,
                       BuildArgs<T>& left, BuildArgs<T>& right, BuildStats& st) {
    uint32_t* I = indices + a.start;
    if (a.depth > st.max_depth) st.max_depth = a.depth;
    if (a.count == 1) {                                         // :95-104
        Node<T>& n = nodes[a.node];
        n.parent = a.parent; n.child_l = U32_MAX; n.child_r = U32_MAX; n.shape = I[0];
        n.l_aabb = aabb_empty<T>(); n.r_aabb = aabb_empty<T>();
        node_index_of_shape[I[0]] = a.node;
        return false;
    }
    st.prim_visits += a.count;
    const int axis = aabb_largest_axis(a.centroid_bounds);       // :107
    const T ext = a.centroid_bounds.max[axis] - a.centroid_bounds.min[axis];   // :108
    Aabb3<T> lab, lcb, rab, rcb;
    uint32_t nl;
    if (ext < std::numeric_limits<T>::epsilon()) {              // :114-124
        st.degenerate_splits++;
        nl = a.count / 2;
        joint_aabb_of_shapes(I, nl, shapes, lab, lcb);
        joint_aabb_of_shapes(I + nl, a.count - nl, shapes, rab, rcb);
    } else {                                                    // build_buckets :183-279
        struct Bucket { uint32_t size; Aabb3<T> aabb, centroid; };    // utils.rs:59-95
        Bucket buckets[6];
        for (auto& b : buckets) { b.size = 0; b.aabb = aabb_empty<T>(); b.centroid = aabb_empty<T>(); }
        for (auto& v : scratch) v.clear();
        const T K = T(6) - T(0.01);                             // :214-215 (T::from(0.01): f64 literal rounded to T)
        for (uint32_t i = 0; i < a.count; ++i) {                // :204-222
            const Aabb3<T>& s = shapes[I[i]];
            T c[3];
            aabb_center(s, c);
            const T rel = (c[axis] - a.centroid_bounds.min[axis]) / ext;
            const T scaled = rel * K;
            // to_usize(): truncation toward zero; NaN / out of range => the reference panics.
            const int b = (int)scaled;
            buckets[b].size += 1;
            buckets[b].aabb = aabb_join(buckets[b].aabb, s);
            buckets[b].centroid = aabb_grow(buckets[b].centroid, c);
            scratch[b].push_back(I[i]);
        }
        int min_bucket = 0;                                     // :225-230
        T min_cost = std::numeric_limits<T>::infinity();
        lab = lcb = rab = rcb = aabb_empty<T>();
        bool any = false;
        for (int s = 0; s < 5; ++s) {                           // :231-247
            Bucket L{0, aabb_empty<T>(), aabb_empty<T>()}, R{0, aabb_empty<T>(), aabb_empty<T>()};
            for (int b = 0; b <= s; ++b) { L.size += buckets[b].size; L.aabb = aabb_join(L.aabb, buckets[b].aabb); L.centroid = aabb_join(L.centroid, buckets[b].centroid); }
            for (int b = s + 1; b < 6; ++b) { R.size += buckets[b].size; R.aabb = aabb_join(R.aabb, buckets[b].aabb); R.centroid = aabb_join(R.centroid, buckets[b].centroid); }
            const T cost = (T(L.size) * aabb_surface_area(L.aabb) + T(R.size) * aabb_surface_area(R.aabb))
                           / aabb_surface_area(a.aabb_bounds);
            if (cost < min_cost) {
                any = true;
                min_bucket = s; min_cost = cost;
                lab = L.aabb; lcb = L.centroid; rab = R.aabb; rcb = R.centroid;
            }
        }
        if (!any) st.nosplit_fallthrough++;
        nl = 0;                                                 // :250-272
        for (int b = 0; b <= min_bucket; ++b) nl += (uint32_t)scratch[b].size();
        uint32_t w = 0;
        for (int b = 0; b < 6; ++b) for (uint32_t v : scratch[b]) I[w++] = v;
    }
    const uint32_t child_l = a.node + 1;                        // :138-142
    const uint32_t child_r = child_l + (2 * nl - 1);
    Node<T>& n = nodes[a.node];                                 // :145-151
    n.parent = a.parent; n.child_l = child_l; n.child_r = child_r; n.shape = a.count;
    n.l_aabb = lab; n.r_aabb = rab;
    left = BuildArgs<T>{a.start, nl, a.node, child_l, a.depth + 1, lab, lcb};               // :154-179
    right = BuildArgs<T>{a.start + nl, a.count - nl, a.node, child_r, a.depth + 1, rab, rcb};
    return true;
}

// Sequential subtree build (BvhNode::build, bvh_node.rs:55-60) with an explicit stack.
template <class T>
inline void build_subtree(const BuildArgs<T>& root, const Aabb3<T>* shapes, uint32_t* indices,
                          Node<T>* nodes, uint32_t* node_index_of_shape, BuildStats& st) {
    std::array<std::vector<uint32_t>, 6> scratch;
    std::vector<BuildArgs<T>> stack;
    stack.push_back(root);
    while (!stack.empty()) {
        BuildArgs<T> a = stack.back();
        stack.pop_back();
        BuildArgs<T> l, r;
        if (prep_build(a, shapes, indices, nodes, node_index_of_shape, scratch, l, r, st)) {
            stack.push_back(r);
            stack.push_back(l);
        }
    }
}

// Bvh::build (src/bvh/bvh_impl.rs:40-96).  nodes must hold 2n-1 entries,
// node_index_of_shape n entries (BHShape::set_bh_node_index, bounding_hierarchy.rs:58).
template <class T>
inline BuildStats build(const Aabb3<T>* shapes, uint32_t n, Node<T>* nodes, uint32_t* node_index_of_shape) {
    BuildStats st;
    if (n == 0) return st;                                      // :57-59
    std::vector<uint32_t> indices(n);
    for (uint32_t i = 0; i < n; ++i) indices[i] = i;            // :61-63
    BuildArgs<T> root{0, n, 0, 0, 0, {}, {}};
    joint_aabb_of_shapes(indices.data(), n, shapes, root.aabb_bounds, root.centroid_bounds);  // :74
    build_subtree(root, shapes, indices.data(), nodes, node_index_of_shape, st);
    return st;
}

// Bvh::build_par analogue (bounding_hierarchy.rs:170-177 + rayon_executor,
// bvh_impl.rs:527-543): children are forked iff their combined shape count is
// >= 64; the result is identical to the sequential build because children
// write disjoint slices.  rayon is replaced by a small shared-queue pool.
template <class T>
inline BuildStats build_par(const Aabb3<T>* shapes, uint32_t n, Node<T>* nodes, uint32_t* node_index_of_shape,
                            unsigned threads) {
    BuildStats st;
    if (n == 0) return st;
    if (threads <= 1) return build(shapes, n, nodes, node_index_of_shape);
    std::vector<uint32_t> indices(n);
    for (uint32_t i = 0; i < n; ++i) indices[i] = i;
    BuildArgs<T> root{0, n, 0, 0, 0, {}, {}};
    joint_aabb_of_shapes(indices.data(), n, shapes, root.aabb_bounds, root.centroid_bounds);

    std::mutex mu;
    std::condition_variable cv;
    std::vector<BuildArgs<T>> queue;
    size_t outstanding = 1;         // tasks queued or running
    queue.push_back(root);
    std::vector<BuildStats> tstats(threads);

    auto worker = [&](unsigned tid) {
        std::array<std::vector<uint32_t>, 6> scratch;
        BuildStats ls;                              // thread-local copy: no false sharing in the hot loop
        for (;;) {
            BuildArgs<T> a;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !queue.empty() || outstanding == 0; });
                if (queue.empty()) { tstats[tid] = ls; return; }
                a = queue.back();
                queue.pop_back();
            }
            // Run this task: descend, forking the right child while the grain rule allows.
            bool have = true;
            while (have) {
                BuildArgs<T> l, r;
                if (!prep_build(a, shapes, indices.data(), nodes, node_index_of_shape, scratch, l, r, ls)) break;
                if (l.count + r.count < 64) {                   // bvh_impl.rs:534
                    build_subtree(l, shapes, indices.data(), nodes, node_index_of_shape, ls);
                    build_subtree(r, shapes, indices.data(), nodes, node_index_of_shape, ls);
                    have = false;
                } else {
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        queue.push_back(r);
                        ++outstanding;
                    }
                    cv.notify_one();
                    a = l;
                }
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                --outstanding;
                if (outstanding == 0) cv.notify_all();
            }
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; ++t) pool.emplace_back(worker, t);
    for (auto& th : pool) th.join();
    for (auto& s : tstats) {
        st.prim_visits += s.prim_visits;
        st.degenerate_splits += s.degenerate_splits;
        st.nosplit_fallthrough += s.nosplit_fallthrough;
        st.max_depth = std::max(st.max_depth, s.max_depth);
    }
    return st;
}

// ----------------------------------------------------------------------------
// L2b flatten: literal restatement of the recursion (src/flat_bvh.rs:60-143,
// 240-251, 312-319) with an explicit stack that reproduces Vec::push order.
// ----------------------------------------------------------------------------
template <class T>
inline void flatten(const Node<T>* nodes, uint32_t n_nodes, std::vector<FlatNode<T>>& out) {
    out.clear();
    if (n_nodes == 0) return;                                   // :245-248
    // Frames emulate flatten_custom(node, next_free) / create_flat_branch.
    struct Frame { uint32_t node; int stage; uint32_t nav_l, nav_r; };
    auto push_leaf = [&](const Node<T>& nd) {                   // :129-141
        FlatNode<T> f;
        f.aabb = aabb_empty<T>();
        f.entry_index = U32_MAX;
        f.exit_index = (uint32_t)out.size() + 1;
        f.shape_index = nd.shape;
        out.push_back(f);
    };
    if (nodes[0].is_leaf()) { push_leaf(nodes[0]); return; }
    std::vector<Frame> stack;
    stack.push_back(Frame{0, 0, 0, 0});
    while (!stack.empty()) {
        Frame& fr = stack.back();
        const Node<T>& nd = nodes[fr.node];
        if (nd.is_leaf()) { push_leaf(nd); stack.pop_back(); continue; }
        if (fr.stage == 0) {                                    // create_flat_branch(child_l): push dummy (:71-74)
            fr.stage = 1;
            fr.nav_l = (uint32_t)out.size();
            out.push_back(FlatNode<T>{aabb_empty<T>(), 0, 0, 0});
            stack.push_back(Frame{nd.child_l, 0, 0, 0});
        } else if (fr.stage == 1) {                             // overwrite left navigator (:80-88), then right dummy
            FlatNode<T>& nav = out[fr.nav_l];
            nav.aabb = nd.l_aabb; nav.entry_index = fr.nav_l + 1; nav.exit_index = (uint32_t)out.size(); nav.shape_index = U32_MAX;
            fr.stage = 2;
            fr.nav_r = (uint32_t)out.size();
            out.push_back(FlatNode<T>{aabb_empty<T>(), 0, 0, 0});
            const uint32_t child = nd.child_r;
            stack.push_back(Frame{child, 0, 0, 0});
        } else {
            FlatNode<T>& nav = out[fr.nav_r];
            nav.aabb = nd.r_aabb; nav.entry_index = fr.nav_r + 1; nav.exit_index = (uint32_t)out.size(); nav.shape_index = U32_MAX;
            stack.pop_back();
        }
    }
}

// ----------------------------------------------------------------------------
// Traversal.
// ----------------------------------------------------------------------------
struct TraverseStats { uint64_t node_visits = 0; uint64_t slab_tests = 0; uint64_t leaf_visits = 0; uint64_t hits = 0; };

// BvhNode::traverse_recursive (src/bvh/bvh_node.rs:288-319) + Bvh::traverse
// (src/bvh/bvh_impl.rs:104-119).  Explicit stack, left child first.
template <class T>
inline void traverse_recursive(const Node<T>* nodes, uint32_t n_nodes, const Aabb3<T>* shapes,
                               const Ray3<T>& ray, std::vector<uint32_t>& out, TraverseStats* st = nullptr) {
    if (n_nodes == 0) return;                                   // bvh_impl.rs:109-112
    if (nodes[0].is_leaf()) {                                   // bvh_node.rs:314 (root leaf re-tests the shape)
        if (st) st->slab_tests++;
        if (ray_intersects_aabb(ray, shapes[nodes[0].shape])) { out.push_back(nodes[0].shape); if (st) st->hits++; }
        return;
    }
    uint32_t stack[128];
    std::vector<uint32_t> big;          // only used if depth exceeds 128
    int sp = 0;
    auto push = [&](uint32_t v) { if (sp < 128) stack[sp++] = v; else { big.push_back(v); ++sp; } };
    auto pop = [&]() -> uint32_t { if (sp > 128) { uint32_t v = big.back(); big.pop_back(); --sp; return v; } return stack[--sp]; };
    push(0);
    while (sp > 0) {
        const uint32_t i = pop();
        const Node<T>& nd = nodes[i];
        if (st) st->node_visits++;
        if (nd.is_leaf()) { out.push_back(nd.shape); if (st) st->hits++; continue; }
        if (st) st->slab_tests += 2;
        const bool hl = ray_intersects_aabb(ray, nd.l_aabb);
        const bool hr = ray_intersects_aabb(ray, nd.r_aabb);
        if (hr) push(nd.child_r);       // popped after the whole left subtree
        if (hl) push(nd.child_l);
    }
}

// Bvh::traverse / FlatBvh::traverse for the non-ray queries (kind 1..3, see query_intersects).
template <class T>
inline void traverse_query(int kind, const T* q, const Node<T>* nodes, uint32_t n_nodes, const FlatNode<T>* flat, uint32_t n_flat,
                           const Aabb3<T>* shapes, bool use_flat, std::vector<uint32_t>& out) {
    if (use_flat) {                                             // src/flat_bvh.rs:396-431
        uint32_t index = 0;
        while (index < n_flat) {
            const FlatNode<T>& node = flat[index];
            if (node.entry_index == U32_MAX) {
                if (query_intersects(kind, q, shapes[node.shape_index])) out.push_back(node.shape_index);
                index = node.exit_index;
            } else index = query_intersects(kind, q, node.aabb) ? node.entry_index : node.exit_index;
        }
        return;
    }
    if (n_nodes == 0) return;
    if (nodes[0].is_leaf()) { if (query_intersects(kind, q, shapes[nodes[0].shape])) out.push_back(nodes[0].shape); return; }
    std::vector<uint32_t> stack{0};
    while (!stack.empty()) {
        const uint32_t i = stack.back();
        stack.pop_back();
        const Node<T>& nd = nodes[i];
        if (nd.is_leaf()) { out.push_back(nd.shape); continue; }
        const bool hl = query_intersects(kind, q, nd.l_aabb), hr = query_intersects(kind, q, nd.r_aabb);
        if (hr) stack.push_back(nd.child_r);
        if (hl) stack.push_back(nd.child_l);
    }
}

// BvhTraverseIterator (src/bvh/iter.rs:6-182), literal state machine with the
// 32-slot stack.  Returns false if the reference would have panicked
// (stack overflow, iter.rs:56-59).
template <class T>
inline bool traverse_iterator(const Node<T>* nodes, uint32_t n_nodes, const Aabb3<T>* shapes,
                              const Ray3<T>& ray, std::vector<uint32_t>& out) {
    bool has_node;
    if (n_nodes == 0) has_node = false;                         // iter.rs:164-182
    else if (nodes[0].is_leaf()) has_node = ray_intersects_aabb(ray, shapes[nodes[0].shape]);
    else has_node = true;
    uint32_t stack[32];
    uint32_t sp = 0, node = 0;
    for (;;) {
        if (sp == 0 && !has_node) break;
        if (has_node) {
            if (sp >= 32) return false;
            stack[sp++] = node;
            const Node<T>& nd = nodes[node];                    // move_left
            if (!nd.is_leaf() && ray_intersects_aabb(ray, nd.l_aabb)) { node = nd.child_l; has_node = true; }
            else has_node = false;
        } else {
            node = stack[--sp];
            const Node<T>& nd = nodes[node];
            if (!nd.is_leaf()) {                                // move_right
                if (ray_intersects_aabb(ray, nd.r_aabb)) { node = nd.child_r; has_node = true; }
                else has_node = false;
            } else {
                has_node = false;
                out.push_back(nd.shape);
            }
        }
    }
    return true;
}

// FlatBvh::traverse (src/flat_bvh.rs:396-431).
template <class T>
inline void traverse_flat(const FlatNode<T>* flat, uint32_t n_flat, const Aabb3<T>* shapes,
                          const Ray3<T>& ray, std::vector<uint32_t>& out, TraverseStats* st = nullptr) {
    uint32_t index = 0;
    while (index < n_flat) {
        const FlatNode<T>& node = flat[index];
        if (st) st->node_visits++;
        if (node.entry_index == U32_MAX) {
            if (st) { st->leaf_visits++; st->slab_tests++; }
            if (ray_intersects_aabb(ray, shapes[node.shape_index])) { out.push_back(node.shape_index); if (st) st->hits++; }
            index = node.exit_index;
        } else {
            if (st) st->slab_tests++;
            index = ray_intersects_aabb(ray, node.aabb) ? node.entry_index : node.exit_index;
        }
    }
}

// ----------------------------------------------------------------------------
// Invariants: Bvh::is_consistent / assert_tight (src/bvh/bvh_impl.rs:280-485).
// ----------------------------------------------------------------------------
template <class T>
inline bool is_consistent(const Node<T>* nodes, uint32_t n_nodes, const Aabb3<T>* shapes) {
    if (n_nodes == 0) return true;
    const T inf = std::numeric_limits<T>::infinity();
    struct F { uint32_t node, parent; Aabb3<T> outer; };
    std::vector<F> stack;
    stack.push_back(F{0, 0, Aabb3<T>{{-inf, -inf, -inf}, {inf, inf, inf}}});
    uint64_t count = 0;
    const T eps = std::numeric_limits<T>::epsilon();
    bool ok = true;
    while (!stack.empty()) {
        F f = stack.back();
        stack.pop_back();
        if (f.node >= n_nodes) return false;
        ++count;
        if (count > n_nodes) return false;
        const Node<T>& nd = nodes[f.node];
        if (nd.parent != f.parent) ok = false;
        if (nd.is_leaf()) {
            if (!aabb_approx_contains_aabb(f.outer, shapes[nd.shape], eps)) ok = false;
        } else {
            if (!aabb_approx_contains_aabb(f.outer, nd.l_aabb, eps)) ok = false;
            if (!aabb_approx_contains_aabb(f.outer, nd.r_aabb, eps)) ok = false;
            stack.push_back(F{nd.child_r, f.node, nd.r_aabb});
            stack.push_back(F{nd.child_l, f.node, nd.l_aabb});
        }
    }
    return ok && count == n_nodes;
}
template <class T>
inline bool is_tight(const Node<T>* nodes, uint32_t n_nodes) {
    if (n_nodes == 0 || nodes[0].is_leaf()) return true;
    struct F { uint32_t node; Aabb3<T> outer; };
    std::vector<F> stack;
    stack.push_back(F{0, aabb_join(nodes[0].l_aabb, nodes[0].r_aabb)});
    while (!stack.empty()) {
        F f = stack.back();
        stack.pop_back();
        const Node<T>& nd = nodes[f.node];
        if (nd.is_leaf()) continue;
        const Aabb3<T> j = aabb_join(nd.l_aabb, nd.r_aabb);
        for (int k = 0; k < 3; ++k) if (!(j.min[k] == f.outer.min[k] && j.max[k] == f.outer.max[k])) return false;
        stack.push_back(F{nd.child_r, nd.r_aabb});
        stack.push_back(F{nd.child_l, nd.l_aabb});
    }
    return true;
}

// Whole-tree SAH cost (definition fixed in SURVEY.md 8d; the reference has no
// whole-tree cost): sum over non-root nodes of SA(child aabb stored in the
// parent) / SA(root aabb), in double.  pseudo = reference's 2*|size|^2,
// geometric = 2(xy+yz+zx).
template <class T>
inline void sah_cost(const Node<T>* nodes, uint32_t n_nodes, double& pseudo, double& geometric) {
    pseudo = geometric = 0.0;
    if (n_nodes == 0 || nodes[0].is_leaf()) return;
    auto sa = [](const Aabb3<T>& a, bool geo) {
        const double x = (double)a.max[0] - (double)a.min[0], y = (double)a.max[1] - (double)a.min[1], z = (double)a.max[2] - (double)a.min[2];
        return geo ? 2.0 * (x * y + y * z + z * x) : 2.0 * (x * x + y * y + z * z);
    };
    const Aabb3<T> root = aabb_join(nodes[0].l_aabb, nodes[0].r_aabb);
    const double rp = sa(root, false), rg = sa(root, true);
    for (uint32_t i = 0; i < n_nodes; ++i) {
        if (nodes[i].is_leaf()) continue;
        pseudo += sa(nodes[i].l_aabb, false) + sa(nodes[i].r_aabb, false);
        geometric += sa(nodes[i].l_aabb, true) + sa(nodes[i].r_aabb, true);
    }
    pseudo /= rp;
    geometric /= rg;
}

// ----------------------------------------------------------------------------
// nearest_to (src/bvh/bvh_impl.rs:221-238, src/bvh/bvh_node.rs:327-372, src/flat_bvh.rs:513-562).
// The shape distance is the caller's PointDistance::distance_squared; two are restated here:
//   DIST_AABB     : the shape's AABB distance, as the reference's UnitBox does (testbase.rs:101-105)
//   DIST_TRIANGLE : the reference's test triangle (testbase.rs:353-443, closest_point_triangle, "adapted from Embree")
// ----------------------------------------------------------------------------
template <class T> inline T aabb_min_distance_squared(const Aabb3<T>& a, const T p[3]) {    // aabb_impl.rs:618-629
    T o[3];
    for (int k = 0; k < 3; ++k) {
        const T hs = (a.max[k] - a.min[k]) * T(0.5);          // half_size(), :479-481
        const T c = a.min[k] + hs;
        const T d = p[k] - c;
        const T q = std::fabs(d) - hs;
        o[k] = q > T(0) ? q : T(0);                            // x.max(0): 0 for NaN as well
    }
    return (o[0] * o[0] + o[1] * o[1]) + o[2] * o[2];
}
template <class T> inline T dot3(const T a[3], const T b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
template <class T> inline void closest_point_segment(const T p[3], const T a[3], const T b[3], T out[3]) {   // testbase.rs:353-363
    T ab[3], ap[3];
    for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ap[k] = p[k] - a[k]; }
    const T m = dot3(ab, ab);
    T s = dot3(ab, ap) / m;
    s = s < T(0) ? T(0) : (s > T(1) ? T(1) : s);               // clamp (NaN propagates as in Rust)
    for (int k = 0; k < 3; ++k) out[k] = a[k] + s * ab[k];
}
template <class T> inline void closest_point_triangle(const T p[3], const T a[3], const T b[3], const T c[3], T out[3]) {   // :367-434
    auto eq = [](const T* x, const T* y) { return x[0] == y[0] && x[1] == y[1] && x[2] == y[2]; };
    const bool e_ab = eq(a, b), e_bc = eq(b, c), e_ac = eq(a, c);
    if (e_ab && e_bc && e_ac) { for (int k = 0; k < 3; ++k) out[k] = a[k]; return; }
    if (e_ab) { closest_point_segment(p, a, c, out); return; }
    if (e_bc) { closest_point_segment(p, a, b, out); return; }
    if (e_ac) { closest_point_segment(p, a, b, out); return; }
    T ab[3], ac[3], ap[3], bp[3], cp[3];
    for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; }
    const T d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= T(0) && d2 <= T(0)) { for (int k = 0; k < 3; ++k) out[k] = a[k]; return; }
    for (int k = 0; k < 3; ++k) bp[k] = p[k] - b[k];
    const T d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= T(0) && d4 <= d3) { for (int k = 0; k < 3; ++k) out[k] = b[k]; return; }
    for (int k = 0; k < 3; ++k) cp[k] = p[k] - c[k];
    const T d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= T(0) && d5 <= d6) { for (int k = 0; k < 3; ++k) out[k] = c[k]; return; }
    const T vc = d1 * d4 - d3 * d2;
    if (vc <= T(0) && d1 >= T(0) && d3 <= T(0)) { const T v = d1 / (d1 - d3); for (int k = 0; k < 3; ++k) out[k] = a[k] + v * ab[k]; return; }
    const T vb = d5 * d2 - d1 * d6;
    if (vb <= T(0) && d2 >= T(0) && d6 <= T(0)) { const T v = d2 / (d2 - d6); for (int k = 0; k < 3; ++k) out[k] = a[k] + v * ac[k]; return; }
    const T va = d3 * d6 - d5 * d4;
    if (va <= T(0) && d4 - d3 >= T(0) && d5 - d6 >= T(0)) {
        const T v = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int k = 0; k < 3; ++k) out[k] = b[k] + v * (c[k] - b[k]);
        return;
    }
    const T denom = T(1) / (va + vb + vc);
    const T v = vb * denom, w = vc * denom;
    for (int k = 0; k < 3; ++k) out[k] = (a[k] + v * ab[k]) + w * ac[k];
}
enum { DIST_AABB = 0, DIST_TRIANGLE = 1 };
template <class T> struct ShapeDist {
    int kind; const Aabb3<T>* shapes; const T* tris;            // tris: 9 scalars per shape (a, b, c) for DIST_TRIANGLE
    T operator()(uint32_t s, const T p[3]) const {
        if (kind == DIST_AABB) return aabb_min_distance_squared(shapes[s], p);
        T q[3], d[3];
        closest_point_triangle(p, tris + 9 * (size_t)s, tris + 9 * (size_t)s + 3, tris + 9 * (size_t)s + 6, q);   // testbase.rs:436-443
        for (int k = 0; k < 3; ++k) d[k] = p[k] - q[k];
        return dot3(d, d);
    }
};
// Bvh::nearest_to: returns the shape index (U32_MAX for an empty tree) and the distance (sqrt of the squared one, :237)
template <class T>
inline uint32_t nearest_to(const Node<T>* nodes, uint32_t n_nodes, const T p[3], const ShapeDist<T>& dist, T& out_dist, uint64_t* evals = nullptr) {
    out_dist = T(0);
    if (n_nodes == 0) return U32_MAX;
    uint32_t best = U32_MAX;
    T best_d = T(0);
    // explicit stack replay of nearest_to_recursive (bvh_node.rs:327-372): a frame = (node, which of the two ordered children is next)
    struct F { uint32_t idx[2]; T d[2]; int next; };
    std::vector<F> stack;
    auto enter = [&](uint32_t i) {
        const Node<T>& nd = nodes[i];
        if (nd.is_leaf()) {
            const T d = dist(nd.shape, p);
            if (evals) ++*evals;
            if (best == U32_MAX || d < best_d) { best = nd.shape; best_d = d; }
            return;
        }
        F f;
        f.idx[0] = nd.child_l; f.d[0] = aabb_min_distance_squared(nd.l_aabb, p);
        f.idx[1] = nd.child_r; f.d[1] = aabb_min_distance_squared(nd.r_aabb, p);
        if (f.d[0] > f.d[1]) { std::swap(f.idx[0], f.idx[1]); std::swap(f.d[0], f.d[1]); }
        f.next = 0;
        stack.push_back(f);
    };
    enter(0);
    while (!stack.empty()) {
        F& f = stack.back();
        if (f.next == 2) { stack.pop_back(); continue; }
        const int k = f.next++;
        const uint32_t child = f.idx[k];
        const T cd = f.d[k];
        if (best == U32_MAX || cd < best_d) enter(child);       // may push: `f` is not used after this
    }
    out_dist = std::sqrt(best_d);
    return best;
}
// FlatBvh::nearest_to (flat_bvh.rs:513-562)
template <class T>
inline uint32_t nearest_to_flat(const FlatNode<T>* flat, uint32_t n_flat, const T p[3], const ShapeDist<T>& dist, T& out_dist) {
    out_dist = T(0);
    if (n_flat == 0) return U32_MAX;
    uint32_t best = U32_MAX;
    T best_d = T(0);
    uint32_t index = 0;
    while (index < n_flat) {
        const FlatNode<T>& nd = flat[index];
        if (nd.entry_index == U32_MAX) {                        // is_leaf()
            const T d = dist(nd.shape_index, p);
            if (best == U32_MAX || d < best_d) { best = nd.shape_index; best_d = d; }
            index = nd.exit_index;
        } else {
            const T md = aabb_min_distance_squared(nd.aabb, p);
            index = (best == U32_MAX || md < best_d) ? nd.entry_index : nd.exit_index;
        }
    }
    out_dist = std::sqrt(best_d);
    return best;
}

// ----------------------------------------------------------------------------
// Bvh::update_shapes = remove_shape* then add_shape* (src/bvh/optimization.rs:70-389): the reference's
// sequential re-insertion.  It appends / swap-removes nodes, so the result is NOT in build's preorder
// layout any more (child_l != i+1 in general); every function above only follows child indices and
// works on it unchanged.  The per-node shape count in `Node::shape` (our extra, not in the reference
// enum) is not maintained by these functions; recount() restores it.
// `shapes` are the CURRENT (already moved) AABBs, as in the reference, where `shapes[i].aabb()` is
// evaluated at the time of the call.
// ----------------------------------------------------------------------------
template <class T> struct DynBvh {
    std::vector<Node<T>> nodes;
    std::vector<uint32_t> node_index;           // BHShape::bh_node_index of every shape
};

template <class T> inline Aabb3<T> get_node_aabb(const DynBvh<T>& b, uint32_t i, const Aabb3<T>* shapes) {   // bvh_node.rs:616-625
    const Node<T>& nd = b.nodes[i];
    return nd.is_leaf() ? shapes[nd.shape] : aabb_join(nd.l_aabb, nd.r_aabb);
}
template <class T> inline Node<T> make_leaf(uint32_t parent, uint32_t shape) {
    Node<T> nd;
    nd.parent = parent; nd.child_l = U32_MAX; nd.child_r = U32_MAX; nd.shape = shape;
    nd.l_aabb = aabb_empty<T>(); nd.r_aabb = aabb_empty<T>();
    return nd;
}
template <class T> inline bool node_is_left_child(const DynBvh<T>& b, uint32_t i) {     // optimization.rs:18-24
    return b.nodes[b.nodes[i].parent].child_l == i;
}
template <class T> inline bool node_is_right_child(const DynBvh<T>& b, uint32_t i) {    // :26-32
    return b.nodes[b.nodes[i].parent].child_r == i;
}
// optimization.rs:34-65
template <class T> inline void connect_nodes(DynBvh<T>& b, uint32_t child, uint32_t parent, bool left_child, const Aabb3<T>* shapes) {
    const Aabb3<T> child_aabb = get_node_aabb(b, child, shapes);
    Node<T>& p = b.nodes[parent];
    if (p.is_leaf()) throw std::logic_error("connect_nodes: parent is a leaf");           // unreachable!() in the reference
    if (left_child) { p.child_l = child; p.l_aabb = child_aabb; }
    else            { p.child_r = child; p.r_aabb = child_aabb; }
    b.nodes[child].parent = parent;
}
// optimization.rs:317-351
template <class T> inline void fix_aabbs_ascending(DynBvh<T>& b, const Aabb3<T>* shapes, uint32_t node_index) {
    uint32_t index_to_fix = node_index;
    auto differs = [](const Aabb3<T>& x, const Aabb3<T>& y) {                              // Aabb: PartialEq on min and max
        for (int k = 0; k < 3; ++k) if (x.min[k] != y.min[k] || x.max[k] != y.max[k]) return true;
        return false;
    };
    while (index_to_fix != 0) {
        const uint32_t parent = b.nodes[index_to_fix].parent;
        Node<T>& pn = b.nodes[parent];
        if (pn.is_leaf()) { index_to_fix = 0; continue; }
        const Aabb3<T> l = get_node_aabb(b, pn.child_l, shapes), r = get_node_aabb(b, pn.child_r, shapes);
        bool stop = true;
        if (differs(l, pn.l_aabb)) { stop = false; pn.l_aabb = l; }
        if (differs(r, pn.r_aabb)) { stop = false; pn.r_aabb = r; }
        index_to_fix = stop ? 0 : parent;
    }
}
// optimization.rs:353-389
template <class T> inline void swap_and_remove_index(DynBvh<T>& b, uint32_t node_index) {
    const uint32_t end = (uint32_t)b.nodes.size() - 1;
    if (node_index != end) {
        b.nodes[node_index] = b.nodes[end];
        const uint32_t parent_index = b.nodes[node_index].parent;
        Node<T>& parent = b.nodes[parent_index];
        if (parent.is_leaf()) throw std::logic_error("swap_and_remove_index: parent is a leaf");
        if (parent.child_l == end) parent.child_l = node_index;
        else {
            if (parent.child_r != end) throw std::logic_error("swap_and_remove_index: moved node is not a child of its parent");
            parent.child_r = node_index;
        }
        const Node<T>& moved = b.nodes[node_index];
        if (moved.is_leaf()) b.node_index[moved.shape] = node_index;
        else { b.nodes[moved.child_l].parent = node_index; b.nodes[moved.child_r].parent = node_index; }
    }
    b.nodes.resize(end);
}
// optimization.rs:70-206
template <class T> inline void add_shape(DynBvh<T>& b, const Aabb3<T>* shapes, uint32_t new_shape_index) {
    uint32_t node_index = 0;
    const Aabb3<T> shape_aabb = shapes[new_shape_index];
    const T shape_sa = aabb_surface_area(shape_aabb);
    if (b.nodes.empty()) {
        b.nodes.push_back(make_leaf<T>(0, new_shape_index));
        b.node_index[new_shape_index] = 0;
        return;
    }
    for (;;) {
        const Node<T> cur = b.nodes[node_index];                // copy: the vector may grow below
        if (!cur.is_leaf()) {
            const Aabb3<T> left_expand = aabb_join(cur.l_aabb, shape_aabb);
            const Aabb3<T> right_expand = aabb_join(cur.r_aabb, shape_aabb);
            const T send_left = aabb_surface_area(cur.r_aabb) + aabb_surface_area(left_expand);
            const T send_right = aabb_surface_area(cur.l_aabb) + aabb_surface_area(right_expand);
            const Aabb3<T> merged_aabb = aabb_join(cur.r_aabb, cur.l_aabb);
            const T merged = aabb_surface_area(merged_aabb) + shape_sa;
            const T min_send = send_left < send_right ? send_left : send_right;
            if (merged < min_send * T(3) / T(10)) {             // "merge is more expensive only do when it's significantly better"
                const uint32_t l_index = (uint32_t)b.nodes.size();
                b.node_index[new_shape_index] = l_index;
                b.nodes.push_back(make_leaf<T>(node_index, new_shape_index));
                const uint32_t r_index = (uint32_t)b.nodes.size();
                Node<T> new_right = cur;
                new_right.parent = node_index;
                b.nodes.push_back(new_right);
                b.nodes[cur.child_r].parent = r_index;
                b.nodes[cur.child_l].parent = r_index;
                Node<T>& here = b.nodes[node_index];
                here.l_aabb = shape_aabb; here.child_l = l_index; here.r_aabb = merged_aabb; here.child_r = r_index; here.parent = cur.parent;
                return;
            } else if (send_left < send_right) {
                if (node_index == cur.child_l) throw std::logic_error("broken loop");
                b.nodes[node_index].l_aabb = left_expand;
                node_index = cur.child_l;
            } else {
                if (node_index == cur.child_r) throw std::logic_error("broken loop");
                b.nodes[node_index].r_aabb = right_expand;
                node_index = cur.child_r;
            }
        } else {                                                  // split the leaf into a node over {new, old}
            const uint32_t l_index = (uint32_t)b.nodes.size();
            b.node_index[new_shape_index] = l_index;
            b.nodes.push_back(make_leaf<T>(node_index, new_shape_index));
            const Aabb3<T> child_r_aabb = shapes[cur.shape];
            const uint32_t child_r_index = (uint32_t)b.nodes.size();
            b.node_index[cur.shape] = child_r_index;
            b.nodes.push_back(make_leaf<T>(node_index, cur.shape));
            Node<T>& here = b.nodes[node_index];
            here.parent = cur.parent; here.child_l = l_index; here.child_r = child_r_index; here.shape = 2;
            here.l_aabb = shape_aabb; here.r_aabb = child_r_aabb;
            fix_aabbs_ascending(b, shapes, cur.parent);
            return;
        }
    }
}
// optimization.rs:208-288 (swap_shape == false: the variant update_shapes uses)
template <class T> inline void remove_shape(DynBvh<T>& b, const Aabb3<T>* shapes, uint32_t deleted_shape_index) {
    if (b.nodes.empty()) throw std::logic_error("can't remove a node from a bvh with only one node");
    const uint32_t dead = b.node_index[deleted_shape_index];
    if (b.nodes.size() == 1) {
        if (dead != 0 || !b.nodes[0].is_leaf()) throw std::logic_error("remove_shape: single node is not this leaf");
        b.nodes.clear();
        return;
    }
    if (!b.nodes[dead].is_leaf()) throw std::logic_error("remove_shape: bh_node_index is not a leaf");
    const uint32_t parent_index = b.nodes[dead].parent;
    const uint32_t gp_index = b.nodes[parent_index].parent;
    uint32_t sibling_index;
    if (node_is_left_child(b, dead)) sibling_index = b.nodes[parent_index].child_r;
    else {
        if (!node_is_right_child(b, dead)) throw std::logic_error("remove_shape: node is neither child of its parent");
        sibling_index = b.nodes[parent_index].child_l;
    }
    if (parent_index == gp_index) {                               // a child of the root goes: the sibling becomes the root
        if (parent_index != 0) throw std::logic_error("Circular node that wasn't root");
        const Node<T> sib = b.nodes[sibling_index];
        if (!sib.is_leaf()) {
            connect_nodes(b, sib.child_l, parent_index, true, shapes);
            connect_nodes(b, sib.child_r, parent_index, false, shapes);
        } else {
            b.nodes[0] = sib;
            b.nodes[0].parent = 0;
            b.node_index[b.nodes[0].shape] = 0;
        }
        swap_and_remove_index(b, std::max(sibling_index, dead));
        swap_and_remove_index(b, std::min(sibling_index, dead));
    } else {
        const bool parent_is_left = node_is_left_child(b, parent_index);
        connect_nodes(b, sibling_index, gp_index, parent_is_left, shapes);
        fix_aabbs_ascending(b, shapes, gp_index);
        swap_and_remove_index(b, std::max(dead, parent_index));
        swap_and_remove_index(b, std::min(parent_index, dead));
    }
}
// optimization.rs:290-302: all removals first, then all insertions, both in the caller's order.
template <class T> inline void update_shapes(DynBvh<T>& b, const uint32_t* changed, uint32_t n_changed, const Aabb3<T>* shapes) {
    for (uint32_t i = 0; i < n_changed; ++i) remove_shape(b, shapes, changed[i]);
    for (uint32_t i = 0; i < n_changed; ++i) add_shape(b, shapes, changed[i]);
}
// Node::shape of inner nodes = number of shapes below (our extra field): recomputed after updates.
template <class T> inline void recount(DynBvh<T>& b) {
    if (b.nodes.empty() || b.nodes[0].is_leaf()) return;
    std::vector<std::pair<uint32_t, int>> stack;
    stack.push_back({0u, 0});
    while (!stack.empty()) {
        auto& f = stack.back();
        Node<T>& nd = b.nodes[f.first];
        if (nd.is_leaf()) { stack.pop_back(); continue; }
        if (f.second == 0) { f.second = 1; const uint32_t l = nd.child_l, r = nd.child_r; stack.push_back({l, 0}); stack.push_back({r, 0}); }
        else {
            auto cnt = [&](uint32_t c) { return b.nodes[c].is_leaf() ? 1u : b.nodes[c].shape; };
            nd.shape = cnt(nd.child_l) + cnt(nd.child_r);
            stack.pop_back();
        }
    }
}

// ----------------------------------------------------------------------------
// Fixtures (src/testbase.rs).
// ----------------------------------------------------------------------------
// ---- closest hit, the way callers of the reference compute it ------------------------------------------------
// kind 0: the shape whose AABB the ray enters first = the first element of a PERFECTLY sorted nearest_traverse_iterator
//         (src/bvh/distance_traverse.rs is best-effort); candidates = Bvh::traverse, key = (entry distance, DFS order).
// kind 1: triangle soups: candidates = Bvh::traverse(ray), each tested with Ray::intersects_triangle, minimum distance; ties go to
//         the lower shape index.  tris: 9 T per shape.  Returns U32_MAX / +inf when nothing is hit.
template <class T>
inline uint32_t closest_hit(const Node<T>* nodes, uint32_t n_nodes, const Aabb3<T>* shapes, const T* tris, int kind, const Ray3<T>& ray,
                            T& dist_out, T& u_out, T& v_out) {
    std::vector<uint32_t> cand;
    traverse_recursive(nodes, n_nodes, shapes, ray, cand);
    uint32_t best = U32_MAX;
    T bd = std::numeric_limits<T>::infinity(), bu = T(0), bv = T(0);
    for (uint32_t s : cand) {
        if (kind == 0) {
            T t0, t1;
            if (!ray_slice_for_aabb(ray, shapes[s], t0, t1)) continue;      // (cannot happen for a tight tree; kept for from_nodes trees)
            if (best == U32_MAX || t0 < bd) { best = s; bd = t0; }
        } else {
            T u, v;
            const T d = ray_intersects_triangle(ray, tris + 9 * (size_t)s, tris + 9 * (size_t)s + 3, tris + 9 * (size_t)s + 6, u, v);
            if (d < bd || (d == bd && d < std::numeric_limits<T>::infinity() && s < best)) { best = s; bd = d; bu = u; bv = v; }
        }
    }
    dist_out = bd; u_out = bu; v_out = bv;
    return best;
}

inline uint64_t splitmix64(uint64_t& x) {                       // testbase.rs:560-566
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline void next_point3_raw(uint64_t& seed, int32_t& a, int32_t& b, int32_t& c) {   // :569-575
    const uint64_t u = splitmix64(seed);
    const int64_t ia = (int64_t)((u >> 32) & 0xFFFFFFFFull) - 0x80000000ll;
    const int64_t ib = (int64_t)(u & 0xFFFFFFFFull) - 0x80000000ll;
    const uint64_t ub = (uint64_t)ib;
    const int64_t rot = (int64_t)((ub << 6) | (ub >> 58));      // i64::rotate_left(6)
    const int64_t ic = ia ^ rot;
    a = (int32_t)ia; b = (int32_t)ib; c = (int32_t)ic;          // `as i32` truncates
}
// next_point3 (:578-597), generic in T so that config 5 can restate it in f64
// from the same integer stream.  T = float is the reference.
template <class T> inline void next_point3(uint64_t& seed, const Aabb3<T>& bounds, T out[3]) {
    int32_t r[3];
    next_point3_raw(seed, r[0], r[1], r[2]);
    const T imax = (T)2147483647;                               // i32::MAX as f32 == 2147483648.0f
    for (int k = 0; k < 3; ++k) {
        const T fv = (((T)r[k] / imax) + T(1)) * T(0.5);
        const T size = bounds.max[k] - bounds.min[k];
        out[k] = bounds.min[k] + fv * size;
    }
}
template <class T> inline Aabb3<T> default_bounds() {           // :600-605
    return Aabb3<T>{{T(-100000), T(-100000), T(-100000)}, {T(100000), T(100000), T(100000)}};
}
// Triangle::new (:347-357): aabb = empty.grow(a).grow(b).grow(c)
template <class T> inline Aabb3<T> triangle_aabb(const T a[3], const T b[3], const T c[3]) {
    return aabb_grow(aabb_grow(aabb_grow(aabb_empty<T>(), a), b), c);
}
// push_cube (:490-556): 12 triangles, vertices pos +- 0.5.  tri_out: 9 T per triangle.
template <class T> inline void push_cube(const T pos[3], std::vector<T>& tri_out) {
    auto V = [&](T sx, T sy, T sz, T v[3]) { v[0] = pos[0] + sx; v[1] = pos[1] + sy; v[2] = pos[2] + sz; };
    const T h = T(0.5);
    T tfr[3], tbr[3], tbl[3], tfl[3], bfr[3], bbr[3], bbl[3], bfl[3];
    V(h, h, -h, tfr); V(h, h, h, tbr); V(-h, h, h, tbl); V(-h, h, -h, tfl);
    V(h, -h, -h, bfr); V(h, -h, h, bbr); V(-h, -h, h, bbl); V(-h, -h, -h, bfl);
    const T* tris[12][3] = {
        {tbr, tfr, tfl}, {tfl, tbl, tbr}, {bfl, bfr, bbr}, {bbr, bbl, bfl},
        {tbl, tfl, bfl}, {bfl, bbl, tbl}, {bfr, tfr, tbr}, {tbr, bbr, bfr},
        {tfl, tfr, bfr}, {bfr, bfl, tfl}, {bbr, tbr, tbl}, {tbl, bbl, bbr}};
    for (auto& t : tris) for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) tri_out.push_back(t[v][k]);
}
template <class T> inline void create_n_cubes(uint32_t n_cubes, const Aabb3<T>& bounds, std::vector<T>& tri_out) {   // :608-615
    uint64_t seed = 0;
    tri_out.clear();
    tri_out.reserve((size_t)n_cubes * 108);
    for (uint32_t i = 0; i < n_cubes; ++i) {
        T p[3];
        next_point3(seed, bounds, p);
        push_cube(p, tri_out);
    }
}
template <class T> inline Ray3<T> create_ray(uint64_t& seed, const Aabb3<T>& bounds) {     // :687-691
    T o[3], d[3];
    next_point3(seed, bounds, o);
    next_point3(seed, bounds, d);
    return ray_new(o, d);
}
// generate_aligned_boxes (:109-116) + UnitBox::aabb (:84-90)
template <class T> inline void aligned_boxes(std::vector<Aabb3<T>>& out) {
    out.clear();
    for (int x = -10; x < 11; ++x) {
        const T pos[3] = {(T)x, T(0), T(0)};
        Aabb3<T> b;
        for (int k = 0; k < 3; ++k) { b.min[k] = pos[k] + T(-0.5); b.max[k] = pos[k] + T(0.5); }
        out.push_back(b);
    }
}

}  // namespace orc
