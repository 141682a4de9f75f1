// bvh_b200/csrc/closest.cu -- closest hit per ray with distance pruning: the form a ray tracer needs.
//
// The reference gives its callers two things: the candidate set (Bvh::traverse, or the best-effort distance-ordered iterators of
// src/bvh/distance_traverse.rs / child_distance_traverse.rs) and the primitive test Ray::intersects_triangle
// (src/ray/ray_impl.rs:154-213); the closest hit is the caller's loop over both (src/bvh/iter.rs:330-365 does exactly that in its
// benchmark).  Here the loop runs on the device, front to back, and subtrees whose AABB is entered behind the best hit so far are
// never opened:
//   AABB mode      result = the shape whose AABB the ray enters first, key (entry distance, DFS order) -- i.e. the first element of a
//                  perfectly sorted nearest_traverse_iterator.  Entry distances are Ray::intersection_slice_for_aabb
//                  (src/ray/ray_impl.rs:118-145) bit for bit; pruning keeps ties (entry == best), so the result is EXACTLY the
//                  minimum over Bvh::traverse's candidates.
//   triangle mode  result = the triangle with the smallest Ray::intersects_triangle distance (Moeller-Trumbore with backface culling,
//                  same operation order, no FMA), ties to the lower shape index.  A subtree is skipped when its entry distance exceeds
//                  best * (1 + 2^-16): the slab distance and the Moeller-Trumbore distance are different roundings of the same
//                  quantity, so pruning at exactly `entry > best` could drop a hit that wins by an ulp; with the margin the result can
//                  differ from the unpruned minimum only between two hits whose distances agree to ~1e-5 relative (stated in the
//                  tests as the tolerance).
// The walk needs no stack: nodes carry parent links, a lane remembers which child it comes back from and re-derives the near / far
// order from the node (same loads, same bits), so any tree depth works (the reference's iterators use a 32-slot stack / a heap).
#include "internal.h"

namespace bvhb200 {

template <class T> __device__ __forceinline__ T cmin(T a, T b);
template <> __device__ __forceinline__ float cmin(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ double cmin(double a, double b) { return fmin(a, b); }
template <class T> __device__ __forceinline__ T cmax(T a, T b);
template <> __device__ __forceinline__ float cmax(float a, float b) { return fmaxf(a, b); }
template <> __device__ __forceinline__ double cmax(double a, double b) { return fmax(a, b); }

// Ray::intersection_slice_for_aabb (src/ray/ray_impl.rs:118-145): entry distance (clamped at 0) or "no intersection".
template <class T>
__device__ __forceinline__ bool slice_entry(const T o[3], const T inv[3], const T mn[3], const T mx[3], T& entry) {
    const T l0 = mul_rn(sub_rn(mn[0], o[0]), inv[0]), r0 = mul_rn(sub_rn(mx[0], o[0]), inv[0]);
    const T l1 = mul_rn(sub_rn(mn[1], o[1]), inv[1]), r1 = mul_rn(sub_rn(mx[1], o[1]), inv[1]);
    const T l2 = mul_rn(sub_rn(mn[2], o[2]), inv[2]), r2 = mul_rn(sub_rn(mx[2], o[2]), inv[2]);
    const bool nan = (l0 != l0) | (r0 != r0) | (l1 != l1) | (r1 != r1) | (l2 != l2) | (r2 != r2);
    const T tmin = cmax(cmax(cmin(l0, r0), cmin(l1, r1)), cmin(l2, r2));
    const T tmax = cmin(cmin(cmax(l0, r0), cmax(l1, r1)), cmax(l2, r2));
    entry = tmin > T(0) ? tmin : T(0);
    return !nan && !(entry > tmax);
}

template <class T> __device__ __forceinline__ void cross_rn(const T a[3], const T b[3], T o[3]) {       // nalgebra 3-D cross
    o[0] = sub_rn(mul_rn(a[1], b[2]), mul_rn(a[2], b[1]));
    o[1] = sub_rn(mul_rn(a[2], b[0]), mul_rn(a[0], b[2]));
    o[2] = sub_rn(mul_rn(a[0], b[1]), mul_rn(a[1], b[0]));
}
template <class T> __device__ __forceinline__ T dot_rn(const T a[3], const T b[3]) { return add_rn(add_rn(mul_rn(a[0], b[0]), mul_rn(a[1], b[1])), mul_rn(a[2], b[2])); }

// Ray::intersects_triangle (src/ray/ray_impl.rs:154-213).  Returns the distance (+inf: miss / back face / behind the origin).
template <class T>
__device__ __forceinline__ T moeller_trumbore(const T o[3], const T dir[3], const T a[3], const T b[3], const T c[3], T& u_out, T& v_out) {
    const T INF = Traits<T>::inf(), EPS = Traits<T>::eps();
    T ab[3], ac[3], uvec[3], ao[3], vvec[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ab[k] = sub_rn(b[k], a[k]); ac[k] = sub_rn(c[k], a[k]); }
    cross_rn(dir, ac, uvec);
    const T det = dot_rn(ab, uvec);
    u_out = T(0); v_out = T(0);
    if (det < EPS) return INF;
    const T inv_det = div_rn(T(1), det);
#pragma unroll
    for (int k = 0; k < 3; ++k) ao[k] = sub_rn(o[k], a[k]);
    const T u = mul_rn(dot_rn(ao, uvec), inv_det);
    u_out = u;
    if (!(u >= T(0) && u <= T(1))) return INF;
    cross_rn(ao, ab, vvec);
    const T v = mul_rn(dot_rn(dir, vvec), inv_det);
    v_out = v;
    if (v < T(0) || add_rn(u, v) > T(1)) return INF;
    const T dist = mul_rn(dot_rn(ac, vvec), inv_det);
    return dist > EPS ? dist : INF;
}

template <class T> struct DTri { T a[3], pa, b[3], pb, c[3], pc; };       // 48 B / 96 B: three vector loads per triangle

template <class T>
__global__ void __launch_bounds__(256) pack_tris_kernel(const T* __restrict__ tris9, uint32_t n, DTri<T>* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DTri<T> t;
    const T* p = tris9 + 9 * (size_t)i;
#pragma unroll
    for (int k = 0; k < 3; ++k) { t.a[k] = p[k]; t.b[k] = p[3 + k]; t.c[k] = p[6 + k]; }
    t.pa = t.pb = t.pc = T(0);
    out[i] = t;
}

template <class T>
__global__ void __launch_bounds__(256) fill_nohit_kernel(uint32_t n, uint32_t* __restrict__ shape, T* __restrict__ dist, T* __restrict__ uv) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    shape[i] = BVH_INVALID; dist[i] = Traits<T>::inf();
    if (uv) { uv[2 * (size_t)i] = T(0); uv[2 * (size_t)i + 1] = T(0); }
}

template <class T, bool TRI>
__global__ void __launch_bounds__(128) closest_kernel(const typename Traits<T>::Node* __restrict__ nodes, uint32_t n_shapes,
                                                      const typename Traits<T>::DAabb* __restrict__ aabb, const DTri<T>* __restrict__ tris,
                                                      const T* __restrict__ rays, uint32_t ray_stride, uint32_t nrays,
                                                      uint32_t* __restrict__ out_shape, T* __restrict__ out_dist, T* __restrict__ out_uv) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrays) return;
    T o[3], dir[3], inv[3];
    {
        const T* p = rays + (size_t)ray_stride * r;
#pragma unroll
        for (int k = 0; k < 3; ++k) { o[k] = __ldg(p + k); dir[k] = __ldg(p + 3 + k); inv[k] = ray_stride == 9 ? __ldg(p + 6 + k) : div_rn(T(1), dir[k]); }
    }
    const T INF = Traits<T>::inf();
    const T margin = TRI ? add_rn(T(1), T(1.0 / 65536.0)) : T(1);
    uint32_t best = BVH_INVALID, best_key = BVH_INVALID;
    T best_d = INF, bu = T(0), bv = T(0);

    auto leaf = [&](uint32_t shape, uint32_t node_idx) {
        if (TRI) {
            const DTri<T>& t = tris[shape];
            T a[3], b[3], c[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { a[k] = __ldg(&t.a[k]); b[k] = __ldg(&t.b[k]); c[k] = __ldg(&t.c[k]); }
            T u, v;
            const T d = moeller_trumbore(o, dir, a, b, c, u, v);
            if (d < best_d || (d == best_d && d < INF && shape < best)) { best = shape; best_d = d; bu = u; bv = v; }
        } else {
            T mn[3], mx[3], e;
            load_aabb(aabb + shape, mn, mx);
            if (slice_entry(o, inv, mn, mx, e)) {
                if (best == BVH_INVALID || e < best_d || (e == best_d && node_idx < best_key)) { best = shape; best_d = e; best_key = node_idx; }
            }
        }
    };

    if (n_shapes == 1) {                                       // root leaf (bvh_node.rs:314 tests the shape's own AABB)
        T mn[3], mx[3], e;
        load_aabb(aabb + nodes[0].shape, mn, mx);
        if (slice_entry(o, inv, mn, mx, e)) leaf(nodes[0].shape, 0u);
    } else {
        uint32_t node = 0, from = BVH_INVALID;                  // from: the child we are coming back from (BVH_INVALID: arriving from the parent)
        for (;;) {
            const uint4 meta = __ldg(reinterpret_cast<const uint4*>(nodes + node));      // parent, child_l, child_r, shape / count
            if (meta.y == BVH_INVALID) {
                leaf(meta.w, node);
                from = node; node = meta.x;
                continue;
            }
            const typename Traits<T>::Node& nd = nodes[node];
            T lmn[3], lmx[3], rmn[3], rmx[3], el, er;
#pragma unroll
            for (int k = 0; k < 3; ++k) { lmn[k] = __ldg(&nd.l_aabb.min[k]); lmx[k] = __ldg(&nd.l_aabb.max[k]); rmn[k] = __ldg(&nd.r_aabb.min[k]); rmx[k] = __ldg(&nd.r_aabb.max[k]); }
            const bool hl = slice_entry(o, inv, lmn, lmx, el), hr = slice_entry(o, inv, rmn, rmx, er);
            if (!hl) el = INF;
            if (!hr) er = INF;
            const bool left_first = el <= er;                   // front to back; ties: left (DFS order)
            const uint32_t near_i = left_first ? meta.y : meta.z, far_i = left_first ? meta.z : meta.y;
            const T near_e = left_first ? el : er, far_e = left_first ? er : el;
            const bool near_ok = left_first ? hl : hr, far_ok = left_first ? hr : hl;
            const T bound = mul_rn(best_d, margin);             // inf stays inf
            uint32_t next = BVH_INVALID;
            if (from == BVH_INVALID) {
                if (near_ok && near_e <= bound) next = near_i;
                else from = near_i;                             // skipped: as if we had just come back from it
            }
            if (next == BVH_INVALID && from == near_i) {
                if (far_ok && far_e <= bound) next = far_i;
                else from = far_i;
            }
            if (next != BVH_INVALID) { node = next; from = BVH_INVALID; continue; }
            if (node == 0) break;                               // back from the far child of the root
            from = node; node = meta.x;
        }
    }
    out_shape[r] = best;
    out_dist[r] = best_d;
    if (out_uv) { out_uv[2 * (size_t)r] = bu; out_uv[2 * (size_t)r + 1] = bv; }
}

template <class T>
int set_triangles(Tree<T>* tree, const T* tris9, size_t n, bool dev_input) {
    bvhgpu_ctx* ctx = tree->ctx;
    if (n != tree->n) { set_error("set_triangles: %zu triangles for a tree over %u shapes", n, tree->n); return BVHGPU_ERR_INVALID; }
    if (n == 0) return BVHGPU_OK;
    Scratch scratch(ctx);
    const T* d_in = tris9;
    if (!dev_input) {
        T* staged = nullptr;
        BVH_TRY(scratch.get(&staged, 9 * n));
        BVH_CUDA_TRY(cudaMemcpyAsync(staged, tris9, sizeof(T) * 9 * n, cudaMemcpyHostToDevice, ctx->stream));
        d_in = staged;
    }
    if (!tree->d_tris) BVH_TRY(dalloc(ctx, &tree->d_tris, sizeof(DTri<T>) * n));
    pack_tris_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_in, (uint32_t)n, reinterpret_cast<DTri<T>*>(tree->d_tris));
    ctx->launches++;
    BVH_CUDA_TRY(cudaGetLastError());
    if (!dev_input) BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));       // the caller's buffer may go away
    return BVHGPU_OK;
}

template <class T>
int closest_hit_device(Tree<T>* tree, const void* d_rays, uint32_t fmt, size_t nrays, int use_triangles, uint32_t* d_shape, T* d_dist, T* d_uv) {
    bvhgpu_ctx* ctx = tree->ctx;
    cudaStream_t st = ctx->stream;
    if (nrays > 0x7FFFFFFFull) { set_error("closest_hit: too many rays"); return BVHGPU_ERR_INVALID; }
    if (fmt != BVHGPU_RAYS_FULL && fmt != BVHGPU_RAYS_OD) { set_error("closest_hit: bad ray layout %u", fmt); return BVHGPU_ERR_INVALID; }
    if (nrays == 0) return BVHGPU_OK;
    BVH_TRY(resolve_status(tree));
    if (tree->n == 0) {                                         // empty tree: no hit
        fill_nohit_kernel<T><<<(unsigned)((nrays + 255) / 256), 256, 0, st>>>((uint32_t)nrays, d_shape, d_dist, d_uv);
        ctx->launches++;
        BVH_CUDA_TRY(cudaGetLastError());
        return BVHGPU_OK;
    }
    if (use_triangles && !tree->d_tris) { set_error("closest_hit: triangle mode needs bvhgpu_tree_set_triangles_* first"); return BVHGPU_ERR_INVALID; }
    const unsigned grid = (unsigned)((nrays + 127) / 128);
    const uint32_t stride = fmt == BVHGPU_RAYS_FULL ? 9u : 6u;
    if (use_triangles)
        closest_kernel<T, true><<<grid, 128, 0, st>>>(tree->d_nodes, tree->n, tree->d_aabb, reinterpret_cast<const DTri<T>*>(tree->d_tris), reinterpret_cast<const T*>(d_rays), stride, (uint32_t)nrays, d_shape, d_dist, d_uv);
    else
        closest_kernel<T, false><<<grid, 128, 0, st>>>(tree->d_nodes, tree->n, tree->d_aabb, nullptr, reinterpret_cast<const T*>(d_rays), stride, (uint32_t)nrays, d_shape, d_dist, d_uv);
    ctx->launches++;
    BVH_CUDA_TRY(cudaGetLastError());
    return BVHGPU_OK;
}

template int set_triangles<float>(Tree<float>*, const float*, size_t, bool);
template int set_triangles<double>(Tree<double>*, const double*, size_t, bool);
template int closest_hit_device<float>(Tree<float>*, const void*, uint32_t, size_t, int, uint32_t*, float*, float*);
template int closest_hit_device<double>(Tree<double>*, const void*, uint32_t, size_t, int, uint32_t*, double*, double*);

}  // namespace bvhb200
