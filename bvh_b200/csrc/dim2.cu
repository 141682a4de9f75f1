// bvh_b200/csrc/dim2.cu -- D = 2 instantiation of build / flatten / traverse (the reference is generic in D and
// ships 2-D slab tests: src/ray/intersect_simd.rs:99-133 (f32), :181-191 (f64); default path src/ray/intersect_default.rs:16-37).
//
// A 2-D scene is run through the 3-D kernels embedded in the plane z = 0, which reproduces the 2-D arithmetic bit for bit:
//   build     every AABB gets z = [0, 0]: centres have z = 0, extents have z = 0, so largest_axis (first strict maximum,
//             aabb_impl.rs:594-596) never picks z; surface_area 2*((sx*sx + sy*sy) + 0) == 2*(sx*sx + sy*sy) exactly
//             (aabb_impl.rs:551-554 with a 2-term dot); buckets, costs and child boxes never see z.  Same topology, same boxes.
//   traverse  the traversal records (and the shape AABBs the FLAT leaf re-test reads) get z = [-1, +1]; rays get origin.z = 0 and
//             inv_direction.z = +inf: the z slab is (-1 - 0) * inf = -inf, (1 - 0) * inf = +inf -- never NaN, and max(tmin, -inf),
//             min(tmax, +inf) are identities, so the 3-D slab test returns exactly what the 2-D one does.
// The 2-D PODs are converted on the device (expand on the way in, drop z on the way out).
#include "internal.h"

namespace bvhb200 {

template <class T> __global__ void __launch_bounds__(256) expand_aabb2_kernel(const T* __restrict__ in /*4 per box*/, uint32_t n, T* __restrict__ out /*6 per box*/) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T* p = in + 4 * (size_t)i;
    T* q = out + 6 * (size_t)i;
    q[0] = p[0]; q[1] = p[1]; q[2] = T(0); q[3] = p[2]; q[4] = p[3]; q[5] = T(0);
}
template <class T> __global__ void __launch_bounds__(256) expand_ray2_kernel(const T* __restrict__ in /*6 per ray*/, uint32_t n, T* __restrict__ out /*9 per ray*/) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T* p = in + 6 * (size_t)i;
    T* q = out + 9 * (size_t)i;
    q[0] = p[0]; q[1] = p[1]; q[2] = T(0);
    q[3] = p[2]; q[4] = p[3]; q[5] = T(0);
    q[6] = p[4]; q[7] = p[5]; q[8] = Traits<T>::inf();
}
// shape AABBs for the FLAT leaf re-test of a 2-D tree: z = [-1, +1]
template <class T> __global__ void __launch_bounds__(256) trav_aabb2_kernel(const typename Traits<T>::DAabb* __restrict__ in, uint32_t n, typename Traits<T>::DAabb* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typename Traits<T>::DAabb d = in[i];
    d.min[2] = T(-1); d.max[2] = T(1);
    out[i] = d;
}
template <class T, class N2> __global__ void __launch_bounds__(256) shrink_nodes_kernel(const typename Traits<T>::Node* __restrict__ in, uint32_t n, N2* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const typename Traits<T>::Node nd = in[i];
    N2 o;
    o.parent = nd.parent; o.child_l = nd.child_l; o.child_r = nd.child_r; o.shape = nd.shape;
    for (int k = 0; k < 2; ++k) { o.l_aabb.min[k] = nd.l_aabb.min[k]; o.l_aabb.max[k] = nd.l_aabb.max[k]; o.r_aabb.min[k] = nd.r_aabb.min[k]; o.r_aabb.max[k] = nd.r_aabb.max[k]; }
    out[i] = o;
}
template <class T, class F2> __global__ void __launch_bounds__(256) shrink_flat_kernel(const typename Traits<T>::Flat* __restrict__ in, size_t n, F2* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const typename Traits<T>::Flat f = in[i];
    F2 o;
    for (int k = 0; k < 2; ++k) { o.aabb.min[k] = f.aabb.min[k]; o.aabb.max[k] = f.aabb.max[k]; }
    o.entry_index = f.entry_index; o.exit_index = f.exit_index; o.shape_index = f.shape_index;
    out[i] = o;
}

template <class T> int dim2_expand_aabbs(bvhgpu_ctx* ctx, const T* d_in4, uint32_t n, T* d_out6) {
    expand_aabb2_kernel<T><<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_in4, n, d_out6);
    ctx->launches++;
    BVH_CUDA_TRY(cudaGetLastError());
    return BVHGPU_OK;
}
template <class T> int dim2_expand_rays(bvhgpu_ctx* ctx, const T* d_in6, uint32_t n, T* d_out9) {
    expand_ray2_kernel<T><<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_in6, n, d_out9);
    ctx->launches++;
    BVH_CUDA_TRY(cudaGetLastError());
    return BVHGPU_OK;
}
template <class T> int dim2_finish_build(Tree<T>* tree) {
    bvhgpu_ctx* ctx = tree->ctx;
    tree->dims = 2;
    if (tree->n == 0) return BVHGPU_OK;
    if (!tree->d_aabb_trav) BVH_TRY(dalloc_t(ctx, &tree->d_aabb_trav, tree->n));
    trav_aabb2_kernel<T><<<(tree->n + 255) / 256, 256, 0, ctx->stream>>>(tree->d_aabb, tree->n, tree->d_aabb_trav);
    ctx->launches++;
    BVH_CUDA_TRY(cudaGetLastError());
    return BVHGPU_OK;
}
template <class T, class N2> int dim2_nodes_out(Tree<T>* tree, N2* d_out) {
    bvhgpu_ctx* ctx = tree->ctx;
    shrink_nodes_kernel<T, N2><<<(tree->n_nodes + 255) / 256, 256, 0, ctx->stream>>>(tree->d_nodes, tree->n_nodes, d_out);
    ctx->launches++;
    BVH_CUDA_TRY(cudaGetLastError());
    return BVHGPU_OK;
}
template <class T, class F2> int dim2_flat_out(Tree<T>* tree, F2* d_out) {
    bvhgpu_ctx* ctx = tree->ctx;
    shrink_flat_kernel<T, F2><<<(unsigned)((tree->n_flat + 255) / 256), 256, 0, ctx->stream>>>(tree->d_flat, tree->n_flat, d_out);
    ctx->launches++;
    BVH_CUDA_TRY(cudaGetLastError());
    return BVHGPU_OK;
}

template int dim2_expand_aabbs<float>(bvhgpu_ctx*, const float*, uint32_t, float*);
template int dim2_expand_aabbs<double>(bvhgpu_ctx*, const double*, uint32_t, double*);
template int dim2_expand_rays<float>(bvhgpu_ctx*, const float*, uint32_t, float*);
template int dim2_expand_rays<double>(bvhgpu_ctx*, const double*, uint32_t, double*);
template int dim2_finish_build<float>(Tree<float>*);
template int dim2_finish_build<double>(Tree<double>*);
template int dim2_nodes_out<float, bvh_node2f>(Tree<float>*, bvh_node2f*);
template int dim2_nodes_out<double, bvh_node2d>(Tree<double>*, bvh_node2d*);
template int dim2_flat_out<float, bvh_flat2f>(Tree<float>*, bvh_flat2f*);
template int dim2_flat_out<double, bvh_flat2d>(Tree<double>*, bvh_flat2d*);

}  // namespace bvhb200
