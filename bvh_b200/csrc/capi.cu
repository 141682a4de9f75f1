// bvh_b200/csrc/capi.cu -- the extern "C" surface declared in include/bvh_b200.h.
#include "internal.h"
#include <cstdarg>
#include <cstring>
#include <vector>
#include <new>
#include <algorithm>
#include <cctype>
#include <sched.h>

namespace bvhb200 {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int dalloc(bvhgpu_ctx* ctx, void** p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) bytes = 16;
    BVH_CUDA_TRY(cudaMallocAsync(p, bytes, ctx->stream));
    return BVHGPU_OK;
}
void dfree(bvhgpu_ctx* ctx, void* p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}

// Resolve the deferred device status of a build (synchronises the stream once).
// A failure is STICKY: the node arrays of a tree whose build / refit / optimize failed are uninitialised or half rewritten,
// so every later entry point on that tree reports the same status again (the reference panicked at this point and the
// Bvh never existed); only bvhgpu_tree_free_* is meaningful afterwards.
template <class T> int resolve_status(Tree<T>* tree) {
    if (tree->failed_status != BVHGPU_OK) { set_error("%s", tree->failed_message.c_str()); return tree->failed_status; }
    if (!tree->status_pending) return BVHGPU_OK;
    bvhgpu_ctx* ctx = tree->ctx;
    BuildStatus h;
    BVH_CUDA_TRY(cudaMemcpyAsync(&h, tree->d_status, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    tree->status_pending = false;
    char msg[512];
    int rc = BVHGPU_OK;
    if (h.nan_found) { snprintf(msg, sizeof msg, "build: NaN coordinate in an input AABB (the reference panics here, src/bvh/bvh_node.rs:214-217); the tree is unusable"); rc = BVHGPU_ERR_NAN; }
    else if (h.error == BVHGPU_ERR_TIMEOUT) { snprintf(msg, sizeof msg, "build: device watchdog fired (tickets=%u leaves=%u/%u); the tree is unusable", h.tickets, h.leaves_done, tree->n); rc = BVHGPU_ERR_TIMEOUT; }
    else if (h.error) { snprintf(msg, sizeof msg, "build: device reported status %u (tickets=%u leaves=%u/%u); the tree is unusable", h.error, h.tickets, h.leaves_done, tree->n); rc = (int)h.error; }
    if (rc != BVHGPU_OK) { tree->failed_status = rc; tree->failed_message = msg; set_error("%s", msg); }
    return rc;
}

template int resolve_status<float>(Tree<float>*);
template int resolve_status<double>(Tree<double>*);

template <class T> static void tree_release(Tree<T>* t) {
    if (!t) return;
    bvhgpu_ctx* ctx = t->ctx;
    if (ctx) {
        dfree(ctx, t->d_aabb); dfree(ctx, t->d_aabb_trav); dfree(ctx, t->d_nodes); dfree(ctx, t->d_node_index); dfree(ctx, t->d_node_start);
        dfree(ctx, t->d_tris); dfree(ctx, t->d_sa_base); dfree(ctx, t->d_arrive); dfree(ctx, t->d_bad); dfree(ctx, t->d_tnodes); dfree(ctx, t->d_top); dfree(ctx, t->d_flat); dfree(ctx, t->d_status); dfree(ctx, t->d_offsets); dfree(ctx, t->d_hits);
    }
}

template <class T, class TreeT>
static int build_impl(bvhgpu_ctx* ctx, const typename Traits<T>::Aabb* aabbs, size_t n, int mode, bool host_input, TreeT** out) {
    if (!ctx || !out || (n && !aabbs)) { set_error("build: null argument"); return BVHGPU_ERR_INVALID; }
    *out = nullptr;
    if (n > (1ull << 30)) { set_error("build: n = %zu exceeds 2^30 shapes (u32 node indices)", n); return BVHGPU_ERR_INVALID; }
    if (mode != BVHGPU_BUILD_EXACT_SAH && mode != BVHGPU_BUILD_LBVH && mode != BVHGPU_BUILD_LBVH_TREELET) { set_error("build: unknown mode %d", mode); return BVHGPU_ERR_INVALID; }
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    TreeT* tree = new (std::nothrow) TreeT();
    if (!tree) { set_error("build: out of host memory"); return BVHGPU_ERR_INTERNAL; }
    tree->ctx = ctx;
    const typename Traits<T>::Aabb* d_in = aabbs;
    typename Traits<T>::Aabb* staged = nullptr;
    int rc = BVHGPU_OK;
    if (host_input && n) {
        rc = dalloc_t(ctx, &staged, n);
        if (rc == BVHGPU_OK) {
            cudaError_t e = cudaMemcpyAsync(staged, aabbs, n * sizeof(*aabbs), cudaMemcpyHostToDevice, ctx->stream);
            if (e != cudaSuccess) { set_error("build: H2D copy failed: %s", cudaGetErrorString(e)); rc = BVHGPU_ERR_CUDA; }
        }
        d_in = staged;
    }
    if (rc == BVHGPU_OK) rc = mode == BVHGPU_BUILD_EXACT_SAH ? build_exact_sah<T>(ctx, d_in, (uint32_t)n, tree) : build_lbvh<T>(ctx, d_in, (uint32_t)n, tree, mode == BVHGPU_BUILD_LBVH_TREELET);
    if (staged) dfree(ctx, staged);
    if (rc == BVHGPU_OK && host_input) rc = resolve_status(tree);     // host entry point reports errors eagerly
    if (rc != BVHGPU_OK) { tree_release(tree); delete tree; return rc; }
    *out = tree;
    return BVHGPU_OK;
}

// Upload of an existing reference-layout Bvh: validate the preorder invariant on the host, derive the
// per-node shape counts / range starts the device kernels rely on.
template <class T, class TreeT>
static int from_nodes_impl(bvhgpu_ctx* ctx, const typename Traits<T>::Node* nodes, size_t n_nodes,
                           const typename Traits<T>::Aabb* aabbs, size_t n, TreeT** out) {
    using Node = typename Traits<T>::Node;
    if (!ctx || !out) { set_error("tree_from_nodes: null argument"); return BVHGPU_ERR_INVALID; }
    *out = nullptr;
    if ((n == 0) != (n_nodes == 0) || (n && n_nodes != 2 * n - 1)) { set_error("tree_from_nodes: n_nodes must be 2n-1"); return BVHGPU_ERR_INVALID; }
    if (n > (1ull << 30)) { set_error("tree_from_nodes: too many shapes"); return BVHGPU_ERR_INVALID; }
    if (n && (!nodes || !aabbs)) { set_error("tree_from_nodes: null argument"); return BVHGPU_ERR_INVALID; }
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    std::vector<Node> fixed(nodes, nodes + n_nodes);
    std::vector<uint32_t> count(n_nodes), start(n_nodes), node_index(n, BVH_INVALID);
    for (size_t ii = n_nodes; ii-- > 0;) {                       // children have larger indices in preorder
        Node& nd = fixed[ii];
        if (nd.child_l == BVH_INVALID) {
            if (nd.shape >= n) { set_error("tree_from_nodes: leaf %zu has shape %u out of range", ii, nd.shape); return BVHGPU_ERR_INVALID; }
            count[ii] = 1;
        } else {
            if (nd.child_l != ii + 1 || nd.child_l >= n_nodes || nd.child_r >= n_nodes || nd.child_r <= nd.child_l) {
                set_error("tree_from_nodes: node %zu is not in the preorder layout Bvh::build emits (child_l must be i+1)", ii);
                return BVHGPU_ERR_UNSUPPORTED;
            }
            if (nd.child_r != ii + 2 * (size_t)count[nd.child_l]) {
                set_error("tree_from_nodes: node %zu: child_r != i + 2*n_l", ii);
                return BVHGPU_ERR_UNSUPPORTED;
            }
            count[ii] = count[nd.child_l] + count[nd.child_r];
            nd.shape = count[ii];
        }
    }
    if (n_nodes && count[0] != n) { set_error("tree_from_nodes: tree does not cover all shapes"); return BVHGPU_ERR_INVALID; }
    if (n_nodes) start[0] = 0;
    for (size_t ii = 0; ii < n_nodes; ++ii) {
        const Node& nd = fixed[ii];
        if (nd.child_l == BVH_INVALID) { node_index[nd.shape] = (uint32_t)ii; continue; }
        start[nd.child_l] = start[ii];
        start[nd.child_r] = start[ii] + count[nd.child_l];
        if (fixed[nd.child_l].parent != ii || fixed[nd.child_r].parent != ii) { set_error("tree_from_nodes: bad parent link under node %zu", ii); return BVHGPU_ERR_INVALID; }
    }
    for (size_t s = 0; s < n; ++s) if (node_index[s] == BVH_INVALID) { set_error("tree_from_nodes: shape %zu is in no leaf", s); return BVHGPU_ERR_INVALID; }

    TreeT* tree = new (std::nothrow) TreeT();
    if (!tree) { set_error("out of host memory"); return BVHGPU_ERR_INTERNAL; }
    tree->ctx = ctx; tree->n = (uint32_t)n; tree->n_nodes = (uint32_t)n_nodes;
    int rc = BVHGPU_OK;
    typename Traits<T>::Aabb* staged = nullptr;
    auto fail = [&](int code) { if (staged) dfree(ctx, staged); tree_release(tree); delete tree; return code; };
    if ((rc = dalloc_t(ctx, &tree->d_status, 1)) != BVHGPU_OK) return fail(rc);
    cudaMemsetAsync(tree->d_status, 0, sizeof(BuildStatus), ctx->stream);
    if (n) {
        if ((rc = dalloc_t(ctx, &tree->d_aabb, n)) != BVHGPU_OK) return fail(rc);
        if ((rc = dalloc_t(ctx, &tree->d_nodes, n_nodes)) != BVHGPU_OK) return fail(rc);
        if ((rc = dalloc_t(ctx, &tree->d_node_index, n)) != BVHGPU_OK) return fail(rc);
        if ((rc = dalloc_t(ctx, &tree->d_node_start, n_nodes)) != BVHGPU_OK) return fail(rc);
        if ((rc = dalloc_t(ctx, &staged, n)) != BVHGPU_OK) return fail(rc);
        cudaMemcpyAsync(staged, aabbs, n * sizeof(*aabbs), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(tree->d_nodes, fixed.data(), n_nodes * sizeof(Node), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(tree->d_node_index, node_index.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(tree->d_node_start, start.data(), n_nodes * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
        if ((rc = convert_aabbs<T>(ctx, staged, (uint32_t)n, tree->d_aabb, &tree->d_status->nan_found)) != BVHGPU_OK) return fail(rc);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);       // the host vectors go out of scope
        if (e != cudaSuccess) { set_error("tree_from_nodes: %s", cudaGetErrorString(e)); return fail(BVHGPU_ERR_CUDA); }
        dfree(ctx, staged);
        staged = nullptr;
        tree->status_pending = true;                              // the NaN flag of convert_aabbs
        if ((rc = resolve_status(tree)) != BVHGPU_OK) return fail(rc);
    }
    *out = tree;
    return BVHGPU_OK;
}

template <class T> static int tree_nodes_impl(Tree<T>* tree, typename Traits<T>::Node* out_nodes, uint32_t* out_node_index) {
    if (!tree) { set_error("tree_nodes: null tree"); return BVHGPU_ERR_INVALID; }
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_TRY(resolve_status(tree));
    if (tree->n == 0) return BVHGPU_OK;
    if (out_nodes) BVH_CUDA_TRY(cudaMemcpyAsync(out_nodes, tree->d_nodes, sizeof(*out_nodes) * tree->n_nodes, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_node_index) BVH_CUDA_TRY(cudaMemcpyAsync(out_node_index, tree->d_node_index, sizeof(uint32_t) * tree->n, cudaMemcpyDeviceToHost, ctx->stream));
    BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return BVHGPU_OK;
}

template <class T> static int flatten_impl(Tree<T>* tree, typename Traits<T>::Flat* out, size_t cap, size_t* len) {
    if (!tree) { set_error("flatten: null tree"); return BVHGPU_ERR_INVALID; }
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    if (out || tree->failed_status != BVHGPU_OK) BVH_TRY(resolve_status(tree));   // never map a node array whose build failed
    BVH_TRY(build_flat(tree));
    if (len) *len = tree->n_flat;
    if (out) {
        if (cap < tree->n_flat) { set_error("flatten: capacity %zu < %zu flat nodes", cap, tree->n_flat); return BVHGPU_ERR_CAPACITY; }
        if (tree->n_flat) {
            BVH_CUDA_TRY(cudaMemcpyAsync(out, tree->d_flat, sizeof(*out) * tree->n_flat, cudaMemcpyDeviceToHost, ctx->stream));
            BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
    }
    return BVHGPU_OK;
}

template <class T> static int ensure_result_buffers(Tree<T>* tree, size_t nrays, size_t hits_cap) {
    bvhgpu_ctx* ctx = tree->ctx;
    if (tree->offsets_cap < nrays + 1) {
        dfree(ctx, tree->d_offsets); tree->d_offsets = nullptr; tree->offsets_cap = 0;
        BVH_TRY(dalloc_t(ctx, &tree->d_offsets, nrays + 1));
        tree->offsets_cap = nrays + 1;
    }
    if (tree->hits_cap < hits_cap) {
        dfree(ctx, tree->d_hits); tree->d_hits = nullptr; tree->hits_cap = 0;
        BVH_TRY(dalloc_t(ctx, &tree->d_hits, hits_cap));
        tree->hits_cap = hits_cap;
    }
    return BVHGPU_OK;
}

template <class T>
static int traverse_host_impl(Tree<T>* tree, int mode, const void* rays, uint32_t fmt, size_t nrays,
                              uint32_t* offsets, uint32_t* hits, size_t cap, size_t* total) {
    if (!tree || (nrays && !rays) || !offsets) { set_error("traverse: null argument"); return BVHGPU_ERR_INVALID; }
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_TRY(resolve_status(tree));
    if (nrays == 0 || tree->n == 0) {                            // nothing to pipeline
        size_t tot0 = 0;
        BVH_TRY(ensure_result_buffers(tree, nrays, 1024));
        BVH_TRY(traverse_device<T>(tree, mode, nullptr, fmt, nrays, tree->d_offsets, tree->d_hits, tree->hits_cap, &tot0));
        BVH_CUDA_TRY(cudaMemcpyAsync(offsets, tree->d_offsets, sizeof(uint32_t) * (nrays + 1), cudaMemcpyDeviceToHost, ctx->stream));
        BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (total) *total = 0;
        return BVHGPU_OK;
    }
    size_t want = std::max<size_t>(std::max<size_t>(tree->hits_cap, 4 * nrays), 1024);
    size_t tot = 0;
    int rc = BVHGPU_OK;
    for (int attempt = 0; attempt < 2; ++attempt) {
        rc = ensure_result_buffers(tree, nrays, want);
        if (rc != BVHGPU_OK) break;
        rc = traverse_host_pipelined<T>(tree, mode, rays, fmt, nrays, offsets, hits, cap, &tot);
        if (rc == BVHGPU_ERR_CAPACITY && tot <= 0xFFFFFFFFull && tot > tree->hits_cap && attempt == 0) { want = tot; continue; }   // grow once and redo
        break;
    }
    if (total) *total = tot;
    if (rc != BVHGPU_OK) return rc;
    if (tot > cap) {
        set_error("traverse: %zu hits do not fit the caller's capacity %zu (use bvhgpu_traverse_fetch_*)", tot, cap);
        return BVHGPU_ERR_CAPACITY;
    }
    return BVHGPU_OK;
}

template <class T>
static int query_host_impl(Tree<T>* tree, int mode, int kind, const T* queries, size_t n, uint32_t* offsets, uint32_t* hits, size_t cap, size_t* total) {
    if (!tree || (n && !queries) || !offsets) { set_error("query: null argument"); return BVHGPU_ERR_INVALID; }
    if (kind < BVHGPU_QUERY_AABB || kind > BVHGPU_QUERY_BALL) { set_error("query: bad kind %d", kind); return BVHGPU_ERR_INVALID; }
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_TRY(resolve_status(tree));
    const size_t stride = kind == BVHGPU_QUERY_AABB ? 6 : (kind == BVHGPU_QUERY_POINT ? 3 : 4);
    T* d_q = nullptr;
    Scratch scratch(ctx);                                           // released on every return path
    if (n) {
        BVH_TRY(scratch.get(&d_q, n * stride));
        BVH_CUDA_TRY(cudaMemcpyAsync(d_q, queries, sizeof(T) * n * stride, cudaMemcpyHostToDevice, ctx->stream));
    }
    size_t want = std::max<size_t>(std::max<size_t>(tree->hits_cap, 16 * n), 1024), tot = 0;
    int rc = BVHGPU_OK;
    for (int attempt = 0; attempt < 2; ++attempt) {
        rc = ensure_result_buffers(tree, n, want);
        if (rc != BVHGPU_OK) break;
        rc = query_device<T>(tree, mode, kind, d_q, n, tree->d_offsets, tree->d_hits, tree->hits_cap, &tot);
        if (rc == BVHGPU_ERR_CAPACITY && tot <= 0xFFFFFFFFull && attempt == 0) { want = tot; continue; }
        break;
    }
    if (total) *total = tot;
    if (rc != BVHGPU_OK) return rc;
    BVH_CUDA_TRY(cudaMemcpyAsync(offsets, tree->d_offsets, sizeof(uint32_t) * (n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    int ret = BVHGPU_OK;
    if (hits && tot <= cap) { if (tot) BVH_CUDA_TRY(cudaMemcpyAsync(hits, tree->d_hits, sizeof(uint32_t) * tot, cudaMemcpyDeviceToHost, ctx->stream)); }
    else if (tot > cap) { set_error("query: %zu hits do not fit the caller's capacity %zu (use bvhgpu_traverse_fetch_*)", tot, cap); ret = BVHGPU_ERR_CAPACITY; }
    BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return ret;
}

template <class T>
static int nearest_host_impl(Tree<T>* tree, int mode, const T* points, size_t n, uint32_t* out_shape, T* out_dist, int use_triangles = 0) {
    if (!tree || (n && (!points || !out_shape || !out_dist))) { set_error("nearest: null argument"); return BVHGPU_ERR_INVALID; }
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_TRY(resolve_status(tree));
    if (n == 0) return BVHGPU_OK;
    T *d_p = nullptr, *d_d = nullptr;
    uint32_t* d_s = nullptr;
    Scratch scratch(ctx);
    BVH_TRY(scratch.get(&d_p, n * 3));
    BVH_TRY(scratch.get(&d_d, n));
    BVH_TRY(scratch.get(&d_s, n));
    BVH_CUDA_TRY(cudaMemcpyAsync(d_p, points, sizeof(T) * n * 3, cudaMemcpyHostToDevice, ctx->stream));
    int rc = nearest_device<T>(tree, mode, d_p, n, d_s, d_d, use_triangles);
    if (rc == BVHGPU_OK) {
        BVH_CUDA_TRY(cudaMemcpyAsync(out_shape, d_s, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
        BVH_CUDA_TRY(cudaMemcpyAsync(out_dist, d_d, sizeof(T) * n, cudaMemcpyDeviceToHost, ctx->stream));
        BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return rc;
}

template <class T>
static int nearest_candidates_host_impl(Tree<T>* tree, const T* points, size_t n, uint32_t* offsets, uint32_t* cand, size_t cap, size_t* total) {
    if (!tree || (n && !points) || !offsets) { set_error("nearest_candidates: null argument"); return BVHGPU_ERR_INVALID; }
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_TRY(resolve_status(tree));
    T* d_p = nullptr;
    Scratch scratch(ctx);
    if (n) {
        BVH_TRY(scratch.get(&d_p, n * 3));
        BVH_CUDA_TRY(cudaMemcpyAsync(d_p, points, sizeof(T) * n * 3, cudaMemcpyHostToDevice, ctx->stream));
    }
    size_t want = std::max<size_t>(std::max<size_t>(tree->hits_cap, 16 * n), 1024), tot = 0;
    int rc = BVHGPU_OK;
    for (int attempt = 0; attempt < 2; ++attempt) {
        rc = ensure_result_buffers(tree, n, want);
        if (rc != BVHGPU_OK) break;
        rc = nearest_candidates_device<T>(tree, d_p, n, tree->d_offsets, tree->d_hits, tree->hits_cap, &tot);
        if (rc == BVHGPU_ERR_CAPACITY && tot <= 0xFFFFFFFFull && attempt == 0) { want = tot; continue; }
        break;
    }
    if (total) *total = tot;
    if (rc != BVHGPU_OK) return rc;
    BVH_CUDA_TRY(cudaMemcpyAsync(offsets, tree->d_offsets, sizeof(uint32_t) * (n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    int ret = BVHGPU_OK;
    if (cand && tot <= cap) { if (tot) BVH_CUDA_TRY(cudaMemcpyAsync(cand, tree->d_hits, sizeof(uint32_t) * tot, cudaMemcpyDeviceToHost, ctx->stream)); }
    else if (tot > cap) { set_error("nearest_candidates: %zu candidates do not fit the caller's capacity %zu (use bvhgpu_traverse_fetch_*)", tot, cap); ret = BVHGPU_ERR_CAPACITY; }
    BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return ret;
}

template <class T>
static int ordered_host_impl(Tree<T>* tree, const typename Traits<T>::Ray* rays, size_t nrays, int ascending,
                             uint32_t* offsets, uint32_t* hits, T* dists, size_t cap, size_t* total) {
    if (!tree || (nrays && !rays) || !offsets || (cap && (!hits || !dists))) { set_error("traverse_ordered: null argument"); return BVHGPU_ERR_INVALID; }
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_TRY(resolve_status(tree));
    typename Traits<T>::Ray* d_rays = nullptr;
    uint32_t *d_off = nullptr, *d_hits = nullptr;
    T* d_dists = nullptr;
    Scratch scratch(ctx);
    BVH_TRY(scratch.get(&d_rays, nrays));
    BVH_TRY(scratch.get(&d_off, nrays + 1));
    BVH_TRY(scratch.get(&d_hits, cap));
    BVH_TRY(scratch.get(&d_dists, cap));
    if (nrays) BVH_CUDA_TRY(cudaMemcpyAsync(d_rays, rays, sizeof(*rays) * nrays, cudaMemcpyHostToDevice, ctx->stream));
    size_t tot = 0;
    int rc = traverse_ordered_device<T>(tree, d_rays, nrays, ascending, d_off, d_hits, d_dists, cap, &tot);
    if (total) *total = tot;
    if (rc == BVHGPU_OK || rc == BVHGPU_ERR_CAPACITY) {
        cudaMemcpyAsync(offsets, d_off, sizeof(uint32_t) * (nrays + 1), cudaMemcpyDeviceToHost, ctx->stream);
        const size_t m = std::min(tot, cap);
        if (m) { cudaMemcpyAsync(hits, d_hits, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, ctx->stream); cudaMemcpyAsync(dists, d_dists, sizeof(T) * m, cudaMemcpyDeviceToHost, ctx->stream); }
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error("traverse_ordered: %s", cudaGetErrorString(e)); rc = BVHGPU_ERR_CUDA; }
    }
    return rc;
}

template <class T>
static int closest_host_impl(Tree<T>* tree, const void* rays, uint32_t fmt, size_t nrays, int use_triangles, uint32_t* out_shape, T* out_dist, T* out_uv) {
    if (!tree || (nrays && (!rays || !out_shape || !out_dist))) { set_error("closest_hit: null argument"); return BVHGPU_ERR_INVALID; }
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_TRY(resolve_status(tree));
    if (nrays == 0) return BVHGPU_OK;
    Scratch scratch(ctx);
    const size_t ray_bytes = (fmt == BVHGPU_RAYS_FULL ? 9 : 6) * sizeof(T);
    unsigned char* d_rays = nullptr;
    uint32_t* d_s = nullptr;
    T *d_d = nullptr, *d_uv = nullptr;
    BVH_TRY(scratch.get(&d_rays, ray_bytes * nrays));
    BVH_TRY(scratch.get(&d_s, nrays));
    BVH_TRY(scratch.get(&d_d, nrays));
    if (out_uv) BVH_TRY(scratch.get(&d_uv, 2 * nrays));
    BVH_CUDA_TRY(cudaMemcpyAsync(d_rays, rays, ray_bytes * nrays, cudaMemcpyHostToDevice, ctx->stream));
    BVH_TRY(closest_hit_device<T>(tree, d_rays, fmt, nrays, use_triangles, d_s, d_d, d_uv));
    BVH_CUDA_TRY(cudaMemcpyAsync(out_shape, d_s, sizeof(uint32_t) * nrays, cudaMemcpyDeviceToHost, ctx->stream));
    BVH_CUDA_TRY(cudaMemcpyAsync(out_dist, d_d, sizeof(T) * nrays, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_uv) BVH_CUDA_TRY(cudaMemcpyAsync(out_uv, d_uv, sizeof(T) * 2 * nrays, cudaMemcpyDeviceToHost, ctx->stream));
    BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return BVHGPU_OK;
}

template <class T> static int fetch_impl(Tree<T>* tree, uint32_t* hits, size_t cap) {
    if (!tree || !hits) { set_error("traverse_fetch: null argument"); return BVHGPU_ERR_INVALID; }
    if (cap < tree->last_total) { set_error("traverse_fetch: capacity %zu < %zu hits", cap, tree->last_total); return BVHGPU_ERR_CAPACITY; }
    if (tree->last_total == 0) return BVHGPU_OK;
    if (!tree->d_hits || tree->hits_cap < tree->last_total) { set_error("traverse_fetch: no retained result"); return BVHGPU_ERR_INVALID; }
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_CUDA_TRY(cudaMemcpyAsync(hits, tree->d_hits, sizeof(uint32_t) * tree->last_total, cudaMemcpyDeviceToHost, ctx->stream));
    BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return BVHGPU_OK;
}

// New shape AABBs for refit / optimize: converted into a TEMPORARY device array and checked for NaN before the tree is touched,
// so a rejected update (BVHGPU_ERR_NAN) leaves the tree exactly as it was.  dev_input: `aabbs` is a device pointer.
template <class T>
static int stage_new_aabbs(Tree<T>* tree, const typename Traits<T>::Aabb* aabbs, size_t n, bool dev_input, const char* who) {
    bvhgpu_ctx* ctx = tree->ctx;
    Scratch scratch(ctx);
    typename Traits<T>::Aabb* staged = nullptr;
    const typename Traits<T>::Aabb* d_in = aabbs;
    if (!dev_input) {
        BVH_TRY(scratch.get(&staged, n));
        BVH_CUDA_TRY(cudaMemcpyAsync(staged, aabbs, n * sizeof(*aabbs), cudaMemcpyHostToDevice, ctx->stream));
        d_in = staged;
    }
    typename Traits<T>::DAabb* fresh = nullptr;
    uint32_t* flag = nullptr;
    BVH_TRY(dalloc_t(ctx, &fresh, n));
    int rc = scratch.get(&flag, 1);
    if (rc == BVHGPU_OK && cudaMemsetAsync(flag, 0, sizeof(uint32_t), ctx->stream) != cudaSuccess) rc = BVHGPU_ERR_CUDA;
    if (rc == BVHGPU_OK) rc = convert_aabbs<T>(ctx, d_in, (uint32_t)n, fresh, flag);
    uint32_t* h = ctx->h_pinned + 200;
    if (rc == BVHGPU_OK && (cudaMemcpyAsync(h, flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                            cudaStreamSynchronize(ctx->stream) != cudaSuccess)) { set_error("%s: CUDA error while staging the AABBs", who); rc = BVHGPU_ERR_CUDA; }
    if (rc == BVHGPU_OK && *h) { set_error("%s: NaN coordinate in an input AABB; the tree was left unchanged", who); rc = BVHGPU_ERR_NAN; }
    if (rc != BVHGPU_OK) { dfree(ctx, fresh); return rc; }
    dfree(ctx, tree->d_aabb);
    tree->d_aabb = fresh;
    return BVHGPU_OK;
}

template <class T>
static int refit_impl(Tree<T>* tree, const typename Traits<T>::Aabb* aabbs, size_t n, bool dev_input) {
    if (!tree || (n && !aabbs)) { set_error("refit: null argument"); return BVHGPU_ERR_INVALID; }
    if (n != tree->n) { set_error("refit: %zu AABBs for a tree over %u shapes", n, tree->n); return BVHGPU_ERR_INVALID; }
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_TRY(resolve_status(tree));
    if (n == 0) return BVHGPU_OK;
    BVH_TRY(stage_new_aabbs<T>(tree, aabbs, n, dev_input, "refit"));
    BVH_CUDA_TRY(cudaMemsetAsync(tree->d_status, 0, sizeof(BuildStatus), ctx->stream));
    BVH_TRY(refit(tree));
    tree->status_pending = true;
    return dev_input ? (int)BVHGPU_OK : resolve_status(tree);
}

template <class T>
static int optimize_impl(Tree<T>* tree, const typename Traits<T>::Aabb* aabbs, size_t n, double max_growth, size_t* rebuilt, bool dev_input) {
    if (!tree || (n && !aabbs)) { set_error("optimize: null argument"); return BVHGPU_ERR_INVALID; }
    if (n != tree->n) { set_error("optimize: %zu AABBs for a tree over %u shapes", n, tree->n); return BVHGPU_ERR_INVALID; }
    if (!(max_growth >= 1.0)) { set_error("optimize: max_growth = %g, must be >= 1", max_growth); return BVHGPU_ERR_INVALID; }
    if (rebuilt) *rebuilt = 0;
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_TRY(resolve_status(tree));
    if (n == 0) return BVHGPU_OK;
    BVH_TRY(stage_new_aabbs<T>(tree, aabbs, n, dev_input, "optimize"));
    BVH_CUDA_TRY(cudaMemsetAsync(tree->d_status, 0, sizeof(BuildStatus), ctx->stream));
    BVH_TRY(optimize(tree, max_growth));
    tree->status_pending = true;
    if (dev_input && !rebuilt) return BVHGPU_OK;                        // asynchronous: errors surface at the next call on the tree
    BuildStatus h;
    BVH_CUDA_TRY(cudaMemcpyAsync(&h, tree->d_status, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    BVH_TRY(resolve_status(tree));                                      // synchronises
    if (rebuilt) *rebuilt = h.rebuilt;
    return BVHGPU_OK;
}

// ---- D = 2 (dim2.cu): host PODs in, converted on the device, run through the 3-D path --------------------------------------
template <class T, class TREE2, class AABB2>
static int build2_impl(bvhgpu_ctx* ctx, const AABB2* aabbs, size_t n, int mode, TREE2** out) {
    if (!ctx || !out || (n && !aabbs)) { set_error("build: null argument"); return BVHGPU_ERR_INVALID; }
    *out = nullptr;
    if (n > (1ull << 30)) { set_error("build: n = %zu exceeds 2^30 shapes", n); return BVHGPU_ERR_INVALID; }
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    Scratch scratch(ctx);
    T *in4 = nullptr, *in6 = nullptr;
    if (n) {
        BVH_TRY(scratch.get(&in4, 4 * n));
        BVH_TRY(scratch.get(&in6, 6 * n));
        BVH_CUDA_TRY(cudaMemcpyAsync(in4, aabbs, sizeof(AABB2) * n, cudaMemcpyHostToDevice, ctx->stream));
        BVH_TRY(dim2_expand_aabbs<T>(ctx, in4, (uint32_t)n, in6));
    }
    TREE2* tree = nullptr;
    BVH_TRY((build_impl<T, TREE2>(ctx, reinterpret_cast<const typename Traits<T>::Aabb*>(in6), n, mode, false, &tree)));
    int rc = dim2_finish_build<T>(tree);
    if (rc == BVHGPU_OK) rc = resolve_status(tree);
    if (rc != BVHGPU_OK) { tree_release<T>(tree); delete tree; return rc; }
    *out = tree;
    return BVHGPU_OK;
}
template <class T, class N2>
static int tree_nodes2_impl(Tree<T>* tree, N2* out_nodes, uint32_t* out_node_index) {
    if (!tree) { set_error("tree_nodes: null tree"); return BVHGPU_ERR_INVALID; }
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_TRY(resolve_status(tree));
    if (tree->n == 0) return BVHGPU_OK;
    Scratch scratch(ctx);
    if (out_nodes) {
        N2* d = nullptr;
        BVH_TRY(scratch.get(&d, tree->n_nodes));
        BVH_TRY((dim2_nodes_out<T, N2>(tree, d)));
        BVH_CUDA_TRY(cudaMemcpyAsync(out_nodes, d, sizeof(N2) * tree->n_nodes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (out_node_index) BVH_CUDA_TRY(cudaMemcpyAsync(out_node_index, tree->d_node_index, sizeof(uint32_t) * tree->n, cudaMemcpyDeviceToHost, ctx->stream));
    BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return BVHGPU_OK;
}
template <class T, class F2>
static int flatten2_impl(Tree<T>* tree, F2* out, size_t cap, size_t* len) {
    if (!tree) { set_error("flatten: null tree"); return BVHGPU_ERR_INVALID; }
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_TRY(resolve_status(tree));
    BVH_TRY(build_flat(tree));
    if (len) *len = tree->n_flat;
    if (out) {
        if (cap < tree->n_flat) { set_error("flatten: capacity %zu < %zu flat nodes", cap, tree->n_flat); return BVHGPU_ERR_CAPACITY; }
        if (tree->n_flat) {
            Scratch scratch(ctx);
            F2* d = nullptr;
            BVH_TRY(scratch.get(&d, tree->n_flat));
            BVH_TRY((dim2_flat_out<T, F2>(tree, d)));
            BVH_CUDA_TRY(cudaMemcpyAsync(out, d, sizeof(F2) * tree->n_flat, cudaMemcpyDeviceToHost, ctx->stream));
            BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
    }
    return BVHGPU_OK;
}
template <class T, class RAY2>
static int traverse2_impl(Tree<T>* tree, int mode, const RAY2* rays, size_t nrays, uint32_t* offsets, uint32_t* hits, size_t cap, size_t* total) {
    if (!tree || (nrays && !rays) || !offsets) { set_error("traverse: null argument"); return BVHGPU_ERR_INVALID; }
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_TRY(resolve_status(tree));
    if (nrays > 0x7FFFFFFFull) { set_error("traverse: too many rays"); return BVHGPU_ERR_INVALID; }
    Scratch scratch(ctx);
    T *r6 = nullptr, *r9 = nullptr;
    if (nrays) {
        BVH_TRY(scratch.get(&r6, 6 * nrays));
        BVH_TRY(scratch.get(&r9, 9 * nrays));
        BVH_CUDA_TRY(cudaMemcpyAsync(r6, rays, sizeof(RAY2) * nrays, cudaMemcpyHostToDevice, ctx->stream));
        BVH_TRY(dim2_expand_rays<T>(ctx, r6, (uint32_t)nrays, r9));
    }
    size_t want = std::max<size_t>(std::max<size_t>(tree->hits_cap, 4 * nrays), 1024), tot = 0;
    int rc = BVHGPU_OK;
    for (int attempt = 0; attempt < 2; ++attempt) {
        rc = ensure_result_buffers(tree, nrays, want);
        if (rc != BVHGPU_OK) break;
        rc = traverse_device<T>(tree, mode, r9, BVHGPU_RAYS_FULL, nrays, tree->d_offsets, tree->d_hits, tree->hits_cap, &tot);
        if (rc == BVHGPU_ERR_CAPACITY && tot <= 0xFFFFFFFFull && tot > tree->hits_cap && attempt == 0) { want = tot; continue; }
        break;
    }
    if (total) *total = tot;
    if (rc != BVHGPU_OK) return rc;
    BVH_CUDA_TRY(cudaMemcpyAsync(offsets, tree->d_offsets, sizeof(uint32_t) * (nrays + 1), cudaMemcpyDeviceToHost, ctx->stream));
    int ret = BVHGPU_OK;
    if (hits && tot <= cap) { if (tot) BVH_CUDA_TRY(cudaMemcpyAsync(hits, tree->d_hits, sizeof(uint32_t) * tot, cudaMemcpyDeviceToHost, ctx->stream)); }
    else if (tot > cap) { set_error("traverse: %zu hits do not fit the caller's capacity %zu", tot, cap); ret = BVHGPU_ERR_CAPACITY; }
    BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return ret;
}

// Bvh::update_shapes(changed_shape_indices, shapes): only the m changed shapes cross the boundary.  The tree is touched only after
// the new AABBs passed the NaN / index check.  max_growth <= 0: refit only (topology kept).
template <class T>
static int update_impl(Tree<T>* tree, const uint32_t* changed, const typename Traits<T>::Aabb* fresh, size_t m, double max_growth, size_t* rebuilt, bool dev_input) {
    if (!tree || (m && (!changed || !fresh))) { set_error("update: null argument"); return BVHGPU_ERR_INVALID; }
    if (max_growth > 0.0 && !(max_growth >= 1.0)) { set_error("update: max_growth = %g, must be >= 1 (or <= 0 for a pure refit)", max_growth); return BVHGPU_ERR_INVALID; }
    if (m > 0xFFFFFFFFull) { set_error("update: too many changed shapes"); return BVHGPU_ERR_INVALID; }
    if (rebuilt) *rebuilt = 0;
    bvhgpu_ctx* ctx = tree->ctx;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_TRY(resolve_status(tree));
    if (m == 0 || tree->n == 0) return BVHGPU_OK;
    Scratch scratch(ctx);
    const uint32_t* d_changed = changed;
    const typename Traits<T>::Aabb* d_fresh = fresh;
    if (!dev_input) {
        uint32_t* c = nullptr;
        typename Traits<T>::Aabb* f = nullptr;
        BVH_TRY(scratch.get(&c, m));
        BVH_TRY(scratch.get(&f, m));
        BVH_CUDA_TRY(cudaMemcpyAsync(c, changed, sizeof(uint32_t) * m, cudaMemcpyHostToDevice, ctx->stream));
        BVH_CUDA_TRY(cudaMemcpyAsync(f, fresh, sizeof(*fresh) * m, cudaMemcpyHostToDevice, ctx->stream));
        d_changed = c; d_fresh = f;
    }
    uint32_t* flags = nullptr;
    BVH_TRY(scratch.get(&flags, 2));
    BVH_CUDA_TRY(cudaMemsetAsync(flags, 0, 2 * sizeof(uint32_t), ctx->stream));
    BVH_TRY(update_changed<T>(tree, d_changed, d_fresh, (uint32_t)m, flags));
    uint32_t* h = ctx->h_pinned + 208;
    BVH_CUDA_TRY(cudaMemcpyAsync(h, flags, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h[1]) { set_error("update: a changed shape index is >= %u; the tree was left unchanged", tree->n); return BVHGPU_ERR_INVALID; }
    if (h[0]) { set_error("update: NaN coordinate in a new AABB; the tree was left unchanged"); return BVHGPU_ERR_NAN; }
    BVH_TRY(update_scatter<T>(tree, d_changed, d_fresh, (uint32_t)m));
    BVH_CUDA_TRY(cudaMemsetAsync(tree->d_status, 0, sizeof(BuildStatus), ctx->stream));
    BVH_TRY(update_incremental<T>(tree, d_changed, (uint32_t)m, max_growth));      // touches the root paths of the changed leaves only
    tree->status_pending = true;
    if (dev_input && !rebuilt) return BVHGPU_OK;                        // asynchronous from here on
    BuildStatus hs;
    BVH_CUDA_TRY(cudaMemcpyAsync(&hs, tree->d_status, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
    BVH_TRY(resolve_status(tree));
    if (rebuilt) *rebuilt = hs.rebuilt;
    return BVHGPU_OK;
}

}  // namespace bvhb200

using namespace bvhb200;

#define BVH_EXPORT extern "C" __attribute__((visibility("default")))

BVH_EXPORT const char* bvhgpu_last_error(void) { return g_last_error.c_str(); }
BVH_EXPORT const char* bvhgpu_version(void) { return "bvh_b200 0.1.0 (sm_100a)"; }

BVH_EXPORT int bvhgpu_create(int device, bvhgpu_ctx** out) {
    if (!out) { set_error("create: null out"); return BVHGPU_ERR_INVALID; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("create: no CUDA device (%s); libbvh_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return BVHGPU_ERR_CUDA;
    }
    if (device < 0 || device >= count) { set_error("create: device %d out of range (0..%d)", device, count - 1); return BVHGPU_ERR_INVALID; }
    BVH_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    BVH_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { set_error("create: device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor); return BVHGPU_ERR_UNSUPPORTED; }
    bvhgpu_ctx* ctx = new (std::nothrow) bvhgpu_ctx();
    if (!ctx) { set_error("create: out of host memory"); return BVHGPU_ERR_INTERNAL; }
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    BVH_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    BVH_CUDA_TRY(cudaMallocHost((void**)&ctx->h_pinned, 256 * sizeof(uint32_t)));
    for (int i = 0; i < 2; ++i) { BVH_CUDA_TRY(cudaEventCreate(&ctx->ev_walk[i])); BVH_CUDA_TRY(cudaEventCreate(&ctx->ev_build[i])); }
    BVH_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    BVH_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    for (unsigned i = 0; i < BVH_MAX_CHUNKS; ++i) BVH_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_emit[i], cudaEventDisableTiming));
    for (int i = 0; i < 5; ++i) BVH_CUDA_TRY(cudaEventCreate(&ctx->ev_e2e[i]));
    BVH_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming));
    BVH_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_total, cudaEventDisableTiming));
    for (unsigned i = 0; i < BVH_MAX_CHUNKS; ++i) BVH_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming));
    BVH_CUDA_TRY(cudaMalloc((void**)&ctx->d_async_err, 256));
    BVH_CUDA_TRY(cudaMemset(ctx->d_async_err, 0, 256));
    ctx->d_ready = ctx->d_async_err + 32;                         // its own 128-byte line of the same small allocation
    {   // NUMA node of the device (bvhgpu_host_alloc): /sys/bus/pci/devices/<domain:bus:dev.fn>/numa_node
        char bus[32] = {0}, path[128];
        if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) == cudaSuccess) {
            for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
            snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
            if (FILE* f = fopen(path, "r")) { int nn = -1; if (fscanf(f, "%d", &nn) == 1) ctx->numa_node = nn; fclose(f); }
        }
    }
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;                       // keep freed blocks cached: alloc/free pairs stay cheap
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = ctx;
    return BVHGPU_OK;
}
BVH_EXPORT void bvhgpu_destroy(bvhgpu_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    for (unsigned i = 0; i < BVH_MAX_CHUNKS; ++i) if (ctx->ev_emit[i]) cudaEventDestroy(ctx->ev_emit[i]);
    if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
    if (ctx->ev_total) cudaEventDestroy(ctx->ev_total);
    for (unsigned i = 0; i < BVH_MAX_CHUNKS; ++i) if (ctx->ev_chunk[i]) cudaEventDestroy(ctx->ev_chunk[i]);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->d_async_err) cudaFree(ctx->d_async_err);
    for (int i = 0; i < 5; ++i) if (ctx->ev_e2e[i]) cudaEventDestroy(ctx->ev_e2e[i]);
    for (int i = 0; i < 2; ++i) { if (ctx->ev_walk[i]) cudaEventDestroy(ctx->ev_walk[i]); if (ctx->ev_build[i]) cudaEventDestroy(ctx->ev_build[i]); }
    delete ctx;
}
BVH_EXPORT int bvhgpu_set_stream(bvhgpu_ctx* ctx, void* cuda_stream) {
    if (!ctx) { set_error("set_stream: null ctx"); return BVHGPU_ERR_INVALID; }
    ctx->stream = (cudaStream_t)cuda_stream;
    return BVHGPU_OK;
}
BVH_EXPORT int bvhgpu_reset_stream(bvhgpu_ctx* ctx) {
    if (!ctx) { set_error("reset_stream: null ctx"); return BVHGPU_ERR_INVALID; }
    ctx->stream = ctx->own_stream;
    return BVHGPU_OK;
}
BVH_EXPORT int bvhgpu_synchronize(bvhgpu_ctx* ctx) {
    if (!ctx) { set_error("synchronize: null ctx"); return BVHGPU_ERR_INVALID; }
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    uint32_t* h = ctx->h_pinned + 204;
    BVH_CUDA_TRY(cudaMemcpyAsync(h, ctx->d_async_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (*h) {                                                     // raised by an asynchronous multi-GPU step; reported once
        const int rc = (int)*h;
        cudaMemsetAsync(ctx->d_async_err, 0, sizeof(uint32_t), ctx->stream);
        set_error("sharded traversal: a peer did not answer within the exchange time-out (status %d); the step's result is invalid", rc);
        return rc;
    }
    return BVHGPU_OK;
}
BVH_EXPORT uint64_t bvhgpu_launch_count(const bvhgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }
BVH_EXPORT int bvhgpu_set_option(bvhgpu_ctx* ctx, const char* name, int64_t value) {
    if (!ctx || !name) { set_error("set_option: null argument"); return BVHGPU_ERR_INVALID; }
    if (!strcmp(name, "traverse_slots")) { ctx->traverse_slots = value; return BVHGPU_OK; }
    if (!strcmp(name, "profile")) { ctx->profile = value; return BVHGPU_OK; }
    if (!strcmp(name, "build_small")) { ctx->build_small = value; return BVHGPU_OK; }
    if (!strcmp(name, "build_gang")) { ctx->build_gang = value; return BVHGPU_OK; }
    if (!strcmp(name, "build_subtree")) { ctx->build_subtree = value; return BVHGPU_OK; }
    if (!strcmp(name, "traverse_persistent")) { ctx->traverse_persistent = value; return BVHGPU_OK; }
    if (!strcmp(name, "traverse_stream")) { ctx->traverse_stream = value; return BVHGPU_OK; }
    if (!strcmp(name, "traverse_top")) { ctx->traverse_top = value; return BVHGPU_OK; }
    if (!strcmp(name, "walk_grid")) { ctx->walk_grid = value <= 0 ? 0 : (int)value; ctx->walk_grid_forced = value > 0; return BVHGPU_OK; }   // CTAs of the persistent walk (0 = one full wave)
    set_error("set_option: unknown option '%s'", name);
    return BVHGPU_ERR_INVALID;
}

BVH_EXPORT int bvhgpu_get_metric(bvhgpu_ctx* ctx, const char* name, double* out) {
    if (!ctx || !name || !out) { set_error("get_metric: null argument"); return BVHGPU_ERR_INVALID; }
    cudaEvent_t* ev = nullptr;
    if (!strncmp(name, "e2e_host_us_", 12) && name[12] >= '0' && name[12] <= '7' && !name[13]) { *out = ctx->host_us[name[12] - '0']; return BVHGPU_OK; }
    if (!strcmp(name, "stream_write_value")) { *out = (double)ctx->wv_ok; return BVHGPU_OK; }
    if (!strcmp(name, "host_streamed")) { *out = (double)ctx->last_streamed; return BVHGPU_OK; }
    if (!strcmp(name, "numa_node")) { *out = (double)ctx->numa_node; return BVHGPU_OK; }
    if (!strcmp(name, "walk_ms") && ctx->have_walk) ev = ctx->ev_walk;
    else if (!strcmp(name, "build_ms") && ctx->have_build) ev = ctx->ev_build;
    if (!ev && ctx->have_e2e && !strncmp(name, "e2e_", 4)) {      // e2e_walk_ms / e2e_h2d_ms / e2e_emit_ms / e2e_d2h_ms: since call start
        int k = !strcmp(name, "e2e_walk_ms") ? 1 : !strcmp(name, "e2e_h2d_ms") ? 2 : !strcmp(name, "e2e_emit_ms") ? 3 : !strcmp(name, "e2e_d2h_ms") ? 4 : 0;
        if (k) {
            BVH_CUDA_TRY(cudaEventSynchronize(ctx->ev_e2e[k]));
            float ms = 0.f;
            BVH_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_e2e[0], ctx->ev_e2e[k]));
            *out = (double)ms;
            return BVHGPU_OK;
        }
    }
    if (!ev) { set_error("get_metric: '%s' not recorded (set option profile=1 and run the call first)", name); return BVHGPU_ERR_INVALID; }
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_CUDA_TRY(cudaEventSynchronize(ev[1]));
    float ms = 0.f;
    BVH_CUDA_TRY(cudaEventElapsedTime(&ms, ev[0], ev[1]));
    *out = (double)ms;
    return BVHGPU_OK;
}

BVH_EXPORT int bvhgpu_peer_alloc(bvhgpu_ctx* ctx, size_t bytes, void** dev_ptr, void* handle64) {
    if (!ctx || !dev_ptr || !handle64) { set_error("peer_alloc: null argument"); return BVHGPU_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == BVHGPU_IPC_HANDLE_BYTES, "ipc handle size");
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    BVH_CUDA_TRY(cudaMalloc(dev_ptr, bytes ? bytes : 16));
    BVH_CUDA_TRY(cudaMemset(*dev_ptr, 0, bytes ? bytes : 16));
    BVH_CUDA_TRY(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, *dev_ptr));
    return BVHGPU_OK;
}
BVH_EXPORT int bvhgpu_peer_open(bvhgpu_ctx* ctx, const void* handle64, void** dev_ptr) {
    if (!ctx || !dev_ptr || !handle64) { set_error("peer_open: null argument"); return BVHGPU_ERR_INVALID; }
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    BVH_CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return BVHGPU_OK;
}
BVH_EXPORT int bvhgpu_peer_close(bvhgpu_ctx* ctx, void* dev_ptr) {
    if (!ctx) { set_error("peer_close: null ctx"); return BVHGPU_ERR_INVALID; }
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    if (dev_ptr) BVH_CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return BVHGPU_OK;
}
BVH_EXPORT int bvhgpu_peer_free(bvhgpu_ctx* ctx, void* dev_ptr) {
    if (!ctx) { set_error("peer_free: null ctx"); return BVHGPU_ERR_INVALID; }
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    if (dev_ptr) BVH_CUDA_TRY(cudaFree(dev_ptr));
    return BVHGPU_OK;
}

BVH_EXPORT int bvhgpu_memcpy_d2h(bvhgpu_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes) {
    if (!ctx || (bytes && (!host_dst || !dev_src))) { set_error("memcpy_d2h: null argument"); return BVHGPU_ERR_INVALID; }
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    if (bytes) BVH_CUDA_TRY(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    BVH_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return BVHGPU_OK;
}

BVH_EXPORT int bvhgpu_memcpy_h2d_async(bvhgpu_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes) {
    if (!ctx || (bytes && (!host_src || !dev_dst))) { set_error("memcpy_h2d_async: null argument"); return BVHGPU_ERR_INVALID; }
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    if (bytes) BVH_CUDA_TRY(cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return BVHGPU_OK;
}

// Pinned host memory on the device's NUMA node: the calling thread is moved onto that node's CPUs while the pages are
// allocated and first touched (first-touch placement; needs no privilege, unlike mbind), then moved back.
BVH_EXPORT int bvhgpu_host_alloc(bvhgpu_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) { set_error("host_alloc: null argument"); return BVHGPU_ERR_INVALID; }
    *out = nullptr;
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    cpu_set_t old_set, node_set;
    bool moved = false;
    if (ctx->numa_node >= 0 && sched_getaffinity(0, sizeof old_set, &old_set) == 0) {
        char path[128];
        snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", ctx->numa_node);
        CPU_ZERO(&node_set);
        int ncpu = 0;
        if (FILE* f = fopen(path, "r")) {                         // "0-31,64-95"
            int a, b;
            char sep;
            while (fscanf(f, "%d", &a) == 1) {
                b = a;
                int c = fgetc(f);
                if (c == '-') { if (fscanf(f, "%d", &b) != 1) break; c = fgetc(f); }
                for (int k = a; k <= b && k < CPU_SETSIZE; ++k) if (CPU_ISSET(k, &old_set)) { CPU_SET(k, &node_set); ++ncpu; }
                if (c != ',') break;
                (void)sep;
            }
            fclose(f);
        }
        if (ncpu > 0 && sched_setaffinity(0, sizeof node_set, &node_set) == 0) moved = true;
    }
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocPortable);
    if (e == cudaSuccess) memset(p, 0, bytes ? bytes : 16);
    if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
    if (e != cudaSuccess) { set_error("host_alloc: cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return BVHGPU_ERR_CUDA; }
    *out = p;
    return BVHGPU_OK;
}
BVH_EXPORT int bvhgpu_host_free(bvhgpu_ctx* ctx, void* p) {
    if (!ctx) { set_error("host_free: null ctx"); return BVHGPU_ERR_INVALID; }
    BVH_CUDA_TRY(cudaSetDevice(ctx->device));
    if (p) BVH_CUDA_TRY(cudaFreeHost(p));
    return BVHGPU_OK;
}

#define DEFINE_API(T, SUF, TREE, AABB, RAY, NODE, FLAT)                                                                   \
    BVH_EXPORT int bvhgpu_build_##SUF(bvhgpu_ctx* ctx, const AABB* aabbs, size_t n, int mode, TREE** out) {              \
        return build_impl<T, TREE>(ctx, aabbs, n, mode, true, out);                                                       \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_build_dev_##SUF(bvhgpu_ctx* ctx, const void* dev_aabbs, size_t n, int mode, TREE** out) {      \
        return build_impl<T, TREE>(ctx, (const AABB*)dev_aabbs, n, mode, false, out);                                     \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_tree_from_nodes_##SUF(bvhgpu_ctx* ctx, const NODE* nodes, size_t n_nodes, const AABB* aabbs,   \
                                                size_t n, TREE** out) {                                                   \
        return from_nodes_impl<T, TREE>(ctx, nodes, n_nodes, aabbs, n, out);                                              \
    }                                                                                                                     \
    BVH_EXPORT void bvhgpu_tree_free_##SUF(TREE* tree) {                                                                  \
        if (!tree) return;                                                                                                \
        if (tree->ctx) cudaSetDevice(tree->ctx->device);                                                                  \
        tree_release<T>(tree);                                                                                            \
        delete tree;                                                                                                      \
    }                                                                                                                     \
    BVH_EXPORT size_t bvhgpu_tree_num_shapes_##SUF(const TREE* tree) { return tree ? tree->n : 0; }                       \
    BVH_EXPORT size_t bvhgpu_tree_num_nodes_##SUF(const TREE* tree) { return tree ? tree->n_nodes : 0; }                  \
    BVH_EXPORT int bvhgpu_tree_nodes_##SUF(TREE* tree, NODE* out_nodes, uint32_t* out_node_index) {                       \
        return tree_nodes_impl<T>(tree, out_nodes, out_node_index);                                                       \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_flatten_##SUF(TREE* tree, FLAT* out, size_t cap, size_t* len) {                                 \
        return flatten_impl<T>(tree, out, cap, len);                                                                      \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_traverse_##SUF(TREE* tree, int mode, const RAY* rays, size_t nrays, uint32_t* offsets,          \
                                         uint32_t* hits, size_t cap, size_t* total) {                                     \
        return traverse_host_impl<T>(tree, mode, rays, BVHGPU_RAYS_FULL, nrays, offsets, hits, cap, total);               \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_traverse_od_##SUF(TREE* tree, int mode, const T* origin_dir, size_t nrays, uint32_t* offsets,   \
                                            uint32_t* hits, size_t cap, size_t* total) {                                  \
        return traverse_host_impl<T>(tree, mode, origin_dir, BVHGPU_RAYS_OD, nrays, offsets, hits, cap, total);           \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_traverse_fetch_##SUF(TREE* tree, uint32_t* hits, size_t cap) { return fetch_impl<T>(tree, hits, cap); } \
    BVH_EXPORT int bvhgpu_traverse_dev_##SUF(TREE* tree, int mode, const void* dev_rays, size_t nrays, void* dev_offsets, \
                                             void* dev_hits, size_t cap, size_t* total) {                                 \
        if (!tree || !dev_offsets || (nrays && !dev_rays)) { set_error("traverse_dev: null argument"); return BVHGPU_ERR_INVALID; } \
        BVH_CUDA_TRY(cudaSetDevice(tree->ctx->device));                                                                   \
        return traverse_device<T>(tree, mode, dev_rays, BVHGPU_RAYS_FULL, nrays, (uint32_t*)dev_offsets, (uint32_t*)dev_hits, cap, total); \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_traverse_od_dev_##SUF(TREE* tree, int mode, const void* dev_origin_dir, size_t nrays, void* dev_offsets, \
                                                void* dev_hits, size_t cap, size_t* total) {                              \
        if (!tree || !dev_offsets || (nrays && !dev_origin_dir)) { set_error("traverse_od_dev: null argument"); return BVHGPU_ERR_INVALID; } \
        BVH_CUDA_TRY(cudaSetDevice(tree->ctx->device));                                                                   \
        return traverse_device<T>(tree, mode, dev_origin_dir, BVHGPU_RAYS_OD, nrays, (uint32_t*)dev_offsets, (uint32_t*)dev_hits, cap, total); \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_traverse_sharded_dev_##SUF(TREE* tree, int mode, const void* dev_rays, size_t nrays, const bvhgpu_shard* shard) { \
        if (!tree || !shard || (nrays && !dev_rays)) { set_error("traverse_sharded: null argument"); return BVHGPU_ERR_INVALID; } \
        if (shard->world < 1 || shard->world > BVHGPU_MAX_PEERS || shard->rank < 0 || shard->rank >= shard->world) {       \
            set_error("traverse_sharded: bad rank/world %d/%d", shard->rank, shard->world); return BVHGPU_ERR_INVALID; }     \
        if (nrays == 0 || tree->n == 0) { set_error("traverse_sharded: every rank needs a non-empty shard and tree"); return BVHGPU_ERR_UNSUPPORTED; } \
        BVH_CUDA_TRY(cudaSetDevice(tree->ctx->device));                                                                   \
        if (!shard->offsets) { set_error("traverse_sharded: shard->offsets is null"); return BVHGPU_ERR_INVALID; }     \
        return traverse_device<T>(tree, mode, dev_rays, (uint32_t)shard->ray_layout, nrays, nullptr, nullptr, shard->cap, nullptr, shard); \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_query_##SUF(TREE* tree, int mode, int kind, const T* queries, size_t n, uint32_t* offsets, uint32_t* hits, \
                                      size_t cap, size_t* total) {                                                       \
        return query_host_impl<T>(tree, mode, kind, queries, n, offsets, hits, cap, total);                              \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_query_dev_##SUF(TREE* tree, int mode, int kind, const void* dev_queries, size_t n, void* dev_offsets, \
                                          void* dev_hits, size_t cap, size_t* total) {                                    \
        if (!tree || !dev_offsets || (n && !dev_queries)) { set_error("query_dev: null argument"); return BVHGPU_ERR_INVALID; } \
        BVH_CUDA_TRY(cudaSetDevice(tree->ctx->device));                                                                   \
        return query_device<T>(tree, mode, kind, (const T*)dev_queries, n, (uint32_t*)dev_offsets, (uint32_t*)dev_hits, cap, total); \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_nearest_##SUF(TREE* tree, int mode, const T* points, size_t n, uint32_t* out_shape, T* out_dist) { \
        return nearest_host_impl<T>(tree, mode, points, n, out_shape, out_dist);                                         \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_nearest_triangles_##SUF(TREE* tree, int mode, const T* points, size_t n, uint32_t* out_shape, T* out_dist) { \
        return nearest_host_impl<T>(tree, mode, points, n, out_shape, out_dist, 1);                                      \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_nearest_candidates_##SUF(TREE* tree, const T* points, size_t n, uint32_t* offsets, uint32_t* cand, \
                                                   size_t cap, size_t* total) {                                           \
        return nearest_candidates_host_impl<T>(tree, points, n, offsets, cand, cap, total);                               \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_traverse_ordered_##SUF(TREE* tree, const RAY* rays, size_t nrays, int ascending, uint32_t* offsets,      \
                                                 uint32_t* hits, T* dists, size_t cap, size_t* total) {                  \
        return ordered_host_impl<T>(tree, rays, nrays, ascending, offsets, hits, dists, cap, total);                      \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_tree_set_triangles_##SUF(TREE* tree, const T* triangles, size_t n) {                            \
        if (!tree || (n && !triangles)) { set_error("set_triangles: null argument"); return BVHGPU_ERR_INVALID; }         \
        BVH_CUDA_TRY(cudaSetDevice(tree->ctx->device));                                                                   \
        return set_triangles<T>(tree, triangles, n, false);                                                               \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_tree_set_triangles_dev_##SUF(TREE* tree, const void* dev_triangles, size_t n) {                 \
        if (!tree || (n && !dev_triangles)) { set_error("set_triangles_dev: null argument"); return BVHGPU_ERR_INVALID; } \
        BVH_CUDA_TRY(cudaSetDevice(tree->ctx->device));                                                                   \
        return set_triangles<T>(tree, (const T*)dev_triangles, n, true);                                                  \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_closest_hit_##SUF(TREE* tree, const RAY* rays, size_t nrays, int use_triangles, uint32_t* out_shape, \
                                            T* out_dist, T* out_uv) {                                                     \
        return closest_host_impl<T>(tree, rays, BVHGPU_RAYS_FULL, nrays, use_triangles, out_shape, out_dist, out_uv);     \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_closest_hit_dev_##SUF(TREE* tree, const void* dev_rays, int ray_layout, size_t nrays, int use_triangles, \
                                                void* dev_shape, void* dev_dist, void* dev_uv) {                          \
        if (!tree || (nrays && (!dev_rays || !dev_shape || !dev_dist))) { set_error("closest_hit_dev: null argument"); return BVHGPU_ERR_INVALID; } \
        BVH_CUDA_TRY(cudaSetDevice(tree->ctx->device));                                                                   \
        return closest_hit_device<T>(tree, dev_rays, (uint32_t)ray_layout, nrays, use_triangles, (uint32_t*)dev_shape, (T*)dev_dist, (T*)dev_uv); \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_traverse_stats_##SUF(TREE* tree, uint64_t* out2) {                                              \
        if (!tree || !out2) { set_error("traverse_stats: null argument"); return BVHGPU_ERR_INVALID; }                    \
        out2[0] = tree->last_visits; out2[1] = tree->last_total;                                                          \
        return BVHGPU_OK;                                                                                                 \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_rays_new_dev_##SUF(bvhgpu_ctx* ctx, const void* dev_origins, const void* dev_directions, size_t n, void* dev_rays) { \
        if (!ctx || (n && (!dev_origins || !dev_directions || !dev_rays))) { set_error("rays_new: null argument"); return BVHGPU_ERR_INVALID; } \
        BVH_CUDA_TRY(cudaSetDevice(ctx->device));                                                                         \
        return rays_new_device<T>(ctx, (const T*)dev_origins, (const T*)dev_directions, n, (RAY*)dev_rays);               \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_sah_cost_##SUF(TREE* tree, double* out2) {                                                      \
        if (!tree || !out2) { set_error("sah_cost: null argument"); return BVHGPU_ERR_INVALID; }                          \
        BVH_CUDA_TRY(cudaSetDevice(tree->ctx->device));                                                                   \
        BVH_TRY(resolve_status<T>(tree));                                                                                 \
        return sah_cost<T>(tree, out2);                                                                                   \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_refit_##SUF(TREE* tree, const AABB* aabbs, size_t n) { return refit_impl<T>(tree, aabbs, n, false); } \
    BVH_EXPORT int bvhgpu_refit_dev_##SUF(TREE* tree, const void* dev_aabbs, size_t n) { return refit_impl<T>(tree, (const AABB*)dev_aabbs, n, true); } \
    BVH_EXPORT int bvhgpu_optimize_##SUF(TREE* tree, const AABB* aabbs, size_t n, double max_growth, size_t* rebuilt) { \
        return optimize_impl<T>(tree, aabbs, n, max_growth, rebuilt, false);                                           \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_optimize_dev_##SUF(TREE* tree, const void* dev_aabbs, size_t n, double max_growth, size_t* rebuilt) { \
        return optimize_impl<T>(tree, (const AABB*)dev_aabbs, n, max_growth, rebuilt, true);                            \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_update_##SUF(TREE* tree, const uint32_t* changed, const AABB* changed_aabbs, size_t m, double max_growth, size_t* rebuilt) { \
        return update_impl<T>(tree, changed, changed_aabbs, m, max_growth, rebuilt, false);                              \
    }                                                                                                                     \
    BVH_EXPORT int bvhgpu_update_dev_##SUF(TREE* tree, const void* dev_changed, const void* dev_changed_aabbs, size_t m, double max_growth, size_t* rebuilt) { \
        return update_impl<T>(tree, (const uint32_t*)dev_changed, (const AABB*)dev_changed_aabbs, m, max_growth, rebuilt, true); \
    }

#define DEFINE_API2(T, SUF, TREE, AABB, RAY, NODE, FLAT)                                                                   \
    BVH_EXPORT int bvhgpu_build_##SUF(bvhgpu_ctx* ctx, const AABB* aabbs, size_t n, int mode, TREE** out) {               \
        return build2_impl<T, TREE, AABB>(ctx, aabbs, n, mode, out);                                                       \
    }                                                                                                                      \
    BVH_EXPORT void bvhgpu_tree_free_##SUF(TREE* tree) {                                                                   \
        if (!tree) return;                                                                                                 \
        if (tree->ctx) cudaSetDevice(tree->ctx->device);                                                                   \
        tree_release<T>(tree);                                                                                             \
        delete tree;                                                                                                       \
    }                                                                                                                      \
    BVH_EXPORT size_t bvhgpu_tree_num_shapes_##SUF(const TREE* tree) { return tree ? tree->n : 0; }                        \
    BVH_EXPORT int bvhgpu_tree_nodes_##SUF(TREE* tree, NODE* out_nodes, uint32_t* out_node_index) {                        \
        return tree_nodes2_impl<T, NODE>(tree, out_nodes, out_node_index);                                                 \
    }                                                                                                                      \
    BVH_EXPORT int bvhgpu_flatten_##SUF(TREE* tree, FLAT* out, size_t cap, size_t* len) { return flatten2_impl<T, FLAT>(tree, out, cap, len); } \
    BVH_EXPORT int bvhgpu_traverse_##SUF(TREE* tree, int mode, const RAY* rays, size_t nrays, uint32_t* offsets, uint32_t* hits, \
                                         size_t cap, size_t* total) {                                                      \
        return traverse2_impl<T, RAY>(tree, mode, rays, nrays, offsets, hits, cap, total);                                 \
    }

DEFINE_API2(float, f32x2, bvhgpu_tree2f, bvh_aabb2f, bvh_ray2f, bvh_node2f, bvh_flat2f)
DEFINE_API2(double, f64x2, bvhgpu_tree2d, bvh_aabb2d, bvh_ray2d, bvh_node2d, bvh_flat2d)
DEFINE_API(float, f32x3, bvhgpu_tree3f, bvh_aabb3f, bvh_ray3f, bvh_node3f, bvh_flat3f)
DEFINE_API(double, f64x3, bvhgpu_tree3d, bvh_aabb3d, bvh_ray3d, bvh_node3d, bvh_flat3d)
