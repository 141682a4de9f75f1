// bvh_b200/csrc/internal.h -- host-side structures shared by the translation units of libbvh_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <cstdio>
#include "common.cuh"

namespace bvhb200 {

void set_error(const char* fmt, ...);

#define BVH_CUDA_TRY(expr)                                                                          \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::bvhb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return BVHGPU_ERR_CUDA;                                                                 \
        }                                                                                           \
    } while (0)

#define BVH_TRY(expr)                  \
    do {                               \
        int _s = (expr);               \
        if (_s != BVHGPU_OK) return _s; \
    } while (0)

}  // namespace bvhb200

// Device status block a build writes (read back lazily).
struct BuildStatus {
    uint32_t error;        // bvhgpu_status raised on the device (timeout / internal)
    uint32_t nan_found;    // prep kernel saw a NaN coordinate
    uint32_t tickets;      // queue tickets handed out (diagnostics)
    uint32_t leaves_done;  // must equal n at the end
    uint32_t rebuilt;      // optimize: shapes in the subtrees that were rebuilt
    uint32_t pad[3];
};

struct bvhgpu_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    int64_t traverse_slots = -1;   // per-ray hit slots of the single-pass traversal (0 = two-pass, -1 = auto by batch size)
    int64_t traverse_persistent = 2;   // 0: one ray per thread, 1: persistent refill kernel, 2: coherence probe decides on the device
    int walk_grid = 0;             // persistent grid size: one full wave (computed once), or the value of option "walk_grid"
    bool walk_grid_forced = false;
    bool top_attr_set = false;     // walk_top_kernel's dynamic shared memory limit has been raised on this device
    int64_t build_gang = -1;       // exact builder: co-resident warp gangs for the top levels (-1 / 1 on, 0 off = queue tiles only)
    int64_t build_subtree = -1;    // exact builder: in-register subtrees for ranges <= 32 shapes (-1 auto, 0 never, 1 always)
    int64_t build_small = -1;      // exact builder: defer ranges <= 16 shapes to the thread-per-range kernel (-1 auto by size, 0 never, 1 always)
    uint32_t* h_pinned = nullptr;  // small pinned read-back area (256 words)
    int64_t profile = 0;           // bracket dominant kernels with events
    cudaEvent_t ev_walk[2] = {nullptr, nullptr};
    cudaEvent_t ev_build[2] = {nullptr, nullptr};
    bool have_walk = false, have_build = false;
    // host-pointer traversal: H2D chunks on a second stream, overlapped with the walks
    cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_order = nullptr, ev_total = nullptr;
    cudaEvent_t ev_chunk[16] = {}, ev_emit[16] = {};
    cudaEvent_t ev_e2e[5] = {};    // profile: start, walk end, last H2D done, last emit, last D2H
    bool have_e2e = false;
    // host-pointer traversal, streaming form (rays consumed by a running kernel while they arrive): only when a kernel launch
    // does not block the host and no tool serialises / replays launches.  -1 = not probed yet, 0 = never stream, 1 = ok.
    int stream_ok = -1;
    double host_us[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // profile: host-side time stamps of the last host-pointer traversal (us since entry)
    uint64_t slot_budget_words = 0; // traversal slot scratch budget (a quarter of the memory that was free at first use, in 4-byte words / 4)
    int last_streamed = -1;        // did the last host-pointer traversal stream (1) or copy-then-walk (0)
    int64_t traverse_stream = -1;  // option "traverse_stream": -1 auto (probe), 0 never, 1 force
    int64_t traverse_top = -1;     // option "traverse_top": shared-memory top-of-tree walk: -1 auto, 0 off, 1 on
    uint32_t* d_ready = nullptr;       // cudaMalloc'ed word the streamed host path bumps with cuStreamWriteValue32 (refused on pool memory)
    int wv_ok = -1;                    // did the last streamed call use stream write-value flags (1) or 4-byte copies (0)
    uint32_t* d_async_err = nullptr;   // sticky device-side error word of asynchronous calls (sharded exchange): surfaced by bvhgpu_synchronize
    int numa_node = -1;            // NUMA node of the device (sysfs), -1 unknown
};
#define BVH_MAX_CHUNKS 16u

namespace bvhb200 {

template <class T> struct Tree {
    using Tr = Traits<T>;
    bvhgpu_ctx* ctx = nullptr;
    uint32_t n = 0;          // shapes
    uint32_t n_nodes = 0;    // 2n-1
    int dims = 3;            // 2: a Bvh<T,2> embedded in the plane z = 0 (dim2.cu)
    typename Tr::DAabb* d_aabb = nullptr;     // [n]      shape AABBs (padded device layout)
    typename Tr::DAabb* d_aabb_trav = nullptr; // [n]     dims == 2 only: the same with z = [-1, +1] for the FLAT leaf re-test
    typename Tr::Node* d_nodes = nullptr;     // [2n-1]   Bvh.nodes, reference preorder layout
    uint32_t* d_node_index = nullptr;         // [n]      leaf node of every shape
    uint32_t* d_node_start = nullptr;         // [2n-1]   first position of the node's shape range (== #leaves before it)
    typename Tr::TNode* d_tnodes = nullptr;   // [n_trec] traversal records
    uint32_t n_trec = 0;
    void* d_top = nullptr;                    // f32: top-of-tree records for walk_top_kernel (32-B header {n_top, C}, lo[n_top], hi[n_top] float4), built lazily
    uint32_t top_cap = 0, top_budget = 0;
    bool top_valid = false;
    uint32_t* d_arrive = nullptr;             // [2n-1] arrival counters of the incremental update (all zero between calls)
    uint8_t* d_bad = nullptr;                 // [2n-1] growth flags of the incremental update (all zero between calls)
    void* d_sa_base = nullptr;                // [2n-1] surface area of every inner node when it was last (re)built: baseline of bvhgpu_optimize / update
    void* d_tris = nullptr;                   // [n] triangle vertices (padded), optional: bvhgpu_tree_set_triangles_*
    typename Tr::Flat* d_flat = nullptr;      // [n_flat] reference-layout FlatBvh (built on demand)
    size_t n_flat = 0;
    bool have_flat = false;
    // deferred build status
    BuildStatus* d_status = nullptr;
    BuildStatus* h_status = nullptr;          // pinned
    bool status_pending = false;
    int failed_status = 0;                    // sticky: first failure of build / refit / optimize (BVHGPU_OK = healthy)
    std::string failed_message;
    // retained result of the last traversal
    uint32_t* d_offsets = nullptr; size_t offsets_cap = 0;
    uint32_t* d_hits = nullptr;    size_t hits_cap = 0;
    size_t last_total = 0, last_nrays = 0;
    uint64_t last_visits = 0;
};

// Resolve the deferred device status of a build / refit (synchronises the stream once).
template <class T> int resolve_status(Tree<T>* tree);

// ---- memory (stream-ordered pool) ----
int dalloc(bvhgpu_ctx* ctx, void** p, size_t bytes);
void dfree(bvhgpu_ctx* ctx, void* p);
template <class P> inline int dalloc_t(bvhgpu_ctx* ctx, P** p, size_t count) { return dalloc(ctx, (void**)p, count * sizeof(P)); }
// Scratch that is released (stream-ordered) when the scope ends, on every return path.
struct Scratch {
    bvhgpu_ctx* ctx;
    void* ptrs[16];
    int n = 0;
    explicit Scratch(bvhgpu_ctx* c) : ctx(c) {}
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    ~Scratch() { for (int i = 0; i < n; ++i) dfree(ctx, ptrs[i]); }
    template <class P> int get(P** p, size_t count) {
        *p = nullptr;
        if (n >= 16) { set_error("internal: scratch table full"); return BVHGPU_ERR_INTERNAL; }
        const int rc = dalloc(ctx, (void**)p, count * sizeof(P));
        if (rc == BVHGPU_OK) ptrs[n++] = (void*)*p;
        return rc;
    }
};

// ---- build_sah.cu ----
// in_aabbs: device pointer to n AABBs in the C-ABI layout (24 B / 48 B).  Fills tree->d_aabb, d_nodes,
// d_node_index, d_node_start (allocated here).  Asynchronous; errors are reported via tree->d_status.
template <class T> int build_exact_sah(bvhgpu_ctx* ctx, const typename Traits<T>::Aabb* in_aabbs, uint32_t n, Tree<T>* tree);
// Converts ABI-layout AABBs to the device layout only (used by tree_from_nodes and refit).
template <class T> int convert_aabbs(bvhgpu_ctx* ctx, const typename Traits<T>::Aabb* in_aabbs, uint32_t n,
                                     typename Traits<T>::DAabb* out, uint32_t* d_nan_flag);

// ---- treelet session (build_sah.cu), used by lbvh.cu for BVHGPU_BUILD_LBVH_TREELET ----
struct BuildCtl;
template <class T> struct QSlot;
template <class T> struct TreeletSession { void* params = nullptr; QSlot<T>* q = nullptr; uint32_t* qseq = nullptr; uint32_t qmask = 0; BuildCtl* ctl = nullptr; };
template <class T> int treelet_begin(bvhgpu_ctx* ctx, Tree<T>* tree, uint32_t* sorted_ids, TreeletSession<T>* S);
template <class T> int treelet_finish(bvhgpu_ctx* ctx, Tree<T>* tree, TreeletSession<T>* S);

// ---- rebuild session (build_sah.cu), used by optimize(): the exact builder restarted from inner nodes.  d_roots[0 .. *d_n_roots)
// are node indices of disjoint subtrees, cb[node][6] the bounds of the shape centres below every node, idx0 the shapes in
// leaf order.  Rewrites d_nodes / d_node_index / d_node_start of those subtrees in place.
// cb_by_root: cb holds 6 values per ROOT (in d_roots order) instead of per node
template <class T> int rebuild_subtrees(bvhgpu_ctx* ctx, Tree<T>* tree, const uint32_t* d_roots, const uint32_t* d_n_roots, const T* cb, uint32_t* idx0, bool cb_by_root);

// ---- lbvh.cu ----
template <class T> int build_lbvh(bvhgpu_ctx* ctx, const typename Traits<T>::Aabb* in_aabbs, uint32_t n, Tree<T>* tree, bool treelets);

// ---- flatten.cu ----
template <class T> int build_traversal_records(Tree<T>* tree);   // d_tnodes
template <class T> int build_flat(Tree<T>* tree);                // d_flat (reference FlatNode layout)
int build_top_records(Tree<float>* tree, uint32_t budget);        // d_top
template <class T> int sah_cost(Tree<T>* tree, double* out2);
template <class T> int refit(Tree<T>* tree);                     // recompute child AABBs bottom-up from d_aabb
template <class T> int optimize(Tree<T>* tree, double max_growth);   // refit + exact rebuild of the degraded subtrees
// update_shapes form: validate (flags[0] NaN, flags[1] bad index) / scatter m changed AABBs (device pointers) into tree->d_aabb
template <class T> int update_changed(Tree<T>* tree, const uint32_t* d_changed, const typename Traits<T>::Aabb* d_fresh, uint32_t m, uint32_t* d_flags);
template <class T> int update_incremental(Tree<T>* tree, const uint32_t* d_changed, uint32_t m, double max_growth);
template <class T> int update_scatter(Tree<T>* tree, const uint32_t* d_changed, const typename Traits<T>::Aabb* d_fresh, uint32_t m);

// ---- traverse.cu ----
// d_rays: rays on the device; fmt: BVHGPU_RAYS_FULL (9 scalars: the C-ABI Ray) or BVHGPU_RAYS_OD (6 scalars: origin, direction).
// shard != nullptr: multi-GPU step, results go to every rank's peer-mapped global CSR (d_offsets / d_hits unused).
template <class T> int traverse_device(Tree<T>* tree, int mode, const void* d_rays, uint32_t fmt,
                                       size_t nrays, uint32_t* d_offsets, uint32_t* d_hits, size_t cap, size_t* total,
                                       const bvhgpu_shard* shard = nullptr);
// Host rays in, host CSR out; H2D / walk+scan+emit / D2H overlapped.  Needs tree->d_offsets / d_hits sized by the caller.
template <class T> int traverse_host_pipelined(Tree<T>* tree, int mode, const void* h_rays, uint32_t fmt, size_t nrays,
                                               uint32_t* h_offsets, uint32_t* h_hits, size_t h_cap, size_t* total);
// hits sorted by entry (ascending) / exit (descending) distance, with the distances; device pointers
template <class T> int traverse_ordered_device(Tree<T>* tree, const typename Traits<T>::Ray* d_rays, size_t nrays, int ascending,
                                               uint32_t* d_offsets, uint32_t* d_hits, T* d_dists, size_t cap, size_t* total);
// Aabb / Point / Ball queries (device pointers); two-pass count / fill.
template <class T> int query_device(Tree<T>* tree, int mode, int kind, const T* d_queries, size_t nq, uint32_t* d_offsets, uint32_t* d_hits, size_t cap, size_t* total);
// nearest_to for a batch of points (device pointers): exact reference walk for AABB-distance shapes; candidate lists for any shape
template <class T> int nearest_device(Tree<T>* tree, int mode, const T* d_points, size_t nq, uint32_t* d_shape, T* d_dist, int use_triangles = 0);
template <class T> int nearest_candidates_device(Tree<T>* tree, const T* d_points, size_t nq, uint32_t* d_offsets, uint32_t* d_cand, size_t cap, size_t* total);
// ---- dim2.cu ----
template <class T> int dim2_expand_aabbs(bvhgpu_ctx* ctx, const T* d_in4, uint32_t n, T* d_out6);
template <class T> int dim2_expand_rays(bvhgpu_ctx* ctx, const T* d_in6, uint32_t n, T* d_out9);
template <class T> int dim2_finish_build(Tree<T>* tree);
template <class T, class N2> int dim2_nodes_out(Tree<T>* tree, N2* d_out);
template <class T, class F2> int dim2_flat_out(Tree<T>* tree, F2* d_out);
// ---- closest.cu ----
template <class T> int set_triangles(Tree<T>* tree, const T* tris9, size_t n, bool dev_input);
template <class T> int closest_hit_device(Tree<T>* tree, const void* d_rays, uint32_t fmt, size_t nrays, int use_triangles, uint32_t* d_shape, T* d_dist, T* d_uv);
template <class T> int rays_new_device(bvhgpu_ctx* ctx, const T* d_origins, const T* d_dirs, size_t n,
                                       typename Traits<T>::Ray* d_rays);

}  // namespace bvhb200

struct bvhgpu_tree3f : bvhb200::Tree<float> {};
struct bvhgpu_tree3d : bvhb200::Tree<double> {};
struct bvhgpu_tree2f : bvhb200::Tree<float> {};
struct bvhgpu_tree2d : bvhb200::Tree<double> {};
