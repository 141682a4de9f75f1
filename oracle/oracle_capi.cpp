// ============================================================================
// oracle/oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY (see bvh_oracle.hpp).
// extern "C" surface over the CPU restatement so that tests/ and bench.py's
// cpu_baseline / --impl reference legs can drive it through ctypes.
// ============================================================================
#include "bvh_oracle.hpp"
#include <cstring>
#include <chrono>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <pthread.h>
#include <sched.h>

using namespace orc;

namespace {

// ---- persistent worker pool for the timed CPU legs (bench.py --impl reference / cpu_baseline) ------------------------------
// Threads are created once, pinned round-robin to the CPUs the process may use, and woken per batch; rays are handed out in
// chunks from an atomic counter (dynamic scheduling: the batch's expensive rays are not evenly spread).  Creating ~100 fresh
// threads per call and splitting the batch statically made the measured CPU rate swing by 3-4x between otherwise equal hosts.
class Pool {
  public:
    static Pool& get() { static Pool p; return p; }
    unsigned size() const { return (unsigned)workers_.size(); }
    // runs fn(worker) on workers [0, nthreads) and returns when all are done
    void run(unsigned nthreads, const std::function<void(unsigned)>& fn) {
        nthreads = std::max(1u, std::min(nthreads, size()));
        std::unique_lock<std::mutex> lk(m_);
        fn_ = &fn; active_ = nthreads; pending_ = nthreads; ++gen_;
        cv_.notify_all();
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }
  private:
    Pool() {
        std::vector<int> cpus;
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof set, &set) == 0)
            for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &set)) cpus.push_back(c);
        unsigned n = std::max(1u, std::thread::hardware_concurrency());
        if (!cpus.empty()) n = std::min<unsigned>(n, (unsigned)cpus.size());
        for (unsigned t = 0; t < n; ++t) {
            workers_.emplace_back([this, t] { loop(t); });
            if (!cpus.empty()) {
                cpu_set_t one;
                CPU_ZERO(&one);
                CPU_SET(cpus[t % cpus.size()], &one);
                pthread_setaffinity_np(workers_.back().native_handle(), sizeof one, &one);
            }
        }
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    void loop(unsigned t) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(unsigned)>* fn = nullptr;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                if (t < active_) fn = fn_;
            }
            if (fn) {
                (*fn)(t);
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(unsigned)>* fn_ = nullptr;
    unsigned active_ = 0, pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

template <class T>
uint64_t traverse_batch(int mode, const void* tree, uint32_t n_tree, const Aabb3<T>* shapes,
                        const Ray3<T>* rays, uint64_t nrays, uint64_t* offsets, uint32_t* hits, uint64_t cap,
                        uint64_t* stats, unsigned threads, int* overflow32) {
    // mode 0: Bvh::traverse (recursive)   1: FlatBvh::traverse   2: BvhTraverseIterator
    // stats == nullptr: no visit counters in the hot loop (the timed legs); stats != nullptr: counters + wall time in stats[4].
    if (threads < 1) threads = 1;
    const uint64_t CHUNK = 2048;
    const uint64_t nchunks = (nrays + CHUNK - 1) / CHUNK;
    std::vector<std::vector<uint32_t>> lists(nchunks);
    std::vector<std::vector<uint32_t>> counts(nchunks);
    std::vector<TraverseStats> tst(std::max(1u, threads));
    std::vector<int> ok(std::max(1u, threads), 1);
    std::atomic<uint64_t> next{0};
    const bool want_stats = stats != nullptr && stats[5] == 0;      // stats[5] != 0 on entry: time only, no counters
    auto work = [&](unsigned t) {
        std::vector<uint32_t> out;
        TraverseStats ls;                        // thread-local: no false sharing in the hot loop
        TraverseStats* lsp = want_stats ? &ls : nullptr;
        int lok = 1;
        for (;;) {
            const uint64_t c = next.fetch_add(1, std::memory_order_relaxed);
            if (c >= nchunks) break;
            const uint64_t lo = c * CHUNK, hi = std::min(nrays, lo + CHUNK);
            std::vector<uint32_t>& mine = lists[c];
            std::vector<uint32_t>& cnt = counts[c];
            cnt.reserve(hi - lo);
            for (uint64_t r = lo; r < hi; ++r) {
                out.clear();
                if (mode == 0) traverse_recursive((const Node<T>*)tree, n_tree, shapes, rays[r], out, lsp);
                else if (mode == 1) traverse_flat((const FlatNode<T>*)tree, n_tree, shapes, rays[r], out, lsp);
                else if (!traverse_iterator((const Node<T>*)tree, n_tree, shapes, rays[r], out)) lok = 0;
                cnt.push_back((uint32_t)out.size());
                mine.insert(mine.end(), out.begin(), out.end());
            }
        }
        tst[t] = ls;
        ok[t] = lok;
    };
    const auto tic = std::chrono::steady_clock::now();
    if (threads == 1) work(0);
    else Pool::get().run(threads, work);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - tic).count();
    // assembly into one CSR: chunk bases serially (a few hundred), then every chunk's offsets / hits in parallel
    std::vector<uint64_t> base(nchunks + 1, 0);
    for (uint64_t c = 0; c < nchunks; ++c) base[c + 1] = base[c] + lists[c].size();
    const uint64_t total = base[nchunks];
    std::atomic<uint64_t> next2{0};
    auto assemble = [&](unsigned) {
        for (;;) {
            const uint64_t c = next2.fetch_add(1, std::memory_order_relaxed);
            if (c >= nchunks) break;
            if (offsets) { uint64_t run = base[c], r = c * CHUNK; for (uint32_t k : counts[c]) { offsets[r++] = run; run += k; } }
            if (hits) { uint64_t w = base[c]; for (uint32_t h : lists[c]) { if (w < cap) hits[w] = h; ++w; } }
        }
    };
    if (threads == 1) assemble(0); else Pool::get().run(threads, assemble);
    if (offsets) offsets[nrays] = total;
    if (stats) {
        stats[0] = stats[1] = stats[2] = stats[3] = 0;
        for (auto& s : tst) { stats[0] += s.node_visits; stats[1] += s.slab_tests; stats[2] += s.leaf_visits; stats[3] += s.hits; }
        stats[4] = (uint64_t)(secs * 1e9);          // wall time of the traversal section (dispatch .. all workers done), ns
    }
    if (overflow32) { *overflow32 = 0; for (int o : ok) if (!o) *overflow32 = 1; }
    return total;
}

template <class T>
void tri_aabbs(const T* tris, uint64_t n, Aabb3<T>* out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = triangle_aabb(tris + 9 * i, tris + 9 * i + 3, tris + 9 * i + 6);
}

}  // namespace

static uint64_t g_last_build_ns = 0;

#define ORC_API extern "C" __attribute__((visibility("default")))

#define DEFINE_FOR(T, SUF)                                                                                   \
    ORC_API void orc_build_##SUF(const Aabb3<T>* shapes, uint32_t n, Node<T>* nodes, uint32_t* node_index,   \
                                 uint64_t* stats4) {                                                         \
        const auto tic = std::chrono::steady_clock::now();                                                   \
        BuildStats st = build(shapes, n, nodes, node_index);                                                 \
        g_last_build_ns = (uint64_t)(std::chrono::duration<double>(std::chrono::steady_clock::now() - tic).count() * 1e9); \
        if (stats4) { stats4[0] = st.prim_visits; stats4[1] = st.degenerate_splits; stats4[2] = st.max_depth; stats4[3] = st.nosplit_fallthrough; } \
    }                                                                                                        \
    ORC_API void orc_build_par_##SUF(const Aabb3<T>* shapes, uint32_t n, Node<T>* nodes, uint32_t* node_index, \
                                     uint64_t* stats4, uint32_t threads) {                                   \
        const auto tic = std::chrono::steady_clock::now();                                                   \
        BuildStats st = build_par(shapes, n, nodes, node_index, threads);                                    \
        g_last_build_ns = (uint64_t)(std::chrono::duration<double>(std::chrono::steady_clock::now() - tic).count() * 1e9); \
        if (stats4) { stats4[0] = st.prim_visits; stats4[1] = st.degenerate_splits; stats4[2] = st.max_depth; stats4[3] = st.nosplit_fallthrough; } \
    }                                                                                                        \
    ORC_API uint64_t orc_flatten_##SUF(const Node<T>* nodes, uint32_t n_nodes, FlatNode<T>* out, uint64_t cap) { \
        std::vector<FlatNode<T>> v;                                                                          \
        flatten(nodes, n_nodes, v);                                                                          \
        if (out) std::memcpy(out, v.data(), sizeof(FlatNode<T>) * std::min<uint64_t>(cap, v.size()));       \
        return v.size();                                                                                     \
    }                                                                                                        \
    ORC_API uint64_t orc_traverse_batch_##SUF(int mode, const void* tree, uint32_t n_tree, const Aabb3<T>* shapes, \
                                              const Ray3<T>* rays, uint64_t nrays, uint64_t* offsets,        \
                                              uint32_t* hits, uint64_t cap, uint64_t* stats5, uint32_t threads, \
                                              int* iter_overflow) {                                          \
        return traverse_batch<T>(mode, tree, n_tree, shapes, rays, nrays, offsets, hits, cap, stats5, threads, iter_overflow); \
    }                                                                                                        \
    ORC_API uint64_t orc_query_batch_##SUF(int kind, int use_flat, const Node<T>* nodes, uint32_t n_nodes, const FlatNode<T>* flat, \
                                           uint32_t n_flat, const Aabb3<T>* shapes, const T* queries, uint64_t nq, uint64_t* offsets, \
                                           uint32_t* hits, uint64_t cap) {                                    \
        const int stride = kind == 1 ? 6 : (kind == 2 ? 3 : 4);                                              \
        std::vector<uint32_t> out;                                                                           \
        uint64_t total = 0;                                                                                  \
        for (uint64_t i = 0; i < nq; ++i) {                                                                  \
            out.clear();                                                                                     \
            traverse_query<T>(kind, queries + i * stride, nodes, n_nodes, flat, n_flat, shapes, use_flat != 0, out); \
            offsets[i] = total;                                                                              \
            for (uint32_t h : out) { if (total < cap) hits[total] = h; ++total; }                            \
        }                                                                                                    \
        offsets[nq] = total;                                                                                 \
        return total;                                                                                        \
    }                                                                                                        \
    ORC_API void orc_closest_hit_batch_##SUF(int kind, const Node<T>* nodes, uint32_t n_nodes, const Aabb3<T>* shapes, const T* tris, \
                                             const Ray3<T>* rays, uint64_t nrays, uint32_t* out_shape, T* out_dist, T* out_uv) { \
        for (uint64_t i = 0; i < nrays; ++i) {                                                               \
            T d, u, v;                                                                                       \
            out_shape[i] = closest_hit(nodes, n_nodes, shapes, tris, kind, rays[i], d, u, v);                \
            out_dist[i] = d;                                                                                 \
            if (out_uv) { out_uv[2 * i] = u; out_uv[2 * i + 1] = v; }                                        \
        }                                                                                                    \
    }                                                                                                        \
    ORC_API T orc_ray_triangle_##SUF(const Ray3<T>* ray, const T* abc9, T* uv2) {                            \
        return ray_intersects_triangle(*ray, abc9, abc9 + 3, abc9 + 6, uv2[0], uv2[1]);                      \
    }                                                                                                        \
    ORC_API int orc_ray_slice_##SUF(const Ray3<T>* ray, const Aabb3<T>* aabb, T* out2) {                      \
        return ray_slice_for_aabb(*ray, *aabb, out2[0], out2[1]) ? 1 : 0;                                    \
    }                                                                                                        \
    ORC_API int orc_is_consistent_##SUF(const Node<T>* nodes, uint32_t n_nodes, const Aabb3<T>* shapes) {    \
        return is_consistent(nodes, n_nodes, shapes) ? 1 : 0;                                                \
    }                                                                                                        \
    ORC_API int orc_is_tight_##SUF(const Node<T>* nodes, uint32_t n_nodes) { return is_tight(nodes, n_nodes) ? 1 : 0; } \
    ORC_API void orc_sah_cost_##SUF(const Node<T>* nodes, uint32_t n_nodes, double* out2) {                  \
        sah_cost(nodes, n_nodes, out2[0], out2[1]);                                                          \
    }                                                                                                        \
    /* nearest_to for nq points; use_flat: FlatBvh::nearest_to; kind 0 = AABB distance, 1 = triangle distance (tris: 9 per shape) */ \
    ORC_API void orc_nearest_batch_##SUF(int use_flat, int kind, const void* tree, uint32_t n_tree, const Aabb3<T>* shapes, const T* tris, \
                                         const T* points, uint64_t nq, uint32_t* out_shape, T* out_dist) {     \
        const ShapeDist<T> sd{kind, shapes, tris};                                                           \
        for (uint64_t i = 0; i < nq; ++i)                                                                    \
            out_shape[i] = use_flat ? nearest_to_flat((const FlatNode<T>*)tree, n_tree, points + 3 * i, sd, out_dist[i]) \
                                    : nearest_to((const Node<T>*)tree, n_tree, points + 3 * i, sd, out_dist[i]); \
    }                                                                                                        \
    ORC_API void orc_shape_distance_##SUF(int kind, const Aabb3<T>* shapes, const T* tris, uint32_t n, const T* point, T* out_d2) { \
        const ShapeDist<T> sd{kind, shapes, tris};                                                           \
        for (uint32_t s = 0; s < n; ++s) out_d2[s] = sd(s, point);                                           \
    }                                                                                                        \
    /* Bvh::update_shapes on (nodes, node_index) in place; returns the node count or U32_MAX on a reference panic path */ \
    ORC_API uint32_t orc_update_shapes_##SUF(Node<T>* nodes, uint32_t n_nodes, uint32_t* node_index, uint32_t n_shapes, \
                                             const Aabb3<T>* shapes, const uint32_t* changed, uint32_t n_changed) { \
        DynBvh<T> b;                                                                                         \
        b.nodes.assign(nodes, nodes + n_nodes);                                                              \
        b.node_index.assign(node_index, node_index + n_shapes);                                              \
        try { update_shapes(b, changed, n_changed, shapes); recount(b); } catch (const std::logic_error&) { return U32_MAX; } \
        if (b.nodes.size() > n_nodes) return U32_MAX;                                                        \
        std::memcpy(nodes, b.nodes.data(), sizeof(Node<T>) * b.nodes.size());                                \
        std::memcpy(node_index, b.node_index.data(), sizeof(uint32_t) * n_shapes);                           \
        return (uint32_t)b.nodes.size();                                                                     \
    }                                                                                                        \
    ORC_API int orc_connect_nodes_##SUF(Node<T>* nodes, uint32_t n_nodes, const Aabb3<T>* shapes, uint32_t child, \
                                        uint32_t parent, int left_child) {                                   \
        DynBvh<T> b;                                                                                         \
        b.nodes.assign(nodes, nodes + n_nodes);                                                              \
        try { connect_nodes(b, child, parent, left_child != 0, shapes); } catch (const std::logic_error&) { return 0; } \
        std::memcpy(nodes, b.nodes.data(), sizeof(Node<T>) * n_nodes);                                       \
        return 1;                                                                                            \
    }                                                                                                        \
    ORC_API void orc_create_n_cubes_##SUF(uint32_t n_cubes, const Aabb3<T>* bounds, T* tris_out /*108 per cube*/, \
                                          Aabb3<T>* aabbs_out /*12 per cube*/) {                             \
        std::vector<T> tris;                                                                                 \
        create_n_cubes(n_cubes, *bounds, tris);                                                              \
        if (tris_out) std::memcpy(tris_out, tris.data(), tris.size() * sizeof(T));                           \
        if (aabbs_out) tri_aabbs(tris.data(), (uint64_t)n_cubes * 12, aabbs_out);                            \
    }                                                                                                        \
    ORC_API void orc_tri_aabbs_##SUF(const T* tris, uint64_t n, Aabb3<T>* out) { tri_aabbs(tris, n, out); }  \
    ORC_API void orc_create_rays_##SUF(uint64_t* seed, const Aabb3<T>* bounds, uint64_t n, Ray3<T>* out) {   \
        for (uint64_t i = 0; i < n; ++i) out[i] = create_ray(*seed, *bounds);                                \
    }                                                                                                        \
    ORC_API void orc_next_points_##SUF(uint64_t* seed, const Aabb3<T>* bounds, uint64_t n, T* out) {         \
        for (uint64_t i = 0; i < n; ++i) next_point3(*seed, *bounds, out + 3 * i);                           \
    }                                                                                                        \
    ORC_API void orc_ray_new_##SUF(const T* origins, const T* dirs, uint64_t n, Ray3<T>* out) {              \
        for (uint64_t i = 0; i < n; ++i) out[i] = ray_new(origins + 3 * i, dirs + 3 * i);                    \
    }                                                                                                        \
    ORC_API int orc_ray_intersects_aabb_##SUF(const Ray3<T>* ray, const Aabb3<T>* aabb) {                    \
        return ray_intersects_aabb(*ray, *aabb) ? 1 : 0;                                                     \
    }                                                                                                        \
    ORC_API void orc_aligned_boxes_##SUF(Aabb3<T>* out21) {                                                  \
        std::vector<Aabb3<T>> v;                                                                             \
        aligned_boxes(v);                                                                                    \
        std::memcpy(out21, v.data(), sizeof(Aabb3<T>) * 21);                                                 \
    }                                                                                                        \
    ORC_API void orc_aabb_ops_##SUF(const Aabb3<T>* a, T* center3, T* surface_area, int* largest_axis) {     \
        aabb_center(*a, center3);                                                                            \
        *surface_area = aabb_surface_area(*a);                                                               \
        *largest_axis = aabb_largest_axis(*a);                                                               \
    }

DEFINE_FOR(float, f32)
DEFINE_FOR(double, f64)

ORC_API uint64_t orc_last_build_ns() { return g_last_build_ns; }
ORC_API uint64_t orc_splitmix64(uint64_t* seed) { return splitmix64(*seed); }
ORC_API uint32_t orc_hardware_threads() { return std::thread::hardware_concurrency(); }
ORC_API uint32_t orc_pool_threads() { return Pool::get().size(); }
ORC_API uint32_t orc_sizeof(int what) {
    switch (what) {
        case 0: return sizeof(Aabb3<float>);
        case 1: return sizeof(Ray3<float>);
        case 2: return sizeof(Node<float>);
        case 3: return sizeof(FlatNode<float>);
        case 4: return sizeof(Aabb3<double>);
        case 5: return sizeof(Ray3<double>);
        case 6: return sizeof(Node<double>);
        case 7: return sizeof(FlatNode<double>);
    }
    return 0;
}
