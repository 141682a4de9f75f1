// bvh_b200/csrc/build_sah.cu -- exact 6-bucket SAH builder, bit-identical to the reference's
// Bvh::build (src/bvh/bvh_impl.rs:53-96, src/bvh/bvh_node.rs:81-279, src/utils.rs:59-109).
//
// B200-first design (DESIGN.md "builder"): the reference recursion is re-expressed as ONE
// persistent kernel in which every warp is a worker pulling tasks from a device-wide ticket
// queue.  A task is a node of the tree = a contiguous range of the shape-index array:
//
//   SEG      (range <= TILE shapes): one warp bins the range into the 6 SAH buckets (shared
//            memory min/max on order-preserving integer keys), evaluates the 5 splits in the
//            reference's operation order, performs the STABLE 6-way partition with warp ballots
//            (the reference rewrites indices bucket by bucket, bvh_node.rs:250-272 -- the order
//            is observable through the degenerate "halve by position" branch, :114-124), writes
//            the node, then continues depth-first with the left child; the right child goes to
//            a per-warp shared-memory stack (small) or to the global queue (large).
//   BIN / SCATTER tiles (range > TILE): the range is cut into TILE-sized tiles processed by
//            different warps; per-tile bucket counts are prefix-summed by the last finishing
//            tile so that the multi-warp partition is still stable; the last SCATTER tile emits
//            the node and creates the children.
//
// No level-wide barriers, no host round trips, one launch for the whole tree.  Node indices
// follow the reference rule child_l = i+1, child_r = i + 2*n_l (bvh_node.rs:138-142), so the
// output array IS the reference's preorder `Bvh.nodes`.
#include "internal.h"
#include "build_types.cuh"
#include <vector>
#include <cstdlib>
#include <cstring>

namespace bvhb200 {

template <class T> struct BuildParams {
    using Tr = Traits<T>;
    const typename Tr::DAabb* aabb;
    uint32_t* idx[2];
    uint8_t* bkt;
    typename Tr::Node* nodes;
    uint32_t* node_index;
    uint32_t* node_start;
    QSlot<T>* q;
    uint32_t* qseq;
    uint32_t qmask;
    BigSeg<T>* big;
    uint32_t* tilecnt;
    BuildCtl* ctl;
    typename Tr::Key* rootkeys;     // 12 keys: AB min3 max3, CB min3 max3
    BuildStatus* status;
    uint32_t n;
    unsigned long long timeout_ns;
    BTask<T>* small;                // deferred small ranges (capacity n/2 + 1)
    uint32_t small_max;             // ranges up to this many shapes are deferred (0 = none): pays off in the throughput regime only
    uint4* trace;                   // optional task log (BVHGPU_TRACE=file): {kind<<28|count, node/sid, t0_ns, t1_ns}
    uint32_t trace_cap;
    uint32_t gang_budget;           // warps that gangs may hold at any time (0 = gang mode off)
    uint32_t sdiv;                  // granularity of the multi-warp segment state / tile-count slots: GT with gangs, else TILE
    uint32_t opt_subtree;           // ranges <= 32 shapes: in-register subtree builder (1) or warp-per-node (0)
};

// ------------------------------------------------------------------------------------------------
template <class S> __device__ __forceinline__ void load_struct_cg(S& dst, const S* src) {
    static_assert(sizeof(S) % 16 == 0, "16-byte granules");
    uint4* d = reinterpret_cast<uint4*>(&dst);
    const uint4* s = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(S) / 16); ++i) d[i] = __ldcg(s + i);
}
template <class S> __device__ __forceinline__ void store_struct_cg(S* dst, const S& src) {
    static_assert(sizeof(S) % 16 == 0, "16-byte granules");
    uint4* d = reinterpret_cast<uint4*>(dst);
    const uint4* s = reinterpret_cast<const uint4*>(&src);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(S) / 16); ++i) __stcg(d + i, s[i]);
}

template <class T> __device__ __forceinline__ bool key_is_min(int e) { const int k = e % 12; return k < 3 || (k >= 6 && k < 9); }

template <class T> __device__ __forceinline__ void zero_bins(WarpScratch<T>* ws) {
    using Tr = Traits<T>;
    for (int e = lane_id(); e < 72; e += 32) ws->keys[e] = key_is_min<T>(e) ? Tr::KEY_POS_INF : Tr::KEY_NEG_INF;
    if (lane_id() < 8) ws->cnt[lane_id()] = 0;
    __syncwarp();
}

// Split axis and extent: Aabb::largest_axis on the centroid bounds (aabb_impl.rs:594-596, first
// maximum wins) and split_axis_size (bvh_node.rs:107-108).
template <class T> __device__ __forceinline__ void split_axis(const BTask<T>& t, int& axis, T& ext, T& cbmin) {
    const T sx = sub_rn(t.cb[3], t.cb[0]), sy = sub_rn(t.cb[4], t.cb[1]), sz = sub_rn(t.cb[5], t.cb[2]);
    axis = 0; ext = sx; cbmin = t.cb[0];
    if (sy > ext) { axis = 1; ext = sy; cbmin = t.cb[1]; }
    if (sz > ext) { axis = 2; ext = sz; cbmin = t.cb[2]; }
}

// Bucket assignment + Bucket::add_aabb (bvh_node.rs:204-222, utils.rs:78-83) for positions [p0,p1); leaves the six
// buckets of the range in ws->keys / ws->cnt.  Two flavours, chosen by regime (measured, tools/build_sweep.py):
//   bin_range         -- no shared-memory atomics: the lanes of a chunk group combine their own keys per bucket, one
//                        REDUX per (bucket, key) folds the warp, lane k keeps key k of every bucket in registers.  More
//                        instructions, shorter dependent chain: used by the GANG tiles, where a handful of warps per SM
//                        sit on the critical path of a tree level (bin phase 8 -> 4.5 us).
//   bin_range_atomic  -- 13 shared-memory atomics per shape; fewer instructions: used by SEG / queue tiles, where all
//                        warps are busy and issue slots are what is scarce (1.2 M shapes: 2.05 ms vs 2.67 ms with REDUX).
template <class T>
__device__ __forceinline__ void bin_range(const BuildParams<T>& P, WarpScratch<T>* ws, const uint32_t* __restrict__ src,
                                          uint32_t seg_start, uint32_t p0, uint32_t p1, int axis, T cbmin, T ext,
                                          bool degenerate, uint32_t half, bool store_bkt, uint32_t& last_id, int& last_b) {
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    const uint32_t lane = lane_id();
    const T K = sub_rn(T(6), T(0.01));                 // T::from(NUM_BUCKETS) - T::from(0.01), bvh_node.rs:214-215
    constexpr int U = sizeof(T) == 8 ? 2 : 4;          // chunks in flight: index loads, then AABB gathers, then the math
    const bool my_min = key_is_min<T>((int)lane);      // lanes 0..11 own key `lane` of each bucket
    Key acc[6];
    uint32_t cnt[6];
#pragma unroll
    for (int bb = 0; bb < 6; ++bb) { acc[bb] = my_min ? Tr::KEY_POS_INF : Tr::KEY_NEG_INF; cnt[bb] = 0; }
    for (uint32_t base = p0; base < p1; base += 32 * U) {
        uint32_t id[U];
        int bk[U];
        Key kv[U][9];                                  // min3, max3, centre3 as keys
        {
            T mn[U][3], mx[U][3];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t pos = base + 32 * u + lane;
                id[u] = pos < p1 ? __ldcg(src + pos) : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t pos = base + 32 * u + lane;
                if (pos < p1) load_aabb(P.aabb + id[u], mn[u], mx[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t pos = base + 32 * u + lane;
                bk[u] = -1;
                if (pos < p1) {
                    T c[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) c[k] = center1(mn[u][k], mx[u][k]);
                    int b;
                    if (degenerate) {
                        b = (pos - seg_start) < half ? 0 : 1;  // indices.split_at_mut(len / 2), bvh_node.rs:117
                    } else {
                        const T ca = axis == 0 ? c[0] : (axis == 1 ? c[1] : c[2]);
                        const T rel = div_rn(sub_rn(ca, cbmin), ext);
                        b = (int)mul_rn(rel, K);               // to_usize(): truncation toward zero
                        b = b < 0 ? 0 : (b > 5 ? 5 : b);       // inert for tight bounds; keeps memory safe
                    }
                    if (store_bkt) P.bkt[pos] = (uint8_t)b;
                    bk[u] = b;
#pragma unroll
                    for (int k = 0; k < 3; ++k) { kv[u][k] = f2key(mn[u][k]); kv[u][3 + k] = f2key(mx[u][k]); kv[u][6 + k] = f2key(c[k]); }
                }
                if (base + 32 * u < p1) { last_id = id[u]; last_b = bk[u] < 0 ? 0 : bk[u]; }
            }
        }
#pragma unroll
        for (int bb = 0; bb < 6; ++bb) {
            bool in[U];
            uint32_t c = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) { in[u] = bk[u] == bb; c += __popc(__ballot_sync(0xffffffffu, in[u])); }
            if (c == 0) continue;                      // warp-uniform
            cnt[bb] += c;
            Key mine = acc[bb];
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const bool is_min = key_is_min<T>(k);
                const int f = k < 6 ? k : (k < 9 ? k : k - 3);     // keys 6..8 (centroid min) and 9..11 (max) share the centre
                Key v = is_min ? Tr::KEY_POS_INF : Tr::KEY_NEG_INF;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const Key x = kv[u][f];
                    if (in[u]) v = is_min ? (x < v ? x : v) : (x > v ? x : v);
                }
                const Key r = is_min ? warp_min_key(v) : warp_max_key(v);
                if ((int)lane == k) mine = is_min ? (r < mine ? r : mine) : (r > mine ? r : mine);
            }
            acc[bb] = mine;
        }
    }
    if (lane < 12) {
#pragma unroll
        for (int bb = 0; bb < 6; ++bb) ws->keys[bb * 12 + lane] = acc[bb];
    }
    if (lane < 6) {
        uint32_t c = cnt[0];
#pragma unroll
        for (int bb = 1; bb < 6; ++bb) c = (int)lane == bb ? cnt[bb] : c;
        ws->cnt[lane] = c;
    }
    __syncwarp();
}

// The same with shared-memory atomics on ws->keys / ws->cnt (zeroed by the caller): fewer instructions, more LSU time.
template <class T>
__device__ __forceinline__ void bin_range_atomic(const BuildParams<T>& P, WarpScratch<T>* ws, const uint32_t* __restrict__ src,
                                          uint32_t seg_start, uint32_t p0, uint32_t p1, int axis, T cbmin, T ext,
                                          bool degenerate, uint32_t half, bool store_bkt, uint32_t& last_id, int& last_b) {
    const uint32_t lane = lane_id();
    const T K = sub_rn(T(6), T(0.01));                 // T::from(NUM_BUCKETS) - T::from(0.01), bvh_node.rs:214-215
    constexpr int U = sizeof(T) == 8 ? 2 : 4;          // chunks in flight: index loads, then AABB gathers, then the math
    for (uint32_t base = p0; base < p1; base += 32 * U) {
        uint32_t id[U];
        T mn[U][3], mx[U][3];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t pos = base + 32 * u + lane;
            id[u] = pos < p1 ? __ldcg(src + pos) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t pos = base + 32 * u + lane;
            if (pos < p1) load_aabb(P.aabb + id[u], mn[u], mx[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t pos = base + 32 * u + lane;
            if (base + 32 * u >= p1) break;            // warp-uniform
            int b = 0;
            if (pos < p1) {
                T c[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) c[k] = center1(mn[u][k], mx[u][k]);
                if (degenerate) {
                    b = (pos - seg_start) < half ? 0 : 1;  // indices.split_at_mut(len / 2), bvh_node.rs:117
                } else {
                    const T ca = axis == 0 ? c[0] : (axis == 1 ? c[1] : c[2]);
                    const T rel = div_rn(sub_rn(ca, cbmin), ext);
                    b = (int)mul_rn(rel, K);               // to_usize(): truncation toward zero
                    b = b < 0 ? 0 : (b > 5 ? 5 : b);       // inert for tight bounds; keeps memory safe
                }
                if (store_bkt) P.bkt[pos] = (uint8_t)b;
                typename Traits<T>::Key* kb = ws->keys + b * 12;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    atomicMin(kb + k, f2key(mn[u][k]));
                    atomicMax(kb + 3 + k, f2key(mx[u][k]));
                    atomicMin(kb + 6 + k, f2key(c[k]));
                    atomicMax(kb + 9 + k, f2key(c[k]));
                }
                atomicAdd(&ws->cnt[b], 1u);
            }
            last_id = id[u];
            last_b = b;
        }
    }
    __syncwarp();
}

// The 5 candidate splits (bvh_node.rs:231-247): lane s evaluates split s; the strict `<`
// first-wins argmin is replayed sequentially over the 5 costs.  Leaves the chosen child bounds in
// ws->child (lab, lcb, rab, rcb) -- Aabb::empty() for all four when no cost is < +inf -- and
// returns the left shape count.
template <class T>
__device__ __forceinline__ uint32_t split_eval(WarpScratch<T>* ws, const T parent_ab[6], bool degenerate) {
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    const uint32_t lane = lane_id();
    const int s = degenerate ? 0 : (lane < 5 ? (int)lane : 4);
    Key L[12], R[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) { L[k] = R[k] = key_is_min<T>(k) ? Tr::KEY_POS_INF : Tr::KEY_NEG_INF; }
    uint32_t nL = 0, nR = 0;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const bool inL = b <= s;
        const uint32_t c = ws->cnt[b];
        if (inL) nL += c; else nR += c;
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const Key v = ws->keys[b * 12 + k];
            if (key_is_min<T>(k)) { if (inL) L[k] = v < L[k] ? v : L[k]; else R[k] = v < R[k] ? v : R[k]; }
            else                  { if (inL) L[k] = v > L[k] ? v : L[k]; else R[k] = v > R[k] ? v : R[k]; }
        }
    }
    T lmn[3], lmx[3], rmn[3], rmx[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { lmn[k] = key2f(L[k]); lmx[k] = key2f(L[3 + k]); rmn[k] = key2f(R[k]); rmx[k] = key2f(R[3 + k]); }
    // cost = (T(nL)*SA(L) + T(nR)*SA(R)) / SA(aabb_bounds), bvh_node.rs:236-238
    const T cost = div_rn(add_rn(mul_rn((T)nL, surface_area(lmn, lmx)), mul_rn((T)nR, surface_area(rmn, rmx))),
                          surface_area(parent_ab, parent_ab + 3));
    int best = 0;
    bool found = degenerate;
    if (!degenerate) {
        T min_cost = Tr::inf();
#pragma unroll
        for (int s2 = 0; s2 < 5; ++s2) {
            const T c = __shfl_sync(0xffffffffu, cost, s2);
            if (c < min_cost) { best = s2; min_cost = c; found = true; }
        }
    }
    if ((int)lane == best) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            ws->child[k]      = found ? key2f(L[k])     : (k < 3 ? Tr::inf() : -Tr::inf());
            ws->child[6 + k]  = found ? key2f(L[6 + k]) : (k < 3 ? Tr::inf() : -Tr::inf());
            ws->child[12 + k] = found ? key2f(R[k])     : (k < 3 ? Tr::inf() : -Tr::inf());
            ws->child[18 + k] = found ? key2f(R[6 + k]) : (k < 3 ? Tr::inf() : -Tr::inf());
        }
    }
    const uint32_t nl = __shfl_sync(0xffffffffu, nL, best);
    __syncwarp();
    return nl;
}

// Stable 6-way partition of positions [p0,p1) (bvh_node.rs:250-272).  base[b] = next free
// destination of bucket b (warp-uniform registers, advanced here).
template <class T>
__device__ __forceinline__ void scatter_range(const BuildParams<T>& P, const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                              uint32_t p0, uint32_t p1, uint32_t base[6], bool use_cached,
                                              uint32_t cached_id, int cached_b) {
    const uint32_t lane = lane_id();
    const uint32_t lt = lanemask_lt();
    constexpr int U = 4;
    for (uint32_t bpos = p0; bpos < p1; bpos += 32 * U) {
        uint32_t id[U];
        int bk[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t pos = bpos + 32 * u + lane;
            id[u] = cached_id; bk[u] = cached_b;
            if (!use_cached && pos < p1) { id[u] = __ldcg(src + pos); bk[u] = (int)__ldcg(P.bkt + pos); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (bpos + 32 * u >= p1) break;            // warp-uniform
            const uint32_t pos = bpos + 32 * u + lane;
            const bool valid = pos < p1;
            uint32_t dest = 0;
#pragma unroll
            for (int bb = 0; bb < 6; ++bb) {
                const uint32_t m = __ballot_sync(0xffffffffu, valid && bk[u] == bb);
                if (bk[u] == bb) dest = base[bb] + __popc(m & lt);
                base[bb] += __popc(m);
            }
            if (valid) __stcg(dst + dest, id[u]);
        }
    }
    __syncwarp();
}

template <class T> __device__ __forceinline__ void write_leaf(const BuildParams<T>& P, uint32_t node, uint32_t parent, uint32_t shape, uint32_t start) {
    using Tr = Traits<T>;
    typename Tr::Node nd;
    nd.parent = parent; nd.child_l = BVH_INVALID; nd.child_r = BVH_INVALID; nd.shape = shape;
#pragma unroll
    for (int k = 0; k < 3; ++k) { nd.l_aabb.min[k] = nd.r_aabb.min[k] = Tr::inf(); nd.l_aabb.max[k] = nd.r_aabb.max[k] = -Tr::inf(); }
    store_struct_cg(P.nodes + node, nd);
    P.node_index[shape] = node;                 // Shapes::set_node_index, bvh_node.rs:103
    P.node_start[node] = start;
}

// Writes the inner node (bvh_node.rs:138-151), emits leaves for single-shape children (:95-104) and
// returns the children that still have to be split in out[0..nc).
template <class T>
__device__ __forceinline__ int finish_node(const BuildParams<T>& P, WarpScratch<T>* ws, const BTask<T>& t, uint32_t nl,
                                           bool moved, BTask<T> out[2], uint32_t& leaves) {
    using Tr = Traits<T>;
    const uint32_t buf = t.parent_buf >> 31, parent = t.parent_buf & 0x7FFFFFFFu;
    const uint32_t nbuf = moved ? (buf ^ 1u) : buf;
    const uint32_t cl = t.node + 1, cr = cl + 2 * nl - 1;
    if (nl == 0u || nl >= t.count) {
        // An empty side cannot come out of a split of NaN-free shapes (buckets 0 and 5 are never empty: tight centroid bounds).
        // It would recurse on the same range forever; raise a device error instead (the workers see ctl->error and leave).
        if (lane_id() == 0) atomicCAS(&P.ctl->error, 0u, (uint32_t)BVHGPU_ERR_INTERNAL);
        leaves = 0;
        return 0;
    }
    if (lane_id() == 0) {
        typename Tr::Node nd;
        nd.parent = parent; nd.child_l = cl; nd.child_r = cr; nd.shape = t.count;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            nd.l_aabb.min[k] = ws->child[k];      nd.l_aabb.max[k] = ws->child[3 + k];
            nd.r_aabb.min[k] = ws->child[12 + k]; nd.r_aabb.max[k] = ws->child[15 + k];
        }
        store_struct_cg(P.nodes + t.node, nd);
        P.node_start[t.node] = t.start;
    }
    int nc = 0;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const uint32_t cstart = side ? t.start + nl : t.start;
        const uint32_t ccount = side ? t.count - nl : nl;
        const uint32_t cnode = side ? cr : cl;
        if (ccount == 1) {
            if (lane_id() == 0) write_leaf(P, cnode, t.node, __ldcg(P.idx[nbuf] + cstart), cstart);
            leaves += 1;
        } else if (ccount <= P.small_max) {
            // The bottom of the tree holds most of the nodes, and a whole warp per 2..16-shape node wastes it: hand the
            // range to small_subtrees_kernel, where one thread replays the reference recursion for it.
            if (lane_id() == 0) {
                BTask<T> c;
                c.start = cstart; c.count = ccount; c.node = cnode; c.parent_buf = t.node | (nbuf << 31);
#pragma unroll
                for (int k = 0; k < 6; ++k) { c.ab[k] = ws->child[side * 12 + k]; c.cb[k] = ws->child[side * 12 + 6 + k]; }
                store_struct_cg(P.small + atomicAdd(&P.ctl->small_count, 1u), c);
            }
            leaves += ccount;                 // accounted for here; the second kernel writes them
        } else {
            BTask<T>& c = out[nc++];
            c.start = cstart; c.count = ccount; c.node = cnode; c.parent_buf = t.node | (nbuf << 31);
#pragma unroll
            for (int k = 0; k < 6; ++k) { c.ab[k] = ws->child[side * 12 + k]; c.cb[k] = ws->child[side * 12 + 6 + k]; }
        }
    }
    return nc;
}

// ---- queue ---------------------------------------------------------------------------------------
template <class T> __device__ __forceinline__ void push_seg(const BuildParams<T>& P, const BTask<T>& t) {
    if (lane_id() == 0) {
        const uint32_t tk = atomicAdd(&P.ctl->tail, 1u);
        const uint32_t slot = tk & P.qmask;
        store_struct_cg(&P.q[slot].t, t);
        __stcg(reinterpret_cast<uint4*>(&P.q[slot].kind), make_uint4(KIND_SEG, 0u, 0u, 0u));
        __threadfence();
        st_release(P.qseq + slot, tk + 1u);
    }
    __syncwarp();
}
template <class T> __device__ __forceinline__ void push_tiles(const BuildParams<T>& P, uint32_t kind, uint32_t sid, uint32_t tiles, const BTask<T>& t) {
    uint32_t base = 0;
    if (lane_id() == 0) base = atomicAdd(&P.ctl->tail, tiles);
    base = __shfl_sync(0xffffffffu, base, 0);
    // payloads first, ONE fence, then the sequence words: a fence per slot would serialise the publication
    for (uint32_t j = lane_id(); j < tiles; j += 32) {
        const uint32_t slot = (base + j) & P.qmask;
        store_struct_cg(&P.q[slot].t, t);              // the task travels in the slot: no dependent load of the segment state
        __stcg(reinterpret_cast<uint4*>(&P.q[slot].kind), make_uint4(kind, sid, j, 0u));
    }
    __threadfence();
    for (uint32_t j = lane_id(); j < tiles; j += 32) {
        const uint32_t tk = base + j;
        st_release(P.qseq + (tk & P.qmask), tk + 1u);
    }
    __syncwarp();
}

__device__ __forceinline__ uint32_t tile_slot(uint32_t sdiv, uint32_t p, bool first) { return 2u * (p / sdiv) + (first ? 1u : 0u); }
// Multi-warp segment state: live segments are >= GANG_MIN > GT shapes long and disjoint, so start / GT is unique among
// them -- except for a gang's parent and child, alive at the same time for a moment, whose starts may share a GT block
// (left child: same start; right child: when the left one is tiny): hence the depth-parity bit.  Queue-mode segments
// are finished before their children exist and use parity 0.
__device__ __forceinline__ uint32_t state_index(uint32_t sdiv, uint32_t start, uint32_t par) { return 2u * (start / sdiv) + (par & 1u); }

template <class T> __device__ __forceinline__ void init_state(BigSeg<T>* B) {
    using Tr = Traits<T>;
    const uint32_t lane = lane_id();
    for (int e = lane; e < 72; e += 32) __stcg(&B->keys[e], key_is_min<T>(e) ? Tr::KEY_POS_INF : Tr::KEY_NEG_INF);
    if (lane < 6) __stcg(&B->cnt[lane], 0u);
    if (lane == 6) atomicAdd(&B->epoch, 1u);      // tells a straggler still polling the slot's previous barrier that it is over
    if (lane == 7) __stcg(&B->bin_done, 0u);
    if (lane == 8) __stcg(&B->scat_done, 0u);
}

template <class T> __device__ __forceinline__ void create_big(const BuildParams<T>& P, const BTask<T>& t) {
    const uint32_t sid = state_index(P.sdiv, t.start, 0u);
    const uint32_t tiles = (t.count + TILE - 1) / TILE;
    init_state(P.big + sid);
    __syncwarp();            // push_tiles fences (every lane) before it publishes: the stores above are covered
    push_tiles(P, KIND_BIN, sid, tiles, t);
}

// A gang takes the segment if enough co-resident warps are left in the budget: one warp per GT shapes, all of them
// spinning on the segment's two barriers, so the reservation is what keeps the queue live.
template <class T> __device__ __forceinline__ bool try_create_gang(const BuildParams<T>& P, const BTask<T>& t) {
    if (P.gang_budget == 0u || t.count < GANG_MIN) return false;
    const uint32_t tiles = t.count / GT;
    uint32_t ok = 0;
    if (lane_id() == 0) {
        const uint32_t used = atomicAdd(&P.ctl->gang_used, tiles);
        ok = used + tiles <= P.gang_budget ? 1u : 0u;
        if (!ok) atomicSub(&P.ctl->gang_used, tiles);
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (!ok) return false;
    init_state(P.big + state_index(P.sdiv, t.start, 0u));
    __syncwarp();
    push_tiles(P, KIND_GANG, 0u, tiles, t);
    return true;
}

template <class T>
__device__ __forceinline__ void dispatch_children(const BuildParams<T>& P, BTask<T> ch[2], int nc) {
    for (int i = 0; i < nc; ++i) {
        if (ch[i].count > (uint32_t)TILE) { if (!try_create_gang(P, ch[i])) create_big(P, ch[i]); }
        else push_seg(P, ch[i]);
    }
}


// ---- in-warp subtree: <= 32 shapes, one per lane, whole levels at a time ------------------------------------------------
// Below ~32 shapes a warp per node is 3 us of dependent round trips for a handful of shapes, and that -- not the top of
// the tree -- is most of the nodes.  Here the range is loaded ONCE and every level handles all of its nodes together,
// in registers, with no bins at all: the shapes of a node are first partitioned by bucket (stable, ballots), so that
// candidate split s is a PREFIX of the node's lanes; a segmented prefix scan and a segmented suffix scan of the 12 keys
// then hold, at every bucket boundary, exactly the two Bucket::join_bucket chains the reference folds
// (bvh_node.rs:231-235; min/max are associative and commutative, the result is the same bits).  Empty buckets make
// consecutive candidates identical, and the reference's strict `<` keeps the first: one evaluation per boundary and
// "lowest lane wins" is the same choice.  The index arrays are not written: they are scratch, leaves carry the shape.
template <class T> __device__ __forceinline__ typename Traits<T>::Key shfl_key_up(typename Traits<T>::Key v, uint32_t d) { return __shfl_up_sync(0xffffffffu, v, d); }
template <class T> __device__ __forceinline__ typename Traits<T>::Key shfl_key_down(typename Traits<T>::Key v, uint32_t d) { return __shfl_down_sync(0xffffffffu, v, d); }

template <class T>
__device__ void process_subtree(const BuildParams<T>& P, WarpScratch<T>* ws, const BTask<T>& t, uint32_t& leaves) {
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    const uint32_t lane = lane_id();
    const uint32_t lt = lanemask_lt();
    const T K = sub_rn(T(6), T(0.01));
    // the shape in this lane
    uint32_t id = 0;
    T mn[3] = {T(0), T(0), T(0)}, mx[3] = {T(0), T(0), T(0)};
    // the node it belongs to (identical in all lanes of the node; scount == 0: lane is done)
    uint32_t snode = t.node, sparent = t.parent_buf & 0x7FFFFFFFu, sstart = 0, scount = lane < t.count ? t.count : 0u;
    T ab[6], cb[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { ab[k] = t.ab[k]; cb[k] = t.cb[k]; }
    if (lane < t.count) {
        id = __ldcg(P.idx[t.parent_buf >> 31] + t.start + lane);
        load_aabb(P.aabb + id, mn, mx);
    }
    while (__any_sync(0xffffffffu, scount >= 2u)) {
        const bool active = scount >= 2u;
        // ---- bucket (bvh_node.rs:204-222) or position half (bvh_node.rs:114-124) ----
        int axis = 0;
        T ext = sub_rn(cb[3], cb[0]), cbmin = cb[0];
        { const T sy = sub_rn(cb[4], cb[1]), sz = sub_rn(cb[5], cb[2]);
          if (sy > ext) { axis = 1; ext = sy; cbmin = cb[1]; }
          if (sz > ext) { axis = 2; ext = sz; cbmin = cb[2]; } }
        const bool degenerate = ext < Tr::eps();
        int b = 0;
        if (active) {
            if (degenerate) {
                b = (lane - sstart) < scount / 2 ? 0 : 1;
            } else {
                const T ca = center1(axis == 0 ? mn[0] : (axis == 1 ? mn[1] : mn[2]), axis == 0 ? mx[0] : (axis == 1 ? mx[1] : mx[2]));
                b = (int)mul_rn(div_rn(sub_rn(ca, cbmin), ext), K);
                b = b < 0 ? 0 : (b > 5 ? 5 : b);
            }
        }
        // ---- stable partition by bucket inside each node (bvh_node.rs:250-272) ----
        const uint32_t segmask = !active ? 0u : ((scount >= 32u ? 0xffffffffu : ((1u << scount) - 1u)) << sstart);
        uint32_t dest = lane;
        {
            uint32_t below = 0, rank = 0;
#pragma unroll
            for (int bb = 0; bb < 6; ++bb) {
                const uint32_t m = __ballot_sync(0xffffffffu, active && b == bb) & segmask;
                if (bb < b) below += __popc(m);
                if (bb == b) rank = __popc(m & lt);
            }
            if (active) dest = sstart + below + rank;
        }
        __syncwarp();
        {
            Staged<T>& d = ws->stage[dest];
            d.id = id; d.b = b;
#pragma unroll
            for (int k = 0; k < 3; ++k) { d.mn[k] = mn[k]; d.mx[k] = mx[k]; }
        }
        __syncwarp();
        {
            const Staged<T>& d = ws->stage[lane];
            id = d.id; b = d.b;
#pragma unroll
            for (int k = 0; k < 3; ++k) { mn[k] = d.mn[k]; mx[k] = d.mx[k]; }
        }
        // ---- segmented scans: Pk = join of the node's lanes up to here, Sk = from here on ----
        T Pk[12], Sk[12];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const T c = center1(mn[k], mx[k]);
            Pk[k] = mn[k]; Pk[3 + k] = mx[k]; Pk[6 + k] = c; Pk[9 + k] = c;
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) Sk[k] = Pk[k];
        const uint32_t send = sstart + scount;
        const uint32_t longest = __reduce_max_sync(0xffffffffu, active ? scount : 0u);     // scan steps: log2 of the longest node
#pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1) {
            if (d >= longest) break;
            const bool okp = active && lane >= sstart + d, oks = active && lane + d < send;
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const T up = __shfl_up_sync(0xffffffffu, Pk[k], d), dn = __shfl_down_sync(0xffffffffu, Sk[k], d);
                if (key_is_min<T>(k)) { if (okp) Pk[k] = min_t(up, Pk[k]); if (oks) Sk[k] = min_t(dn, Sk[k]); }
                else                  { if (okp) Pk[k] = max_t(up, Pk[k]); if (oks) Sk[k] = max_t(dn, Sk[k]); }
            }
        }
        // ---- candidate boundaries: between this lane and the next, where the bucket changes ----
        const int b_next = __shfl_down_sync(0xffffffffu, b, 1);
        const uint32_t o1 = lane - sstart + 1u;                     // left count of the boundary after this lane
        const bool cand = active && o1 < scount && (degenerate ? o1 == scount / 2 : b_next != b);
        T Rn[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) Rn[k] = __shfl_down_sync(0xffffffffu, Sk[k], 1);
        T cost = Tr::inf();
        if (cand) {
            cost = div_rn(add_rn(mul_rn((T)o1, surface_area(Pk, Pk + 3)), mul_rn((T)(scount - o1), surface_area(Rn, Rn + 3))),
                          surface_area(ab, ab + 3));                // bvh_node.rs:236-238
        }
        // first-wins argmin of the node's candidates (bvh_node.rs:239-246) on order-preserving keys; a candidate only
        // counts if cost < +inf (NaN never does)
        const bool valid = cand && cost < Tr::inf();
        const Key ck = valid ? f2key(cost) : ~Key(0);
        {
            Key m = ck;
#pragma unroll
            for (uint32_t d = 1; d < 32; d <<= 1) {
                if (d >= longest) break;
                const Key up = shfl_key_up<T>(m, d);
                if (active && lane >= sstart + d) m = up < m ? up : m;
            }
            const uint32_t lastl = active ? send - 1u : lane;
            m = __shfl_sync(0xffffffffu, m, lastl);
            const uint32_t win = __ballot_sync(0xffffffffu, valid && ck == m) & segmask;
            const uint32_t cm = __ballot_sync(0xffffffffu, cand) & segmask;
            bool found = win != 0u;
            uint32_t q = found ? (uint32_t)__ffs(win) - 1u : (cm ? (uint32_t)__ffs(cm) - 1u : sstart);
            if (degenerate) found = true;                            // the half split is unconditional; q is its only candidate
            if (!active) q = lane;
            // ---- children (bvh_node.rs:126-151): lane q holds the left join, lane q + 1 the right one ----
            const uint32_t nl = q - sstart + 1u;
            if (active && lane == q) {
                typename Tr::Node nd;
                nd.parent = sparent; nd.child_l = snode + 1u; nd.child_r = snode + 2u * nl; nd.shape = scount;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    nd.l_aabb.min[k] = found ? Pk[k] : Tr::inf();  nd.l_aabb.max[k] = found ? Pk[3 + k] : -Tr::inf();
                    nd.r_aabb.min[k] = found ? Rn[k] : Tr::inf();  nd.r_aabb.max[k] = found ? Rn[3 + k] : -Tr::inf();
                }
                store_struct_cg(P.nodes + snode, nd);
                P.node_start[snode] = t.start + sstart;
            }
            const bool right = lane > q;
            const uint32_t srcl = !active ? lane : (right ? q + 1u : q);
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const T pub = lane == q ? Pk[k] : Sk[k];
                const T got = __shfl_sync(0xffffffffu, pub, srcl);
                const T v = found ? got : (key_is_min<T>(k) ? Tr::inf() : -Tr::inf());
                if (k < 6) ab[k] = v; else cb[k - 6] = v;
            }
            if (active) {
                sparent = snode;
                snode = right ? snode + 2u * nl : snode + 1u;
                const uint32_t ccount = right ? scount - nl : nl;
                sstart = right ? q + 1u : sstart;
                scount = ccount;
                if (ccount == 1u) {                                  // bvh_node.rs:95-104
                    write_leaf(P, snode, sparent, id, t.start + lane);
                    scount = 0;
                }
            }
        }
    }
    __syncwarp();
    leaves += t.count;
}

// ---- SEG: one warp owns the whole range; depth-first continuation ------------------------------------
template <class T>
__device__ void process_seg(const BuildParams<T>& P, WarpScratch<T>* ws, BTask<T> t, uint32_t& leaves) {
    using Tr = Traits<T>;
    int sp = 0;
    for (;;) {
        if (t.count <= SUBW && P.opt_subtree) {
            process_subtree(P, ws, t, leaves);
            if (sp == 0) return;
            --sp;
            t = ws->stack[sp];
            __syncwarp();
            continue;
        }
        int axis; T ext, cbmin;
        split_axis(t, axis, ext, cbmin);
        const bool degenerate = ext < Tr::eps();            // bvh_node.rs:114
        const uint32_t buf = t.parent_buf >> 31;
        const uint32_t p0 = t.start, p1 = t.start + t.count;
        const bool single = t.count <= 32;
        uint32_t cid = 0; int cb = 0;
        zero_bins(ws);
        bin_range_atomic(P, ws, P.idx[buf], t.start, p0, p1, axis, cbmin, ext, degenerate, t.count / 2, !single && !degenerate, cid, cb);
        const uint32_t nl = split_eval(ws, t.ab, degenerate);
        if (!degenerate) {
            uint32_t base[6];
            uint32_t run = t.start;
#pragma unroll
            for (int b = 0; b < 6; ++b) { base[b] = run; run += ws->cnt[b]; }
            scatter_range(P, P.idx[buf], P.idx[buf ^ 1u], p0, p1, base, single, cid, cb);
        }
        BTask<T> ch[2];
        const int nc = finish_node(P, ws, t, nl, !degenerate, ch, leaves);
        if (nc == 2) {
            if (ch[1].count <= LOCAL_MAX && sp < LOCAL_STACK) {
                if (lane_id() == 0) ws->stack[sp] = ch[1];
                ++sp;
                __syncwarp();
            } else {
                push_seg(P, ch[1]);
            }
            t = ch[0];
        } else if (nc == 1) {
            t = ch[0];
        } else {
            if (sp == 0) return;
            --sp;
            t = ws->stack[sp];
            __syncwarp();
        }
    }
}

// ---- BIN tile of a multi-warp segment -------------------------------------------------------------
template <class T>
__device__ void finish_big(const BuildParams<T>& P, WarpScratch<T>* ws, const BTask<T>& t, uint32_t nl, bool moved, uint32_t& leaves) {
    BTask<T> ch[2];
    const int nc = finish_node(P, ws, t, nl, moved, ch, leaves);
    dispatch_children(P, ch, nc);
}

template <class T>
__device__ void process_bin_tile(const BuildParams<T>& P, WarpScratch<T>* ws, const BTask<T>& t, uint32_t sid, uint32_t k, uint32_t& leaves) {
    using Tr = Traits<T>;
    BigSeg<T>* B = P.big + sid;
    const uint32_t lane = lane_id();
    const uint32_t tiles = (t.count + TILE - 1) / TILE;
    int axis; T ext, cbmin;
    split_axis(t, axis, ext, cbmin);
    const bool degenerate = ext < Tr::eps();
    const uint32_t buf = t.parent_buf >> 31;
    const uint32_t p0 = t.start + k * TILE;
    const uint32_t pend = t.start + t.count;
    const uint32_t p1 = p0 + TILE < pend ? p0 + TILE : pend;
    uint32_t cid; int cb;
    zero_bins(ws);
    bin_range_atomic(P, ws, P.idx[buf], t.start, p0, p1, axis, cbmin, ext, degenerate, t.count / 2, !degenerate, cid, cb);
    // flush this tile's buckets into the segment's global buckets
    for (int e = lane; e < 72; e += 32) {
        const typename Tr::Key v = ws->keys[e];
        if (key_is_min<T>(e)) { if (v != Tr::KEY_POS_INF) atomicMin(&B->keys[e], v); }
        else                  { if (v != Tr::KEY_NEG_INF) atomicMax(&B->keys[e], v); }
    }
    const uint32_t slot = tile_slot(P.sdiv, p0, k == 0);
    if (lane < 6) {
        const uint32_t c = ws->cnt[lane];
        __stcg(&P.tilecnt[slot * 8 + lane], c);
        if (c) atomicAdd(&B->cnt[lane], c);
    }
    __threadfence();
    __syncwarp();
    uint32_t last = 0;
    if (lane == 0) last = (atomicAdd(&B->bin_done, 1u) == tiles - 1) ? 1u : 0u;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    // ---- last tile: choose the split for the whole segment ----
    for (int e = lane; e < 72; e += 32) ws->keys[e] = __ldcg(&B->keys[e]);
    if (lane < 6) ws->cnt[lane] = __ldcg(&B->cnt[lane]);
    __syncwarp();
    const uint32_t nl = split_eval(ws, t.ab, degenerate);
    if (degenerate) {                       // halves stay in place: children can be created right away
        finish_big(P, ws, t, nl, false, leaves);
        return;
    }
    uint32_t run = t.start;
    uint32_t basev = 0;
#pragma unroll
    for (int b = 0; b < 6; ++b) { if ((int)lane == b) basev = run; run += ws->cnt[b]; }
    if (lane < 6) __stcg(&B->base[lane], basev);
    if (lane == 6) __stcg(&B->nl, nl);
    if (lane < 24) __stcg(&B->child[lane], ws->child[lane]);
    // exclusive prefix of the per-tile bucket counts (keeps the multi-warp partition stable).  Blocked: lane L owns
    // tiles [L*m, (L+1)*m); all loads of a pass are independent, so they pipeline instead of forming a latency chain.
    {
        const uint32_t m = (tiles + 31) / 32;
        uint32_t sum[6] = {0, 0, 0, 0, 0, 0};
        for (uint32_t q = 0; q < m; ++q) {
            const uint32_t j = lane * m + q;
            if (j < tiles) {
                const uint32_t sj = tile_slot(P.sdiv, t.start + j * TILE, j == 0);
                const uint4 a = __ldcg(reinterpret_cast<const uint4*>(P.tilecnt + sj * 8));
                const uint2 c = __ldcg(reinterpret_cast<const uint2*>(P.tilecnt + sj * 8 + 4));
                sum[0] += a.x; sum[1] += a.y; sum[2] += a.z; sum[3] += a.w; sum[4] += c.x; sum[5] += c.y;
            }
        }
        uint32_t run[6];
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            uint32_t incl = sum[b];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += v; }
            run[b] = incl - sum[b];
        }
        for (uint32_t q = 0; q < m; ++q) {
            const uint32_t j = lane * m + q;
            if (j < tiles) {
                const uint32_t sj = tile_slot(P.sdiv, t.start + j * TILE, j == 0);
                const uint4 a = __ldcg(reinterpret_cast<const uint4*>(P.tilecnt + sj * 8));
                const uint2 c = __ldcg(reinterpret_cast<const uint2*>(P.tilecnt + sj * 8 + 4));
                __stcg(reinterpret_cast<uint4*>(P.tilecnt + sj * 8), make_uint4(run[0], run[1], run[2], run[3]));
                __stcg(reinterpret_cast<uint2*>(P.tilecnt + sj * 8 + 4), make_uint2(run[4], run[5]));
                run[0] += a.x; run[1] += a.y; run[2] += a.z; run[3] += a.w; run[4] += c.x; run[5] += c.y;
            }
        }
    }
    __syncwarp();            // ordered by the fence inside push_tiles
    push_tiles(P, KIND_SCATTER, sid, tiles, t);
}

template <class T>
__device__ void process_scatter_tile(const BuildParams<T>& P, WarpScratch<T>* ws, const BTask<T>& t, uint32_t sid, uint32_t k, uint32_t& leaves) {
    BigSeg<T>* B = P.big + sid;
    const uint32_t lane = lane_id();
    const uint32_t tiles = (t.count + TILE - 1) / TILE;
    const uint32_t buf = t.parent_buf >> 31;
    const uint32_t p0 = t.start + k * TILE;
    const uint32_t pend = t.start + t.count;
    const uint32_t p1 = p0 + TILE < pend ? p0 + TILE : pend;
    const uint32_t slot = tile_slot(P.sdiv, p0, k == 0);
    uint32_t base[6];
#pragma unroll
    for (int b = 0; b < 6; ++b) base[b] = __ldcg(&B->base[b]) + __ldcg(&P.tilecnt[slot * 8 + b]);
    scatter_range(P, P.idx[buf], P.idx[buf ^ 1u], p0, p1, base, false, 0u, 0);
    __threadfence();
    __syncwarp();
    uint32_t last = 0;
    if (lane == 0) last = (atomicAdd(&B->scat_done, 1u) == tiles - 1) ? 1u : 0u;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    const uint32_t nl = __ldcg(&B->nl);
    if (lane < 24) ws->child[lane] = __ldcg(&B->child[lane]);
    __syncwarp();
    finish_big(P, ws, t, nl, true, leaves);
}


// ---- GANG: the top of the tree, level by level, by a set of co-resident warps -----------------------------------------
// Queue hops (publish -> poll -> pop, ~3 us each, two per level plus the last-tile epilogues) dominate the latency of
// the first levels, where there is one segment and nothing else to do.  A gang is `count / GT` warps that stay with
// the segment: each bins its own GT shapes, ALL of them evaluate the split after barrier 1 (redundantly -- no
// broadcast), each computes the stable-partition prefix of its own tile from the per-tile counts and scatters it, and
// after barrier 2 the warps re-map themselves onto the children: warp k takes tile k of the left child while
// k < tiles(left), else tile k - tiles(left) of the right child, else it leaves.  floor() tiling makes
// tiles(left) + tiles(right) <= tiles, so a gang only ever shrinks and never waits for a warp that is not running.
template <class T>
__device__ __forceinline__ bool gang_barrier(const BuildParams<T>& P, uint32_t* ctr, uint32_t tiles, const uint32_t* epoch, uint32_t e0) {
    __threadfence();
    __syncwarp();
    uint32_t ok = 1;
    if (lane_id() == 0) {
        atomicAdd(ctr, 1u);
        uint32_t spins = 0, ns = 32;
        // The state slot is recycled two levels down (same start, same node parity) by warps that may be that far ahead
        // of this one once the barrier has opened: a changed epoch means exactly that.
        while (ld_relaxed(ctr) < tiles && ld_relaxed(epoch) == e0) {
            if ((++spins & 63u) == 0u) {
                if (ld_relaxed(&P.ctl->error) != 0u) { ok = 0; break; }
                if ((spins & 4095u) == 0u &&
                    global_timer_ns() - *(volatile unsigned long long*)&P.ctl->t_start > P.timeout_ns) {
                    atomicExch(&P.ctl->error, (uint32_t)BVHGPU_ERR_TIMEOUT);
                    ok = 0;
                    break;
                }
            }
            __nanosleep(ns);
            if (ns < 256) ns <<= 1;                  // hundreds of warps poll one line: back off
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    __threadfence();
    return ok != 0u;
}

template <class T>
__device__ __forceinline__ void make_child(BTask<T>& c, const WarpScratch<T>* ws, const BTask<T>& t, int side, uint32_t nl, uint32_t nbuf) {
    c.start = side ? t.start + nl : t.start;
    c.count = side ? t.count - nl : nl;
    c.node = side ? t.node + 2 * nl : t.node + 1;
    c.parent_buf = t.node | (nbuf << 31);
#pragma unroll
    for (int k = 0; k < 6; ++k) { c.ab[k] = ws->child[side * 12 + k]; c.cb[k] = ws->child[side * 12 + 6 + k]; }
}

template <class T>
__device__ void process_gang(const BuildParams<T>& P, WarpScratch<T>* ws, BTask<T> t, uint32_t k, uint32_t& leaves) {
    using Tr = Traits<T>;
    const uint32_t lane = lane_id();
    for (uint32_t par = 0;; par ^= 1u) {
        const uint32_t tiles = t.count / GT;
        BigSeg<T>* B = P.big + state_index(P.sdiv, t.start, par);
        int axis; T ext, cbmin;
        split_axis(t, axis, ext, cbmin);
        const bool degenerate = ext < Tr::eps();
        const uint32_t buf = t.parent_buf >> 31;
        const uint32_t p0 = t.start + k * GT;
        const uint32_t p1 = k + 1 == tiles ? t.start + t.count : p0 + GT;
        unsigned long long tr0 = 0;
        if (P.trace && k == 0) tr0 = global_timer_ns();
        zero_bins(ws);
        uint32_t cid; int cb;
        bin_range(P, ws, P.idx[buf], t.start, p0, p1, axis, cbmin, ext, degenerate, t.count / 2, !degenerate, cid, cb);
        for (int e = lane; e < 72; e += 32) {
            const typename Tr::Key v = ws->keys[e];
            if (key_is_min<T>(e)) { if (v != Tr::KEY_POS_INF) atomicMin(&B->keys[e], v); }
            else                  { if (v != Tr::KEY_NEG_INF) atomicMax(&B->keys[e], v); }
        }
        if (lane < 6) {
            const uint32_t c = ws->cnt[lane];
            __stcg(&P.tilecnt[tile_slot(P.sdiv, p0, k == 0) * 8 + lane], c);
            if (c) atomicAdd(&B->cnt[lane], c);
        }
        uint32_t e0 = 0;
        if (lane == 0) e0 = ld_relaxed(&B->epoch);
        unsigned long long tr1 = 0, tr2 = 0, tr3 = 0;
        if (P.trace && k == 0) tr1 = global_timer_ns();
        if (!gang_barrier(P, &B->bin_done, tiles, &B->epoch, e0)) return;
        if (P.trace && k == 0) tr2 = global_timer_ns();
        // ---- every warp: the segment's buckets -> the split ----
        for (int e = lane; e < 72; e += 32) ws->keys[e] = __ldcg(&B->keys[e]);
        if (lane < 6) ws->cnt[lane] = __ldcg(&B->cnt[lane]);
        __syncwarp();
        const uint32_t nl = split_eval(ws, t.ab, degenerate);
        const uint32_t nr = t.count - nl;
        const uint32_t tL = nl >= GANG_MIN ? nl / GT : 0u, tR = nr >= GANG_MIN ? nr / GT : 0u;
        const uint32_t nbuf = degenerate ? buf : (buf ^ 1u);
        if (k == 0) {                           // children's barrier / bucket state, published by barrier 2
            if (tL) init_state(P.big + state_index(P.sdiv, t.start, par ^ 1u));
            if (tR) init_state(P.big + state_index(P.sdiv, t.start + nl, par ^ 1u));
        }
        if (!degenerate) {
            // exclusive prefix of the bucket counts of tiles 0..k-1: what makes the multi-warp partition stable
            uint32_t pre[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll 4
            for (uint32_t j = lane; j < k; j += 32) {
                const uint32_t sj = tile_slot(P.sdiv, t.start + j * GT, j == 0);
                const uint4 a = __ldcg(reinterpret_cast<const uint4*>(P.tilecnt + sj * 8));
                const uint2 c = __ldcg(reinterpret_cast<const uint2*>(P.tilecnt + sj * 8 + 4));
                pre[0] += a.x; pre[1] += a.y; pre[2] += a.z; pre[3] += a.w; pre[4] += c.x; pre[5] += c.y;
            }
            uint32_t base[6];
            uint32_t run = t.start;
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                uint32_t v = pre[b];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                base[b] = run + v;
                run += ws->cnt[b];
            }
            scatter_range(P, P.idx[buf], P.idx[buf ^ 1u], p0, p1, base, false, 0u, 0);
        }
        if (P.trace && k == 0) tr3 = global_timer_ns();
        if (!gang_barrier(P, &B->scat_done, tiles, &B->epoch, e0)) return;
        if (P.trace && k == 0 && lane == 0) {
            const uint32_t slot = P.trace_cap - 1u - atomicAdd(&P.ctl->gang_trace, 2u);
            const unsigned long long base = *(volatile unsigned long long*)&P.ctl->t_start;
            P.trace[slot] = make_uint4((4u << 28) | tiles, t.count, (uint32_t)(tr0 - base), (uint32_t)(global_timer_ns() - base));
            P.trace[slot - 1] = make_uint4((5u << 28) | tiles, (uint32_t)(tr1 - base), (uint32_t)(tr2 - base), (uint32_t)(tr3 - base));
        }
        if (k + 1 == tiles) {                   // the closer: node, leaves, SEG children, give back the warps that leave
            BTask<T> ch[2];
            const int nc = finish_node(P, ws, t, nl, !degenerate, ch, leaves);
            for (int i = 0; i < nc; ++i)
                if (ch[i].count < GANG_MIN) push_seg(P, ch[i]);
            if (lane == 0 && tiles > tL + tR) atomicSub(&P.ctl->gang_used, tiles - tL - tR);
        }
        if (k < tL) {
            BTask<T> c; make_child(c, ws, t, 0, nl, nbuf); t = c;
        } else if (k < tL + tR) {
            BTask<T> c; make_child(c, ws, t, 1, nl, nbuf); t = c; k -= tL;
        } else {
            return;
        }
        __syncwarp();
    }
}

// ---- the persistent kernel -------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, BUILD_MIN_CTAS) build_kernel(BuildParams<T> P) {
    __shared__ WarpScratch<T> wsall[WARPS_PER_CTA];
    WarpScratch<T>* ws = &wsall[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    uint32_t leaves = 0;
    for (;;) {
        uint32_t ticket = 0, stop = 0;
        if (lane == 0) {
            ticket = atomicAdd(&P.ctl->head, 1u);
            const uint32_t* seq = P.qseq + (ticket & P.qmask);
            uint32_t spins = 0, ns = 64;
            for (;;) {
                // relaxed poll: an acquire load would invalidate this SM's whole L1 (CCTL.IVALL) on every spin
                if (ld_relaxed(seq) == ticket + 1u) break;
                if ((spins & 7u) == 7u) {
                    if (ld_relaxed(&P.ctl->leaves_done) >= P.n || ld_relaxed(&P.ctl->error) != 0u) { stop = 1; break; }
                }
                if ((++spins & 1023u) == 0u) {
                    if (global_timer_ns() - *(volatile unsigned long long*)&P.ctl->t_start > P.timeout_ns) {
                        atomicExch(&P.ctl->error, (uint32_t)BVHGPU_ERR_TIMEOUT);
                        stop = 1;
                        break;
                    }
                }
                __nanosleep(ns);
                if (ns < 256) ns <<= 1;
            }
        }
        stop = __shfl_sync(0xffffffffu, stop, 0);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (stop) break;
        __threadfence();
        QSlot<T> s;
        load_struct_cg(s, P.q + (ticket & P.qmask));
        unsigned long long tr0 = 0;
        if (P.trace) tr0 = global_timer_ns();
        if (s.kind == KIND_SEG) process_seg(P, ws, s.t, leaves);
        else if (s.kind == KIND_BIN) process_bin_tile(P, ws, s.t, s.a, s.b, leaves);
        else if (s.kind == KIND_SCATTER) process_scatter_tile(P, ws, s.t, s.a, s.b, leaves);
        else process_gang(P, ws, s.t, s.b, leaves);
        if (P.trace && lane == 0 && ticket < P.trace_cap) {
            const unsigned long long base = *(volatile unsigned long long*)&P.ctl->t_start;
            P.trace[ticket] = make_uint4((s.kind << 28) | (s.kind == KIND_SEG ? s.t.count : s.b), s.kind == KIND_SEG ? s.t.node : s.t.count,
                                         (uint32_t)(tr0 - base), (uint32_t)(global_timer_ns() - base));
        }
        if (leaves) {
            __threadfence();
            if (lane == 0) atomicAdd(&P.ctl->leaves_done, leaves);
            leaves = 0;
        }
    }
}

// ---- small_subtrees_kernel: one THREAD per deferred range of <= SMALL shapes ---------------------------------------
// A literal, sequential replay of BvhNode::prep_build / build_buckets (bvh_node.rs:81-279) on thread-private data:
// the shapes of the range live in local arrays, the recursion is an explicit stack.  32 ranges per warp run
// concurrently, which is ~25x cheaper in issue slots than a warp per node for the 2..16-shape nodes that make up
// most of a tree.
template <class T>
__global__ void __launch_bounds__(128) small_subtrees_kernel(BuildParams<T> P) {
    using Tr = Traits<T>;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P.ctl->small_count) return;
    BTask<T> t;
    load_struct_cg(t, P.small + tid);
    const uint32_t buf = t.parent_buf >> 31;
    uint32_t id[SMALL];
    T mn[SMALL][3], mx[SMALL][3], ct[SMALL][3];
    for (uint32_t i = 0; i < t.count; ++i) {
        id[i] = __ldcg(P.idx[buf] + t.start + i);
        load_aabb(P.aabb + id[i], mn[i], mx[i]);
        for (int k = 0; k < 3; ++k) ct[i][k] = center1(mn[i][k], mx[i][k]);
    }
    struct Frame { uint32_t s, c, node, parent; T ab[6], cb[6]; };
    Frame stack[SMALL];
    int sp = 0;
    {
        Frame& f = stack[sp++];
        f.s = 0; f.c = t.count; f.node = t.node; f.parent = t.parent_buf & 0x7FFFFFFFu;
        for (int k = 0; k < 6; ++k) { f.ab[k] = t.ab[k]; f.cb[k] = t.cb[k]; }
    }
    const T INF = Tr::inf();
    const T K6 = sub_rn(T(6), T(0.01));
    while (sp > 0) {
        const Frame f = stack[--sp];
        // split axis (aabb_impl.rs:594-596) and extent (bvh_node.rs:107-108)
        int axis = 0;
        T ext = sub_rn(f.cb[3], f.cb[0]);
        { const T sy = sub_rn(f.cb[4], f.cb[1]), sz = sub_rn(f.cb[5], f.cb[2]);
          if (sy > ext) { axis = 1; ext = sy; }
          if (sz > ext) { axis = 2; ext = sz; } }
        const T cbmin = f.cb[axis];
        T L[12], R[12];                                   // chosen children: aabb min3 max3, centroid min3 max3
        for (int k = 0; k < 12; ++k) { const bool isMin = (k % 6) < 3; L[k] = R[k] = isMin ? INF : -INF; }
        uint32_t nl;
        if (ext < Tr::eps()) {                            // bvh_node.rs:114-124: halve by position
            nl = f.c / 2;
            for (uint32_t i = 0; i < f.c; ++i) {
                T* D = i < nl ? L : R;
                const uint32_t j = f.s + i;
                for (int k = 0; k < 3; ++k) {
                    D[k] = mn[j][k] < D[k] ? mn[j][k] : D[k];         D[3 + k] = mx[j][k] > D[3 + k] ? mx[j][k] : D[3 + k];
                    D[6 + k] = ct[j][k] < D[6 + k] ? ct[j][k] : D[6 + k]; D[9 + k] = ct[j][k] > D[9 + k] ? ct[j][k] : D[9 + k];
                }
            }
        } else {                                          // build_buckets, bvh_node.rs:183-279
            T bk[6][12];
            uint32_t bn[6];
            uint8_t bid[SMALL];
            for (int b = 0; b < 6; ++b) { bn[b] = 0; for (int k = 0; k < 12; ++k) bk[b][k] = ((k % 6) < 3) ? INF : -INF; }
            for (uint32_t i = 0; i < f.c; ++i) {
                const uint32_t j = f.s + i;
                int b = (int)mul_rn(div_rn(sub_rn(ct[j][axis], cbmin), ext), K6);
                b = b < 0 ? 0 : (b > 5 ? 5 : b);
                bid[i] = (uint8_t)b;
                bn[b]++;
                for (int k = 0; k < 3; ++k) {
                    bk[b][k] = mn[j][k] < bk[b][k] ? mn[j][k] : bk[b][k];         bk[b][3 + k] = mx[j][k] > bk[b][3 + k] ? mx[j][k] : bk[b][3 + k];
                    bk[b][6 + k] = ct[j][k] < bk[b][6 + k] ? ct[j][k] : bk[b][6 + k]; bk[b][9 + k] = ct[j][k] > bk[b][9 + k] ? ct[j][k] : bk[b][9 + k];
                }
            }
            const T sap = surface_area(f.ab, f.ab + 3);
            int best = 0;
            bool found = false;
            T min_cost = INF;
            for (int s = 0; s < 5; ++s) {                 // bvh_node.rs:231-247
                T l[12], r[12];
                uint32_t cl = 0, cr = 0;
                for (int k = 0; k < 12; ++k) { const bool isMin = (k % 6) < 3; l[k] = r[k] = isMin ? INF : -INF; }
                for (int b = 0; b < 6; ++b) {
                    T* D = b <= s ? l : r;
                    if (b <= s) cl += bn[b]; else cr += bn[b];
                    for (int k = 0; k < 12; ++k) { const bool isMin = (k % 6) < 3; D[k] = isMin ? (bk[b][k] < D[k] ? bk[b][k] : D[k]) : (bk[b][k] > D[k] ? bk[b][k] : D[k]); }
                }
                const T cost = div_rn(add_rn(mul_rn((T)cl, surface_area(l, l + 3)), mul_rn((T)cr, surface_area(r, r + 3))), sap);
                if (cost < min_cost) {
                    best = s; min_cost = cost; found = true;
                    for (int k = 0; k < 12; ++k) { L[k] = l[k]; R[k] = r[k]; }
                }
            }
            (void)found;                                   // not found: L / R stay Aabb::empty(), best = 0 (bvh_node.rs:225-230)
            // stable 6-way partition of the range (bvh_node.rs:250-272)
            uint32_t tid2[SMALL];
            T tmn[SMALL][3], tmx[SMALL][3], tct[SMALL][3];
            uint32_t w = 0;
            nl = 0;
            for (int b = 0; b < 6; ++b) {
                if (b <= best) nl += bn[b];
                for (uint32_t i = 0; i < f.c; ++i) {
                    if (bid[i] != b) continue;
                    const uint32_t j = f.s + i;
                    tid2[w] = id[j];
                    for (int k = 0; k < 3; ++k) { tmn[w][k] = mn[j][k]; tmx[w][k] = mx[j][k]; tct[w][k] = ct[j][k]; }
                    ++w;
                }
            }
            for (uint32_t i = 0; i < f.c; ++i) {
                const uint32_t j = f.s + i;
                id[j] = tid2[i];
                for (int k = 0; k < 3; ++k) { mn[j][k] = tmn[i][k]; mx[j][k] = tmx[i][k]; ct[j][k] = tct[i][k]; }
            }
        }
        const uint32_t cl = f.node + 1, cr = f.node + 2 * nl;
        {
            typename Tr::Node nd;
            nd.parent = f.parent; nd.child_l = cl; nd.child_r = cr; nd.shape = f.c;
            for (int k = 0; k < 3; ++k) { nd.l_aabb.min[k] = L[k]; nd.l_aabb.max[k] = L[3 + k]; nd.r_aabb.min[k] = R[k]; nd.r_aabb.max[k] = R[3 + k]; }
            store_struct_cg(P.nodes + f.node, nd);
            P.node_start[f.node] = t.start + f.s;
        }
        // children: right first onto the stack so that the left is processed next (order is irrelevant for the result)
        for (int side = 1; side >= 0; --side) {
            const uint32_t cs = side ? f.s + nl : f.s, cc = side ? f.c - nl : nl, cn = side ? cr : cl;
            if (cc == 1) { write_leaf(P, cn, f.node, id[cs], t.start + cs); continue; }
            Frame& g = stack[sp++];
            g.s = cs; g.c = cc; g.node = cn; g.parent = f.node;
            const T* S = side ? R : L;
            for (int k = 0; k < 6; ++k) { g.ab[k] = S[k]; g.cb[k] = S[6 + k]; }
        }
    }
}

// ---- prep: ABI layout -> device layout, NaN check, scene bounds (joint_aabb_of_shapes, utils.rs:97-109) ---
template <class T>
__global__ void __launch_bounds__(256) prep_kernel(const typename Traits<T>::Aabb* __restrict__ in, uint32_t n,
                                                   typename Traits<T>::DAabb* __restrict__ out, uint32_t* __restrict__ idx0,
                                                   typename Traits<T>::Key* rootkeys, uint32_t* nan_flag) {
    using Tr = Traits<T>;
    using Key = typename Tr::Key;
    __shared__ Key sk[12];
    if (threadIdx.x < 12) sk[threadIdx.x] = key_is_min<T>(threadIdx.x) ? Tr::KEY_POS_INF : Tr::KEY_NEG_INF;
    __syncthreads();
    Key k[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) k[e] = key_is_min<T>(e) ? Tr::KEY_POS_INF : Tr::KEY_NEG_INF;
    bool nan = false;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const T* p = reinterpret_cast<const T*>(in + i);
        T mn[3], mx[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { mn[c] = p[c]; mx[c] = p[3 + c]; nan |= (mn[c] != mn[c]) | (mx[c] != mx[c]); }
        typename Tr::DAabb d;
#pragma unroll
        for (int c = 0; c < 3; ++c) { d.min[c] = mn[c]; d.max[c] = mx[c]; }
        if constexpr (sizeof(T) == 4) { d.pad0 = 0; d.pad1 = 0; }
        out[i] = d;
        if (idx0) idx0[i] = i;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const T ctr = center1(mn[c], mx[c]);
            const Key kmn = f2key(mn[c]), kmx = f2key(mx[c]), kc = f2key(ctr);
            k[c] = kmn < k[c] ? kmn : k[c];
            k[3 + c] = kmx > k[3 + c] ? kmx : k[3 + c];
            k[6 + c] = kc < k[6 + c] ? kc : k[6 + c];
            k[9 + c] = kc > k[9 + c] ? kc : k[9 + c];
        }
    }
    if (rootkeys) {
#pragma unroll
        for (int e = 0; e < 12; ++e) {
            const Key r = key_is_min<T>(e) ? warp_min_key(k[e]) : warp_max_key(k[e]);
            if (lane_id() == 0) { if (key_is_min<T>(e)) atomicMin(&sk[e], r); else atomicMax(&sk[e], r); }
        }
        __syncthreads();
        if (threadIdx.x < 12) {
            if (key_is_min<T>(threadIdx.x)) atomicMin(&rootkeys[threadIdx.x], sk[threadIdx.x]);
            else atomicMax(&rootkeys[threadIdx.x], sk[threadIdx.x]);
        }
    }
    if (nan) atomicExch(nan_flag, 1u);
}

template <class T>
__global__ void init_keys_kernel(typename Traits<T>::Key* rootkeys, BuildCtl* ctl, BuildStatus* status) {
    using Tr = Traits<T>;
    if (threadIdx.x < 12) rootkeys[threadIdx.x] = key_is_min<T>(threadIdx.x) ? Tr::KEY_POS_INF : Tr::KEY_NEG_INF;
    if (threadIdx.x == 0) {
        ctl->head = ctl->tail = ctl->leaves_done = ctl->error = 0;
        ctl->small_count = 0;
        ctl->gang_used = 0;
        ctl->gang_trace = 0;
        ctl->rebuilt = 0;
        ctl->t_start = 0;
        status->error = status->nan_found = status->tickets = status->leaves_done = 0;
    }
}

template <class T> __global__ void init_root_kernel(BuildParams<T> P) {
    BTask<T> t;
    t.start = 0; t.count = P.n; t.node = 0; t.parent_buf = 0;      // root's parent_index is 0, bvh_impl.rs:80
#pragma unroll
    for (int k = 0; k < 6; ++k) { t.ab[k] = key2f(P.rootkeys[k]); t.cb[k] = key2f(P.rootkeys[6 + k]); }
    if (lane_id() == 0) P.ctl->t_start = global_timer_ns();
    __syncwarp();
    if (P.status->nan_found) {                                      // reference panics on NaN centroids: build nothing
        if (lane_id() == 0) { P.ctl->error = (uint32_t)BVHGPU_ERR_NAN; }
        return;
    }
    if (t.count > (uint32_t)TILE) { if (!try_create_gang(P, t)) create_big(P, t); } else push_seg(P, t);
}

template <class T> __global__ void single_leaf_kernel(BuildParams<T> P) {
    if (threadIdx.x == 0) { write_leaf(P, 0u, 0u, 0u, 0u); P.ctl->leaves_done = 1; }
}

template <class T> __global__ void finish_status_kernel(BuildCtl* ctl, BuildStatus* status, uint32_t n) {
    if (threadIdx.x == 0) {
        status->tickets = ctl->tail;
        status->leaves_done = ctl->leaves_done;
        status->rebuilt = ctl->rebuilt;
        uint32_t e = ctl->error;
        if (status->nan_found) e = (uint32_t)BVHGPU_ERR_NAN;
        if (e == 0 && ctl->leaves_done != n) e = (uint32_t)BVHGPU_ERR_INTERNAL;
        status->error = e;
    }
}

template <class T> __global__ void init_rootkeys_kernel(typename Traits<T>::Key* rootkeys) {
    using Tr = Traits<T>;
    if (threadIdx.x < 12) rootkeys[threadIdx.x] = key_is_min<T>(threadIdx.x) ? Tr::KEY_POS_INF : Tr::KEY_NEG_INF;
}

// ABI -> device AABB layout + NaN check + scene bounds (AABB and centroid) as keys; shared with lbvh.cu.
template <class T>
int prep_only(bvhgpu_ctx* ctx, const typename Traits<T>::Aabb* in_aabbs, uint32_t n, typename Traits<T>::DAabb* out,
              typename Traits<T>::Key* rootkeys, BuildStatus* status) {
    init_rootkeys_kernel<T><<<1, 32, 0, ctx->stream>>>(rootkeys);
    const int blocks = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 8);
    prep_kernel<T><<<blocks, 256, 0, ctx->stream>>>(in_aabbs, n, out, nullptr, rootkeys, &status->nan_found);
    ctx->launches += 2;
    BVH_CUDA_TRY(cudaGetLastError());
    return BVHGPU_OK;
}
template int prep_only<float>(bvhgpu_ctx*, const bvh_aabb3f*, uint32_t, DAabbF*, uint32_t*, BuildStatus*);
template int prep_only<double>(bvhgpu_ctx*, const bvh_aabb3d*, uint32_t, DAabbD*, unsigned long long*, BuildStatus*);

static uint32_t next_pow2(uint64_t v) { uint64_t p = 1; while (p < v) p <<= 1; return (uint32_t)p; }

template <class T>
int convert_aabbs(bvhgpu_ctx* ctx, const typename Traits<T>::Aabb* in_aabbs, uint32_t n,
                  typename Traits<T>::DAabb* out, uint32_t* d_nan_flag) {
    if (n == 0) return BVHGPU_OK;
    const int blocks = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 8);
    prep_kernel<T><<<blocks, 256, 0, ctx->stream>>>(in_aabbs, n, out, nullptr, nullptr, d_nan_flag);
    ctx->launches++;
    BVH_CUDA_TRY(cudaGetLastError());
    return BVHGPU_OK;
}

template <class T>
int build_exact_sah(bvhgpu_ctx* ctx, const typename Traits<T>::Aabb* in_aabbs, uint32_t n, Tree<T>* tree) {
    using Tr = Traits<T>;
    cudaStream_t st = ctx->stream;
    tree->ctx = ctx;
    tree->n = n;
    tree->n_nodes = n ? 2 * n - 1 : 0;
    BVH_TRY(dalloc_t(ctx, &tree->d_status, 1));
    BVH_CUDA_TRY(cudaMemsetAsync(tree->d_status, 0, sizeof(BuildStatus), st));
    if (n == 0) return BVHGPU_OK;
    BVH_TRY(dalloc_t(ctx, &tree->d_aabb, n));
    BVH_TRY(dalloc_t(ctx, &tree->d_nodes, tree->n_nodes));
    BVH_TRY(dalloc_t(ctx, &tree->d_node_index, n));
    BVH_TRY(dalloc_t(ctx, &tree->d_node_start, tree->n_nodes));

    BuildParams<T> P{};
    P.aabb = tree->d_aabb;
    P.nodes = tree->d_nodes;
    P.node_index = tree->d_node_index;
    P.node_start = tree->d_node_start;
    P.n = n;
    P.status = tree->d_status;
    P.timeout_ns = 20ull * 1000ull * 1000ull * 1000ull;
    const uint32_t qcap = next_pow2(std::max<uint64_t>(n, 1024) * 2);
    P.qmask = qcap - 1;
    // launch shape of the persistent kernel, and whether gangs will be used (they need finer-grained segment state)
    int occ = 1;
    BVH_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, build_kernel<T>, WARPS_PER_CTA * 32, 0));
    if (occ < 1) occ = 1;
    // enough warps that every one has ~16 shapes of work, capped by what is co-resident
    uint64_t want = ((uint64_t)n / 16 + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    if (want < 1) want = 1;
    const int grid = (int)std::min<uint64_t>(want, (uint64_t)ctx->sm_count * occ);
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
    // gangs may hold at most half of the co-resident warps: the other half keeps the task queue draining
    P.gang_budget = (ctx->build_gang != 0 && coop) ? (uint32_t)grid * WARPS_PER_CTA / 2u : 0u;
    // Gang warps spin instead of draining the queue: worth it only while the machine would otherwise idle, i.e. when
    // the whole root fits one gang (measured at 1.2 M shapes: 2.7 ms without gangs, 4.5 ms with).
    if (ctx->build_gang < 0 && (uint64_t)n / GT > P.gang_budget) P.gang_budget = 0;
    P.sdiv = P.gang_budget ? GT : TILE;
    const size_t nbig = 2 * ((size_t)n / P.sdiv + 2);
    uint32_t *idx0 = nullptr, *idx1 = nullptr;
    BVH_TRY(dalloc_t(ctx, &idx0, n));
    BVH_TRY(dalloc_t(ctx, &idx1, n));
    BVH_TRY(dalloc_t(ctx, &P.bkt, n));
    BVH_TRY(dalloc_t(ctx, &P.q, qcap));
    BVH_TRY(dalloc_t(ctx, &P.qseq, qcap));
    BVH_TRY(dalloc_t(ctx, &P.big, nbig));
    BVH_TRY(dalloc_t(ctx, &P.tilecnt, nbig * 8));
    BVH_TRY(dalloc_t(ctx, &P.ctl, 1));
    BVH_TRY(dalloc_t(ctx, &P.rootkeys, 12));
    // Deferring the bottom of the tree to the thread-per-range kernel adds that kernel's own latency (~0.1-0.4 ms tail) but
    // removes most warp-per-node work: measured slower below ~0.5 M shapes (120 k: 0.62 -> 0.67 ms), faster above
    // (1.2 M f32: 3.33 -> 2.32 ms, 10 M f64: 48 -> 28 ms).
    // Bottom of the tree (measured, tools/build_sweep.py): f32 -- the in-register subtree builder everywhere (120 k:
    // 0.51 -> 0.35 ms; 1.2 M: 2.05 ms, the same as deferring <= 16-shape ranges to the thread-per-range kernel).
    // f64 -- 64-bit shuffles and DMNMX make the subtree builder the slower one (10 M: 36.7 ms vs 23.5 ms): large f64
    // scenes defer to the thread-per-range kernel instead, whose own ~0.1-0.4 ms tail only pays off above ~0.4 M shapes.
    const bool defer_small = ctx->build_small < 0 ? (sizeof(T) == 8 && n >= 400000u) : ctx->build_small != 0;
    P.small_max = defer_small ? SMALL : 0u;
    P.opt_subtree = ctx->build_subtree < 0 ? !defer_small : ctx->build_subtree != 0;
    BVH_TRY(dalloc_t(ctx, &P.small, P.small_max ? (size_t)n / 2 + 1 : 1));
    P.idx[0] = idx0;
    P.idx[1] = idx1;
    const char* trace_path = getenv("BVHGPU_TRACE");
    if (trace_path && n > 1) {
        P.trace_cap = 4u * n + 4096u;
        BVH_TRY(dalloc_t(ctx, &P.trace, P.trace_cap));
        BVH_CUDA_TRY(cudaMemsetAsync(P.trace, 0, sizeof(uint4) * P.trace_cap, st));
    }

    init_keys_kernel<T><<<1, 32, 0, st>>>(P.rootkeys, P.ctl, P.status);
    ctx->launches++;
    const int pblocks = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 8);
    prep_kernel<T><<<pblocks, 256, 0, st>>>(in_aabbs, n, tree->d_aabb, idx0, P.rootkeys, &P.status->nan_found);
    ctx->launches++;
    if (n == 1) {
        single_leaf_kernel<T><<<1, 32, 0, st>>>(P);
        ctx->launches++;
    } else {
        BVH_CUDA_TRY(cudaMemsetAsync(P.qseq, 0, sizeof(uint32_t) * qcap, st));
        init_root_kernel<T><<<1, 32, 0, st>>>(P);
        ctx->launches++;
        if (ctx->profile) cudaEventRecord(ctx->ev_build[0], st);
        if (P.gang_budget) {
            // gangs spin on each other: the grid must be co-resident as a whole, which is what a cooperative launch
            // guarantees (two concurrent builds are then serialised by the scheduler instead of starving each other)
            void* kargs[] = {&P};
            BVH_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)build_kernel<T>, dim3(grid), dim3(WARPS_PER_CTA * 32), kargs, 0, st));
        } else {
            build_kernel<T><<<grid, WARPS_PER_CTA * 32, 0, st>>>(P);
        }
        ctx->launches++;
        if (P.small_max) {
            const unsigned sgrid = (unsigned)(((size_t)n / 2 + 1 + 127) / 128);
            small_subtrees_kernel<T><<<sgrid, 128, 0, st>>>(P);
            ctx->launches++;
        }
        if (ctx->profile) { cudaEventRecord(ctx->ev_build[1], st); ctx->have_build = true; }
    }
    finish_status_kernel<T><<<1, 32, 0, st>>>(P.ctl, P.status, n);
    ctx->launches++;
    BVH_CUDA_TRY(cudaGetLastError());
    if (P.trace) {                          // debugging aid: dump the task log (synchronises)
        std::vector<uint4> h(P.trace_cap);
        BVH_CUDA_TRY(cudaMemcpyAsync(h.data(), P.trace, sizeof(uint4) * P.trace_cap, cudaMemcpyDeviceToHost, st));
        BVH_CUDA_TRY(cudaStreamSynchronize(st));
        if (FILE* f = fopen(trace_path, "wb")) { fwrite(h.data(), sizeof(uint4), h.size(), f); fclose(f); }
        dfree(ctx, P.trace);
    }
    dfree(ctx, idx0); dfree(ctx, idx1); dfree(ctx, P.bkt); dfree(ctx, P.q); dfree(ctx, P.qseq);
    dfree(ctx, P.big); dfree(ctx, P.tilecnt); dfree(ctx, P.ctl); dfree(ctx, P.rootkeys); dfree(ctx, P.small);
    tree->status_pending = true;
    return BVHGPU_OK;
}

// ---- treelet session: lbvh.cu pre-fills the task queue with SEG tasks (one per treelet of <= TILE shapes) and this runs
// the same persistent kernel over them: binned-SAH re-optimisation of the bottom of an LBVH tree, staged in shared memory.
template <class T>
int treelet_begin(bvhgpu_ctx* ctx, Tree<T>* tree, uint32_t* sorted_ids, TreeletSession<T>* S) {
    cudaStream_t st = ctx->stream;
    const uint32_t n = tree->n;
    BuildParams<T>* P = new BuildParams<T>();
    memset(P, 0, sizeof(*P));
    P->aabb = tree->d_aabb; P->nodes = tree->d_nodes; P->node_index = tree->d_node_index; P->node_start = tree->d_node_start;
    P->n = n; P->status = tree->d_status; P->timeout_ns = 20ull * 1000ull * 1000ull * 1000ull;
    const uint32_t qcap = next_pow2(std::max<uint64_t>(n, 1024) * 2);
    P->qmask = qcap - 1;
    P->idx[0] = sorted_ids;
    BVH_TRY(dalloc_t(ctx, &P->idx[1], n));
    BVH_TRY(dalloc_t(ctx, &P->bkt, n));
    BVH_TRY(dalloc_t(ctx, &P->q, qcap));
    BVH_TRY(dalloc_t(ctx, &P->qseq, qcap));
    P->sdiv = TILE;
    BVH_TRY(dalloc_t(ctx, &P->big, 2));
    BVH_TRY(dalloc_t(ctx, &P->tilecnt, 32));
    BVH_TRY(dalloc_t(ctx, &P->ctl, 1));
    const bool defer_small = ctx->build_small < 0 ? (sizeof(T) == 8 && n >= 400000u) : ctx->build_small != 0;
    P->small_max = defer_small ? SMALL : 0u;
    P->opt_subtree = ctx->build_subtree < 0 ? !defer_small : ctx->build_subtree != 0;
    BVH_TRY(dalloc_t(ctx, &P->small, P->small_max ? (size_t)n / 2 + 1 : 1));
    BVH_CUDA_TRY(cudaMemsetAsync(P->qseq, 0, sizeof(uint32_t) * qcap, st));
    BVH_CUDA_TRY(cudaMemsetAsync(P->ctl, 0, sizeof(BuildCtl), st));
    S->params = P; S->q = P->q; S->qseq = P->qseq; S->qmask = P->qmask; S->ctl = P->ctl;
    return BVHGPU_OK;
}
template <class T> __global__ void treelet_start_kernel(BuildCtl* ctl, const BuildStatus* status) {
    if (threadIdx.x == 0) {
        ctl->t_start = global_timer_ns();
        if (status->nan_found) ctl->error = (uint32_t)BVHGPU_ERR_NAN;     // NaN shapes: build nothing (the reference panics), the workers leave at once
    }
}

template <class T>
int treelet_finish(bvhgpu_ctx* ctx, Tree<T>* tree, TreeletSession<T>* S) {
    cudaStream_t st = ctx->stream;
    BuildParams<T>* P = static_cast<BuildParams<T>*>(S->params);
    const uint32_t n = tree->n;
    treelet_start_kernel<T><<<1, 32, 0, st>>>(P->ctl, P->status);
    int occ = 1;
    BVH_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, build_kernel<T>, WARPS_PER_CTA * 32, 0));
    if (occ < 1) occ = 1;
    uint64_t want = ((uint64_t)n / 16 + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    if (want < 1) want = 1;
    const int grid = (int)std::min<uint64_t>(want, (uint64_t)ctx->sm_count * occ);
    build_kernel<T><<<grid, WARPS_PER_CTA * 32, 0, st>>>(*P);
    ctx->launches += 2;
    if (P->small_max) {
        const unsigned sgrid = (unsigned)(((size_t)n / 2 + 1 + 127) / 128);
        small_subtrees_kernel<T><<<sgrid, 128, 0, st>>>(*P);
        ctx->launches++;
    }
    finish_status_kernel<T><<<1, 32, 0, st>>>(P->ctl, P->status, n);
    ctx->launches++;
    BVH_CUDA_TRY(cudaGetLastError());
    dfree(ctx, P->idx[1]); dfree(ctx, P->bkt); dfree(ctx, P->q); dfree(ctx, P->qseq); dfree(ctx, P->big); dfree(ctx, P->tilecnt);
    dfree(ctx, P->ctl); dfree(ctx, P->small);
    delete P;
    S->params = nullptr;
    return BVHGPU_OK;
}

// ---- rebuild session: the exact builder restarted from inner nodes (bvhgpu_optimize, flatten.cu: optimize) ---------------
// One task per rebuild root: its shape range in leaf order (index buffer 0), its node index, its parent, its AABB (the
// join of its two stored child AABBs, tight after the refit) and the bounds of its shape centres (carried up by the
// refit).  From there on it is the ordinary build: the subtree of a node with k shapes occupies the node range
// [i, i + 2k - 1) and the leaf range [start, start + k), and child_l = i + 1 / child_r = i + 2 n_l stay inside it.
template <class T>
__global__ void __launch_bounds__(256) rebuild_push_kernel(BuildParams<T> P, const uint32_t* __restrict__ roots, const uint32_t* __restrict__ n_roots,
                                                           const T* __restrict__ cb, bool cb_by_root) {
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t nr = *n_roots;
    if (P.status->nan_found) return;                                   // (cannot happen through the C ABI: new AABBs are checked before the tree is touched)
    for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < nr; i += warps) {
        const uint32_t r = roots[i];
        const typename Traits<T>::Node& nd = P.nodes[r];
        BTask<T> t;
        t.start = P.node_start[r]; t.count = nd.shape; t.node = r; t.parent_buf = nd.parent;          // range lives in index buffer 0
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            t.ab[k] = min_t(nd.l_aabb.min[k], nd.r_aabb.min[k]); t.ab[3 + k] = max_t(nd.l_aabb.max[k], nd.r_aabb.max[k]);
            const size_t at = 6 * (size_t)(cb_by_root ? i : r);
            t.cb[k] = cb[at + k]; t.cb[3 + k] = cb[at + 3 + k];
        }
        if (lane_id() == 0) { atomicSub(&P.ctl->leaves_done, t.count); atomicAdd(&P.ctl->rebuilt, t.count); }
        __syncwarp();
        if (t.count > (uint32_t)TILE) { if (!try_create_gang(P, t)) create_big(P, t); } else push_seg(P, t);
    }
}
template <class T> __global__ void rebuild_start_kernel(BuildCtl* ctl, uint32_t n, const BuildStatus* status) {
    if (threadIdx.x == 0) {
        ctl->t_start = global_timer_ns(); ctl->leaves_done = n;
        if (status->nan_found) ctl->error = (uint32_t)BVHGPU_ERR_NAN;
    }
}

template <class T>
int rebuild_subtrees(bvhgpu_ctx* ctx, Tree<T>* tree, const uint32_t* d_roots, const uint32_t* d_n_roots, const T* cb, uint32_t* idx0, bool cb_by_root) {
    cudaStream_t st = ctx->stream;
    const uint32_t n = tree->n;
    BuildParams<T> P{};
    P.aabb = tree->d_aabb; P.nodes = tree->d_nodes; P.node_index = tree->d_node_index; P.node_start = tree->d_node_start;
    P.n = n; P.status = tree->d_status; P.timeout_ns = 20ull * 1000ull * 1000ull * 1000ull;
    const uint32_t qcap = next_pow2(std::max<uint64_t>(n, 1024) * 2);
    P.qmask = qcap - 1;
    int occ = 1;
    BVH_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, build_kernel<T>, WARPS_PER_CTA * 32, 0));
    if (occ < 1) occ = 1;
    uint64_t want = ((uint64_t)n / 16 + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    if (want < 1) want = 1;
    const int grid = (int)std::min<uint64_t>(want, (uint64_t)ctx->sm_count * occ);
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
    P.gang_budget = (ctx->build_gang != 0 && coop) ? (uint32_t)grid * WARPS_PER_CTA / 2u : 0u;
    if (ctx->build_gang < 0 && (uint64_t)n / GT > P.gang_budget) P.gang_budget = 0;
    P.sdiv = P.gang_budget ? GT : TILE;
    const size_t nbig = 2 * ((size_t)n / P.sdiv + 2);
    // a rebuild usually covers a small part of the tree: the thread-per-range kernel (one thread replays a <= 16-shape range) only pays
    // when millions of ranges keep every SM busy; here the warp-per-subtree builder finishes the job sooner (10 M f64, 2.4 M shapes
    // rebuilt: 4.2 ms in small_subtrees_kernel)
    const bool defer_small = ctx->build_small < 0 ? false : ctx->build_small != 0;
    P.small_max = defer_small ? SMALL : 0u;
    P.opt_subtree = ctx->build_subtree < 0 ? !defer_small : ctx->build_subtree != 0;
    P.idx[0] = idx0;
    BVH_TRY(dalloc_t(ctx, &P.idx[1], n));
    BVH_TRY(dalloc_t(ctx, &P.bkt, n));
    BVH_TRY(dalloc_t(ctx, &P.q, qcap));
    BVH_TRY(dalloc_t(ctx, &P.qseq, qcap));
    BVH_TRY(dalloc_t(ctx, &P.big, nbig));
    BVH_TRY(dalloc_t(ctx, &P.tilecnt, nbig * 8));
    BVH_TRY(dalloc_t(ctx, &P.ctl, 1));
    BVH_TRY(dalloc_t(ctx, &P.small, P.small_max ? (size_t)n / 2 + 1 : 1));
    BVH_CUDA_TRY(cudaMemsetAsync(P.qseq, 0, sizeof(uint32_t) * qcap, st));
    BVH_CUDA_TRY(cudaMemsetAsync(P.ctl, 0, sizeof(BuildCtl), st));
    rebuild_start_kernel<T><<<1, 32, 0, st>>>(P.ctl, n, P.status);
    const int pblocks = (int)std::min<uint64_t>(((uint64_t)n + 63) / 64, (uint64_t)ctx->sm_count * 4);
    rebuild_push_kernel<T><<<pblocks, 256, 0, st>>>(P, d_roots, d_n_roots, cb, cb_by_root);
    if (P.gang_budget) {
        void* kargs[] = {&P};
        BVH_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)build_kernel<T>, dim3(grid), dim3(WARPS_PER_CTA * 32), kargs, 0, st));
    } else {
        build_kernel<T><<<grid, WARPS_PER_CTA * 32, 0, st>>>(P);
    }
    ctx->launches += 3;
    if (P.small_max) {
        const unsigned sgrid = (unsigned)(((size_t)n / 2 + 1 + 127) / 128);
        small_subtrees_kernel<T><<<sgrid, 128, 0, st>>>(P);
        ctx->launches++;
    }
    finish_status_kernel<T><<<1, 32, 0, st>>>(P.ctl, P.status, n);
    ctx->launches++;
    BVH_CUDA_TRY(cudaGetLastError());
    dfree(ctx, P.idx[1]); dfree(ctx, P.bkt); dfree(ctx, P.q); dfree(ctx, P.qseq); dfree(ctx, P.big); dfree(ctx, P.tilecnt);
    dfree(ctx, P.ctl); dfree(ctx, P.small);
    tree->status_pending = true;
    return BVHGPU_OK;
}
template int rebuild_subtrees<float>(bvhgpu_ctx*, Tree<float>*, const uint32_t*, const uint32_t*, const float*, uint32_t*, bool);
template int rebuild_subtrees<double>(bvhgpu_ctx*, Tree<double>*, const uint32_t*, const uint32_t*, const double*, uint32_t*, bool);

template int treelet_begin<float>(bvhgpu_ctx*, Tree<float>*, uint32_t*, TreeletSession<float>*);
template int treelet_begin<double>(bvhgpu_ctx*, Tree<double>*, uint32_t*, TreeletSession<double>*);
template int treelet_finish<float>(bvhgpu_ctx*, Tree<float>*, TreeletSession<float>*);
template int treelet_finish<double>(bvhgpu_ctx*, Tree<double>*, TreeletSession<double>*);

template int build_exact_sah<float>(bvhgpu_ctx*, const bvh_aabb3f*, uint32_t, Tree<float>*);
template int build_exact_sah<double>(bvhgpu_ctx*, const bvh_aabb3d*, uint32_t, Tree<double>*);
template int convert_aabbs<float>(bvhgpu_ctx*, const bvh_aabb3f*, uint32_t, DAabbF*, uint32_t*);
template int convert_aabbs<double>(bvhgpu_ctx*, const bvh_aabb3d*, uint32_t, DAabbD*, uint32_t*);

}  // namespace bvhb200
