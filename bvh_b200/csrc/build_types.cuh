// bvh_b200/csrc/build_types.cuh -- task / queue records of the persistent SAH build kernel, shared with lbvh.cu
// (the LBVH + treelet mode pre-fills the queue with one SEG task per treelet).
#pragma once
#include "common.cuh"

namespace bvhb200 {

constexpr int TILE = 512;              // shapes per tile task of a multi-warp segment (measured: 128 / 256 / 512 / 1024 -> 0.67 / 0.60 / 0.59 / 0.64 ms at 120 k)
constexpr int WARPS_PER_CTA = 8;
#ifndef BUILD_MIN_CTAS
#define BUILD_MIN_CTAS 2
#endif
constexpr int LOCAL_STACK = 24;        // per-warp DFS stack (entries)
constexpr uint32_t LOCAL_MAX = 6;      // right children up to this size stay on the warp's own stack; larger ones go to
                                       // the global queue: idle warps are plentiful, the critical path is what matters
constexpr uint32_t KIND_SEG = 0, KIND_BIN = 1, KIND_SCATTER = 2, KIND_GANG = 3;
constexpr int GT = 128;                // shapes per warp of a GANG (co-resident warps that walk the top of the tree level by
                                       // level with device-wide barriers instead of queue hops); the last warp takes the remainder
constexpr uint32_t GANG_MIN = 512;     // ranges below this leave the gang and become SEG tasks

template <class T> struct __align__(16) BTask {
    uint32_t start, count, node, parent_buf;   // parent_buf: bit31 = which index buffer holds the range
    T ab[6];                                   // aabb_bounds      (min xyz, max xyz)
    T cb[6];                                   // centroid_bounds
};
template <class T> struct __align__(16) QSlot {
    BTask<T> t;
    uint32_t kind, a, b, pad;                  // tile tasks: a = big-segment id, b = tile index
};
template <class T> struct __align__(16) BigSeg {
    using Key = typename Traits<T>::Key;
    BTask<T> t;
    Key keys[72];                              // [bucket][12]: aabb min3, aabb max3, centroid min3, centroid max3
    uint32_t cnt[6]; uint32_t epoch; uint32_t bin_done;
    uint32_t scat_done; uint32_t nl; uint32_t pad0[2];
    uint32_t base[6]; uint32_t pad1[2];
    T child[24];                               // lab, lcb, rab, rcb of the chosen split
};
// In-warp subtree builder (process_subtree): a range of <= 32 shapes lives in registers, one shape per lane.
constexpr uint32_t SUBW = 32;
template <class T> struct __align__(16) Staged { uint32_t id; int b; T mn[3], mx[3]; };
template <class T> struct __align__(16) WarpScratch {
    using Key = typename Traits<T>::Key;
    Key keys[72];
    uint32_t cnt[8];
    T child[24];
    BTask<T> stack[LOCAL_STACK];
    Staged<T> stage[32];                        // lane permutation of process_subtree
};
// Control block.  Each group lives in its own 128-byte line: idle warps poll `leaves_done` / `error` all the time, and that
// traffic must not queue in front of the ticket atomics on `head` / `tail`.
struct __align__(128) BuildCtl {
    uint32_t head;        uint32_t pad0[31];
    uint32_t tail;        uint32_t pad1[31];
    uint32_t leaves_done, error; uint32_t pad2[30];
    unsigned long long t_start;
    uint32_t small_count;            // ranges of <= SMALL shapes deferred to small_subtrees_kernel
    uint32_t gang_used;              // warps currently reserved by gangs (bounded by BuildParams::gang_budget)
    uint32_t gang_trace;             // BVHGPU_TRACE: gang levels are logged from the end of the trace buffer
    uint32_t rebuilt;                // rebuild session: shapes in the rebuilt subtrees
    uint32_t pad3[26];
};
constexpr uint32_t SMALL = 16;       // ranges this small are finished by ONE THREAD each in a second kernel


}  // namespace bvhb200
