/* ============================================================================
 * include/bvh_b200.h -- C ABI of libbvh_b200.so
 *
 * B200-native (sm_100a) replacement for the ONE data-parallel hot path of the
 * Rust crate svenstaro/bvh 0.12.0:
 *
 *     Bvh::build  ->  Bvh::flatten  ->  batched Ray traversal
 *
 * This header is the drop-in boundary: plain pointers and sizes only, so a Rust
 * shim (`impl BoundingHierarchy<T,3> for GpuBvh<T>`) can bind
 * it with `extern "C"`.  Every entry point cites the reference interface it
 * replaces (file:line relative to the reference checkout).
 *
 * There is NO CPU fallback behind any of these calls: without a CUDA device
 * `bvhgpu_create` fails with BVHGPU_ERR_CUDA.
 *
 * Conventions
 *   - all functions return a bvhgpu_status (0 = ok); `bvhgpu_last_error()` gives
 *     the message for the calling thread.  The reference panics on the same
 *     conditions (NaN centroid: src/bvh/bvh_node.rs:214-217); the shim turns a
 *     non-zero status into `panic!`.
 *   - `*_dev_*` variants take DEVICE pointers, enqueue on the context's stream
 *     and do not synchronise unless they must return a host value.
 *   - n == 0 and n == 1 behave as the reference does (empty tree / root leaf,
 *     src/bvh/bvh_impl.rs:57-59, src/bvh/bvh_node.rs:95-104, 314).
 * ========================================================================== */
#ifndef BVH_B200_H
#define BVH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVHGPU_INVALID_INDEX 0xFFFFFFFFu /* u32::MAX sentinel, src/flat_bvh.rs:36-45 */

/* ---- POD mirrors (Rust's Aabb / Ray / BvhNode / FlatNode are not repr(C)) ---- */

/* Aabb<f32,3> / Aabb<f64,3>: src/aabb/aabb_impl.rs:10-16 */
typedef struct { float min[3]; float max[3]; } bvh_aabb3f;    /* 24 B */
typedef struct { double min[3]; double max[3]; } bvh_aabb3d;  /* 48 B */

/* Ray<T,3>: src/ray/ray_impl.rs:17-29 (direction already normalised, inv = 1/direction) */
typedef struct { float origin[3]; float direction[3]; float inv_direction[3]; } bvh_ray3f;     /* 36 B */
typedef struct { double origin[3]; double direction[3]; double inv_direction[3]; } bvh_ray3d;  /* 72 B */

/* BvhNode<T,3> (enum, src/bvh/bvh_node.rs:21-47) flattened:
 *   Leaf: child_l == child_r == BVHGPU_INVALID_INDEX, shape = shape_index, AABBs = Aabb::empty()
 *   Node: child_l / child_r = child node indices, shape = number of shapes under the node
 *         (extra information the reference does not store), l_aabb / r_aabb = child AABBs.
 * Indices are u32 (the reference uses usize; FlatNode already limits trees to u32). */
typedef struct { uint32_t parent, child_l, child_r, shape; bvh_aabb3f l_aabb, r_aabb; } bvh_node3f;  /*  64 B */
typedef struct { uint32_t parent, child_l, child_r, shape; bvh_aabb3d l_aabb, r_aabb; } bvh_node3d;  /* 112 B */

/* FlatNode<T,3>: src/flat_bvh.rs:17-46 */
typedef struct { bvh_aabb3f aabb; uint32_t entry_index, exit_index, shape_index; } bvh_flat3f;             /* 36 B */
typedef struct { bvh_aabb3d aabb; uint32_t entry_index, exit_index, shape_index, _pad; } bvh_flat3d;      /* 64 B */

/* D = 2 twins (the reference is generic in D; 2-D slab tests: src/ray/intersect_simd.rs:99-133, 181-191) */
typedef struct { float min[2]; float max[2]; } bvh_aabb2f;                                                   /* 16 B */
typedef struct { double min[2]; double max[2]; } bvh_aabb2d;                                                 /* 32 B */
typedef struct { float origin[2]; float direction[2]; float inv_direction[2]; } bvh_ray2f;                   /* 24 B */
typedef struct { double origin[2]; double direction[2]; double inv_direction[2]; } bvh_ray2d;                /* 48 B */
typedef struct { uint32_t parent, child_l, child_r, shape; bvh_aabb2f l_aabb, r_aabb; } bvh_node2f;          /* 48 B */
typedef struct { uint32_t parent, child_l, child_r, shape; bvh_aabb2d l_aabb, r_aabb; } bvh_node2d;          /* 80 B */
typedef struct { bvh_aabb2f aabb; uint32_t entry_index, exit_index, shape_index; } bvh_flat2f;               /* 28 B */
typedef struct { bvh_aabb2d aabb; uint32_t entry_index, exit_index, shape_index, _pad; } bvh_flat2d;         /* 48 B */

typedef enum {
    BVHGPU_OK = 0,
    BVHGPU_ERR_INVALID = 1,     /* bad argument */
    BVHGPU_ERR_CUDA = 2,        /* CUDA runtime error / no device */
    BVHGPU_ERR_NAN = 3,         /* NaN in an input AABB (reference: panic, bvh_node.rs:214-217) */
    BVHGPU_ERR_CAPACITY = 4,    /* caller buffer too small; *total / *len holds the needed size */
    BVHGPU_ERR_TIMEOUT = 5,     /* device watchdog fired */
    BVHGPU_ERR_UNSUPPORTED = 6,
    BVHGPU_ERR_INTERNAL = 7
} bvhgpu_status;

typedef enum {
    BVHGPU_BUILD_EXACT_SAH = 0, /* bit-identical to Bvh::build (6-bucket SAH, src/bvh/bvh_node.rs:81-279) */
    BVHGPU_BUILD_LBVH = 1,      /* Morton/Karras LBVH: same hit sets, different topology */
    BVHGPU_BUILD_LBVH_TREELET = 2 /* LBVH top + every subtree of <= 512 shapes rebuilt with the reference's 6-bucket SAH (shared memory) */
} bvhgpu_build_mode;

typedef enum {
    BVHGPU_TRAVERSE_BVH = 0,    /* Bvh::traverse semantics     (src/bvh/bvh_node.rs:288-319): leaves are not re-tested */
    BVHGPU_TRAVERSE_FLAT = 1    /* FlatBvh::traverse semantics (src/flat_bvh.rs:396-431): reached leaves re-test the shape AABB */
} bvhgpu_traverse_mode;

/* Ray batch layouts at the boundary.  FULL is the crate's Ray (bvh_ray3f / bvh_ray3d).  OD carries only what Ray::new keeps
 * besides the reciprocal: 6 scalars per ray {origin[3], direction[3]} with `direction` exactly as Ray stores it (normalised);
 * the device recomputes inv_direction = 1/direction with the same IEEE division Ray::new performs (src/ray/ray_impl.rs:76-78),
 * so both layouts give bit-identical results -- OD moves 24 instead of 36 bytes per f32 ray across PCIe. */
typedef enum { BVHGPU_RAYS_FULL = 0, BVHGPU_RAYS_OD = 1 } bvhgpu_ray_layout;

typedef struct bvhgpu_ctx bvhgpu_ctx;       /* one per device: stream, scratch pool           */
typedef struct bvhgpu_tree3f bvhgpu_tree3f; /* device-resident Bvh<f32,3> (+ FlatBvh, shape AABBs) */
typedef struct bvhgpu_tree3d bvhgpu_tree3d; /* device-resident Bvh<f64,3>                      */
typedef struct bvhgpu_tree2f bvhgpu_tree2f; /* device-resident Bvh<f32,2>                      */
typedef struct bvhgpu_tree2d bvhgpu_tree2d; /* device-resident Bvh<f64,2>                      */

/* ---- context ------------------------------------------------------------------ */
int bvhgpu_create(int device, bvhgpu_ctx** out);
void bvhgpu_destroy(bvhgpu_ctx* ctx);
const char* bvhgpu_last_error(void);
const char* bvhgpu_version(void);
/* Enqueue on an externally owned cudaStream_t (e.g. torch's current stream; 0 is CUDA's legacy default
 * stream and is honoured as such).  bvhgpu_reset_stream returns to the context's own stream. */
int bvhgpu_set_stream(bvhgpu_ctx* ctx, void* cuda_stream);
int bvhgpu_reset_stream(bvhgpu_ctx* ctx);
/* Waits for the context's stream.  Also the point where errors of asynchronous multi-GPU steps surface: a peer that never
 * answered (BVHGPU_ERR_TIMEOUT) is reported here, once, and the step's result must not be used. */
int bvhgpu_synchronize(bvhgpu_ctx* ctx);
/* Pinned host memory for ray / result staging, placed on the NUMA node the device hangs off (sysfs numa_node of the
 * PCI function) -- on a two-socket host a buffer pinned on the far socket moves at a fraction of the PCIe rate. */
int bvhgpu_host_alloc(bvhgpu_ctx* ctx, size_t bytes, void** out);
int bvhgpu_host_free(bvhgpu_ctx* ctx, void* p);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t bvhgpu_launch_count(const bvhgpu_ctx* ctx);
/* Tunables: "traverse_slots" (per-ray hit slots of the single-pass path, 0 = two-pass count/fill, -1 = auto),
 * "traverse_persistent" (0 = one ray per thread, 1 = persistent refill kernel, 2 = decided per batch by a coherence probe),
 * "build_small" (exact builder: finish ranges of <= 16 shapes with one thread each in a second kernel; -1 auto by type and size, 0 never, 1 always),
 * "build_subtree" (exact builder: build ranges of <= 32 shapes in registers, one warp per subtree; -1 auto, 0 never, 1 always),
 * "build_gang" (exact builder: co-resident warp gangs walk the top levels behind device-wide barriers; -1 auto by size, 0 never, 1 always),
 * -- every combination produces the same bits; the switches exist for measurement (tools/build_sweep.py) --
 * "traverse_stream" (host-pointer traversal: consume the rays in a running kernel while the copy is still in flight;
 *   -1 auto = only when launches are asynchronous and no profiler / debugger / sanitizer is attached, 0 never, 1 force),
 * "traverse_top" (f32 persistent walk with the top of the tree in shared memory: -1 auto = trees with >= 5 MB of traversal records,
 *   0 never, 1 always, > 1 = always with at most that many top entries; the results are the same bits either way),
 * "walk_grid" (CTAs of the plain persistent walk; 0 = automatic: 4..8 per SM by batch size),
 * "profile" (1: bracket the dominant kernels with CUDA events, read back with bvhgpu_get_metric). */
int bvhgpu_set_option(bvhgpu_ctx* ctx, const char* name, int64_t value);
/* Measurements of the last profiled call on this context: "walk_ms" (traversal walk kernel),
 * "build_ms" (persistent SAH build kernel).  Synchronises on the recorded events. */
int bvhgpu_get_metric(bvhgpu_ctx* ctx, const char* name, double* out);

/* ---- build: replaces BoundingHierarchy::build / build_par for Bvh ----------------
 * (src/bounding_hierarchy.rs:158-177, src/bvh/bvh_impl.rs:40-96, src/bvh/bvh_node.rs:81-279).
 * `aabbs[i]` is `shapes[i].aabb()` (Bounded::aabb, src/aabb/aabb_impl.rs:55) gathered by the shim.
 * The tree stays on the device; fetch Bvh.nodes / the per-shape node indices with
 * bvhgpu_tree_nodes_*.  build_par == build (rayon is off the hot path). */
int bvhgpu_build_f32x3(bvhgpu_ctx* ctx, const bvh_aabb3f* aabbs, size_t n, int mode, bvhgpu_tree3f** out);
int bvhgpu_build_f64x3(bvhgpu_ctx* ctx, const bvh_aabb3d* aabbs, size_t n, int mode, bvhgpu_tree3d** out);
int bvhgpu_build_dev_f32x3(bvhgpu_ctx* ctx, const void* dev_aabbs, size_t n, int mode, bvhgpu_tree3f** out);
int bvhgpu_build_dev_f64x3(bvhgpu_ctx* ctx, const void* dev_aabbs, size_t n, int mode, bvhgpu_tree3d** out);

/* Upload an existing reference-layout Bvh (`Bvh.nodes` in the preorder layout Bvh::build emits:
 * child_l == i+1, src/bvh/bvh_node.rs:138-142) together with the shapes' AABBs. */
int bvhgpu_tree_from_nodes_f32x3(bvhgpu_ctx* ctx, const bvh_node3f* nodes, size_t n_nodes,
                                 const bvh_aabb3f* aabbs, size_t n, bvhgpu_tree3f** out);
int bvhgpu_tree_from_nodes_f64x3(bvhgpu_ctx* ctx, const bvh_node3d* nodes, size_t n_nodes,
                                 const bvh_aabb3d* aabbs, size_t n, bvhgpu_tree3d** out);

void bvhgpu_tree_free_f32x3(bvhgpu_tree3f* tree);
void bvhgpu_tree_free_f64x3(bvhgpu_tree3d* tree);
size_t bvhgpu_tree_num_shapes_f32x3(const bvhgpu_tree3f* tree);
size_t bvhgpu_tree_num_shapes_f64x3(const bvhgpu_tree3d* tree);
size_t bvhgpu_tree_num_nodes_f32x3(const bvhgpu_tree3f* tree);   /* 2n-1 (0 for n == 0) */
size_t bvhgpu_tree_num_nodes_f64x3(const bvhgpu_tree3d* tree);

/* Materialise `Bvh.nodes` (pub field, src/bvh/bvh_impl.rs:27-33) and the leaf node index of every
 * shape (BHShape::set_bh_node_index, src/bounding_hierarchy.rs:58; written at src/bvh/bvh_node.rs:103).
 * Either pointer may be NULL. */
int bvhgpu_tree_nodes_f32x3(bvhgpu_tree3f* tree, bvh_node3f* out_nodes, uint32_t* out_node_index);
int bvhgpu_tree_nodes_f64x3(bvhgpu_tree3d* tree, bvh_node3d* out_nodes, uint32_t* out_node_index);

/* ---- D = 2: Bvh<T,2>::build / nodes / flatten / traverse.  Host pointers; semantics, modes, error codes and
 * the CSR output exactly as the 3-D entry points above.  The scene is embedded in the plane z = 0 of the 3-D kernels in a way that
 * reproduces the 2-D arithmetic bit for bit (dim2.cu). */
int bvhgpu_build_f32x2(bvhgpu_ctx* ctx, const bvh_aabb2f* aabbs, size_t n, int mode, bvhgpu_tree2f** out);
int bvhgpu_build_f64x2(bvhgpu_ctx* ctx, const bvh_aabb2d* aabbs, size_t n, int mode, bvhgpu_tree2d** out);
void bvhgpu_tree_free_f32x2(bvhgpu_tree2f* tree);
void bvhgpu_tree_free_f64x2(bvhgpu_tree2d* tree);
size_t bvhgpu_tree_num_shapes_f32x2(const bvhgpu_tree2f* tree);
size_t bvhgpu_tree_num_shapes_f64x2(const bvhgpu_tree2d* tree);
int bvhgpu_tree_nodes_f32x2(bvhgpu_tree2f* tree, bvh_node2f* out_nodes, uint32_t* out_node_index);
int bvhgpu_tree_nodes_f64x2(bvhgpu_tree2d* tree, bvh_node2d* out_nodes, uint32_t* out_node_index);
int bvhgpu_flatten_f32x2(bvhgpu_tree2f* tree, bvh_flat2f* out, size_t cap, size_t* len);
int bvhgpu_flatten_f64x2(bvhgpu_tree2d* tree, bvh_flat2d* out, size_t cap, size_t* len);
int bvhgpu_traverse_f32x2(bvhgpu_tree2f* tree, int mode, const bvh_ray2f* rays, size_t nrays,
                          uint32_t* offsets, uint32_t* hits, size_t cap, size_t* total);
int bvhgpu_traverse_f64x2(bvhgpu_tree2d* tree, int mode, const bvh_ray2d* rays, size_t nrays,
                          uint32_t* offsets, uint32_t* hits, size_t cap, size_t* total);

/* ---- flatten: replaces Bvh::flatten (src/flat_bvh.rs:60-143, 240-251, 312-319) -----
 * Writes the FlatBvh (3n-2 FlatNodes for n >= 2, 1 for n == 1, 0 for n == 0) into `out`
 * (may be NULL to only build the device copy) and its length into *len. */
int bvhgpu_flatten_f32x3(bvhgpu_tree3f* tree, bvh_flat3f* out, size_t cap, size_t* len);
int bvhgpu_flatten_f64x3(bvhgpu_tree3d* tree, bvh_flat3d* out, size_t cap, size_t* len);

/* ---- traverse: batched Bvh::traverse / FlatBvh::traverse for Ray queries -----------
 * (src/bvh/bvh_impl.rs:104-119, src/bvh/bvh_node.rs:288-319, src/flat_bvh.rs:396-431,
 *  slab test src/ray/intersect_default.rs:16-37).
 * Output is CSR: hits of ray r are hits[offsets[r] .. offsets[r+1]) = shape indices in the
 * reference's DFS (left-first) order.  If the hit list does not fit `cap`, offsets and *total
 * are still valid, the call returns BVHGPU_ERR_CAPACITY and bvhgpu_traverse_fetch_* can copy the
 * retained result without traversing again.  Shape AABBs are the ones given at build time. */
int bvhgpu_traverse_f32x3(bvhgpu_tree3f* tree, int mode, const bvh_ray3f* rays, size_t nrays,
                          uint32_t* offsets, uint32_t* hits, size_t cap, size_t* total);
int bvhgpu_traverse_f64x3(bvhgpu_tree3d* tree, int mode, const bvh_ray3d* rays, size_t nrays,
                          uint32_t* offsets, uint32_t* hits, size_t cap, size_t* total);
int bvhgpu_traverse_fetch_f32x3(bvhgpu_tree3f* tree, uint32_t* hits, size_t cap);
int bvhgpu_traverse_fetch_f64x3(bvhgpu_tree3d* tree, uint32_t* hits, size_t cap);
/* The same with the compact ray layout BVHGPU_RAYS_OD: `origin_dir` holds 6 scalars per ray. */
int bvhgpu_traverse_od_f32x3(bvhgpu_tree3f* tree, int mode, const float* origin_dir, size_t nrays,
                             uint32_t* offsets, uint32_t* hits, size_t cap, size_t* total);
int bvhgpu_traverse_od_f64x3(bvhgpu_tree3d* tree, int mode, const double* origin_dir, size_t nrays,
                             uint32_t* offsets, uint32_t* hits, size_t cap, size_t* total);
/* Device-resident variant: rays / offsets / hits are device pointers.  `total` may be NULL
 * (no host synchronisation); hits beyond `cap` are dropped and reported through *total. */
int bvhgpu_traverse_dev_f32x3(bvhgpu_tree3f* tree, int mode, const void* dev_rays, size_t nrays,
                              void* dev_offsets, void* dev_hits, size_t cap, size_t* total);
int bvhgpu_traverse_dev_f64x3(bvhgpu_tree3d* tree, int mode, const void* dev_rays, size_t nrays,
                              void* dev_offsets, void* dev_hits, size_t cap, size_t* total);
int bvhgpu_traverse_od_dev_f32x3(bvhgpu_tree3f* tree, int mode, const void* dev_origin_dir, size_t nrays,
                                 void* dev_offsets, void* dev_hits, size_t cap, size_t* total);
int bvhgpu_traverse_od_dev_f64x3(bvhgpu_tree3d* tree, int mode, const void* dev_origin_dir, size_t nrays,
                                 void* dev_offsets, void* dev_hits, size_t cap, size_t* total);
/* ---- multi-GPU ray sharding with the exchange fused into the traversal (no NCCL on the data path) -------------
 * Every rank owns a contiguous shard of the ray batch and a replica of the tree.  Every rank ends the step with its own
 * copy of the GLOBAL CSR in original ray order -- the all-gather of hit lists north_star asks for -- built over peer
 * memory (buffers allocated with bvhgpu_peer_alloc and opened on the other ranks through CUDA IPC):
 *   1. the scan kernel behind the local walk stores every 2048-ray tile's hit COUNTS, narrowed to 1 / 2 / 4 bytes by the
 *      tile's largest count, into every rank's staging (8/16-byte P2P stores over NVLink); its last block adds the table of
 *      tile offsets and publishes the rank's hit total in all mailboxes;
 *   2. the emit kernel waits for the peers' posts (which fixes its hit base), writes its hit lists into its own copy of
 *      the global hit buffer, and every block ships its contiguous piece to all peers with whole 16-byte P2P stores
 *      (4-byte stores scattered straight from the emit loop reached a small fraction of the NVLink rate: 12.9 ms per
 *      16 M-ray Sponza step on 4 GPUs against 2.7 ms this way on 8); the last block raises the done flags;
 *   3. extra blocks of the same launch rebuild the global u32 offsets on every rank from the staged counts (1 byte per ray
 *      crossed NVLink instead of 4) and end the step by waiting for the peers' done flags: when the stream reaches the end of
 *      the step, this rank's copy of the global CSR is complete.  Same number of launches as a single-GPU step.
 * `seq` must increase by one per call on all ranks.  No host synchronisation; failures (a peer that never answers)
 * are reported by bvhgpu_synchronize.  Mailbox layout (trace words for diagnostics included): traverse.cu. */
#define BVHGPU_MAX_PEERS 8
#define BVHGPU_MAILBOX_BYTES 65536
#define BVHGPU_IPC_HANDLE_BYTES 64
#define BVHGPU_SHARD_STAGE_BYTES(nrays_global) ((8200 * ((size_t)(nrays_global) / 2048 + 2 * BVHGPU_MAX_PEERS) + 255) & ~(size_t)255)
typedef struct {
    int rank, world;
    void* peer_counts[BVHGPU_MAX_PEERS];    /* 2 * BVHGPU_SHARD_STAGE_BYTES(nrays_global) on every rank (index = rank): two halves, alternating per step */
    void* peer_hits[BVHGPU_MAX_PEERS];      /* u32[cap] on every rank: the global hit lists                          */
    void* peer_mailbox[BVHGPU_MAX_PEERS];   /* BVHGPU_MAILBOX_BYTES on every rank, zero-initialised                  */
    void* offsets;                          /* LOCAL device memory, u32[nrays_global + 1]: the global CSR offsets     */
    uint64_t seq;                           /* 1, 2, 3, ... identical on all ranks for the same step                 */
    size_t shard_rays[BVHGPU_MAX_PEERS];    /* rays of every rank's shard (shard_rays[rank] == nrays of the call)     */
    size_t cap;                             /* capacity of the global hit buffers                                    */
    int ray_layout;                         /* bvhgpu_ray_layout of dev_rays                                         */
} bvhgpu_shard;
int bvhgpu_traverse_sharded_dev_f32x3(bvhgpu_tree3f* tree, int mode, const void* dev_rays, size_t nrays, const bvhgpu_shard* shard);
int bvhgpu_traverse_sharded_dev_f64x3(bvhgpu_tree3d* tree, int mode, const void* dev_rays, size_t nrays, const bvhgpu_shard* shard);
/* Peer-mappable device memory (plain cudaMalloc + cudaIpcGetMemHandle / cudaIpcOpenMemHandle). */
int bvhgpu_peer_alloc(bvhgpu_ctx* ctx, size_t bytes, void** dev_ptr, void* handle64);
int bvhgpu_peer_open(bvhgpu_ctx* ctx, const void* handle64, void** dev_ptr);
int bvhgpu_peer_close(bvhgpu_ctx* ctx, void* dev_ptr);
int bvhgpu_peer_free(bvhgpu_ctx* ctx, void* dev_ptr);
/* Synchronous device -> host copy on the context's stream (lets a binding read peer-allocated buffers). */
int bvhgpu_memcpy_d2h(bvhgpu_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes);
/* Asynchronous host -> device copy on the context's stream (host memory should be pinned: bvhgpu_host_alloc). */
int bvhgpu_memcpy_h2d_async(bvhgpu_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes);

/* Counters of the last traversal on this tree: [0] node records visited, [1] hits. */
int bvhgpu_traverse_stats_f32x3(bvhgpu_tree3f* tree, uint64_t* out2);
int bvhgpu_traverse_stats_f64x3(bvhgpu_tree3d* tree, uint64_t* out2);

/* ---- the other IntersectsAabb implementors as batched queries ------------------------------
 * Bvh::traverse / FlatBvh::traverse with an Aabb (src/aabb/aabb_impl.rs:240-248, src/aabb/intersection.rs:35-39),
 * a Point (Aabb::contains, src/aabb/aabb_impl.rs:175-177, intersection.rs:41-45) or a Ball (src/ball.rs:85-106) as the
 * query.  `queries` holds n records of 6 T {min,max}, 3 T {point} or 4 T {center, radius}.  Output: CSR as for rays. */
typedef enum { BVHGPU_QUERY_AABB = 1, BVHGPU_QUERY_POINT = 2, BVHGPU_QUERY_BALL = 3 } bvhgpu_query_kind;
int bvhgpu_query_f32x3(bvhgpu_tree3f* tree, int mode, int kind, const float* queries, size_t n,
                       uint32_t* offsets, uint32_t* hits, size_t cap, size_t* total);
int bvhgpu_query_f64x3(bvhgpu_tree3d* tree, int mode, int kind, const double* queries, size_t n,
                       uint32_t* offsets, uint32_t* hits, size_t cap, size_t* total);
int bvhgpu_query_dev_f32x3(bvhgpu_tree3f* tree, int mode, int kind, const void* dev_queries, size_t n,
                           void* dev_offsets, void* dev_hits, size_t cap, size_t* total);
int bvhgpu_query_dev_f64x3(bvhgpu_tree3d* tree, int mode, int kind, const void* dev_queries, size_t n,
                           void* dev_offsets, void* dev_hits, size_t cap, size_t* total);

/* ---- distance-ordered traversal: batched counterpart of Bvh::nearest_traverse_iterator /
 * farthest_traverse_iterator (src/bvh/distance_traverse.rs, src/bvh/bvh_impl.rs).  Per ray: the shapes whose AABB the ray
 * hits (same set as bvhgpu_traverse_*, BVH semantics), sorted by AABB entry distance ascending (`ascending` != 0) or by
 * exit distance descending, with that distance in `dists` (Ray::intersection_slice_for_aabb, src/ray/ray_impl.rs:118-145).
 * The reference iterator is best-effort ("not necessarily perfectly sorted"); this result is perfectly sorted, ties in the
 * reference's DFS order.  Host pointers; `cap` entries in hits and dists. */
int bvhgpu_traverse_ordered_f32x3(bvhgpu_tree3f* tree, const bvh_ray3f* rays, size_t nrays, int ascending,
                                  uint32_t* offsets, uint32_t* hits, float* dists, size_t cap, size_t* total);
int bvhgpu_traverse_ordered_f64x3(bvhgpu_tree3d* tree, const bvh_ray3d* rays, size_t nrays, int ascending,
                                  uint32_t* offsets, uint32_t* hits, double* dists, size_t cap, size_t* total);

/* ---- closest hit with distance pruning: what callers of the reference build from Bvh::traverse (or the distance
 * iterators, src/bvh/distance_traverse.rs, child_distance_traverse.rs) + Ray::intersects_triangle (src/ray/ray_impl.rs:154-213; the
 * loop itself: src/bvh/iter.rs:330-365) -- per ray, front to back, subtrees entered behind the best hit are never opened.
 *   use_triangles == 0: out_shape = the shape whose AABB the ray enters first, key (entry distance as
 *       Ray::intersection_slice_for_aabb, src/ray/ray_impl.rs:118-145, then DFS order) = the first element of a perfectly sorted
 *       nearest_traverse_iterator; out_dist = that entry distance.  Exact (ties are never pruned).
 *   use_triangles != 0: the triangles given with bvhgpu_tree_set_triangles_* (9 scalars per shape: a, b, c; shape i's AABB must
 *       contain triangle i); out_shape = the triangle with the smallest Moeller-Trumbore distance (backface culled, the reference's
 *       operation order, no FMA), ties to the lower index; out_dist = that distance, out_uv (may be NULL) = its u, v.  A subtree is
 *       skipped when its entry distance exceeds best * (1 + 2^-16): results can differ from the unpruned minimum only between hits
 *       whose distances agree to ~1e-5 relative.
 * No hit: out_shape = BVHGPU_INVALID_INDEX, out_dist = +inf. */
int bvhgpu_tree_set_triangles_f32x3(bvhgpu_tree3f* tree, const float* triangles, size_t n);
int bvhgpu_tree_set_triangles_f64x3(bvhgpu_tree3d* tree, const double* triangles, size_t n);
int bvhgpu_tree_set_triangles_dev_f32x3(bvhgpu_tree3f* tree, const void* dev_triangles, size_t n);
int bvhgpu_tree_set_triangles_dev_f64x3(bvhgpu_tree3d* tree, const void* dev_triangles, size_t n);
int bvhgpu_closest_hit_f32x3(bvhgpu_tree3f* tree, const bvh_ray3f* rays, size_t nrays, int use_triangles,
                             uint32_t* out_shape, float* out_dist, float* out_uv);
int bvhgpu_closest_hit_f64x3(bvhgpu_tree3d* tree, const bvh_ray3d* rays, size_t nrays, int use_triangles,
                             uint32_t* out_shape, double* out_dist, double* out_uv);
int bvhgpu_closest_hit_dev_f32x3(bvhgpu_tree3f* tree, const void* dev_rays, int ray_layout, size_t nrays, int use_triangles,
                                 void* dev_shape, void* dev_dist, void* dev_uv);
int bvhgpu_closest_hit_dev_f64x3(bvhgpu_tree3d* tree, const void* dev_rays, int ray_layout, size_t nrays, int use_triangles,
                                 void* dev_shape, void* dev_dist, void* dev_uv);

/* ---- nearest_to: batched Bvh::nearest_to (src/bvh/bvh_impl.rs:221-238, src/bvh/bvh_node.rs:327-372) and
 * FlatBvh::nearest_to (src/flat_bvh.rs:513-562).  The reference calls the shape's own PointDistance::distance_squared at the
 * leaves (user code), so there are two forms.  `points`: 3 T per query point, host pointers.
 *   bvhgpu_nearest_*            for shapes whose distance IS their AABB's (as the reference's UnitBox, src/testbase.rs:101-105):
 *                               the reference's walk replayed exactly (children ordered by Aabb::min_distance_squared,
 *                               src/aabb/aabb_impl.rs:618-629; strict `<`; first minimum kept).  out_shape[i] = shape index
 *                               (BVHGPU_INVALID_INDEX for an empty tree), out_dist[i] = distance (sqrt, as the reference returns).
 *                               `mode` selects Bvh (BVHGPU_TRAVERSE_BVH) or FlatBvh (BVHGPU_TRAVERSE_FLAT) visiting order.
 *   bvhgpu_nearest_candidates_* for ANY shape contained in its AABB: CSR lists that are guaranteed to contain the nearest shape
 *                               of every point (all shapes whose AABB is at most as far as the smallest farthest-corner
 *                               distance of any shape's AABB); the shim evaluates distance_squared on that short list and
 *                               keeps the minimum. */
int bvhgpu_nearest_f32x3(bvhgpu_tree3f* tree, int mode, const float* points, size_t n, uint32_t* out_shape, float* out_dist);
int bvhgpu_nearest_f64x3(bvhgpu_tree3d* tree, int mode, const double* points, size_t n, uint32_t* out_shape, double* out_dist);
/* The same walk with the TRIANGLE's own distance at the leaves -- Triangle::distance_squared of the reference's test shape
 * (closest_point_triangle, src/testbase.rs:353-443), operation for operation; triangles from bvhgpu_tree_set_triangles_*.  This is
 * the PointDistance of every benchmark scene of the reference, evaluated on the device: same shape, bit-identical distance. */
int bvhgpu_nearest_triangles_f32x3(bvhgpu_tree3f* tree, int mode, const float* points, size_t n, uint32_t* out_shape, float* out_dist);
int bvhgpu_nearest_triangles_f64x3(bvhgpu_tree3d* tree, int mode, const double* points, size_t n, uint32_t* out_shape, double* out_dist);
int bvhgpu_nearest_candidates_f32x3(bvhgpu_tree3f* tree, const float* points, size_t n, uint32_t* offsets, uint32_t* cand,
                                    size_t cap, size_t* total);
int bvhgpu_nearest_candidates_f64x3(bvhgpu_tree3d* tree, const double* points, size_t n, uint32_t* offsets, uint32_t* cand,
                                    size_t cap, size_t* total);

/* Ray::new for a batch (src/ray/ray_impl.rs:70-80): normalise, inv = 1/direction. Device pointers. */
int bvhgpu_rays_new_dev_f32x3(bvhgpu_ctx* ctx, const void* dev_origins, const void* dev_directions, size_t n, void* dev_rays);
int bvhgpu_rays_new_dev_f64x3(bvhgpu_ctx* ctx, const void* dev_origins, const void* dev_directions, size_t n, void* dev_rays);

/* ---- whole-tree SAH cost (definition: DESIGN.md; the reference only has the per-split cost,
 * src/bvh/bvh_node.rs:236-238).  out2[0]: with the reference's surface_area (2*|size|^2,
 * src/aabb/aabb_impl.rs:551-554), out2[1]: geometric area. */
int bvhgpu_sah_cost_f32x3(bvhgpu_tree3f* tree, double* out2);
int bvhgpu_sah_cost_f64x3(bvhgpu_tree3d* tree, double* out2);

/* ---- refit: bottom-up AABB update after shapes moved (the data-parallel part of
 * Bvh::update_shapes, src/bvh/optimization.rs:304-351 fix_aabbs_ascending).  Topology is kept. */
int bvhgpu_refit_f32x3(bvhgpu_tree3f* tree, const bvh_aabb3f* aabbs, size_t n);
int bvhgpu_refit_f64x3(bvhgpu_tree3d* tree, const bvh_aabb3d* aabbs, size_t n);
/* The AABBs are already on the device (C-ABI layout): nothing is uploaded.  A NaN in the new AABBs is rejected
 * (BVHGPU_ERR_NAN) before the tree is touched, in every refit / optimize / update variant. */
int bvhgpu_refit_dev_f32x3(bvhgpu_tree3f* tree, const void* dev_aabbs, size_t n);
int bvhgpu_refit_dev_f64x3(bvhgpu_tree3d* tree, const void* dev_aabbs, size_t n);

/* ---- optimize: replaces Bvh::update_shapes after shapes moved (src/bvh/optimization.rs:290-302) ----
 * The reference removes and re-inserts every changed shape sequentially (remove_shape :208-288, add_shape :70-206).
 * The data-parallel counterpart: refit, then rebuild -- in place, with the exact 6-bucket SAH builder -- the
 * outermost subtrees that contain a node whose surface area grew by more than `max_growth` (>= 1; e.g. 1.5).
 * `aabbs` are the CURRENT AABBs of all n shapes (no list of changed indices is needed: unchanged subtrees are
 * found by the growth test).  The node array stays in Bvh::build's preorder layout (the reference's does not,
 * it appends and swap-removes nodes), node indices of shapes in rebuilt subtrees change: fetch them with
 * bvhgpu_tree_nodes_* and pass them to BHShape::set_bh_node_index.  *rebuilt (may be NULL) = number of shapes in
 * the rebuilt subtrees (0: the call was a pure refit).  Not the reference's tree: parity is on the invariants
 * (assert_consistent, assert_tight), on hit sets, and on SAH cost against the oracle's update_shapes. */
int bvhgpu_optimize_f32x3(bvhgpu_tree3f* tree, const bvh_aabb3f* aabbs, size_t n, double max_growth, size_t* rebuilt);
int bvhgpu_optimize_f64x3(bvhgpu_tree3d* tree, const bvh_aabb3d* aabbs, size_t n, double max_growth, size_t* rebuilt);
int bvhgpu_optimize_dev_f32x3(bvhgpu_tree3f* tree, const void* dev_aabbs, size_t n, double max_growth, size_t* rebuilt);
int bvhgpu_optimize_dev_f64x3(bvhgpu_tree3d* tree, const void* dev_aabbs, size_t n, double max_growth, size_t* rebuilt);
/* The same with Bvh::update_shapes' own signature (src/bvh/optimization.rs:304-315: the indices of the changed shapes + the shapes):
 * `changed[i]` is a shape index, `changed_aabbs[i]` its new AABB -- only the m changed shapes cross the boundary (10 M f64 shapes,
 * 1 % moved: 5 MB instead of 480 MB).  max_growth >= 1: refit + rebuild of the degraded subtrees as bvhgpu_optimize_*;
 * max_growth <= 0: refit only.  Indices >= n or NaN AABBs are rejected before the tree is touched.  Growth is judged against the
 * surface area every node had when it was last (re)built, so slow drift over many calls adds up and is rebuilt eventually. */
int bvhgpu_update_f32x3(bvhgpu_tree3f* tree, const uint32_t* changed, const bvh_aabb3f* changed_aabbs, size_t m, double max_growth, size_t* rebuilt);
int bvhgpu_update_f64x3(bvhgpu_tree3d* tree, const uint32_t* changed, const bvh_aabb3d* changed_aabbs, size_t m, double max_growth, size_t* rebuilt);
int bvhgpu_update_dev_f32x3(bvhgpu_tree3f* tree, const void* dev_changed, const void* dev_changed_aabbs, size_t m, double max_growth, size_t* rebuilt);
int bvhgpu_update_dev_f64x3(bvhgpu_tree3d* tree, const void* dev_changed, const void* dev_changed_aabbs, size_t m, double max_growth, size_t* rebuilt);

#ifdef __cplusplus
}
#endif
#endif /* BVH_B200_H */
