/* nbk.h -- C ABI of the B200-native kd-tree build + batched kNN query (libnbk.so).
 *
 * This is the drop-in boundary for nbodyhpc's kdtree hot path.  Plain pointers and sizes only; no
 * CUDA or torch types cross it (streams are passed as opaque `void*` = cudaStream_t).  Each entry
 * point names the reference interface it replaces (paths relative to the reference's kdtree/).
 *
 * Conventions
 *   - every function returning `int` returns NBK_OK (0) or an NBK_ERR_* code; the message of the
 *     last failure on the calling thread is nbk_last_error().  Messages of argument errors use the
 *     reference's wording (pybind.cpp:16-18,42-46,92-98; kdtree.cpp:98-108) so that the host layers
 *     can rethrow them unchanged as std::runtime_error / Python RuntimeError.
 *   - "host" pointers may be pageable or pinned; "device" pointers must live on the tree's device.
 *   - there is no CPU fallback: without a usable sm_100 device every compute entry point fails with
 *     NBK_ERR_CUDA.
 *   - results: squared distances are evaluated in float32 without FMA contraction in the
 *     reference's operation order (kdtree_asm_systemv.asm:76-119), neighbours are the exact top-k
 *     under the total order (d2, index), rows come back ascending with sqrt applied
 *     (kdtree.cpp:149-156); unfilled slots are (sqrt(FLT_MAX), 0xFFFFFFFF) (kdtree_impl.hpp:210).
 */
#ifndef NBK_H
#define NBK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NBK_API __attribute__((visibility("default")))
#else
#define NBK_API
#endif

enum {
    NBK_OK = 0,
    NBK_ERR_INVALID = 1, /* bad argument; message uses the reference's wording where one exists */
    NBK_ERR_CUDA = 2,    /* CUDA runtime / device failure (including "no device")               */
    NBK_ERR_NOMEM = 3    /* host or device allocation failure                                   */
};

typedef struct nbk_tree nbk_tree;

/* Node record, byte-identical to the reference's KDTree::KDTreeNode (kdtree.hpp:149-163). */
typedef struct nbk_node {
    int32_t dim;    /* split dimension, -1 for a leaf                                   */
    float split;    /* split coordinate (internal nodes)                                */
    uint32_t left;  /* leaf: first point position;   internal: left child node index    */
    uint32_t right; /* leaf: one past the last point; internal: right child node index  */
} nbk_node;

/* Fixed-size description of a built tree; what a replica needs besides the arena bytes. */
typedef struct nbk_tree_meta {
    uint64_t n_points;    /* points supplied by the caller                                  */
    uint64_t n_padded;    /* rounded up to a multiple of block_size (pybind.cpp:23)         */
    uint64_t n_nodes;
    uint64_t arena_bytes; /* size of the packed device arena [nodes | 128-byte point tiles]  */
    int32_t leaf_size;    /* as given by the caller (effective = max(leaf_size, 2*block))   */
    int32_t block_size;
    int32_t periodic;
    float box_size;       /* 0 when open (pybind.cpp:80)                                    */
    float lo[3], hi[3];   /* bounding box of the real points (query ordering only)          */
    int32_t n_levels;
    int32_t reserved;
} nbk_tree_meta;

NBK_API const char *nbk_last_error(void);

/* Number of CUDA kernels this library has launched so far in this process (all threads). */
NBK_API uint64_t nbk_launch_count(void);

/* Number of visible CUDA devices, or -1 with nbk_last_error() set. */
NBK_API int nbk_device_count(void);

/* ---- build -------------------------------------------------------------------------------- */

/* Replaces PyKDTree::PyKDTree = make_positions_and_indices + KDTree::KDTree (pybind.cpp:14-56,76-81;
 * kdtree.cpp:95-131).  xyz_aos: n x 3 float32 on the host.  periodic != 0 validates
 * 0 <= x <= box_size on the device.  device = -1 uses the current device.  Returns NULL on failure
 * with *status set. */
NBK_API nbk_tree *nbk_tree_build(const float *xyz_aos, uint64_t n, int leaf_size, int block_size,
                                 int periodic, float box_size, int device, int *status);

/* Same, input already on the device; work is enqueued on `stream` and the call returns once the
 * tree is complete (the topology upload needs one synchronisation). */
NBK_API nbk_tree *nbk_tree_build_device(const float *d_xyz_aos, uint64_t n, int leaf_size,
                                        int block_size, int periodic, float box_size, int device,
                                        void *stream, int *status);

/* Replaces KDTree::KDTree(PositionAndIndexArray<3>, config) (kdtree.cpp:95-131): host SoA columns
 * plus caller-supplied indices, already padded (n_padded % block_size must be 0). */
NBK_API nbk_tree *nbk_tree_build_soa(const float *x, const float *y, const float *z,
                                     const uint32_t *idx, uint64_t n_padded, int leaf_size,
                                     int block_size, int periodic, float box_size, int device,
                                     int *status);

/* Host-only: the tree topology the build will produce for n points (it is a function of counts
 * only, kdtree_impl.hpp:91,101-110).  Writes n_nodes / n_levels and, if `nodes` is not NULL, the
 * node records in the reference's pre-order with dim/left/right filled and split = 0.  Needs no
 * GPU; backs the `size` property before/without a build and the host-logic tests. */
NBK_API int nbk_plan_topology(uint64_t n_points, int leaf_size, int block_size, nbk_node *nodes,
                              uint64_t *n_nodes, int *n_levels);

NBK_API void nbk_tree_free(nbk_tree *tree);

/* ---- introspection ------------------------------------------------------------------------ */

/* Backs the properties n / size / periodic / boxsize (pybind.cpp:71-74,212-215). */
NBK_API int nbk_tree_get_meta(const nbk_tree *tree, nbk_tree_meta *meta);
NBK_API int nbk_tree_device(const nbk_tree *tree);

/* Backs KDTree::nodes() (kdtree.hpp:191): n_nodes records in the reference's pre-order. */
NBK_API int nbk_tree_copy_nodes(const nbk_tree *tree, nbk_node *nodes);

/* Backs KDTree::positions() (kdtree.hpp:192): leaf-ordered SoA columns and the permutation. */
NBK_API int nbk_tree_copy_points(const nbk_tree *tree, float *x, float *y, float *z, uint32_t *idx);

/* ---- query -------------------------------------------------------------------------------- */

/* Replaces PyKDTree::query (pybind.cpp:90-189) and, with m = 1, KDTree::find_closest
 * (kdtree.cpp:133-159).  q_aos: m x 3 float32; out_dist / out_idx: m x k, row-major, HOST memory.
 * Queries are streamed through the device in chunks; copies are inside the call. */
NBK_API int nbk_tree_query(const nbk_tree *tree, const float *q_aos, uint64_t m, int k,
                           float *out_dist, uint32_t *out_idx);

/* Same on DEVICE buffers, enqueued on `stream`, no synchronisation: the benchmark path. */
NBK_API int nbk_tree_query_device(const nbk_tree *tree, const float *d_q_aos, uint64_t m, int k,
                                  float *d_out_dist, uint32_t *d_out_idx, void *stream);

/* As nbk_tree_query, but with the metric chosen per call like the Distance argument of
 * KDTree::find_closest<Distance> (kdtree.hpp:207-210): periodic = 0 -> L2Distance,
 * periodic != 0 -> L2PeriodicDistance<float>{box_size}; periodic < 0 -> the tree's own metric. */
NBK_API int nbk_tree_query_ex(const nbk_tree *tree, const float *q_aos, uint64_t m, int k,
                              int periodic, float box_size, float *out_dist, uint32_t *out_idx);

/* flags of the entry points below */
enum { NBK_QUERY_SQUARED = 1 /* rows hold the squared distances d2 the search ranks by (bit-exact to the
                                reference's L2Distance / L2PeriodicDistance value BEFORE postprocess(),
                                kdtree.hpp:22-31,71-84) instead of sqrt(d2) */ };

/* nbk_tree_query_ex plus `flags` (NBK_QUERY_*). */
NBK_API int nbk_tree_query_ex2(const nbk_tree *tree, const float *q_aos, uint64_t m, int k,
                               int periodic, float box_size, int flags, float *out_dist, uint32_t *out_idx);

/* nbk_tree_query_device plus the per-call metric and `flags`. */
NBK_API int nbk_tree_query_device_ex(const nbk_tree *tree, const float *d_q_aos, uint64_t m, int k,
                                     int periodic, float box_size, int flags, float *d_out_dist,
                                     uint32_t *d_out_idx, void *stream);

/* The leaf scan + top-k container in isolation, on ONE flat block of points: the twin of the
 * reference's C-ABI leaf kernels wenda_insert_closest_l2_avx2 / _periodic_avx2
 * (kdtree_opt_asm.hpp:12-19,39-63; pinned by tests/test_asm.cpp:97-199 and
 * tests/test_inserters.cpp:159-220).  x, y, z, idx: n host values each, n % 8 == 0 (the leaf kernels'
 * own precondition); every query is answered by scanning all n points with the production leaf-scan
 * code, no tree and no traversal.  Rows as nbk_tree_query_ex2.  device = -1: current. */
NBK_API int nbk_scan_block(const float *x, const float *y, const float *z, const uint32_t *idx, uint64_t n,
                           const float *q_aos, uint64_t m, int k, int periodic, float box_size, int flags,
                           float *out_dist, uint32_t *out_idx, int device);

/* Fused kNN-CDF (the consumer of the (M,k) distance rows PyKDTree::query returns,
 * pybind.cpp:179-188).  For every ks[i] (distinct, >= 1) the distance d to the ks[i]-th neighbour
 * of every query is histogrammed on the device instead of being written out:
 *     counts[i * n_bins + b] += #{queries : edges[b] <= d < edges[b+1]}   (last bin closed)
 * i.e. numpy.histogram(dist[:, ks[i]-1], edges) of the rows nbk_tree_query would return, with one
 * traversal at k = max(ks) for all of them and no per-query output traffic.  edges: n_bins + 1
 * non-decreasing float32.  `counts` is accumulated into (the caller zeroes it). */
NBK_API int nbk_tree_knn_cdf(const nbk_tree *tree, const float *q_aos, uint64_t m, const int *ks, int n_ks,
                             const float *edges, int n_bins, uint64_t *counts);

/* Same on DEVICE buffers (queries, edges, counts); ks is a host array.  Enqueued on `stream`; returns
 * after the work has completed. */
NBK_API int nbk_tree_knn_cdf_device(const nbk_tree *tree, const float *d_q_aos, uint64_t m, const int *ks,
                                    int n_ks, const float *d_edges, int n_bins,
                                    unsigned long long *d_counts, void *stream);

/* KDTreeQueryStatistics (kdtree.hpp:124-131; kdtree.cpp:143-147) summed over m host queries:
 * out3 = {nodes_visited, nodes_pruned, points_visited} of the reference's closer-first traversal
 * (kdtree_impl.hpp:226-268) run on this tree. */
NBK_API int nbk_tree_stats(const nbk_tree *tree, const float *q_aos, uint64_t m, int k,
                           int periodic, float box_size, uint64_t *out3);

/* ---- replication (one process per GPU; the bytes move with an NCCL broadcast) -------------- */

/* Device address of the packed arena [nodes | 128-byte point tiles] (meta.arena_bytes bytes). */
NBK_API int nbk_tree_arena(const nbk_tree *tree, void **d_arena, uint64_t *bytes);

/* Allocates an empty tree with the given meta on `device`; the caller fills its arena (e.g. as the
 * destination of ncclBroadcast) before querying it. */
NBK_API nbk_tree *nbk_tree_alloc_replica(const nbk_tree_meta *meta, int device, int *status);

/* A byte-identical copy of `tree` on another device of this process (peer copy over NVLink): the
 * single-process alternative to alloc_replica + ncclBroadcast for callers that drive several GPUs
 * from one process (the reference fans queries out to threads of one process, pybind.cpp:164-172). */
NBK_API nbk_tree *nbk_tree_clone_to_device(const nbk_tree *tree, int device, int *status);

/* ---- device-side timing of the library's own kernels (bench.py's roofline numbers) ---------- */
enum { NBK_SECTION_QUERY_ORDER = 0, /* Morton keys + radix sort of the queries */
       NBK_SECTION_KNN_KERNEL = 1,  /* the kNN traversal kernel                */
       NBK_SECTION_COUNT = 2 };

/* When enabled, every query brackets its sections with CUDA events on the launching stream. */
NBK_API void nbk_profile_enable(int on);

/* Waits for the recorded events, returns the summed duration (ms) and number of recordings of
 * `section` since the last read, and resets it. */
NBK_API int nbk_profile_read(int section, double *total_ms, uint64_t *count);

/* ---- device buffers for callers without a GPU array library (results left on the device) ------ */
NBK_API void *nbk_device_alloc(uint64_t bytes); /* current device; NULL + nbk_last_error() on failure */
NBK_API void *nbk_device_alloc_on(int device, uint64_t bytes); /* on `device` (-1 = current)          */
/* Device ordinal a pointer belongs to, -1 for host / unknown memory. */
NBK_API int nbk_pointer_device(const void *ptr);
NBK_API void nbk_device_free(void *ptr);
/* kind: 0 = host -> device, 1 = device -> host; synchronous */
NBK_API int nbk_device_copy(void *dst, const void *src, uint64_t bytes, int kind);
NBK_API int nbk_device_zero(void *dst, uint64_t bytes);

/* ---- pinned host staging (optional; speeds up the host-buffer entry points) ---------------- */
/* How large PAGEABLE host buffers crossed the boundary so far in this process:
 * out4 = {downloads staged through the pinned ring, downloads copied directly (slow path),
 *         uploads staged, uploads copied directly}.  Buffers below the staging threshold are not counted. */
NBK_API int nbk_host_path_stats(uint64_t *out4);
NBK_API void *nbk_host_alloc(uint64_t bytes);
NBK_API void nbk_host_free(void *ptr);
/* Page-locks / releases a caller-owned host range (cudaHostRegister, portable) so that the host-buffer entry
 * points copy to and from it directly.  `device`: a device whose context exists (-1 = current). */
NBK_API int nbk_host_register(void *ptr, uint64_t bytes, int device);
NBK_API int nbk_host_unregister(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* NBK_H */
