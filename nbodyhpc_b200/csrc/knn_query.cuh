// Batched kNN query kernels (sm_100a).
//
// Replaces PyKDTree::query -> KDTree::find_closest -> KDTreeQuery::compute -> asm leaf scan
// (pybind.cpp:90-189, kdtree.cpp:133-159, kdtree_impl.hpp:212-268, kdtree_asm_systemv.asm:121-248).
//
// Queries are ordered along a Morton curve, so the 32 lanes of a warp hold spatial neighbours.
//
//   knn_lane_kernel   (default) one query per lane, every lane walks the tree on its own; neighbouring
//                     lanes mostly visit the same nodes and leaf tiles at the same time, so a warp's
//                     loads collapse to a few broadcast sectors.
//   knn_packet_kernel (NBK_KERNEL=packet) one warp walks the tree once for its 32 queries with a
//                     warp-uniform stack; kept as the measured alternative.
//
// The top-k of a lane is a register-resident sorted list of 64-bit keys (d2 bits << 32 | index),
// so the result is the exact top-k under the total order (d2, index) whatever the visiting order --
// which is what makes the answer independent of the traversal and equal to the reference's (whose
// own result is traversal dependent only for exact d2 ties).
//
// Arithmetic contract: d2 is computed with __fsub_rn/__fmul_rn/__fadd_rn (never contracted into
// FMA) in the reference's order ((dx2 + dy2) + dz2) with d = p - q (kdtree_asm_systemv.asm:76-87),
// periodic axis term min(d^2, (d+L)^2, (d-L)^2) (kdtree_asm_systemv.asm:89-119).
#pragma once

#include "common.cuh"

namespace nbk {

// Leaf-ordered points live in 128-byte tiles of 8 points: {x[8], y[8], z[8], idx[8]}.  Leaves start
// on multiples of 8 points and hold a multiple of 8 (block_size 8, kdtree_impl.hpp:108-110), so a
// leaf is a run of whole tiles and one tile is exactly one cache line.
constexpr int kTilePoints = 8;
constexpr int kTileFloat4 = 8; // float4 per tile: x0 x1 y0 y1 z0 z1 i0 i1

struct QueryTree {
    const nbk_node *nodes;
    const float4 *tiles;
    bool periodic;
    float box;          // periodic box size
    float lo[3], hi[3]; // root cell: periodic [0, box]; open [-FLT_MAX, FLT_MAX]
};

__device__ __forceinline__ float tile_coord(const float4 *tiles, uint32_t p, int axis) {
    return reinterpret_cast<const float *>(tiles)[(uint64_t)(p >> 3) * 32 + axis * 8 + (p & 7)];
}

#ifndef NBK_QUERY_THREADS
#define NBK_QUERY_THREADS 128
#endif
#ifndef NBK_MERGED_INSERT
#define NBK_MERGED_INSERT 0 // 1: one insertion site per half tile (measured alternative, see scan_half_tile)
#endif
constexpr int kQueryThreads = NBK_QUERY_THREADS;
constexpr int kQueryWarps = kQueryThreads / 32;

// ---- Morton ordering of the queries --------------------------------------------------------------
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__device__ __forceinline__ uint32_t morton_key(const float *__restrict__ q_aos, uint64_t i, float lo0, float lo1,
                                               float lo2, float s0, float s1, float s2) {
    const float qx = q_aos[3 * i], qy = q_aos[3 * i + 1], qz = q_aos[3 * i + 2];
    const int cx = min(max((int)((qx - lo0) * s0), 0), 1023);
    const int cy = min(max((int)((qy - lo1) * s1), 0), 1023);
    const int cz = min(max((int)((qz - lo2) * s2), 0), 1023);
    return spread10((uint32_t)cx) | (spread10((uint32_t)cy) << 1) | (spread10((uint32_t)cz) << 2);
}

__global__ void __launch_bounds__(256)
morton_keys_kernel(const float *__restrict__ q_aos, uint64_t m, float lo0, float lo1, float lo2,
                   float s0, float s1, float s2, uint32_t *__restrict__ keys,
                   uint32_t *__restrict__ vals) {
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    keys[i] = morton_key(q_aos, i, lo0, lo1, lo2, s0, s1, s2);
    vals[i] = (uint32_t)i;
}

// Keys + the digit totals of every ordering pass (rs::sweep_order): 4096 queries per CTA, the
// histograms privatised in shared memory.
constexpr int kKeysPerCta = 4096;
constexpr int kMaxOrderPasses = 4;
__global__ void __launch_bounds__(256)
morton_keys_totals_kernel(const float *__restrict__ q_aos, uint64_t m, float lo0, float lo1, float lo2,
                          float s0, float s1, float s2, int first_bit, int passes,
                          uint32_t *__restrict__ keys, uint32_t *__restrict__ digit_totals) {
    __shared__ uint32_t h[kMaxOrderPasses][256];
    for (int p = 0; p < passes; ++p) h[p][threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kKeysPerCta;
#pragma unroll 4
    for (int it = 0; it < kKeysPerCta / 256; ++it) {
        const uint64_t i = base + it * 256 + threadIdx.x;
        if (i < m) {
            const uint32_t key = morton_key(q_aos, i, lo0, lo1, lo2, s0, s1, s2);
            keys[i] = key;
            for (int p = 0; p < passes; ++p) atomicAdd(&h[p][(key >> (first_bit + 8 * p)) & 255u], 1u);
        }
    }
    __syncthreads();
    for (int p = 0; p < passes; ++p) {
        const uint32_t c = h[p][threadIdx.x];
        if (c) atomicAdd(&digit_totals[p * 256 + threadIdx.x], c);
    }
}

// ---- point distance -------------------------------------------------------------------------------
__device__ __forceinline__ float d2_open(float px, float py, float pz, float qx, float qy,
                                         float qz) {
    float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// min(d^2, (d+L)^2, (d-L)^2) == min(d^2, (d - copysign(L, d))^2) bit for bit: the dropped candidate
// has magnitude >= |d| (monotone rounding), so it can never be the strict minimum.
__device__ __forceinline__ float axis_periodic(float p, float q, float L) {
    float d = __fsub_rn(p, q);
    float w = __fsub_rn(d, copysignf(L, d));
    return fminf(__fmul_rn(d, d), __fmul_rn(w, w));
}

__device__ __forceinline__ float d2_periodic(float px, float py, float pz, float qx, float qy,
                                             float qz, float L) {
    return __fadd_rn(__fadd_rn(axis_periodic(px, qx, L), axis_periodic(py, qy, L)),
                     axis_periodic(pz, qz, L));
}

template <bool WRAP>
__device__ __forceinline__ void d2x4(float4 const &X, float4 const &Y, float4 const &Z, float qx,
                                     float qy, float qz, float L, float d[4]) {
    if (WRAP) {
        d[0] = d2_periodic(X.x, Y.x, Z.x, qx, qy, qz, L);
        d[1] = d2_periodic(X.y, Y.y, Z.y, qx, qy, qz, L);
        d[2] = d2_periodic(X.z, Y.z, Z.z, qx, qy, qz, L);
        d[3] = d2_periodic(X.w, Y.w, Z.w, qx, qy, qz, L);
    } else {
        d[0] = d2_open(X.x, Y.x, Z.x, qx, qy, qz);
        d[1] = d2_open(X.y, Y.y, Z.y, qx, qy, qz);
        d[2] = d2_open(X.z, Y.z, Z.z, qx, qy, qz);
        d[3] = d2_open(X.w, Y.w, Z.w, qx, qy, qz);
    }
}

// Out of line on purpose: taken only for points more than L/2 away along some axis, and keeping it
// out of the leaf loop keeps that loop's register footprint at the open-metric level.
__device__ __noinline__ float4 d2x4_wrapped(float4 X, float4 Y, float4 Z, float qx, float qy, float qz,
                                            float L) {
    float d[4];
    d2x4<true>(X, Y, Z, qx, qy, qz, L, d);
    return make_float4(d[0], d[1], d[2], d[3]);
}

// ---- per-thread binding of a top-k container ---------------------------------------------------------
struct TopBind {
    unsigned long long *smem; // the CTA's dynamic shared memory (HeapT<K > 0>)
    unsigned long long *gmem; // global heap scratch [k][gcolumns] (HeapT<0>: any k)
    uint32_t gcolumns;
};

// ---- register-resident top-k ----------------------------------------------------------------------
// keys sorted ascending; key = (d2 bits << 32) | index, held as two 32-bit halves.  d2 >= +0 so the
// bit pattern orders like the float.  Empty slots hold (FLT_MAX, 0): a candidate at d2 == FLT_MAX
// never replaces one (the reference inserts only if d2 < FLT_MAX, kdtree_impl.hpp:210 + strict <).
// k < K (k = 3, 5, 6, 7): the first K - k slots hold the minimal key (0, 0) and never move, so that
// worst() is the k-th best and not the K-th (the reference prunes with the k-th, kdtree_impl.hpp:243-262).
template <int K> struct TopK {
    uint32_t hi[K], lo[K];
    __device__ __forceinline__ void init(int k) {
        const int pad = K - k;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            hi[j] = j < pad ? 0u : kFltMaxBits;
            lo[j] = 0u;
        }
    }
    __device__ __forceinline__ float worst() const { return __uint_as_float(hi[K - 1]); }
    __device__ __forceinline__ bool beats_worst(uint32_t chi, uint32_t clo) const {
        return chi < hi[K - 1] || (chi == hi[K - 1] && clo < lo[K - 1]);
    }
    __device__ __forceinline__ bool contains(uint32_t chi, uint32_t clo, int k) const {
        const int pad = K - k;
        bool hit = false;
#pragma unroll
        for (int j = 0; j < K; ++j) hit = hit || (j >= pad && hi[j] == chi && lo[j] == clo);
        return hit;
    }
    static constexpr int kSize = K;
    static constexpr bool kShared = false;
    static constexpr bool kGlobal = false;
    static constexpr bool kNeedsIndex = true;
    static constexpr int kKeyBytes = 8;
    __device__ __forceinline__ void bind(TopBind const &) {}
    // f(j, d2 bits) for every rank j (0 = nearest); any order
    template <typename F> __device__ __forceinline__ void for_each_rank(int k, F &&f) {
        const int pad = K - k;
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (j >= pad) f(j - pad, hi[j]);
    }
    // raw: squared distances (the d2 the search ranks by) instead of postprocess()'s sqrt
    __device__ __forceinline__ void write_row(uint32_t qid, int k, bool raw, float *__restrict__ out_d,
                                              uint32_t *__restrict__ out_i) const {
        const int pad = K - k;
        float *od = out_d + (uint64_t)qid * k - pad;
        uint32_t *oi = out_i + (uint64_t)qid * k - pad;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (j >= pad) {
                const float d2 = __uint_as_float(hi[j]);
                od[j] = raw ? d2 : __fsqrt_rn(d2); // postprocess, kdtree.cpp:154-156
                oi[j] = hi[j] == kFltMaxBits ? 0xFFFFFFFFu : lo[j];
            }
        }
    }
    // the inverse of write_row(raw = true): resume from a row a previous pass left in the output
    __device__ __forceinline__ void load_row(uint32_t qid, int k, const float *out_d, const uint32_t *out_i) {
        const int pad = K - k;
        const float *od = out_d + (uint64_t)qid * k - pad;
        const uint32_t *oi = out_i + (uint64_t)qid * k - pad;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            hi[j] = 0u;
            lo[j] = 0u;
            if (j >= pad) {
                hi[j] = __float_as_uint(od[j]);
                lo[j] = hi[j] == kFltMaxBits ? 0u : oi[j];
            }
        }
    }
    // precondition: beats_worst(chi, clo).  One 64-bit compare (2 instructions) + 4 selects per step.
    __device__ __forceinline__ void insert(uint32_t chi, uint32_t clo) {
        hi[K - 1] = chi;
        lo[K - 1] = clo;
#pragma unroll
        for (int j = K - 1; j > 0; --j) {
            uint32_t ah = hi[j - 1], al = lo[j - 1], bh = hi[j], bl = lo[j];
            asm("{\n\t.reg .pred p;\n\t.reg .b64 ka, kb;\n\t"
                "mov.b64 ka, {%4, %5};\n\tmov.b64 kb, {%6, %7};\n\t"
                "setp.lt.u64 p, kb, ka;\n\t"
                "selp.b32 %0, %6, %4, p;\n\tselp.b32 %1, %7, %5, p;\n\t"
                "selp.b32 %2, %4, %6, p;\n\tselp.b32 %3, %5, %7, p;\n\t}"
                : "=&r"(lo[j - 1]), "=&r"(hi[j - 1]), "=&r"(lo[j]), "=&r"(hi[j])
                : "r"(al), "r"(ah), "r"(bl), "r"(bh));
        }
    }
};

// ---- heap top-k for large k ---------------------------------------------------------------------------
// For k > 8 a sorted register list costs O(k) per insertion and most of the register file.  This
// variant keeps each lane's k best in a binary MAX-heap of exactly k slots, slot-major so that any
// per-lane slot pattern is conflict free / coalesced, and replaces the current worst in O(log k) like
// the reference's tournament tree (tournament_tree.hpp:49-91); the root (= current k-th best) is
// mirrored in registers.  The row is produced by a heap sort in the epilogue.
//   KS > 0: heap in shared memory, heap[slot * blockDim + thread], k <= KS
//   KS = 0: heap in global memory (L2-resident scratch), heap[slot * columns + column]: any k, like the
//           reference's queue (kdtree.cpp:133-141 accepts every k)
template <int KS> struct HeapT {
    static constexpr int kSize = KS;
    static constexpr bool kShared = KS > 0;
    static constexpr bool kGlobal = KS == 0;
    static constexpr bool kNeedsIndex = true;
    static constexpr int kKeyBytes = 8;
    unsigned long long *heap; // this thread's column
    uint32_t stride;
    int n;            // heap size = k
    uint32_t rhi, rlo; // root
    __device__ __forceinline__ void bind(TopBind const &b) {
        if (KS > 0) {
            heap = b.smem + threadIdx.x;
            stride = blockDim.x;
        } else {
            heap = b.gmem + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
            stride = b.gcolumns;
        }
    }
    __device__ __forceinline__ unsigned long long &at(int slot) const { return heap[(uint64_t)slot * stride]; }
    __device__ __forceinline__ void init(int k) {
        n = k;
        for (int j = 0; j < n; ++j) at(j) = (unsigned long long)kFltMaxBits << 32;
        rhi = kFltMaxBits;
        rlo = 0u;
    }
    __device__ __forceinline__ float worst() const { return __uint_as_float(rhi); }
    __device__ __forceinline__ bool beats_worst(uint32_t chi, uint32_t clo) const {
        return chi < rhi || (chi == rhi && clo < rlo);
    }
    __device__ __forceinline__ bool contains(uint32_t chi, uint32_t clo, int) const {
        const unsigned long long c = ((unsigned long long)chi << 32) | clo;
        bool hit = false;
        for (int j = 0; j < n; ++j) hit = hit || at(j) == c;
        return hit;
    }
    // sift `c` down from the root of a heap of `size` slots
    __device__ __forceinline__ void sift_down(unsigned long long c, int size) {
        int i = 0;
        while (true) {
            const int l = 2 * i + 1;
            if (l >= size) break;
            const int r = l + 1;
            const unsigned long long kl = at(l);
            const unsigned long long kr = r < size ? at(r) : 0ull;
            const bool right = kr > kl;
            const unsigned long long kb = right ? kr : kl;
            if (kb <= c) break;
            at(i) = kb;
            i = right ? r : l;
        }
        at(i) = c;
    }
    // precondition: beats_worst(chi, clo): the root (current worst) is replaced
    __device__ __forceinline__ void insert(uint32_t chi, uint32_t clo) {
        sift_down(((unsigned long long)chi << 32) | clo, n);
        const unsigned long long root = at(0);
        rhi = (uint32_t)(root >> 32);
        rlo = (uint32_t)root;
    }
    // f(j, d2 bits) for every rank j (0 = nearest), largest first; consumes the heap
    template <typename F> __device__ __forceinline__ void for_each_rank(int, F &&f) {
        for (int size = n; size > 0; --size) {
            const unsigned long long top = at(0);
            f(size - 1, (uint32_t)(top >> 32));
            if (size > 1) sift_down(at(size - 1), size - 1);
        }
    }
    __device__ __forceinline__ void write_row(uint32_t qid, int k, bool raw, float *__restrict__ out_d,
                                              uint32_t *__restrict__ out_i) {
        float *od = out_d + (uint64_t)qid * k;
        uint32_t *oi = out_i + (uint64_t)qid * k;
        // heap sort: the maximum leaves first and lands at the end of the row
        for (int size = n; size > 0; --size) {
            const unsigned long long top = at(0);
            const uint32_t h = (uint32_t)(top >> 32);
            const float d2 = __uint_as_float(h);
            od[size - 1] = raw ? d2 : __fsqrt_rn(d2); // postprocess, kdtree.cpp:154-156
            oi[size - 1] = h == kFltMaxBits ? 0xFFFFFFFFu : (uint32_t)top;
            if (size > 1) sift_down(at(size - 1), size - 1);
        }
    }
    // the inverse of write_row(raw = true): a row in descending order is a valid max-heap
    __device__ __forceinline__ void load_row(uint32_t qid, int k, const float *out_d, const uint32_t *out_i) {
        n = k;
        const float *od = out_d + (uint64_t)qid * k;
        const uint32_t *oi = out_i + (uint64_t)qid * k;
        for (int j = 0; j < n; ++j) {
            const uint32_t h = __float_as_uint(od[n - 1 - j]);
            at(j) = ((unsigned long long)h << 32) | (h == kFltMaxBits ? 0u : oi[n - 1 - j]);
        }
        const unsigned long long root = at(0);
        rhi = (uint32_t)(root >> 32);
        rlo = (uint32_t)root;
    }
};
template <int K> using HeapK = HeapT<K>;

// ---- distance-only containers: the fast pass of the fused kNN-CDF -------------------------------------------
// The histograms need the k smallest DISTANCES of a query, as a multiset, and nothing else: no indices, and
// equal distances need no order (inserting a candidate that ties the current worst or not gives the same
// multiset).  So the keys are the 32-bit d2 patterns: half the shared memory per lane, one compare instead of
// two per step, no index loads in the leaf scan.  A point must not be met twice (there is nothing to
// recognise it by), which holds for the fast pass (one image); the boundary pass keeps the (d2, index)
// containers above.
template <int K> struct DistTopK { // k <= 8: sorted register list, two min/max per insertion step
    uint32_t hi[K];
    static constexpr int kSize = K;
    static constexpr bool kShared = false;
    static constexpr bool kGlobal = false;
    static constexpr bool kNeedsIndex = false;
    static constexpr int kKeyBytes = 4;
    __device__ __forceinline__ void bind(TopBind const &) {}
    __device__ __forceinline__ void init(int k) {
        const int pad = K - k;
#pragma unroll
        for (int j = 0; j < K; ++j) hi[j] = j < pad ? 0u : kFltMaxBits;
    }
    __device__ __forceinline__ float worst() const { return __uint_as_float(hi[K - 1]); }
    __device__ __forceinline__ bool beats_worst(uint32_t chi, uint32_t) const { return chi < hi[K - 1]; }
    __device__ __forceinline__ bool contains(uint32_t, uint32_t, int) const { return false; }
    __device__ __forceinline__ void insert(uint32_t chi, uint32_t) {
        hi[K - 1] = chi;
#pragma unroll
        for (int j = K - 1; j > 0; --j) {
            const uint32_t a = hi[j - 1], b = hi[j];
            hi[j - 1] = min(a, b);
            hi[j] = max(a, b);
        }
    }
    template <typename F> __device__ __forceinline__ void for_each_rank(int k, F &&f) {
        const int pad = K - k;
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (j >= pad) f(j - pad, hi[j]);
    }
    __device__ __forceinline__ void write_row(uint32_t, int, bool, float *, uint32_t *) const {} // CDF only
    __device__ __forceinline__ void load_row(uint32_t, int, const float *, const uint32_t *) {}
};

template <int KS> struct DistHeapT { // 8 < k <= KS: max-heap of 32-bit keys in shared memory
    static constexpr int kSize = KS;
    static constexpr bool kShared = true;
    static constexpr bool kGlobal = false;
    static constexpr bool kNeedsIndex = false;
    static constexpr int kKeyBytes = 4;
    uint32_t *heap; // this thread's column, heap[slot * blockDim + thread]
    uint32_t stride;
    int n;
    uint32_t root;
    __device__ __forceinline__ void bind(TopBind const &b) {
        heap = reinterpret_cast<uint32_t *>(b.smem) + threadIdx.x;
        stride = blockDim.x;
    }
    __device__ __forceinline__ uint32_t &at(int slot) const { return heap[(uint32_t)slot * stride]; }
    __device__ __forceinline__ void init(int k) {
        n = k;
        for (int j = 0; j < n; ++j) at(j) = kFltMaxBits;
        root = kFltMaxBits;
    }
    __device__ __forceinline__ float worst() const { return __uint_as_float(root); }
    __device__ __forceinline__ bool beats_worst(uint32_t chi, uint32_t) const { return chi < root; }
    __device__ __forceinline__ bool contains(uint32_t, uint32_t, int) const { return false; }
    __device__ __forceinline__ void sift_down(uint32_t c, int size) {
        int i = 0;
        while (true) {
            const int l = 2 * i + 1;
            if (l >= size) break;
            const int r = l + 1;
            const uint32_t kl = at(l);
            const uint32_t kr = r < size ? at(r) : 0u;
            const bool right = kr > kl;
            const uint32_t kb = right ? kr : kl;
            if (kb <= c) break;
            at(i) = kb;
            i = right ? r : l;
        }
        at(i) = c;
    }
    __device__ __forceinline__ void insert(uint32_t chi, uint32_t) {
        sift_down(chi, n);
        root = at(0);
    }
    template <typename F> __device__ __forceinline__ void for_each_rank(int, F &&f) { // largest first; consumes the heap
        for (int size = n; size > 0; --size) {
            f(size - 1, at(0));
            if (size > 1) sift_down(at(size - 1), size - 1);
        }
    }
    __device__ __forceinline__ void write_row(uint32_t, int, bool, float *, uint32_t *) {} // CDF only
    __device__ __forceinline__ void load_row(uint32_t, int, const float *, const uint32_t *) {}
};

// Scans the tiles [begin, end) of one leaf for this lane's query.  Per 4 points: x, y, z and the
// indices are four 16-byte loads from one 128-byte tile.  PERIODIC: a wrapped image can only win on
// an axis with |p - q| > L/2, i.e. when the open d2 is >= (L/2)^2 = wrap_d2; only then is the
// 3-image formula evaluated.  `dedupe` is set while a shifted image is searched.
template <typename Top, bool PERIODIC, bool STAGED = false>
__device__ __forceinline__ void scan_half_tile(const float4 *g, float qx, float qy, float qz, float L,
                                               float wrap_d2, bool dedupe, int k, Top &top) {
    // STAGED: g points into shared memory (a leaf staged by a TMA bulk copy, see NBK_STAGE_HOME)
    const float4 X = STAGED ? g[0] : __ldg(g), Y = STAGED ? g[2] : __ldg(g + 2), Z = STAGED ? g[4] : __ldg(g + 4);
    float d[4];
    d2x4<false>(X, Y, Z, qx, qy, qz, L, d);
    if (PERIODIC) {
        const float dmax = fmaxf(fmaxf(d[0], d[1]), fmaxf(d[2], d[3]));
        if (dmax >= wrap_d2) {
            const float4 w = d2x4_wrapped(X, Y, Z, qx, qy, qz, L);
            d[0] = w.x; d[1] = w.y; d[2] = w.z; d[3] = w.w;
        }
    }
    const float dmin = fminf(fminf(d[0], d[1]), fminf(d[2], d[3]));
    if (dmin <= top.worst()) {
        // the indices are needed by about one half-tile in five: loading them here (same 128-byte line
        // as the coordinates, an L1 hit) instead of up front takes a quarter off the L1 wavefronts, which
        // ncu shows at 76 % of peak -- the kernel's co-limiter next to instruction issue
        uint4 I = make_uint4(0u, 0u, 0u, 0u);
        if (Top::kNeedsIndex) I = STAGED ? *reinterpret_cast<const uint4 *>(g + 6) : __ldg(reinterpret_cast<const uint4 *>(g + 6));
        const uint32_t idx[4] = {I.x, I.y, I.z, I.w};
#if NBK_MERGED_INSERT
        // ONE insertion site per half tile, fed by a per-lane loop over the lane's candidates (selected
        // with SEL chains, nothing indexed dynamically): the site runs max-over-lanes(#candidates) times
        // instead of once per point position at which any lane has a candidate.
        const float w = top.worst();
        unsigned m = (d[0] <= w ? 1u : 0u) | (d[1] <= w ? 2u : 0u) | (d[2] <= w ? 4u : 0u) | (d[3] <= w ? 8u : 0u);
        while (m) {
            const unsigned low = m & (0u - m);
            m ^= low;
            const float dj = low == 1u ? d[0] : (low == 2u ? d[1] : (low == 4u ? d[2] : d[3]));
            const uint32_t ij = low == 1u ? idx[0] : (low == 2u ? idx[1] : (low == 4u ? idx[2] : idx[3]));
            const uint32_t chi = __float_as_uint(dj);
            if (top.beats_worst(chi, ij) && !(dedupe && top.contains(chi, ij, k))) top.insert(chi, ij);
        }
#else
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t chi = __float_as_uint(d[j]);
            if (top.beats_worst(chi, idx[j]) && !(dedupe && top.contains(chi, idx[j], k)))
                top.insert(chi, idx[j]);
        }
#endif
    }
}

template <typename Top, bool PERIODIC>
__device__ __forceinline__ void scan_leaf(QueryTree const &t, uint32_t begin, uint32_t end, float qx,
                                          float qy, float qz, float wrap_d2, bool dedupe, int k, Top &top) {
    const float4 *tp = t.tiles + (uint64_t)(begin >> 3) * kTileFloat4;
    const float4 *const te = t.tiles + (uint64_t)(end >> 3) * kTileFloat4;
    for (; tp != te; tp += kTileFloat4) {
        scan_half_tile<Top, PERIODIC>(tp, qx, qy, qz, t.box, wrap_d2, dedupe, k, top);
        scan_half_tile<Top, PERIODIC>(tp + 1, qx, qy, qz, t.box, wrap_d2, dedupe, k, top);
    }
}

// ---- candidate queue: the scan appends, the warp inserts together ---------------------------------------
// With insertion inside the scan loop a 44-instruction insertion site runs whenever ANY lane of the warp
// has a candidate at that point position (ncu, round 1: 138 site executions per warp for 30 insertions
// per lane, 7 of 32 lanes active).  Here the scan only appends (d2, index) keys that pass the float test
// against the lane's current k-th distance to a small per-lane queue in shared memory, and the warp then
// drains all queues together: one insertion site, executed max-over-lanes(queue length) times with most
// lanes busy.  The k-th distance is stale while a queue fills, so a few candidates are appended that
// an immediate insertion would have rejected; the exact 64-bit test is repeated at the drain.  In the
// first (home-leaf) round, where the bound falls fastest, the warp drains after every 8-point tile.
#ifndef NBK_QUEUE_CAP
#define NBK_QUEUE_CAP 0 // 0: insert inside the scan loop (no queue)
#endif
constexpr int kQueueCap = NBK_QUEUE_CAP;
#ifndef NBK_QUEUE_HOME
#define NBK_QUEUE_HOME 1 // 0: the home-leaf round inserts directly, only the later leaves are queued
#endif

struct CandQueue { // slot-major like the heaps: queue[slot * blockDim + thread]
    unsigned long long *col;
    int cnt;
    __device__ __forceinline__ void bind(unsigned long long *base) {
        col = base + threadIdx.x;
        cnt = 0;
    }
    __device__ __forceinline__ void push(uint32_t chi, uint32_t clo) {
        col[(uint32_t)cnt * kQueryThreads] = ((unsigned long long)chi << 32) | clo;
        ++cnt;
    }
};

template <typename Top>
__device__ __forceinline__ void drain_own(Top &top, CandQueue &cq, bool dedupe, int k) {
    while (cq.cnt > 0) {
        --cq.cnt;
        const unsigned long long key = cq.col[(uint32_t)cq.cnt * kQueryThreads];
        const uint32_t chi = (uint32_t)(key >> 32), clo = (uint32_t)key;
        if (top.beats_worst(chi, clo) && !(dedupe && top.contains(chi, clo, k))) top.insert(chi, clo);
    }
}

// all 32 lanes call this together
template <typename Top>
__device__ __forceinline__ void drain_warp(Top &top, CandQueue &cq, bool dedupe, int k) {
    while (__any_sync(0xffffffffu, cq.cnt > 0)) {
        if (cq.cnt > 0) {
            --cq.cnt;
            const unsigned long long key = cq.col[(uint32_t)cq.cnt * kQueryThreads];
            const uint32_t chi = (uint32_t)(key >> 32), clo = (uint32_t)key;
            if (top.beats_worst(chi, clo) && !(dedupe && top.contains(chi, clo, k))) top.insert(chi, clo);
        }
    }
}

template <typename Top, bool PERIODIC>
__device__ __forceinline__ void append_half_tile(const float4 *g, float qx, float qy, float qz, float L,
                                                 float wrap_d2, float bound, CandQueue &cq) {
    const float4 X = __ldg(g), Y = __ldg(g + 2), Z = __ldg(g + 4);
    float d[4];
    d2x4<false>(X, Y, Z, qx, qy, qz, L, d);
    if (PERIODIC) {
        const float dmax = fmaxf(fmaxf(d[0], d[1]), fmaxf(d[2], d[3]));
        if (dmax >= wrap_d2) {
            const float4 w = d2x4_wrapped(X, Y, Z, qx, qy, qz, L);
            d[0] = w.x; d[1] = w.y; d[2] = w.z; d[3] = w.w;
        }
    }
    const float dmin = fminf(fminf(d[0], d[1]), fminf(d[2], d[3]));
    if (dmin <= bound) {
        const uint4 I = __ldg(reinterpret_cast<const uint4 *>(g + 6));
        const uint32_t idx[4] = {I.x, I.y, I.z, I.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (d[j] <= bound) cq.push(__float_as_uint(d[j]), idx[j]); // non-strict: ties are settled at the drain
    }
}

// All 32 lanes call this together; lanes without a leaf pass begin == end.  `every_tile`: warp-uniform.
template <typename Top, bool PERIODIC>
__device__ __forceinline__ void scan_leaf_queued(QueryTree const &t, uint32_t begin, uint32_t end, float qx, float qy,
                                                 float qz, float wrap_d2, bool dedupe, bool every_tile, int k,
                                                 Top &top, CandQueue &cq) {
    const float4 *tp = t.tiles + (uint64_t)(begin >> 3) * kTileFloat4;
    const float4 *const te = t.tiles + (uint64_t)(end >> 3) * kTileFloat4;
    if (every_tile) {
        while (__any_sync(0xffffffffu, tp != te)) {
            if (tp != te) {
                const float bound = top.worst();
                append_half_tile<Top, PERIODIC>(tp, qx, qy, qz, t.box, wrap_d2, bound, cq);
                append_half_tile<Top, PERIODIC>(tp + 1, qx, qy, qz, t.box, wrap_d2, bound, cq);
                tp += kTileFloat4;
            }
            drain_warp(top, cq, dedupe, k);
        }
    } else {
        for (; tp != te; tp += kTileFloat4) {
            if (cq.cnt > kQueueCap - 8) drain_own(top, cq, dedupe, k); // rare: room for one more tile
            const float bound = top.worst();
            append_half_tile<Top, PERIODIC>(tp, qx, qy, qz, t.box, wrap_d2, bound, cq);
            append_half_tile<Top, PERIODIC>(tp + 1, qx, qy, qz, t.box, wrap_d2, bound, cq);
        }
        __syncwarp();
        drain_warp(top, cq, dedupe, k);
    }
}

// ---- measured alternative: the home leaf staged in shared memory by TMA (NBK_STAGE_HOME=1) -------------
// When all queries of a warp live in the same leaf (one leaf holds ~1.5 warps of queries at the headline
// density), one elected lane issues ONE bulk copy global -> shared of the leaf's tiles (cp.async.bulk with
// an mbarrier transaction count), the warp waits on the mbarrier and all 32 lanes scan the staged copy with
// broadcast shared-memory loads.  Other warps and all later leaves use the global loads.
#ifndef NBK_STAGE_HOME
#define NBK_STAGE_HOME 0
#endif
constexpr uint32_t kStagePoints = 128;                      // largest leaf that is staged
constexpr uint32_t kStageBytesPerWarp = kStagePoints * 16;  // tiles: 16 bytes per point
constexpr size_t kStageSmem = NBK_STAGE_HOME ? (size_t)kQueryWarps * kStageBytesPerWarp + 8 * kQueryWarps : 0;

struct StageBuf {
    float4 *tiles; // this warp's staging area
    uint32_t mbar; // shared-memory address of this warp's mbarrier
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_init(StageBuf &sb, unsigned long long *smem) {
    unsigned char *base = reinterpret_cast<unsigned char *>(smem);
    const int warp = threadIdx.x >> 5;
    sb.tiles = reinterpret_cast<float4 *>(base + (size_t)warp * kStageBytesPerWarp);
    sb.mbar = smem_addr(base + (size_t)kQueryWarps * kStageBytesPerWarp + 8 * warp);
    if ((threadIdx.x & 31) == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb.mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
}

// one lane: arm the barrier with the byte count, start the bulk copy
__device__ __forceinline__ void stage_load(StageBuf const &sb, const void *src, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb.mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(sb.tiles)), "l"(src), "r"(bytes), "r"(sb.mbar)
                 : "memory");
}

__device__ __forceinline__ void stage_wait(StageBuf const &sb, uint32_t phase) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(sb.mbar), "r"(phase)
                     : "memory");
        if (!done && ++spins > (1u << 22)) __trap(); // never hang the device
    }
}

// ---- the lane kernel: one query per lane, independent traversal -----------------------------------
//
// Pruning bound.  A cell is described by per-axis terms (t0, t1, t2) whose ordered sum
// fl(fl(t0 + t1) + t2) never exceeds the d2 of any point inside.  Descending, only the split axis
// changes: for the far child the term becomes w^2 with w = fl(fl(split - q) + o), the same monotone
// operations the point distance applies to fl(p - q) (o = 0 for the open metric and for the primary
// periodic image), so it equals the reference's point-to-box distance (kdtree.hpp:34-45).
// Periodic metric: min over the three images per axis == min over the 27 image shifts
// (o0, o1, o2) in {0, +L, -L}^3 of the shifted sum, so the tree is searched once per image whose
// root bound can still beat the current k-th distance.  Almost always that is only the primary
// image: the fast instantiation (IMAGES = false) searches just that one and, if a shifted image could
// still matter (the search ball reaches through a face of the box), hands the query to a work list
// that the general instantiation (IMAGES = true) finishes: it resumes from the row the fast pass left
// in the output (raw d2) and searches only the shifted images, or, when no rows are written (kNN-CDF),
// answers the query from scratch.  Leaf distances are always the TRUE periodic d2, so a point met
// again through another image carries the same key and is recognised as a duplicate.
constexpr int kLaneStack = 32; // one pending far child per level; depth <= log2(2^32 / 16)
constexpr uint32_t kNoNode = 0xFFFFFFFFu;

__device__ __forceinline__ float root_term(float q, float o, float L) {
    // points of the periodic root cell have p in [0, L]
    float a = __fadd_rn(__fsub_rn(0.0f, q), o), b = __fadd_rn(__fsub_rn(L, q), o);
    float v = fmaxf(fmaxf(a, -b), 0.0f);
    return __fmul_rn(v, v);
}

// smallest root bound of any shifted image = smallest single-axis shifted term (the other axes
// contribute >= 0)
__device__ __forceinline__ float min_shifted_root_term(float qx, float qy, float qz, float L) {
    return fminf(fminf(fminf(root_term(qx, L, L), root_term(qx, -L, L)),
                       fminf(root_term(qy, L, L), root_term(qy, -L, L))),
                 fminf(root_term(qz, L, L), root_term(qz, -L, L)));
}

#ifndef NBK_LANE_MIN_BLOCKS
#define NBK_LANE_MIN_BLOCKS 9
#endif

struct DeferList {
    uint32_t *slots; // sorted-order slots of deferred queries
    uint32_t *count; // [0] entries appended by the fast pass, [1] entries handed out by the general pass
};

// What one launch answers.  flags: NBK_QUERY_SQUARED -> rows hold d2 instead of sqrt(d2).
struct QueryBatch {
    const float *q_aos;    // (m_total, 3) queries as the caller gave them
    const uint32_t *order; // this launch's slots -> query ids
    uint64_t m;            // slots of this launch
    int k;
    int flags;
    float *out_d;          // (m_total, k) rows at the query's own position; null: kNN-CDF only
    uint32_t *out_i;
    unsigned long long *gheap; // HeapT<0> scratch
    uint32_t gcolumns;
};

// Fused kNN-CDF epilogue: instead of the (M,k) rows, histogram the distance to the
// j-th neighbour for every rank j with row_of_rank[j] >= 0 over `n_bins` bins with edges[0..n_bins]:
// counts[row_of_rank[j]][b] += 1 for edges[b] <= d < edges[b+1], the last bin closed, exactly
// np.histogram(dist[:, j], edges) of the rows the query would have written.
struct CdfArgs {
    const float *edges;          // null: write rows
    unsigned long long *counts;  // [rows][n_bins]
    const int *row_of_rank;      // [k]
    int n_bins;
};

__device__ __forceinline__ void cdf_accumulate(CdfArgs const &c, int row, uint32_t d2_bits, bool emit) {
    // all 32 lanes arrive here together; `emit` says whether this lane has a row
    uint32_t b = 0xFFFFFFFFu;
    if (emit) {
        const float d = __fsqrt_rn(__uint_as_float(d2_bits)); // postprocess, kdtree.cpp:154-156
        if (d >= __ldg(c.edges) && d <= __ldg(c.edges + c.n_bins)) {
            int lo = 0, hi = c.n_bins; // largest lo with edges[lo] <= d
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(c.edges + mid) <= d) lo = mid;
                else hi = mid;
            }
            b = (uint32_t)lo;
        }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    if (b != 0xFFFFFFFFu && (int)(threadIdx.x & 31) == __ffs(peers) - 1)
        atomicAdd(&c.counts[(uint64_t)row * c.n_bins + b], (unsigned long long)__popc(peers));
}

// One query per lane, all 32 lanes of the warp call this together (`valid` = this lane has a query).
template <typename Top, bool PERIODIC, bool IMAGES>
__device__ __forceinline__ void lane_query(QueryTree const &t, QueryBatch const &a, DeferList const &defer,
                                           CdfArgs const &cdf, Top &top, uint64_t slot, bool valid,
                                           StageBuf const &stage = StageBuf{nullptr, 0u}) {
    // (tiny batches are not ordered: order == null means slot == query id)
    const uint64_t safe_slot = valid ? slot : a.m - 1;
    const uint32_t qid = a.order ? a.order[safe_slot] : (uint32_t)safe_slot;
    const float qx = a.q_aos[3 * (uint64_t)qid], qy = a.q_aos[3 * (uint64_t)qid + 1],
                qz = a.q_aos[3 * (uint64_t)qid + 2];
    const float L = t.box;
    const float wrap_d2 = __fmul_rn(0.5f * L, 0.5f * L);
    const int k = a.k;
    const bool raw = (a.flags & NBK_QUERY_SQUARED) != 0;
    // rows mode: the fast pass leaves its (raw) row in the output and the general pass resumes from it
    const bool carry = a.out_d != nullptr;

    float4 stack[kLaneStack]; // (t0, t1, t2, node bits)
    int sp = 0;
    int img = 0; // 0 = primary image; 1..26 = shifted images (IMAGES only)
    float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f;
    float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
    if (PERIODIC) {
        t0 = root_term(qx, 0.0f, L);
        t1 = root_term(qy, 0.0f, L);
        t2 = root_term(qz, 0.0f, L);
    }
    uint32_t cur = 0; // root
    bool done = !valid;
    uint32_t lbeg = 0, lend = 0;
    uint32_t home = kNoNode; // first point of the leaf that was scanned up front
    constexpr bool kQueued = kQueueCap > 0 && !Top::kShared && !Top::kGlobal;
    CandQueue cq;
    if (kQueued) {
        extern __shared__ unsigned long long queue_smem[];
        cq.bind(queue_smem);
    }
    bool first_round = !IMAGES; // the home-leaf round of the fast pass

    if (IMAGES && carry) {
        if (valid) top.load_row(qid, k, a.out_d, a.out_i);
        else top.init(k);
        cur = kNoNode; // the primary image is done: the walk below goes straight to the shifted ones
    } else {
        top.init(k);
    }

    if (!IMAGES) {
        // Home leaf first: plain descent to the leaf holding the query (no bounds, nothing pushed);
        // the loop below scans it and only then starts the bounded traversal with a realistic k-th distance,
        // so that almost none of the ~log2(N/leaf) far children on the path is ever pushed.
        if (valid) {
            uint32_t nd = 0;
            while (true) {
                const int4 raw_node = __ldg(reinterpret_cast<const int4 *>(t.nodes) + nd);
                if (raw_node.x < 0) {
                    lbeg = (uint32_t)raw_node.z;
                    lend = (uint32_t)raw_node.w;
                    break;
                }
                const float qd = raw_node.x == 0 ? qx : (raw_node.x == 1 ? qy : qz);
                nd = __fsub_rn(__int_as_float(raw_node.y), qd) < 0.0f ? (uint32_t)raw_node.w : (uint32_t)raw_node.z;
            }
            home = lbeg; // the main loop scans [lbeg, lend) first, then starts at the root
        }
    }

    while (true) {
        // ---- walk until this lane has a leaf to scan (or is finished) -------------------------
        while (!done && lbeg == lend) {
            if (cur == kNoNode) {
                if (sp == 0) {
                    if (!IMAGES) {
                        done = true;
                        break;
                    }
                    // next image whose root bound can still beat the k-th distance
                    bool found = false;
                    while (!found && ++img < 27) {
                        const int s0 = img % 3, s1 = (img / 3) % 3, s2 = img / 9;
                        o0 = s0 == 0 ? 0.0f : (s0 == 1 ? L : -L);
                        o1 = s1 == 0 ? 0.0f : (s1 == 1 ? L : -L);
                        o2 = s2 == 0 ? 0.0f : (s2 == 1 ? L : -L);
                        t0 = root_term(qx, o0, L);
                        t1 = root_term(qy, o1, L);
                        t2 = root_term(qz, o2, L);
                        found = __fadd_rn(__fadd_rn(t0, t1), t2) <= top.worst();
                    }
                    if (!found) {
                        done = true;
                        break;
                    }
                    cur = 0;
                    continue;
                }
                const float4 e = stack[--sp];
                if (!(__fadd_rn(__fadd_rn(e.x, e.y), e.z) <= top.worst())) continue;
                t0 = e.x;
                t1 = e.y;
                t2 = e.z;
                cur = __float_as_uint(e.w);
            }
            const int4 raw_node = __ldg(reinterpret_cast<const int4 *>(t.nodes) + cur);
            const int dim = raw_node.x;
            if (dim < 0) {
                cur = kNoNode;
                if (!IMAGES && (uint32_t)raw_node.z == home) continue; // already scanned
                lbeg = (uint32_t)raw_node.z;
                lend = (uint32_t)raw_node.w;
                break;
            }
            const float split = __int_as_float(raw_node.y);
            const float qd = dim == 0 ? qx : (dim == 1 ? qy : qz);
            float w = __fsub_rn(split, qd);
            if (IMAGES) w = __fadd_rn(w, dim == 0 ? o0 : (dim == 1 ? o1 : o2));
            // w > 0: the (shifted) query lies left of the plane -> left child first
            const bool left_first = !(w < 0.0f);
            const uint32_t near = left_first ? (uint32_t)raw_node.z : (uint32_t)raw_node.w;
            const uint32_t far = left_first ? (uint32_t)raw_node.w : (uint32_t)raw_node.z;
            const float ft = __fmul_rn(w, w);
            const float f0 = dim == 0 ? ft : t0, f1 = dim == 1 ? ft : t1, f2 = dim == 2 ? ft : t2;
            // non-strict: an equal-distance point with a smaller index must still be found
            if (__fadd_rn(__fadd_rn(f0, f1), f2) <= top.worst())
                stack[sp++] = make_float4(f0, f1, f2, __uint_as_float(far));
            cur = near;
        }
        if (!__any_sync(0xffffffffu, lbeg != lend)) break;
        // ---- scan it ------------------------------------------------------------------------------
        constexpr bool kStaged = NBK_STAGE_HOME != 0 && !IMAGES && !Top::kShared && !Top::kGlobal && kQueueCap == 0;
        if (kStaged && first_round) {
            first_round = false;
            // do all queries of the warp live in one (small enough) leaf?
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            const int src = __ffs(vm) - 1;
            const uint32_t b0 = __shfl_sync(0xffffffffu, lbeg, src), e0 = __shfl_sync(0xffffffffu, lend, src);
            const bool one_leaf = __all_sync(0xffffffffu, !valid || (lbeg == b0 && lend == e0)) && e0 > b0 &&
                                  e0 - b0 <= kStagePoints;
            if (one_leaf) {
                if ((threadIdx.x & 31) == 0)
                    stage_load(stage, t.tiles + (uint64_t)(b0 >> 3) * kTileFloat4, (e0 - b0) * 16u);
                stage_wait(stage, 0u); // one use per kernel: phase 0
                if (valid) {
                    const float4 *sp_tile = stage.tiles;
                    for (uint32_t p = b0; p != e0; p += 8u, sp_tile += kTileFloat4) {
                        scan_half_tile<Top, PERIODIC, true>(sp_tile, qx, qy, qz, t.box, wrap_d2, false, k, top);
                        scan_half_tile<Top, PERIODIC, true>(sp_tile + 1, qx, qy, qz, t.box, wrap_d2, false, k, top);
                    }
                }
            } else {
                scan_leaf<Top, PERIODIC>(t, lbeg, lend, qx, qy, qz, wrap_d2, false, k, top);
            }
        } else if (kQueued) {
#if NBK_QUEUE_HOME
            scan_leaf_queued<Top, PERIODIC>(t, lbeg, lend, qx, qy, qz, wrap_d2, IMAGES && img > 0, first_round, k, top, cq);
#else
            // home-leaf round: direct insertion (its insertion sites are busy anyway); queue + drain afterwards
            if (first_round) scan_leaf<Top, PERIODIC>(t, lbeg, lend, qx, qy, qz, wrap_d2, false, k, top);
            else scan_leaf_queued<Top, PERIODIC>(t, lbeg, lend, qx, qy, qz, wrap_d2, IMAGES && img > 0, false, k, top, cq);
#endif
            first_round = false;
        } else {
            scan_leaf<Top, PERIODIC>(t, lbeg, lend, qx, qy, qz, wrap_d2, IMAGES && img > 0, k, top);
        }
        lbeg = lend = 0;
    }

    bool emit = valid, finished = true;
    if (PERIODIC && !IMAGES) {
        // queries whose search ball reaches through a face of the box: one list position per query, the
        // positions of a warp contiguous (its queries are neighbours, so are their shifted searches)
        const bool deferred = valid && min_shifted_root_term(qx, qy, qz, L) <= top.worst();
        const unsigned dmask = __ballot_sync(0xffffffffu, deferred);
        if (dmask) {
            const int lane = threadIdx.x & 31;
            uint32_t base = 0;
            if (lane == __ffs(dmask) - 1) base = atomicAdd(defer.count, (uint32_t)__popc(dmask));
            base = __shfl_sync(0xffffffffu, base, __ffs(dmask) - 1);
            if (deferred) {
                defer.slots[base + __popc(dmask & ((1u << lane) - 1u))] = (uint32_t)slot;
                finished = false;
            }
        }
    }
    if (cdf.edges) {
        // the whole warp walks the ranks together (match_any inside)
        emit = emit && finished;
        top.for_each_rank(k, [&](int j, uint32_t d2_bits) {
            const int row = __ldg(cdf.row_of_rank + j);
            if (row >= 0) cdf_accumulate(cdf, row, d2_bits, emit);
        });
        return;
    }
    if (emit) top.write_row(qid, k, raw || !finished, a.out_d, a.out_i);
}

template <typename Top, bool PERIODIC, bool IMAGES>
// resident CTAs per SM the register allocation aims for: 9 (56 registers) measured best for K = 8
// (8: 81.2, 9: 78.8, 10: 78.8 ms per 10^8 queries; 64- and 256-thread CTAs: 79.1 / 81.6); the short
// lists of K <= 4 fit 48 registers (57.1 -> 54.7 ms at k = 4)
__global__ void __launch_bounds__(kQueryThreads, (IMAGES || Top::kShared || Top::kGlobal) ? 1
                                                 : (Top::kSize <= 4 ? NBK_LANE_MIN_BLOCKS + 1 : NBK_LANE_MIN_BLOCKS))
knn_lane_kernel(QueryTree t, QueryBatch a, DeferList defer, CdfArgs cdf) {
    static_assert(PERIODIC || !IMAGES, "image shifts exist only for the periodic metric");
    extern __shared__ unsigned long long heap_smem[];
    Top top;
    top.bind(TopBind{heap_smem, a.gheap, a.gcolumns});
    if (IMAGES) {
        // second pass: a persistent grid drains the work list, 32 consecutive entries per warp and turn,
        // handed out by an atomic ticket so that the warps finish together whatever the list length
        const uint32_t n = defer.count[0];
        const uint32_t lane = threadIdx.x & 31u;
        while (true) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&defer.count[1], 32u);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= n) break;
            const uint32_t pos = base + lane;
            const bool valid = pos < n;
            lane_query<Top, PERIODIC, IMAGES>(t, a, defer, cdf, top, defer.slots[valid ? pos : base], valid);
        }
    } else {
        const uint64_t slot = (uint64_t)blockIdx.x * kQueryThreads + threadIdx.x;
        const bool valid = slot < a.m;
        if (!__any_sync(0xffffffffu, valid)) return;
        StageBuf stage{nullptr, 0u};
        if (NBK_STAGE_HOME != 0 && !Top::kShared && !Top::kGlobal && kQueueCap == 0) stage_init(stage, heap_smem);
        lane_query<Top, PERIODIC, IMAGES>(t, a, defer, cdf, top, slot, valid, stage);
    }
}

// ---- flat block scan: the leaf scan + top-k in isolation (test hook) --------------------------------
// The twin of the reference's leaf kernel called on one flat block (wenda_insert_closest_l2[_periodic]_avx2,
// kdtree_opt_asm.hpp:12-19; pinned by tests/test_asm.cpp:97-199 and tests/test_inserters.cpp:159-220):
// every thread scans ALL n points (tiles [0, n)) for its query with the production scan_leaf + container
// and writes the row.  No tree, no traversal.
template <typename Top, bool PERIODIC>
__global__ void __launch_bounds__(kQueryThreads)
scan_block_kernel(QueryTree t, uint32_t n, QueryBatch a) {
    extern __shared__ unsigned long long heap_smem[];
    const uint64_t slot = (uint64_t)blockIdx.x * kQueryThreads + threadIdx.x;
    if (slot >= a.m) return;
    Top top;
    top.bind(TopBind{heap_smem, a.gheap, a.gcolumns});
    top.init(a.k);
    const float qx = a.q_aos[3 * slot], qy = a.q_aos[3 * slot + 1], qz = a.q_aos[3 * slot + 2];
    const float wrap_d2 = __fmul_rn(0.5f * t.box, 0.5f * t.box);
    scan_leaf<Top, PERIODIC>(t, 0u, n, qx, qy, qz, wrap_d2, false, a.k, top);
    top.write_row((uint32_t)slot, a.k, (a.flags & NBK_QUERY_SQUARED) != 0, a.out_d, a.out_i);
}

// ---- the packet kernel -----------------------------------------------------------------------------
constexpr int kMaxStack = 48; // pushes two, pops one per internal node: depth + 1 entries

struct __align__(16) StackEntry {
    float lo0, lo1, lo2;
    uint32_t node;
    float hi0, hi1, hi2;
    uint32_t pad;
};

// open metric: identical to L2Distance::box_distance (kdtree.hpp:34-45)
__device__ __forceinline__ float axis_lb_open(float lo, float hi, float q) {
    float dl = fmaxf(__fsub_rn(lo, q), 0.0f);
    float dr = fmaxf(__fsub_rn(q, hi), 0.0f);
    return __fadd_rn(__fmul_rn(dl, dl), __fmul_rn(dr, dr));
}

// periodic metric: lower bound of min(d^2, (d+L)^2, (d-L)^2) over d = fl(p - q) in [a, b]
// (a = fl(lo - q), b = fl(hi - q)): each image interval is the same monotone operation applied to
// the end points, and the distance of an interval [u, v] from zero is max(u, -v, 0).
__device__ __forceinline__ float axis_lb_periodic(float a, float b, float L) {
    float v0 = fmaxf(fmaxf(a, -b), 0.0f);
    float ap = __fadd_rn(a, L), bp = __fadd_rn(b, L);
    float vp = fmaxf(fmaxf(ap, -bp), 0.0f);
    float am = __fsub_rn(a, L), bm = __fsub_rn(b, L);
    float vm = fmaxf(fmaxf(am, -bm), 0.0f);
    float v = fminf(v0, fminf(vp, vm));
    return __fmul_rn(v, v);
}

template <int K, bool PERIODIC>
__global__ void __launch_bounds__(kQueryThreads)
knn_packet_kernel(QueryTree t, const float *__restrict__ q_aos, const uint32_t *__restrict__ order,
                  uint64_t m, int k_out, bool raw, float *__restrict__ out_d, uint32_t *__restrict__ out_i) {
    __shared__ StackEntry stack[kQueryWarps][kMaxStack];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t slot = (uint64_t)blockIdx.x * kQueryThreads + threadIdx.x;
    const bool valid = slot < m;
    if (!__any_sync(0xffffffffu, valid)) return;
    const uint32_t qid = order[valid ? slot : m - 1];
    const float qx = q_aos[3 * (uint64_t)qid], qy = q_aos[3 * (uint64_t)qid + 1],
                qz = q_aos[3 * (uint64_t)qid + 2];
    const float wrap_d2 = __fmul_rn(0.5f * t.box, 0.5f * t.box);

    TopK<K> top;
    top.init(k_out);

    StackEntry *st = stack[warp];
    if (lane == 0) st[0] = StackEntry{t.lo[0], t.lo[1], t.lo[2], 0u, t.hi[0], t.hi[1], t.hi[2], 0u};
    __syncwarp();
    int sp = 1;
    while (sp > 0) {
        --sp;
        const StackEntry e = st[sp];
        __syncwarp(); // everyone has read the entry before lane 0 overwrites the slot
        float lb;
        if (PERIODIC) {
            // d = fl(p - q) of every point of the cell lies in [a, b] on each axis
            lb = __fadd_rn(__fadd_rn(axis_lb_periodic(__fsub_rn(e.lo0, qx), __fsub_rn(e.hi0, qx), t.box),
                                     axis_lb_periodic(__fsub_rn(e.lo1, qy), __fsub_rn(e.hi1, qy), t.box)),
                           axis_lb_periodic(__fsub_rn(e.lo2, qz), __fsub_rn(e.hi2, qz), t.box));
        } else {
            lb = __fadd_rn(__fadd_rn(axis_lb_open(e.lo0, e.hi0, qx), axis_lb_open(e.lo1, e.hi1, qy)),
                           axis_lb_open(e.lo2, e.hi2, qz));
        }
        // non-strict: an equal-distance point with a smaller index must still be found
        const bool need = valid && lb <= top.worst();
        const unsigned need_mask = __ballot_sync(0xffffffffu, need);
        if (need_mask == 0u) continue;
        const nbk_node nd = *reinterpret_cast<const nbk_node *>(
            &reinterpret_cast<const int4 *>(t.nodes)[e.node]);
        if (nd.dim < 0) {
            scan_leaf<TopK<K>, PERIODIC>(t, nd.left, nd.right, qx, qy, qz, wrap_d2, false, k_out, top);
            continue;
        }
        const float qd = nd.dim == 0 ? qx : (nd.dim == 1 ? qy : qz);
        const unsigned right_mask = __ballot_sync(0xffffffffu, need && qd > nd.split);
        const bool right_first = 2 * __popc(right_mask) > __popc(need_mask);
        if (lane == 0) {
            StackEntry l = e, r = e;
            l.node = nd.left;
            r.node = nd.right;
            if (nd.dim == 0) { l.hi0 = nd.split; r.lo0 = nd.split; }
            else if (nd.dim == 1) { l.hi1 = nd.split; r.lo1 = nd.split; }
            else { l.hi2 = nd.split; r.lo2 = nd.split; }
            st[sp] = right_first ? l : r;     // visited second
            st[sp + 1] = right_first ? r : l; // visited first
        }
        __syncwarp();
        sp += 2;
    }
    if (valid) top.write_row(qid, k_out, raw, out_d, out_i);
}

// ---- KDTreeQueryStatistics: the reference's own traversal, one thread per query -----------------
// kdtree_impl.hpp:226-268 with the reference's box_distance (kdtree.hpp:34-45,88-107) and strict-<
// insertion; counters summed over the batch.  Debug/observability path, not the hot path.
template <bool PERIODIC>
__device__ __forceinline__ float ref_box_distance(const float q[3], const float b[6], float L) {
    float r = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (!PERIODIC) {
            float dl = fmaxf(__fsub_rn(b[2 * i], q[i]), 0.0f);
            float dr = fmaxf(__fsub_rn(q[i], b[2 * i + 1]), 0.0f);
            r = __fadd_rn(r, __fadd_rn(__fmul_rn(dl, dl), __fmul_rn(dr, dr)));
        } else if (q[i] < b[2 * i]) {
            float d = __fsub_rn(b[2 * i], q[i]);
            float dw = __fsub_rn(__fadd_rn(q[i], L), b[2 * i + 1]);
            float mn = fminf(d, dw);
            r = __fadd_rn(r, __fmul_rn(mn, mn));
        } else if (q[i] > b[2 * i + 1]) {
            float d = __fsub_rn(q[i], b[2 * i + 1]);
            float dw = __fsub_rn(__fadd_rn(b[2 * i], L), q[i]);
            float mn = fminf(d, dw);
            r = __fadd_rn(r, __fmul_rn(mn, mn));
        }
    }
    return r;
}

constexpr int kStatsMaxK = 64;

template <bool PERIODIC>
__global__ void __launch_bounds__(128)
stats_kernel(QueryTree t, const float *__restrict__ q_aos, uint64_t m, int k, float *__restrict__ gbest,
             unsigned long long *__restrict__ out3) {
    uint64_t i = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    unsigned long long nv = 0, np = 0, pv = 0;
    if (i < m) {
        const float q[3] = {q_aos[3 * i], q_aos[3 * i + 1], q_aos[3 * i + 2]};
        float local_best[kStatsMaxK]; // unsorted, track the maximum like a replace-top queue
        float *best = k <= kStatsMaxK ? local_best : gbest + i * (uint64_t)k; // k > 64: global scratch row
        for (int j = 0; j < k; ++j) best[j] = FLT_MAX;
        int top = 0;
        // explicit stack of (node, bounds, stage): stage 0 = entering, 1 = closer child done
        struct Frame { uint32_t node; float b[6]; int stage; };
        Frame stack[40];
        int sp = 0;
        stack[0].node = 0;
        for (int d = 0; d < 3; ++d) { stack[0].b[2 * d] = t.lo[d]; stack[0].b[2 * d + 1] = t.hi[d]; }
        stack[0].stage = 0;
        sp = 1;
        while (sp > 0) {
            Frame &f = stack[sp - 1];
            const nbk_node nd = t.nodes[f.node];
            if (f.stage == 0) {
                nv += 1;
                if (nd.dim < 0) {
                    for (uint32_t p = nd.left; p < nd.right; ++p) {
                        const float px = tile_coord(t.tiles, p, 0), py = tile_coord(t.tiles, p, 1),
                                    pz = tile_coord(t.tiles, p, 2);
                        float d = PERIODIC ? d2_periodic(px, py, pz, q[0], q[1], q[2], t.box)
                                           : d2_open(px, py, pz, q[0], q[1], q[2]);
                        if (d < best[top]) {
                            best[top] = d;
                            top = 0;
                            for (int j = 1; j < k; ++j) if (best[top] < best[j]) top = j;
                        }
                    }
                    pv += nd.right - nd.left;
                    --sp;
                    continue;
                }
                f.stage = 1;
                const bool right_close = q[nd.dim] > nd.split;
                Frame c;
                c.node = right_close ? nd.right : nd.left;
                for (int d = 0; d < 6; ++d) c.b[d] = f.b[d];
                c.b[right_close ? 2 * nd.dim : 2 * nd.dim + 1] = nd.split;
                c.stage = 0;
                if (ref_box_distance<PERIODIC>(q, c.b, t.box) < best[top]) stack[sp++] = c;
                else np += 1;
            } else {
                const bool right_close = q[nd.dim] > nd.split;
                Frame c;
                c.node = right_close ? nd.left : nd.right;
                for (int d = 0; d < 6; ++d) c.b[d] = f.b[d];
                c.b[right_close ? 2 * nd.dim + 1 : 2 * nd.dim] = nd.split;
                c.stage = 0;
                --sp; // this frame is finished either way
                if (best[top] < ref_box_distance<PERIODIC>(q, c.b, t.box)) np += 1;
                else stack[sp++] = c;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        nv += __shfl_xor_sync(0xffffffffu, nv, o);
        np += __shfl_xor_sync(0xffffffffu, np, o);
        pv += __shfl_xor_sync(0xffffffffu, pv, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out3[0], nv);
        atomicAdd(&out3[1], np);
        atomicAdd(&out3[2], pv);
    }
}

// ---- tile -> SoA (backs nbk_tree_copy_points) -------------------------------------------------------
__global__ void __launch_bounds__(256)
untile_kernel(const float4 *__restrict__ tiles, uint64_t n, float *__restrict__ x,
              float *__restrict__ y, float *__restrict__ z, uint32_t *__restrict__ idx) {
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float *f = reinterpret_cast<const float *>(tiles) + (i >> 3) * 32 + (i & 7);
    x[i] = f[0];
    y[i] = f[8];
    z[i] = f[16];
    idx[i] = __float_as_uint(f[24]);
}

} // namespace nbk
