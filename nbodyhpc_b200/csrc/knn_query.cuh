// Batched kNN query kernels (sm_100a).
//
// Replaces PyKDTree::query -> KDTree::find_closest -> KDTreeQuery::compute -> asm leaf scan
// (pybind.cpp:90-189, kdtree.cpp:133-159, kdtree_impl.hpp:212-268, kdtree_asm_systemv.asm:121-248).
//
// Design: queries are ordered along a Morton curve, so 32 consecutive queries are
// spatial neighbours.  One warp walks the tree ONCE for its 32 queries (one query per lane): the
// node stack is warp-uniform and lives in shared memory, a subtree is entered when ANY lane's
// exact lower bound to its cell does not exceed that lane's current k-th distance, and every leaf
// point is fetched once per warp with warp-uniform 16-byte loads and evaluated by all 32 lanes.
// The top-k of each lane is a register-resident sorted list of 64-bit (d2 bits, index) keys, so the
// result is the exact top-k under the total order (d2, index) no matter in which order nodes are
// visited -- which is what makes the answer independent of the traversal and equal to the
// reference's (whose own result is traversal dependent only for exact d2 ties).
//
// Arithmetic contract: d2 is computed with __fsub_rn/__fmul_rn/__fadd_rn (never contracted into
// FMA) in the reference's order ((dx2 + dy2) + dz2) with d = p - q (kdtree_asm_systemv.asm:76-87),
// periodic axis term min(d^2, (d+L)^2, (d-L)^2) (kdtree_asm_systemv.asm:89-119).
#pragma once

#include "common.cuh"

namespace nbk {

struct QueryTree {
    const nbk_node *nodes;
    const float *x, *y, *z;
    const uint32_t *idx;
    bool periodic;
    float box;      // periodic box size
    float lo[3], hi[3]; // root cell: periodic [0, box]; open [-FLT_MAX, FLT_MAX]
};

constexpr int kQueryThreads = 128;
constexpr int kQueryWarps = kQueryThreads / 32;
constexpr int kMaxStack = 48; // pushes two, pops one per internal node: depth + 1 entries

struct __align__(16) StackEntry {
    float lo0, lo1, lo2;
    uint32_t node;
    float hi0, hi1, hi2;
    uint32_t pad;
};

// ---- Morton ordering of the queries --------------------------------------------------------------
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void __launch_bounds__(256)
morton_keys_kernel(const float *__restrict__ q_aos, uint64_t m, float lo0, float lo1, float lo2,
                   float s0, float s1, float s2, uint32_t *__restrict__ keys,
                   uint32_t *__restrict__ vals) {
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    float qx = q_aos[3 * i], qy = q_aos[3 * i + 1], qz = q_aos[3 * i + 2];
    int cx = min(max((int)((qx - lo0) * s0), 0), 1023);
    int cy = min(max((int)((qy - lo1) * s1), 0), 1023);
    int cz = min(max((int)((qz - lo2) * s2), 0), 1023);
    keys[i] = spread10((uint32_t)cx) | (spread10((uint32_t)cy) << 1) | (spread10((uint32_t)cz) << 2);
    vals[i] = (uint32_t)i;
}

// ---- exact lower bounds from a query to an axis-aligned cell ------------------------------------
// Both bounds never exceed the d2 this file computes for any point inside the cell, because every
// step is the same monotone float operation applied to the end points of the interval the point's
// coordinate lies in ("pruning is exact").

// open metric: identical to L2Distance::box_distance (kdtree.hpp:34-45)
__device__ __forceinline__ float axis_lb_open(float lo, float hi, float q) {
    float dl = fmaxf(__fsub_rn(lo, q), 0.0f);
    float dr = fmaxf(__fsub_rn(q, hi), 0.0f);
    return __fadd_rn(__fmul_rn(dl, dl), __fmul_rn(dr, dr));
}

// periodic metric: lower bound of min(d^2, (d+L)^2, (d-L)^2) over d = fl(p - q), p in [lo, hi].
// `wrap` is set when some point of the cell may need a wrapped image for this query.
__device__ __forceinline__ float axis_lb_periodic(float lo, float hi, float q, float L, float halfL,
                                                  bool &wrap) {
    float a = __fsub_rn(lo, q), b = __fsub_rn(hi, q); // d in [a, b]
    wrap = wrap || (a < -halfL) || (b > halfL);
    float v0 = fmaxf(fmaxf(a, -b), 0.0f);
    float ap = __fadd_rn(a, L), bp = __fadd_rn(b, L);
    float vp = fmaxf(fmaxf(ap, -bp), 0.0f);
    float am = __fsub_rn(a, L), bm = __fsub_rn(b, L);
    float vm = fmaxf(fmaxf(am, -bm), 0.0f);
    float v = fminf(v0, fminf(vp, vm));
    return __fmul_rn(v, v);
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

// ---- register-resident top-k ----------------------------------------------------------------------
// keys sorted ascending; key = (d2 bits << 32) | index.  d2 >= +0 so the bit pattern orders like the
// float.  Empty slots hold (FLT_MAX, 0): a candidate at d2 == FLT_MAX never replaces one
// (the reference inserts only if d2 < FLT_MAX, kdtree_impl.hpp:210 + strict <).
template <int K> struct TopK {
    unsigned long long key[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < K; ++j) key[j] = (unsigned long long)kFltMaxBits << 32;
    }
    __device__ __forceinline__ float worst() const {
        return __uint_as_float((uint32_t)(key[K - 1] >> 32));
    }
    __device__ __forceinline__ void insert(unsigned long long cand) {
        // precondition: cand < key[K-1]
        key[K - 1] = cand;
#pragma unroll
        for (int j = K - 1; j > 0; --j) {
            unsigned long long a = key[j - 1], b = key[j];
            bool sw = b < a;
            key[j - 1] = sw ? b : a;
            key[j] = sw ? a : b;
        }
    }
};

// ---- the packet kernel ------------------------------------------------------------------------------
template <int K, bool PERIODIC>
__device__ __forceinline__ void scan_leaf(QueryTree const &t, uint32_t begin, uint32_t end, float qx,
                                          float qy, float qz, bool wrap, TopK<K> &top) {
    // leaves start on multiples of 8 points and hold a multiple of 8: two float4 per column/step
    for (uint32_t p = begin; p < end; p += 4) {
        const float4 X = __ldg(reinterpret_cast<const float4 *>(t.x + p));
        const float4 Y = __ldg(reinterpret_cast<const float4 *>(t.y + p));
        const float4 Z = __ldg(reinterpret_cast<const float4 *>(t.z + p));
        float d[4];
        if (PERIODIC && wrap) {
            d[0] = d2_periodic(X.x, Y.x, Z.x, qx, qy, qz, t.box);
            d[1] = d2_periodic(X.y, Y.y, Z.y, qx, qy, qz, t.box);
            d[2] = d2_periodic(X.z, Y.z, Z.z, qx, qy, qz, t.box);
            d[3] = d2_periodic(X.w, Y.w, Z.w, qx, qy, qz, t.box);
        } else {
            d[0] = d2_open(X.x, Y.x, Z.x, qx, qy, qz);
            d[1] = d2_open(X.y, Y.y, Z.y, qx, qy, qz);
            d[2] = d2_open(X.z, Y.z, Z.z, qx, qy, qz);
            d[3] = d2_open(X.w, Y.w, Z.w, qx, qy, qz);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (d[j] <= top.worst()) {
                unsigned long long cand =
                    ((unsigned long long)__float_as_uint(d[j]) << 32) | __ldg(t.idx + p + j);
                if (cand < top.key[K - 1]) top.insert(cand);
            }
        }
    }
}

template <int K, bool PERIODIC>
__global__ void __launch_bounds__(kQueryThreads)
knn_packet_kernel(QueryTree t, const float *__restrict__ q_aos, const uint32_t *__restrict__ order,
                  uint64_t m, int k_out, float *__restrict__ out_d, uint32_t *__restrict__ out_i) {
    __shared__ StackEntry stack[kQueryWarps][kMaxStack];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t slot = (uint64_t)blockIdx.x * kQueryThreads + threadIdx.x;
    const bool valid = slot < m;
    if (!__any_sync(0xffffffffu, valid)) return;
    const uint32_t qid = order[valid ? slot : m - 1];
    const float qx = q_aos[3 * (uint64_t)qid], qy = q_aos[3 * (uint64_t)qid + 1],
                qz = q_aos[3 * (uint64_t)qid + 2];
    const float halfL = 0.5f * t.box;

    TopK<K> top;
    top.init();

    StackEntry *st = stack[warp];
    if (lane == 0) st[0] = StackEntry{t.lo[0], t.lo[1], t.lo[2], 0u, t.hi[0], t.hi[1], t.hi[2], 0u};
    __syncwarp();
    int sp = 1;
    while (sp > 0) {
        --sp;
        const StackEntry e = st[sp];
        __syncwarp(); // everyone has read the entry before lane 0 overwrites the slot
        bool wrap = false;
        float lb;
        if (PERIODIC) {
            lb = __fadd_rn(__fadd_rn(axis_lb_periodic(e.lo0, e.hi0, qx, t.box, halfL, wrap),
                                     axis_lb_periodic(e.lo1, e.hi1, qy, t.box, halfL, wrap)),
                           axis_lb_periodic(e.lo2, e.hi2, qz, t.box, halfL, wrap));
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
            // lanes that do not need the leaf cannot be hurt by the cheaper open formula: their
            // true d2 >= lb > worst, and the open d2 is never below the periodic one
            const bool any_wrap = PERIODIC && (__ballot_sync(0xffffffffu, need && wrap) != 0u);
            scan_leaf<K, PERIODIC>(t, nd.left, nd.right, qx, qy, qz, any_wrap, top);
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

    if (valid) {
        float *od = out_d + (uint64_t)qid * k_out;
        uint32_t *oi = out_i + (uint64_t)qid * k_out;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (j < k_out) {
                uint32_t bits = (uint32_t)(top.key[j] >> 32);
                od[j] = __fsqrt_rn(__uint_as_float(bits)); // postprocess, kdtree.cpp:154-156
                oi[j] = bits == kFltMaxBits ? 0xFFFFFFFFu : (uint32_t)top.key[j];
            }
        }
    }
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
stats_kernel(QueryTree t, const float *__restrict__ q_aos, uint64_t m, int k,
             unsigned long long *__restrict__ out3) {
    uint64_t i = (uint64_t)blockIdx.x * 128 + threadIdx.x;
    unsigned long long nv = 0, np = 0, pv = 0;
    if (i < m) {
        const float q[3] = {q_aos[3 * i], q_aos[3 * i + 1], q_aos[3 * i + 2]};
        float best[kStatsMaxK]; // unsorted, track the maximum like a replace-top queue
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
                        float d = PERIODIC ? d2_periodic(t.x[p], t.y[p], t.z[p], q[0], q[1], q[2], t.box)
                                           : d2_open(t.x[p], t.y[p], t.z[p], q[0], q[1], q[2]);
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
    // block reduction
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

} // namespace nbk
