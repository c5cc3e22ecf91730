// kd-tree construction, select-and-partition form (default build path).
//
// Replaces the reference's recursive builder + Floyd-Rivest/AVX2 selection
// (kdtree_impl.hpp:94-157, kdtree_selection.cpp:34-200,322-368,475-494).  The reference selects the
// element of rank median_offset along one axis and partitions the segment around it, level by
// level; so does this code, for ALL segments of a level at once:
//
//   top phase (segments larger than one CTA's shared memory), per level, all in HBM:
//     hist      per-segment histogram of the split coordinate over range-adaptive bins
//               (fused into the previous level's partition pass from level 1 on)
//     bucket    the bin holding the rank-median element, rank inside it, zero the histogram
//     compact   the elements of that bin -> candidate list of (orderable coordinate << 32 | id)
//     select    exact rank selection among the candidates -> pivot, node record, child bounds
//     partition one streaming pass: (coordinate, id) < pivot goes left, the rest right
//   bottom phase: one CTA per remaining segment (<= 8192 points) finishes its whole sub-tree in
//   shared memory (tree_bottom.cuh) and writes the final 128-byte point tiles and node records.
//
// The total order is (coordinate as orderable uint32, id) with id = position in the caller's
// array, so the left/right SETS are unique whatever order the passes write elements in, and the
// split is the coordinate of the rank-median element exactly as the reference's selection leaves
// it (kdtree_impl.hpp:108-125).  Topology (node numbering, leaf ranges) is a function of counts
// only and is planned on the host for the top levels and from a small lookup table below them.
#pragma once

#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace nbk {
namespace td {

constexpr uint32_t kNoSplit = 0xFFFFFFFFu;
constexpr int kChunk = 2048; // points handled by one CTA of the streaming passes
constexpr int kChunkThreads = 256;
constexpr int kChunkItems = kChunk / kChunkThreads;
constexpr int kChunkWarps = kChunkThreads / 32;
#ifndef NBK_BOTTOM_CAP
#define NBK_BOTTOM_CAP 8192
#endif
constexpr int kBottomCap = NBK_BOTTOM_CAP; // largest segment the bottom kernel splits in shared memory
constexpr int kSelectCap = 2048; // candidates sorted in shared memory by the select kernel

// One segment of one level.  Level l has 2^l slots in heap order (children of s: 2s, 2s+1); a
// leaf met above the bottom phase is carried down as child 2s (mid == kNoSplit), 2s+1 stays empty.
struct alignas(16) Seg {
    uint32_t begin, count;
    uint32_t mid;        // median offset ((count/2)/block)*block, kdtree_impl.hpp:108-110
    uint32_t node;       // pre-order index of this segment's node
    uint32_t right_node; // pre-order index of the right child (left child = node + 1)
    uint32_t pad[3];
};

struct alignas(16) Sel {
    uint32_t bin; // histogram bin holding the rank-`mid` element; kNoSplit if the segment is not split
    uint32_t r;   // rank of that element among the bin's elements
    uint32_t c;   // elements in the bin
    uint32_t pad;
};

// ---- host plan ------------------------------------------------------------------------------------
struct CountRules {
    uint32_t leaf, block; // leaf = max(leaf_size, 2*block) (kdtree_impl.hpp:91)
    std::unordered_map<uint32_t, uint32_t> nodes_memo, depth_memo;
    uint32_t median(uint32_t c) const { return (c / 2 / block) * block; }
    bool splits(uint32_t c) const { return c > leaf; } // kdtree_impl.hpp:101
    uint32_t nodes_of(uint32_t c) { // nodes of the sub-tree over c points
        if (!splits(c)) return 1;
        auto it = nodes_memo.find(c);
        if (it != nodes_memo.end()) return it->second;
        uint32_t m = median(c);
        uint32_t v = 1 + nodes_of(m) + nodes_of(c - m);
        nodes_memo.emplace(c, v);
        return v;
    }
    uint32_t depth_of(uint32_t c) { // split levels of the sub-tree over c points
        if (!splits(c)) return 0;
        auto it = depth_memo.find(c);
        if (it != depth_memo.end()) return it->second;
        uint32_t m = median(c);
        uint32_t v = 1 + std::max(depth_of(m), depth_of(c - m));
        depth_memo.emplace(c, v);
        return v;
    }
};

struct TopPlan {
    int top_levels = 0;            // levels [0, top_levels) are split by the streaming passes
    std::vector<Seg> segs;         // heap array: level l at [(1<<l)-1, (1<<(l+1))-1), l <= top_levels
    std::vector<uint32_t> biggest;  // per level: points of the largest segment
    std::vector<uint32_t> lut;     // nodes_of(8*i), i <= kBottomCap/8
    uint64_t n_nodes = 0;
    int n_levels = 0;
};

inline TopPlan plan_top(uint64_t n_padded, int leaf_size, int block_size) {
    TopPlan plan;
    CountRules rules;
    rules.block = (uint32_t)block_size;
    rules.leaf = (uint32_t)std::max(leaf_size, 2 * block_size);
    const uint32_t n = (uint32_t)n_padded;
    plan.n_nodes = rules.nodes_of(n);
    plan.n_levels = (int)rules.depth_of(n);
    plan.lut.resize(kBottomCap / 8 + 1);
    for (uint32_t i = 0; i < plan.lut.size(); ++i) plan.lut[i] = rules.nodes_of(8 * i);

    auto fill = [&](Seg &s) {
        s.mid = rules.splits(s.count) ? rules.median(s.count) : kNoSplit;
        s.right_node = s.mid != kNoSplit ? s.node + 1 + rules.nodes_of(s.mid) : 0u;
    };
    Seg root{};
    root.begin = 0;
    root.count = n;
    root.node = 0;
    fill(root);
    plan.segs.push_back(root);
    for (int l = 0;; ++l) {
        const size_t base = ((size_t)1 << l) - 1, cnt = (size_t)1 << l;
        bool any_big = false;
        uint32_t biggest = 0;
        for (size_t s = 0; s < cnt; ++s) {
            Seg const &sg = plan.segs[base + s];
            biggest = std::max(biggest, sg.count);
            any_big = any_big || (sg.mid != kNoSplit && sg.count > (uint32_t)kBottomCap);
        }
        plan.biggest.push_back(std::max(biggest, 1u));
        if (!any_big) {
            plan.top_levels = l;
            break;
        }
        plan.segs.resize(base + cnt + 2 * cnt);
        for (size_t s = 0; s < cnt; ++s) {
            Seg const sg = plan.segs[base + s];
            Seg a{}, b{};
            if (sg.mid != kNoSplit) {
                a.begin = sg.begin;
                a.count = sg.mid;
                a.node = sg.node + 1;
                b.begin = sg.begin + sg.mid;
                b.count = sg.count - sg.mid;
                b.node = sg.right_node;
            } else {
                a.begin = sg.begin;
                a.count = sg.count;
                a.node = sg.node;
                b.begin = sg.begin + sg.count;
            }
            if (a.count) fill(a); else a.mid = kNoSplit;
            if (b.count) fill(b); else b.mid = kNoSplit;
            plan.segs[base + cnt + 2 * s] = a;
            plan.segs[base + cnt + 2 * s + 1] = b;
        }
    }
    return plan;
}

// ---- monotone binning -----------------------------------------------------------------------------
// bin_of is non-decreasing in the orderable key of v (NaNs sit at the two ends like their keys), which
// is all the selection needs: every element of a lower bin orders before every element of a higher
// one.  lo/hi are the segment's cell bounds along the axis, so uniform-ish data spreads evenly.
__device__ __forceinline__ float bin_scale(float lo, float hi, uint32_t nb) {
    const float ext = hi - lo;
    return (ext > 0.0f && ext <= FLT_MAX) ? (float)nb / ext : 0.0f;
}
__device__ __forceinline__ uint32_t bin_of(float v, float lo, float scale, uint32_t nb) {
    if (v != v) return (__float_as_uint(v) >> 31) ? 0u : nb - 1u;
    const float t = (v - lo) * scale;
    const int b = max(__float2int_rz(t), 0); // NaN -> 0, saturating
    return min((uint32_t)b, nb - 1u);
}
__device__ __forceinline__ unsigned long long composite(float v, uint32_t id) {
    return ((unsigned long long)float_to_ordered(__float_as_uint(v)) << 32) | id;
}

// bounds: two float4 per segment {lo0, lo1, lo2, -} {hi0, hi1, hi2, -}
__global__ void init_bounds_kernel(const uint32_t *__restrict__ bounds6, float4 *__restrict__ bounds) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float v[6];
    for (int i = 0; i < 6; ++i) v[i] = __uint_as_float(ordered_to_float(bounds6[i]));
    bounds[0] = make_float4(v[0], v[1], v[2], 0.0f);
    bounds[1] = make_float4(v[3], v[4], v[5], 0.0f);
}

__device__ __forceinline__ void seg_range(const float4 *bounds, uint32_t s, int dim, float &lo, float &hi) {
    const float *b = reinterpret_cast<const float *>(bounds + 2 * (uint64_t)s);
    lo = b[dim];
    hi = b[4 + dim];
}

// ---- hist (stand-alone form, level 0) ----------------------------------------------------------------
__global__ void __launch_bounds__(kChunkThreads)
hist_kernel(const float *__restrict__ coord, const Seg *__restrict__ segs,
            const float4 *__restrict__ bounds, int dim, uint32_t max_chunks, uint32_t nb,
            uint32_t *__restrict__ hist) {
    extern __shared__ uint32_t sh_hist[];
    const uint32_t s = blockIdx.x / max_chunks, c = blockIdx.x - s * max_chunks;
    const Seg sg = segs[s];
    const uint32_t off = c * kChunk;
    if (sg.mid == kNoSplit || off >= sg.count) return;
    const uint32_t end = min(off + (uint32_t)kChunk, sg.count);
    for (uint32_t b = threadIdx.x; b < nb; b += kChunkThreads) sh_hist[b] = 0;
    __syncthreads();
    float lo, hi;
    seg_range(bounds, s, dim, lo, hi);
    const float scale = bin_scale(lo, hi, nb);
#pragma unroll
    for (int r = 0; r < kChunkItems; ++r) {
        const uint32_t i = off + r * kChunkThreads + threadIdx.x;
        if (i < end) atomicAdd(&sh_hist[bin_of(coord[(uint64_t)sg.begin + i], lo, scale, nb)], 1u);
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nb; b += kChunkThreads) {
        const uint32_t v = sh_hist[b];
        if (v) atomicAdd(&hist[(uint64_t)s * nb + b], v);
    }
}

// ---- bucket: the bin holding the rank-`mid` element (one warp per segment) ----------------------------
// Reads and clears the segment's histogram row and its three cursors.
__device__ __forceinline__ Sel find_bucket(const Seg &sg, uint32_t s, uint32_t nb, uint32_t *__restrict__ hist,
                                           uint32_t *__restrict__ cursors, int lane) {
    Sel out{kNoSplit, 0u, 0u, 0u};
    if (lane < 3) cursors[3 * (uint64_t)s + lane] = 0;
    if (sg.mid == kNoSplit) return out;
    uint32_t running = 0;
    bool found = false;
    for (uint32_t base = 0; base < nb; base += 32) {
        uint32_t *hp = hist + (uint64_t)s * nb + base + lane;
        const uint32_t h = *hp;
        *hp = 0;
        uint32_t incl = h;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, running + incl > sg.mid);
        if (!found && hit) {
            const int L = __ffs(hit) - 1;
            const uint32_t hL = __shfl_sync(0xffffffffu, h, L);
            const uint32_t iL = __shfl_sync(0xffffffffu, incl, L);
            out.bin = base + L;
            out.r = sg.mid - (running + iL - hL);
            out.c = hL;
            found = true;
        }
        running += __shfl_sync(0xffffffffu, incl, 31);
    }
    return out;
}

__global__ void __launch_bounds__(256)
bucket_kernel(const Seg *__restrict__ segs, uint32_t nseg, uint32_t nb, uint32_t *__restrict__ hist,
              Sel *__restrict__ sel, uint32_t *__restrict__ cursors) {
    const uint32_t s = (blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= nseg) return;
    const Sel out = find_bucket(segs[s], s, nb, hist, cursors, lane);
    if (lane == 0) sel[s] = out;
}

// exclusive scan of kChunkItems*kChunkWarps smem counters by one warp; returns the total
constexpr int kChunkCounters = kChunkItems * kChunkWarps;
static_assert(kChunkCounters % 32 == 0, "counters are scanned 32 lanes at a time");
__device__ __forceinline__ uint32_t warp_scan_counters(uint32_t *cnt, int lane) {
    constexpr int PER = kChunkCounters / 32;
    uint32_t a[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        a[j] = cnt[lane * PER + j];
        sum += a[j];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    uint32_t run = incl - sum;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        cnt[lane * PER + j] = run;
        run += a[j];
    }
    return __shfl_sync(0xffffffffu, incl, 31);
}

// ---- compact: elements of the selected bin -> the segment's candidate list ------------------------------
// One CTA streams kCompactSpan points of one segment without any barrier; candidates (about 1 in nb)
// are collected in shared memory and appended to the segment's list with one global atomic at the end.
// If the buffer overflows (clustered data) the remaining candidates are appended warp by warp.
constexpr int kCompactThreads = 512;
constexpr int kCompactItems = 8;
constexpr int kCompactSpan = 32768;
constexpr int kCompactBuf = 4096;
__global__ void __launch_bounds__(kCompactThreads)
compact_kernel(const float *__restrict__ coord, const uint32_t *__restrict__ id,
               const Seg *__restrict__ segs, const float4 *__restrict__ bounds,
               const Sel *__restrict__ sel, int dim, uint32_t max_spans, uint32_t nb,
               uint32_t *__restrict__ cursors, unsigned long long *__restrict__ cand) {
    __shared__ unsigned long long sbuf[kCompactBuf];
    __shared__ uint32_t s_n, s_fail, s_base;
    const uint32_t s = blockIdx.x / max_spans, c = blockIdx.x - s * max_spans;
    const Seg sg = segs[s];
    const uint32_t off = c * kCompactSpan;
    if (sg.mid == kNoSplit || off >= sg.count) return;
    const uint32_t end = min(off + (uint32_t)kCompactSpan, sg.count);
    const Sel se = sel[s];
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    float lo, hi;
    seg_range(bounds, s, dim, lo, hi);
    const float scale = bin_scale(lo, hi, nb);
    if (threadIdx.x == 0) {
        s_n = 0;
        s_fail = 0xFFFFFFFFu;
    }
    __syncthreads();
    for (uint32_t sub = off; sub < end; sub += kCompactThreads * kCompactItems) {
        float v[kCompactItems];
#pragma unroll
        for (int r = 0; r < kCompactItems; ++r) {
            const uint32_t i = sub + r * kCompactThreads + threadIdx.x;
            v[r] = i < end ? __ldcs(coord + (uint64_t)sg.begin + i) : 0.0f;
        }
#pragma unroll
        for (int r = 0; r < kCompactItems; ++r) {
            const uint32_t i = sub + r * kCompactThreads + threadIdx.x;
            const bool f = i < end && bin_of(v[r], lo, scale, nb) == se.bin;
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (bal) {
                const uint32_t cnt = __popc(bal);
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&s_n, cnt);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base + cnt <= (uint32_t)kCompactBuf) {
                    if (f) sbuf[base + __popc(bal & lt)] = composite(v[r], id[(uint64_t)sg.begin + i]);
                } else {
                    uint32_t g = 0;
                    if (lane == 0) {
                        atomicMin(&s_fail, base);
                        g = atomicAdd(&cursors[3 * (uint64_t)s], cnt);
                    }
                    g = __shfl_sync(0xffffffffu, g, 0);
                    if (f)
                        cand[(uint64_t)sg.begin + g + __popc(bal & lt)] =
                            composite(v[r], id[(uint64_t)sg.begin + i]);
                }
            }
        }
    }
    __syncthreads();
    const uint32_t have = min(s_n, s_fail); // reservations are contiguous up to the first failed one
    if (have) {
        if (threadIdx.x == 0) s_base = atomicAdd(&cursors[3 * (uint64_t)s], have);
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < have; j += kCompactThreads)
            cand[(uint64_t)sg.begin + s_base + j] = sbuf[j];
    }
}

// ---- select: exact rank among the candidates, one CTA per segment -------------------------------------
__global__ void __launch_bounds__(1024)
select_kernel(const Seg *__restrict__ segs, const Sel *__restrict__ sel,
              const unsigned long long *__restrict__ cand, int dim, const float4 *__restrict__ bounds,
              float4 *__restrict__ bounds_next, unsigned long long *__restrict__ pivot,
              nbk_node *__restrict__ nodes) {
    __shared__ unsigned long long sbuf[kSelectCap];
    __shared__ uint32_t sh[256];
    __shared__ unsigned long long s_red[2][32];
    __shared__ uint32_t s_pick[3]; // digit, elements before it, elements in it
    __shared__ uint32_t s_cnt;
    const uint32_t s = blockIdx.x;
    const Seg sg = segs[s];
    if (sg.count == 0) return;
    const int tid = threadIdx.x, bd = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const float4 blo = bounds[2 * (uint64_t)s], bhi = bounds[2 * (uint64_t)s + 1];
    if (sg.mid == kNoSplit) { // carried leaf: child 2s inherits the cell
        if (tid == 0) {
            bounds_next[4 * (uint64_t)s] = blo;
            bounds_next[4 * (uint64_t)s + 1] = bhi;
        }
        return;
    }
    const Sel se = sel[s];
    const unsigned long long *cp = cand + sg.begin;
    const uint32_t c_all = se.c;
    uint32_t c = se.c, r = se.r;
    unsigned long long base = 0ull, span = ~0ull;
    if (c > (uint32_t)kSelectCap) {
        unsigned long long mn = ~0ull, mx = 0ull;
        for (uint32_t i0 = tid; i0 < c_all; i0 += 4u * bd) {
            unsigned long long v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = i0 + u * bd < c_all ? cp[i0 + u * bd] : cp[i0];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                mn = v[u] < mn ? v[u] : mn;
                mx = v[u] > mx ? v[u] : mx;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, o);
            const unsigned long long b = __shfl_xor_sync(0xffffffffu, mx, o);
            mn = a < mn ? a : mn;
            mx = b > mx ? b : mx;
        }
        if (lane == 0) {
            s_red[0][warp] = mn;
            s_red[1][warp] = mx;
        }
        __syncthreads();
        mn = ~0ull;
        mx = 0ull;
        for (int w = 0; w < (bd >> 5); ++w) {
            mn = s_red[0][w] < mn ? s_red[0][w] : mn;
            mx = s_red[1][w] > mx ? s_red[1][w] : mx;
        }
        base = mn;
        span = mx - mn;
        // narrow [base, base + span] by 8 bits of (value - base) per pass until it fits the sort
        while (c > (uint32_t)kSelectCap && span > 0ull) {
            const int bits = 64 - __clzll((long long)span);
            const int shift = bits > 8 ? bits - 8 : 0;
            for (int b = tid; b < 256; b += bd) sh[b] = 0;
            __syncthreads();
            for (uint32_t i0 = tid; i0 < c_all; i0 += 4u * bd) {
                unsigned long long v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = i0 + u * bd < c_all ? cp[i0 + u * bd] : ~0ull;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i0 + u * bd < c_all && v[u] >= base && v[u] - base <= span)
                        atomicAdd(&sh[(uint32_t)((v[u] - base) >> shift)], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t cum = 0, d = 0;
                for (; d < 255; ++d) {
                    if (r < cum + sh[d]) break;
                    cum += sh[d];
                }
                s_pick[0] = d;
                s_pick[1] = cum;
                s_pick[2] = sh[d];
            }
            __syncthreads();
            const unsigned long long step = (unsigned long long)s_pick[0] << shift;
            r -= s_pick[1];
            c = s_pick[2];
            const unsigned long long rest = span - step;
            const unsigned long long width = shift ? ((1ull << shift) - 1ull) : 0ull;
            base += step;
            span = rest < width ? rest : width;
            __syncthreads();
        }
    }
    // gather the remaining range
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    for (uint32_t i = tid; i < c_all; i += bd) {
        const unsigned long long v = cp[i];
        if (v >= base && v - base <= span) {
            const uint32_t p = atomicAdd(&s_cnt, 1u);
            if (p < (uint32_t)kSelectCap) sbuf[p] = v;
        }
    }
    __syncthreads();
    // sort it, pick rank r
    c = min(c, (uint32_t)kSelectCap); // (never exceeded: composites are unique)
    uint32_t p2 = 1;
    while (p2 < c) p2 <<= 1;
    for (uint32_t i = c + tid; i < p2; i += bd) sbuf[i] = ~0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= p2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < p2; i += bd) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = sbuf[i], b = sbuf[ixj];
                    const bool asc = (i & k) == 0;
                    if ((a > b) == asc) {
                        sbuf[i] = b;
                        sbuf[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        const unsigned long long pk = sbuf[r];
        pivot[s] = pk;
        const float split = __uint_as_float(ordered_to_float((uint32_t)(pk >> 32)));
        // kdtree_impl.hpp:116-125: split = coordinate of the rank-median element
        nbk_node nd;
        nd.dim = dim;
        nd.split = split;
        nd.left = sg.node + 1;
        nd.right = sg.right_node;
        nodes[sg.node] = nd;
        float lo[3] = {blo.x, blo.y, blo.z}, hi[3] = {bhi.x, bhi.y, bhi.z};
        float l_hi[3] = {hi[0], hi[1], hi[2]}, r_lo[3] = {lo[0], lo[1], lo[2]};
        l_hi[dim] = split;
        r_lo[dim] = split;
        bounds_next[4 * (uint64_t)s] = blo;
        bounds_next[4 * (uint64_t)s + 1] = make_float4(l_hi[0], l_hi[1], l_hi[2], 0.0f);
        bounds_next[4 * (uint64_t)s + 2] = make_float4(r_lo[0], r_lo[1], r_lo[2], 0.0f);
        bounds_next[4 * (uint64_t)s + 3] = bhi;
    }
}

// ---- partition (+ the next level's histogram) ----------------------------------------------------------
struct Columns {
    float *x, *y, *z;
    uint32_t *id;
};

template <bool FUSE_HIST>
__global__ void __launch_bounds__(kChunkThreads, 4)
partition_kernel(Columns in, Columns out, const Seg *__restrict__ segs,
                 const unsigned long long *__restrict__ pivot, uint32_t *__restrict__ cursors, int dim,
                 uint32_t max_chunks, const Seg *__restrict__ next_segs,
                 const float4 *__restrict__ bounds, uint32_t nb, uint32_t *__restrict__ hist_next) {
    extern __shared__ uint32_t sh_hist[]; // FUSE_HIST: [2][nb]
    __shared__ uint32_t cntL[kChunkCounters], cntR[kChunkCounters];
    __shared__ uint32_t s_base[2];
    const uint32_t s = blockIdx.x / max_chunks, c = blockIdx.x - s * max_chunks;
    const Seg sg = segs[s];
    const uint32_t off = c * kChunk;
    if (off >= sg.count) return;
    const uint32_t end = min(off + (uint32_t)kChunk, sg.count);
    const bool split = sg.mid != kNoSplit;
    const unsigned long long pk = split ? pivot[s] : ~0ull;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t gbase = sg.begin;

    bool hist_left = false, hist_right = false;
    float hlo = 0.0f, hscale = 0.0f;
    const int ndim = dim == 2 ? 0 : dim + 1;
    if (FUSE_HIST) {
        hist_left = next_segs[2 * (uint64_t)s].mid != kNoSplit;
        hist_right = next_segs[2 * (uint64_t)s + 1].mid != kNoSplit;
        for (uint32_t b = threadIdx.x; b < 2 * nb; b += kChunkThreads) sh_hist[b] = 0;
        float lo, hi;
        seg_range(bounds, s, ndim, lo, hi); // the split along `dim` leaves this axis' cell unchanged
        hlo = lo;
        hscale = bin_scale(lo, hi, nb);
        __syncthreads();
    }

    float px[kChunkItems], py[kChunkItems], pz[kChunkItems];
    uint32_t pid[kChunkItems];
    uint32_t rank[kChunkItems]; // bit 31: goes left
    // all 32 loads of the thread are issued before the first use (the pass is latency bound otherwise)
#pragma unroll
    for (int r = 0; r < kChunkItems; ++r) {
        const uint32_t i = off + r * kChunkThreads + threadIdx.x;
        const bool ok = i < end;
        px[r] = ok ? __ldcs(in.x + gbase + i) : 0.0f;
        py[r] = ok ? __ldcs(in.y + gbase + i) : 0.0f;
        pz[r] = ok ? __ldcs(in.z + gbase + i) : 0.0f;
        pid[r] = ok ? __ldcs(in.id + gbase + i) : 0u;
    }
#pragma unroll
    for (int r = 0; r < kChunkItems; ++r) {
        const uint32_t i = off + r * kChunkThreads + threadIdx.x;
        const bool ok = i < end;
        const float key = dim == 0 ? px[r] : (dim == 1 ? py[r] : pz[r]);
        const bool left = ok && (!split || composite(key, pid[r]) < pk);
        const unsigned bl = __ballot_sync(0xffffffffu, left);
        const unsigned br = __ballot_sync(0xffffffffu, ok && !left);
        rank[r] = left ? (0x80000000u | __popc(bl & lt)) : __popc(br & lt);
        if (lane == 0) {
            cntL[r * kChunkWarps + warp] = __popc(bl);
            cntR[r * kChunkWarps + warp] = __popc(br);
        }
        if (FUSE_HIST) {
            if (ok && (left ? hist_left : hist_right)) {
                const float nk = ndim == 0 ? px[r] : (ndim == 1 ? py[r] : pz[r]);
                atomicAdd(&sh_hist[(left ? 0u : nb) + bin_of(nk, hlo, hscale, nb)], 1u);
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        const uint32_t total = warp_scan_counters(cntL, lane);
        if (lane == 0) s_base[0] = total ? atomicAdd(&cursors[3 * (uint64_t)s + 1], total) : 0u;
    } else if (warp == 1) {
        const uint32_t total = warp_scan_counters(cntR, lane);
        if (lane == 0) s_base[1] = total ? atomicAdd(&cursors[3 * (uint64_t)s + 2], total) : 0u;
    }
    __syncthreads();
    const uint32_t right_start = split ? sg.mid : sg.count;
#pragma unroll
    for (int r = 0; r < kChunkItems; ++r) {
        const uint32_t i = off + r * kChunkThreads + threadIdx.x;
        if (i < end) {
            const bool left = rank[r] >> 31;
            const uint32_t rk = rank[r] & 0x7FFFFFFFu;
            const uint64_t dst = gbase + (left ? s_base[0] + cntL[r * kChunkWarps + warp] + rk
                                               : right_start + s_base[1] + cntR[r * kChunkWarps + warp] + rk);
            __stcs(out.x + dst, px[r]); // written once, read once by the next pass: keep it out of the way
            __stcs(out.y + dst, py[r]);
            __stcs(out.z + dst, pz[r]);
            __stcs(out.id + dst, pid[r]);
        }
    }
    if (FUSE_HIST) {
        for (uint32_t b = threadIdx.x; b < 2 * nb; b += kChunkThreads) {
            const uint32_t v = sh_hist[b];
            if (v) {
                const uint32_t child = b >= nb ? 1u : 0u;
                atomicAdd(&hist_next[(2 * (uint64_t)s + child) * nb + (b - child * nb)], v);
            }
        }
    }
}

} // namespace td
} // namespace nbk
