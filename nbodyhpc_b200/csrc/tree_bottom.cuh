// Bottom phase of the kd-tree build: one CTA finishes the whole sub-tree of one segment
// (<= 8192 points) in shared memory.
//
// The reference recurses: select the rank-median element along dim, split, recurse with the next
// dim (kdtree_impl.hpp:98-157).  Here the three coordinate orders of the segment are established
// ONCE, as three lists of local element ids sorted by (orderable coordinate, id).  A level of the
// recursion then moves no point data at all:
//   * every sub-segment is the SAME position range [beg, beg+cnt) in all three lists;
//   * in the list of the level's split dimension the rank-median element simply sits at position
//     beg + median_offset: it supplies the split (kdtree_impl.hpp:108-125) and everything from there
//     on is the right child;
//   * the other two lists are stably partitioned inside every range by that left/right flag (one
//     block-wide prefix sum per level), which keeps them sorted inside both children.
// At the end the x-list is the final leaf order; every point is read once more and written once into
// its 128-byte tile.
//
// Sorting: one counting pass over 8192 value-range bins (shared-memory atomics) plus an exact
// (coordinate, id) ranking inside each bin.  Degenerate distributions (a bin with more than kRunMax
// elements: heavy ties, extreme clustering) take a stable LSD radix sort started from id order
// instead.  Either way the order is the total order (coordinate, id) that the top phase uses, so the
// result does not depend on the order in which the top phase happened to write the segment.
#pragma once

#include "tree_topdown.cuh"

namespace nbk {
namespace td {

constexpr int kBotThreads = kBottomCap / 8;
constexpr int kBotWarps = kBotThreads / 32;
constexpr int kBotItems = kBottomCap / kBotThreads; // 8 = one tile / one block_size unit per thread
constexpr int kBotMaxIds = kBottomCap / 8;          // heap ids of sub-segments (leaves hold >= 16 points)
constexpr uint16_t kNoSplit16 = 0xFFFFu;
constexpr uint32_t kSortBins = kBottomCap;
constexpr uint32_t kRunMax = 64;
static_assert(kBotItems == 8, "a thread owns 8 consecutive list positions (one tile)");

// shared-memory carve-up (bytes)
constexpr size_t kOffList = 0;                                   // u16[2][3][8192]  lists [buffer][dim]
constexpr size_t kOffTmpKey = kOffList + 6 * 2 * kBottomCap;     // u32[8192]        sort scratch
constexpr size_t kOffTmpLid = kOffTmpKey + 4 * kBottomCap;       // u16[8192]
constexpr size_t kOffHist = kOffTmpLid + 2 * kBottomCap;         // u32[8193]        (radix: u16[32][256])
constexpr size_t kOffSegPos = kOffHist + 4 * (kSortBins + 4);    // u16[8192]        sub-segment of a position
constexpr size_t kOffSide = kOffSegPos + 2 * kBottomCap;         // u8[8192]         1 = right child
constexpr size_t kOffTabCnt = kOffSide + kBottomCap;             // u16[1024]
constexpr size_t kOffTabBeg = kOffTabCnt + 2 * kBotMaxIds;       // u16[1024]
constexpr size_t kOffTabMed = kOffTabBeg + 2 * kBotMaxIds;       // u16[1024]
constexpr size_t kOffTabNode = kOffTabMed + 2 * kBotMaxIds;      // u32[1024]
constexpr size_t kOffThreadP = kOffTabNode + 4 * kBotMaxIds;     // u32[1024]        packed prefix per thread
constexpr size_t kOffDstart = kOffThreadP + 4 * kBotThreads;     // u32[256]
constexpr size_t kOffMisc = kOffDstart + 4 * 256;                // u32[80]
constexpr size_t kBottomSmem = kOffMisc + 4 * 80;
// radix fallback aliases: keyA = tmpKey, lidA = tmpLid, whist = hist; keyB over the second buffers of
// the x and y lists (contiguous 32 KB), lidB over the second buffer of the z list, idord over segpos:
// all unused while sorting
static_assert(kBottomSmem <= 227 * 1024, "bottom kernel shared memory");
static_assert(kBotThreads % 32 == 0 && kBotThreads <= 1024 && kSortBins % kBotThreads == 0, "bottom kernel geometry");

// ---- radix fallback (stable, 8 bits per pass) ----------------------------------------------------------
// Lanes of the warp holding the same digit.  One ballot per bit: the MATCH.ANY instruction
// serialises over the distinct values of a warp and is several times slower here.
__device__ __forceinline__ unsigned match_digit(uint32_t d, bool ok, int nbits) {
    unsigned peers = __ballot_sync(0xffffffffu, ok);
    for (int b = 0; b < nbits; ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

// Block-wide stable rank of every element inside its digit class.  Elements are laid out
// warp-blocked: warp w owns positions [w*256, w*256+256), round r covers w*256 + r*32 + lane.
// whist: u16[kBotWarps][256].  rank[r] = number of elements with the same digit at smaller
// positions; totals[d] = size of class d.
__device__ __forceinline__ void block_rank(const uint32_t (&digit)[kBotItems], const bool (&ok)[kBotItems],
                                           uint16_t *whist, uint32_t *totals, uint32_t (&rank)[kBotItems]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    {
        uint32_t *w32 = reinterpret_cast<uint32_t *>(whist);
        for (uint32_t i = tid; i < 256u * kBotWarps / 2; i += kBotThreads) w32[i] = 0u;
    }
    __syncthreads();
    uint16_t *wh = whist + (uint32_t)warp * 256u;
#pragma unroll
    for (int r = 0; r < kBotItems; ++r) {
        const uint32_t d = ok[r] ? digit[r] : 0u;
        const unsigned peers = match_digit(d, ok[r], 8);
        const uint32_t pre = ok[r] ? wh[d] : 0u;
        __syncwarp();
        if (ok[r] && (peers & lt) == 0u) wh[d] = (uint16_t)(pre + __popc(peers));
        __syncwarp();
        rank[r] = pre + __popc(peers & lt);
    }
    __syncthreads();
    for (uint32_t d = tid; d < 256u; d += kBotThreads) {
        uint32_t run = 0;
#pragma unroll 8
        for (int w = 0; w < kBotWarps; ++w) {
            const uint32_t t = whist[(uint32_t)w * 256u + d];
            whist[(uint32_t)w * 256u + d] = (uint16_t)run;
            run += t;
        }
        totals[d] = run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kBotItems; ++r)
        if (ok[r]) rank[r] += wh[digit[r]];
}

// Stable LSD radix sort of (key, lid) pairs, n <= 8192, input in (keyA, lidA); the sorted lids go to
// `out`.  Byte positions on which all keys agree are skipped.
__device__ __noinline__ void block_sort(uint32_t *keyA, uint32_t *keyB, uint16_t *lidA, uint16_t *lidB,
                                        uint32_t n, uint16_t *whist, uint32_t *dstart, uint32_t *misc,
                                        uint16_t *out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) misc[0] = 0u;
    __syncthreads();
    {
        const uint32_t k0 = keyA[0];
        uint32_t diff = 0;
        for (uint32_t i = tid; i < n; i += kBotThreads) diff |= keyA[i] ^ k0;
        diff = __reduce_or_sync(0xffffffffu, diff);
        if (lane == 0 && diff) atomicOr(&misc[0], diff);
    }
    __syncthreads();
    const uint32_t varying = misc[0];
    uint32_t *kin = keyA, *kout = keyB;
    uint16_t *lin = lidA, *lout = lidB;
    for (int shift = 0; shift < 32; shift += 8) {
        if (((varying >> shift) & 0xFFu) == 0u) continue;
        uint32_t k[kBotItems], digit[kBotItems], rank[kBotItems];
        uint16_t l[kBotItems];
        bool ok[kBotItems];
#pragma unroll
        for (int r = 0; r < kBotItems; ++r) {
            const uint32_t i = (uint32_t)warp * 256u + r * 32u + lane;
            ok[r] = i < n;
            k[r] = ok[r] ? kin[i] : 0u;
            l[r] = ok[r] ? lin[i] : (uint16_t)0;
            digit[r] = (k[r] >> shift) & 0xFFu;
        }
        block_rank(digit, ok, whist, dstart, rank);
        if (warp == 0) { // exclusive scan of the 256 class sizes, 8 per lane
            uint32_t a[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a[j] = dstart[lane * 8 + j];
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
            for (int j = 0; j < 8; ++j) {
                dstart[lane * 8 + j] = run;
                run += a[j];
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kBotItems; ++r) {
            if (ok[r]) {
                const uint32_t pos = dstart[digit[r]] + rank[r];
                kout[pos] = k[r];
                lout[pos] = l[r];
            }
        }
        __syncthreads();
        uint32_t *tk = kin; kin = kout; kout = tk;
        uint16_t *tl = lin; lin = lout; lout = tl;
    }
    for (uint32_t i = tid; i < n; i += kBotThreads) out[i] = lin[i];
    __syncthreads();
}

// Block-wide exclusive prefix sum of one uint32 per thread; wsum: u32[32] scratch.
__device__ __forceinline__ uint32_t block_exclusive_sum(uint32_t v, uint32_t *wsum, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    const uint32_t w = lane < kBotWarps ? wsum[lane] : 0u;
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
    }
    total = __shfl_sync(0xffffffffu, winc, 31);
    const uint32_t wofs = __shfl_sync(0xffffffffu, winc - w, warp);
    __syncthreads();
    return wofs + incl - v;
}

struct BottomArgs {
    const float *x, *y, *z;  // the segment's columns (output of the last partition pass)
    const uint32_t *id;      // position of each point in the caller's array
    const uint32_t *idx0;    // caller-supplied indices (nbk_tree_build_soa) or null
    const Seg *segs;         // level `level` of the plan
    uint32_t nseg;
    int level;               // dim of the first split below = level % 3
    uint32_t leaf;           // max(leaf_size, 2*block)
    uint32_t block;
    const uint32_t *lut;     // nodes of the sub-tree over 8*i points
    nbk_node *nodes;
    float *tiles;
    uint32_t *error;         // set to 1 if a sub-tree is deeper than the tables allow
};

__device__ __forceinline__ void put_tile(float *tiles, uint64_t p, float x, float y, float z, uint32_t idx) {
    float *f = tiles + (p >> 3) * 32 + (p & 7);
    f[0] = x;
    f[8] = y;
    f[16] = z;
    f[24] = __uint_as_float(idx);
}

__global__ void __launch_bounds__(kBotThreads, kBotThreads <= 512 ? 2 : 1) bottom_kernel(BottomArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Seg sg = a.segs[blockIdx.x];
    if (sg.count == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t gbase = sg.begin;
    const uint32_t n = sg.count;

    if (n <= a.leaf || n > (uint32_t)kBottomCap) {
        // a leaf (of any size): node record + points in arrival order (kdtree_impl.hpp:101-105)
        if (tid == 0) {
            nbk_node nd;
            nd.dim = -1;
            nd.split = 0.0f;
            nd.left = sg.begin;
            nd.right = sg.begin + n;
            a.nodes[sg.node] = nd;
        }
        for (uint32_t i = tid; i < n; i += kBotThreads) {
            const uint32_t id = a.id[gbase + i];
            put_tile(a.tiles, gbase + i, a.x[gbase + i], a.y[gbase + i], a.z[gbase + i],
                     a.idx0 ? a.idx0[id] : id);
        }
        return;
    }

    uint16_t *lists = reinterpret_cast<uint16_t *>(smem + kOffList); // [buf][dim][8192]
    uint32_t *tmp_key = reinterpret_cast<uint32_t *>(smem + kOffTmpKey);
    uint16_t *tmp_lid = reinterpret_cast<uint16_t *>(smem + kOffTmpLid);
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem + kOffHist);
    uint16_t *seg_of_pos = reinterpret_cast<uint16_t *>(smem + kOffSegPos);
    uint8_t *side = reinterpret_cast<uint8_t *>(smem + kOffSide);
    uint16_t *t_cnt = reinterpret_cast<uint16_t *>(smem + kOffTabCnt);
    uint16_t *t_beg = reinterpret_cast<uint16_t *>(smem + kOffTabBeg);
    uint16_t *t_med = reinterpret_cast<uint16_t *>(smem + kOffTabMed);
    uint32_t *t_node = reinterpret_cast<uint32_t *>(smem + kOffTabNode);
    uint32_t *thread_p = reinterpret_cast<uint32_t *>(smem + kOffThreadP);
    uint32_t *dstart = reinterpret_cast<uint32_t *>(smem + kOffDstart);
    uint32_t *misc = reinterpret_cast<uint32_t *>(smem + kOffMisc);
    auto list_of = [&](int d, int buf) { return lists + ((size_t)buf * 3 + d) * kBottomCap; };

    // ---- the three coordinate orders --------------------------------------------------------------------
    bool have_idord = false;
    uint16_t *idord = seg_of_pos; // radix fallback only
#pragma unroll 1
    for (int d = 0; d < 3; ++d) {
        const float *col = d == 0 ? a.x : (d == 1 ? a.y : a.z);
        uint16_t *out = list_of(d, 0);
        float v[kBotItems];
        uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
#pragma unroll
        for (int r = 0; r < kBotItems; ++r) {
            const uint32_t e = r * kBotThreads + tid;
            v[r] = e < n ? col[gbase + e] : 0.0f;
            if (e < n) {
                const uint32_t key = float_to_ordered(__float_as_uint(v[r]));
                kmin = min(kmin, key);
                kmax = max(kmax, key);
            }
        }
        kmin = __reduce_min_sync(0xffffffffu, kmin);
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        if (lane == 0) {
            misc[8 + warp] = kmin;
            misc[40 + warp] = kmax;
        }
        for (uint32_t b = tid; b < kSortBins + 1; b += kBotThreads) hist[b] = 0u;
        __syncthreads();
        kmin = __reduce_min_sync(0xffffffffu, lane < kBotWarps ? misc[8 + lane] : 0xFFFFFFFFu);
        kmax = __reduce_max_sync(0xffffffffu, lane < kBotWarps ? misc[40 + lane] : 0u);
        const float lo = __uint_as_float(ordered_to_float(kmin)), hi = __uint_as_float(ordered_to_float(kmax));
        const float scale = bin_scale(lo, hi, kSortBins);
        uint32_t bin[kBotItems], lr[kBotItems];
#pragma unroll
        for (int r = 0; r < kBotItems; ++r) {
            const uint32_t e = r * kBotThreads + tid;
            bin[r] = bin_of(v[r], lo, scale, kSortBins);
            lr[r] = e < n ? atomicAdd(&hist[bin[r]], 1u) : 0u;
        }
        __syncthreads();
        // exclusive scan of the bin counts (8 per thread) -> bin starts; hist[kSortBins] = n
        bool long_run = false;
        {
            constexpr int PER = kSortBins / kBotThreads;
            uint32_t c[PER], sum = 0;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                c[i] = hist[tid * PER + i];
                sum += c[i];
                long_run = long_run || c[i] > kRunMax;
            }
            uint32_t total;
            uint32_t run = block_exclusive_sum(sum, misc + 8, total);
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                hist[tid * PER + i] = run;
                run += c[i];
            }
            if (tid == 0) hist[kSortBins] = n;
        }
        if (__syncthreads_or(long_run ? 1 : 0)) {
            // degenerate distribution: stable radix sort by coordinate, starting from id order
            uint32_t *keyA = tmp_key, *keyB = reinterpret_cast<uint32_t *>(list_of(0, 1)); // x+y second buffers
            uint16_t *lidA = tmp_lid, *lidB = list_of(2, 1);
            uint16_t *whist = reinterpret_cast<uint16_t *>(hist);
            if (!have_idord) {
                for (uint32_t i = tid; i < n; i += kBotThreads) {
                    keyA[i] = a.id[gbase + i];
                    lidA[i] = (uint16_t)i;
                }
                __syncthreads();
                block_sort(keyA, keyB, lidA, lidB, n, whist, dstart, misc, idord);
                have_idord = true;
            }
            for (uint32_t i = tid; i < n; i += kBotThreads) {
                const uint16_t e = idord[i];
                keyA[i] = float_to_ordered(__float_as_uint(col[gbase + e]));
                lidA[i] = e;
            }
            __syncthreads();
            block_sort(keyA, keyB, lidA, lidB, n, whist, dstart, misc, out);
            continue;
        }
        // scatter by bin (arrival order inside a bin), then rank exactly inside each bin
#pragma unroll
        for (int r = 0; r < kBotItems; ++r) {
            const uint32_t e = r * kBotThreads + tid;
            if (e < n) {
                const uint32_t pos = hist[bin[r]] + lr[r];
                tmp_key[pos] = float_to_ordered(__float_as_uint(v[r]));
                tmp_lid[pos] = (uint16_t)e;
            }
        }
        __syncthreads();
#pragma unroll 2
        for (int r = 0; r < kBotItems; ++r) {
            const uint32_t p = r * kBotThreads + tid;
            if (p < n) {
                const uint32_t key = tmp_key[p];
                const uint16_t lid = tmp_lid[p];
                const uint32_t b = bin_of(__uint_as_float(ordered_to_float(key)), lo, scale, kSortBins);
                const uint32_t s0 = hist[b], s1 = hist[b + 1];
                uint32_t lt = 0, eq = 0;
                for (uint32_t q = s0; q < s1; ++q) {
                    const uint32_t kq = tmp_key[q];
                    lt += kq < key ? 1u : 0u;
                    eq += kq == key ? 1u : 0u;
                }
                if (eq > 1u) { // equal coordinates: by id (rare)
                    const uint32_t my_id = a.id[gbase + lid];
                    for (uint32_t q = s0; q < s1; ++q)
                        if (tmp_key[q] == key && a.id[gbase + tmp_lid[q]] < my_id) ++lt;
                }
                out[s0 + lt] = lid;
            }
        }
        __syncthreads();
    }

    // ---- root of the sub-tree ------------------------------------------------------------------------
    const uint32_t block = a.block;
    if (tid == 0) {
        const uint32_t med = (n / 2 / block) * block;
        t_cnt[1] = (uint16_t)n; // n <= 8192 < 65535
        t_beg[1] = 0;
        t_med[1] = (uint16_t)med;
        t_node[1] = sg.node;
        nbk_node *nd = a.nodes + sg.node;
        nd->dim = a.level % 3;
        nd->left = sg.node + 1;
        nd->right = sg.node + 1 + a.lut[med >> 3];
    }
    for (uint32_t i = tid; i < n; i += kBotThreads) seg_of_pos[i] = 1;
    __syncthreads();

    // ---- levels ------------------------------------------------------------------------------------------
    // Thread t owns list positions [8t, 8t+8): sub-segment boundaries and medians are multiples of
    // block_size (a multiple of 8), so the 8 positions always lie in one sub-segment and on one side.
    const uint32_t p0 = (uint32_t)tid * 8u;
    const bool mine = p0 < n;
    int cur0 = 0, cur1 = 0, cur2 = 0; // which buffer holds the current x / y / z list
    int j = 0;
    while (true) {
        const int dim = (a.level + j) % 3;
        const int d1 = dim == 2 ? 0 : dim + 1, d2 = d1 == 2 ? 0 : d1 + 1;
        const int cdim = dim == 0 ? cur0 : (dim == 1 ? cur1 : cur2);
        const int c1 = d1 == 0 ? cur0 : (d1 == 1 ? cur1 : cur2);
        const int c2 = d2 == 0 ? cur0 : (d2 == 1 ? cur1 : cur2);
        const uint32_t first = 1u << j; // heap ids of this level: [first, 2*first)
        // A: side of every element from the split dimension's list; split value; new sub-segment ids
        if (mine) {
            const uint32_t h = seg_of_pos[p0];
            const uint32_t med = t_med[h];
            uint32_t nh = 2u * h;
            if (med != kNoSplit16) {
                const uint32_t rel = p0 - t_beg[h];
                const bool right = rel >= med;
                const uint4 pk = *reinterpret_cast<const uint4 *>(list_of(dim, cdim) + p0);
                const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
                const uint8_t flag = right ? 1 : 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    side[w[i] & 0xFFFFu] = flag;
                    side[w[i] >> 16] = flag;
                }
                if (rel == med) {
                    const float *col = dim == 0 ? a.x : (dim == 1 ? a.y : a.z);
                    a.nodes[t_node[h]].split = col[gbase + (w[0] & 0xFFFFu)]; // kdtree_impl.hpp:116-125
                }
                if (right) nh += 1u;
            }
            const uint32_t two = nh | (nh << 16);
            *reinterpret_cast<uint4 *>(seg_of_pos + p0) = make_uint4(two, two, two, two);
        }
        __syncthreads();
        // B: stable partition of the other two lists inside every splitting sub-segment
        {
            uint32_t e1[4] = {0, 0, 0, 0}, e2[4] = {0, 0, 0, 0};
            uint32_t left1 = 0, left2 = 0; // bit i: element i goes left
            uint32_t med = kNoSplit16, beg = 0;
            if (mine) {
                const uint32_t h = seg_of_pos[p0] >> 1;
                med = t_med[h];
                beg = t_beg[h];
                const uint4 a1 = *reinterpret_cast<const uint4 *>(list_of(d1, c1) + p0);
                const uint4 a2 = *reinterpret_cast<const uint4 *>(list_of(d2, c2) + p0);
                e1[0] = a1.x; e1[1] = a1.y; e1[2] = a1.z; e1[3] = a1.w;
                e2[0] = a2.x; e2[1] = a2.y; e2[2] = a2.z; e2[3] = a2.w;
                if (med != kNoSplit16) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        left1 |= (side[e1[i] & 0xFFFFu] ? 0u : 1u) << (2 * i);
                        left1 |= (side[e1[i] >> 16] ? 0u : 1u) << (2 * i + 1);
                        left2 |= (side[e2[i] & 0xFFFFu] ? 0u : 1u) << (2 * i);
                        left2 |= (side[e2[i] >> 16] ? 0u : 1u) << (2 * i + 1);
                    }
                }
            }
            uint32_t total;
            const uint32_t packed = __popc(left1) | (__popc(left2) << 16); // both prefix sums at once (<= 8192)
            const uint32_t myp = block_exclusive_sum(packed, misc + 8, total);
            thread_p[tid] = myp;
            __syncthreads();
            if (mine) {
                uint16_t *o1 = list_of(d1, c1 ^ 1), *o2 = list_of(d2, c2 ^ 1);
                if (med != kNoSplit16) {
                    const uint32_t at_beg = thread_p[beg >> 3];
                    uint32_t l1 = (myp & 0xFFFFu) - (at_beg & 0xFFFFu); // lefts of this sub-segment before p0
                    uint32_t l2 = (myp >> 16) - (at_beg >> 16);
                    const uint32_t rel = p0 - beg;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t v1 = (i & 1) ? (e1[i >> 1] >> 16) : (e1[i >> 1] & 0xFFFFu);
                        const uint32_t v2 = (i & 1) ? (e2[i >> 1] >> 16) : (e2[i >> 1] & 0xFFFFu);
                        const bool g1 = (left1 >> i) & 1u, g2 = (left2 >> i) & 1u;
                        const uint32_t q1 = g1 ? beg + l1 : beg + med + (rel + i - l1);
                        const uint32_t q2 = g2 ? beg + l2 : beg + med + (rel + i - l2);
                        o1[q1] = (uint16_t)v1;
                        o2[q2] = (uint16_t)v2;
                        l1 += g1 ? 1u : 0u;
                        l2 += g2 ? 1u : 0u;
                    }
                } else {
                    *reinterpret_cast<uint4 *>(o1 + p0) = make_uint4(e1[0], e1[1], e1[2], e1[3]);
                    *reinterpret_cast<uint4 *>(o2 + p0) = make_uint4(e2[0], e2[1], e2[2], e2[3]);
                }
            }
            if (d1 == 0) cur0 ^= 1; else if (d1 == 1) cur1 ^= 1; else cur2 ^= 1;
            if (d2 == 0) cur0 ^= 1; else if (d2 == 1) cur1 ^= 1; else cur2 ^= 1;
        }
        // C: tables of level j+1
        bool splits_more = false;
        if (2u * first < (uint32_t)kBotMaxIds) {
            const int ndim = (a.level + j + 1) % 3;
            for (uint32_t h = 2u * first + tid; h < 4u * first; h += kBotThreads) {
                const uint32_t p = h >> 1;
                const uint32_t pmed = t_med[p];
                uint32_t cnt, beg, node;
                bool fresh;
                if (pmed != kNoSplit16) {
                    fresh = true;
                    if ((h & 1u) == 0u) {
                        cnt = pmed;
                        beg = t_beg[p];
                        node = t_node[p] + 1u;
                    } else {
                        cnt = (uint32_t)t_cnt[p] - pmed;
                        beg = (uint32_t)t_beg[p] + pmed;
                        node = t_node[p] + 1u + a.lut[pmed >> 3];
                    }
                } else {
                    fresh = false;
                    cnt = (h & 1u) == 0u ? (uint32_t)t_cnt[p] : 0u;
                    beg = t_beg[p];
                    node = t_node[p];
                }
                const bool sp = cnt > a.leaf;
                const uint32_t med = sp ? (cnt / 2 / block) * block : (uint32_t)kNoSplit16;
                t_cnt[h] = (uint16_t)cnt;
                t_beg[h] = (uint16_t)beg;
                t_med[h] = (uint16_t)med;
                t_node[h] = node;
                if (fresh) {
                    nbk_node *nd = a.nodes + node;
                    if (sp) {
                        nd->dim = ndim;
                        nd->left = node + 1u;
                        nd->right = node + 1u + a.lut[med >> 3];
                    } else {
                        nbk_node leafnode;
                        leafnode.dim = -1;
                        leafnode.split = 0.0f;
                        leafnode.left = sg.begin + beg;
                        leafnode.right = sg.begin + beg + cnt;
                        *nd = leafnode;
                    }
                }
                splits_more = splits_more || sp;
            }
        } else if (tid == 0) {
            // cannot happen for block >= 8, leaf >= 16, n <= 8192 (depth <= 9); refuse loudly
            bool any = false;
            for (uint32_t h = first; h < 2u * first; ++h) any = any || t_med[h] != kNoSplit16;
            if (any) *a.error = 1u;
        }
        ++j;
        if (!__syncthreads_or(splits_more ? 1 : 0)) break;
    }

    // ---- the x-list is the final order -------------------------------------------------------------------------
    // pos_of = inverse of the x-list; points are read coalesced by element id, placed into staged
    // 128-byte tiles in shared memory (over the lists and the sort keys, all dead now), copied out.
    uint16_t *pos_of = tmp_lid;
    float *stage = reinterpret_cast<float *>(smem);
    static_assert(kOffTmpLid >= 16 * (size_t)kBottomCap, "staging area must end before pos_of");
    if (mine) {
        const uint4 pk = *reinterpret_cast<const uint4 *>(list_of(0, cur0) + p0);
        const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            pos_of[w[i] & 0xFFFFu] = (uint16_t)(p0 + 2 * i);
            pos_of[w[i] >> 16] = (uint16_t)(p0 + 2 * i + 1);
        }
    }
    __syncthreads();
    {
        float px[kBotItems], py[kBotItems], pz[kBotItems];
        uint32_t pi[kBotItems], pp[kBotItems];
#pragma unroll
        for (int r = 0; r < kBotItems; ++r) {
            const uint32_t e = r * kBotThreads + tid;
            const bool ok = e < n;
            px[r] = ok ? a.x[gbase + e] : 0.0f;
            py[r] = ok ? a.y[gbase + e] : 0.0f;
            pz[r] = ok ? a.z[gbase + e] : 0.0f;
            pi[r] = ok ? a.id[gbase + e] : 0u;
            pp[r] = ok ? pos_of[e] : 0u;
        }
        if (a.idx0) {
#pragma unroll
            for (int r = 0; r < kBotItems; ++r)
                if (r * kBotThreads + tid < n) pi[r] = a.idx0[pi[r]];
        }
        __syncthreads(); // every thread has read its pos_of / list entries: the staging area may be overwritten
#pragma unroll
        for (int r = 0; r < kBotItems; ++r)
            if (r * kBotThreads + tid < n) put_tile(stage, pp[r], px[r], py[r], pz[r], pi[r]);
    }
    // The staged tiles are one contiguous run of n x 16 bytes in shared memory and in the arena: one TMA
    // bulk store (cp.async.bulk, shared -> global) issued by a single thread moves it; the generic-proxy
    // writes above are made visible to the async proxy first.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        float *dst = a.tiles + (gbase >> 3) * 32;
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(stage);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                     "cp.async.bulk.commit_group;\n"
                     "cp.async.bulk.wait_group.read 0;\n"
                     :
                     : "l"(dst), "r"(src), "r"(n * 16u)
                     : "memory");
    }
}

} // namespace td
} // namespace nbk
