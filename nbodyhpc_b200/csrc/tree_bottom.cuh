// Bottom phase of the kd-tree build: one CTA finishes the whole sub-tree of one segment
// (<= 8192 points) in shared memory.
//
// The reference recurses: select the rank-median element along dim, split, recurse with the next
// dim (kdtree_impl.hpp:98-157).  Here the three coordinate orders of the segment are established
// ONCE (three stable LSD radix sorts of (orderable coordinate, local id) in shared memory); after
// that a level of the recursion is a single ranking pass: walking the order of that level's
// dimension, every element counts the elements of its own sub-segment that precede it, which is its
// rank inside the sub-segment -- rank < median_offset goes left, the element of rank ==
// median_offset supplies the split (kdtree_impl.hpp:108-125).  No element moves until the end, when
// every point is written once into its final 128-byte tile.
//
// Ties between equal coordinates are broken by id (the sorts are stable and start from id order), the
// same total order the top phase uses, so the result does not depend on the order in which the top
// phase happened to write the segment.
#pragma once

#include "tree_topdown.cuh"

namespace nbk {
namespace td {

constexpr int kBotThreads = 1024;
constexpr int kBotWarps = kBotThreads / 32;
constexpr int kBotItems = kBottomCap / kBotThreads; // 8
constexpr int kBotMaxIds = 1024;                    // heap ids of sub-segments: 10 levels
constexpr uint16_t kNoSplit16 = 0xFFFFu;

// shared-memory carve-up (bytes)
constexpr size_t kOffKeyA = 0;                                  // u32[8192]
constexpr size_t kOffKeyB = kOffKeyA + 4 * kBottomCap;          // u32[8192]
constexpr size_t kOffLidA = kOffKeyB + 4 * kBottomCap;          // u16[8192]
constexpr size_t kOffLidB = kOffLidA + 2 * kBottomCap;          // u16[8192]
constexpr size_t kOffList = kOffLidB + 2 * kBottomCap;          // u16[3][8192]
constexpr size_t kOffIdOrd = kOffList + 3 * 2 * kBottomCap;     // u16[8192]
constexpr size_t kOffSubseg = kOffIdOrd + 2 * kBottomCap;       // u16[8192]
constexpr size_t kOffWhist = kOffSubseg + 2 * kBottomCap;       // u16[32][256] (sort passes)
constexpr size_t kOffTabCnt = kOffWhist + 2 * kBotWarps * 256;  // u16[1024]
constexpr size_t kOffTabBeg = kOffTabCnt + 2 * kBotMaxIds;      // u16[1024]
constexpr size_t kOffTabMed = kOffTabBeg + 2 * kBotMaxIds;      // u16[1024]
constexpr size_t kOffTabNode = kOffTabMed + 2 * kBotMaxIds;     // u32[1024]
constexpr size_t kOffDstart = kOffTabNode + 4 * kBotMaxIds;     // u32[256]
constexpr size_t kOffMisc = kOffDstart + 4 * 256;               // u32[8]
constexpr size_t kBottomSmem = kOffMisc + 4 * 8;
// aliases: ranking histogram u16[32][512] over keyA (dead after the sorts); staging tiles
// (16 B x 8192 = 128 KB) over [0, kOffList + 2*2*kBottomCap) = keyA..list[1]; final positions over list[2]
static_assert(kOffList + 2 * 2 * kBottomCap >= 16 * (size_t)kBottomCap, "staging area too small");
static_assert(kBottomSmem <= 227 * 1024, "bottom kernel shared memory");

// Block-wide stable rank of every element inside its digit class.  Elements are laid out
// warp-blocked: warp w owns positions [w*256, w*256+256), round r covers w*256 + r*32 + lane.
// whist: u16[kBotWarps][nslots].  Returns rank[r] = number of elements with the same digit at
// smaller positions.  If `totals` is not null, totals[d] = size of class d.
// Lanes of the warp holding the same digit (nbits wide).  One ballot per bit: the MATCH.ANY
// instruction serialises over the distinct values of a warp and is several times slower here.
__device__ __forceinline__ unsigned match_digit(uint32_t d, bool ok, int nbits) {
    unsigned peers = __ballot_sync(0xffffffffu, ok);
    for (int b = 0; b < nbits; ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

__device__ __forceinline__ void block_rank(const uint32_t (&digit)[kBotItems], const bool (&ok)[kBotItems],
                                           uint32_t nslots, int nbits, uint16_t *whist, uint32_t *totals,
                                           uint32_t (&rank)[kBotItems]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    {
        uint32_t *w32 = reinterpret_cast<uint32_t *>(whist);
        for (uint32_t i = tid; i < nslots * kBotWarps / 2; i += kBotThreads) w32[i] = 0u;
    }
    __syncthreads();
    uint16_t *wh = whist + (uint32_t)warp * nslots;
#pragma unroll
    for (int r = 0; r < kBotItems; ++r) {
        const uint32_t d = ok[r] ? digit[r] : 0u;
        const unsigned peers = match_digit(d, ok[r], nbits);
        const uint32_t pre = ok[r] ? wh[d] : 0u;
        __syncwarp();
        if (ok[r] && (peers & lt) == 0u) wh[d] = (uint16_t)(pre + __popc(peers));
        __syncwarp();
        rank[r] = pre + __popc(peers & lt);
    }
    __syncthreads();
    for (uint32_t d = tid; d < nslots; d += kBotThreads) {
        uint32_t run = 0;
#pragma unroll 8
        for (int w = 0; w < kBotWarps; ++w) {
            const uint32_t t = whist[(uint32_t)w * nslots + d];
            whist[(uint32_t)w * nslots + d] = (uint16_t)run;
            run += t;
        }
        if (totals) totals[d] = run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kBotItems; ++r)
        if (ok[r]) rank[r] += wh[digit[r]];
}

// Stable LSD radix sort of (key, lid) pairs, n <= 8192, input in (keyA, lidA); the sorted lids go to
// `out`.  Byte positions on which all keys agree are skipped.
__device__ __forceinline__ void block_sort(uint32_t *keyA, uint32_t *keyB, uint16_t *lidA, uint16_t *lidB,
                                           uint32_t n, uint16_t *whist, uint32_t *dstart, uint32_t *misc,
                                           uint16_t *out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // which key bits vary at all
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
        block_rank(digit, ok, 256u, 8, whist, dstart, rank);
        // exclusive scan of the 256 class sizes (warp 0, 8 per lane)
        if (warp == 0) {
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

__global__ void __launch_bounds__(kBotThreads, 1) bottom_kernel(BottomArgs a) {
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

    uint32_t *keyA = reinterpret_cast<uint32_t *>(smem + kOffKeyA);
    uint32_t *keyB = reinterpret_cast<uint32_t *>(smem + kOffKeyB);
    uint16_t *lidA = reinterpret_cast<uint16_t *>(smem + kOffLidA);
    uint16_t *lidB = reinterpret_cast<uint16_t *>(smem + kOffLidB);
    uint16_t *list = reinterpret_cast<uint16_t *>(smem + kOffList);
    uint16_t *idord = reinterpret_cast<uint16_t *>(smem + kOffIdOrd);
    uint16_t *subseg = reinterpret_cast<uint16_t *>(smem + kOffSubseg);
    uint16_t *whist = reinterpret_cast<uint16_t *>(smem + kOffWhist);
    uint16_t *t_cnt = reinterpret_cast<uint16_t *>(smem + kOffTabCnt);
    uint16_t *t_beg = reinterpret_cast<uint16_t *>(smem + kOffTabBeg);
    uint16_t *t_med = reinterpret_cast<uint16_t *>(smem + kOffTabMed);
    uint32_t *t_node = reinterpret_cast<uint32_t *>(smem + kOffTabNode);
    uint32_t *dstart = reinterpret_cast<uint32_t *>(smem + kOffDstart);
    uint32_t *misc = reinterpret_cast<uint32_t *>(smem + kOffMisc);
    uint16_t *rhist = reinterpret_cast<uint16_t *>(smem + kOffKeyA); // ranking passes (keys are dead)
    float *stage = reinterpret_cast<float *>(smem);
    uint16_t *pos_of = list + 2 * kBottomCap;                         // over list[2]

    // ---- id order, then the three coordinate orders -------------------------------------------------
    for (uint32_t i = tid; i < n; i += kBotThreads) {
        keyA[i] = a.id[gbase + i];
        lidA[i] = (uint16_t)i;
    }
    __syncthreads();
    block_sort(keyA, keyB, lidA, lidB, n, whist, dstart, misc, idord);
#pragma unroll 1
    for (int d = 0; d < 3; ++d) {
        const float *col = d == 0 ? a.x : (d == 1 ? a.y : a.z);
        for (uint32_t i = tid; i < n; i += kBotThreads) {
            const uint16_t e = idord[i];
            keyA[i] = float_to_ordered(__float_as_uint(col[gbase + e]));
            lidA[i] = e;
        }
        __syncthreads();
        block_sort(keyA, keyB, lidA, lidB, n, whist, dstart, misc, list + d * kBottomCap);
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
    for (uint32_t i = tid; i < n; i += kBotThreads) subseg[i] = 1;
    __syncthreads();

    // ---- one ranking pass per level -------------------------------------------------------------------
    int j = 0;
    while (true) {
        const int dim = (a.level + j) % 3;
        const float *col = dim == 0 ? a.x : (dim == 1 ? a.y : a.z);
        const uint16_t *order = list + dim * kBottomCap;
        const uint32_t first = 1u << j; // heap ids of this level: [first, 2*first)
        uint32_t digit[kBotItems], rank[kBotItems];
        uint16_t e[kBotItems];
        bool ok[kBotItems];
#pragma unroll
        for (int r = 0; r < kBotItems; ++r) {
            const uint32_t i = (uint32_t)warp * 256u + r * 32u + lane;
            ok[r] = i < n;
            e[r] = ok[r] ? order[i] : (uint16_t)0;
            digit[r] = ok[r] ? (uint32_t)subseg[e[r]] - first : 0u;
        }
        block_rank(digit, ok, first, j, rhist, nullptr, rank);
#pragma unroll
        for (int r = 0; r < kBotItems; ++r) {
            if (ok[r]) {
                const uint32_t h = digit[r] + first;
                const uint32_t med = t_med[h];
                uint32_t nh = 2u * h;
                if (med != kNoSplit16) {
                    if (rank[r] >= med) nh += 1u;
                    if (rank[r] == med) a.nodes[t_node[h]].split = col[gbase + e[r]];
                }
                subseg[e[r]] = (uint16_t)nh;
            }
        }
        // tables of level j+1
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

    // ---- final position inside the leaf: rank in id order among the leaf's points ------------------------
    {
        const uint32_t first = 1u << j;
        const bool have_tables = first < (uint32_t)kBotMaxIds;
        uint32_t digit[kBotItems], rank[kBotItems];
        uint16_t e[kBotItems];
        bool ok[kBotItems];
#pragma unroll
        for (int r = 0; r < kBotItems; ++r) {
            const uint32_t i = (uint32_t)warp * 256u + r * 32u + lane;
            ok[r] = have_tables && i < n;
            e[r] = ok[r] ? idord[i] : (uint16_t)0;
            digit[r] = ok[r] ? (uint32_t)subseg[e[r]] - first : 0u;
        }
        block_rank(digit, ok, have_tables ? first : 1u, have_tables ? j : 0, rhist, nullptr, rank);
#pragma unroll
        for (int r = 0; r < kBotItems; ++r)
            if (ok[r]) pos_of[e[r]] = (uint16_t)(t_beg[digit[r] + first] + rank[r]);
        __syncthreads();
    }
    // ---- stage the tiles in shared memory, then one contiguous copy ---------------------------------------
    const bool have_positions = (1u << j) < (uint32_t)kBotMaxIds; // false only after *a.error was set
    uint32_t my_pos[kBotItems];
#pragma unroll
    for (int r = 0; r < kBotItems; ++r) {
        const uint32_t i = r * kBotThreads + tid;
        my_pos[r] = i < n ? (have_positions ? (uint32_t)pos_of[i] : i) : 0u;
    }
    __syncthreads(); // pos_of (list[2]) is outside the staging area, rhist (keyA) is inside: all reads done
#pragma unroll
    for (int r = 0; r < kBotItems; ++r) {
        const uint32_t i = r * kBotThreads + tid;
        if (i < n) {
            const uint32_t id = a.id[gbase + i];
            put_tile(stage, my_pos[r], a.x[gbase + i], a.y[gbase + i], a.z[gbase + i],
                     a.idx0 ? a.idx0[id] : id);
        }
    }
    __syncthreads();
    float4 *dst = reinterpret_cast<float4 *>(a.tiles + (gbase >> 3) * 32);
    const float4 *src = reinterpret_cast<const float4 *>(stage);
    for (uint32_t i = tid; i < n; i += kBotThreads) dst[i] = src[i]; // n points = n float4
}

} // namespace td
} // namespace nbk
