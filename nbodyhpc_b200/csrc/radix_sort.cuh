// Stable LSD radix sort of (key, uint32 payload) pairs, 8 bits per pass, hand-written for sm_100a.
//
// Used by the tree build (64-bit keys: segment id | orderable coordinate) and by the query path
// (32-bit Morton keys, payload = query id).  Per pass: one histogram kernel, one exclusive scan of
// the [digit][tile] table, one scatter kernel that ranks every tile stably with warp match-any,
// reorders it in shared memory and writes each digit run contiguously (full 32-byte sectors).
#pragma once

#include "common.cuh"

namespace nbk {
namespace rs {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

template <typename KeyT> struct Items { static constexpr int value = sizeof(KeyT) == 8 ? 12 : 16; };

// ---- exclusive scan of uint32 (3-phase, recursive on the block sums) --------------------------
constexpr int kScanItems = 16;
constexpr int kScanTile = kThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums,
                                                         uint32_t &block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t wofs = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        uint32_t s = warp_sums[w];
        if (w < warp) wofs += s;
        total += s;
    }
    __syncthreads();
    block_total = total;
    return wofs + incl - v;
}

__global__ void __launch_bounds__(kThreads)
scan_reduce_kernel(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ sums) {
    __shared__ uint32_t warp_sums[kWarps];
    uint64_t base = (uint64_t)blockIdx.x * kScanTile;
    uint32_t acc = 0;
#pragma unroll
    for (int it = 0; it < kScanItems; ++it) {
        uint64_t i = base + (uint64_t)it * kThreads + threadIdx.x;
        if (i < n) acc += in[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kWarps; ++w) t += warp_sums[w];
        sums[blockIdx.x] = t;
    }
}

// In-place exclusive scan of one tile per block; `offsets` (may be null) holds the scanned sums.
__global__ void __launch_bounds__(kThreads)
scan_apply_kernel(uint32_t *__restrict__ data, uint64_t n, const uint32_t *__restrict__ offsets) {
    __shared__ uint32_t warp_sums[kWarps];
    // thread-contiguous items so the in-thread running sum follows the array order
    uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t sum = 0;
#pragma unroll
    for (int it = 0; it < kScanItems; ++it) {
        uint64_t i = base + it;
        v[it] = i < n ? data[i] : 0u;
        sum += v[it];
    }
    uint32_t total;
    uint32_t ofs = block_exclusive_scan(sum, warp_sums, total);
    if (offsets) ofs += offsets[blockIdx.x];
#pragma unroll
    for (int it = 0; it < kScanItems; ++it) {
        uint64_t i = base + it;
        if (i < n) data[i] = ofs;
        ofs += v[it];
    }
}

// workspace needed by exclusive_scan for n entries (uint32 count)
inline uint64_t scan_workspace_entries(uint64_t n) {
    uint64_t total = 0;
    while (n > (uint64_t)kScanTile) {
        n = div_up(n, kScanTile);
        total += n;
    }
    return total + 1;
}

inline void exclusive_scan(uint32_t *data, uint64_t n, uint32_t *work, cudaStream_t stream) {
    if (n == 0) return;
    uint64_t nb = div_up(n, kScanTile);
    if (nb == 1) {
        scan_apply_kernel<<<1, kThreads, 0, stream>>>(data, n, nullptr);
        NBK_LAUNCHED();
        return;
    }
    scan_reduce_kernel<<<(unsigned)nb, kThreads, 0, stream>>>(data, n, work);
    NBK_LAUNCHED();
    exclusive_scan(work, nb, work + nb, stream);
    scan_apply_kernel<<<(unsigned)nb, kThreads, 0, stream>>>(data, n, work);
    NBK_LAUNCHED();
}

// ---- per-tile digit histogram -------------------------------------------------------------------
template <typename KeyT>
__global__ void __launch_bounds__(kThreads)
hist_kernel(const KeyT *__restrict__ keys, uint64_t n, int shift, uint32_t *__restrict__ hist,
            uint32_t ntiles) {
    constexpr int ITEMS = Items<KeyT>::value;
    __shared__ uint32_t h[kRadix];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint64_t base = (uint64_t)blockIdx.x * (kThreads * ITEMS);
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        uint64_t i = base + (uint64_t)it * kThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & (kRadix - 1)], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// ---- scatter ------------------------------------------------------------------------------------
// STABLE = false ranks the items of a warp with shared-memory atomics instead of the ballot match: items
// of one 32-item row that share the digit may then swap places, i.e. the result is sorted by the LAST
// digit and only nearly sorted (displacements < 32 positions per pass) by the earlier ones.  That is all
// the query ordering needs -- any permutation gives the same answers, the order only has to keep spatial
// neighbours in the same warp -- and it makes the pass cheaper; the build's sorts stay stable.
template <typename KeyT, bool STABLE = true>
__global__ void __launch_bounds__(kThreads)
scatter_kernel(const KeyT *__restrict__ kin, const uint32_t *__restrict__ vin,
               KeyT *__restrict__ kout, uint32_t *__restrict__ vout, uint64_t n, int shift,
               const uint32_t *__restrict__ offsets, uint32_t ntiles) {
    constexpr int ITEMS = Items<KeyT>::value;
    constexpr int TILE = kThreads * ITEMS;
    __shared__ KeyT skeys[TILE];
    __shared__ uint32_t svals[TILE];
    __shared__ uint32_t whist[kWarps * kRadix]; // per-warp digit counts, then warp start offsets
    __shared__ uint32_t dstart[kRadix];         // tile-local start of every digit run
    __shared__ uint32_t gofs[kRadix];           // global start of this tile's run of every digit
    __shared__ uint32_t warp_sums[kWarps];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < kWarps * kRadix; i += kThreads) whist[i] = 0;
    __syncthreads();

    const uint64_t tile_base = (uint64_t)blockIdx.x * TILE;
    const uint64_t warp_base = tile_base + (uint64_t)warp * (32 * ITEMS);
    KeyT k[ITEMS];
    uint32_t v[ITEMS];
    uint32_t r[ITEMS];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        uint64_t i = warp_base + it * 32 + lane;
        bool ok = i < n;
        k[it] = ok ? kin[i] : (KeyT)0;
        v[it] = ok ? vin[i] : 0u;
    }
    uint32_t *wh = whist + warp * kRadix;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        uint64_t i = warp_base + it * 32 + lane;
        bool ok = i < n;
        uint32_t d = ok ? ((uint32_t)(k[it] >> shift) & (kRadix - 1)) : 0xFFFFu;
        if (!STABLE) {
            r[it] = ok ? atomicAdd(&wh[d], 1u) : 0u;
            continue;
        }
        // lanes with the same digit: one ballot per bit (MATCH.ANY serialises over the distinct values of
        // the warp and is several times slower for 8-bit digits)
        uint32_t peers = __ballot_sync(0xffffffffu, ok);
#pragma unroll
        for (int b = 0; b < kRadixBits; ++b) {
            const bool bit = (d >> b) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? bal : ~bal;
        }
        uint32_t pre = ok ? wh[d] : 0u;
        __syncwarp();
        uint32_t rank = __popc(peers & lt_mask);
        if (ok && rank == 0) wh[d] = pre + __popc(peers);
        __syncwarp();
        r[it] = pre + rank;
    }
    (void)lt_mask;
    __syncthreads();
    {
        // thread t owns digit t: turn per-warp counts into per-warp start offsets
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            uint32_t c = whist[w * kRadix + threadIdx.x];
            whist[w * kRadix + threadIdx.x] = total;
            total += c;
        }
        uint32_t tile_total;
        uint32_t start = block_exclusive_scan(total, warp_sums, tile_total);
        dstart[threadIdx.x] = start;
        gofs[threadIdx.x] = offsets[(uint64_t)threadIdx.x * ntiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        uint64_t i = warp_base + it * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t)(k[it] >> shift) & (kRadix - 1);
            uint32_t pos = dstart[d] + wh[d] + r[it];
            skeys[pos] = k[it];
            svals[pos] = v[it];
        }
    }
    __syncthreads();
    uint64_t remaining = n - tile_base;
    uint32_t count = remaining < (uint64_t)TILE ? (uint32_t)remaining : (uint32_t)TILE;
    for (uint32_t j = threadIdx.x; j < count; j += kThreads) {
        KeyT key = skeys[j];
        uint32_t d = (uint32_t)(key >> shift) & (kRadix - 1);
        uint64_t dst = (uint64_t)gofs[d] + (j - dstart[d]);
        kout[dst] = key;
        vout[dst] = svals[j];
    }
}

// ---- single-sweep passes (query ordering) -------------------------------------------------------------
// The three-kernel pass above (per-tile histogram, scan of the [digit][tile] table, scatter) reads the
// keys twice and needs five launches.  For the query ordering the digit totals of ALL passes are counted
// once, by the kernel that makes the keys; a pass is then ONE kernel: tiles are taken in order through
// an atomic ticket, a tile publishes its digit counts, finds the counts of all earlier tiles by a
// decoupled look-back (Merrill & Garland's single-pass scan: a tile first publishes its own aggregate,
// so nothing it waits for ever waits for it), and scatters.  Ranking inside a warp uses shared-memory
// atomics (see STABLE = false above).  FIRST: the payload is the position itself (not read);
// LAST: only the payload is written.
constexpr uint32_t kStateAggregate = 1u << 30, kStatePrefix = 2u << 30, kStateCount = (1u << 30) - 1u;
#ifndef NBK_SWEEP_ITEMS
#define NBK_SWEEP_ITEMS 16
#endif
#ifndef NBK_SWEEP_MIN_BLOCKS
#define NBK_SWEEP_MIN_BLOCKS 4 // 64 registers: the pass is bound by load latency (ncu: 24 resident warps at 80 registers,
#endif                         // long-scoreboard stalls 5.4 per issue); measured 2.33 -> 2.17 ms per 10^8 queries; 8 or 12
                               // keys per thread instead of 16: 3.57 / 3.19 ms
constexpr int kSweepItems = NBK_SWEEP_ITEMS;
constexpr int kSweepTile = kThreads * kSweepItems;

inline uint64_t sweep_tiles(uint64_t n) { return div_up(n, (uint64_t)kSweepTile); }

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(kThreads, NBK_SWEEP_MIN_BLOCKS)
sweep_kernel(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint32_t *__restrict__ kout,
             uint32_t *__restrict__ vout, uint64_t n, int shift, const uint32_t *__restrict__ digit_totals,
             uint32_t *state, uint32_t *ticket) {
    constexpr int ITEMS = kSweepItems;
    constexpr int TILE = kSweepTile;
    __shared__ uint32_t skeys[TILE];
    __shared__ uint32_t svals[TILE];
    __shared__ uint32_t whist[kWarps * kRadix]; // per-warp digit counts, then warp start offsets
    __shared__ uint32_t dstart[kRadix];         // tile-local start of every digit run
    __shared__ uint32_t gofs[kRadix];           // global start of this tile's run of every digit
    __shared__ uint32_t warp_sums[kWarps];
    __shared__ uint32_t s_tile;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < kWarps * kRadix; i += kThreads) whist[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tile_base = (uint64_t)tile * TILE;
    const uint64_t warp_base = tile_base + (uint64_t)warp * (32 * ITEMS);
    uint32_t k[ITEMS], v[ITEMS], r[ITEMS];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const uint64_t i = warp_base + it * 32 + lane;
        const bool ok = i < n;
        k[it] = ok ? kin[i] : 0u;
        v[it] = FIRST ? (uint32_t)i : (ok ? vin[i] : 0u);
    }
    uint32_t *wh = whist + warp * kRadix;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const uint64_t i = warp_base + it * 32 + lane;
        r[it] = i < n ? atomicAdd(&wh[(k[it] >> shift) & (kRadix - 1)], 1u) : 0u;
    }
    __syncthreads();
    // thread t owns digit t
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const uint32_t c = whist[w * kRadix + threadIdx.x];
        whist[w * kRadix + threadIdx.x] = total;
        total += c;
    }
    volatile uint32_t *st = state;
    st[(uint64_t)tile * kRadix + threadIdx.x] = total | kStateAggregate; // published before anything is waited for
    // global start of the digit = digits below it (all tiles) + this digit in earlier tiles
    uint32_t all_tiles;
    const uint32_t below = block_exclusive_scan(digit_totals[threadIdx.x], warp_sums, all_tiles);
    {
        uint32_t tile_total;
        const uint32_t start = block_exclusive_scan(total, warp_sums, tile_total);
        dstart[threadIdx.x] = start;
    }
    __syncthreads();
    // the tile is brought into digit order in shared memory first: by the time the look-back starts, the
    // earlier tiles have had that long to publish their counts
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const uint64_t i = warp_base + it * 32 + lane;
        if (i < n) {
            const uint32_t d = (k[it] >> shift) & (kRadix - 1);
            const uint32_t pos = dstart[d] + wh[d] + r[it];
            skeys[pos] = k[it];
            svals[pos] = v[it];
        }
    }
    {
        uint32_t earlier = 0;
        for (int64_t prev = (int64_t)tile - 1; prev >= 0; --prev) {
            uint32_t seen, spins = 0;
            while (((seen = st[(uint64_t)prev * kRadix + threadIdx.x]) >> 30) == 0u)
                if (++spins > (1u << 26)) __trap(); // cannot happen (tickets are handed out in order); never hang
            earlier += seen & kStateCount;
            if ((seen >> 30) == 2u) break;
        }
        st[(uint64_t)tile * kRadix + threadIdx.x] = ((earlier + total) & kStateCount) | kStatePrefix;
        gofs[threadIdx.x] = below + earlier;
    }
    __syncthreads();
    const uint64_t remaining = n - tile_base;
    const uint32_t count = remaining < (uint64_t)TILE ? (uint32_t)remaining : (uint32_t)TILE;
    for (uint32_t j = threadIdx.x; j < count; j += kThreads) {
        const uint32_t key = skeys[j];
        const uint32_t d = (key >> shift) & (kRadix - 1);
        const uint64_t dst = (uint64_t)gofs[d] + (j - dstart[d]);
        if (!LAST) kout[dst] = key;
        vout[dst] = svals[j];
    }
}

// uint32 entries of workspace: digit totals [passes][256], tickets [passes], tile states [passes][tiles][256]
inline uint64_t sweep_workspace_entries(uint64_t n, int passes) {
    return (uint64_t)passes * kRadix + 32 + (uint64_t)passes * sweep_tiles(n) * kRadix;
}

struct SweepWorkspace {
    uint32_t *digit_totals; // filled by the caller's key kernel: [pass][256]
    uint32_t *tickets;
    uint32_t *states;
};

inline SweepWorkspace sweep_workspace(uint32_t *work, uint64_t n, int passes, cudaStream_t stream) {
    SweepWorkspace w{work, work + (uint64_t)passes * kRadix, work + (uint64_t)passes * kRadix + 32};
    NBK_CUDA(cudaMemsetAsync(work, 0, sweep_workspace_entries(n, passes) * 4, stream));
    return w;
}

// Orders (keys_a, positions) by key bits [begin_bit, begin_bit + 8 * passes); the digit totals must be in
// place.  Returns 0 if the ordered positions are in vals_a, 1 if in vals_b (the keys are not kept).
inline int sweep_order(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, uint64_t n,
                       int begin_bit, int passes, SweepWorkspace const &w, cudaStream_t stream) {
    if (n == 0) return 0;
    const unsigned tiles = (unsigned)sweep_tiles(n);
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        uint32_t *kin = cur ? keys_b : keys_a, *kout = cur ? keys_a : keys_b;
        uint32_t *vin = cur ? vals_b : vals_a, *vout = cur ? vals_a : vals_b;
        const int shift = begin_bit + 8 * p;
        const uint32_t *totals = w.digit_totals + p * kRadix;
        uint32_t *state = w.states + (uint64_t)p * tiles * kRadix;
        const bool first = p == 0, last = p == passes - 1;
        if (first && last) sweep_kernel<true, true><<<tiles, kThreads, 0, stream>>>(kin, vin, kout, vout, n, shift, totals, state, w.tickets + p);
        else if (first) sweep_kernel<true, false><<<tiles, kThreads, 0, stream>>>(kin, vin, kout, vout, n, shift, totals, state, w.tickets + p);
        else if (last) sweep_kernel<false, true><<<tiles, kThreads, 0, stream>>>(kin, vin, kout, vout, n, shift, totals, state, w.tickets + p);
        else sweep_kernel<false, false><<<tiles, kThreads, 0, stream>>>(kin, vin, kout, vout, n, shift, totals, state, w.tickets + p);
        NBK_LAUNCHED();
        cur ^= 1;
    }
    return cur;
}

template <typename KeyT> inline uint64_t sort_tiles(uint64_t n) {
    return div_up(n, (uint64_t)kThreads * Items<KeyT>::value);
}

// uint32 entries of workspace needed to sort n pairs
template <typename KeyT> inline uint64_t sort_workspace_entries(uint64_t n) {
    uint64_t table = sort_tiles<KeyT>(n) * kRadix;
    return table + scan_workspace_entries(table);
}

// Sorts by key bits [begin_bit, end_bit).  Buffers ping-pong; returns 0 if the result is in
// (keys_a, vals_a), 1 if it is in (keys_b, vals_b).
template <typename KeyT, bool STABLE = true>
inline int sort_pairs(KeyT *keys_a, uint32_t *vals_a, KeyT *keys_b, uint32_t *vals_b, uint64_t n,
                      int begin_bit, int end_bit, uint32_t *work, cudaStream_t stream) {
    if (n == 0) return 0;
    uint64_t ntiles = sort_tiles<KeyT>(n);
    uint64_t table = ntiles * kRadix;
    int cur = 0;
    for (int shift = begin_bit; shift < end_bit; shift += kRadixBits) {
        KeyT *kin = cur ? keys_b : keys_a, *kout = cur ? keys_a : keys_b;
        uint32_t *vin = cur ? vals_b : vals_a, *vout = cur ? vals_a : vals_b;
        hist_kernel<KeyT><<<(unsigned)ntiles, kThreads, 0, stream>>>(kin, n, shift, work,
                                                                      (uint32_t)ntiles);
        NBK_LAUNCHED();
        exclusive_scan(work, table, work + table, stream);
        scatter_kernel<KeyT, STABLE><<<(unsigned)ntiles, kThreads, 0, stream>>>(
            kin, vin, kout, vout, n, shift, work, (uint32_t)ntiles);
        NBK_LAUNCHED();
        cur ^= 1;
    }
    return cur;
}

} // namespace rs
} // namespace nbk
