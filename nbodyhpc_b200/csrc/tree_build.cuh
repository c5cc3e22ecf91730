// kd-tree construction on the device.
//
// Replaces the reference's recursive builder + Floyd-Rivest/AVX2 selection
// (kdtree_impl.hpp:94-157, kdtree_selection.cpp:34-200,322-368,475-494) with a level-synchronous
// build.  The topology is a function of point COUNTS only (leaf iff count <= max(leaf, 2*block);
// median_offset = ((count/2)/block)*block; kdtree_impl.hpp:91,101-110), so the host lays out every
// node, segment boundary and split position up front (plan_topology) and the device only has to
// order the points: per level one segmented radix sort by (segment, coordinate[level % 3]) moving a
// permutation, after which the element at each segment's median position IS the rank-median_offset
// element the reference selects, and its coordinate is the node's split.
#pragma once

#include <chrono>
#include <vector>

#include "common.cuh"
#include "radix_sort.cuh"
#include "tree_bottom.cuh"
#include "tree_topdown.cuh"

namespace nbk {

// ---- host-side topology plan ---------------------------------------------------------------------
struct LevelPlan {
    // indexed by segment id at this level (root = 0, children of s are 2s and 2s+1)
    std::vector<uint32_t> mid;   // absolute position of the median element; UINT32_MAX = no split
    std::vector<int32_t> node;   // node index receiving the split; -1 = no split
};

struct TopologyPlan {
    std::vector<nbk_node> nodes; // pre-order, left subtree first; splits filled by the device
    std::vector<LevelPlan> levels;
};

inline uint32_t plan_node(TopologyPlan &plan, uint64_t leaf, uint64_t block, int level,
                          uint64_t seg, int dim, uint32_t left, uint32_t count) {
    uint32_t me = (uint32_t)plan.nodes.size();
    if (count <= leaf) { // kdtree_impl.hpp:101-105
        plan.nodes.push_back(nbk_node{-1, 0.0f, left, left + count});
        return me;
    }
    uint32_t median = (uint32_t)((count / 2 / block) * block); // kdtree_impl.hpp:108-110
    plan.nodes.push_back(nbk_node{dim, 0.0f, 0u, 0u});
    if ((int)plan.levels.size() <= level) plan.levels.resize(level + 1);
    LevelPlan &lp = plan.levels[level];
    if (lp.mid.empty()) {
        lp.mid.assign((size_t)1 << level, 0xFFFFFFFFu);
        lp.node.assign((size_t)1 << level, -1);
    }
    lp.mid[seg] = left + median;
    lp.node[seg] = (int32_t)me;
    uint32_t l = plan_node(plan, leaf, block, level + 1, 2 * seg, (dim + 1) % 3, left, median);
    uint32_t r = plan_node(plan, leaf, block, level + 1, 2 * seg + 1, (dim + 1) % 3,
                           left + median, count - median);
    plan.nodes[me].left = l;
    plan.nodes[me].right = r;
    return me;
}

inline TopologyPlan plan_topology(uint64_t n_padded, int leaf_size, int block_size) {
    TopologyPlan plan;
    uint64_t leaf = (uint64_t)std::max(leaf_size, 2 * block_size); // kdtree_impl.hpp:91
    plan_node(plan, leaf, (uint64_t)block_size, 0, 0, 0, 0u, (uint32_t)n_padded);
    return plan;
}

// ---- kernels -------------------------------------------------------------------------------------

// Block-wide min/max of the orderable coordinates of the real points, then ONE atomic per block and
// value (a per-warp atomic on six addresses serialises the whole grid on them).
__device__ __forceinline__ void block_bounds(uint32_t (&lo)[3], uint32_t (&hi)[3], uint32_t *bounds6) {
    __shared__ uint32_t s_b[6][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], s));
            hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], s));
        }
        if (lane == 0) {
            s_b[d][warp] = lo[d];
            s_b[3 + d][warp] = hi[d];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const bool is_min = threadIdx.x < 3;
        uint32_t v = is_min ? 0xFFFFFFFFu : 0u;
        for (int w = 0; w < 8; ++w) v = is_min ? min(v, s_b[threadIdx.x][w]) : max(v, s_b[threadIdx.x][w]);
        if (is_min ? v != 0xFFFFFFFFu : v != 0u) {
            if (is_min) atomicMin(&bounds6[threadIdx.x], v);
            else atomicMax(&bounds6[threadIdx.x], v);
        }
    }
}

// AoS (n,3) -> SoA columns padded with FLT_MAX to n_padded, identity permutation
// (pybind.cpp:14-56).  flags[0] |= 1 if a coordinate is outside [0, box] (periodic only);
// bounds6 = {min x,y,z, max x,y,z} of the real points as orderable uint32.  Grid-stride.
__global__ void __launch_bounds__(256)
ingest_aos_kernel(const float *__restrict__ aos, uint64_t n, uint64_t n_padded,
                  float *__restrict__ x, float *__restrict__ y, float *__restrict__ z,
                  uint32_t *__restrict__ perm, int periodic, float box, uint32_t *flags,
                  uint32_t *bounds6) {
    constexpr int kTile = 1024; // points per iteration: 12 independent loads per thread in flight
    __shared__ float tile[kTile * 3];
    uint32_t lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
    bool bad = false;
    for (uint64_t base = (uint64_t)blockIdx.x * kTile; base < n_padded; base += (uint64_t)gridDim.x * kTile) {
        const uint64_t in_tile = n > base ? (n - base < (uint64_t)kTile ? n - base : (uint64_t)kTile) : 0;
        float v[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const uint32_t f = j * 256 + threadIdx.x;
            v[j] = f < in_tile * 3 ? __ldcs(aos + base * 3 + f) : FLT_MAX;
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) tile[j * 256 + threadIdx.x] = v[j];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t p = j * 256 + threadIdx.x;
            const uint64_t i = base + p;
            const bool real = i < n;
            // padding points are (FLT_MAX, FLT_MAX, FLT_MAX) (pybind.cpp:23-33): the tile was filled with it
            const float px = tile[p * 3], py = tile[p * 3 + 1], pz = tile[p * 3 + 2];
            if (i < n_padded) {
                x[i] = px;
                y[i] = py;
                z[i] = pz;
                perm[i] = (uint32_t)i;
            }
            bad = bad || (real && periodic &&
                          !(px >= 0.0f && px <= box && py >= 0.0f && py <= box && pz >= 0.0f && pz <= box));
            if (real) {
                const float c[3] = {px, py, pz};
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const uint32_t o = float_to_ordered(__float_as_uint(c[d]));
                    lo[d] = min(lo[d], o);
                    hi[d] = max(hi[d], o);
                }
            }
        }
        __syncthreads();
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flags, 1u);
    block_bounds(lo, hi, bounds6);
}

// Bounding box + identity permutation for a tree handed over as SoA columns.  Grid-stride.
__global__ void __launch_bounds__(256)
scan_soa_kernel(const float *__restrict__ x, const float *__restrict__ y,
                const float *__restrict__ z, uint64_t n, uint32_t *__restrict__ perm,
                uint32_t *bounds6) {
    uint32_t lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256) {
        const float p[3] = {x[i], y[i], z[i]};
        perm[i] = (uint32_t)i;
        // padding points (FLT_MAX) do not belong to the bounding box
        if (!(p[0] == FLT_MAX && p[1] == FLT_MAX && p[2] == FLT_MAX)) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const uint32_t o = float_to_ordered(__float_as_uint(p[d]));
                lo[d] = min(lo[d], o);
                hi[d] = max(hi[d], o);
            }
        }
    }
    block_bounds(lo, hi, bounds6);
}

// key[i] = (segment(i) << 32) | orderable(coord[perm[i]]).  seg_prev/mid_prev describe the level
// above: an element of segment s goes to 2s (left of the median position) or 2s+1.
__global__ void __launch_bounds__(256)
make_keys_kernel(const float *__restrict__ coord, const uint32_t *__restrict__ perm,
                 uint32_t *__restrict__ seg, const uint32_t *__restrict__ mid_prev, int level,
                 uint64_t n, uint64_t *__restrict__ keys) {
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint32_t s = 0;
    if (level > 0) {
        uint32_t sp = seg[i];
        s = 2 * sp + ((uint32_t)i >= mid_prev[sp] ? 1u : 0u);
    }
    seg[i] = s;
    uint32_t o = float_to_ordered(__float_as_uint(coord[perm[i]]));
    keys[i] = ((uint64_t)s << 32) | o;
}

// nodes[node[s]].split = coordinate of the element now sitting at mid[s]
__global__ void __launch_bounds__(256)
record_splits_kernel(const uint64_t *__restrict__ sorted_keys, const uint32_t *__restrict__ mid,
                     const int32_t *__restrict__ node, uint32_t nseg, nbk_node *__restrict__ nodes) {
    uint32_t s = blockIdx.x * 256 + threadIdx.x;
    if (s >= nseg) return;
    int32_t nd = node[s];
    if (nd < 0) return;
    uint32_t o = (uint32_t)sorted_keys[mid[s]];
    nodes[nd].split = __uint_as_float(ordered_to_float(o));
}

// Final placement: point at leaf-order position i goes into tile i/8, slot i%8 of the arena's
// 128-byte tiles {x[8], y[8], z[8], idx[8]} (see knn_query.cuh).
__global__ void __launch_bounds__(256)
gather_points_kernel(const float *__restrict__ x0, const float *__restrict__ y0,
                     const float *__restrict__ z0, const uint32_t *__restrict__ idx0,
                     const uint32_t *__restrict__ perm, uint64_t n, float *__restrict__ tiles) {
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint32_t p = perm[i];
    float *f = tiles + (i >> 3) * 32 + (i & 7);
    f[0] = x0[p];
    f[8] = y0[p];
    f[16] = z0[p];
    f[24] = __uint_as_float(idx0 ? idx0[p] : p);
}

// ---- driver --------------------------------------------------------------------------------------
struct TreeArena {
    nbk_node *nodes;
    float *tiles; // n_padded / 8 tiles of 32 floats
};

// Select-and-partition build (tree_topdown.cuh + tree_bottom.cuh): orders the points held in
// (x0,y0,z0[,idx0]) into the arena and fills the node array.  perm must hold the identity on entry;
// the four columns are used as one of the two ping-pong buffers.  Everything is enqueued on `stream`
// without host synchronisation except the final error check.
inline uint64_t build_select(td::TopPlan const &plan, uint64_t n_padded, int leaf_size, int block_size,
                         float *x0, float *y0, float *z0, uint32_t *perm, const uint32_t *idx0,
                         const uint32_t *d_bounds6, TreeArena const &arena, cudaStream_t stream) {
    using namespace td;
    if (n_padded == 0) {
        const nbk_node leaf{-1, 0.0f, 0u, 0u};
        NBK_CUDA(cudaMemcpyAsync(arena.nodes, &leaf, sizeof leaf, cudaMemcpyHostToDevice, stream));
        NBK_CUDA(cudaStreamSynchronize(stream));
        return 0;
    }
    static const uint32_t nb = [] {
        const char *v = std::getenv("NBK_BUILD_BINS");
        uint32_t n = v ? (uint32_t)std::atoi(v) : 0u;
        return (n >= 32 && n <= 2048 && (n & (n - 1)) == 0) ? n : 256u;
    }();
    const int T = plan.top_levels;
    // NBK_BUILD_TRACE=1: per-phase device times (CUDA events) and host wall clock on stderr
    static const bool trace = [] {
        const char *v = std::getenv("NBK_BUILD_TRACE");
        return v && v[0] == '1';
    }();
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    const auto host0 = std::chrono::steady_clock::now();
    if (trace) {
        for (auto &e : ev) NBK_CUDA(cudaEventCreate(&e));
        NBK_CUDA(cudaEventRecord(ev[0], stream));
    }
    Scratch scratch(stream);
    {
        const uint64_t max_segs = T > 0 ? (uint64_t)1 << (T - 1) : 0;
        uint64_t bytes = Scratch::padded(plan.segs.size() * sizeof(Seg)) + Scratch::padded(plan.lut.size() * 4) + 256;
        if (T > 0)
            bytes += 4 * Scratch::padded(n_padded * 4) + Scratch::padded(n_padded * 8) +
                     2 * Scratch::padded(max_segs * nb * 4) + Scratch::padded(max_segs * sizeof(Sel)) +
                     Scratch::padded(max_segs * 8) + Scratch::padded(3 * max_segs * 4) +
                     2 * Scratch::padded(4 * max_segs * sizeof(float4));
        scratch.reserve(bytes);
    }
    Seg *d_segs = scratch.get<Seg>(plan.segs.size());
    uint32_t *d_lut = scratch.get<uint32_t>(plan.lut.size());
    uint32_t *d_err = scratch.get<uint32_t>(1);
    NBK_CUDA(cudaMemcpyAsync(d_segs, plan.segs.data(), plan.segs.size() * sizeof(Seg),
                             cudaMemcpyHostToDevice, stream));
    NBK_CUDA(cudaMemcpyAsync(d_lut, plan.lut.data(), plan.lut.size() * 4, cudaMemcpyHostToDevice, stream));
    NBK_CUDA(cudaMemsetAsync(d_err, 0, 4, stream));
    Columns cur{x0, y0, z0, perm}, alt{nullptr, nullptr, nullptr, nullptr};
    if (T > 0) {
        alt.x = scratch.get<float>(n_padded);
        alt.y = scratch.get<float>(n_padded);
        alt.z = scratch.get<float>(n_padded);
        alt.id = scratch.get<uint32_t>(n_padded);
        const uint64_t max_segs = (uint64_t)1 << (T - 1);
        uint32_t *hist_cur = scratch.get<uint32_t>(max_segs * nb);
        uint32_t *hist_alt = scratch.get<uint32_t>(max_segs * nb);
        Sel *sel = scratch.get<Sel>(max_segs);
        unsigned long long *pivot = scratch.get<unsigned long long>(max_segs);
        uint32_t *cursors = scratch.get<uint32_t>(3 * max_segs);
        float4 *bounds_cur = scratch.get<float4>(4 * max_segs);
        float4 *bounds_alt = scratch.get<float4>(4 * max_segs);
        unsigned long long *cand = scratch.get<unsigned long long>(n_padded);
        NBK_CUDA(cudaMemsetAsync(hist_cur, 0, max_segs * nb * 4, stream));
        NBK_CUDA(cudaMemsetAsync(hist_alt, 0, max_segs * nb * 4, stream));
        init_bounds_kernel<<<1, 32, 0, stream>>>(d_bounds6, bounds_cur);
        NBK_LAUNCHED();
        const uint32_t chunks0 = (uint32_t)div_up(plan.biggest[0], kChunk);
        hist_kernel<<<chunks0, kChunkThreads, nb * 4, stream>>>(cur.x, d_segs, bounds_cur, 0, chunks0, nb,
                                                                hist_cur);
        NBK_LAUNCHED();
        for (int l = 0; l < T; ++l) {
            const uint32_t nseg = 1u << l;
            const Seg *segs_l = d_segs + (nseg - 1);
            const Seg *segs_next = d_segs + (2 * nseg - 1);
            const int dim = l % 3;
            const uint32_t mc = (uint32_t)div_up(plan.biggest[l], kChunk);
            const unsigned grid = nseg * mc;
            const float *col = dim == 0 ? cur.x : (dim == 1 ? cur.y : cur.z);
            bucket_kernel<<<(unsigned)div_up((uint64_t)nseg * 32, 256), 256, 0, stream>>>(segs_l, nseg, nb,
                                                                                       hist_cur, sel, cursors);
            NBK_LAUNCHED();
            const uint32_t spans = (uint32_t)div_up(plan.biggest[l], kCompactSpan);
            compact_kernel<<<nseg * spans, kCompactThreads, 0, stream>>>(col, cur.id, segs_l, bounds_cur, sel,
                                                                       dim, spans, nb, cursors, cand);
            NBK_LAUNCHED();
            const uint64_t expect = plan.biggest[l] / nb; // candidates per segment for uniform data
            const unsigned sel_threads = expect > 2048 ? 1024u : (expect > 128 ? 256u : 128u);
            select_kernel<<<nseg, sel_threads, 0, stream>>>(segs_l, sel, cand, dim, bounds_cur, bounds_alt, pivot,
                                                            arena.nodes);
            NBK_LAUNCHED();
            if (l + 1 < T) {
                partition_kernel<true><<<grid, kChunkThreads, 2 * nb * 4, stream>>>(
                    cur, alt, segs_l, pivot, cursors, dim, mc, segs_next, bounds_cur, nb, hist_alt);
            } else {
                partition_kernel<false><<<grid, kChunkThreads, 0, stream>>>(
                    cur, alt, segs_l, pivot, cursors, dim, mc, segs_next, bounds_cur, nb, hist_alt);
            }
            NBK_LAUNCHED();
            std::swap(cur, alt);
            std::swap(hist_cur, hist_alt);
            std::swap(bounds_cur, bounds_alt);
        }
    }
    if (trace) NBK_CUDA(cudaEventRecord(ev[1], stream));
    BottomArgs args;
    args.x = cur.x;
    args.y = cur.y;
    args.z = cur.z;
    args.id = cur.id;
    args.idx0 = idx0;
    args.segs = d_segs + (((size_t)1 << T) - 1);
    args.nseg = 1u << T;
    args.level = T;
    args.leaf = (uint32_t)std::max(leaf_size, 2 * block_size);
    args.block = (uint32_t)block_size;
    args.lut = d_lut;
    args.nodes = arena.nodes;
    args.tiles = arena.tiles;
    args.error = d_err;
    // function attributes are per device: set it on whichever device this build runs on
    NBK_CUDA(cudaFuncSetAttribute(bottom_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBottomSmem));
    bottom_kernel<<<args.nseg, kBotThreads, kBottomSmem, stream>>>(args);
    NBK_LAUNCHED();
    if (trace) NBK_CUDA(cudaEventRecord(ev[2], stream));
    const auto host1 = std::chrono::steady_clock::now();
    uint32_t err = 0;
    NBK_CUDA(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, stream));
    NBK_CUDA(cudaStreamSynchronize(stream)); // also keeps the plan's host tables alive long enough
    if (trace) {
        float top_ms = 0, bottom_ms = 0;
        cudaEventElapsedTime(&top_ms, ev[0], ev[1]);
        cudaEventElapsedTime(&bottom_ms, ev[1], ev[2]);
        const double enq = std::chrono::duration<double, std::milli>(host1 - host0).count();
        const double all = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host0).count();
        fprintf(stderr, "[nbk build] n=%llu top_levels=%d: top %.3f ms, bottom %.3f ms (device); host enqueue %.3f ms, "
                        "host total %.3f ms; scratch %.1f MB in %d allocation(s)\n", (unsigned long long)n_padded, T,
                top_ms, bottom_ms, enq, all, scratch.total_bytes / 1e6, scratch.count);
        for (auto &e : ev) cudaEventDestroy(e);
    }
    if (err) throw Error(NBK_ERR_CUDA, "kd-tree build: sub-tree deeper than the bottom kernel's tables");
    return scratch.total_bytes;
}

// Sort-based build (NBK_BUILD=sort; the first implementation, kept as the cross-check of the
// select-and-partition build).  Orders the points held in (x0,y0,z0[,idx0]) into the arena and fills
// the node array.  perm must hold the identity on entry.
inline void build_levels(TopologyPlan const &plan, uint64_t n_padded, const float *x0,
                         const float *y0, const float *z0, const uint32_t *idx0, uint32_t *perm,
                         TreeArena const &arena, cudaStream_t stream) {
    const int n_levels = (int)plan.levels.size();
    NBK_CUDA(cudaMemcpyAsync(arena.nodes, plan.nodes.data(), plan.nodes.size() * sizeof(nbk_node),
                             cudaMemcpyHostToDevice, stream));
    const unsigned grid = (unsigned)div_up(n_padded ? n_padded : 1, 256);
    if (n_levels > 0) {
        Scratch scratch(stream);
        uint64_t *keys_a = scratch.get<uint64_t>(n_padded);
        uint64_t *keys_b = scratch.get<uint64_t>(n_padded);
        uint32_t *perm_b = scratch.get<uint32_t>(n_padded);
        uint32_t *seg = scratch.get<uint32_t>(n_padded);
        uint32_t *work = scratch.get<uint32_t>(rs::sort_workspace_entries<uint64_t>(n_padded));
        // per-level segment tables, one upload
        uint64_t table_entries = 0;
        for (auto const &lp : plan.levels) table_entries += lp.mid.size();
        std::vector<uint32_t> h_mid(table_entries);
        std::vector<int32_t> h_node(table_entries);
        std::vector<uint64_t> level_ofs(n_levels);
        uint64_t ofs = 0;
        for (int l = 0; l < n_levels; ++l) {
            level_ofs[l] = ofs;
            std::copy(plan.levels[l].mid.begin(), plan.levels[l].mid.end(), h_mid.begin() + ofs);
            std::copy(plan.levels[l].node.begin(), plan.levels[l].node.end(), h_node.begin() + ofs);
            ofs += plan.levels[l].mid.size();
        }
        uint32_t *d_mid = scratch.get<uint32_t>(table_entries);
        int32_t *d_node = scratch.get<int32_t>(table_entries);
        NBK_CUDA(cudaMemcpyAsync(d_mid, h_mid.data(), table_entries * 4, cudaMemcpyHostToDevice,
                                 stream));
        NBK_CUDA(cudaMemcpyAsync(d_node, h_node.data(), table_entries * 4, cudaMemcpyHostToDevice,
                                 stream));
        uint32_t *perm_cur = perm, *perm_alt = perm_b;
        for (int l = 0; l < n_levels; ++l) {
            const float *coord = (l % 3 == 0) ? x0 : ((l % 3 == 1) ? y0 : z0);
            make_keys_kernel<<<grid, 256, 0, stream>>>(
                coord, perm_cur, seg, l > 0 ? d_mid + level_ofs[l - 1] : nullptr, l, n_padded,
                keys_a);
            NBK_LAUNCHED();
            int where = rs::sort_pairs<uint64_t>(keys_a, perm_cur, keys_b, perm_alt, n_padded, 0,
                                                 32 + l, work, stream);
            const uint64_t *sorted = where ? keys_b : keys_a;
            if (where) std::swap(perm_cur, perm_alt);
            uint32_t nseg = (uint32_t)plan.levels[l].mid.size();
            record_splits_kernel<<<(unsigned)div_up(nseg, 256), 256, 0, stream>>>(
                sorted, d_mid + level_ofs[l], d_node + level_ofs[l], nseg, arena.nodes);
            NBK_LAUNCHED();
        }
        gather_points_kernel<<<grid, 256, 0, stream>>>(x0, y0, z0, idx0, perm_cur, n_padded,
                                                       arena.tiles);
        NBK_LAUNCHED();
        // the host tables must outlive the async copies
        NBK_CUDA(cudaStreamSynchronize(stream));
    } else {
        if (n_padded) {
            gather_points_kernel<<<grid, 256, 0, stream>>>(x0, y0, z0, idx0, perm, n_padded,
                                                           arena.tiles);
            NBK_LAUNCHED();
        }
        NBK_CUDA(cudaStreamSynchronize(stream));
    }
}

} // namespace nbk
