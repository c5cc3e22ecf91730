// libnbk.so -- C ABI (include/nbk.h) over the sm_100a build and query kernels.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "host_stage.cuh"
#include "knn_query.cuh"
#include "radix_sort.cuh"
#include "tree_build.cuh"

namespace nbk {
std::atomic<uint64_t> g_launches{0};
thread_local std::string g_error;

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int device) {
        NBK_CUDA(cudaGetDevice(&prev));
        if (device >= 0 && device != prev) NBK_CUDA(cudaSetDevice(device));
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

constexpr uint64_t kAlign = 256;

struct ArenaLayout {
    uint64_t nodes, tiles, total;
};

// [ nodes: n_nodes x 16 B | tiles: n_padded/8 x 128 B {x[8], y[8], z[8], idx[8]} ]
inline ArenaLayout arena_layout(uint64_t n_padded, uint64_t n_nodes) {
    ArenaLayout l;
    l.nodes = 0;
    l.tiles = align_up(n_nodes * sizeof(nbk_node), kAlign);
    l.total = l.tiles + align_up(std::max<uint64_t>(n_padded, 8) * 16, kAlign);
    return l;
}
} // namespace nbk

struct nbk_tree {
    nbk_tree_meta meta{};
    int device = 0;
    char *arena = nullptr;
    bool arena_cached = false; // from the library's block cache (big trees) instead of the stream-ordered pool
    nbk::TreeArena view{};
    // streams that have read this tree since it was built, with an event marking their last use: the
    // arena is released behind them instead of behind a device-wide synchronisation
    mutable std::mutex use_mutex;
    mutable std::vector<std::pair<cudaStream_t, cudaEvent_t>> uses;

    nbk_tree() = default;
    nbk_tree(nbk_tree const &) = delete;
    ~nbk_tree() {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(device);
        for (auto &u : uses) {
            cudaStreamWaitEvent(nullptr, u.second, 0);
            cudaEventDestroy(u.second);
        }
        if (arena && arena_cached) {
            // back to the block cache; whoever takes the block next waits for the readers collected above
            cudaEvent_t ready = nullptr;
            if (cudaEventCreateWithFlags(&ready, cudaEventDisableTiming) == cudaSuccess) cudaEventRecord(ready, nullptr);
            else cudaStreamSynchronize(nullptr);
            nbk::BlockCache::release(arena, ready);
        } else if (arena) {
            cudaFreeAsync(arena, nullptr);
        }
        if (prev >= 0) cudaSetDevice(prev);
    }
    // called after work reading the tree has been enqueued on `stream`
    void mark_use(cudaStream_t stream) const {
        if (stream == nullptr || stream == cudaStreamLegacy) return; // the release itself is ordered on this stream
        std::lock_guard<std::mutex> lock(use_mutex);
        for (auto &u : uses)
            if (u.first == stream) {
                cudaEventRecord(u.second, stream);
                return;
            }
        if (uses.size() >= 64) {
            // a caller cycling through short-lived streams: settle the old ones instead of growing
            for (auto &u : uses) {
                cudaEventSynchronize(u.second);
                cudaEventDestroy(u.second);
            }
            uses.clear();
        }
        cudaEvent_t ev = nullptr;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return;
        cudaEventRecord(ev, stream);
        uses.emplace_back(stream, ev);
    }

    void bind() {
        nbk::ArenaLayout l = nbk::arena_layout(meta.n_padded, meta.n_nodes);
        view.nodes = reinterpret_cast<nbk_node *>(arena + l.nodes);
        view.tiles = reinterpret_cast<float *>(arena + l.tiles);
    }
    // periodic < 0: the tree's own metric
    nbk::QueryTree query_view(int periodic = -1, float box_size = 0.0f) const {
        if (periodic < 0) {
            periodic = meta.periodic;
            box_size = meta.box_size;
        }
        nbk::QueryTree q;
        q.nodes = view.nodes;
        q.tiles = reinterpret_cast<const float4 *>(view.tiles);
        q.periodic = periodic != 0;
        q.box = periodic ? box_size : 0.0f;
        for (int d = 0; d < 3; ++d) {
            // initial_box: kdtree.hpp:51-61 (open) / :111-120 (periodic)
            q.lo[d] = periodic ? 0.0f : -FLT_MAX;
            q.hi[d] = periodic ? box_size : FLT_MAX;
        }
        return q;
    }
};

namespace nbk {

template <typename F> int guarded(F &&f) {
    try {
        f();
        return NBK_OK;
    } catch (Error const &e) {
        g_error = e.what();
        return e.code;
    } catch (std::bad_alloc const &) {
        g_error = "host allocation failed";
        return NBK_ERR_NOMEM;
    } catch (std::exception const &e) {
        g_error = e.what();
        return NBK_ERR_INVALID;
    }
}

void require_sm100(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw Error(NBK_ERR_CUDA, std::string("no CUDA device available (libnbk has no CPU fallback): ") +
                                      cudaGetErrorString(e));
    int dev = device;
    if (dev < 0) NBK_CUDA(cudaGetDevice(&dev));
    if (dev >= count) throw Error(NBK_ERR_INVALID, "device index out of range");
    int major = 0;
    NBK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10)
        throw Error(NBK_ERR_CUDA, "libnbk is built for sm_100a only (compute capability 10.x required)");
    // keep stream-ordered scratch memory cached in the pool instead of returning it at every sync
    static std::atomic<uint64_t> pool_ready{0};
    if (dev < 64 && !(pool_ready.load() & (1ull << dev))) {
        cudaMemPool_t pool;
        NBK_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t threshold = UINT64_MAX;
        NBK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
        pool_ready.fetch_or(1ull << dev);
    }
}

void check_build_args(uint64_t n_padded, int block_size, bool soa) {
    // kdtree.cpp:98-108, same order and wording
    if (n_padded > 0xFFFFFFFFull) throw Error(NBK_ERR_INVALID, "More than uint32_t points are not supported.");
    if (block_size <= 0 || block_size % 8 != 0) throw Error(NBK_ERR_INVALID, "block_size must be a multiple of 8.");
    if (soa && n_padded % (uint64_t)block_size != 0)
        throw Error(NBK_ERR_INVALID, "block_size must divide the number of points.");
}

// Small arenas come from the stream-ordered pool, allocated on the stream that fills them; big ones from the
// library's block cache (cudaMalloc once, reused by the next tree of that size): growing the pool by a
// 17.7 GB arena took 2.9 s on the first 1024^3 build, 20 ms for the 2.2 GB of 512^3, cudaMalloc takes 3 ms.
// Either way a tree freed and rebuilt reuses the same memory, and ~nbk_tree releases it behind every
// stream that used it.
std::unique_ptr<nbk_tree> alloc_tree(nbk_tree_meta const &meta, cudaStream_t stream) {
    auto t = std::make_unique<nbk_tree>();
    t->meta = meta;
    NBK_CUDA(cudaGetDevice(&t->device));
    ArenaLayout l = arena_layout(meta.n_padded, meta.n_nodes);
    t->meta.arena_bytes = l.total;
    if (l.total >= Scratch::kCacheFrom) {
        t->arena = static_cast<char *>(BlockCache::acquire(l.total, stream));
        t->arena_cached = true;
    } else {
        NBK_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&t->arena), l.total, stream));
    }
    t->mark_use(stream);
    t->bind();
    return t;
}

// Common tail of the three build entry points.  (x0,y0,z0[,idx0]) are device SoA columns of
// n_padded points, perm holds the identity, bounds6 the orderable bounding box.
std::unique_ptr<nbk_tree> finish_build(uint64_t n, uint64_t n_padded, int leaf_size, int block_size,
                                       int periodic, float box_size, float *x0, float *y0, float *z0,
                                       const uint32_t *idx0, bool &trim_after,
                                       uint32_t *perm, const uint32_t *d_bounds6, int device,
                                       cudaStream_t stream) {
    // NBK_BUILD=sort selects the per-level radix-sort build (cross-check); default: select + partition
    static const bool sort_build = [] {
        const char *v = std::getenv("NBK_BUILD");
        return v && std::string(v) == "sort";
    }();
    const char *tv = std::getenv("NBK_BUILD_TRACE");
    const bool trace = tv && tv[0] == '1';
    const auto host0 = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host0).count(); };
    td::TopPlan top = td::plan_top(n_padded, leaf_size, block_size);
    if (trace) fprintf(stderr, "[nbk build]   finish_build +%.3f ms: host plan\n", since());
    nbk_tree_meta meta{};
    meta.n_points = n;
    meta.n_padded = n_padded;
    meta.n_nodes = top.n_nodes;
    meta.leaf_size = leaf_size;
    meta.block_size = block_size;
    meta.periodic = periodic ? 1 : 0;
    meta.box_size = periodic ? box_size : 0.0f;
    meta.n_levels = top.n_levels;
    (void)device;
    auto tree = alloc_tree(meta, stream);
    if (trace) fprintf(stderr, "[nbk build]   finish_build +%.3f ms: arena allocated\n", since());
    uint64_t scratch_bytes = n_padded * 16; // the caller's columns
    if (sort_build) {
        TopologyPlan plan = plan_topology(n_padded, leaf_size, block_size);
        build_levels(plan, n_padded, x0, y0, z0, idx0, perm, tree->view, stream);
        scratch_bytes += n_padded * 32;
    } else {
        scratch_bytes += build_select(top, n_padded, leaf_size, block_size, x0, y0, z0, perm, idx0, d_bounds6,
                                      tree->view, stream);
    }
    if (trace) fprintf(stderr, "[nbk build]   finish_build +%.3f ms: build_select returned\n", since());
    uint32_t b[6];
    NBK_CUDA(cudaMemcpyAsync(b, d_bounds6, sizeof b, cudaMemcpyDeviceToHost, stream));
    NBK_CUDA(cudaStreamSynchronize(stream));
    {
        // A big build's scratch (several times the tree) goes back to the device so that query outputs
        // can use it; a small one stays cached in the pool, which makes the next build allocation-free.
        // (the device's memory size is looked up once: cudaMemGetInfo walks the pools and cost ~9 ms per build)
        static std::atomic<uint64_t> total_mem[64];
        int dev_now = 0;
        NBK_CUDA(cudaGetDevice(&dev_now));
        uint64_t total_b = dev_now < 64 ? total_mem[dev_now].load() : 0;
        if (!total_b) {
            size_t free_b = 0, tb = 0;
            NBK_CUDA(cudaMemGetInfo(&free_b, &tb));
            total_b = tb;
            if (dev_now < 64) total_mem[dev_now].store(total_b);
        }
        // Handing tens of GB back costs ~0.3 s of cudaFree per build (1024^3: 43 GB), so it is done only when
        // the device is actually getting full: big scratch AND less than a quarter of the memory still free.
        trim_after = false;
        if (scratch_bytes > total_b / 8) {
            size_t free_b = 0, tb = 0;
            NBK_CUDA(cudaMemGetInfo(&free_b, &tb));
            trim_after = free_b < tb / 4;
        }
        if (trim_after) {
            cudaMemPool_t pool;
            int dev = 0;
            NBK_CUDA(cudaGetDevice(&dev));
            NBK_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
            NBK_CUDA(cudaMemPoolTrimTo(pool, 0));
        }
    }
    if (trace) fprintf(stderr, "[nbk build]   finish_build +%.3f ms: pool policy done\n", since());
    for (int d = 0; d < 3; ++d) {
        bool empty = b[d] == 0xFFFFFFFFu && b[3 + d] == 0u;
        uint32_t lo = ordered_to_float(b[d]), hi = ordered_to_float(b[3 + d]);
        std::memcpy(&tree->meta.lo[d], &lo, 4);
        std::memcpy(&tree->meta.hi[d], &hi, 4);
        if (empty) tree->meta.lo[d] = tree->meta.hi[d] = 0.0f;
    }
    return tree;
}

std::unique_ptr<nbk_tree> build_from_device_aos(const float *d_aos, uint64_t n, int leaf_size,
                                                int block_size, int periodic, float box_size,
                                                int device, cudaStream_t stream) {
    const auto host0 = std::chrono::steady_clock::now();
    struct Trace {
        std::chrono::steady_clock::time_point t0;
        ~Trace() {
            const char *v = std::getenv("NBK_BUILD_TRACE");
            if (v && v[0] == '1')
                fprintf(stderr, "[nbk build] whole call (ingest + plan + build + pool policy): %.3f ms host\n",
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        }
    } trace_whole{host0};
    check_build_args(n, block_size, false);
    uint64_t n_padded = div_up(n, block_size) * block_size; // pybind.cpp:23
    check_build_args(n_padded, block_size, false);
    bool trim_after = false;
    std::unique_ptr<nbk_tree> tree;
    const char *tv = std::getenv("NBK_BUILD_TRACE");
    const bool trace = tv && tv[0] == '1';
    auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host0).count(); };
    {
        Scratch scratch(stream);
        uint64_t cols = std::max<uint64_t>(n_padded, 1);
        scratch.reserve(4 * Scratch::padded(cols * 4) + 256);
        if (trace) fprintf(stderr, "[nbk build] +%.3f ms: column block reserved\n", since());
        float *x0 = scratch.get<float>(cols), *y0 = scratch.get<float>(cols), *z0 = scratch.get<float>(cols);
        uint32_t *perm = scratch.get<uint32_t>(cols);
        uint32_t *aux = scratch.get<uint32_t>(8); // [0] flags, [1..6] bounds
        const uint32_t init[8] = {0u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u, 0u};
        NBK_CUDA(cudaMemcpyAsync(aux, init, sizeof init, cudaMemcpyHostToDevice, stream));
        if (n_padded) {
            ingest_aos_kernel<<<(unsigned)std::min<uint64_t>(div_up(n_padded, 1024), 148 * 8), 256, 0, stream>>>(
                d_aos, n, n_padded, x0, y0, z0, perm, periodic, box_size, aux, aux + 1);
            NBK_LAUNCHED();
        }
        uint32_t flags = 0;
        NBK_CUDA(cudaMemcpyAsync(&flags, aux, 4, cudaMemcpyDeviceToHost, stream));
        NBK_CUDA(cudaStreamSynchronize(stream));
        if (trace) fprintf(stderr, "[nbk build] +%.3f ms: ingest done\n", since());
        if (flags & 1u) // pybind.cpp:42-46
            throw Error(NBK_ERR_INVALID, "When using periodic boundary conditions, all points must be "
                                         "within the box (0 <= x <= box_size).");
        tree = finish_build(n, n_padded, leaf_size, block_size, periodic, box_size, x0, y0, z0, nullptr,
                            trim_after, perm, aux + 1, device, stream);
    }
    if (trim_after) {
        int dev = 0;
        NBK_CUDA(cudaGetDevice(&dev));
        BlockCache::trim(dev);
    }
    return tree;
}

// ---- optional event timing of sections ------------------------------------------------------------
std::atomic<int> g_profile{0};
std::mutex g_profile_mutex;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_profile_events[NBK_SECTION_COUNT];

struct SectionTimer {
    int section;
    cudaStream_t stream;
    cudaEvent_t start = nullptr, stop = nullptr;
    SectionTimer(int sec, cudaStream_t s) : section(sec), stream(s) {
        if (!g_profile.load(std::memory_order_relaxed)) return;
        NBK_CUDA(cudaEventCreate(&start));
        NBK_CUDA(cudaEventCreate(&stop));
        NBK_CUDA(cudaEventRecord(start, stream));
    }
    void finish() {
        if (!start) return;
        NBK_CUDA(cudaEventRecord(stop, stream));
        std::lock_guard<std::mutex> lock(g_profile_mutex);
        g_profile_events[section].emplace_back(start, stop);
        start = nullptr;
    }
    ~SectionTimer() {
        if (start) {
            cudaEventDestroy(start);
            cudaEventDestroy(stop);
        }
    }
};

// ---- query -----------------------------------------------------------------------------------------
// NBK_KERNEL=packet selects the warp-packet traversal kernel; default is the per-lane kernel.
inline bool use_packet_kernel() {
    static const bool packet = [] {
        const char *v = std::getenv("NBK_KERNEL");
        return v && std::string(v) == "packet";
    }();
    return packet;
}

constexpr int kMaxSharedK = 64; // above: heaps in global memory
// NBK_MAX_SHARED_K (8..64): experiments with the global-memory heap for smaller k
inline int max_shared_k() {
    static const int v = [] {
        const char *e = std::getenv("NBK_MAX_SHARED_K");
        const int n = e ? std::atoi(e) : kMaxSharedK;
        return (n >= 8 && n <= kMaxSharedK) ? n : kMaxSharedK;
    }();
    return v;
}

// second (shifted images) pass: a persistent grid over the device-side work list
constexpr unsigned kImagesGridMax = 148 * 8;

template <typename Top> constexpr size_t lane_smem() {
    if (Top::kShared) return (size_t)Top::kSize * kQueryThreads * Top::kKeyBytes;
    if (kQueueCap > 0 && !Top::kGlobal) return (size_t)kQueueCap * kQueryThreads * sizeof(unsigned long long); // candidate queues
    if (kStageSmem > 0 && !Top::kGlobal) return kStageSmem; // NBK_STAGE_HOME: one staged leaf + one mbarrier per warp
    return 0;
}

// TopFast: container of the first pass; Top: container of the boundary pass (they differ for the kNN-CDF)
template <typename TopFast, typename Top, bool P>
void launch_lane(QueryTree const &qt, QueryBatch const &qb, DeferList defer, CdfArgs cdf, cudaStream_t stream) {
    unsigned grid = (unsigned)div_up(qb.m, kQueryThreads);
    constexpr size_t smem_fast = lane_smem<TopFast>(), smem = lane_smem<Top>();
    auto fast = knn_lane_kernel<TopFast, P, false>;
    auto general = knn_lane_kernel<Top, P, P>;
    if (smem_fast > 48 * 1024)
        NBK_CUDA(cudaFuncSetAttribute(fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast));
    if (smem > 48 * 1024)
        NBK_CUDA(cudaFuncSetAttribute(general, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fast<<<grid, kQueryThreads, smem_fast, stream>>>(qt, qb, defer, cdf);
    NBK_LAUNCHED();
    if (P) {
        // queries whose search ball reaches through a face of the box (usually ~1 %): the kernel reads
        // their number on the device and walks the list with a grid that fits the machine once
        general<<<std::min(grid, kImagesGridMax), kQueryThreads, smem, stream>>>(qt, qb, defer, cdf);
        NBK_LAUNCHED();
    }
}

template <int K, bool P>
void launch_knn(QueryTree const &qt, QueryBatch const &qb, DeferList defer, CdfArgs cdf, cudaStream_t stream) {
    if (use_packet_kernel()) {
        if (cdf.edges) throw Error(NBK_ERR_INVALID, "the kNN-CDF epilogue needs the default (lane) kernel");
        if constexpr (K > 0 && K <= 8) {
            unsigned grid = (unsigned)div_up(qb.m, kQueryThreads);
            knn_packet_kernel<K, P><<<grid, kQueryThreads, 0, stream>>>(qt, qb.q_aos, qb.order, qb.m, qb.k,
                                                                        (qb.flags & NBK_QUERY_SQUARED) != 0,
                                                                        qb.out_d, qb.out_i);
            NBK_LAUNCHED();
            return;
        }
    }
    // NBK_CDF_KEYS=64: the fused kNN-CDF keeps (d2, index) keys in its first pass too (the previous form)
    static const bool dist_only_cdf = [] {
        const char *v = std::getenv("NBK_CDF_KEYS");
        return !(v && std::string(v) == "64");
    }();
    const bool dist_only = cdf.edges != nullptr && dist_only_cdf;
    if constexpr (K == 0) {
        launch_lane<HeapT<0>, HeapT<0>, P>(qt, qb, defer, cdf, stream);
    } else if constexpr (K >= 16) {
        if (dist_only) launch_lane<DistHeapT<K>, HeapT<K>, P>(qt, qb, defer, cdf, stream);
        else launch_lane<HeapT<K>, HeapT<K>, P>(qt, qb, defer, cdf, stream);
    } else {
        if (dist_only) launch_lane<DistTopK<K>, TopK<K>, P>(qt, qb, defer, cdf, stream);
        else launch_lane<TopK<K>, TopK<K>, P>(qt, qb, defer, cdf, stream);
    }
}

template <bool P>
void dispatch_knn(QueryTree const &qt, QueryBatch const &qb, DeferList defer, CdfArgs cdf, cudaStream_t stream) {
    const int k = qb.k;
    if (k > max_shared_k()) {
        launch_knn<0, P>(qt, qb, defer, cdf, stream);
        return;
    }
    if (k <= 1) launch_knn<1, P>(qt, qb, defer, cdf, stream);
    else if (k <= 2) launch_knn<2, P>(qt, qb, defer, cdf, stream);
    else if (k <= 4) launch_knn<4, P>(qt, qb, defer, cdf, stream);
    else if (k <= 8) launch_knn<8, P>(qt, qb, defer, cdf, stream);
    else if (k <= 16) launch_knn<16, P>(qt, qb, defer, cdf, stream);
    else if (k <= 32) launch_knn<32, P>(qt, qb, defer, cdf, stream);
    else if (k <= 64) launch_knn<64, P>(qt, qb, defer, cdf, stream);
    else launch_knn<0, P>(qt, qb, defer, cdf, stream);
}

template <typename Top, bool P>
void launch_scan_block(QueryTree const &qt, uint32_t n, QueryBatch const &qb, cudaStream_t stream) {
    size_t smem = Top::kShared ? (size_t)Top::kSize * kQueryThreads * Top::kKeyBytes : 0;
    auto kern = scan_block_kernel<Top, P>;
    if (smem > 48 * 1024) NBK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)div_up(qb.m, kQueryThreads), kQueryThreads, smem, stream>>>(qt, n, qb);
    NBK_LAUNCHED();
}

// the same k -> container choice as dispatch_knn
template <bool P> void dispatch_scan_block(QueryTree const &qt, uint32_t n, QueryBatch const &qb, cudaStream_t stream) {
    const int k = qb.k;
    if (k <= 1) launch_scan_block<TopK<1>, P>(qt, n, qb, stream);
    else if (k <= 2) launch_scan_block<TopK<2>, P>(qt, n, qb, stream);
    else if (k <= 4) launch_scan_block<TopK<4>, P>(qt, n, qb, stream);
    else if (k <= 8) launch_scan_block<TopK<8>, P>(qt, n, qb, stream);
    else if (k <= 16) launch_scan_block<HeapT<16>, P>(qt, n, qb, stream);
    else if (k <= 32) launch_scan_block<HeapT<32>, P>(qt, n, qb, stream);
    else if (k <= 64) launch_scan_block<HeapT<64>, P>(qt, n, qb, stream);
    else launch_scan_block<HeapT<0>, P>(qt, n, qb, stream);
}

// k > 64: heaps of exactly k slots in a global scratch of `columns` heaps, answered batch by batch.
// The scratch is sized to stay L2-resident-ish (<= 256 MB) while a batch still fills the machine.
inline uint64_t global_heap_columns(int k) {
    const uint64_t by_bytes = (256ull << 20) / ((uint64_t)k * 8);
    const uint64_t cols = std::max<uint64_t>(148ull * 4 * kQueryThreads, std::min<uint64_t>(by_bytes, 1ull << 20));
    return align_up(cols, kQueryThreads);
}

void query_device(nbk_tree const &tree, const float *d_q, uint64_t m, int k, float *d_out_d,
                  uint32_t *d_out_i, cudaStream_t stream, int periodic = -1, float box_size = 0.0f,
                  CdfArgs cdf = CdfArgs{nullptr, nullptr, nullptr, 0}, int flags = 0) {
    if (k <= 0) throw Error(NBK_ERR_INVALID, "k must be positive integer"); // pybind.cpp:92-94
    if (flags & ~NBK_QUERY_SQUARED) throw Error(NBK_ERR_INVALID, "unknown query flag");
    if (m == 0) return;
    if (m > 0xFFFFFFFFull) throw Error(NBK_ERR_INVALID, "more than 2^32-1 queries per call");
    // Morton grid over the box (periodic) or the points' bounding box (open)
    float lo[3], scale[3];
    QueryTree qt = tree.query_view(periodic, box_size);
    for (int d = 0; d < 3; ++d) {
        float l = qt.periodic ? 0.0f : tree.meta.lo[d];
        float h = qt.periodic ? qt.box : tree.meta.hi[d];
        float ext = h - l;
        lo[d] = l;
        scale[d] = (ext > 0.0f && std::isfinite(ext)) ? 1024.0f / ext : 0.0f;
    }
    // the lowest two bits per axis only order queries inside one 1/256-box cell (a leaf spans several
    // cells at any realistic density): three 8-bit passes over bits [6, 30) instead of four
    static const int first_bit = [] {
        const char *v = std::getenv("NBK_MORTON_FIRST_BIT");
        const int b = v ? std::atoi(v) : 6;
        return (b >= 0 && b < 30) ? b : 6;
    }();
    // NBK_ORDER=passes selects the three-kernel radix passes (rs::sort_pairs) instead of the single-sweep ones
    static const bool sweep = [] {
        const char *v = std::getenv("NBK_ORDER");
        return !(v && std::string(v) == "passes");
    }();
    const int passes = (30 - first_bit + 7) / 8;
    const bool use_sweep = sweep && m < (1ull << 30); // the look-back words hold 30-bit counts
    // a handful of warps gain nothing from an ordering that costs six launches: find_closest (m = 1) and
    // other tiny batches go straight to the kernel
    const bool ordered = m > 4096 || use_packet_kernel();
    Scratch scratch(stream);
    uint32_t *keys_a = scratch.get<uint32_t>(m), *keys_b = nullptr, *vals_a = nullptr, *vals_b = nullptr;
    const uint32_t *order = nullptr;
    int where = 0;
    SectionTimer t_order(NBK_SECTION_QUERY_ORDER, stream);
    if (ordered) {
        keys_b = scratch.get<uint32_t>(m);
        vals_a = scratch.get<uint32_t>(m);
        vals_b = scratch.get<uint32_t>(m);
        uint32_t *work = scratch.get<uint32_t>(use_sweep ? rs::sweep_workspace_entries(m, passes)
                                                     : rs::sort_workspace_entries<uint32_t>(m));
        if (use_sweep) {
            rs::SweepWorkspace w = rs::sweep_workspace(work, m, passes, stream);
            morton_keys_totals_kernel<<<(unsigned)div_up(m, kKeysPerCta), 256, 0, stream>>>(
                d_q, m, lo[0], lo[1], lo[2], scale[0], scale[1], scale[2], first_bit, passes, keys_a, w.digit_totals);
            NBK_LAUNCHED();
            where = rs::sweep_order(keys_a, vals_a, keys_b, vals_b, m, first_bit, passes, w, stream);
        } else {
            morton_keys_kernel<<<(unsigned)div_up(m, 256), 256, 0, stream>>>(d_q, m, lo[0], lo[1], lo[2], scale[0],
                                                                           scale[1], scale[2], keys_a, vals_a);
            NBK_LAUNCHED();
#ifndef NBK_SORT_STABLE
#define NBK_SORT_STABLE 0
#endif
            where = rs::sort_pairs<uint32_t, NBK_SORT_STABLE != 0>(keys_a, vals_a, keys_b, vals_b, m, first_bit, 30, work, stream);
        }
        order = where ? vals_b : vals_a;
    }
    t_order.finish();
    SectionTimer t_knn(NBK_SECTION_KNN_KERNEL, stream);
    DeferList defer{nullptr, nullptr};
    if (qt.periodic) { // (the packet kernel ignores it, but k > 8 falls through to the lane kernels even then)
        // the sort's input buffers are free again: reuse one as the deferred-query list
        defer.slots = (ordered && !where) ? keys_b : keys_a;
        defer.count = scratch.get<uint32_t>(2);
    }
    QueryBatch qb{d_q, order, m, k, flags, d_out_d, d_out_i, nullptr, 0u};
    uint64_t batch = m;
    if (k > max_shared_k()) {
        const uint64_t cols = global_heap_columns(k);
        qb.gheap = scratch.get<unsigned long long>(cols * (uint64_t)k);
        qb.gcolumns = (uint32_t)cols;
        batch = cols;
    }
    for (uint64_t begin = 0; begin < m; begin += batch) {
        qb.order = order + begin;
        qb.m = std::min(batch, m - begin);
        if (defer.count) NBK_CUDA(cudaMemsetAsync(defer.count, 0, 8, stream));
        if (qt.periodic) dispatch_knn<true>(qt, qb, defer, cdf, stream);
        else dispatch_knn<false>(qt, qb, defer, cdf, stream);
    }
    t_knn.finish();
}

// kNN-CDF: ks may come in any order; row i of `d_counts` belongs to ks[i].
struct CdfPlan {
    int kmax = 0;
    std::vector<int> row_of_rank; // [kmax]: row of the histogram of rank j (= k - 1), or -1
};

CdfPlan plan_cdf(const int *ks, int n_ks, int n_bins) {
    if (!ks || n_ks <= 0) throw Error(NBK_ERR_INVALID, "ks must hold at least one k");
    if (n_bins <= 0) throw Error(NBK_ERR_INVALID, "n_bins must be positive");
    CdfPlan p;
    for (int i = 0; i < n_ks; ++i) {
        if (ks[i] <= 0) throw Error(NBK_ERR_INVALID, "k must be positive integer");
        p.kmax = std::max(p.kmax, ks[i]);
    }
    p.row_of_rank.assign(p.kmax, -1);
    for (int i = 0; i < n_ks; ++i) {
        if (p.row_of_rank[ks[i] - 1] >= 0) throw Error(NBK_ERR_INVALID, "ks must be distinct");
        p.row_of_rank[ks[i] - 1] = i;
    }
    return p;
}

// d_counts[i][b] += histogram of the ks[i]-th neighbour distance; everything on `stream`
void knn_cdf_device(nbk_tree const &tree, const float *d_q, uint64_t m, CdfPlan const &plan, int n_ks,
                    const float *d_edges, int n_bins, unsigned long long *d_counts, cudaStream_t stream) {
    (void)n_ks;
    if (m == 0) return;
    Scratch scratch(stream);
    int *d_row = scratch.get<int>(plan.kmax);
    NBK_CUDA(cudaMemcpyAsync(d_row, plan.row_of_rank.data(), plan.kmax * sizeof(int), cudaMemcpyHostToDevice, stream));
    CdfArgs cdf{d_edges, d_counts, d_row, n_bins};
    query_device(tree, d_q, m, plan.kmax, nullptr, nullptr, stream, -1, 0.0f, cdf);
    NBK_CUDA(cudaStreamSynchronize(stream)); // plan.row_of_rank must outlive the copy
}

} // namespace nbk

using namespace nbk;

extern "C" {

const char *nbk_last_error(void) { return g_error.c_str(); }

uint64_t nbk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int nbk_device_count(void) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) {
        g_error = cudaGetErrorString(e);
        return -1;
    }
    return count;
}

nbk_tree *nbk_tree_build(const float *xyz_aos, uint64_t n, int leaf_size, int block_size,
                         int periodic, float box_size, int device, int *status) {
    nbk_tree *out = nullptr;
    int st = guarded([&] {
        check_build_args(n, block_size, false);
        require_sm100(device);
        DeviceGuard guard(device);
        cudaStream_t stream = nullptr; // legacy default stream
        const auto t0 = std::chrono::steady_clock::now();
        auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
        const char *tv = std::getenv("NBK_BUILD_TRACE");
        const bool trace = tv && tv[0] == '1';
        std::unique_ptr<nbk_tree> tree;
        {
            // the staging copy of the caller's points comes from the block cache too (kept between builds)
            Scratch staging(stream);
            staging.reserve(Scratch::padded(std::max<uint64_t>(n, 1) * 12));
            float *d_aos = staging.get<float>(std::max<uint64_t>(n, 1) * 3);
            const double t_alloc = since();
            {
                const bool big_pageable = n * 12ull >= (32u << 20) && is_pageable_host(xyz_aos);
                int dev = 0;
                NBK_CUDA(cudaGetDevice(&dev));
                PinnedRing *ring = big_pageable ? PinnedRing::acquire(dev) : nullptr;
                if (big_pageable) g_host_path[ring ? 2 : 3].fetch_add(1);
                if (ring) {
                    struct Release {
                        PinnedRing *r;
                        ~Release() { r->release(); }
                    } release{ring};
                    static const int threads = (int)std::min(8u, std::max(2u, std::thread::hardware_concurrency() / 2));
                    StagedUpload up(ring, dev, threads);
                    up.upload(d_aos, xyz_aos, n * 12, stream);
                } else {
                    NBK_CUDA(cudaMemcpyAsync(d_aos, xyz_aos, n * 12, cudaMemcpyHostToDevice, stream));
                    NBK_CUDA(cudaStreamSynchronize(stream));
                }
            }
            const double t_copy = since();
            tree = build_from_device_aos(d_aos, n, leaf_size, block_size, periodic, box_size, device, stream);
            if (trace)
                fprintf(stderr, "[nbk build] host entry: staging block %.3f ms, upload of %.1f MB %.3f ms, build %.3f ms\n",
                        t_alloc, n * 12 / 1e6, t_copy - t_alloc, since() - t_copy);
        }
        if (trace) fprintf(stderr, "[nbk build] host entry total %.3f ms\n", since());
        out = tree.release();
    });
    if (status) *status = st;
    return out;
}

nbk_tree *nbk_tree_build_device(const float *d_xyz_aos, uint64_t n, int leaf_size, int block_size,
                                int periodic, float box_size, int device, void *stream,
                                int *status) {
    nbk_tree *out = nullptr;
    int st = guarded([&] {
        check_build_args(n, block_size, false);
        require_sm100(device);
        DeviceGuard guard(device);
        out = build_from_device_aos(d_xyz_aos, n, leaf_size, block_size, periodic, box_size, device,
                                    static_cast<cudaStream_t>(stream))
                  .release();
    });
    if (status) *status = st;
    return out;
}

nbk_tree *nbk_tree_build_soa(const float *x, const float *y, const float *z, const uint32_t *idx,
                             uint64_t n_padded, int leaf_size, int block_size, int periodic,
                             float box_size, int device, int *status) {
    nbk_tree *out = nullptr;
    int st = guarded([&] {
        check_build_args(n_padded, block_size, true);
        require_sm100(device);
        DeviceGuard guard(device);
        cudaStream_t stream = nullptr;
        bool trim_after = false;
        {
            Scratch scratch(stream);
            uint64_t cols = std::max<uint64_t>(n_padded, 1);
            scratch.reserve(5 * Scratch::padded(cols * 4) + 256);
            float *x0 = scratch.get<float>(cols), *y0 = scratch.get<float>(cols), *z0 = scratch.get<float>(cols);
            uint32_t *idx0 = scratch.get<uint32_t>(cols), *perm = scratch.get<uint32_t>(cols);
            uint32_t *aux = scratch.get<uint32_t>(8);
            const uint32_t init[8] = {0u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u, 0u};
            NBK_CUDA(cudaMemcpyAsync(aux, init, sizeof init, cudaMemcpyHostToDevice, stream));
            NBK_CUDA(cudaMemcpyAsync(x0, x, n_padded * 4, cudaMemcpyHostToDevice, stream));
            NBK_CUDA(cudaMemcpyAsync(y0, y, n_padded * 4, cudaMemcpyHostToDevice, stream));
            NBK_CUDA(cudaMemcpyAsync(z0, z, n_padded * 4, cudaMemcpyHostToDevice, stream));
            NBK_CUDA(cudaMemcpyAsync(idx0, idx, n_padded * 4, cudaMemcpyHostToDevice, stream));
            if (n_padded) {
                scan_soa_kernel<<<(unsigned)std::min<uint64_t>(div_up(n_padded, 256), 148 * 16), 256, 0, stream>>>(
                    x0, y0, z0, n_padded, perm, aux + 1);
                NBK_LAUNCHED();
            }
            out = finish_build(n_padded, n_padded, leaf_size, block_size, periodic, box_size, x0, y0, z0, idx0,
                               trim_after, perm, aux + 1, device, stream)
                      .release();
        }
        if (trim_after) {
            int dev = 0;
            NBK_CUDA(cudaGetDevice(&dev));
            BlockCache::trim(dev);
        }
    });
    if (status) *status = st;
    return out;
}

int nbk_plan_topology(uint64_t n_points, int leaf_size, int block_size, nbk_node *nodes,
                      uint64_t *n_nodes, int *n_levels) {
    return guarded([&] {
        check_build_args(n_points, block_size, false);
        uint64_t n_padded = div_up(n_points, block_size) * block_size;
        check_build_args(n_padded, block_size, false);
        // the memoised count rules the build uses (no O(n_nodes) work) ...
        td::TopPlan top = td::plan_top(n_padded, leaf_size, block_size);
        if (n_nodes) *n_nodes = top.n_nodes;
        if (n_levels) *n_levels = top.n_levels;
        if (nodes) {
            // ... and the explicit recursion (kdtree_impl.hpp:98-157) when the records are wanted; the two
            // must agree
            TopologyPlan plan = plan_topology(n_padded, leaf_size, block_size);
            if (plan.nodes.size() != top.n_nodes || (int)plan.levels.size() != top.n_levels)
                throw Error(NBK_ERR_INVALID, "internal error: topology plans disagree");
            std::copy(plan.nodes.begin(), plan.nodes.end(), nodes);
        }
    });
}

void nbk_tree_free(nbk_tree *tree) {
    // ~nbk_tree orders the release of the arena behind the last use on every stream that queried the
    // tree (no device-wide synchronisation: other trees' work keeps running)
    delete tree;
}

int nbk_tree_get_meta(const nbk_tree *tree, nbk_tree_meta *meta) {
    return guarded([&] {
        if (!tree || !meta) throw Error(NBK_ERR_INVALID, "null argument");
        *meta = tree->meta;
    });
}

int nbk_tree_device(const nbk_tree *tree) { return tree ? tree->device : -1; }

int nbk_tree_copy_nodes(const nbk_tree *tree, nbk_node *nodes) {
    return guarded([&] {
        if (!tree || !nodes) throw Error(NBK_ERR_INVALID, "null argument");
        DeviceGuard guard(tree->device);
        NBK_CUDA(cudaMemcpy(nodes, tree->view.nodes, tree->meta.n_nodes * sizeof(nbk_node),
                            cudaMemcpyDeviceToHost));
    });
}

int nbk_tree_copy_points(const nbk_tree *tree, float *x, float *y, float *z, uint32_t *idx) {
    return guarded([&] {
        if (!tree) throw Error(NBK_ERR_INVALID, "null argument");
        DeviceGuard guard(tree->device);
        const uint64_t n = tree->meta.n_padded;
        if (n == 0) return;
        cudaStream_t stream = nullptr;
        Scratch scratch(stream);
        float *dx = scratch.get<float>(n), *dy = scratch.get<float>(n), *dz = scratch.get<float>(n);
        uint32_t *di = scratch.get<uint32_t>(n);
        untile_kernel<<<(unsigned)div_up(n, 256), 256, 0, stream>>>(
            reinterpret_cast<const float4 *>(tree->view.tiles), n, dx, dy, dz, di);
        NBK_LAUNCHED();
        if (x) NBK_CUDA(cudaMemcpyAsync(x, dx, n * 4, cudaMemcpyDeviceToHost, stream));
        if (y) NBK_CUDA(cudaMemcpyAsync(y, dy, n * 4, cudaMemcpyDeviceToHost, stream));
        if (z) NBK_CUDA(cudaMemcpyAsync(z, dz, n * 4, cudaMemcpyDeviceToHost, stream));
        if (idx) NBK_CUDA(cudaMemcpyAsync(idx, di, n * 4, cudaMemcpyDeviceToHost, stream));
        NBK_CUDA(cudaStreamSynchronize(stream));
    });
}

int nbk_tree_query_device(const nbk_tree *tree, const float *d_q_aos, uint64_t m, int k,
                          float *d_out_dist, uint32_t *d_out_idx, void *stream) {
    return nbk_tree_query_device_ex(tree, d_q_aos, m, k, -1, 0.0f, 0, d_out_dist, d_out_idx, stream);
}

int nbk_tree_query_device_ex(const nbk_tree *tree, const float *d_q_aos, uint64_t m, int k, int periodic,
                             float box_size, int flags, float *d_out_dist, uint32_t *d_out_idx, void *stream) {
    return guarded([&] {
        if (!tree) throw Error(NBK_ERR_INVALID, "null argument");
        DeviceGuard guard(tree->device);
        query_device(*tree, d_q_aos, m, k, d_out_dist, d_out_idx, static_cast<cudaStream_t>(stream), periodic,
                     box_size, CdfArgs{nullptr, nullptr, nullptr, 0}, flags);
        tree->mark_use(static_cast<cudaStream_t>(stream)); // the call does not synchronise
    });
}

int nbk_tree_query(const nbk_tree *tree, const float *q_aos, uint64_t m, int k, float *out_dist,
                   uint32_t *out_idx) {
    return nbk_tree_query_ex(tree, q_aos, m, k, -1, 0.0f, out_dist, out_idx);
}

int nbk_tree_query_ex(const nbk_tree *tree, const float *q_aos, uint64_t m, int k, int periodic,
                      float box_size, float *out_dist, uint32_t *out_idx) {
    return nbk_tree_query_ex2(tree, q_aos, m, k, periodic, box_size, 0, out_dist, out_idx);
}

int nbk_tree_query_ex2(const nbk_tree *tree, const float *q_aos, uint64_t m, int k, int periodic,
                       float box_size, int flags, float *out_dist, uint32_t *out_idx) {
    return guarded([&] {
        if (!tree) throw Error(NBK_ERR_INVALID, "null argument");
        if (k <= 0) throw Error(NBK_ERR_INVALID, "k must be positive integer");
        if (flags & ~NBK_QUERY_SQUARED) throw Error(NBK_ERR_INVALID, "unknown query flag");
        if (m == 0) return;
        DeviceGuard guard(tree->device);
        if (m <= 4096) {
            // find_closest and other tiny batches: no pipeline to set up, one round trip on the calling
            // thread's own stream
            cudaStream_t st = cudaStreamPerThread;
            Scratch scratch(st);
            float *dq = scratch.get<float>(m * 3), *dd = scratch.get<float>(m * (uint64_t)k);
            uint32_t *di = scratch.get<uint32_t>(m * (uint64_t)k);
            NBK_CUDA(cudaMemcpyAsync(dq, q_aos, m * 12, cudaMemcpyHostToDevice, st));
            query_device(*tree, dq, m, k, dd, di, st, periodic, box_size, CdfArgs{nullptr, nullptr, nullptr, 0}, flags);
            NBK_CUDA(cudaMemcpyAsync(out_dist, dd, m * (uint64_t)k * 4, cudaMemcpyDeviceToHost, st));
            NBK_CUDA(cudaMemcpyAsync(out_idx, di, m * (uint64_t)k * 4, cudaMemcpyDeviceToHost, st));
            NBK_CUDA(cudaStreamSynchronize(st));
            return;
        }
        // Up to three slices in flight, one stream each: slice c+2 uploads while slice c+1 computes
        // and slice c downloads.  A slice is ordered (Morton sort) on its own, and the kernel gets slower
        // per query as slices get smaller (fewer queries per leaf), but the pipeline is bound by the copies
        // over PCIe (8k bytes of rows per query against ~1 ns of kernel), so that slack is free and small
        // slices win: the first rows leave after a few hundred microseconds and the last download, which
        // nothing overlaps, is short.  Measured, 10^8 queries from pinned buffers: k = 8: 2^24-query slices
        // 0.134 s, 2^23 0.124 s, 2^22 with a 2^19 first slice 0.119 s = 0.84 G queries/s against a PCIe
        // ceiling of 0.85; k = 4 / 16 / 32: 0.078 / 0.240 / 0.479 s against 0.080 / 0.247 / 0.489 s.
        static const uint64_t slice_cfg = [] {
            const char *v = std::getenv("NBK_HOST_SLICE");
            uint64_t n = v ? std::strtoull(v, nullptr, 10) : 0;
            return n ? n : (1ull << 22);
        }();
        static const uint64_t first_slice = [] {
            const char *v = std::getenv("NBK_HOST_FIRST_SLICE");
            uint64_t n = v ? std::strtoull(v, nullptr, 10) : 0;
            return n ? n : (1ull << 19);
        }();
        // wide rows: keep the three in-flight result buffers at <= 1 GB each
        const uint64_t by_rows = std::max<uint64_t>(1ull << 16, (1ull << 27) / (uint64_t)k);
        const uint64_t slice = std::min<uint64_t>(m, std::min(slice_cfg, by_rows));
        const int nbuf = (int)std::min<uint64_t>(3, div_up(m, std::min(slice, first_slice)));
        cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
        float *d_q[3] = {}, *d_d[3] = {};
        uint32_t *d_i[3] = {};
        auto cleanup = [&] {
            for (int s = 0; s < nbuf; ++s) {
                if (!streams[s]) continue;
                if (d_q[s]) cudaFreeAsync(d_q[s], streams[s]);
                if (d_d[s]) cudaFreeAsync(d_d[s], streams[s]);
                if (d_i[s]) cudaFreeAsync(d_i[s], streams[s]);
                cudaStreamSynchronize(streams[s]);
                cudaStreamDestroy(streams[s]);
            }
        };
        // Results for pageable memory (fresh numpy arrays) go through a pinned ring and host threads
        // (host_stage.cuh); pinned or registered destinations are written by the copy engine directly.
        static const int host_threads = [] {
            const char *v = std::getenv("NBK_HOST_THREADS");
            int n = v ? std::atoi(v) : 0;
            if (n <= 0) n = (int)std::min(8u, std::max(2u, std::thread::hardware_concurrency() / 2));
            return n;
        }();
        std::unique_ptr<StagedDownload> staged;
        PinnedRing *staged_ring = nullptr;
        // (decided per array: one of the two may be a recycled, page-locked buffer while the other is fresh)
        const bool big_rows = m * (uint64_t)k * 8 >= (64u << 20);
        const bool stage_dist = big_rows && is_pageable_host(out_dist), stage_idx = big_rows && is_pageable_host(out_idx);
        if (stage_dist || stage_idx) {
            if ((staged_ring = PinnedRing::acquire(tree->device))) {
                try {
                    staged = std::make_unique<StagedDownload>(staged_ring, tree->device, host_threads);
                } catch (...) {
                    staged_ring->release();
                    throw;
                }
            }
            g_host_path[staged ? 0 : 1].fetch_add(1);
        }
        std::unique_ptr<StagedUpload> uploader;
        PinnedRing *upload_ring = nullptr;
        if (m * 12ull >= (32u << 20) && is_pageable_host(q_aos)) {
            // shares the ring of the staged download if there is one, otherwise takes it for itself
            if (staged) upload_ring = staged_ring;
            else upload_ring = PinnedRing::acquire(tree->device);
            g_host_path[upload_ring ? 2 : 3].fetch_add(1);
        }
        struct RingRelease {
            PinnedRing *ring;
            ~RingRelease() { if (ring) ring->release(); }
        } upload_ring_release{staged ? nullptr : upload_ring};
        if (upload_ring) uploader = std::make_unique<StagedUpload>(upload_ring, tree->device, std::max(2, host_threads / 2));
        auto download = [&](int s, uint64_t begin, uint64_t cnt) {
            const uint64_t bytes = cnt * (uint64_t)k * 4;
            if (staged && stage_dist) staged->download(out_dist + begin * k, d_d[s], bytes, streams[s]);
            else NBK_CUDA(cudaMemcpyAsync(out_dist + begin * k, d_d[s], bytes, cudaMemcpyDeviceToHost, streams[s]));
            if (staged && stage_idx) staged->download(out_idx + begin * k, d_i[s], bytes, streams[s]);
            else NBK_CUDA(cudaMemcpyAsync(out_idx + begin * k, d_i[s], bytes, cudaMemcpyDeviceToHost, streams[s]));
        };
        try {
            for (int s = 0; s < nbuf; ++s) {
                NBK_CUDA(cudaStreamCreateWithFlags(&streams[s], cudaStreamNonBlocking));
                NBK_CUDA(cudaMallocAsync(&d_q[s], slice * 12, streams[s]));
                NBK_CUDA(cudaMallocAsync(&d_d[s], slice * (uint64_t)k * 4, streams[s]));
                NBK_CUDA(cudaMallocAsync(&d_i[s], slice * (uint64_t)k * 4, streams[s]));
            }
            int s = 0;
            uint64_t step = std::min(slice, first_slice);
            // the download of slice c is enqueued after the kernel of slice c+1, so that the device has
            // work while this thread waits for ring slots
            int pend_s = -1;
            uint64_t pend_begin = 0, pend_cnt = 0;
            for (uint64_t begin = 0; begin < m; s = (s + 1) % nbuf) {
                const uint64_t cnt = std::min(step, m - begin);
                cudaStream_t st = streams[s];
                if (uploader) uploader->upload(d_q[s], q_aos + begin * 3, cnt * 12, st);
                else NBK_CUDA(cudaMemcpyAsync(d_q[s], q_aos + begin * 3, cnt * 12, cudaMemcpyHostToDevice, st));
                query_device(*tree, d_q[s], cnt, k, d_d[s], d_i[s], st, periodic, box_size,
                             CdfArgs{nullptr, nullptr, nullptr, 0}, flags);
                if (pend_s >= 0) download(pend_s, pend_begin, pend_cnt);
                pend_s = s;
                pend_begin = begin;
                pend_cnt = cnt;
                begin += cnt;
                step = std::min(slice, step * 2);
            }
            if (pend_s >= 0) download(pend_s, pend_begin, pend_cnt);
            if (staged) staged->finish();
            for (int t = 0; t < nbuf; ++t) NBK_CUDA(cudaStreamSynchronize(streams[t]));
        } catch (...) {
            uploader.reset();
            staged.reset();
            cleanup();
            throw;
        }
        uploader.reset();
        staged.reset();
        cleanup();
    });
}

int nbk_scan_block(const float *x, const float *y, const float *z, const uint32_t *idx, uint64_t n,
                   const float *q_aos, uint64_t m, int k, int periodic, float box_size, int flags,
                   float *out_dist, uint32_t *out_idx, int device) {
    return guarded([&] {
        if (k <= 0) throw Error(NBK_ERR_INVALID, "k must be positive integer");
        if (n % 8 != 0) throw Error(NBK_ERR_INVALID, "block_size must be a multiple of 8.");
        if (n > 0xFFFFFFFFull) throw Error(NBK_ERR_INVALID, "More than uint32_t points are not supported.");
        if (flags & ~NBK_QUERY_SQUARED) throw Error(NBK_ERR_INVALID, "unknown query flag");
        if (m == 0) return;
        if ((n && (!x || !y || !z || !idx)) || !q_aos || !out_dist || !out_idx) throw Error(NBK_ERR_INVALID, "null argument");
        require_sm100(device);
        DeviceGuard guard(device);
        cudaStream_t stream = nullptr;
        // the block in the tree's tile layout {x[8], y[8], z[8], idx[8]}
        std::vector<float> tiles(std::max<uint64_t>(n, 8) * 4);
        for (uint64_t p = 0; p < n; ++p) {
            float *t = tiles.data() + (p >> 3) * 32 + (p & 7);
            t[0] = x[p];
            t[8] = y[p];
            t[16] = z[p];
            std::memcpy(&t[24], &idx[p], 4);
        }
        Scratch scratch(stream);
        float *d_tiles = scratch.get<float>(tiles.size());
        float *d_q = scratch.get<float>(m * 3);
        float *d_d = scratch.get<float>(m * (uint64_t)k);
        uint32_t *d_i = scratch.get<uint32_t>(m * (uint64_t)k);
        NBK_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * 4, cudaMemcpyHostToDevice, stream));
        NBK_CUDA(cudaMemcpyAsync(d_q, q_aos, m * 12, cudaMemcpyHostToDevice, stream));
        QueryTree qt{};
        qt.nodes = nullptr;
        qt.tiles = reinterpret_cast<const float4 *>(d_tiles);
        qt.periodic = periodic != 0;
        qt.box = periodic ? box_size : 0.0f;
        QueryBatch qb{d_q, nullptr, m, k, flags, d_d, d_i, nullptr, 0u};
        if (k > kMaxSharedK) {
            const uint64_t cols = align_up(m, kQueryThreads);
            qb.gheap = scratch.get<unsigned long long>(cols * (uint64_t)k);
            qb.gcolumns = (uint32_t)cols;
        }
        if (qt.periodic) dispatch_scan_block<true>(qt, (uint32_t)n, qb, stream);
        else dispatch_scan_block<false>(qt, (uint32_t)n, qb, stream);
        NBK_CUDA(cudaMemcpyAsync(out_dist, d_d, m * (uint64_t)k * 4, cudaMemcpyDeviceToHost, stream));
        NBK_CUDA(cudaMemcpyAsync(out_idx, d_i, m * (uint64_t)k * 4, cudaMemcpyDeviceToHost, stream));
        NBK_CUDA(cudaStreamSynchronize(stream));
    });
}

int nbk_tree_knn_cdf_device(const nbk_tree *tree, const float *d_q_aos, uint64_t m, const int *ks, int n_ks,
                            const float *d_edges, int n_bins, unsigned long long *d_counts, void *stream) {
    return guarded([&] {
        if (!tree || !d_edges || !d_counts) throw Error(NBK_ERR_INVALID, "null argument");
        CdfPlan plan = plan_cdf(ks, n_ks, n_bins);
        DeviceGuard guard(tree->device);
        knn_cdf_device(*tree, d_q_aos, m, plan, n_ks, d_edges, n_bins, d_counts, static_cast<cudaStream_t>(stream));
    });
}

int nbk_tree_knn_cdf(const nbk_tree *tree, const float *q_aos, uint64_t m, const int *ks, int n_ks,
                     const float *edges, int n_bins, uint64_t *counts) {
    return guarded([&] {
        if (!tree || !edges || !counts) throw Error(NBK_ERR_INVALID, "null argument");
        CdfPlan plan = plan_cdf(ks, n_ks, n_bins);
        for (int b = 0; b < n_bins; ++b)
            if (!(edges[b] <= edges[b + 1])) throw Error(NBK_ERR_INVALID, "bin edges must increase monotonically");
        DeviceGuard guard(tree->device);
        cudaStream_t stream = nullptr;
        Scratch scratch(stream);
        const uint64_t cells = (uint64_t)n_ks * n_bins;
        float *d_edges = scratch.get<float>(n_bins + 1);
        unsigned long long *d_counts = scratch.get<unsigned long long>(cells);
        NBK_CUDA(cudaMemcpyAsync(d_edges, edges, (n_bins + 1) * 4, cudaMemcpyHostToDevice, stream));
        NBK_CUDA(cudaMemsetAsync(d_counts, 0, cells * 8, stream));
        // queries stream through the device in slices; only the histogram comes back
        const uint64_t slice = std::min<uint64_t>(std::max<uint64_t>(m, 1), 1ull << 25);
        float *d_q = scratch.get<float>(slice * 3);
        for (uint64_t begin = 0; begin < m; begin += slice) {
            const uint64_t cnt = std::min(slice, m - begin);
            NBK_CUDA(cudaMemcpyAsync(d_q, q_aos + begin * 3, cnt * 12, cudaMemcpyHostToDevice, stream));
            knn_cdf_device(*tree, d_q, cnt, plan, n_ks, d_edges, n_bins, d_counts, stream);
        }
        std::vector<unsigned long long> h(cells);
        NBK_CUDA(cudaMemcpyAsync(h.data(), d_counts, cells * 8, cudaMemcpyDeviceToHost, stream));
        NBK_CUDA(cudaStreamSynchronize(stream));
        for (uint64_t i = 0; i < cells; ++i) counts[i] += h[i];
    });
}

int nbk_tree_stats(const nbk_tree *tree, const float *q_aos, uint64_t m, int k, int periodic,
                   float box_size, uint64_t *out3) {
    return guarded([&] {
        if (!tree || !out3) throw Error(NBK_ERR_INVALID, "null argument");
        if (k <= 0) throw Error(NBK_ERR_INVALID, "k must be positive integer");
        out3[0] = out3[1] = out3[2] = 0;
        if (m == 0) return;
        DeviceGuard guard(tree->device);
        cudaStream_t stream = nullptr;
        Scratch scratch(stream);
        float *d_q = scratch.get<float>(m * 3);
        unsigned long long *d_out = scratch.get<unsigned long long>(3);
        // k > 64: the replace-top queue of every query lives in a global scratch row
        float *d_best = k > kStatsMaxK ? scratch.get<float>(m * (uint64_t)k) : nullptr;
        NBK_CUDA(cudaMemcpyAsync(d_q, q_aos, m * 12, cudaMemcpyHostToDevice, stream));
        NBK_CUDA(cudaMemsetAsync(d_out, 0, 24, stream));
        QueryTree qt = tree->query_view(periodic, box_size);
        unsigned grid = (unsigned)div_up(m, 128);
        if (qt.periodic) stats_kernel<true><<<grid, 128, 0, stream>>>(qt, d_q, m, k, d_best, d_out);
        else stats_kernel<false><<<grid, 128, 0, stream>>>(qt, d_q, m, k, d_best, d_out);
        NBK_LAUNCHED();
        unsigned long long h[3];
        NBK_CUDA(cudaMemcpyAsync(h, d_out, 24, cudaMemcpyDeviceToHost, stream));
        NBK_CUDA(cudaStreamSynchronize(stream));
        for (int c = 0; c < 3; ++c) out3[c] = h[c];
    });
}

int nbk_tree_arena(const nbk_tree *tree, void **d_arena, uint64_t *bytes) {
    return guarded([&] {
        if (!tree) throw Error(NBK_ERR_INVALID, "null argument");
        if (d_arena) *d_arena = tree->arena;
        if (bytes) *bytes = tree->meta.arena_bytes;
    });
}

nbk_tree *nbk_tree_alloc_replica(const nbk_tree_meta *meta, int device, int *status) {
    nbk_tree *out = nullptr;
    int st = guarded([&] {
        if (!meta) throw Error(NBK_ERR_INVALID, "null argument");
        require_sm100(device);
        DeviceGuard guard(device);
        ArenaLayout l = arena_layout(meta->n_padded, meta->n_nodes);
        if (meta->arena_bytes != l.total) throw Error(NBK_ERR_INVALID, "tree meta does not match this library's arena layout");
        out = alloc_tree(*meta, nullptr).release();
    });
    if (status) *status = st;
    return out;
}

nbk_tree *nbk_tree_clone_to_device(const nbk_tree *tree, int device, int *status) {
    nbk_tree *out = nullptr;
    int st = guarded([&] {
        if (!tree) throw Error(NBK_ERR_INVALID, "null argument");
        require_sm100(device);
        std::unique_ptr<nbk_tree> copy;
        {
            DeviceGuard guard(device);
            copy = alloc_tree(tree->meta, nullptr);
            NBK_CUDA(cudaDeviceSynchronize());
        }
        DeviceGuard guard(tree->device);
        NBK_CUDA(cudaDeviceSynchronize()); // the source tree is complete
        NBK_CUDA(cudaMemcpyPeer(copy->arena, copy->device, tree->arena, tree->device, tree->meta.arena_bytes));
        out = copy.release();
    });
    if (status) *status = st;
    return out;
}

void nbk_profile_enable(int on) { g_profile.store(on ? 1 : 0); }

int nbk_profile_read(int section, double *total_ms, uint64_t *count) {
    return guarded([&] {
        if (section < 0 || section >= NBK_SECTION_COUNT) throw Error(NBK_ERR_INVALID, "bad section");
        std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
        {
            std::lock_guard<std::mutex> lock(g_profile_mutex);
            events.swap(g_profile_events[section]);
        }
        double total = 0.0;
        for (auto &e : events) {
            NBK_CUDA(cudaEventSynchronize(e.second));
            float ms = 0.0f;
            NBK_CUDA(cudaEventElapsedTime(&ms, e.first, e.second));
            total += ms;
            cudaEventDestroy(e.first);
            cudaEventDestroy(e.second);
        }
        if (total_ms) *total_ms = total;
        if (count) *count = events.size();
    });
}

void *nbk_device_alloc(uint64_t bytes) { return nbk_device_alloc_on(-1, bytes); }

void *nbk_device_alloc_on(int device, uint64_t bytes) {
    void *p = nullptr;
    int st = guarded([&] {
        DeviceGuard guard(device);
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e != cudaSuccess) {
            p = nullptr;
            throw Error(NBK_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
        }
    });
    return st == NBK_OK ? p : nullptr;
}

int nbk_pointer_device(const void *ptr) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) ? attr.device : -1;
}

int nbk_host_path_stats(uint64_t *out4) {
    return guarded([&] {
        if (!out4) throw Error(NBK_ERR_INVALID, "null argument");
        for (int i = 0; i < 4; ++i) out4[i] = g_host_path[i].load();
    });
}

void nbk_device_free(void *ptr) {
    if (ptr) cudaFree(ptr);
}

int nbk_device_copy(void *dst, const void *src, uint64_t bytes, int kind) {
    return guarded([&] {
        if (kind != 0 && kind != 1) throw Error(NBK_ERR_INVALID, "kind must be 0 (h2d) or 1 (d2h)");
        NBK_CUDA(cudaMemcpy(dst, src, bytes, kind == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost));
    });
}

int nbk_device_zero(void *dst, uint64_t bytes) {
    // complete on return: work the caller enqueues afterwards on ANY stream (non-blocking ones included)
    // sees the zeros
    return guarded([&] {
        NBK_CUDA(cudaMemsetAsync(dst, 0, bytes, nullptr));
        NBK_CUDA(cudaStreamSynchronize(nullptr));
    });
}

void *nbk_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        g_error = "cudaHostAlloc failed";
        return nullptr;
    }
    return p;
}

void nbk_host_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

int nbk_host_register(void *ptr, uint64_t bytes, int device) {
    return guarded([&] {
        if (!ptr || !bytes) throw Error(NBK_ERR_INVALID, "null argument");
        DeviceGuard guard(device);
        NBK_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    });
}

int nbk_host_unregister(void *ptr) {
    return guarded([&] {
        if (!ptr) throw Error(NBK_ERR_INVALID, "null argument");
        NBK_CUDA(cudaHostUnregister(ptr));
    });
}

} // extern "C"
