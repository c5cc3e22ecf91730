// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
//
// Thin C-ABI harness around the UNMODIFIED reference sources under /root/reference/kdtree
// (compiled where they lie by oracle/Makefile; outputs go to oracle/_ref/ only).  It exists so
// that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can
// drive the reference's own tree builder and query traversal from Python (ctypes).
//
// What runs here is the reference's code, not a restatement:
//   * build  : wenda::kdtree::KDTree(PositionAndIndexArray<3>, config)        kdtree.cpp:95-131
//              on positions padded exactly like pybind.cpp:14-56 (FLT_MAX pad to a multiple of 8)
//   * query  : detail::KDTreeQuery<Distance, TournamentTree<..>, InsertShorterDistanceAVX>
//              + copy_values + std::sort(PairLessFirst) + postprocess, i.e. the body of
//              KDTree::find_closest (kdtree.cpp:133-159) with the inserter swapped from the NASM
//              kernel (nasm is not installed in this image) to the reference's own AVX2-intrinsics
//              inserter, which the reference's typed tests pin equal to the asm one
//              (tests/test_inserters.cpp:91-99,128-224).
//   * threads: wenda::thread_pool::parallelize_loop over contiguous query chunks  pybind.cpp:164-172
//   * fixtures: make_random_position_and_index (Philox4x32)                    kdtree_utils.hpp:16-46
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <kdtree/kdtree.hpp>
#include <kdtree/kdtree_impl.hpp>
#include <kdtree/kdtree_opt.hpp>
#include <kdtree/kdtree_utils.hpp>
#include <kdtree/tournament_tree.hpp>
#include <thread_pool.hpp>

namespace kdt = wenda::kdtree;

// The two NASM entry points kdtree.cpp refers to (kdtree_opt_asm.hpp:12-19).  The harness never
// routes a query through KDTree::find_closest, so they are never reached.
extern "C" void wenda_insert_closest_l2_avx2(float const *const *, size_t, float const *, void *,
                                             uint32_t const *) {
    std::abort();
}
extern "C" void wenda_insert_closest_l2_periodic_avx2(float const *const *, size_t, float const *,
                                                      void *, uint32_t const *, float) {
    std::abort();
}

namespace {

struct RefTree {
    kdt::KDTree tree;
    bool periodic;
    float box;
};

thread_local std::string g_error;

typedef std::pair<float, uint32_t> result_t;
typedef kdt::TournamentTree<result_t, kdt::PairLessFirst> queue_t;

template <typename Distance>
void query_range(RefTree const &t, Distance const &distance, const float *q, size_t begin,
                 size_t end, int k, bool squared, float *out_d, uint32_t *out_i, uint64_t *stats) {
    std::vector<result_t> result(k);
    uint64_t nv = 0, np = 0, pv = 0;
    for (size_t i = begin; i < end; ++i) {
        std::array<float, 3> query = {q[3 * i], q[3 * i + 1], q[3 * i + 2]};
        kdt::detail::KDTreeQuery<Distance, queue_t, kdt::InsertShorterDistanceAVX> search(
            t.tree.nodes(), t.tree.positions(), distance, query, k);
        search.compute(&t.tree.nodes()[0]);
        search.distances_.copy_values(result.begin());
        std::sort(result.begin(), result.end(), kdt::PairLessFirst{});
        for (int j = 0; j < k; ++j) {
            // squared: the Distance functor's value as the search ranks it, before postprocess()
            out_d[i * k + j] = squared ? result[j].first : distance.postprocess(result[j].first);
            out_i[i * k + j] = result[j].second;
        }
        nv += search.num_nodes_visited;
        np += search.num_nodes_pruned;
        pv += search.num_points_visited;
    }
    if (stats) {
        __atomic_fetch_add(&stats[0], nv, __ATOMIC_RELAXED);
        __atomic_fetch_add(&stats[1], np, __ATOMIC_RELAXED);
        __atomic_fetch_add(&stats[2], pv, __ATOMIC_RELAXED);
    }
}

} // namespace

extern "C" {

const char *ref_last_error() { return g_error.c_str(); }

int ref_hardware_concurrency() { return (int)std::thread::hardware_concurrency(); }

// Philox fixtures of the reference's tests, AoS (n,3) float32.
void ref_philox_points(uint32_t n, unsigned seed, float boxsize, float *out_aos) {
    auto pts = kdt::make_random_position_and_index<3>(n, seed, boxsize);
    for (uint32_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d)
            out_aos[3 * (size_t)i + d] = pts[i].position[d];
}

// Builds with the reference's constructor.  box < 0 => open boundaries.
void *ref_tree_build(const float *xyz_aos, uint64_t n, int leaf_size, float box,
                     double *build_seconds) {
    try {
        const int block = 8;
        size_t size_up = (n + block - 1) / block * block;
        kdt::PositionAndIndexArray<3, float, uint32_t> positions(size_up);
        std::iota(positions.indices_.begin(), positions.indices_.end(), 0);
        for (int d = 0; d < 3; ++d) {
            float *col = positions.positions_[d];
            for (size_t i = 0; i < n; ++i)
                col[i] = xyz_aos[3 * i + d];
            if (box >= 0) {
                for (size_t i = 0; i < n; ++i)
                    if (!(col[i] >= 0.0f && col[i] <= box))
                        throw std::runtime_error(
                            "When using periodic boundary conditions, all points must be "
                            "within the box (0 <= x <= box_size).");
            }
            std::fill(col + n, col + size_up, std::numeric_limits<float>::max());
        }
        auto t0 = std::chrono::steady_clock::now();
        kdt::KDTreeConfiguration config;
        config.leaf_size = leaf_size;
        config.max_threads = 0;
        config.block_size = block;
        auto *t = new RefTree{kdt::KDTree(std::move(positions), config), box >= 0,
                              box >= 0 ? box : 0.0f};
        auto t1 = std::chrono::steady_clock::now();
        if (build_seconds)
            *build_seconds = std::chrono::duration<double>(t1 - t0).count();
        return t;
    } catch (std::exception const &e) {
        g_error = e.what();
        return nullptr;
    }
}

void ref_tree_free(void *handle) { delete static_cast<RefTree *>(handle); }

uint64_t ref_tree_num_points(void *handle) {
    return static_cast<RefTree *>(handle)->tree.positions().size();
}
uint64_t ref_tree_num_nodes(void *handle) {
    return static_cast<RefTree *>(handle)->tree.nodes().size();
}

// nodes16: n_nodes records of {int32 dim, float split, uint32 left, uint32 right} (kdtree.hpp:149-163)
void ref_tree_copy_nodes(void *handle, void *nodes16) {
    auto nodes = static_cast<RefTree *>(handle)->tree.nodes();
    static_assert(sizeof(kdt::KDTree::KDTreeNode) == 16);
    std::memcpy(nodes16, nodes.data(), nodes.size() * 16);
}

void ref_tree_copy_points(void *handle, float *x, float *y, float *z, uint32_t *idx) {
    auto const &p = static_cast<RefTree *>(handle)->tree.positions();
    size_t n = p.size();
    std::memcpy(x, p.positions_[0], n * 4);
    std::memcpy(y, p.positions_[1], n * 4);
    std::memcpy(z, p.positions_[2], n * 4);
    std::memcpy(idx, p.indices_.data(), n * 4);
}

// Batched query, threaded like pybind.cpp:164-172.  stats (optional) = {nodes_visited,
// nodes_pruned, points_visited} summed over all queries.  Returns 0 on success.
int ref_tree_query_ex(void *handle, const float *q_aos, uint64_t m, int k, int workers, int squared,
                      float *out_d, uint32_t *out_i, uint64_t *stats);

int ref_tree_query(void *handle, const float *q_aos, uint64_t m, int k, int workers,
                   float *out_d, uint32_t *out_i, uint64_t *stats) {
    return ref_tree_query_ex(handle, q_aos, m, k, workers, 0, out_d, out_i, stats);
}

// squared != 0: rows hold the squared distances (postprocess skipped)
int ref_tree_query_ex(void *handle, const float *q_aos, uint64_t m, int k, int workers, int squared,
                      float *out_d, uint32_t *out_i, uint64_t *stats) {
    auto &t = *static_cast<RefTree *>(handle);
    if (k <= 0) {
        g_error = "k must be positive integer";
        return 1;
    }
    if (stats)
        stats[0] = stats[1] = stats[2] = 0;
    auto run = [&](size_t begin, size_t end) {
        if (t.periodic)
            query_range(t, kdt::L2PeriodicDistance<float>{t.box}, q_aos, begin, end, k, squared != 0,
                        out_d, out_i, stats);
        else
            query_range(t, kdt::L2Distance{}, q_aos, begin, end, k, squared != 0, out_d, out_i, stats);
    };
    if (workers == 1 || m == 0) {
        run(0, m);
    } else {
        wenda::thread_pool pool(workers > 0 ? workers : std::thread::hardware_concurrency());
        pool.parallelize_loop((size_t)0, (size_t)m, run);
    }
    return 0;
}

// The reference's metric helpers, exposed for the oracle-restatement tests (kdtree.hpp:20-121).
float ref_box_distance(const float *point, const float *box6, float boxsize) {
    std::array<float, 3> p = {point[0], point[1], point[2]};
    std::array<float, 6> b;
    std::copy(box6, box6 + 6, b.begin());
    if (boxsize >= 0)
        return kdt::L2PeriodicDistance<float>{boxsize}.box_distance(p, b);
    return kdt::L2Distance{}.box_distance(p, b);
}

float ref_point_distance(const float *a, const float *b, float boxsize) {
    std::array<float, 3> l = {a[0], a[1], a[2]}, r = {b[0], b[1], b[2]};
    if (boxsize >= 0)
        return kdt::L2PeriodicDistance<float>{boxsize}(l, r);
    return kdt::L2Distance{}(l, r);
}

} // extern "C"
