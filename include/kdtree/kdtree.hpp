// Drop-in C++ entry points of nbodyhpc's kdtree package, re-based on the B200 C ABI (nbk.h).
//
// Mirrors the reference's public header kdtree/src/cpp/include/kdtree/kdtree.hpp: same namespace,
// type names, member names, defaults and exception behaviour
//   L2Distance                     kdtree.hpp:20-62
//   L2PeriodicDistance<T>          kdtree.hpp:66-121
//   KDTreeQueryStatistics          kdtree.hpp:124-131
//   KDTreeConfiguration            kdtree.hpp:134-141
//   KDTree, KDTree::KDTreeNode     kdtree.hpp:144-211
// but the tree is built by, lives on and is searched by the GPU.  Header-only; link with -lnbk.
// Errors from the C ABI surface as std::runtime_error carrying the reference's messages
// (kdtree.cpp:98-108).  There is no CPU fallback: without a B200 every constructor throws.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <mutex>
#include <span>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../nbk.h"
#include "../span.hpp" // tcb::span, the name the reference's signatures use (kdtree.hpp:11)
#include "position_array.hpp"

namespace wenda {
namespace kdtree {

//! Squared Euclidean distance, float32, evaluated term by term without FMA (kdtree.hpp:20-62).
struct L2Distance {
    static constexpr bool is_periodic = false;
    float box_size() const { return 0.0f; }

    template <typename T, size_t R>
    T operator()(std::array<T, R> const &a, std::array<T, R> const &b) const {
        T acc = 0;
        for (size_t i = 0; i < R; ++i) {
            T d = a[i] - b[i];
            acc += d * d;
        }
        return acc;
    }

    //! Squared distance from `point` to the box {lo0, hi0, lo1, hi1, ...}.
    template <typename T, size_t R>
    T box_distance(std::array<T, R> const &point, std::array<T, 2 * R> const &box) const {
        T acc = 0;
        for (size_t i = 0; i < R; ++i) {
            T below = std::max(box[2 * i] - point[i], T{0});
            T above = std::max(point[i] - box[2 * i + 1], T{0});
            acc += below * below + above * above;
        }
        return acc;
    }

    template <typename T> T postprocess(T value) const { return std::sqrt(value); }

    template <typename T, size_t R> std::array<T, 2 * R> initial_box(std::array<T, R> const &) const {
        std::array<T, 2 * R> box;
        for (size_t i = 0; i < R; ++i) {
            box[2 * i] = std::numeric_limits<T>::lowest();
            box[2 * i + 1] = std::numeric_limits<T>::max();
        }
        return box;
    }
};

//! Squared distance under periodic boundaries: per axis the smallest of the three images
//! d, d + L, d - L (kdtree.hpp:66-121).
template <typename T> struct L2PeriodicDistance {
    static constexpr bool is_periodic = true;
    T box_size_;
    float box_size() const { return static_cast<float>(box_size_); }

    template <size_t R> T operator()(std::array<T, R> const &a, std::array<T, R> const &b) const {
        T acc = 0;
        for (size_t i = 0; i < R; ++i) {
            T d = a[i] - b[i];
            T up = d + box_size_, down = d - box_size_;
            acc += std::min({d * d, up * up, down * down});
        }
        return acc;
    }

    //! Box must lie inside [0, L] and not straddle the boundary.
    template <size_t R>
    T box_distance(std::array<T, R> const &point, std::array<T, 2 * R> const &box) const {
        T acc = 0;
        for (size_t i = 0; i < R; ++i) {
            T lo = box[2 * i], hi = box[2 * i + 1], p = point[i];
            if (p < lo) {
                T m = std::min(lo - p, p + box_size_ - hi);
                acc += m * m;
            } else if (p > hi) {
                T m = std::min(p - hi, lo + box_size_ - p);
                acc += m * m;
            }
        }
        return acc;
    }

    T postprocess(T value) const { return std::sqrt(value); }

    template <size_t R> std::array<T, 2 * R> initial_box(std::array<T, R> const &) const {
        std::array<T, 2 * R> box;
        for (size_t i = 0; i < R; ++i) {
            box[2 * i] = 0;
            box[2 * i + 1] = box_size_;
        }
        return box;
    }
};

//! Counters of the reference's closer-first traversal for one query (kdtree.hpp:124-131).
struct KDTreeQueryStatistics {
    size_t nodes_visited;
    size_t nodes_pruned;
    size_t points_visited;
};

//! Build configuration (kdtree.hpp:134-141).  max_threads is accepted and ignored: the build runs
//! on the GPU.
struct KDTreeConfiguration {
    int leaf_size = 64;
    int max_threads = 0;
    int block_size = 8;
};

class KDTree {
  public:
    //! Byte-identical to nbk_node and to the reference's node record (kdtree.hpp:149-163).
    struct KDTreeNode {
        int dimension_;
        float split_;
        uint32_t left_;
        uint32_t right_;
    };
    static_assert(sizeof(KDTreeNode) == sizeof(nbk_node), "node layout");

    //! Builds from AoS positions; pads to a multiple of block_size with FLT_MAX points
    //! (kdtree.cpp:64-93).
    KDTree(tcb::span<const std::array<float, 3>> positions, KDTreeConfiguration const &config = {})
        : config_(config) {
        int status = NBK_OK;
        handle_ = nbk_tree_build(positions.empty() ? nullptr : positions.data()->data(), positions.size(),
                                 config.leaf_size, config.block_size, 0, 0.0f, -1, &status);
        check(status);
    }

    //! Builds from pre-padded SoA positions with caller-chosen indices (kdtree.cpp:95-131).
    KDTree(PositionAndIndexArray<3> positions, KDTreeConfiguration const &config = {}) : config_(config) {
        int status = NBK_OK;
        handle_ = nbk_tree_build_soa(positions.positions_[0], positions.positions_[1], positions.positions_[2],
                                     positions.indices_.data(), positions.size(), config.leaf_size,
                                     config.block_size, 0, 0.0f, -1, &status);
        check(status);
    }

    KDTree(KDTree &&other) noexcept
        : handle_(other.handle_), config_(other.config_), nodes_(std::move(other.nodes_)),
          positions_(std::move(other.positions_)), have_nodes_(other.have_nodes_),
          have_positions_(other.have_positions_) {
        other.handle_ = nullptr;
    }
    KDTree(KDTree const &) = delete;
    KDTree &operator=(KDTree const &) = delete;
    ~KDTree() { nbk_tree_free(handle_); }

    //! Node array in the reference's pre-order numbering (copied from the device on first use).
    tcb::span<const KDTreeNode> nodes() const {
        std::lock_guard<std::mutex> lock(mutex_);
        if (!have_nodes_) {
            nbk_tree_meta meta;
            check(nbk_tree_get_meta(handle_, &meta));
            nodes_.resize(meta.n_nodes);
            check(nbk_tree_copy_nodes(handle_, reinterpret_cast<nbk_node *>(nodes_.data())));
            have_nodes_ = true;
        }
        return nodes_;
    }

    //! Leaf-ordered positions and the permutation (copied from the device on first use).
    PositionAndIndexArray<3> const &positions() const {
        std::lock_guard<std::mutex> lock(mutex_);
        if (!have_positions_) {
            nbk_tree_meta meta;
            check(nbk_tree_get_meta(handle_, &meta));
            positions_ = PositionAndIndexArray<3>(meta.n_padded);
            check(nbk_tree_copy_points(handle_, positions_.positions_[0], positions_.positions_[1],
                                       positions_.positions_[2], positions_.indices_.data()));
            have_positions_ = true;
        }
        return positions_;
    }

    KDTreeConfiguration const &config() const noexcept { return config_; }

    //! Opaque C-ABI handle (for batched queries through nbk_tree_query).
    nbk_tree const *handle() const noexcept { return handle_; }

    //! k nearest neighbours of one position, ascending, as (distance, original index); a batch of
    //! one on the GPU (kdtree.hpp:207-210, kdtree.cpp:133-159).  Distance is L2Distance or
    //! L2PeriodicDistance<float>.
    template <typename Distance>
    std::vector<std::pair<float, uint32_t>>
    find_closest(std::array<float, 3> const &position, size_t k, Distance const &distance = {},
                 KDTreeQueryStatistics *statistics = nullptr) const {
        std::vector<float> d(k);
        std::vector<uint32_t> i(k);
        check(nbk_tree_query_ex(handle_, position.data(), 1, static_cast<int>(k), Distance::is_periodic ? 1 : 0,
                                distance.box_size(), d.data(), i.data()));
        if (statistics) {
            uint64_t s[3];
            check(nbk_tree_stats(handle_, position.data(), 1, static_cast<int>(k), Distance::is_periodic ? 1 : 0,
                                 distance.box_size(), s));
            statistics->nodes_visited = s[0];
            statistics->nodes_pruned = s[1];
            statistics->points_visited = s[2];
        }
        std::vector<std::pair<float, uint32_t>> result(k);
        for (size_t j = 0; j < k; ++j) result[j] = {d[j], i[j]};
        return result;
    }

    //! Batched form of find_closest (extension; what PyKDTree::query does per row, pybind.cpp:111-161):
    //! out_dist / out_idx are positions.size() x k, row-major.  If `statistics` is given it receives
    //! the counters SUMMED over the batch.
    template <typename Distance>
    void find_closest_batch(tcb::span<const std::array<float, 3>> positions, size_t k, float *out_dist,
                            uint32_t *out_idx, Distance const &distance = {},
                            KDTreeQueryStatistics *statistics = nullptr) const {
        const float *q = positions.empty() ? nullptr : positions.data()->data();
        check(nbk_tree_query_ex(handle_, q, positions.size(), static_cast<int>(k), Distance::is_periodic ? 1 : 0,
                                distance.box_size(), out_dist, out_idx));
        if (statistics) {
            uint64_t s[3];
            check(nbk_tree_stats(handle_, q, positions.size(), static_cast<int>(k), Distance::is_periodic ? 1 : 0,
                                 distance.box_size(), s));
            statistics->nodes_visited = s[0];
            statistics->nodes_pruned = s[1];
            statistics->points_visited = s[2];
        }
    }

  protected:
    //! For wrappers that build the handle themselves (the pybind layer).
    KDTree(nbk_tree *handle, KDTreeConfiguration const &config) : handle_(handle), config_(config) {}

    static void check(int status) {
        if (status != NBK_OK) throw std::runtime_error(nbk_last_error());
    }

    nbk_tree *handle_ = nullptr;
    KDTreeConfiguration config_;

  private:
    mutable std::mutex mutex_;
    mutable std::vector<KDTreeNode> nodes_;
    mutable PositionAndIndexArray<3> positions_;
    mutable bool have_nodes_ = false, have_positions_ = false;
};

} // namespace kdtree
} // namespace wenda
