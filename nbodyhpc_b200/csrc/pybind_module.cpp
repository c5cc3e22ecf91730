// pybind11 module `_impl`: same class, argument names, defaults, property names and error
// messages as the reference's kdtree/src/cpp/pybind.cpp:196-216, over the B200 C ABI.
#include <cstdlib>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <optional>
#include <unordered_map>
#include <vector>

#include <sys/mman.h>

#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <kdtree/kdtree.hpp>
#include <kdtree/kdtree_utils.hpp>

namespace py = pybind11;

namespace {

// Result arrays are allocated by the library and handed to numpy through a capsule, like the
// reference's (pybind.cpp:103-104,174-188).  2 MB alignment + MADV_HUGEPAGE: a 10^8 x 8 result is
// 3.2 GB per array, and filling it through 4 KB first-touch faults costs more than the query.
//
// Big result buffers are RECYCLED: when numpy drops an array, its buffer (>= 64 MB) goes into a small
// cache instead of back to the OS, and the next result of the same size takes it.  A fresh 6.4 GB
// allocation costs ~0.15 s of page faults and zeroing per 10^8-query call -- as much as the whole
// PCIe transfer -- while memory that has been touched before is filled at copy speed.  On its way into
// the cache a buffer is also page-locked (cudaHostRegister, by a background thread: pinning 3 GB takes
// longer than a query), so that a recycled buffer is written by the copy engine directly, like the
// pinned buffers of bench.py's `e2e` leg, with no staging copy at all.  The cache holds at most
// NBK_RESULT_CACHE_MB (default 8192; 0 disables it) and only buffers the caller has released;
// NBK_RESULT_PIN=0 keeps them pageable.
class ResultCache {
  public:
    static ResultCache &instance() {
        static ResultCache *cache = new ResultCache(); // leaked on purpose: arrays may outlive static destruction
        return *cache;
    }
    static constexpr size_t kMinBytes = (size_t)64 << 20;

    void *take(size_t alloc) {
        std::lock_guard<std::mutex> lock(mutex_);
        int best = -1;
        for (size_t i = 0; i < idle_.size(); ++i)
            if (idle_[i].alloc == alloc && (best < 0 || (idle_[i].pinned && !idle_[best].pinned))) best = (int)i;
        if (best < 0) return nullptr;
        Block b = idle_[best];
        idle_.erase(idle_.begin() + best);
        held_bytes_ -= b.alloc;
        live_[b.p] = b;
        return b.p;
    }
    void remember(void *p, size_t alloc) {
        std::lock_guard<std::mutex> lock(mutex_);
        if (!live_.count(p)) live_[p] = Block{p, alloc, false};
    }
    void give_back(void *p) {
        Block b{p, 0, false};
        bool keep = false, pin = false;
        {
            std::lock_guard<std::mutex> lock(mutex_);
            auto it = live_.find(p);
            if (it != live_.end()) {
                b = it->second;
                live_.erase(it);
            }
            keep = b.alloc >= kMinBytes && held_bytes_ + b.alloc <= cap_;
            if (keep) {
                held_bytes_ += b.alloc;
                pin = pin_enabled_ && !b.pinned && !stopping_.load();
                if (pin) to_pin_.push_back(b);
                else idle_.push_back(b);
            }
        }
        if (pin) {
            start_worker();
            cv_.notify_one();
        } else if (!keep) {
            release(b);
        }
    }
    void set_device(int device) { device_.store(device); }

  private:
    struct Block {
        void *p;
        size_t alloc;
        bool pinned;
    };
    ResultCache() {
        const char *v = std::getenv("NBK_RESULT_CACHE_MB");
        cap_ = (v ? std::strtoull(v, nullptr, 10) : 8192ull) << 20;
        const char *pin = std::getenv("NBK_RESULT_PIN");
        pin_enabled_ = !(pin && pin[0] == '0');
    }
    static void release(Block const &b) {
        if (b.pinned) nbk_host_unregister(b.p);
        std::free(b.p);
    }
    void start_worker() {
        std::lock_guard<std::mutex> lock(mutex_);
        if (worker_started_) return;
        worker_started_ = true;
        // At process exit no registration may be in flight while the CUDA runtime shuts down: this handler is
        // registered after CUDA's own (it exists by the time a result is recycled), so it runs before it.
        std::atexit([] {
            ResultCache &c = ResultCache::instance();
            c.stopping_.store(true);
            std::lock_guard<std::mutex> wait_for_registration(c.pinning_);
        });
        std::thread([this] {
            while (true) {
                Block b;
                {
                    std::unique_lock<std::mutex> lock(mutex_);
                    cv_.wait(lock, [&] { return !to_pin_.empty(); });
                    b = to_pin_.front();
                    to_pin_.erase(to_pin_.begin());
                }
                {
                    std::lock_guard<std::mutex> busy(pinning_);
                    if (!stopping_.load()) b.pinned = nbk_host_register(b.p, b.alloc, device_.load()) == NBK_OK;
                }
                std::lock_guard<std::mutex> lock(mutex_);
                idle_.push_back(b);
            }
        }).detach();
    }
    std::mutex mutex_, pinning_;
    std::atomic<bool> stopping_{false};
    std::condition_variable cv_;
    std::vector<Block> idle_, to_pin_;
    std::unordered_map<void *, Block> live_;
    size_t held_bytes_ = 0, cap_ = 0;
    bool pin_enabled_ = true, worker_started_ = false;
    std::atomic<int> device_{0};
};

template <typename T> py::array_t<T> make_result(py::ssize_t rows, py::ssize_t cols) {
    constexpr size_t kHuge = (size_t)2 << 20;
    const size_t bytes = (size_t)rows * (size_t)cols * sizeof(T);
    const bool big = bytes >= kHuge;
    const size_t alloc = big ? (bytes + kHuge - 1) / kHuge * kHuge : (std::max<size_t>(bytes, 1) + 63) / 64 * 64;
    ResultCache &cache = ResultCache::instance();
    void *p = alloc >= ResultCache::kMinBytes ? cache.take(alloc) : nullptr;
    if (!p) {
        p = std::aligned_alloc(big ? kHuge : 64, alloc);
        if (!p) throw std::bad_alloc();
#ifdef MADV_HUGEPAGE
        if (big) madvise(p, alloc, MADV_HUGEPAGE);
#endif
    }
    cache.remember(p, alloc);
    py::capsule owner(p, [](void *q) { ResultCache::instance().give_back(q); });
    return py::array_t<T>({rows, cols}, static_cast<T *>(p), owner);
}

void require_n_by_3(py::array_t<float> const &a) {
    // pybind.cpp:16-18 / :96-98
    if (a.ndim() != 2 || a.shape(1) != 3) throw std::runtime_error("positions must be a 2D array of shape (N, 3)");
}

class PyKDTree : public wenda::kdtree::KDTree {
    bool periodic_;
    float box_size_;

    static nbk_tree *build(py::array_t<float, py::array::c_style | py::array::forcecast> const &points,
                           int leaf_size, std::optional<float> box_size, int device) {
        int status = NBK_OK;
        nbk_tree *h = nbk_tree_build(points.data(), static_cast<uint64_t>(points.shape(0)), leaf_size, 8,
                                     box_size.has_value() ? 1 : 0, box_size.value_or(0.0f), device, &status);
        if (status != NBK_OK) throw std::runtime_error(nbk_last_error());
        return h;
    }

    static nbk_tree *build_device(uintptr_t d_points, uint64_t n, int leaf_size, std::optional<float> box_size,
                                  int device, uintptr_t stream) {
        int status = NBK_OK;
        nbk_tree *h = nbk_tree_build_device(reinterpret_cast<const float *>(d_points), n, leaf_size, 8,
                                            box_size.has_value() ? 1 : 0, box_size.value_or(0.0f), device,
                                            reinterpret_cast<void *>(stream), &status);
        if (status != NBK_OK) throw std::runtime_error(nbk_last_error());
        return h;
    }

  public:
    PyKDTree(PyKDTree &&) noexcept = default;

    // points already resident on the device (row-major (n,3) float32): no host round trip
    PyKDTree(uintptr_t d_points, uint64_t n, int leaf_size, int max_threads, std::optional<float> box_size,
             int device, uintptr_t stream)
        : KDTree(build_device(d_points, n, leaf_size, box_size, device, stream),
                 {.leaf_size = leaf_size, .max_threads = max_threads, .block_size = 8}),
          periodic_(box_size.has_value()), box_size_(box_size.value_or(0.0f)) {}

    static PyKDTree from_device(uintptr_t d_points, uint64_t n, int leaf_size, int max_threads,
                                std::optional<float> box_size, int device, uintptr_t stream) {
        py::gil_scoped_release nogil;
        return PyKDTree(d_points, n, leaf_size, max_threads, box_size, device, stream);
    }

    PyKDTree(py::array_t<float, py::array::c_style | py::array::forcecast> const &points, int leaf_size,
             int max_threads, std::optional<float> box_size, int device)
        : KDTree(build(points, leaf_size, box_size, device),
                 {.leaf_size = leaf_size, .max_threads = max_threads, .block_size = 8}),
          periodic_(box_size.has_value()), box_size_(box_size.value_or(0.0f)) {}

    static PyKDTree from_points(py::array_t<float, py::array::c_style | py::array::forcecast> points,
                                int leaf_size, int max_threads, std::optional<float> box_size, int device) {
        require_n_by_3(points);
        py::gil_scoped_release nogil; // pybind.cpp:86
        return PyKDTree(points, leaf_size, max_threads, box_size, device);
    }

    nbk_tree_meta meta() const {
        nbk_tree_meta m;
        check(nbk_tree_get_meta(handle_, &m));
        return m;
    }
    size_t num_points() const { return meta().n_padded; } // padded count, pybind.cpp:71
    size_t num_nodes() const { return meta().n_nodes; }   // pybind.cpp:72
    bool periodic() const noexcept { return periodic_; }
    float box_size() const noexcept { return box_size_; }
    uintptr_t raw_handle() const noexcept { return reinterpret_cast<uintptr_t>(handle_); }
    int device() const noexcept { return nbk_tree_device(handle_); }

    std::pair<py::array_t<float>, py::array_t<uint32_t>>
    query(py::array_t<float, py::array::c_style | py::array::forcecast> points, int k, int workers, bool squared) {
        (void)workers; // the batch runs on the GPU; kept for call compatibility (pybind.cpp:210)
        if (k <= 0) throw std::runtime_error("k must be positive integer"); // pybind.cpp:92-94
        require_n_by_3(points);
        const uint64_t m = static_cast<uint64_t>(points.shape(0));
        ResultCache::instance().set_device(device());
        py::array_t<float> dist = make_result<float>((py::ssize_t)m, (py::ssize_t)k);
        py::array_t<uint32_t> idx = make_result<uint32_t>((py::ssize_t)m, (py::ssize_t)k);
        const float *q = points.data();
        float *od = dist.mutable_data();
        uint32_t *oi = idx.mutable_data();
        // Pieces of 2^27 queries so that Ctrl-C is honoured between them (pybind.cpp:128-133); the
        // library pipelines each piece in slices of its own.
        const uint64_t slice = 1ull << 27;
        for (uint64_t begin = 0; begin < m; begin += slice) {
            uint64_t cnt = std::min(slice, m - begin);
            int status;
            {
                py::gil_scoped_release nogil;
                status = nbk_tree_query_ex2(handle_, q + begin * 3, cnt, k, -1, 0.0f, squared ? NBK_QUERY_SQUARED : 0,
                                            od + begin * k, oi + begin * k);
            }
            if (status != NBK_OK) throw std::runtime_error(nbk_last_error());
            if (PyErr_CheckSignals() != 0) throw py::error_already_set();
        }
        return {dist, idx};
    }

    // device buffers in, device buffers out, enqueued on `stream` (no synchronisation)
    void query_device(uintptr_t d_q, uint64_t m, int k, uintptr_t d_dist, uintptr_t d_idx, uintptr_t stream,
                      bool squared) {
        if (k <= 0) throw std::runtime_error("k must be positive integer");
        py::gil_scoped_release nogil;
        check(nbk_tree_query_device_ex(handle_, reinterpret_cast<const float *>(d_q), m, k, -1, 0.0f,
                                       squared ? NBK_QUERY_SQUARED : 0, reinterpret_cast<float *>(d_dist),
                                       reinterpret_cast<uint32_t *>(d_idx), reinterpret_cast<void *>(stream)));
    }

    // numpy.histogram(dist[:, k-1], edges) for every k in ks without materialising the rows
    py::array_t<uint64_t> knn_cdf(py::array_t<float, py::array::c_style | py::array::forcecast> points,
                                  std::vector<int> ks,
                                  py::array_t<float, py::array::c_style | py::array::forcecast> edges) {
        require_n_by_3(points);
        if (edges.ndim() != 1 || edges.shape(0) < 2) throw std::runtime_error("edges must be a 1D array of at least 2 values");
        const int n_bins = static_cast<int>(edges.shape(0)) - 1;
        py::array_t<uint64_t> counts({(py::ssize_t)ks.size(), (py::ssize_t)n_bins});
        std::memset(counts.mutable_data(), 0, sizeof(uint64_t) * ks.size() * n_bins);
        int status;
        {
            py::gil_scoped_release nogil;
            status = nbk_tree_knn_cdf(handle_, points.data(), static_cast<uint64_t>(points.shape(0)), ks.data(),
                                      static_cast<int>(ks.size()), edges.data(), n_bins, counts.mutable_data());
        }
        if (status != NBK_OK) throw std::runtime_error(nbk_last_error());
        return counts;
    }

    void knn_cdf_device(uintptr_t d_q, uint64_t m, std::vector<int> ks, uintptr_t d_edges, int n_bins,
                        uintptr_t d_counts, uintptr_t stream) {
        py::gil_scoped_release nogil;
        check(nbk_tree_knn_cdf_device(handle_, reinterpret_cast<const float *>(d_q), m, ks.data(),
                                      static_cast<int>(ks.size()), reinterpret_cast<const float *>(d_edges), n_bins,
                                      reinterpret_cast<unsigned long long *>(d_counts),
                                      reinterpret_cast<void *>(stream)));
    }

    py::array nodes_array() const {
        auto n = nodes();
        py::list fields;
        py::array_t<uint8_t> raw({(py::ssize_t)(n.size() * sizeof(nbk_node))});
        std::memcpy(raw.mutable_data(), n.data(), n.size() * sizeof(nbk_node));
        return raw;
    }

    std::array<uint64_t, 3> stats(py::array_t<float, py::array::c_style | py::array::forcecast> points, int k) {
        require_n_by_3(points);
        uint64_t out[3];
        check(nbk_tree_stats(handle_, points.data(), static_cast<uint64_t>(points.shape(0)), k, -1, 0.0f, out));
        return {out[0], out[1], out[2]};
    }
};

} // namespace

PYBIND11_MODULE(_impl, m) {
    m.doc() = "Fast KD-tree for spatial data, including periodic boundary conditions (B200-native).";

    // the C++ fixture generators of include/kdtree/kdtree_utils.hpp (host only), for the tests
    m.def("_make_random_positions", [](uint32_t n, unsigned int seed, float boxsize) {
        auto pts = wenda::kdtree::make_random_position_and_index<3>(n, seed, boxsize);
        py::array_t<float> out({(py::ssize_t)n, (py::ssize_t)3});
        float *o = out.mutable_data();
        for (uint32_t i = 0; i < n; ++i)
            for (int d = 0; d < 3; ++d) o[3 * i + d] = pts[i].position[d];
        return out;
    });
    m.def("_fill_random_positions", [](uint32_t n, unsigned int seed) {
        auto pts = wenda::kdtree::fill_random_positions(n, seed);
        py::array_t<float> out({(py::ssize_t)n, (py::ssize_t)3});
        if (n) std::memcpy(out.mutable_data(), pts.data(), sizeof(float) * 3 * n);
        return out;
    });

    py::class_<PyKDTree>(m, "KDTree")
        // device-pointer overload first: an integer is never taken for an array
        .def(py::init(&PyKDTree::from_device), py::arg("device_pointer"), py::arg("n"), py::arg("leafsize") = 64,
             py::arg("max_threads") = -1, py::arg("boxsize") = std::nullopt, py::arg("device") = -1,
             py::arg("stream") = 0)
        .def(py::init(&PyKDTree::from_points), py::arg("points"), py::arg("leafsize") = 64,
             py::arg("max_threads") = -1, py::arg("boxsize") = std::nullopt, py::arg("device") = -1)
        // (squared: extension, SURVEY.md 8f-4 `return_squared`; the reference's three arguments keep their
        // names, order and defaults)
        .def("query", &PyKDTree::query, py::arg("points"), py::arg("k") = 1, py::arg("workers") = 1,
             py::arg("squared") = false)
        .def_property_readonly("n", &PyKDTree::num_points)
        .def_property_readonly("size", &PyKDTree::num_nodes)
        .def_property_readonly("periodic", &PyKDTree::periodic)
        .def_property_readonly("boxsize", &PyKDTree::box_size)
        // additions (not in the reference): access for tooling and tests
        .def_property_readonly("device", &PyKDTree::device)
        .def_property_readonly("_handle", &PyKDTree::raw_handle)
        .def("_nodes_bytes", &PyKDTree::nodes_array)
        .def("_stats", &PyKDTree::stats, py::arg("points"), py::arg("k") = 1)
        .def("_query_device", &PyKDTree::query_device, py::arg("d_q"), py::arg("m"), py::arg("k"), py::arg("d_dist"),
             py::arg("d_idx"), py::arg("stream") = 0, py::arg("squared") = false)
        .def("_knn_cdf", &PyKDTree::knn_cdf, py::arg("points"), py::arg("ks"), py::arg("edges"))
        .def("_knn_cdf_device", &PyKDTree::knn_cdf_device);
}
