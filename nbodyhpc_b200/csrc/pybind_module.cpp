// pybind11 module `_impl`: same class, argument names, defaults, property names and error
// messages as the reference's kdtree/src/cpp/pybind.cpp:196-216, over the B200 C ABI.
#include <optional>

#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <kdtree/kdtree.hpp>

namespace py = pybind11;

namespace {

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

  public:
    PyKDTree(PyKDTree &&) noexcept = default;

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
    query(py::array_t<float, py::array::c_style | py::array::forcecast> points, int k, int workers) {
        (void)workers; // the batch runs on the GPU; kept for call compatibility (pybind.cpp:210)
        if (k <= 0) throw std::runtime_error("k must be positive integer"); // pybind.cpp:92-94
        require_n_by_3(points);
        const uint64_t m = static_cast<uint64_t>(points.shape(0));
        py::array_t<float> dist({(py::ssize_t)m, (py::ssize_t)k});
        py::array_t<uint32_t> idx({(py::ssize_t)m, (py::ssize_t)k});
        const float *q = points.data();
        float *od = dist.mutable_data();
        uint32_t *oi = idx.mutable_data();
        // Slices of 2^24 queries so that Ctrl-C is honoured between them (pybind.cpp:128-133).
        const uint64_t slice = 1ull << 24;
        for (uint64_t begin = 0; begin < m; begin += slice) {
            uint64_t cnt = std::min(slice, m - begin);
            int status;
            {
                py::gil_scoped_release nogil;
                status = nbk_tree_query(handle_, q + begin * 3, cnt, k, od + begin * k, oi + begin * k);
            }
            if (status != NBK_OK) throw std::runtime_error(nbk_last_error());
            if (PyErr_CheckSignals() != 0) throw py::error_already_set();
        }
        return {dist, idx};
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

    py::class_<PyKDTree>(m, "KDTree")
        .def(py::init(&PyKDTree::from_points), py::arg("points"), py::arg("leafsize") = 64,
             py::arg("max_threads") = -1, py::arg("boxsize") = std::nullopt, py::arg("device") = -1)
        .def("query", &PyKDTree::query, py::arg("points"), py::arg("k") = 1, py::arg("workers") = 1)
        .def_property_readonly("n", &PyKDTree::num_points)
        .def_property_readonly("size", &PyKDTree::num_nodes)
        .def_property_readonly("periodic", &PyKDTree::periodic)
        .def_property_readonly("boxsize", &PyKDTree::box_size)
        // additions (not in the reference): access for tooling and tests
        .def_property_readonly("device", &PyKDTree::device)
        .def_property_readonly("_handle", &PyKDTree::raw_handle)
        .def("_nodes_bytes", &PyKDTree::nodes_array)
        .def("_stats", &PyKDTree::stats, py::arg("points"), py::arg("k") = 1);
}
