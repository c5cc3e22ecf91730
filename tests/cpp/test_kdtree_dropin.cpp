// The reference's C++ test cases (kdtree/src/cpp/tests/test.cpp:43-114) replayed against the drop-in
// headers: same fixture calls, same tree configurations, same queries, same pass criterion (the
// (distance, index) pairs equal an exhaustive scan's, exactly).  Plain main() instead of gtest
// (gtest is a network fetch in the reference's build); exit code = number of failed checks.
#include <algorithm>
#include <cstdio>
#include <limits>
#include <vector>

#include <kdtree/kdtree.hpp>
#include <kdtree/kdtree_utils.hpp>

namespace wk = wenda::kdtree;

namespace {

int failures = 0;
#define CHECK(cond)                                                                         \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                   \
            ++failures;                                                                     \
        }                                                                                   \
    } while (0)

// exhaustive k nearest under `distance` (squared metric + postprocess), ascending (distance, index)
template <typename Positions, typename Distance>
std::vector<std::pair<float, uint32_t>> scan_all(Positions const &positions, std::array<float, 3> const &query,
                                                  size_t k, Distance const &distance) {
    std::vector<std::pair<float, uint32_t>> all;
    for (auto const &p : positions) {
        const float d = distance(p.position, query);
        if (d < std::numeric_limits<float>::max()) all.emplace_back(d, p.index); // padding is at +inf
    }
    std::sort(all.begin(), all.end());
    all.resize(std::min(all.size(), k));
    for (auto &e : all) e.first = distance.postprocess(e.first);
    return all;
}

bool leaves_are_block_multiples(wk::KDTree const &tree, uint32_t block) {
    auto nodes = tree.nodes();
    return std::all_of(nodes.begin(), nodes.end(),
                       [block](auto const &n) { return n.dimension_ != -1 || (n.right_ - n.left_) % block == 0; });
}

void build_and_find_nearest(uint32_t n) { // test.cpp:43-65
    const int block_size = 8;
    const std::array<float, 3> query = {0.4f, 0.5f, 0.6f};
    auto positions = wk::make_random_position_and_index_array(n, 42, 1.0, block_size);
    auto expected = scan_all(positions, query, 4, wk::L2Distance{});
    wk::KDTree tree(std::move(positions), {.leaf_size = 32, .block_size = block_size});
    CHECK(leaves_are_block_multiples(tree, block_size));
    wk::KDTreeQueryStatistics statistics;
    auto result = tree.find_closest(query, 4, wk::L2Distance{}, &statistics);
    CHECK(std::is_sorted(result.begin(), result.end()));
    CHECK(result == expected);
    CHECK(statistics.nodes_visited >= 1 && statistics.points_visited >= 8);
}

void build_and_find_nearest_default_config(uint32_t n) { // test.cpp:67-87
    wk::KDTreeConfiguration config{};
    const std::array<float, 3> query = {0.5f, 0.5f, 0.5f};
    auto positions = wk::make_random_position_and_index_array(n, 42, 1.0, config.block_size);
    auto expected = scan_all(positions, query, 4, wk::L2Distance{});
    wk::KDTree tree(std::move(positions), config);
    CHECK(leaves_are_block_multiples(tree, 8));
    auto result = tree.find_closest(query, 4, wk::L2Distance{});
    CHECK(std::is_sorted(result.begin(), result.end()));
    CHECK(result == expected);
}

void build_and_find_nearest_periodic(uint32_t n) { // test.cpp:89-111
    wk::KDTreeConfiguration config{};
    const float boxsize = 2.0f;
    auto positions = wk::make_random_position_and_index_array(n, 42, boxsize, config.block_size);
    auto queries = wk::make_random_position_and_index(100, 43, boxsize);
    const wk::L2PeriodicDistance<float> distance{boxsize};
    std::vector<std::vector<std::pair<float, uint32_t>>> expected;
    for (auto const &q : queries) expected.push_back(scan_all(positions, q.position, 4, distance));
    wk::KDTree tree(std::move(positions), config);
    for (size_t i = 0; i < queries.size(); ++i) {
        auto result = tree.find_closest(queries[i].position, 4, distance);
        CHECK(std::is_sorted(result.begin(), result.end()));
        CHECK(result == expected[i]);
    }
    // the batched form gives the same rows
    std::vector<std::array<float, 3>> q(queries.size());
    for (size_t i = 0; i < q.size(); ++i) q[i] = queries[i].position;
    std::vector<float> d(q.size() * 4);
    std::vector<uint32_t> idx(q.size() * 4);
    tree.find_closest_batch(q, 4, d.data(), idx.data(), distance);
    for (size_t i = 0; i < q.size(); ++i)
        for (size_t j = 0; j < expected[i].size(); ++j)
            CHECK(d[4 * i + j] == expected[i][j].first && idx[4 * i + j] == expected[i][j].second);
}

void span_constructor_and_errors() { // kdtree.cpp:64-108
    auto pts = wk::fill_random_positions(1000, 42);
    wk::KDTree tree(tcb::span<const std::array<float, 3>>(pts.data(), pts.size()));
    auto r = tree.find_closest(pts[17], 1, wk::L2Distance{});
    CHECK(r.size() == 1 && r[0].first == 0.0f && r[0].second == 17);
    CHECK(tree.positions().size() == 1000);
    bool threw = false;
    try {
        auto bad = wk::make_random_position_and_index_array(100, 1, 1.0, -1); // 100 is not a multiple of 8
        wk::KDTree t(std::move(bad), {.leaf_size = 32, .block_size = 8});
    } catch (std::runtime_error const &e) {
        threw = std::string(e.what()) == "block_size must divide the number of points.";
    }
    CHECK(threw);
}

void any_k_and_position_helper() { // kdtree.cpp:64-90 (make_position_and_indices), :133-141 (any k)
    auto pts = wk::fill_random_positions(5000, 9);
    auto soa = wk::make_position_and_indices(tcb::span<const std::array<float, 3>>(pts.data(), pts.size()), 8);
    CHECK(soa.size() == 5000 && soa.indices_[4999] == 4999);
    const std::array<float, 3> query = {0.3f, 0.6f, 0.9f};
    for (size_t k : {65u, 200u}) {
        auto expected = scan_all(soa, query, k, wk::L2Distance{});
        wk::KDTree tree(wk::make_position_and_indices(tcb::span<const std::array<float, 3>>(pts.data(), pts.size()), 8));
        auto result = tree.find_closest(query, k, wk::L2Distance{});
        CHECK(result.size() == k && result == expected);
        const wk::L2PeriodicDistance<float> periodic{1.0f};
        CHECK(tree.find_closest(query, k, periodic) == scan_all(soa, query, k, periodic));
    }
}

} // namespace

int main() {
    for (uint32_t n : {10u, 100u, 1000u}) { // test.cpp:113-114
        build_and_find_nearest(n);
        build_and_find_nearest_default_config(n);
        build_and_find_nearest_periodic(n);
    }
    span_constructor_and_errors();
    any_k_and_position_helper();
    std::printf(failures ? "%d check(s) failed\n" : "all checks passed\n", failures);
    return failures;
}
