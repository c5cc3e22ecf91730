// Host-only checks of the drop-in C++ containers and helpers (no GPU needed):
//   PositionAndIndexArray proxy / random-access iterators   (reference position_array.hpp:53-161,273-352)
//   OffsetRangeContainerWrapper                              (position_array.hpp:26-46)
//   make_position_and_indices                                (kdtree.cpp:64-90, kdtree_utils.hpp:117-118)
//   <span.hpp> / tcb::span                                   (kdtree.hpp:11)
// STL and ranges algorithms must run over the SoA container the way they do over the reference's
// (its selection code sorts and partitions through these iterators, tests/test_floyd_rivest.cpp).
#include <span.hpp>

#include <algorithm>
#include <cstdio>
#include <iterator>
#include <numeric>
#include <vector>

#include <kdtree/kdtree_utils.hpp>
#include <kdtree/position_array.hpp>

namespace wk = wenda::kdtree;

static int failures = 0;
#define CHECK(cond)                                                       \
    do {                                                                  \
        if (!(cond)) {                                                    \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            ++failures;                                                   \
        }                                                                 \
    } while (0)

using Array = wk::PositionAndIndexArray<3, float, uint32_t>;
static_assert(std::random_access_iterator<Array::iterator>);
static_assert(std::random_access_iterator<Array::const_iterator>);
static_assert(std::sortable<Array::iterator, std::ranges::less, float (*)(wk::PositionAndIndex<3, float> const &)>);
static_assert(std::is_same_v<std::iter_value_t<Array::iterator>, wk::PositionAndIndex<3, float>>);

static float x_of(wk::PositionAndIndex<3, float> const &p) { return p.position[0]; }

int main() {
    // ---- make_position_and_indices: columns, iota indices, max() padding --------------------------
    std::vector<std::array<float, 3>> pts = wk::fill_random_positions(21, 7);
    tcb::span<const std::array<float, 3>> view(pts.data(), pts.size());
    Array soa = wk::make_position_and_indices(view, 8);
    CHECK(soa.size() == 24);
    for (size_t i = 0; i < 21; ++i)
        for (size_t d = 0; d < 3; ++d) CHECK(soa.positions_[d][i] == pts[i][d]);
    for (size_t i = 21; i < 24; ++i)
        for (size_t d = 0; d < 3; ++d) CHECK(soa.positions_[d][i] == std::numeric_limits<float>::max());
    for (size_t i = 0; i < 24; ++i) CHECK(soa.indices_[i] == i);
    CHECK(wk::make_position_and_indices(view).size() == 21);     // block_size <= 0: no padding
    CHECK(wk::make_position_and_indices(view, 7).size() == 21);  // already a multiple
    for (size_t d = 0; d < 3; ++d) CHECK(reinterpret_cast<uintptr_t>(soa.positions_[d]) % 64 == 0);

    // ---- proxies write through; const access yields values -----------------------------------------
    Array a = wk::make_random_position_and_index_array(100, 42, 1.0, 8); // 104 rows
    const Array &ca = a;
    wk::PositionAndIndex<3, float> v = ca[5];
    a[6] = v;
    CHECK(a.positions_[1][6] == a.positions_[1][5] && a.indices_[6] == 5);
    a[6] = a[7];
    CHECK(a.positions_[2][6] == a.positions_[2][7] && a.indices_[6] == 7);
    a.indices_[6] = 6;
    using std::swap;
    const float x0 = a.positions_[0][0], x1 = a.positions_[0][1];
    swap(a[0], a[1]);
    CHECK(a.positions_[0][0] == x1 && a.positions_[0][1] == x0 && a.indices_[0] == 1 && a.indices_[1] == 0);
    std::ranges::iter_swap(a.begin(), a.begin() + 1);
    CHECK(a.positions_[0][0] == x0 && a.indices_[0] == 0);

    // ---- iterator arithmetic -----------------------------------------------------------------------
    auto it = a.begin();
    CHECK(a.end() - it == 104 && (it + 3)[2].index == 5);
    CHECK((*(3 + it)).index == 3 && ca.begin()[4].index == 4);
    CHECK((it += 10) - a.begin() == 10 && it > a.begin() && it <= a.end() && --it == a.begin() + 9);
    CHECK(std::distance(ca.begin(), ca.end()) == 104);
    CHECK(std::count_if(ca.begin(), ca.end(), [](auto const &p) { return p.position[0] < 0.5f; }) > 20);

    // ---- algorithms over the SoA rows: whole tuples move together -----------------------------------
    std::vector<wk::PositionAndIndex<3, float>> before(ca.begin(), ca.end());
    std::ranges::sort(a, std::ranges::less{}, x_of);
    CHECK(std::is_sorted(a.positions_[0], a.positions_[0] + a.size()));
    for (size_t i = 0; i < a.size(); ++i) {
        auto const &orig = before[a.indices_[i]];
        CHECK(orig.position[0] == a.positions_[0][i] && orig.position[1] == a.positions_[1][i] &&
              orig.position[2] == a.positions_[2][i]);
    }
    std::ranges::nth_element(a.begin(), a.begin() + 40, a.end(), std::ranges::less{},
                             [](wk::PositionAndIndex<3, float> const &p) { return p.position[1]; });
    const float pivot = a.positions_[1][40];
    CHECK(std::all_of(a.positions_[1], a.positions_[1] + 40, [&](float y) { return y <= pivot; }));
    CHECK(std::all_of(a.positions_[1] + 41, a.positions_[1] + 104, [&](float y) { return y >= pivot; }));
    std::vector<uint32_t> seen(a.indices_);
    std::sort(seen.begin(), seen.end());
    for (size_t i = 0; i < seen.size(); ++i) CHECK(seen[i] == i);

    // ---- a window of the container ---------------------------------------------------------------------
    wk::OffsetRangeContainerWrapper<Array> window(a, 8, 16);
    CHECK(window.size() == 16 && window.end() - window.begin() == 16);
    CHECK(window[3].index == a.indices_[11]);
    std::ranges::sort(window.begin(), window.end(), std::ranges::less{}, x_of);
    CHECK(std::is_sorted(a.positions_[0] + 8, a.positions_[0] + 24));
    wk::OffsetRangeContainerWrapper<Array> all(a);
    CHECK(all.size() == a.size());

    // ---- copies are deep, moves steal --------------------------------------------------------------------
    Array copy(a);
    copy[0] = copy[1];
    CHECK(copy.positions_[0] != a.positions_[0] && a.indices_[0] != a.indices_[1]);
    Array moved(std::move(copy));
    CHECK(copy.positions_[0] == nullptr && moved.size() == 104);

    std::printf(failures ? "%d check(s) failed\n" : "all checks passed\n", failures);
    return failures;
}
