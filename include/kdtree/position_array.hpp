// Host-side SoA container of the drop-in C++ API.
//
// Same public surface as the reference's wenda::kdtree::PositionAndIndexArray<R,T,IndexT>
// (position_array.hpp:166-271: `positions_` = R column pointers, 64-byte aligned; `indices_`;
// size(); operator[]; swap_elements) so code that fills or inspects one keeps compiling.  Here it
// is only a staging/inspection buffer: the tree itself lives in HBM (see include/nbk.h).
#pragma once

#include <algorithm>
#include <array>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <iterator>
#include <utility>
#include <vector>

namespace wenda {
namespace kdtree {

template <size_t R = 3, typename T = float> struct PositionAndIndex {
    std::array<T, R> position;
    uint32_t index;
};

template <size_t R = 3, typename T = float, typename IndexT = uint32_t> struct PositionAndIndexArray {
    static const size_t dimension = R;
    typedef T element_type;
    typedef PositionAndIndex<R, T> value_type;

    std::array<T *, R> positions_{};
    std::vector<IndexT> indices_;

    PositionAndIndexArray() = default;

    explicit PositionAndIndexArray(size_t n) : indices_(n) { allocate(n); }

    template <typename Container,
              typename std::enable_if<!std::is_integral<Container>::value, bool>::type = true>
    explicit PositionAndIndexArray(Container const &points) : PositionAndIndexArray(std::size(points)) {
        size_t i = 0;
        for (auto const &p : points) {
            for (size_t d = 0; d < R; ++d) positions_[d][i] = p.position[d];
            indices_[i] = p.index;
            ++i;
        }
    }

    PositionAndIndexArray(PositionAndIndexArray const &other) : indices_(other.indices_) {
        allocate(indices_.size());
        for (size_t d = 0; d < R; ++d)
            std::copy(other.positions_[d], other.positions_[d] + indices_.size(), positions_[d]);
    }

    PositionAndIndexArray(PositionAndIndexArray &&other) noexcept
        : positions_(other.positions_), indices_(std::move(other.indices_)) {
        other.positions_.fill(nullptr);
    }

    PositionAndIndexArray &operator=(PositionAndIndexArray const &) = delete;

    PositionAndIndexArray &operator=(PositionAndIndexArray &&other) noexcept {
        std::swap(positions_, other.positions_);
        std::swap(indices_, other.indices_);
        return *this;
    }

    ~PositionAndIndexArray() noexcept {
        for (auto &p : positions_) {
            std::free(p);
            p = nullptr;
        }
    }

    size_t size() const noexcept { return indices_.size(); }

    value_type operator[](size_t i) const noexcept {
        value_type v;
        for (size_t d = 0; d < R; ++d) v.position[d] = positions_[d][i];
        v.index = static_cast<uint32_t>(indices_[i]);
        return v;
    }

    void swap_elements(size_t i, size_t j) noexcept {
        std::swap(indices_[i], indices_[j]);
        for (size_t d = 0; d < R; ++d) std::swap(positions_[d][i], positions_[d][j]);
    }

    // Read-only forward iteration over value_type (enough for range-for and the test helpers).
    struct const_iterator {
        typedef std::forward_iterator_tag iterator_category;
        typedef PositionAndIndex<R, T> value_type;
        typedef std::ptrdiff_t difference_type;
        typedef value_type const *pointer;
        typedef value_type reference;
        PositionAndIndexArray const *array;
        size_t offset;
        value_type operator*() const { return (*array)[offset]; }
        const_iterator &operator++() { ++offset; return *this; }
        const_iterator operator++(int) { const_iterator c = *this; ++offset; return c; }
        bool operator==(const_iterator const &o) const { return offset == o.offset; }
        bool operator!=(const_iterator const &o) const { return offset != o.offset; }
    };
    const_iterator begin() const noexcept { return {this, 0}; }
    const_iterator end() const noexcept { return {this, size()}; }

  private:
    void allocate(size_t n) {
        size_t bytes = (sizeof(T) * n + 63) / 64 * 64;
        if (bytes == 0) bytes = 64;
        for (size_t d = 0; d < R; ++d) positions_[d] = static_cast<T *>(std::aligned_alloc(64, bytes));
    }
};

} // namespace kdtree
} // namespace wenda
