// Host-side SoA container of the drop-in C++ API.
//
// Same public surface as the reference's wenda::kdtree::PositionAndIndexArray<R,T,IndexT>
// (position_array.hpp:166-271: `positions_` = R column pointers, 64-byte aligned; `indices_`;
// size(); operator[] (proxy / value); begin()/end() as random-access proxy iterators; swap_elements)
// plus OffsetRangeContainerWrapper (position_array.hpp:26-46) and the iterator customisation points
// iter_swap / iter_move / swap (position_array.hpp:327-352), so code that fills, inspects or runs
// STL / ranges algorithms over one keeps compiling.  Here it is only a staging/inspection buffer:
// the tree itself lives in HBM (see include/nbk.h).
#pragma once

#include <algorithm>
#include <array>
#include <compare>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <iterator>
#include <type_traits>
#include <utility>
#include <vector>

namespace wenda {
namespace kdtree {

template <size_t R = 3, typename T = float> struct PositionAndIndex {
    std::array<T, R> position;
    uint32_t index;
};

//! A window [offset, offset + count) of a random-access container (position_array.hpp:26-46).
template <typename Container> struct OffsetRangeContainerWrapper {
    Container &container_;
    size_t offset;
    size_t count;

    OffsetRangeContainerWrapper(Container &c) : container_(c), offset(0), count(c.size()) {}
    OffsetRangeContainerWrapper(Container &c, size_t first, size_t n) : container_(c), offset(first), count(n) {}

    decltype(auto) begin() { return container_.begin() + offset; }
    decltype(auto) begin() const { return std::as_const(container_).begin() + offset; }
    decltype(auto) end() { return begin() + count; }
    decltype(auto) end() const { return begin() + count; }
    decltype(auto) operator[](size_t i) { return container_[offset + i]; }
    decltype(auto) operator[](size_t i) const { return std::as_const(container_)[offset + i]; }
    size_t size() const { return count; }
};

template <size_t R, typename T, typename IndexT> struct PositionAndIndexArray;

namespace detail {

//! Writable stand-in for one element of the SoA array: references into the R columns and the
//! index column (position_array.hpp:130-161).  Assigning through it writes the columns.
template <size_t R, typename T, typename IndexT> struct PositionAndIndexProxy {
    std::array<std::reference_wrapper<T>, R> position;
    IndexT &index;

    operator PositionAndIndex<R, T>() const noexcept {
        PositionAndIndex<R, T> v;
        for (size_t d = 0; d < R; ++d) v.position[d] = position[d].get();
        v.index = static_cast<uint32_t>(index);
        return v;
    }
    PositionAndIndexProxy const &operator=(PositionAndIndex<R, T> const &v) const noexcept {
        for (size_t d = 0; d < R; ++d) position[d].get() = v.position[d];
        index = static_cast<IndexT>(v.index);
        return *this;
    }
    PositionAndIndexProxy const &operator=(PositionAndIndexProxy const &o) const noexcept {
        for (size_t d = 0; d < R; ++d) position[d].get() = o.position[d].get();
        index = o.index;
        return *this;
    }
};

//! Random-access iterator over an SoA array: (array, offset).  Const = true yields values,
//! Const = false yields proxies.  A proxy iterator in the C++20 sense (std::random_access_iterator
//! holds); the legacy category tag is kept because the reference's has it (position_array.hpp:53-125,277-325).
template <size_t R, typename T, typename IndexT, bool Const> struct SoaIterator {
    using Array = std::conditional_t<Const, PositionAndIndexArray<R, T, IndexT> const, PositionAndIndexArray<R, T, IndexT>>;
    using iterator_category = std::random_access_iterator_tag;
    using iterator_concept = std::random_access_iterator_tag;
    using difference_type = std::ptrdiff_t;
    using value_type = PositionAndIndex<R, T>;
    using reference = std::conditional_t<Const, value_type, PositionAndIndexProxy<R, T, IndexT>>;
    using pointer = void;

    std::ptrdiff_t offset_ = 0;
    Array *array_ = nullptr;

    SoaIterator() = default;
    SoaIterator(Array *array, std::ptrdiff_t offset) : offset_(offset), array_(array) {}

    reference operator*() const { return (*array_)[static_cast<size_t>(offset_)]; }
    reference operator[](difference_type n) const { return (*array_)[static_cast<size_t>(offset_ + n)]; }

    SoaIterator &operator++() { ++offset_; return *this; }
    SoaIterator operator++(int) { SoaIterator c = *this; ++offset_; return c; }
    SoaIterator &operator--() { --offset_; return *this; }
    SoaIterator operator--(int) { SoaIterator c = *this; --offset_; return c; }
    SoaIterator &operator+=(difference_type n) { offset_ += n; return *this; }
    SoaIterator &operator-=(difference_type n) { offset_ -= n; return *this; }
    friend SoaIterator operator+(SoaIterator it, difference_type n) { return it += n; }
    friend SoaIterator operator+(difference_type n, SoaIterator it) { return it += n; }
    friend SoaIterator operator-(SoaIterator it, difference_type n) { return it -= n; }
    friend difference_type operator-(SoaIterator const &a, SoaIterator const &b) { return a.offset_ - b.offset_; }
    friend bool operator==(SoaIterator const &a, SoaIterator const &b) { return a.offset_ == b.offset_; }
    friend auto operator<=>(SoaIterator const &a, SoaIterator const &b) { return a.offset_ <=> b.offset_; }

    //! std::ranges::iter_swap / iter_move customisation points (position_array.hpp:327-339)
    friend void iter_swap(SoaIterator const &a, SoaIterator const &b) noexcept
        requires(!Const)
    {
        a.array_->swap_elements(static_cast<size_t>(a.offset_), static_cast<size_t>(b.offset_));
    }
    friend value_type iter_move(SoaIterator const &it) {
        return std::as_const(*it.array_)[static_cast<size_t>(it.offset_)];
    }
};

template <size_t R, typename T, typename IndexT> using PositionAndIndexIterator = SoaIterator<R, T, IndexT, false>;
template <size_t R, typename T, typename IndexT> using ConstPositionAndIndexIterator = SoaIterator<R, T, IndexT, true>;

//! Swapping two proxies swaps the elements they stand for (position_array.hpp:341-350).
template <size_t R, typename T, typename IndexT>
void swap(PositionAndIndexProxy<R, T, IndexT> a, PositionAndIndexProxy<R, T, IndexT> b) {
    for (size_t d = 0; d < R; ++d) std::swap(a.position[d].get(), b.position[d].get());
    std::swap(a.index, b.index);
}

} // namespace detail

template <size_t R = 3, typename T = float, typename IndexT = uint32_t> struct PositionAndIndexArray {
    static const size_t dimension = R;
    typedef T element_type;
    typedef PositionAndIndex<R, T> value_type;
    typedef detail::PositionAndIndexIterator<R, T, IndexT> iterator;
    typedef detail::ConstPositionAndIndexIterator<R, T, IndexT> const_iterator;
    typedef detail::PositionAndIndexProxy<R, T, IndexT> PositionAndIndexProxy;

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

    //! mutable element access goes through a proxy that writes the columns
    PositionAndIndexProxy operator[](size_t i) noexcept {
        return make_proxy(i, std::make_index_sequence<R>{});
    }

    iterator begin() noexcept { return iterator(this, 0); }
    iterator end() noexcept { return iterator(this, static_cast<std::ptrdiff_t>(size())); }
    const_iterator begin() const noexcept { return const_iterator(this, 0); }
    const_iterator end() const noexcept { return const_iterator(this, static_cast<std::ptrdiff_t>(size())); }
    const_iterator cbegin() const noexcept { return begin(); }
    const_iterator cend() const noexcept { return end(); }

  private:
    template <size_t... D> PositionAndIndexProxy make_proxy(size_t i, std::index_sequence<D...>) noexcept {
        return PositionAndIndexProxy{{std::ref(positions_[D][i])...}, indices_[i]};
    }
    void allocate(size_t n) {
        size_t bytes = (sizeof(T) * n + 63) / 64 * 64;
        if (bytes == 0) bytes = 64;
        for (size_t d = 0; d < R; ++d) positions_[d] = static_cast<T *>(std::aligned_alloc(64, bytes));
    }
};

} // namespace kdtree
} // namespace wenda
