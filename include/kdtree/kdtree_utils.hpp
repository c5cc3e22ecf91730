// Fixture generators of the drop-in C++ API: same names, arguments and VALUES as the reference's
// kdtree_utils.hpp:16-90 (make_random_position_and_index[_array]) and as the CLI's
// fill_random_positions (main.cpp:14-35), so that the reference's seeded tests and benchmarks can be
// replayed against this library bit for bit.
//
// The reference draws from Random123's Philox4x32-10 (third_party/random123, not vendored here).
// This is a restatement of the published algorithm (Salmon, Moraes, Dror, Shaw: "Parallel random
// numbers: as easy as 1, 2, 3", SC'11): 10 rounds of two 32x32->64 multiplies with the constants
// below and a Weyl key schedule; u01<float>(x) = x * 2^-32 + 2^-33 (Random123/uniform.hpp:174-184).
// tests/test_oracle.py pins it against values produced by the compiled reference.
#pragma once

#include <array>
#include <cstdint>
#include <limits>
#include <numeric>
#include <vector>

#include "kdtree.hpp"

namespace wenda {
namespace kdtree {

namespace philox {
inline std::array<uint32_t, 4> philox4x32_10(std::array<uint32_t, 4> ctr, std::array<uint32_t, 2> key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)M0 * ctr[0], p1 = (uint64_t)M1 * ctr[2];
        ctr = {(uint32_t)(p1 >> 32) ^ ctr[1] ^ key[0], (uint32_t)p1, (uint32_t)(p0 >> 32) ^ ctr[3] ^ key[1],
               (uint32_t)p0};
        key[0] += W0;
        key[1] += W1;
    }
    return ctr;
}
inline float u01(uint32_t x) { return (float)x * (1.0f / 4294967296.0f) + (0.5f / 4294967296.0f); }
} // namespace philox

//! kdtree_utils.hpp:16-46: counter {dim, i, 0, 0}, key {seed, 0}, lane 0.
template <size_t R = 3>
inline std::vector<PositionAndIndex<R>> make_random_position_and_index(uint32_t n, unsigned int seed,
                                                                       float boxsize = 1.0) {
    std::vector<PositionAndIndex<R>> positions(n);
    for (uint32_t i = 0; i < n; ++i) {
        for (size_t dim = 0; dim < R; ++dim)
            positions[i].position[dim] =
                philox::u01(philox::philox4x32_10({(uint32_t)dim, i, 0u, 0u}, {seed, 0u})[0]) * boxsize;
        positions[i].index = i;
    }
    return positions;
}

//! kdtree_utils.hpp:52-90: the same values as SoA columns, padded with max() to a multiple of block_size.
template <size_t R = 3, typename T = float, typename IndexT = uint32_t>
inline PositionAndIndexArray<R, T, IndexT> make_random_position_and_index_array(uint32_t n, unsigned int seed,
                                                                                float boxsize = 1.0,
                                                                                int block_size = -1) {
    const uint32_t size_up = block_size <= 0 ? n : (n + block_size - 1) / block_size * block_size;
    PositionAndIndexArray<R, T, IndexT> result(size_up);
    for (size_t dim = 0; dim < R; ++dim) {
        for (uint32_t i = 0; i < n; ++i)
            result.positions_[dim][i] =
                philox::u01(philox::philox4x32_10({(uint32_t)dim, i, 0u, 0u}, {seed, 0u})[0]) * boxsize;
        for (uint32_t i = n; i < size_up; ++i) result.positions_[dim][i] = std::numeric_limits<T>::max();
    }
    std::iota(result.indices_.begin(), result.indices_.end(), 0);
    return result;
}

//! main.cpp:14-35 (the CLI's points): counter {i, 0, 0, 0}, key {seed, 0}, lanes 0..2.
inline std::vector<std::array<float, 3>> fill_random_positions(uint32_t n, unsigned int seed) {
    std::vector<std::array<float, 3>> positions(n);
    for (uint32_t i = 0; i < n; ++i) {
        const auto r = philox::philox4x32_10({i, 0u, 0u, 0u}, {seed, 0u});
        positions[i] = {philox::u01(r[0]), philox::u01(r[1]), philox::u01(r[2])};
    }
    return positions;
}

//! kdtree.cpp:64-90 (declared kdtree_utils.hpp:117-118): AoS positions -> SoA columns with indices
//! 0..n-1; if block_size > 0 the array is padded to a multiple of it with max() positions (the padding
//! points carry the indices n.. like every other row).  Defined inline here: the reference compiles it
//! into its library, the drop-in is header-only above the C ABI.
inline PositionAndIndexArray<3, float, uint32_t>
make_position_and_indices(tcb::span<const std::array<float, 3>> const &positions, int block_size = -1) {
    const size_t n = positions.size();
    const size_t block = block_size > 0 ? static_cast<size_t>(block_size) : 1;
    const size_t padded = (n + block - 1) / block * block;
    PositionAndIndexArray<3, float, uint32_t> out(padded);
    for (size_t dim = 0; dim < 3; ++dim) {
        float *column = out.positions_[dim];
        for (size_t i = 0; i < n; ++i) column[i] = positions[i][dim];
        std::fill(column + n, column + padded, std::numeric_limits<float>::max());
    }
    std::iota(out.indices_.begin(), out.indices_.end(), 0u);
    return out;
}

} // namespace kdtree
} // namespace wenda
