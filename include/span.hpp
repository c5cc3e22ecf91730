// Forwarding header for code that includes <span.hpp> the way the reference's sources do
// (kdtree/src/cpp/include/kdtree/kdtree.hpp:11, kdtree.cpp:23: a vendored C++11 span in namespace tcb,
// third_party/misc/span.hpp).  The drop-in headers need C++20, which has the real thing.
#pragma once

#include <span>

namespace tcb {
using std::span;
inline constexpr std::size_t dynamic_extent = std::dynamic_extent;
} // namespace tcb
