// Shared host/device helpers for libnbk (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "nbk.h"

namespace nbk {

struct Error : std::runtime_error {
    int code;
    Error(int c, std::string const &msg) : std::runtime_error(msg), code(c) {}
};

inline void cuda_check(cudaError_t e, const char *what, const char *file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error at %s:%d (%s): %s", file, line, what,
                 cudaGetErrorString(e));
        throw Error(e == cudaErrorMemoryAllocation ? NBK_ERR_NOMEM : NBK_ERR_CUDA, buf);
    }
}
#define NBK_CUDA(expr) ::nbk::cuda_check((expr), #expr, __FILE__, __LINE__)

extern std::atomic<uint64_t> g_launches;
// Counts the launch (nbk_launch_count) and surfaces launch-configuration errors immediately.
#define NBK_LAUNCHED()                                                                             \
    do {                                                                                           \
        ::nbk::g_launches.fetch_add(1, std::memory_order_relaxed);                                 \
        NBK_CUDA(cudaGetLastError());                                                              \
    } while (0)

inline uint64_t div_up(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
inline uint64_t align_up(uint64_t a, uint64_t b) { return div_up(a, b) * b; }

// Stream-ordered scratch allocation (cudaMallocAsync pool: after the first call of a given size the
// memory comes back from the pool without a device synchronisation).
struct Scratch {
    cudaStream_t stream;
    void *ptrs[32];
    int count = 0;
    explicit Scratch(cudaStream_t s) : stream(s) {}
    Scratch(Scratch const &) = delete;
    template <typename T> T *get(uint64_t n) {
        void *p = nullptr;
        uint64_t bytes = n * sizeof(T);
        if (bytes == 0) bytes = sizeof(T);
        NBK_CUDA(cudaMallocAsync(&p, bytes, stream));
        if (count >= 32) throw Error(NBK_ERR_INVALID, "scratch table overflow");
        ptrs[count++] = p;
        return static_cast<T *>(p);
    }
    ~Scratch() {
        for (int i = count - 1; i >= 0; --i) cudaFreeAsync(ptrs[i], stream);
    }
};

// ---- float <-> order-preserving uint32 --------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t float_to_ordered(uint32_t bits) {
    return bits ^ ((bits & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ __forceinline__ uint32_t ordered_to_float(uint32_t o) {
    return o ^ ((o & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu);
}

constexpr uint32_t kFltMaxBits = 0x7F7FFFFFu;

} // namespace nbk
