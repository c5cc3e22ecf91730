// Shared host/device helpers for libnbk (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <vector>
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

// Library-owned cache of big scratch blocks (per device).  A tree build needs several GB of scratch
// for a few milliseconds; taking it from the driver (or from the stream-ordered pool, which re-maps
// physical memory when block sizes interleave) costs more than the build itself, so the block of the
// last build is kept and handed to the next one.  trim() returns it to the device.
struct BlockCache {
    struct Entry {
        void *ptr;
        uint64_t bytes;
        int device;
        bool busy;
        cudaEvent_t ready; // work still reading the block when it was released (a freed tree's queries), or null
    };
    static std::mutex &mutex() {
        static std::mutex m;
        return m;
    }
    static std::vector<Entry> &entries() {
        static std::vector<Entry> e;
        return e;
    }
    // `stream`: where the caller will first use the block (made to wait for the block's previous readers)
    static void *acquire(uint64_t bytes, cudaStream_t stream) {
        int dev = 0;
        NBK_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(mutex());
        auto &es = entries();
        // best fit among this device's idle blocks that are not more than twice too big
        Entry *best = nullptr;
        for (auto &e : es)
            if (!e.busy && e.device == dev && e.bytes >= bytes && e.bytes / 2 <= bytes && (!best || e.bytes < best->bytes))
                best = &e;
        if (best) {
            best->busy = true;
            if (best->ready) {
                cudaStreamWaitEvent(stream, best->ready, 0);
                cudaEventDestroy(best->ready);
                best->ready = nullptr;
            }
            return best->ptr;
        }
        // miss: if the cache is full of wrong-sized idle blocks, drop them, then take a fresh one
        size_t idle = 0;
        for (auto &e : es) idle += (!e.busy && e.device == dev) ? 1 : 0;
        if (idle >= 4) {
            for (size_t i = 0; i < es.size();) {
                if (!es[i].busy && es[i].device == dev) {
                    drop(es[i]);
                    es.erase(es.begin() + i);
                } else {
                    ++i;
                }
            }
        }
        void *p = nullptr;
        cudaError_t err = cudaMalloc(&p, bytes);
        if (err != cudaSuccess) {
            // out of memory with idle blocks cached: give them back and retry once
            cudaGetLastError();
            for (size_t i = 0; i < es.size();) {
                if (!es[i].busy && es[i].device == dev) {
                    drop(es[i]);
                    es.erase(es.begin() + i);
                } else {
                    ++i;
                }
            }
            NBK_CUDA(cudaMalloc(&p, bytes));
        }
        es.push_back(Entry{p, bytes, dev, true, nullptr});
        return p;
    }
    // `ready` (optional, ownership passes to the cache): an event after which the block is no longer read
    static void release(void *ptr, cudaEvent_t ready = nullptr) {
        std::lock_guard<std::mutex> lock(mutex());
        for (auto &e : entries())
            if (e.ptr == ptr) {
                e.busy = false;
                e.ready = ready;
                return;
            }
        if (ready) cudaEventDestroy(ready);
    }
    static void drop(Entry &e) { // cudaFree waits for the device, so a pending `ready` is settled by it
        if (e.ready) cudaEventDestroy(e.ready);
        e.ready = nullptr;
        cudaFree(e.ptr);
    }
    static void trim(int device) {
        std::lock_guard<std::mutex> lock(mutex());
        auto &es = entries();
        for (size_t i = 0; i < es.size();) {
            if (!es[i].busy && es[i].device == device) {
                drop(es[i]);
                es.erase(es.begin() + i);
            } else {
                ++i;
            }
        }
    }
};

// Scratch for one call.  get() takes stream-ordered allocations from the cudaMallocAsync pool;
// reserve(bytes) first makes the following get() calls carve one block from the BlockCache instead
// (the caller must have synchronised `stream` with the last use before the Scratch dies: the
// destructor synchronises it to be safe).
struct Scratch {
    cudaStream_t stream;
    void *ptrs[32];
    int count = 0;
    char *block = nullptr;
    uint64_t block_bytes = 0, block_used = 0, total_bytes = 0;
    explicit Scratch(cudaStream_t s) : stream(s) {}
    Scratch(Scratch const &) = delete;
    static uint64_t padded(uint64_t bytes) { return (bytes + 255) / 256 * 256; }
    static constexpr uint64_t kCacheFrom = 64ull << 20; // smaller blocks: the pool is quick and exact
    bool block_cached = false;
    void reserve(uint64_t bytes) {
        if (block || bytes == 0) return;
        if (bytes >= kCacheFrom) {
            block = static_cast<char *>(BlockCache::acquire(bytes, stream));
            block_cached = true;
        } else {
            void *p = nullptr;
            NBK_CUDA(cudaMallocAsync(&p, bytes, stream));
            ptrs[count++] = p;
            block = static_cast<char *>(p);
        }
        block_bytes = bytes;
        total_bytes += bytes;
    }
    template <typename T> T *get(uint64_t n) {
        uint64_t bytes = n * sizeof(T);
        if (bytes == 0) bytes = sizeof(T);
        if (block && block_used + padded(bytes) <= block_bytes) {
            T *p = reinterpret_cast<T *>(block + block_used);
            block_used += padded(bytes);
            return p;
        }
        void *p = nullptr;
        NBK_CUDA(cudaMallocAsync(&p, bytes, stream));
        if (count >= 32) throw Error(NBK_ERR_INVALID, "scratch table overflow");
        ptrs[count++] = p;
        total_bytes += bytes;
        return static_cast<T *>(p);
    }
    ~Scratch() {
        for (int i = count - 1; i >= 0; --i) cudaFreeAsync(ptrs[i], stream);
        if (block && block_cached) {
            cudaStreamSynchronize(stream);
            BlockCache::release(block);
        }
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
