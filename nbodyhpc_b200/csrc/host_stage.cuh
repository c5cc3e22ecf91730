// Staged download of results into PAGEABLE host memory.
//
// The reference hands freshly allocated arrays to numpy (pybind.cpp:103-104,174-188); so does the
// drop-in, and such memory is pageable and untouched.  A plain cudaMemcpy into it runs at ~20 GB/s and
// takes one page fault per 4 KB on a single driver thread: 1.6 s for the 6.4 GB of a 10^8 x 8 result,
// twenty times the kernel.  Here the device-to-host copy goes at full PCIe speed into a small ring of
// pinned slots, and a few host threads move finished slots into the destination (first touch and
// copy in parallel) while the GPU keeps computing the next slice.
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace nbk {

// {staged downloads, direct pageable downloads, staged uploads, direct pageable uploads} (nbk_host_path_stats)
inline std::atomic<uint64_t> g_host_path[4];

class PinnedRing { // one per device, allocated on first use; a staged call holds it for its duration
  public:
#ifndef NBK_RING_SLOT_MB
#define NBK_RING_SLOT_MB 32 // measured: 16 MB slots with 16 threads were slower (0.235 s vs 0.185 s per 10^8 x 8 rows)
#endif
    // 384 MB of pinned memory per device.  A slot is the unit one host thread copies, so the number of
    // download slots bounds how many threads can work on results at once.
    static constexpr size_t kSlotBytes = (size_t)NBK_RING_SLOT_MB << 20;
    static constexpr int kSlots = (int)(((size_t)384 << 20) / kSlotBytes); // [0, kUploadSlots): uploads, the rest: downloads
    static constexpr int kUploadSlots = kSlots / 6;
    static constexpr int kMaxDevices = 64;
    // Blocks while another call on the same device is staging (the transfers of both would share the
    // same PCIe link anyway); calls on different devices -- KDTree(devices=[...]) runs one host thread
    // per replica -- have a ring each.  nullptr only if pinned memory cannot be had.
    static PinnedRing *acquire(int device) {
        static PinnedRing rings[kMaxDevices];
        if (device < 0 || device >= kMaxDevices) return nullptr;
        PinnedRing &ring = rings[device];
        ring.lock_.lock();
        if (!ring.base_) {
            void *p = nullptr;
            if (cudaHostAlloc(&p, kSlots * kSlotBytes, cudaHostAllocPortable) != cudaSuccess) {
                cudaGetLastError();
                ring.lock_.unlock();
                return nullptr;
            }
            ring.base_ = static_cast<char *>(p);
        }
        return &ring;
    }
    void release() { lock_.unlock(); }
    char *slot(int i) const { return base_ + (size_t)i * kSlotBytes; }

  private:
    PinnedRing() = default;
    std::mutex lock_;
    char *base_ = nullptr;
};

class StagedDownload {
  public:
    StagedDownload(PinnedRing *ring, int device, int threads) : ring_(ring) {
        for (int i = 0; i < PinnedRing::kSlots; ++i) {
            NBK_CUDA(cudaEventCreateWithFlags(&events_[i], cudaEventDisableTiming));
            free_[i] = i >= PinnedRing::kUploadSlots; // the first slots belong to StagedUpload
        }
        for (int t = 0; t < threads; ++t) workers_.emplace_back([this, device] { work(device); });
    }
    ~StagedDownload() {
        finish();
        for (auto &e : events_) cudaEventDestroy(e);
        ring_->release();
    }
    // `bytes` from device memory (ready in `stream` order) to pageable `dst`; blocks only while the ring is full
    void download(void *dst, const void *d_src, size_t bytes, cudaStream_t stream) {
        for (size_t off = 0; off < bytes; off += PinnedRing::kSlotBytes) {
            const size_t n = std::min(PinnedRing::kSlotBytes, bytes - off);
            int slot = -1;
            {
                std::unique_lock<std::mutex> lock(mutex_);
                cv_free_.wait(lock, [&] {
                    for (int i = 0; i < PinnedRing::kSlots; ++i)
                        if (free_[i]) {
                            slot = i;
                            return true;
                        }
                    return false;
                });
                free_[slot] = false;
            }
            NBK_CUDA(cudaMemcpyAsync(ring_->slot(slot), static_cast<const char *>(d_src) + off, n,
                                     cudaMemcpyDeviceToHost, stream));
            NBK_CUDA(cudaEventRecord(events_[slot], stream));
            {
                std::lock_guard<std::mutex> lock(mutex_);
                tasks_.push_back(Task{slot, static_cast<char *>(dst) + off, n});
            }
            cv_task_.notify_one();
        }
    }
    void finish() {
        {
            std::lock_guard<std::mutex> lock(mutex_);
            if (done_) return;
            done_ = true;
        }
        cv_task_.notify_all();
        for (auto &w : workers_) w.join();
        workers_.clear();
    }

  private:
    struct Task {
        int slot;
        char *dst;
        size_t bytes;
    };
    void work(int device) {
        cudaSetDevice(device);
        while (true) {
            Task t;
            {
                std::unique_lock<std::mutex> lock(mutex_);
                cv_task_.wait(lock, [&] { return done_ || !tasks_.empty(); });
                if (tasks_.empty()) return; // done_ and drained
                t = tasks_.front();
                tasks_.pop_front();
            }
            cudaEventSynchronize(events_[t.slot]);
            std::memcpy(t.dst, ring_->slot(t.slot), t.bytes);
            {
                std::lock_guard<std::mutex> lock(mutex_);
                free_[t.slot] = true;
            }
            cv_free_.notify_one();
        }
    }
    PinnedRing *ring_;
    cudaEvent_t events_[PinnedRing::kSlots];
    bool free_[PinnedRing::kSlots];
    std::vector<std::thread> workers_;
    std::mutex mutex_;
    std::condition_variable cv_task_, cv_free_;
    std::deque<Task> tasks_;
    bool done_ = false;
};

// The other direction: pageable host memory -> device.  Host threads copy chunks into pinned slots
// and issue the H2D copy of each; upload() returns when the data is on the device.
class StagedUpload {
  public:
    StagedUpload(PinnedRing *ring, int device, int threads) : ring_(ring) {
        for (int i = 0; i < PinnedRing::kUploadSlots; ++i) free_[i] = true;
        for (int t = 0; t < threads; ++t) workers_.emplace_back([this, device] { work(device); });
    }
    ~StagedUpload() {
        {
            std::lock_guard<std::mutex> lock(mutex_);
            done_ = true;
        }
        cv_task_.notify_all();
        for (auto &w : workers_) w.join();
    }
    void upload(void *d_dst, const void *src, size_t bytes, cudaStream_t stream) {
        for (size_t off = 0; off < bytes; off += PinnedRing::kSlotBytes) {
            const size_t n = std::min(PinnedRing::kSlotBytes, bytes - off);
            std::unique_lock<std::mutex> lock(mutex_);
            int slot = -1;
            cv_free_.wait(lock, [&] {
                for (int i = 0; i < PinnedRing::kUploadSlots; ++i)
                    if (free_[i]) {
                        slot = i;
                        return true;
                    }
                return false;
            });
            free_[slot] = false;
            ++pending_;
            tasks_.push_back(Task{slot, static_cast<char *>(d_dst) + off, static_cast<const char *>(src) + off, n, stream});
            lock.unlock();
            cv_task_.notify_one();
        }
        std::unique_lock<std::mutex> lock(mutex_);
        cv_free_.wait(lock, [&] { return pending_ == 0; });
        if (failed_) throw Error(NBK_ERR_CUDA, "staged upload failed");
    }

  private:
    struct Task {
        int slot;
        char *d_dst;
        const char *src;
        size_t bytes;
        cudaStream_t stream;
    };
    void work(int device) {
        cudaSetDevice(device);
        cudaEvent_t ev;
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        while (true) {
            Task t;
            {
                std::unique_lock<std::mutex> lock(mutex_);
                cv_task_.wait(lock, [&] { return done_ || !tasks_.empty(); });
                if (tasks_.empty()) break;
                t = tasks_.front();
                tasks_.pop_front();
            }
            std::memcpy(ring_->slot(t.slot), t.src, t.bytes);
            bool ok = cudaMemcpyAsync(t.d_dst, ring_->slot(t.slot), t.bytes, cudaMemcpyHostToDevice, t.stream) == cudaSuccess;
            ok = ok && cudaEventRecord(ev, t.stream) == cudaSuccess && cudaEventSynchronize(ev) == cudaSuccess;
            {
                std::lock_guard<std::mutex> lock(mutex_);
                free_[t.slot] = true;
                --pending_;
                failed_ = failed_ || !ok;
            }
            cv_free_.notify_all();
        }
        cudaEventDestroy(ev);
    }
    PinnedRing *ring_;
    bool free_[PinnedRing::kUploadSlots];
    std::vector<std::thread> workers_;
    std::mutex mutex_;
    std::condition_variable cv_task_, cv_free_;
    std::deque<Task> tasks_;
    size_t pending_ = 0;
    bool done_ = false, failed_ = false;
};

inline bool is_pageable_host(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

} // namespace nbk
