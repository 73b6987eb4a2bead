// pinned_pool.h -- process-wide pinned staging pool shared by the translation units of libb200ba.so.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace b200ba {

// Process-wide pinned staging pool.  The large host->device / device->host copies of a BA call (point-order measurement
// and camera-index arrays, the point block) go through page-locked memory: SequenceSFM calls the BA once per keyframe from
// one worker thread, so the pool is allocated once (cudaHostAlloc of ~100 MB costs ~50 ms) and reused by every later call,
// where it turns 12 GB/s pageable copies into ~50 GB/s ones.  Lease = exclusive use until release (a mutex: handles on
// other threads simply wait); when pinning fails the lease falls back to ordinary memory.
struct PinnedPool {
    std::mutex mu; char* base = nullptr; size_t cap = 0;
    struct Lease {
        PinnedPool* pool = nullptr; char* p = nullptr; std::unique_ptr<char[]> fallback;
        ~Lease() { if (pool) pool->mu.unlock(); }
    };
    void lease(Lease& L, size_t bytes)
    {
        mu.lock(); L.pool = this;
        if (bytes > cap && getenv("B200BA_NO_PINNED") == nullptr) {
            if (base) { cudaFreeHost(base); base = nullptr; cap = 0; }
            void* q = nullptr; const size_t want = bytes + bytes / 4;
            if (cudaHostAlloc(&q, want, cudaHostAllocDefault) == cudaSuccess) { base = (char*)q; cap = want; } else cudaGetLastError();
        }
        if (bytes <= cap) L.p = base; else { L.fallback.reset(new char[bytes]); L.p = L.fallback.get(); }
    }
};

// Moves a caller's pageable array into / out of the pool with a few host threads (one memcpy saturates ~10 GB/s, the
// PCIe DMA that follows or precedes it runs at ~50 GB/s).
inline void threaded_copy(void* dst, const void* src, size_t bytes)
{
    const unsigned hw = std::thread::hardware_concurrency();
    const int T = bytes >= ((size_t)4 << 20) ? (int)std::max(1u, std::min(8u, hw)) : 1;
    if (T == 1) { std::memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) {
        const size_t a = bytes * t / T, b = bytes * (t + 1) / T;
        th.emplace_back([=] { std::memcpy((char*)dst + a, (const char*)src + a, b - a); });
    }
    std::memcpy(dst, src, bytes / T);
    for (auto& x : th) x.join();
}

// the one pool of the library (defined in ba_engine.cu)
__attribute__((visibility("hidden"))) extern PinnedPool g_pinned;

}  // namespace b200ba
