// backend_cuda.cuh -- device memory, copies, sort and kernel launch for the product build (nvcc, sm_100a).
// tests/emu/backend_emu.h provides the same interface on the host for kernel-logic unit tests.
#pragma once
#include <sys/mman.h>
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include "vor_common.cuh"

namespace vor {
namespace be {

typedef cudaStream_t Stream;

struct CudaError : std::runtime_error {
    int code;
    CudaError(const std::string &m, int c) : std::runtime_error(m), code(c) {}
};

inline void check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        const int code = (e == cudaErrorMemoryAllocation) ? ERR_OOM : ERR_CUDA;
        throw CudaError(std::string(what) + ": " + cudaGetErrorString(e), code);
    }
}
#define VOR_CUDA(x) ::vor::be::check((x), #x)

inline void set_device(int dev) { VOR_CUDA(cudaSetDevice(dev)); }
// Caching device allocator (per process, per device): a triangulation allocates ~1.2 KB per 3D point in a handful of
// large arrays; cudaMalloc / cudaFree of multi-GB blocks cost tens of milliseconds, so freed blocks are kept and
// handed out again to the next tree (vor_release_memory() returns them to the driver).
struct Pool {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void *> cache;   // (device, size) -> block
    std::unordered_map<void *, std::pair<int, size_t>> live;
    size_t cached = 0;
    size_t limit = (size_t)96 << 30;
    static size_t round_up(size_t b) {
        const size_t g = b >= ((size_t)1 << 20) ? ((size_t)2 << 20) : 512;
        return (b + g - 1) / g * g;
    }
    void *alloc(size_t bytes) {
        const size_t need = round_up(bytes ? bytes : 16);
        int dev = 0;
        VOR_CUDA(cudaGetDevice(&dev));
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = cache.lower_bound({dev, need});
            if (it != cache.end() && it->first.first == dev && it->first.second <= need + need / 4 + ((size_t)2 << 20)) {
                void *p = it->second;
                live[p] = it->first;
                cached -= it->first.second;
                cache.erase(it);
                return p;
            }
        }
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, need);
        if (e == cudaErrorMemoryAllocation) {   // give the cache back and retry once
            cudaGetLastError();
            release();
            e = cudaMalloc(&p, need);
        }
        check(e, "cudaMalloc");
        std::lock_guard<std::mutex> lk(mu);
        live[p] = {dev, need};
        return p;
    }
    void free(void *p) {
        if (!p) return;
        cudaDeviceSynchronize();   // same guarantee as cudaFree: no kernel still uses the block
        std::lock_guard<std::mutex> lk(mu);
        auto it = live.find(p);
        if (it == live.end()) { cudaFree(p); return; }
        const auto key = it->second;
        live.erase(it);
        if (cached + key.second > limit) { cudaFree(p); return; }
        cache.emplace(key, p);
        cached += key.second;
    }
    void release() {
        std::lock_guard<std::mutex> lk(mu);
        for (auto &kv : cache) cudaFree(kv.second);
        cache.clear();
        cached = 0;
    }
};
extern Pool g_pool;
// Caching allocator of host blocks for results that are handed to the caller (vor_tree_edges_host).  A recycled block
// is already faulted in, so the copy into it does not pay the first-touch page faults that dominate a copy into a
// fresh 620 MB buffer (10M-point edge list: ~50 ms -> ~20 ms).  Blocks are pageable by default (the copy is staged
// through the two pinned chunks below); with VOR_PINNED_RESULTS=1 they are page-locked and the DMA writes them directly
// (~12 ms), at the price of ~270 ms of cudaMallocHost the first time a size is seen.  vor_release_memory() frees them.
struct HostPool {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void *> cache;          // (pinned, size) -> block
    std::unordered_map<void *, std::pair<int, size_t>> live;
    // page-locked result blocks by default (the DMA writes them directly: 621 MB of edges in ~12 ms instead of ~30 ms through
    // the staging chunks); VOR_PINNED_RESULTS=0 for pageable blocks.  The first block of a size costs a cudaMallocHost
    // (~270 ms for 621 MB): vor_delaunay warms the pool on a side thread while the points are inserted.
    static bool want_pinned() { static const bool p = [] { const char *e = getenv("VOR_PINNED_RESULTS"); return !e || atoi(e) != 0; }(); return p; }
    bool has_block(size_t bytes) {
        const size_t g = (size_t)2 << 20;
        const size_t need = (std::max(bytes, (size_t)16) + g - 1) / g * g;
        const int pin = want_pinned() ? 1 : 0;
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.lower_bound({pin, need});
        return it != cache.end() && it->first.first == pin && it->first.second <= need + need / 4 + g;
    }
    void *alloc(size_t bytes, bool *pinned_out) {
        const size_t g = (size_t)2 << 20;
        const size_t need = (std::max(bytes, (size_t)16) + g - 1) / g * g;
        const int pin = want_pinned() ? 1 : 0;
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = cache.lower_bound({pin, need});
            if (it != cache.end() && it->first.first == pin && it->first.second <= need + need / 4 + g) {
                void *p = it->second;
                live[p] = it->first;
                cache.erase(it);
                *pinned_out = pin != 0;
                return p;
            }
        }
        void *p = nullptr;
        if (pin) VOR_CUDA(cudaMallocHost(&p, need));
        else {
            if (posix_memalign(&p, g, need) != 0) throw CudaError("out of host memory", 5);
            madvise(p, need, MADV_HUGEPAGE);   // first touch faults 2 MB at a time (what numpy does for large arrays)
        }
        std::lock_guard<std::mutex> lk(mu);
        live[p] = {pin, need};
        *pinned_out = pin != 0;
        return p;
    }
    bool free(void *p) {
        if (!p) return true;
        std::lock_guard<std::mutex> lk(mu);
        auto it = live.find(p);
        if (it == live.end()) return false;
        cache.emplace(it->second, p);
        live.erase(it);
        return true;
    }
    void release() {
        std::lock_guard<std::mutex> lk(mu);
        for (auto &kv : cache) { if (kv.first.first) cudaFreeHost(kv.second); else ::free(kv.second); }
        cache.clear();
    }
};
extern HostPool g_hostpool;
inline void *dmalloc(size_t bytes) { return g_pool.alloc(bytes); }
inline void dfree(void *p) { g_pool.free(p); }
inline void release_cached() { g_pool.release(); g_hostpool.release(); }
inline void dmemset(void *p, int byte, size_t n, Stream s) { VOR_CUDA(cudaMemsetAsync(p, byte, n, s)); }
inline void h2d(void *d, const void *h, size_t n, Stream s) { VOR_CUDA(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); }
inline void d2h(void *h, const void *d, size_t n, Stream s) { VOR_CUDA(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); }
inline void d2d(void *d, const void *s_, size_t n, Stream s) { VOR_CUDA(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s)); }
inline void sync(Stream s) { VOR_CUDA(cudaStreamSynchronize(s)); }

// Large copies between pageable host memory and the device go through two pinned staging buffers so that the DMA of
// chunk i+1 overlaps the host memcpy of chunk i (a plain cudaMemcpy on pageable memory runs at a few GB/s).
// One staging set per DEVICE: the events are created under the device that is current at the first large copy, and an
// event may only be recorded on a stream of its own device; per-device mutexes also keep the copies of different
// devices (vor_delaunay_batch: one host thread per device) from serialising each other.
struct Staging {
    static constexpr size_t CHUNK = (size_t)32 << 20;
    char *buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2];
    std::mutex mu;
    void init() {
        if (buf[0]) return;
        for (int i = 0; i < 2; i++) {
            VOR_CUDA(cudaMallocHost((void **)&buf[i], CHUNK));
            VOR_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
    }
};
constexpr int MAX_DEVICES = 64;
extern Staging g_staging_dev[MAX_DEVICES];
inline Staging &staging_here() {
    int dev = 0;
    VOR_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) throw CudaError("device index beyond the staging table", ERR_ARG);
    return g_staging_dev[dev];
}
// host memcpy on a few threads: the destination of a result copy is usually freshly allocated memory, so the copy is
// dominated by first-touch page faults, which scale with threads
inline int copy_threads() {
    static const int T = [] {
        int t = (int)std::thread::hardware_concurrency() / 2;
        if (const char *e = getenv("VOR_COPY_THREADS")) t = atoi(e);
        return std::max(1, std::min(t, 16));
    }();
    return T;
}
inline void par_memcpy(char *dst, const char *src, size_t n) {
    const int T = copy_threads();
    if (T == 1 || n < ((size_t)4 << 20)) { memcpy(dst, src, n); return; }
    std::thread th[16];
    const size_t part = (n / T + 4095) & ~(size_t)4095;
    for (int i = 1; i < T; i++) {
        const size_t off = std::min(n, part * i), len = std::min(n - off, part);
        th[i - 1] = std::thread([=] { if (len) memcpy(dst + off, src + off, len); });
    }
    memcpy(dst, src, std::min(n, part));
    for (int i = 0; i < T - 1; i++) th[i].join();
}
inline void d2h_big(void *h, const void *d, size_t n, Stream s) {
    if (n < 4 * Staging::CHUNK) { VOR_CUDA(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); VOR_CUDA(cudaStreamSynchronize(s)); return; }
    Staging &g_staging = staging_here();
    std::lock_guard<std::mutex> lk(g_staging.mu);
    g_staging.init();
    const size_t nch = (n + Staging::CHUNK - 1) / Staging::CHUNK;
    auto len = [&](size_t c) { return std::min(Staging::CHUNK, n - c * Staging::CHUNK); };
    VOR_CUDA(cudaMemcpyAsync(g_staging.buf[0], d, len(0), cudaMemcpyDeviceToHost, s));
    VOR_CUDA(cudaEventRecord(g_staging.ev[0], s));
    for (size_t c = 0; c < nch; c++) {
        const int cur = (int)(c & 1), nxt = cur ^ 1;
        if (c + 1 < nch) {
            VOR_CUDA(cudaMemcpyAsync(g_staging.buf[nxt], (const char *)d + (c + 1) * Staging::CHUNK, len(c + 1), cudaMemcpyDeviceToHost, s));
            VOR_CUDA(cudaEventRecord(g_staging.ev[nxt], s));
        }
        VOR_CUDA(cudaEventSynchronize(g_staging.ev[cur]));
        par_memcpy((char *)h + c * Staging::CHUNK, g_staging.buf[cur], len(c));
    }
}
inline void h2d_big(void *d, const void *h, size_t n, Stream s) {
    if (n < 4 * Staging::CHUNK) { VOR_CUDA(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); VOR_CUDA(cudaStreamSynchronize(s)); return; }
    Staging &g_staging = staging_here();
    std::lock_guard<std::mutex> lk(g_staging.mu);
    g_staging.init();
    const size_t nch = (n + Staging::CHUNK - 1) / Staging::CHUNK;
    auto len = [&](size_t c) { return std::min(Staging::CHUNK, n - c * Staging::CHUNK); };
    for (size_t c = 0; c < nch; c++) {
        const int cur = (int)(c & 1);
        if (c >= 2) VOR_CUDA(cudaEventSynchronize(g_staging.ev[cur]));   // the DMA that used this buffer is done
        par_memcpy(g_staging.buf[cur], (const char *)h + c * Staging::CHUNK, len(c));
        VOR_CUDA(cudaMemcpyAsync((char *)d + c * Staging::CHUNK, g_staging.buf[cur], len(c), cudaMemcpyHostToDevice, s));
        VOR_CUDA(cudaEventRecord(g_staging.ev[cur], s));
    }
    VOR_CUDA(cudaStreamSynchronize(s));
}
inline void *hmalloc_pinned(size_t bytes) {
    void *p = nullptr;
    VOR_CUDA(cudaMallocHost(&p, bytes ? bytes : 16));
    return p;
}
inline void hfree_pinned(void *p) { if (p) cudaFreeHost(p); }

// library plumbing (not a hot-path kernel): CUB LSD radix sort of (u64 key, u32 value) pairs
inline void sort_pairs(uint64_t *keys_in, uint64_t *keys_out, uint32_t *vals_in, uint32_t *vals_out, size_t n, Stream s) {
    size_t tmp = 0;
    VOR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, keys_in, keys_out, vals_in, vals_out, (int)n, 0, 64, s));
    void *d = dmalloc(tmp);
    VOR_CUDA(cub::DeviceRadixSort::SortPairs(d, tmp, keys_in, keys_out, vals_in, vals_out, (int)n, 0, 64, s));
    VOR_CUDA(cudaStreamSynchronize(s));
    dfree(d);
}
inline void sort_keys(uint64_t *keys_in, uint64_t *keys_out, size_t n, Stream s) {
    size_t tmp = 0;
    VOR_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp, keys_in, keys_out, (int)n, 0, 64, s));
    void *d = dmalloc(tmp);
    VOR_CUDA(cub::DeviceRadixSort::SortKeys(d, tmp, keys_in, keys_out, (int)n, 0, 64, s));
    VOR_CUDA(cudaStreamSynchronize(s));
    dfree(d);
}

extern std::atomic<unsigned long long> g_launches;   // kernels of this library launched so far (bench.py gpu_launches)
// a launch-configuration failure must surface at the launch, not as "no progress" hundreds of rounds later
inline void check_launch(const char *what) { check(cudaGetLastError(), what); }

// launch with the programmatic-stream-serialization attribute (see pdl_wait / pdl_trigger in coop_kernels.cuh); pdl == false:
// a plain launch (the kernel's griddepcontrol instructions are no-ops then)
template <class... KArgs, class... Args>
inline void launch_pdl(bool pdl, void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, Stream s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    VOR_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

// per-kernel-class CUDA-event timing on the launching stream (bench.py roofline; off unless option "profile")
struct Prof {
    bool on = false;
    std::vector<cudaEvent_t> ev;   // pairs
    std::vector<int> cls;
    double ms[4] = {0, 0, 0, 0};
    double cnt[4] = {0, 0, 0, 0};
    void start(int c, Stream s) {
        if (!on) return;
        cudaEvent_t a, b;
        VOR_CUDA(cudaEventCreate(&a));
        VOR_CUDA(cudaEventCreate(&b));
        ev.push_back(a); ev.push_back(b); cls.push_back(c);
        VOR_CUDA(cudaEventRecord(a, s));
    }
    void stop(Stream s) {
        if (!on) return;
        VOR_CUDA(cudaEventRecord(ev.back(), s));
    }
    void resolve(Stream s) {
        if (!on) return;
        VOR_CUDA(cudaStreamSynchronize(s));
        for (size_t i = 0; i < cls.size(); i++) {
            float t = 0;
            VOR_CUDA(cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]));
            ms[cls[i]] += t;
            cnt[cls[i]] += 1;
            cudaEventDestroy(ev[2 * i]);
            cudaEventDestroy(ev[2 * i + 1]);
        }
        ev.clear();
        cls.clear();
    }
};

} // namespace be

// generic one-thread-per-item kernel around a body function
template <class Args, void (*Body)(const Args &, int)>
__global__ void __launch_bounds__(256) k_items(Args a, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) Body(a, i);
}
// variant without early exit (bodies that use full-warp collectives)
template <class Args, void (*Body)(const Args &, int, bool)>
__global__ void __launch_bounds__(256) k_items_full(Args a, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    Body(a, i, i < n);
}

#define VOR_LAUNCH(ArgsT, body, n, args, stream)                                                         \
    do {                                                                                                 \
        const int _n = (int)(n);                                                                         \
        if (_n > 0) {                                                                                    \
            ::vor::k_items<ArgsT, body><<<(_n + 255) / 256, 256, 0, stream>>>(args, _n);                 \
            ::vor::be::check_launch(#body);                                                              \
            ::vor::be::g_launches++;                                                                     \
        }                                                                                                \
    } while (0)
#define VOR_LAUNCH_FULL(ArgsT, body, n, args, stream)                                                    \
    do {                                                                                                 \
        const int _n = (int)(n);                                                                         \
        if (_n > 0) {                                                                                    \
            ::vor::k_items_full<ArgsT, body><<<(_n + 255) / 256, 256, 0, stream>>>(args, _n);            \
            ::vor::be::check_launch(#body);                                                              \
            ::vor::be::g_launches++;                                                                     \
        }                                                                                                \
    } while (0)

} // namespace vor
