// tests/emu/backend_emu.h -- TEST INFRASTRUCTURE.  Host stand-in for backend_cuda.cuh so that the kernel
// bodies and the host round loop of voronoids_b200/csrc can be unit-tested in the CPU-only container.
// Kernels run as sequential loops; "device memory" is the heap.  Never linked into libvoronoids_b200.so.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>
#include "vor_common.cuh"

namespace vor {
namespace be {
typedef int Stream;
struct CudaError : std::runtime_error {
    int code;
    CudaError(const std::string &m, int c) : std::runtime_error(m), code(c) {}
};
inline void set_device(int) {}
inline void *dmalloc(size_t bytes) { return std::malloc(bytes ? bytes : 16); }
inline void dfree(void *p) { std::free(p); }
inline void release_cached() {}
inline void dmemset(void *p, int byte, size_t n, Stream) { std::memset(p, byte, n); }
inline void h2d(void *d, const void *h, size_t n, Stream) { std::memcpy(d, h, n); }
inline void d2h(void *h, const void *d, size_t n, Stream) { std::memcpy(h, d, n); }
inline void d2d(void *d, const void *s, size_t n, Stream) { std::memcpy(d, s, n); }
inline void sync(Stream) {}
inline void d2h_big(void *h, const void *d, size_t n, Stream) { std::memcpy(h, d, n); }
inline void h2d_big(void *d, const void *h, size_t n, Stream) { std::memcpy(d, h, n); }
inline void *hmalloc_pinned(size_t bytes) { return std::malloc(bytes ? bytes : 16); }
inline void hfree_pinned(void *p) { std::free(p); }
struct HostPool {   // stand-in for the caching pinned allocator of backend_cuda.cuh
    std::vector<void *> live;
    void *alloc(size_t bytes, bool *pinned) { void *p = std::malloc(bytes ? bytes : 16); live.push_back(p); *pinned = false; return p; }
    bool free(void *p) {
        if (!p) return true;
        auto it = std::find(live.begin(), live.end(), p);
        if (it == live.end()) return false;
        live.erase(it);
        std::free(p);
        return true;
    }
    void release() {}
};
static HostPool g_hostpool;
inline void sort_pairs(uint64_t *ki, uint64_t *ko, uint32_t *vi, uint32_t *vo, size_t n, Stream) {
    std::vector<size_t> idx(n);
    std::iota(idx.begin(), idx.end(), (size_t)0);
    std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return ki[a] < ki[b]; });
    for (size_t i = 0; i < n; i++) { ko[i] = ki[idx[i]]; vo[i] = vi[idx[i]]; }
}
inline void sort_keys(uint64_t *ki, uint64_t *ko, size_t n, Stream) {
    std::copy(ki, ki + n, ko);
    std::sort(ko, ko + n);
}
static unsigned long long g_launches = 0;
inline void check_launch(const char *) {}
struct Prof {
    bool on = false;
    double ms[4] = {0, 0, 0, 0};
    double cnt[4] = {0, 0, 0, 0};
    void start(int, Stream) {}
    void stop(Stream) {}
    void resolve(Stream) {}
};
} // namespace be
#define VOR_LAUNCH(ArgsT, body, n, args, stream)                      \
    do {                                                              \
        const int _n = (int)(n);                                      \
        for (int _i = 0; _i < _n; _i++) body(args, _i);               \
        ::vor::be::g_launches++;                                      \
    } while (0)
#define VOR_LAUNCH_FULL(ArgsT, body, n, args, stream)                 \
    do {                                                              \
        const int _n = (int)(n);                                      \
        for (int _i = 0; _i < _n; _i++) body(args, _i, true);         \
        ::vor::be::g_launches++;                                      \
    } while (0)
} // namespace vor
