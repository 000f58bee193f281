// vor_common.cuh -- shared definitions for the sm_100a Delaunay engine.
//
// Every kernel is written as a `*_body(args, tid)` function plus a thin
// `__global__` wrapper.  tests/emu/ compiles the same bodies with g++
// (-DVOR_EMU) and runs them as sequential loops so the kernel LOGIC can be
// unit-tested in the CPU-only container; that emulation is test infrastructure
// and is never linked into libvoronoids_b200.so (no CPU fallback in the product).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__) && !defined(VOR_EMU)
#include <cuda_runtime.h>
#define VOR_HD __host__ __device__ __forceinline__
#define VOR_HD_NOINLINE __host__ __device__ __noinline__ inline
#define VOR_GPU 1
#else
#define VOR_HD inline
#define VOR_HD_NOINLINE inline
#define VOR_GPU 0
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(32) double4 { double x, y, z, w; };
#endif

namespace vor {

// ---- status codes (mirrors include/voronoids_b200.h)
enum : int {
    ERR_NONE = 0,
    ERR_NO_CONFLICT = 1,   // reference: panic "No simplex found" delaunay_tree.rs:53
    ERR_DEGENERATE = 2,    // reference: LU unwrap panic geometry.rs:49 (here: flat new simplex)
    ERR_DUPLICATE = 3,     // duplicate input point (undefined behaviour in the reference)
    ERR_CUDA = 4,
    ERR_OOM = 5,
    ERR_CAPACITY = 6,      // cavity larger than the overflow scratch
    ERR_RANGE = 7,         // coordinate dynamic range beyond the exact-arithmetic capacity
    ERR_OUTSIDE = 8,       // point outside the super simplex
    ERR_WALK = 9,          // visibility walk did not terminate
    ERR_ARG = 10,
};

constexpr int OWNER_FREE = 0x7fffffff;

// ---- atomics: CUDA on device, plain on host (emulation is sequential)
VOR_HD int atomic_min_i(int *p, int v) {
#ifdef __CUDA_ARCH__
    return atomicMin(p, v);
#else
    int o = *p; if (v < o) *p = v; return o;
#endif
}
VOR_HD int atomic_add_i(int *p, int v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    int o = *p; *p = o + v; return o;
#endif
}
VOR_HD unsigned atomic_add_u(unsigned *p, unsigned v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    unsigned o = *p; *p = o + v; return o;
#endif
}
VOR_HD unsigned long long atomic_add_ull(unsigned long long *p, unsigned long long v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    unsigned long long o = *p; *p = o + v; return o;
#endif
}
VOR_HD int atomic_max_i(int *p, int v) {
#ifdef __CUDA_ARCH__
    return atomicMax(p, v);
#else
    int o = *p; if (v > o) *p = v; return o;
#endif
}
VOR_HD int atomic_cas_i(int *p, int cmp, int v) {
#ifdef __CUDA_ARCH__
    return atomicCAS(p, cmp, v);
#else
    int o = *p; if (o == cmp) *p = v; return o;
#endif
}

VOR_HD uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// bijective hash on `bits` bits (odd multiply / xorshift are both invertible mod 2^bits)
VOR_HD uint32_t bij_hash(uint32_t x, int bits, uint32_t salt) {
    const uint32_t mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
    const int h = bits > 1 ? bits / 2 : 1;
    x = (x + salt) & mask;
    x = (x * 0x9E3779B1u) & mask;
    x ^= x >> h;
    x = (x * 0x85EBCA6Bu) & mask;
    x ^= x >> h;
    x = (x * 0xC2B2AE35u) & mask;
    x ^= x >> h;
    return x;
}

// int4 component access by runtime index (kept in registers by the compiler via selects)
VOR_HD int get4(const int4 &v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
VOR_HD void set4(int4 &v, int i, int val) {
    if (i == 0) v.x = val; else if (i == 1) v.y = val; else if (i == 2) v.z = val; else v.w = val;
}

// device counters, one struct per engine, lives in device memory
struct Counters {
    int ntets;        // next free simplex slot (bump allocator)
    int nslots;       // attempt slots claimed this round
    int nwinners;     // winners this round
    int nbig;         // overflow scratch slots claimed this round
    int err;          // first error code
    int ndup;         // duplicate points dropped
    int nact_out;     // compaction output count
    int oom_soft;     // a winner could not get simplex slots (host grows the store at the next sync)
    unsigned long long walk_steps;   // W: visibility-walk steps
    unsigned long long tests;        // E: in-sphere tests
    unsigned long long killed;       // K: simplices killed
    unsigned long long created;      // C: simplices created
    unsigned long long exact_calls;  // predicates that needed the exact path
    unsigned long long exact_zero;   // exact predicates that evaluated to zero
    unsigned long long attempts;     // attempt slots used (all rounds)
    unsigned long long aborted;      // attempts that lost during the flood
    unsigned long long win_total;    // points inserted (all rounds)
    unsigned long long sel_total;    // attempt slots claimed (all rounds)
    unsigned long long tests_ok;     // in-sphere tests of the attempts that completed their flood
    unsigned long long created_all;  // simplices created since the tree was made (slots are recycled: ntets is not this)
    // privatised copies of (win_total, created_all) for the cooperative commit kernel: 10M same-address atomics cost
    // ~9 ms each on the B200 (L2 serialises them); winner g adds (1 << 40 | created) to part[g % NPART][0], each part
    // in its own 32 B sector.  The host folds them (Engine::win_total / created_all).
    unsigned long long part[128][4];
    int nflag_set, nflag_done;       // points handed to / finished by the exact twin of the attempt kernel
    int sph_lo, nslow;                // simplex slots below sph_lo have their sphere block (k_spheres); slots queued for the exact twin
    int q_attempt, q_commit;          // tile queues of the tiled round kernels (each is reset by the other kernel)
    unsigned long long sph_undecided; // conflict tests the stored sphere filter left to the determinant predicate
};
constexpr int NPART = 128;

VOR_HD void set_err(Counters *c, int code) { atomic_cas_i(&c->err, 0, code); }

} // namespace vor
