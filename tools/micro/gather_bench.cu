// Scratch micro-benchmark (not part of the product): throughput of independent random gathers of 32 / 64 / 128 B records
// from footprints of 0.25 .. 16 GB on one B200.  Answers: is the engine's record layout (32 B record + 8 B owner pair in
// two arrays) bound by DRAM access RATE rather than bytes?   nvcc -arch=sm_100a -O3 gather_bench.cu -o gather_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
template <int Q>   // Q = int4 per record (2 = 32 B, 4 = 64 B, 8 = 128 B)
__global__ void gather(const int4 *a, uint64_t nrec, int iters, int *out) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    int acc = 0;
    for (int it = 0; it < iters; it++) {
        const uint64_t r = mix(tid * 0x9E3779B97F4A7C15ULL + it) % nrec;
        const int4 *p = a + r * Q;
#pragma unroll
        for (int q = 0; q < Q; q++) { const int4 v = __ldcg(p + q); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    }
    if (acc == 0x12345678) out[0] = acc;
}
// 32 B records fetched with ONE 256-bit load per record (sm_100: ld.global.v8.b32)
__device__ __forceinline__ void ld256(const void *p, int4 &a, int4 &b) {
    asm volatile("ld.global.cg.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
template <int Q>   // Q = 256-bit words per record
__global__ void gather256(const int4 *a, uint64_t nrec, int iters, int *out) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    int acc = 0;
    for (int it = 0; it < iters; it++) {
        const uint64_t r = mix(tid * 0x9E3779B97F4A7C15ULL + it) % nrec;
        const int4 *p = a + r * 2 * Q;
#pragma unroll
        for (int q = 0; q < Q; q++) { int4 v0, v1; ld256(p + 2 * q, v0, v1); acc ^= v0.x ^ v0.y ^ v0.z ^ v0.w ^ v1.x ^ v1.y ^ v1.z ^ v1.w; }
    }
    if (acc == 0x12345678) out[0] = acc;
}
// two arrays: 32 B record + 8 B owner pair at the same index (the engine's current layout)
__global__ void gather_split(const int4 *a, const int2 *o, uint64_t nrec, int iters, int *out) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    int acc = 0;
    for (int it = 0; it < iters; it++) {
        const uint64_t r = mix(tid * 0x9E3779B97F4A7C15ULL + it) % nrec;
        const int4 v0 = __ldcg(a + r * 2), v1 = __ldcg(a + r * 2 + 1);
        const int2 w = __ldcg(o + r);
        acc ^= v0.x ^ v0.y ^ v0.z ^ v0.w ^ v1.x ^ v1.y ^ v1.z ^ v1.w ^ w.x ^ w.y;
    }
    if (acc == 0x12345678) out[0] = acc;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    const size_t maxBytes = 16ULL << 30;
    int4 *a; int2 *o; int *out;
    if (cudaMalloc(&a, maxBytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&o, maxBytes / 4); cudaMalloc(&out, 4);
    cudaMemset(a, 1, maxBytes); cudaMemset(o, 1, maxBytes / 4);
    const int threads = 256, blocks = 148 * 16, iters = 64;
    const double nacc = (double)threads * blocks * iters;
    for (double gb : {0.25, 1.0, 4.0, 16.0}) {
        const size_t bytes = (size_t)(gb * (1ULL << 30));
        float ms;
        ms = timeit([&] { gather<2><<<blocks, threads>>>(a, bytes / 32, iters, out); });
        printf("footprint %5.2f GB  32B records : %7.2f G acc/s  %7.1f GB/s\n", gb, nacc / ms / 1e6, nacc * 32 / ms / 1e6);
        ms = timeit([&] { gather256<1><<<blocks, threads>>>(a, bytes / 32, iters, out); });
        printf("footprint %5.2f GB  32B rec, 1 x 256-bit load : %7.2f G acc/s  %7.1f GB/s\n", gb, nacc / ms / 1e6, nacc * 32 / ms / 1e6);
        ms = timeit([&] { gather256<2><<<blocks, threads>>>(a, bytes / 64, iters, out); });
        printf("footprint %5.2f GB  64B rec, 2 x 256-bit load : %7.2f G acc/s  %7.1f GB/s\n", gb, nacc / ms / 1e6, nacc * 64 / ms / 1e6);
        ms = timeit([&] { gather<4><<<blocks, threads>>>(a, bytes / 64, iters, out); });
        printf("footprint %5.2f GB  64B records : %7.2f G acc/s  %7.1f GB/s\n", gb, nacc / ms / 1e6, nacc * 64 / ms / 1e6);
        ms = timeit([&] { gather<8><<<blocks, threads>>>(a, bytes / 128, iters, out); });
        printf("footprint %5.2f GB 128B records : %7.2f G acc/s  %7.1f GB/s\n", gb, nacc / ms / 1e6, nacc * 128 / ms / 1e6);
        ms = timeit([&] { gather_split<<<blocks, threads>>>(a, o, bytes / 32, iters, out); });
        printf("footprint %5.2f GB 32B+8B split : %7.2f G acc/s  %7.1f GB/s (40 B useful)\n", gb * 1.25, nacc / ms / 1e6, nacc * 40 / ms / 1e6);
    }
    return 0;
}
