// tools/experiments/atomic_rate.cu -- how many warp-aggregated queue reservations per second does ONE address take?
// k_shade reserves its output slots with one 64-bit atomicAdd per warp iteration on a single word (and one 32-bit
// atomicAdd on the traverse-queue count): 0.5 M of each per launch of the first bounce.  This measures the rate at which
// the L2 serves same-address atomics whose result is needed (lane 0 of every warp, result broadcast by shuffle), with
// some independent arithmetic per iteration so that the warps do not simply queue up.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o atomic_rate atomic_rate.cu && ./atomic_rate
#include <cuda_runtime.h>
#include <cstdio>
template <int KIND>
__global__ void k(unsigned long long* c64, unsigned* c32, int iters, int work, float* sink) {
    const unsigned lane = threadIdx.x & 31;
    float acc = threadIdx.x;
    unsigned long long got = 0;
    for (int i = 0; i < iters; ++i) {
        unsigned long long r = 0;
        if (lane == 0) {
            if (KIND == 0) r = atomicAdd(c64, 17ull | (15ull << 32));
            if (KIND == 1) { r = atomicAdd(c64, 17ull | (15ull << 32)); r += atomicAdd(c32, 20u); }
            if (KIND == 2) r = atomicAdd(c64 + 32 * (blockIdx.x & 7), 17ull);          // eight addresses in different lines
        }
        for (int w = 0; w < work; ++w) acc = fmaf(acc, 1.0001f, 0.5f);                  // independent work under the atomic
        r = __shfl_sync(0xFFFFFFFFu, r, 0);
        got += r;
    }
    if (got == 0x1234567 && acc == 3.f) *sink = acc;
}
int main() {
    unsigned long long* c64; unsigned* c32; float* sink;
    cudaMalloc(&c64, 4096 * 8); cudaMalloc(&c32, 4096); cudaMalloc(&sink, 4);
    cudaMemset(c64, 0, 4096 * 8); cudaMemset(c32, 0, 4096);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = 148 * 9, iters = 100;   // k_shade's residency: 9 blocks of 4 warps per SM
    for (int work : {0, 200, 900}) {
        for (int kind = 0; kind < 3; ++kind) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a);
                if (kind == 0) k<0><<<blocks, 128>>>(c64, c32, iters, work, sink);
                if (kind == 1) k<1><<<blocks, 128>>>(c64, c32, iters, work, sink);
                if (kind == 2) k<2><<<blocks, 128>>>(c64, c32, iters, work, sink);
                cudaEventRecord(b); cudaEventSynchronize(b);
            }
            float ms; cudaEventElapsedTime(&ms, a, b);
            const double n = (double)blocks * 4 * iters;
            printf("work %4d FFMA per iteration, %s: %.3f ms for %.0f warp reservations = %.1f M/s (%.2f ns each)\n", work,
                   kind == 0 ? "one 64-bit atomic, one address " : kind == 1 ? "64-bit + 32-bit, two addresses " : "one 64-bit atomic, 8 addresses  ",
                   ms, n, n / ms / 1e3, ms * 1e6 / n);
        }
    }
    return 0;
}
