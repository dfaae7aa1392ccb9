// Memory-system probe for the feasibility-scan access pattern: every warp repeatedly streams its own private
// 60.8 KB column-major matrix (n=50 columns x 152 rows fp64) with 128-bit loads, DEPTH loads in flight per lane.
// Reports GB/s for several depths / warps-per-SM / cache policies. Not part of the product.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void ld2(const double* p, double& a, double& b, uint64_t pol, int hint) {
    if (hint) asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(a), "=d"(b) : "l"(p), "l"(pol));
    else asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
}

template <int DEPTH>
__global__ void probe(const double* __restrict__ base, int P, int elems, int visits, int delay, int hint,
                      int* counter, double* sink) {
    const int lane = threadIdx.x & 31;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    double acc = 0;
    const int nvec = elems / 2;          // double2 per region
    for (;;) {
        int p = 0;
        if (lane == 0) p = atomicAdd(counter, 1);
        p = __shfl_sync(0xffffffffu, p, 0);
        if (p >= P) break;
        const double* reg = base + (size_t)p * elems;
        for (int v = 0; v < visits; v++) {
            double bx[DEPTH], by[DEPTH];
            int idx = lane;
#pragma unroll
            for (int i = 0; i < DEPTH; i++) { bx[i] = by[i] = 0; if (idx + 32 * i < nvec) ld2(reg + 2 * (idx + 32 * i), bx[i], by[i], pol, hint); }
            for (; idx < nvec; idx += 32 * DEPTH) {
#pragma unroll
                for (int i = 0; i < DEPTH; i++) {
                    acc += bx[i] * 1.0000001 + by[i];
                    const int nx = idx + 32 * (i + DEPTH);
                    bx[i] = by[i] = 0;
                    if (nx < nvec) ld2(reg + 2 * nx, bx[i], by[i], pol, hint);
                }
            }
            // emulate the non-streaming phases of an iteration
            long long t0 = clock64();
            while (clock64() - t0 < delay) { }
        }
    }
    if (acc == 12345.678) sink[0] = acc;
}

int main(int argc, char** argv) {
    const int P = 20000, elems = 50 * 152, visits = 20;
    double* d; int* counter; double* sink;
    cudaMalloc(&d, (size_t)P * elems * sizeof(double));
    cudaMemset(d, 0, (size_t)P * elems * sizeof(double));
    cudaMalloc(&counter, 4); cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)P * elems * 8 * visits;
    for (int hint = 0; hint < 2; hint++)
    for (int delay : {0, 20000})
    for (int warps : {8, 16, 32})
    for (int depth : {4, 8, 12, 16, 24}) {
        cudaMemset(counter, 0, 4);
        cudaEventRecord(e0);
        switch (depth) {
            case 4: probe<4><<<148, 32 * warps>>>(d, P, elems, visits, delay, hint, counter, sink); break;
            case 8: probe<8><<<148, 32 * warps>>>(d, P, elems, visits, delay, hint, counter, sink); break;
            case 12: probe<12><<<148, 32 * warps>>>(d, P, elems, visits, delay, hint, counter, sink); break;
            case 16: probe<16><<<148, 32 * warps>>>(d, P, elems, visits, delay, hint, counter, sink); break;
            default: probe<24><<<148, 32 * warps>>>(d, P, elems, visits, delay, hint, counter, sink); break;
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("hint=%d delay=%5d warps/SM=%2d depth=%2d  %8.2f ms  %7.1f GB/s  (%s)\n", hint, delay, warps, depth, ms,
               bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
