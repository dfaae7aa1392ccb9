// Standalone timing probe for init_active_kernel (knobs: -DIA_ROWS=.. -DIA_WARPS_N=..). Built in the container, run on
// the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DIA_ROWS=24 -o build_variants/ia_r24 scripts/ia_probe.cu
#include "../daqp_b200/csrc/warmstart_kernel.cuh"
#include <cstdio>
using namespace dq;
int main() {
    const int N = 100000, n = 50, m = 150, ms = 0;
    double *A, *x, *bu, *bl; int* se;
    cudaMalloc(&A, (size_t)N * m * n * 8); cudaMalloc(&x, (size_t)N * n * 8);
    cudaMalloc(&bu, (size_t)N * m * 8); cudaMalloc(&bl, (size_t)N * m * 8); cudaMalloc(&se, (size_t)N * m * 4);
    cudaMemset(A, 0, (size_t)N * m * n * 8); cudaMemset(x, 0, (size_t)N * n * 8);
    cudaMemset(bu, 0, (size_t)N * m * 8); cudaMemset(bl, 0, (size_t)N * m * 8); cudaMemset(se, 0, (size_t)N * m * 4);
    InitActiveArgs ia{N, n, m, ms, x, nullptr, A, bu, bl, se};
    const size_t smem = ia_smem_per_warp(n) * IA_WARPS;
    cudaFuncSetAttribute(init_active_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, init_active_kernel, 32 * IA_WARPS, smem);
    const int grid = sms * occ;
    for (int i = 0; i < 3; i++) init_active_kernel<<<grid, 32 * IA_WARPS, smem>>>(ia);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; i++) init_active_kernel<<<grid, 32 * IA_WARPS, smem>>>(ia);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float t; cudaEventElapsedTime(&t, e0, e1); t /= 10;
    const double bytes = (double)N * (m * n * 8 + n * 8 + 2 * m * 8 + 2 * m * 4);
    printf("rows %d warps %d occ %d grid %d: %.3f ms  %.0f GB/s  err=%s\n", IA_ROWS, IA_WARPS, occ, grid, t, bytes / t / 1e6,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
