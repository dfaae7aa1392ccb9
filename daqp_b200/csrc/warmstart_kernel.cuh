// daqp_b200/csrc/warmstart_kernel.cuh -- batched warm-start initialisers: the callers on the INPUT side of the hot path.
//
// The reference derives the starting working set of a warm-started solve from a primal or a dual iterate by setting bits
// in DAQPProblem.sense (daqp_primal_init_active / daqp_dual_init_active, src/api.c:577-631; used by daqp.solve(...,
// primal_start=, dual_start=), interfaces/daqp-python/daqp.pyx:24-38,404-414, and quadprog(...) in api.jl:244-248);
// the solve then activates those rows first (daqp_activate_constraints, src/auxiliary.c:399-479 -- row a13 of the solve
// kernel). Batched here so that a closed loop can stay on the device: solve -> init_active(x or lam of the previous
// step) -> solve.
//
//   primal (api.c:579-617): row i is marked ACTIVE at its upper bound when |a_i'x - bupper_i| < 1e-9, else ACTIVE + LOWER
//                           when |a_i'x - blower_i| < 1e-9; IMMUTABLE rows are left alone. a_i'x is the plain
//                           left-to-right sum of daqp_dot_inline (include/factorization.h:13-17), reproduced here with
//                           separate multiply and add so that the bits match the reference built without contraction.
//   dual   (api.c:620-631): lam_i > 1e-12 -> ACTIVE (upper), lam_i < -1e-12 -> ACTIVE + LOWER.
//
// One warp per problem. The primal pass is a single read of A (HBM-bound: m-ms rows of n doubles per problem). Tiles of
// IA_ROWS rows x up to 64 columns -- for n <= 64 a tile is IA_ROWS WHOLE rows, one contiguous piece of A -- are copied
// with cp.async straight into a padded shared-memory tile (no staging registers, every row of the tile in flight), then
// lane r walks row r of the tile left to right: the reference's summation order with coalesced global traffic.
//
// Measured on 100k C3 problems (n = 50, 150 rows; scripts/ia_probe.cu on a B200, algorithmic bytes 6.4 GB per launch):
//   32 x 32 tiles, loads through registers, 1 / 4 / 8 / 16 / 32 loads in flight per warp   3.90 / 2.25 / 1.95 / 1.60 / 1.80 ms
//   32 x 32 tiles, cp.async                                                                 1.46 ms
//   whole-row tiles, cp.async, 32 / 24 / 16 / 8 rows per tile                               1.30 / 1.14 / 1.32 / 1.33 ms
//   whole-row tiles, two or three tile buffers per warp (copy of tile t+1 behind the walk of tile t)  1.44 / 2.10 ms
// i.e. resident warps beat per-warp overlap (a second buffer halves the warps per SM), and 24 rows per tile is the sweet
// spot between the number of copies a lane has in flight and the number of tile phases: 5.6 TB/s in the probe (zero-filled
// arrays), 1.24 ms = 5.2 TB/s = 79 % of the measured HBM peak (6553 GB/s) with random data through the product entry point.
#pragma once
#include "common.cuh"

namespace dq {

struct InitActiveArgs {
    int N, n, m, ms;
    const double* x;      // [N][n] or nullptr
    const double* lam;    // [N][m] or nullptr (used when x == nullptr)
    const double* A;      // [N][m - ms][n]
    const double *bupper, *blower; // [N][m]
    int* sense;           // [N][m] in / out
};

#ifndef IA_WARPS_N
#define IA_WARPS_N 8
#endif
#ifndef IA_ROWS
#define IA_ROWS 24 // rows per tile; lanes >= IA_ROWS idle in the row walk (experiment knob, scripts/ia_probe.cu)
#endif
constexpr int IA_WARPS = IA_WARPS_N;

// doubles per tile row: odd, so that the lanes, each walking its own row, hit different banks
__host__ __device__ inline int ia_pitch(int n) { return (n < 64 ? n : 64) | 1; }
__host__ __device__ inline size_t ia_smem_per_warp(int n) { return ((size_t)IA_ROWS * ia_pitch(n) + n) * sizeof(double); }

__device__ __forceinline__ int init_active_bits(int s, double ax, double bu, double bl) {
    const double tol = 1e-9; // api.c:582
    if (s & B_IMMUTABLE) return s;
    double slack = ax - bu;
    if (slack < tol && slack > -tol) return (s | B_ACTIVE) & ~B_LOWER;
    slack = ax - bl;
    if (slack < tol && slack > -tol) return s | (B_ACTIVE + B_LOWER);
    return s;
}

__global__ void __launch_bounds__(32 * IA_WARPS) init_active_kernel(const InitActiveArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int n = a.n, m = a.m, ms = a.ms, mA = m - ms, pitch = ia_pitch(n);
    double* tile = reinterpret_cast<double*>(smem_raw + ia_smem_per_warp(n) * wib);
    double* xs = tile + (size_t)IA_ROWS * pitch;
    const int nchunk = (n + 63) >> 6;
    for (int p = blockIdx.x * IA_WARPS + wib; p < a.N; p += gridDim.x * IA_WARPS) {
        int* se = a.sense + (size_t)p * m;
        if (a.x == nullptr) { // dual iterate: api.c:620-631
            const double* lam = a.lam + (size_t)p * m;
            for (int i = lane; i < m; i += 32) {
                int s = se[i];
                if (!(s & B_IMMUTABLE)) {
                    const double l = lam[i];
                    if (l > 1e-12) s = (s | B_ACTIVE) & ~B_LOWER;
                    else if (l < -1e-12) s |= B_ACTIVE + B_LOWER;
                    se[i] = s;
                }
            }
            continue;
        }
        const double* x = a.x + (size_t)p * n;
        const double* bu = a.bupper + (size_t)p * m;
        const double* bl = a.blower + (size_t)p * m;
        const double* A = a.A + (size_t)p * mA * n;
        for (int i = lane; i < n; i += 32) xs[i] = x[i];
        __syncwarp();
        for (int i = lane; i < ms; i += 32) se[i] = init_active_bits(se[i], xs[i], bu[i], bl[i]); // api.c:585-598
        for (int r0 = 0; r0 < mA; r0 += IA_ROWS) {
            const int nr = min(IA_ROWS, mA - r0);
            double acc = 0.0;
            for (int ch = 0; ch < nchunk; ch++) {
                const int c0 = ch << 6, nc = min(64, n - c0);
                // rows r0 .. r0+nr-1, columns c0 .. c0+nc-1: asynchronous copies straight into the tile
                {
                    const unsigned dst = smem_u32(tile);
                    const double* src = A + (size_t)r0 * n + c0;
                    for (int rr = 0; rr < nr; rr++) {
                        if (lane < nc)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * (rr * pitch + lane)), "l"(src + (size_t)rr * n + lane) : "memory");
                        if (lane + 32 < nc)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * (rr * pitch + lane + 32)), "l"(src + (size_t)rr * n + lane + 32) : "memory");
                    }
                    cp_async_commit();
                    cp_async_wait<0>();
                }
                __syncwarp();
                if (lane < nr) {
                    const double* row = tile + lane * pitch;
                    for (int c = 0; c < nc; c++) acc = __dadd_rn(acc, __dmul_rn(xs[c0 + c], row[c]));
                }
                __syncwarp();
            }
            if (lane < nr) { // api.c:601-616
                const int i = ms + r0 + lane;
                se[i] = init_active_bits(se[i], acc, bu[i], bl[i]);
            }
        }
        __syncwarp();
    }
}

// ---- daqp_first_violating for a batch of points (reference src/api.c:562-574): index of the first constraint of
// {bl <= [I(ms); A] x <= bu} that x violates by more than tol, or m. One warp per point; lane r walks rows r, r + 32, ...
// left to right with separate multiply and add (the reference's loop order, bit for bit), keeps the smallest violated
// index it meets, and the warp takes the minimum.
__global__ void __launch_bounds__(256) first_violating_kernel(int N, int n, int m, int ms, const double* x, const double* A,
                                                               const double* bu, const double* bl, double tol, int* out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double* xs = reinterpret_cast<double*>(smem_raw) + (size_t)wib * n;
    for (int p = blockIdx.x * wpb + wib; p < N; p += gridDim.x * wpb) {
        const double* xp = x + (size_t)p * n;
        for (int j = lane; j < n; j += 32) xs[j] = xp[j];
        __syncwarp();
        int first = m;
        for (int i = lane; i < m && first == m; i += 32) {
            double v;
            if (i < ms) v = xs[i];
            else {
                const double* row = A + (size_t)(i - ms) * n;
                v = 0;
                for (int j = 0; j < n; j++) v = __dadd_rn(v, __dmul_rn(__ldg(row + j), xs[j]));
            }
            if (v > bu[i] + tol || v < bl[i] - tol) first = i;
        }
        first = __reduce_min_sync(FULL, first);
        if (lane == 0) out[p] = first;
        __syncwarp();
    }
}

} // namespace dq
