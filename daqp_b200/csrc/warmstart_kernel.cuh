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
// One warp per problem. The primal pass is a single read of A (HBM-bound: m-ms rows of n doubles per problem): 32 x 32
// tiles go through a padded shared-memory tile with coalesced 256-byte row segments, then lane r walks row r of the tile
// left to right -- the reference's summation order with coalesced global loads.
// Measured on 100k C3 problems (scripts/ia_probe.cu, B200): tile loads through registers, 1 / 4 / 8 / 16 / 32 loads in
// flight per warp: 3.90 / 2.25 / 1.95 / 1.60 / 1.80 ms; cp.async straight into the tile (all rows in flight, no staging
// registers): 1.46 ms = 4.4 TB/s, 67 % of the measured HBM peak; 24 resident warps per SM in both (16: 2.2 ms). What is
// left is the per-warp alternation of a copy phase and a compute phase (next: double-buffered tiles).
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
#ifndef IA_UNROLL
#define IA_UNROLL 4 // loads of a tile kept in flight per warp (experiment knob, scripts/ia_probe.cu)
#endif
#ifndef IA_MODE
#define IA_MODE 1 // 0: tile loads through registers, 1: cp.async into the tile
#endif
constexpr int IA_WARPS = IA_WARPS_N;
constexpr int IA_UNROLL_C = IA_UNROLL;
constexpr int IA_PITCH = 33; // doubles per tile row: odd, so that the 32 lanes, each walking its own row, hit different banks

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
    const int n = a.n, m = a.m, ms = a.ms, mA = m - ms;
    double* tile = reinterpret_cast<double*>(smem_raw) + (size_t)wib * (32 * IA_PITCH + n);
    double* xs = tile + 32 * IA_PITCH;
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
        for (int i = lane; i < n; i += 32) xs[i] = x[i];
        __syncwarp();
        for (int i = lane; i < ms; i += 32) se[i] = init_active_bits(se[i], xs[i], bu[i], bl[i]); // api.c:585-598
        const double* A = a.A + (size_t)p * mA * n;
        for (int r0 = 0; r0 < mA; r0 += 32) {
            const int nr = min(32, mA - r0);
            double acc = 0.0;
            for (int c0 = 0; c0 < n; c0 += 32) {
                const int nc = min(32, n - c0);
                // rows r0 .. r0+nr-1, columns c0 .. c0+nc-1: one coalesced segment per row. (Issuing all 32 loads of a tile
                // before the first store -- fully unrolled -- measured SLOWER: 2.20 ms vs 1.75 ms per 100k C3 problems.)
#if IA_MODE == 0
#pragma unroll IA_UNROLL_C
                for (int rr = 0; rr < nr; rr++)
                    if (lane < nc) tile[rr * IA_PITCH + lane] = __ldg(A + (size_t)(r0 + rr) * n + c0 + lane);
#else
                // asynchronous copies straight into the tile (LDGSTS): no staging registers, so all rows of the tile are in
                // flight at once
                {
                    const unsigned dst = smem_u32(tile) + 8u * lane;
                    const double* src = A + (size_t)r0 * n + c0 + lane;
                    if (lane < nc)
                        for (int rr = 0; rr < nr; rr++)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * IA_PITCH * rr), "l"(src + (size_t)rr * n) : "memory");
                    cp_async_commit();
                    cp_async_wait<0>();
                }
#endif
                __syncwarp();
                if (lane < nr)
                    for (int c = 0; c < nc; c++) acc = __dadd_rn(acc, __dmul_rn(xs[c0 + c], tile[lane * IA_PITCH + c]));
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

} // namespace dq
