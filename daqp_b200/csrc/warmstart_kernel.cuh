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

constexpr int IA_WARPS = 8;
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
                // rows r0 .. r0+nr-1, columns c0 .. c0+nc-1: one coalesced segment per row
                for (int rr = 0; rr < nr; rr++)
                    if (lane < nc) tile[rr * IA_PITCH + lane] = __ldg(A + (size_t)(r0 + rr) * n + c0 + lane);
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
