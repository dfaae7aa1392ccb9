// daqp_b200/csrc/minrep_kernel.cuh -- batched minimal representation of polyhedra.
//
// The reference's daqp_minrep (src/api.c:531-556) wraps the polyhedron {x : [I(ms); A] x <= b} in a workspace with
// M = A (NOT normalised, scaling == NULL), Rinv == NULL (simple bounds are unit rows), v == NULL, dlower = -inf, and
// daqp_minrep_work (src/utils.c:808-835) then solves ONE LDP PER CONSTRAINT, one after the other: constraint i is made
// an active equality (sense = ACTIVE + IMMUTABLE, added with lam = 1) and the dual active-set loop decides whether
// {a_i' u = b_i} meets the polyhedron; DAQP_EXIT_INFEASIBLE marks the constraint redundant.
//
// Here the m LDPs of a polyhedron -- and of every polyhedron of a batch -- run CONCURRENTLY in the solve kernel, one
// warp per LDP, in its shared-matrix mode (LdpArgs::grp = m): the kernels below write the constraint matrix of each
// polyhedron ONCE in the two layouts the solve kernel streams (row-major Mr, column-major Mt; no fp32 screening
// copy, whose error bound assumes unit rows) and, per LDP, only the bounds, the sense bytes and the hand-over flag.
// The m warps of a polyhedron then read the same matrix out of L2.
//
// Sequential vs concurrent: the reference drops a constraint it has found redundant from the LDPs that follow (it stays
// IMMUTABLE and inactive, i.e. ignored) and skips constraints that were active at an earlier optimum. For a NON-EMPTY
// polyhedron both orders name the same constraints: if H_i misses P without i and H_j misses P without j, a point of
// H_j inside P without {i, j} would, on the segment to any point of P, cross H_i inside P without i. A constraint active
// at an optimum of another LDP supports the polyhedron, so skipping it changes nothing either. For an EMPTY polyhedron
// every probe is infeasible and the reference's answer depends on its order: it drops constraints from the front until
// what is left is non-empty. The host entry point reproduces that with `dropped`: a polyhedron whose probes all came
// back infeasible is run again with its first remaining constraint marked dropped (sense = IMMUTABLE, not probed), until
// some probe is feasible. Tests pin the result on the reference's own daqp_minrep output.
#pragma once
#include "common.cuh"

namespace dq {

struct MinrepArgs {
    int P, n, m, ms, ldm, ldn;
    const double* A; // [P][m - ms][n]
    const double* b; // [P][m]
    const unsigned char* dropped; // [P][m] or nullptr: constraints already found redundant and taken out (see below)
    // per polyhedron
    double *Mt, *Mr, *scaling, *Rinv;
    // per LDP (P * m of them)
    double *dupper, *dlower;
    unsigned char* sense;
    int *setup_flag, *exitflag, *iter;
};

// One CTA per polyhedron (grid-stride). All stores are coalesced; A is read once from HBM (the transposed read of
// the column-major copy hits L1/L2).
__global__ void __launch_bounds__(256) minrep_prep_kernel(const MinrepArgs a) {
    const int n = a.n, m = a.m, ms = a.ms, mA = m - ms, ldm = a.ldm, ldn = a.ldn, ntri = n * (n + 1) / 2;
    for (int q = blockIdx.x; q < a.P; q += gridDim.x) {
        const double* A = a.A + (size_t)q * mA * n;
        const double* b = a.b + (size_t)q * m;
        double* Mr = a.Mr + (size_t)q * m * ldn;
        double* Mt = a.Mt + (size_t)q * n * ldm;
        double* sc = a.scaling + (size_t)q * ldm;
        double* Ri = a.Rinv + (size_t)q * ntri;
        for (int idx = threadIdx.x; idx < m * ldn; idx += blockDim.x) {
            const int r = idx / ldn, c = idx - r * ldn;
            Mr[idx] = (c >= n) ? 0.0 : (r < ms ? (c == r ? 1.0 : 0.0) : A[(size_t)(r - ms) * n + c]);
        }
        for (int idx = threadIdx.x; idx < n * ldm; idx += blockDim.x) {
            const int c = idx / ldm, r = idx - c * ldm;
            Mt[idx] = (r >= m) ? 0.0 : (r < ms ? (c == r ? 1.0 : 0.0) : A[(size_t)(r - ms) * n + c]);
        }
        for (int r = threadIdx.x; r < ldm; r += blockDim.x) sc[r] = 1.0; // scaling == NULL: bound = -primal_tol
        for (int idx = threadIdx.x; idx < ntri; idx += blockDim.x) Ri[idx] = 0.0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) Ri[roff(i, n) + i] = 1.0; // Rinv == NULL: x = u
        // the m LDPs: bounds of the polyhedron with row i turned into an equality at its upper bound
        const size_t l0 = (size_t)q * m;
        const unsigned char* dr = a.dropped ? a.dropped + l0 : nullptr;
        for (int idx = threadIdx.x; idx < m * ldm; idx += blockDim.x) {
            const int i = idx / ldm, r = idx - i * ldm;
            const double br = r < m ? b[r] : 0.0;
            a.dupper[l0 * ldm + idx] = br;
            a.dlower[l0 * ldm + idx] = (r == i) ? br : (r < m ? -1e30 : 0.0);
            // utils.c:821-824: a constraint found redundant earlier stays IMMUTABLE and inactive -- every scan skips it
            const bool gone = dr && r < m && dr[r];
            a.sense[l0 * ldm + idx] = (r == i) ? (unsigned char)(B_ACTIVE + B_IMMUTABLE) : (unsigned char)(gone ? B_IMMUTABLE : 0);
        }
        // every LDP goes to the solve kernel with its probed row pre-activated (utils.c:817-818). No shortcut for
        // degenerate rows: the reference's first LDL' append returns before its singularity test (factorization.c:56), so
        // an all-zero row runs the loop with a zero pivot and whatever that produces is what the solve kernel reproduces.
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
            const bool gone = dr && dr[i]; // not probed again (utils.c:815): reported redundant
            a.setup_flag[l0 + i] = gone ? EXIT_INFEASIBLE : SETUP_SOLVE_ACTIVATE;
            a.exitflag[l0 + i] = gone ? EXIT_INFEASIBLE : 0;
            a.iter[l0 + i] = 0;
        }
        __syncthreads();
    }
}

// ---- raw LDPs, one per polyhedron: what daqp_ldp runs on a hand-filled workspace (interfaces/daqp-julia/src/api.jl:
// 440-459 -- M = A as given, no normalisation, Rinv == NULL, v == NULL). Same matrix layouts as above, then ONE problem
// per polyhedron with the caller's bounds and sense bits; rows with the ACTIVE bit are activated first (a13). Soft and
// binary bits are not part of this form: flagged -8.
__global__ void __launch_bounds__(256) ldp_prep_kernel(const MinrepArgs a, const double* bl, const int* sense_in, double* vzero) {
    const int n = a.n, m = a.m, ms = a.ms, mA = m - ms, ldm = a.ldm, ldn = a.ldn, ntri = n * (n + 1) / 2;
    __shared__ int bad;
    for (int q = blockIdx.x; q < a.P; q += gridDim.x) {
        const double* A = a.A + (size_t)q * mA * n;
        double* Mr = a.Mr + (size_t)q * m * ldn;
        double* Mt = a.Mt + (size_t)q * n * ldm;
        double* sc = a.scaling + (size_t)q * ldm;
        double* Ri = a.Rinv + (size_t)q * ntri;
        if (threadIdx.x == 0) bad = 0;
        for (int idx = threadIdx.x; idx < m * ldn; idx += blockDim.x) {
            const int r = idx / ldn, c = idx - r * ldn;
            Mr[idx] = (c >= n) ? 0.0 : (r < ms ? (c == r ? 1.0 : 0.0) : A[(size_t)(r - ms) * n + c]);
        }
        for (int idx = threadIdx.x; idx < n * ldm; idx += blockDim.x) {
            const int c = idx / ldm, r = idx - c * ldm;
            Mt[idx] = (r >= m) ? 0.0 : (r < ms ? (c == r ? 1.0 : 0.0) : A[(size_t)(r - ms) * n + c]);
        }
        for (int idx = threadIdx.x; idx < ntri; idx += blockDim.x) Ri[idx] = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) vzero[(size_t)q * n + i] = 0.0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) Ri[roff(i, n) + i] = 1.0;
        for (int r = threadIdx.x; r < ldm; r += blockDim.x) {
            sc[r] = 1.0;
            const int sb = (r < m && sense_in) ? sense_in[(size_t)q * m + r] : 0;
            if (sb & (B_SOFT | B_BINARY | ~63)) bad = 1;
            a.dupper[(size_t)q * ldm + r] = r < m ? a.b[(size_t)q * m + r] : 0.0;
            a.dlower[(size_t)q * ldm + r] = r < m ? bl[(size_t)q * m + r] : 0.0;
            a.sense[(size_t)q * ldm + r] = (unsigned char)(sb & 63);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            a.setup_flag[q] = bad ? EXIT_UNSUPPORTED : SETUP_SOLVE_ACTIVATE;
            a.exitflag[q] = bad ? EXIT_UNSUPPORTED : 0;
            a.iter[q] = 0;
        }
        __syncthreads();
    }
}

// is_redundant = 1 iff the LDP ended DAQP_EXIT_INFEASIBLE (utils.c:821-824); anything else is 0 (utils.c:825-831)
__global__ void minrep_finish_kernel(const int* exitflag, int* is_redundant, size_t count) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        is_redundant[i] = exitflag[i] == EXIT_INFEASIBLE ? 1 : 0;
}

} // namespace dq
