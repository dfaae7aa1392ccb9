// daqp_b200/csrc/setup_kernel.cuh -- batched QP -> LDP transform, one warp per problem.
//
// Covers SURVEY.md §8(a) rows a17..a19 exactly as daqp_quadprog() drives them (reference src/utils.c:58-221 with
// mask = Rinv|M|v|d|sense|unconstrained, src/api.c:62-79,163-209):
//   check_bounds (utils.c:546-567) -> symmetrise + Cholesky + triangular inverse (utils.c:223-391)
//   -> v = R^-T f (utils.c:474-497) -> unconstrained-optimum shortcut (utils.c:618-687)
//   -> M = A R^-1, row normalisation (utils.c:434-472,586-613) -> normalise simple-bound rows (utils.c:569-585)
//   -> d (utils.c:151-158 after the shortcut, utils.c:499-544 otherwise).
// Output is the device-side LDP the solve kernel consumes: the constraint matrix in BOTH layouts (column-major
// Mt for the streaming scan, row-major Mr for active-row passes), bounds, scaling, packed R^-1, v, sense bytes.
#pragma once
#include "common.cuh"

namespace dq {

template <typename T>
struct SetupArgs {
    int P, n, m, ms, ldm, ldn;
    const T *H, *f, *A, *bupper, *blower; // packed inputs; f may be nullptr
    const int* sense_in;                  // [P][m] or nullptr
    T *Mt, *Mr, *dupper, *dlower, *scaling, *Rinv, *v;
    float* Mt32;                          // [P][ceil(n/4)][m][4] fp32 copy of M, quad layout (screening scan) or nullptr
    unsigned char* sense;                 // [P][ldm]
    int* setup_flag;                      // [P]
    T *x, *lam, *fval;                    // final outputs (written here only for problems finished by the setup)
    int *exitflag, *iter;
    int* work_counter;
    T* soft_slack;                        // [P] or nullptr
    // persistent-workspace mode (reference setup_daqp: no unconstrained shortcut, d = b * scaling + M v, utils.c:499-544)
    int no_shortcut;
    unsigned char* sense_static;          // [P][ldm] or nullptr: user sense bits with zero rows marked IMMUTABLE, WITHOUT
                                          // the equality marks check_bounds derives from the current bounds
    int ns_max;                           // upper bound on soft constraints per problem (sizes the solve kernel's factor)
    // hand-over between the two kernels of the split transform (setup2_kernel.cuh); unused by the fused kernel
    T* xu;                                // [P][n]  unconstrained optimum
    int* info;                            // [P][4]  {flag, SI_* bits, -, -}
    T* vnorm;                             // [P]     |v|^2
    DevSettings<T> st;
};

template <typename T> struct Pair;
template <> struct Pair<double> { using type = double2; };
template <> struct Pair<float> { using type = float2; };
constexpr int SETUP_RB = 8; // rows of A processed together (two 32-byte sectors of Mt per lane and column)

template <typename T>
__host__ __device__ inline size_t setup_smem_per_warp(int n) {
    size_t e = (size_t)n * (n + 1) / 2 + (size_t)n * (2 + SETUP_RB);
    return (e * sizeof(T) + 15) / 16 * 16;
}

// Cooperative asynchronous copy of a contiguous block of `bytes` (a multiple of 8) from global to shared memory:
// 16-byte cp.async when source and size allow it, 8-byte otherwise (odd n, or caller arrays that are only 8-byte
// aligned). One commit group; the caller waits with cp_async_wait + __syncwarp.
__device__ __forceinline__ void async_block_copy(unsigned dst, const char* src, unsigned bytes, int lane) {
    if (((reinterpret_cast<size_t>(src) | bytes) & 15) == 0) {
        for (unsigned off = 16u * lane; off < bytes; off += 512u) cp_async16(dst + off, src + off);
    } else {
        for (unsigned off = 8u * lane; off < bytes; off += 256u)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + off), "l"(src + off) : "memory");
    }
    cp_async_commit();
}

template <typename T, int NGS>
__global__ void __launch_bounds__(512, 1) qp_setup_kernel(const SetupArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int n = a.n, m = a.m, ms = a.ms, mA = a.m - a.ms, ldm = a.ldm, ldn = a.ldn;
    const int ntri = n * (n + 1) / 2;
    T* arow = reinterpret_cast<T*>(smem_raw + setup_smem_per_warp<T>(n) * wib); // SETUP_RB staged rows of A (16-byte aligned)
    T* R = arow + (size_t)n * SETUP_RB;
    T* vv = R + ntri;          // f, then v = R^-T f
    T* xu = vv + n;            // unconstrained optimum
    const DevSettings<T>& st = a.st;

    for (;;) {
        int p = 0;
        if (lane == 0) p = atomicAdd(a.work_counter, 1);
        p = __shfl_sync(FULL, p, 0);
        if (p >= a.P) break;

        const T* H = a.H + (size_t)p * n * n;
        const T* f = a.f ? a.f + (size_t)p * n : nullptr;
        const T* A = a.A + (size_t)p * mA * n;
        const T* bu = a.bupper + (size_t)p * m;
        const T* bl = a.blower + (size_t)p * m;
        const int* sin = a.sense_in ? a.sense_in + (size_t)p * m : nullptr;
        T* Mt = a.Mt + (size_t)p * n * ldm;
        float* Mt32 = a.Mt32 ? a.Mt32 + (size_t)p * ((n + 3) / 4) * m * 4 : nullptr; // quad layout, see scan_screen
        T* Mr = a.Mr + (size_t)p * m * ldn;
        T* du = a.dupper + (size_t)p * ldm;
        T* dl = a.dlower + (size_t)p * ldm;
        T* sc = a.scaling + (size_t)p * ldm;
        unsigned char* so = a.sense + (size_t)p * ldm;
        T* Rg = a.Rinv + (size_t)p * ntri;
        int flag = SETUP_SOLVE;
        // pull this problem's A into L2 now: the factorisation below gives the prefetch ~20k instructions of cover, so the
        // row blocks of the product loop are L2 hits instead of exposed HBM round trips
        if (lane == 0 && mA > 0) {
            const char* ab = reinterpret_cast<const char*>(A);
            const size_t abytes = ((size_t)mA * n * sizeof(T)) & ~(size_t)15;
            const size_t mis = (16 - (reinterpret_cast<size_t>(ab) & 15)) & 15; // bulk prefetch needs 16-byte alignment
            for (size_t off = mis; off + 16 <= abytes; off += 32768)
                bulk_prefetch_l2(ab + off, (unsigned)(((abytes - off) < 32768 ? (abytes - off) : 32768) & ~(size_t)15));
        }
        __syncwarp();
        // ... and the first row block into its staging buffer (nothing else touches it before the product loop)
        const unsigned stage = smem_u32(arow);
        if (mA > 0) async_block_copy(stage, reinterpret_cast<const char*>(A), (unsigned)(min(SETUP_RB, mA) * n * sizeof(T)), lane);

        // ---- sense copy + check_bounds (utils.c:84-98, 546-567)
        int any_fixed = 0, bad = 0, unsupported = 0, nsoft = 0;
        for (int i = lane; i < ldm; i += 32) {
            int s = 0;
            if (i < m) {
                s = sin ? sin[i] : 0;
                if ((s & B_BINARY) || (s & ~63)) unsupported = 1;
                if (s & B_SOFT) nsoft++;
                if (!(s & B_IMMUTABLE)) {
                    const T diff = bu[i] - bl[i];
                    if (diff < -st.primal_tol) bad = 1;
                    else if (diff < st.zero_tol && !(s & B_SOFT)) s |= B_ACTIVE + B_IMMUTABLE; // utils.c:560-563
                }
                if (s & (B_ACTIVE + B_IMMUTABLE)) any_fixed = 1;
            }
            so[i] = (unsigned char)s;
            if (a.sense_static) a.sense_static[(size_t)p * ldm + i] = (unsigned char)((i < m && sin) ? sin[i] : 0);
        }
        any_fixed = __any_sync(FULL, any_fixed);
        nsoft = __reduce_add_sync(FULL, nsoft);
        if (__any_sync(FULL, unsupported) || nsoft > a.ns_max) flag = EXIT_UNSUPPORTED; // ns_max sizes the factor storage
        // workspace mode: conflicting bounds do not stop the transform (new bounds may arrive with the next update); the
        // problem is flagged once its LDP is in place
        const bool bad_bounds = __any_sync(FULL, bad);
        if (flag >= 0 && bad_bounds && !a.no_shortcut) flag = EXIT_INFEASIBLE;

        // ---- Hessian factor (utils.c:223-391)
        bool is_diag = true;
        T hscale = 0;
        if (flag >= 0) {
            if (st.eps_prox > 0) flag = EXIT_UNSUPPORTED; // forced proximal mode is a different driver
            int nd = 0;
            for (int idx = lane; idx < n * n; idx += 32) {
                const int i = idx / n, j = idx - i * n;
                const T h = H[idx];
                if (j > i && (h > st.zero_tol || h < -st.zero_tol)) nd = 1;
                if (j == i) hscale = fmax(hscale, fabs(h));
                if (j >= i) R[roff(i, n) + j] = (j == i) ? h : (T)0.5 * (h + H[(size_t)j * n + i]);
            }
            is_diag = !__any_sync(FULL, nd);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) hscale = fmax(hscale, __shfl_xor_sync(FULL, hscale, o));
            __syncwarp();
        }
        if (flag >= 0 && is_diag) { // utils.c:284-312
            const T factor_tol = hscale > 0 ? st.zero_tol * hscale : st.zero_tol;
            int prox = 0, nonconvex = 0;
            for (int i = lane; i < n; i += 32) {
                T Hi = R[roff(i, n) + i];
                if (Hi <= factor_tol) {
                    prox = 1;
                    T eps = st.eps_prox < 0 ? -st.eps_prox : st.eps_prox;
                    const T fl = sqrt(st.zero_tol) * hscale;
                    if (eps > 0 && eps < fl) eps = fl;
                    Hi += eps;
                }
                if (Hi <= st.zero_tol) nonconvex = 1;
                Hi = sqrt(Hi);
                T* Ri = R + roff(i, n);
                for (int j = i + 1; j < n; j++) Ri[j] = 0;
                Ri[i] = 1 / Hi;
            }
            if (__any_sync(FULL, nonconvex)) flag = EXIT_NONCONVEX;
            else if (__any_sync(FULL, prox)) flag = EXIT_UNSUPPORTED; // reference hands over to daqp_prox
            __syncwarp();
        } else if (flag >= 0) {
            // upper Cholesky by rows, 1/r_ii kept on the diagonal (utils.c:337-352)
            T min_piv = (T)1e30, max_piv = 0;
            bool singular = false;
            for (int i = 0; i < n; i++) {
                T* Ri = R + roff(i, n);
                T part = 0;
                for (int kk = lane; kk < i; kk += 32) { const T t = R[roff(kk, n) + i]; part += t * t; }
                T di = Ri[i] - warp_sum(part);
                __syncwarp(); // every lane has read the pivot before lane 0 replaces it below
                if (di <= st.zero_tol) { singular = true; break; }
                min_piv = fmin(min_piv, di);
                max_piv = fmax(max_piv, di);
                di = rsqrt_exact<T>(di);
                for (int j = i + 1 + lane; j < n; j += 32) {
                    T s = Ri[j];
                    const T* Rk = R; // running row pointer: roff(kk+1) = roff(kk) + (n - kk - 1)
                    for (int kk = 0; kk < i; kk++) { s -= Rk[i] * Rk[j]; Rk += n - kk - 1; }
                    Ri[j] = s * di;
                }
                if (lane == 0) Ri[i] = di;
                __syncwarp();
            }
            if (singular || min_piv <= st.zero_tol * max_piv) { // utils.c:356-377: shift + proximal driver
                T eps = st.eps_prox < 0 ? -st.eps_prox : st.eps_prox;
                flag = (eps <= 0) ? EXIT_NONCONVEX : EXIT_UNSUPPORTED;
            } else {
                // R -> R^-1 in place (utils.c:380-389). Row k of the inverse needs only row k and the ORIGINAL rows
                // i > k, so all rows advance in lock-step over i: at step i rows k < i consume row i, then row i
                // starts its own inversion (and is never read again).
                for (int i = 0; i < n; i++) {
                    const T* Ri = R + roff(i, n);
                    const T rii = Ri[i];
                    for (int k0 = lane; k0 < i; k0 += 32) {
                        T* Rk = R + roff(k0, n);
                        const T t = Rk[i] * rii;
                        Rk[i] = t;
                        for (int j = i + 1; j < n; j++) Rk[j] -= Ri[j] * t;
                    }
                    __syncwarp();
                    for (int j = i + 1 + lane; j < n; j += 32) R[roff(i, n) + j] *= -rii;
                    __syncwarp();
                }
            }
        }

        if (flag < 0) { // setup failure: exit flag only, x untouched (api.c:69-72)
            if (lane == 0) { a.setup_flag[p] = flag; a.exitflag[p] = flag; a.iter[p] = 0; }
            cp_async_wait<0>(); // the staged row block must land before the next problem reuses the buffer
            __syncwarp();
            continue;
        }

        // ---- v = R^-T f (utils.c:474-497); zeros when f == NULL
        for (int i = lane; i < n; i += 32) {
            T s = 0;
            if (f) {
                s = R[roff(i, n) + i] * f[i];
                for (int j = i - 1; j >= 0; j--) s += R[roff(j, n) + i] * f[j];
            }
            vv[i] = s;
        }
        __syncwarp();
        T vnorm = 0;
        for (int i = lane; i < n; i += 32) vnorm += vv[i] * vv[i];
        vnorm = warp_sum(vnorm);
        if (a.v) for (int i = lane; i < n; i += 32) a.v[(size_t)p * n + i] = vv[i];

        // ---- unconstrained optimum x = -R^-1 v (utils.c:633-660), only when nothing is pre-activated/immutable
        const bool unc = !any_fixed && !a.no_shortcut;
        int infeasible_pt = 0;
        if (unc) {
            for (int i = lane; i < n; i += 32) {
                const T* Ri = R + roff(i, n);
                T s = 0;
                for (int j = i; j < n; j++) s += Ri[j] * vv[j];
                xu[i] = -s;
            }
            __syncwarp();
        }

        // ---- general rows: M = A R^-1, normalise, d (utils.c:434-472, 586-613, 663-678 / 499-544)
        int zero_row_infeasible = 0;
        // The row blocks of A stream through ONE staging buffer with cp.async: the copy of block k+1 is issued as soon as
        // the product loop of block k has read the buffer, and lands while block k's norms / d / stores are computed.
        for (int r0 = 0; r0 < mA; r0 += SETUP_RB) {
            const int nr = min(SETUP_RB, mA - r0);
            // the block's bounds, one row per lane, fetched before the product so that their latency is covered
            T bub = 0, blb = 0;
            if (lane < nr) { bub = bu[ms + r0 + lane]; blb = bl[ms + r0 + lane]; }
            cp_async_wait<0>();
            if (nr < SETUP_RB) // last, partial block: the missing rows read as zeros
                for (int idx = nr * n + lane; idx < SETUP_RB * n; idx += 32) arow[idx] = 0;
            __syncwarp();
            T acc[SETUP_RB][NGS];
#pragma unroll
            for (int b = 0; b < SETUP_RB; b++)
#pragma unroll
                for (int g = 0; g < NGS; g++) acc[b][g] = 0;
            // M[r][c] = sum_{i<=c} A[r][i] Rinv[i][c], accumulated from i = c down to 0 as the reference does
            {
                const T* Ri = R + roff(n - 1, n); // running row pointer: roff(i-1) = roff(i) - (n - i)
                const T* ap = arow + (n - 1); // column i of the staged rows: ap[b * n]
                int lim[NGS]; // column index, or -1 for lanes without a column in this segment
#pragma unroll
                for (int g = 0; g < NGS; g++) { const int c = lane + 32 * g; lim[g] = c < n ? c : -1; }
#pragma unroll 2
                for (int i = n - 1; i >= 0; i--) {
                    T ai[SETUP_RB];
#pragma unroll
                    for (int b = 0; b < SETUP_RB; b++) ai[b] = ap[b * n]; // broadcast reads
#pragma unroll
                    for (int g = 0; g < NGS; g++) {
                        if (lim[g] >= i) {
                            const T rv = Ri[lane + 32 * g];
#pragma unroll
                            for (int b = 0; b < SETUP_RB; b++) acc[b][g] += rv * ai[b];
                        }
                    }
                    Ri -= n - i;
                    ap -= 1;
                }
            }
            // Row norms, scaling and d for the RB rows of the block, arranged in three branch-free phases so that the
            // 2 RB warp reductions overlap instead of forming one long dependent chain (same operations per value).
            T nrmv[SETUP_RB], scal[SETUP_RB], dotv[SETUP_RB];
#pragma unroll
            for (int b = 0; b < SETUP_RB; b++) { // A_row . x_unc needs the staged rows: take it before the buffer is refilled
                T t = 0;
                if (unc) {
#pragma unroll
                    for (int g = 0; g < NGS; g++) { const int c = lane + 32 * g; if (c < n) t += arow[b * n + c] * xu[c]; }
                }
                dotv[b] = t;
            }
            __syncwarp(); // every lane is done with the staged rows
            if (r0 + SETUP_RB < mA)
                async_block_copy(stage, reinterpret_cast<const char*>(A + (size_t)(r0 + SETUP_RB) * n),
                                 (unsigned)(min(SETUP_RB, mA - r0 - SETUP_RB) * n * sizeof(T)), lane);
#pragma unroll
            for (int b = 0; b < SETUP_RB; b++) {
                T t = 0;
#pragma unroll
                for (int g = 0; g < NGS; g++) t += acc[b][g] * acc[b][g];
                nrmv[b] = t;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int b = 0; b < SETUP_RB; b++) nrmv[b] += __shfl_xor_sync(FULL, nrmv[b], o);
#pragma unroll
            for (int b = 0; b < SETUP_RB; b++) {
                const bool scale = b < nr && !(nrmv[b] < st.zero_tol); // utils.c:595-606: zero rows are not scaled
                const T sv = rsqrt_exact<T>(scale ? nrmv[b] : (T)1);   // exactly 1 for unscaled rows
                scal[b] = sv;
                T t = dotv[b];
#pragma unroll
                for (int g = 0; g < NGS; g++) {
                    acc[b][g] *= sv;
                    const int c = lane + 32 * g;
                    if (!unc && c < n) t += acc[b][g] * vv[c];
                }
                dotv[b] = t;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int b = 0; b < SETUP_RB; b++) dotv[b] += __shfl_xor_sync(FULL, dotv[b], o);
            {
                // lane b finishes row b (its bounds are already in bub / blb): four coalesced stores for the block
                T s_my = 1, d_my = 0, n_my = 1;
#pragma unroll
                for (int b = 0; b < SETUP_RB; b++) if (lane == b) { s_my = scal[b]; d_my = dotv[b]; n_my = nrmv[b]; }
                if (lane < nr) {
                    const int row = ms + r0 + lane; // constraint index
                    int sb = so[row];
                    if (n_my < st.zero_tol) {
                        if ((bub < -st.zero_tol || blb > st.zero_tol) && !(sb & B_IMMUTABLE) && !(sb & B_SOFT)) zero_row_infeasible = 1;
                        sb = B_IMMUTABLE;
                    }
                    T u_ = bub, l_ = blb;
                    if (unc) {
                        u_ -= d_my; l_ -= d_my;
                        if (u_ < -st.primal_tol || l_ > st.primal_tol) infeasible_pt = 1;
                        u_ *= s_my; l_ *= s_my;
                    } else {
                        u_ = u_ * s_my + d_my; l_ = l_ * s_my + d_my;
                    }
                    du[row] = u_; dl[row] = l_; sc[row] = s_my; so[row] = (unsigned char)sb;
                    if (a.sense_static && n_my < st.zero_tol) a.sense_static[(size_t)p * ldm + row] = (unsigned char)B_IMMUTABLE;
                }
            }
            // row-major copy (coalesced), column-major copy (one 32-byte sector per lane and column)
#pragma unroll
            for (int g = 0; g < NGS; g++) {
                const int c = lane + 32 * g;
                if (c < ldn) {
#pragma unroll
                    for (int b = 0; b < SETUP_RB; b++)
                        if (b < nr) Mr[(size_t)(ms + r0 + b) * ldn + c] = (c < n) ? acc[b][g] : (T)0;
                }
                if (c < n) {
                    T* dst = Mt + (size_t)c * ldm + ms + r0;
#pragma unroll
                    for (int b = 0; b < SETUP_RB; b++)
                        if (ms + r0 + b < ldm) dst[b] = (b < nr) ? acc[b][g] : (T)0;
                }
                if (Mt32 && c < ((n + 3) & ~3)) { // columns n..4 ceil(n/4)-1 are zero padding
                    float* d32 = Mt32 + ((size_t)(c >> 2) * m + ms + r0) * 4 + (c & 3);
#pragma unroll
                    for (int b = 0; b < SETUP_RB; b++)
                        if (b < nr) d32[4 * b] = (c < n) ? (float)acc[b][g] : 0.f;
                }
            }
            __syncwarp();
        }

        // ---- simple bounds: rows 0..ms-1 of R^-1, normalised (utils.c:569-585); identity rows for diagonal H
        for (int i = 0; i < ms; i++) {
            T* Ri = R + roff(i, n);
            T s;
            if (is_diag) { // utils.c:308-309: scaling = sqrt(H_ii), row = e_i
                s = sqrt(H[(size_t)i * n + i]);
                __syncwarp();
                if (lane == 0) Ri[i] = 1;
            } else {
                T part = 0;
                for (int j = i + lane; j < n; j += 32) part += Ri[j] * Ri[j];
                s = rsqrt_exact<T>(warp_sum(part));
                for (int j = i + lane; j < n; j += 32) Ri[j] *= s;
            }
            __syncwarp();
            T dotd = 0;
            if (!unc) { for (int j = i + lane; j < n; j += 32) dotd += Ri[j] * vv[j]; dotd = warp_sum(dotd); }
            for (int c = lane; c < max(ldn, (n + 3) & ~3); c += 32) {
                const T val = (c >= i && c < n) ? Ri[c] : (T)0;
                if (c < ldn) Mr[(size_t)i * ldn + c] = val;
                if (c < n) Mt[(size_t)c * ldm + i] = val;
                if (Mt32 && c < ((n + 3) & ~3)) Mt32[((size_t)(c >> 2) * m + i) * 4 + (c & 3)] = (float)val;
            }
            if (lane == 0) {
                T u_ = bu[i], l_ = bl[i];
                if (unc) {
                    u_ -= xu[i]; l_ -= xu[i];
                    if (u_ < -st.primal_tol || l_ > st.primal_tol) infeasible_pt = 1;
                    u_ *= s; l_ *= s;
                } else {
                    u_ = u_ * s + dotd; l_ = l_ * s + dotd;
                }
                du[i] = u_; dl[i] = l_; sc[i] = s;
            }
        }
        // pad rows m..ldm-1 of the column-major copy and of the per-row vectors
        for (int c = lane; c < n; c += 32)
            for (int r = m; r < ldm; r++) Mt[(size_t)c * ldm + r] = 0;
        for (int r = m + lane; r < ldm; r += 32) { du[r] = 0; dl[r] = 0; sc[r] = 1; }
        __syncwarp();
        for (int idx = lane; idx < ntri; idx += 32) Rg[idx] = R[idx];

        infeasible_pt = __any_sync(FULL, infeasible_pt);
        zero_row_infeasible = __any_sync(FULL, zero_row_infeasible);
        if (unc && !infeasible_pt) { // utils.c:679-683 + api.c:40-45,455-495: solve is skipped
            for (int i = lane; i < n; i += 32) a.x[(size_t)p * n + i] = xu[i];
            if (a.lam) for (int i = lane; i < m; i += 32) a.lam[(size_t)p * m + i] = 0;
            if (lane == 0) {
                if (f) a.fval[p] = (T)-0.5 * vnorm;
                if (a.soft_slack) a.soft_slack[p] = 0;
                a.exitflag[p] = EXIT_OPTIMAL;
                a.iter[p] = 1;
                a.setup_flag[p] = SETUP_UNCONSTRAINED;
            }
        } else if (zero_row_infeasible || bad_bounds) {
            if (lane == 0) { a.setup_flag[p] = EXIT_INFEASIBLE; a.exitflag[p] = EXIT_INFEASIBLE; a.iter[p] = 0; }
        } else {
            if (lane == 0) a.setup_flag[p] = (sin != nullptr || any_fixed) ? SETUP_SOLVE_ACTIVATE : SETUP_SOLVE;
        }
        __syncwarp();
    }
}

} // namespace dq
