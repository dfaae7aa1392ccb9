// daqp_b200/csrc/setup2_kernel.cuh -- the QP -> LDP transform as TWO kernels (fp64, n <= 127).
//
//   qp_factor_kernel   check_bounds, symmetrise + Cholesky + triangular inverse, v = R^-T f, unconstrained optimum
//                      (reference src/utils.c:84-98,223-391,474-497,546-567,633-660): dependent chains, shared-memory
//                      resident packed triangle -- one warp per problem (n <= 64) or one CTA of four warps (n > 64).
//   qp_product_kernel  M = A R^-1 (utils.c:434-472) on the fp64 TENSOR CORES (mma.sync m8n8k4: the one dense contraction
//                      of the path), fused with the row normalisation (utils.c:586-613), d (utils.c:151-158 / 499-544),
//                      the normalised simple-bound rows (utils.c:569-585) and the stores of the three layouts the solve
//                      kernel reads (row-major fp64, column-major fp64, fp32 quad copy). One CTA of four warps per
//                      problem: R^-1 sits in shared memory as ready-made B fragments (upper triangular: the k-steps
//                      beyond a column tile are skipped), every warp takes 8-row blocks of A straight from global memory
//                      as A fragments; no staging, 56 DMMA instead of 800 DFMA + 500 LDS per block at n = 50.
//
// The single fused kernel of setup_kernel.cuh remains for fp32 and n > 127; it was 16 % of a C3 step and 5x above its
// HBM floor, two thirds of it in the scalar product loop and in the lane-serial factorisation.
// What may differ from the fused kernel: the ORDER of the sums inside M = A R^-1 (k in groups of four, ascending) --
// a reduction order, which the reference itself leaves to the compiler (-fassociative-math).
#pragma once
#include "setup_kernel.cuh"

namespace dq {

constexpr int FACTOR_MAX_WARPS = 20; // problems per CTA of the factor kernel (one warp each): 20 x 32 threads x 102 registers = the register file
constexpr int SI_UNC = 1, SI_DIAG = 2, SI_FIXED = 4, SI_BADBOUNDS = 8, SI_HAS_SENSE = 16;

template <typename T>
__host__ __device__ inline size_t factor_smem_per_team(int n) {
    size_t e = (size_t)n * (n + 1) / 2 + 2 * (size_t)n;
    return (e * sizeof(T) + 15) / 16 * 16 + 16;
}

// D = A (8x4, row) * B (4x8, col) + D on the fp64 tensor cores. Lane l holds A[l/4][l%4], B[l%4][l/4],
// D[l/4][2(l%4)], D[l/4][2(l%4)+1].
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// ---- kernel 1: Hessian factor --------------------------------------------------------------------------------------
template <typename T, int TW>
__global__ void __launch_bounds__(TW > 1 ? 32 * TW : 32 * FACTOR_MAX_WARPS, 1) qp_factor_kernel(const SetupArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NT = 32 * TW;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int tid = TW > 1 ? (int)threadIdx.x : lane;
    const int n = a.n, m = a.m, ldm = a.ldm;
    const int ntri = n * (n + 1) / 2;
    unsigned char* base = smem_raw + (TW > 1 ? 0 : factor_smem_per_team<T>(n) * wib);
    int* sh = reinterpret_cast<int*>(base); // [0] problem index (team), [1] soft count (team)
    T* R = reinterpret_cast<T*>(base + 16);
    T* vv = R + ntri; // f, then v = R^-T f
    T* xu = vv + n;   // unconstrained optimum
    const DevSettings<T>& st = a.st;
    auto sync = [] { if constexpr (TW > 1) __syncthreads(); else __syncwarp(); };
    auto any = [](int v) -> bool { if constexpr (TW > 1) return __syncthreads_or(v) != 0; else return __any_sync(FULL, v) != 0; };

    for (;;) {
        int p = 0;
        if constexpr (TW > 1) {
            if (tid == 0) { sh[0] = atomicAdd(a.work_counter, 1); sh[1] = 0; }
            __syncthreads();
            p = sh[0];
        } else {
            if (lane == 0) p = atomicAdd(a.work_counter, 1);
            p = __shfl_sync(FULL, p, 0);
        }
        if (p >= a.P) break;

        const T* H = a.H + (size_t)p * n * n;
        const T* f = a.f ? a.f + (size_t)p * n : nullptr;
        const T* bu = a.bupper + (size_t)p * m;
        const T* bl = a.blower + (size_t)p * m;
        const int* sin = a.sense_in ? a.sense_in + (size_t)p * m : nullptr;
        unsigned char* so = a.sense + (size_t)p * ldm;
        T* Rg = a.Rinv + (size_t)p * ntri;
        int flag = SETUP_SOLVE;

        // ---- sense copy + check_bounds (utils.c:84-98, 546-567)
        int any_fixed = 0, bad = 0, unsupported = 0, nsoft = 0;
        for (int i = tid; i < ldm; i += NT) {
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
        if constexpr (TW > 1) { if (nsoft) atomicAdd(&sh[1], nsoft); }
        any_fixed = any(any_fixed);
        if constexpr (TW > 1) nsoft = sh[1]; else nsoft = __reduce_add_sync(FULL, nsoft);
        if (any(unsupported) || nsoft > a.ns_max) flag = EXIT_UNSUPPORTED; // ns_max sizes the factor storage
        const bool bad_bounds = any(bad);
        if (flag >= 0 && bad_bounds && !a.no_shortcut) flag = EXIT_INFEASIBLE;

        // ---- Hessian factor (utils.c:223-391)
        bool is_diag = true;
        T hscale = 0;
        if (flag >= 0) {
            if (st.eps_prox > 0) flag = EXIT_UNSUPPORTED; // forced proximal mode is a different driver
            // (a two-pass variant that reads H coalesced both times and averages the lower elements into their mirror
            // images was slower at both sizes: 7.1 -> 7.4 ms at C3, 60 -> 64 ms at C4 -- the strided read overlaps)
            int nd = 0;
            for (int idx = tid; idx < n * n; idx += NT) {
                const int i = idx / n, j = idx - i * n;
                const T h = H[idx];
                if (j > i && (h > st.zero_tol || h < -st.zero_tol)) nd = 1;
                if (j >= i) R[roff(i, n) + j] = (j == i) ? h : (T)0.5 * (h + H[(size_t)j * n + i]);
            }
            is_diag = !any(nd);
            sync(); // the writes of R are ordered before the reads below (a vote alone is not a memory barrier)
            for (int i = lane; i < n; i += 32) hscale = fmax(hscale, fabs(R[roff(i, n) + i])); // every warp: the whole diagonal
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) hscale = fmax(hscale, __shfl_xor_sync(FULL, hscale, o));
            sync();
        }
        if (flag >= 0 && is_diag) { // utils.c:284-312
            const T factor_tol = hscale > 0 ? st.zero_tol * hscale : st.zero_tol;
            int prox = 0, nonconvex = 0;
            for (int i = tid; i < n; i += NT) {
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
            if (any(nonconvex)) flag = EXIT_NONCONVEX;
            else if (any(prox)) flag = EXIT_UNSUPPORTED; // reference hands over to daqp_prox
            sync();
        } else if (flag >= 0) {
            // upper Cholesky by rows, 1/r_ii kept on the diagonal (utils.c:337-352). Rows are taken CB at a time: the part
            // of their sums that runs over the rows above the block (kk < i0) is accumulated for all CB rows in one
            // pass -- one load of R[kk][j] serves CB products -- then the rows are finished one after the other with
            // the terms of the block's own rows. Per element the products and their order (kk ascending) are unchanged.
            T min_piv = (T)1e30, max_piv = 0;
            bool singular = false;
            constexpr int CB = 4, JC = TW > 1 ? 1 : 2; // thread tid owns columns i0 + 1 + tid + NT c (n - 1 <= NT JC)
            for (int i0 = 0; i0 < n && !singular; i0 += CB) {
                T s[JC][CB];
#pragma unroll
                for (int c = 0; c < JC; c++) {
                    const int j = i0 + 1 + tid + NT * c;
#pragma unroll
                    for (int r = 0; r < CB; r++) s[c][r] = (j < n && j > i0 + r) ? R[roff(i0 + r, n) + j] : (T)0;
                }
                {
                    const T* Rk = R; // running row pointer: roff(kk+1) = roff(kk) + (n - kk - 1)
                    for (int kk = 0; kk < i0; kk++) {
                        T b[CB];
#pragma unroll
                        for (int r = 0; r < CB; r++) b[r] = Rk[min(i0 + r, n - 1)];
#pragma unroll
                        for (int c = 0; c < JC; c++) {
                            if (c == 0 || i0 + 1 + NT * c < n) { // (uniform)
                                const int j = i0 + 1 + tid + NT * c;
                                const T x = Rk[min(j, n - 1)];
#pragma unroll
                                for (int r = 0; r < CB; r++) s[c][r] -= b[r] * x;
                            }
                        }
                        Rk += n - kk - 1;
                    }
                }
#pragma unroll
                for (int r = 0; r < CB; r++) {
                    const int i = i0 + r;
                    if (i < n && !singular) { // (uniform)
                        T* Ri = R + roff(i, n);
                        T part = 0; // every warp sums the whole column: no cross-warp reduction
                        for (int kk = lane; kk < i; kk += 32) { const T t = R[roff(kk, n) + i]; part += t * t; }
                        T di = Ri[i] - warp_sum(part);
                        sync(); // every thread has read the pivot before it is replaced below
                        if (di <= st.zero_tol) singular = true;
                        else {
                            min_piv = fmin(min_piv, di);
                            max_piv = fmax(max_piv, di);
                            di = rsqrt_exact<T>(di);
#pragma unroll
                            for (int c = 0; c < JC; c++) {
                                const int j = i0 + 1 + tid + NT * c;
                                if (j < n && j > i) {
                                    T v = s[c][r];
                                    for (int kk = i0; kk < i; kk++) v -= R[roff(kk, n) + i] * R[roff(kk, n) + j];
                                    Ri[j] = v * di;
                                }
                            }
                            if (tid == 0) Ri[i] = di;
                        }
                        sync();
                    }
                }
            }
            if (singular || min_piv <= st.zero_tol * max_piv) { // utils.c:356-377: shift + proximal driver
                T eps = st.eps_prox < 0 ? -st.eps_prox : st.eps_prox;
                flag = (eps <= 0) ? EXIT_NONCONVEX : EXIT_UNSUPPORTED;
            } else {
                // R -> R^-1 in place (utils.c:380-389): all rows advance in lock-step over the pivot i. The update of pivot
                // i is the outer product (rows k0 < i) x (columns j > i): tall and narrow late, short and wide early. Early
                // pivots give a WARP a row and its lanes the columns (row i cached in registers), late ones a THREAD a row
                // (the round-1 mapping); the switch is per pivot by instruction count. Same products, same order.
                const int wrp = TW > 1 ? wib : 0;
                for (int i = 0; i < n; i++) {
                    const T* Ri = R + roff(i, n);
                    const T rii = Ri[i];
                    const int w = n - i - 1;
                    const int cost_row = ((i + NT - 1) / NT) * w * 4, cost_col = ((i + TW - 1) / TW) * (((w + 31) >> 5) * 3 + 8);
                    if (cost_row <= cost_col) {
                        for (int k0 = tid; k0 < i; k0 += NT) {
                            T* Rk = R + roff(k0, n);
                            const T t = Rk[i] * rii;
                            Rk[i] = t;
                            for (int j = i + 1; j < n; j++) Rk[j] -= Ri[j] * t;
                        }
                    } else {
                        T rj[4];
#pragma unroll
                        for (int c = 0; c < 4; c++) { const int j = i + 1 + lane + 32 * c; rj[c] = j < n ? Ri[j] : (T)0; }
                        for (int k0 = wrp; k0 < i; k0 += TW) {
                            T* Rk = R + roff(k0, n);
                            const T t = Rk[i] * rii;
                            __syncwarp(); // every lane has read the row's pivot entry before lane 0 replaces it
                            if (lane == 0) Rk[i] = t;
#pragma unroll
                            for (int c = 0; c < 4; c++) {
                                const int j = i + 1 + lane + 32 * c;
                                if (32 * c < w && j < n) Rk[j] -= rj[c] * t;
                            }
                        }
                    }
                    sync();
                    for (int j = i + 1 + tid; j < n; j += NT) R[roff(i, n) + j] *= -rii;
                    sync();
                }
            }
        }

        if (flag < 0) { // setup failure: exit flag only, x untouched (api.c:69-72)
            if (tid == 0) { a.setup_flag[p] = flag; a.exitflag[p] = flag; a.iter[p] = 0; a.info[4 * (size_t)p] = flag; }
            sync();
            continue;
        }

        // ---- v = R^-T f (utils.c:474-497); zeros when f == NULL
        for (int i = tid; i < n; i += NT) {
            T s = 0;
            if (f) {
                s = R[roff(i, n) + i] * f[i];
                for (int j = i - 1; j >= 0; j--) s += R[roff(j, n) + i] * f[j];
            }
            vv[i] = s;
        }
        sync();
        T vnorm = 0;
        for (int i = lane; i < n; i += 32) vnorm += vv[i] * vv[i];
        vnorm = warp_sum(vnorm);
        for (int i = tid; i < n; i += NT) a.v[(size_t)p * n + i] = vv[i];

        // ---- unconstrained optimum x = -R^-1 v (utils.c:633-660), only when nothing is pre-activated/immutable
        const bool unc = !any_fixed && !a.no_shortcut;
        if (unc) {
            for (int i = tid; i < n; i += NT) {
                const T* Ri = R + roff(i, n);
                T s = 0;
                for (int j = i; j < n; j++) s += Ri[j] * vv[j];
                a.xu[(size_t)p * n + i] = -s;
            }
        }
        for (int idx = tid; idx < ntri; idx += NT) Rg[idx] = R[idx];
        if (tid == 0) {
            a.info[4 * (size_t)p] = SETUP_SOLVE;
            a.info[4 * (size_t)p + 1] = (unc ? SI_UNC : 0) | (is_diag ? SI_DIAG : 0) | (any_fixed ? SI_FIXED : 0) |
                                        (bad_bounds ? SI_BADBOUNDS : 0) | (sin ? SI_HAS_SENSE : 0);
            a.vnorm[p] = vnorm;
        }
        sync();
    }
}

// ---- kernel 2: M = A R^-1 on the tensor cores + everything per constraint row -------------------------------------
inline size_t product_smem(int n) {
    const int S = (n + 3) >> 2, nct = (n + 8) >> 3;
    size_t nfrag = 0;
    for (int t = 0; t < nct; t++) nfrag += (size_t)std::min(S, 2 * t + 2);
    return nfrag * 256 + 2 * (size_t)(8 * nct) * sizeof(double) + 32 + (nfrag * 4 + 15) / 16 * 16;
}

template <int NCT>
__global__ void __launch_bounds__(128, NCT <= 9 ? 4 : 3) qp_product_kernel(const SetupArgs<double> a) {
    typedef double T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, tid = threadIdx.x;
    const int n = a.n, m = a.m, ms = a.ms, mA = a.m - a.ms, ldm = a.ldm, ldn = a.ldn;
    const int S = (n + 3) >> 2;   // k-steps of four
    const int nct = (n + 8) >> 3; // column tiles of eight, column n (the unconstrained optimum) included
    int nfrag = 0;
    for (int t = 0; t < nct; t++) nfrag += min(S, 2 * t + 2);
    T* Bf = reinterpret_cast<T*>(smem_raw);          // [nfrag][32] B fragments of R^-1 (column n: x_unc)
    T* vs = Bf + (size_t)nfrag * 32;                  // v, zero padded to 8 nct
    T* xus = vs + 8 * nct;                            // x_unc, zero padded
    int* sh = reinterpret_cast<int*>(xus + 8 * nct);  // [0] problem index
    short2* fts = reinterpret_cast<short2*>(sh + 8);  // [nfrag] (column tile, k-step) of every B fragment
    const DevSettings<T>& st = a.st;
    const int g4 = lane >> 2, q4 = lane & 3;          // row within a block, column pair within a tile
    const int tx = n >> 3, jx = n & 7;                // tile / column of the x_unc column

    {
        int fb = 0;
        for (int t = 0; t < nct; t++) {
            const int ns_t = min(S, 2 * t + 2);
            for (int s = tid; s < ns_t; s += 128) fts[fb + s] = make_short2((short)t, (short)s);
            fb += ns_t;
        }
    }
    for (;;) {
        __syncthreads(); // the previous problem's fragments are no longer read
        if (tid == 0) sh[0] = atomicAdd(a.work_counter + 1, 1);
        __syncthreads();
        const int p = sh[0];
        if (p >= a.P) break;
        if (a.info[4 * (size_t)p] < 0) continue; // finished by the factor kernel
        const int bits = a.info[4 * (size_t)p + 1];
        const bool unc = bits & SI_UNC, is_diag = bits & SI_DIAG, any_fixed = bits & SI_FIXED, bad_bounds = bits & SI_BADBOUNDS;

        const T* H = a.H + (size_t)p * n * n;
        const T* A = a.A + (size_t)p * mA * n;
        const T* bu = a.bupper + (size_t)p * m;
        const T* bl = a.blower + (size_t)p * m;
        T* Mt = a.Mt + (size_t)p * n * ldm;
        float* Mt32 = a.Mt32 ? a.Mt32 + (size_t)p * ((n + 3) / 4) * m * 4 : nullptr;
        T* Mr = a.Mr + (size_t)p * m * ldn;
        T* du = a.dupper + (size_t)p * ldm;
        T* dl = a.dlower + (size_t)p * ldm;
        T* sc = a.scaling + (size_t)p * ldm;
        unsigned char* so = a.sense + (size_t)p * ldm;
        T* Rg = a.Rinv + (size_t)p * (n * (n + 1) / 2);
        const T* vg = a.v + (size_t)p * n;
        const T* xug = a.xu + (size_t)p * n;

        for (int i = tid; i < 8 * nct; i += 128) { vs[i] = i < n ? vg[i] : (T)0; xus[i] = (unc && i < n) ? xug[i] : (T)0; }
        { // B fragments: element (k, c) of [R^-1 | x_unc], k = 4 s + lane % 4, c = 8 t + lane / 4; zero below the diagonal.
          // Four fragments per warp and trip: four independent gathers in flight per thread (one at a time left the
          // global-load latency of this phase exposed: 19 % of the kernel's stall samples).
            for (int f0 = wid; f0 < nfrag; f0 += 16) {
                T val[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int f = f0 + 4 * q;
                    val[q] = 0;
                    if (f < nfrag) {
                        const short2 ts = fts[f];
                        const int k = 4 * ts.y + q4, c = 8 * ts.x + g4;
                        if (k < n) {
                            if (c < n) { if (k <= c) val[q] = Rg[roff(k, n) + c]; }
                            else if (c == n && unc) val[q] = xug[k];
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int f = f0 + 4 * q;
                    if (f < nfrag) Bf[(size_t)f * 32 + lane] = val[q];
                }
            }
        }
        __syncthreads();

        const int nbg = (mA + 7) >> 3, nbs = (ms + 7) >> 3;
        int infeasible_pt = 0, zero_row_infeasible = 0;
        for (int blk = wid; blk < nbg + nbs; blk += 4) {
            const bool simple = blk >= nbg;
            const int row = 8 * (simple ? blk - nbg : blk) + g4; // row within its group (simple bounds / general rows)
            const bool rvalid = row < (simple ? ms : mA);
            const int ci = simple ? row : ms + row;               // constraint index
            // (bounds and sense of the row, asked for before the product so that the epilogue does not wait for them)
            const bool owner = q4 == 0 && rvalid;
            const T bub = owner ? bu[ci] : (T)0, blb = owner ? bl[ci] : (T)0;
            int sb = owner ? so[ci] : 0;
            T C[NCT][2];
#pragma unroll
            for (int t = 0; t < NCT; t++) { C[t][0] = 0; C[t][1] = 0; }
            if (!simple) {
                T af[2 * NCT];
                const T* Arow = A + (size_t)min(row, mA - 1) * n + q4;
#pragma unroll
                for (int s = 0; s < 2 * NCT; s++) af[s] = (rvalid && 4 * s + q4 < n) ? __ldg(Arow + 4 * s) : (T)0;
                int fb = 0;
#pragma unroll
                for (int t = 0; t < NCT; t++) {
                    if (t < nct) {
                        const int ns_t = min(S, 2 * t + 2);
#pragma unroll
                        for (int s = 0; s < 2 * t + 2; s++)
                            if (s < ns_t) dmma884(C[t], af[s], Bf[(size_t)(fb + s) * 32 + lane]);
                        fb += ns_t;
                    }
                }
            } else if (rvalid) { // row `row` of R^-1 (not yet normalised); the unit vector when H is diagonal
                const T* Ri = Rg + roff(row, n);
#pragma unroll
                for (int t = 0; t < NCT; t++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int c = 8 * t + 2 * q4 + e;
                        if (c < n && c >= row) C[t][e] = is_diag ? (c == row ? (T)1 : (T)0) : Ri[c];
                    }
                }
            }
            // A_row . x_unc sits in column n: take it out of the tile (it is not part of the row of M)
            T dotx = 0;
#pragma unroll
            for (int t = 0; t < NCT; t++) {
                if (t == tx) {
                    const T mine = (jx & 1) ? C[t][1] : C[t][0];
                    dotx = __shfl_sync(FULL, mine, (lane & ~3) | (jx >> 1));
                    if (q4 == (jx >> 1)) { if (jx & 1) C[t][1] = 0; else C[t][0] = 0; }
                }
            }
            if (simple) dotx = xus[min(row, n - 1)]; // a simple bound's row is e_i in x-space (utils.c:663-678)
            // row norm, scaling (utils.c:586-613 / 569-585)
            T nrm = 0;
#pragma unroll
            for (int t = 0; t < NCT; t++) nrm += C[t][0] * C[t][0] + C[t][1] * C[t][1];
            nrm += __shfl_xor_sync(FULL, nrm, 1);
            nrm += __shfl_xor_sync(FULL, nrm, 2);
            T srow, sbnd; // factor applied to the row / to the bounds
            if (simple && is_diag) { srow = 1; sbnd = sqrt(H[(size_t)min(row, n - 1) * n + min(row, n - 1)]); } // utils.c:308-309
            else {
                const bool scale = rvalid && !(nrm < st.zero_tol) ; // utils.c:595-606: zero rows are not scaled
                srow = sbnd = rsqrt_exact<T>(scale ? nrm : (T)1);
            }
            T dotv = 0; // M_row . v with the normalised row (utils.c:499-544)
#pragma unroll
            for (int t = 0; t < NCT; t++) {
                C[t][0] *= srow; C[t][1] *= srow;
                if (!unc && t < nct) {
                    const double2 v2 = *reinterpret_cast<const double2*>(vs + 8 * t + 2 * q4);
                    dotv += C[t][0] * v2.x; dotv += C[t][1] * v2.y;
                }
            }
            dotv += __shfl_xor_sync(FULL, dotv, 1);
            dotv += __shfl_xor_sync(FULL, dotv, 2);
            if (owner) { // one lane per row: bounds, scaling, sense
                if (!simple && nrm < st.zero_tol) {
                    if ((bub < -st.zero_tol || blb > st.zero_tol) && !(sb & B_IMMUTABLE) && !(sb & B_SOFT)) zero_row_infeasible = 1;
                    sb = B_IMMUTABLE;
                    if (a.sense_static) a.sense_static[(size_t)p * ldm + ci] = (unsigned char)B_IMMUTABLE;
                }
                T u_ = bub, l_ = blb;
                if (unc) {
                    u_ -= dotx; l_ -= dotx;
                    if (u_ < -st.primal_tol || l_ > st.primal_tol) infeasible_pt = 1;
                    u_ *= sbnd; l_ *= sbnd;
                } else {
                    u_ = u_ * sbnd + dotv; l_ = l_ * sbnd + dotv;
                }
                du[ci] = u_; dl[ci] = l_; sc[ci] = sbnd; so[ci] = (unsigned char)sb;
            }
            if (rvalid) { // the three layouts of the row (zeros beyond column n-1 come out of the zero fragments)
#pragma unroll
                for (int t = 0; t < NCT; t++) {
                    const int c = 8 * t + 2 * q4;
                    if (c < ldn) *reinterpret_cast<double2*>(Mr + (size_t)ci * ldn + c) = make_double2(C[t][0], C[t][1]);
                    if (c < n) Mt[(size_t)c * ldm + ci] = C[t][0];
                    if (c + 1 < n) Mt[(size_t)(c + 1) * ldm + ci] = C[t][1];
                    if (Mt32 && c < ((n + 3) & ~3))
                        *reinterpret_cast<float2*>(Mt32 + ((size_t)(c >> 2) * m + ci) * 4 + (c & 3)) = make_float2((float)C[t][0], (float)C[t][1]);
                    if (simple) { // the first ms rows of R^-1 stay normalised for x = R^-1 (u - v) (daqp.c:120-131)
                        if (c < n && c >= row) Rg[roff(row, n) + c] = C[t][0];
                        if (c + 1 < n && c + 1 >= row) Rg[roff(row, n) + c + 1] = C[t][1];
                    }
                }
            }
        }
        // pad rows m..ldm-1 of the column-major copy and of the per-row vectors
        for (int c = tid; c < n; c += 128)
            for (int r = m; r < ldm; r++) Mt[(size_t)c * ldm + r] = 0;
        for (int r = m + tid; r < ldm; r += 128) { du[r] = 0; dl[r] = 0; sc[r] = 1; }

        infeasible_pt = __syncthreads_or(infeasible_pt);
        zero_row_infeasible = __syncthreads_or(zero_row_infeasible);
        if (unc && !infeasible_pt) { // utils.c:679-683 + api.c:40-45,455-495: solve is skipped
            for (int i = tid; i < n; i += 128) a.x[(size_t)p * n + i] = xus[i];
            if (a.lam) for (int i = tid; i < m; i += 128) a.lam[(size_t)p * m + i] = 0;
            if (tid == 0) {
                if (a.f) a.fval[p] = (T)-0.5 * a.vnorm[p];
                if (a.soft_slack) a.soft_slack[p] = 0;
                a.exitflag[p] = EXIT_OPTIMAL;
                a.iter[p] = 1;
                a.setup_flag[p] = SETUP_UNCONSTRAINED;
            }
        } else if (zero_row_infeasible || bad_bounds) {
            if (tid == 0) { a.setup_flag[p] = EXIT_INFEASIBLE; a.exitflag[p] = EXIT_INFEASIBLE; a.iter[p] = 0; }
        } else {
            if (tid == 0) a.setup_flag[p] = ((bits & SI_HAS_SENSE) || any_fixed) ? SETUP_SOLVE_ACTIVATE : SETUP_SOLVE;
        }
    }
}

} // namespace dq
