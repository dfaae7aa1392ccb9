// daqp_b200/csrc/ldp_kernel.cuh -- the hot path: batched dual active-set LDP solve, one warp per problem.
//
// Covers SURVEY.md §8(a) rows a2..a16: LDL' rank-1 add / remove, CSP triangular solves, dual ratio test, primal
// update, feasibility scan, singular direction, pivoting, warm-start activation, refinement, the daqp_ldp state
// machine, and the LDP->QP back-transform + result extraction.
//
// Memory plan per problem
//   shared (warp-private, lives for the whole solve): packed L, D, lam, lam*, xldl, zldl, u, WS, sense bits
//   global, streamed every feasibility scan : Mt  = constraint matrix, column-major [n][ldm]  (128-bit coalesced)
//   global, gathered per add / primal update: Mr  = same matrix, row-major [m][ldn]           (128-bit coalesced rows)
//   global, once                            : dupper, dlower, scaling, Rinv (packed), v
// Both matrix copies are written by the setup kernel; rows 0..ms-1 are the normalised rows of R^-1 (simple bounds,
// zero-filled below the diagonal), rows ms..m-1 the normalised rows of A R^-1.
#pragma once
#include "common.cuh"
#include <limits.h>

namespace dq {

template <typename T>
struct LdpArgs {
    int P, n, m, ms, ldm, ldn, cap;
    const T* Mt;              // [P][n][ldm]
    const T* Mr;              // [P][m][ldn]
    const T* dupper;          // [P][ldm]
    const T* dlower;          // [P][ldm]
    const T* scaling;         // [P][ldm]
    const T* Rinv;            // [P][n(n+1)/2]
    const T* v;               // [P][n] or nullptr (f == NULL)
    const unsigned char* sense; // [P][ldm]   bits after setup (check_bounds / zero rows applied)
    const int* setup_flag;    // [P]
    T* x;                     // [P][n]
    T* lam;                   // [P][m] or nullptr
    T* fval;                  // [P]
    int* exitflag;            // [P]
    int* iter;                // [P]
    int* ws_out;              // [P][cap] or nullptr : final working set (factor order)
    int* nact_out;            // [P] or nullptr
    int* counts_out;          // [P][4] or nullptr  : scans, adds, removes, csp solves
    unsigned char* sense_out; // [P][ldm] or nullptr: final sense bits
    int* work_counter;        // dynamic problem queue
    int* pst_id;              // [total warps][cap] pivot stack (rare path)
    T* pst_lam;               // [total warps][cap]
    DevSettings<T> st;
};

// bytes of shared memory one warp needs
template <typename T>
__host__ __device__ inline size_t ldp_smem_per_warp(int n, int m, int cap) {
    const int V = VecOf<T>::N;
    size_t e = (size_t)loff(cap) + 5 * (size_t)cap + (size_t)round_up(n, V);
    size_t b = e * sizeof(T) + (size_t)cap * sizeof(int) + (size_t)round_up(m, V);
    return (b + 15) / 16 * 16;
}

template <typename T, int NG>
struct Warp {
    static constexpr int V = VecOf<T>::N;
    static constexpr int ROWB = (NG == 1) ? 8 : (NG == 2 ? 4 : 2); // active rows fetched per batch
    uint64_t pol_keep, pol_stream; // L2 policies: active rows are re-read every iteration, the scan is a pure stream
    // shared memory
    T *L, *D, *lam, *lams, *xl, *zl, *u;
    int* WS;
    unsigned char* sense;
    // problem
    const T *Mt, *Mr, *du, *dl, *sc;
    int n, m, ldm, ldn, cap, lane;
    const DevSettings<T>* st;
    int *pst_id; T* pst_lam;
    // warp-uniform state
    int k, reuse, sing;
    T fval, soft_slack;
    int n_scan, n_add, n_remove, n_csp;

    __device__ __forceinline__ void reset() { sing = EMPTY_IND; k = 0; reuse = 0; } // daqp.c:142-146

    // ---- a2: LDL' row append (factorization.c:21-111) + bookkeeping of daqp_add_constraint (auxiliary.c:27-41)
    __device__ void raw_add(int add, T lamval) {
        n_add++;
        if (lane == 0) sense[add] |= B_ACTIVE;
        sing = EMPTY_IND;
        T mi[NG][V];
        const T* rowi = Mr + (size_t)add * ldn;
        T part = 0;
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const int c = V * (lane + 32 * g);
#pragma unroll
            for (int e = 0; e < V; e++) mi[g][e] = 0;
            if (c < ldn) ldg_vec_hint<T>(rowi + c, mi[g], pol_keep);
#pragma unroll
            for (int e = 0; e < V; e++) part += mi[g][e] * mi[g][e];
        }
        T d = warp_sum(part);
        const int kk = k;
        T* Lk = L + loff(kk);
        if (kk > 0) {
            // l_j = M_{WS[j]} . m_add : ROWB rows per batch -> ROWB*NG independent 128-bit loads in flight, then one
            // transposed butterfly sums all ROWB dot products at once
            for (int j0 = 0; j0 < kk; j0 += ROWB) {
                T tv[ROWB][NG][V];
#pragma unroll
                for (int b = 0; b < ROWB; b++) {
                    const int j = min(j0 + b, kk - 1); // clamp keeps the address valid; surplus results are dropped
                    const T* rowj = Mr + (size_t)WS[j] * ldn;
#pragma unroll
                    for (int g = 0; g < NG; g++) {
                        const int c = V * (lane + 32 * g);
#pragma unroll
                        for (int e = 0; e < V; e++) tv[b][g][e] = 0;
                        if (c < ldn) ldg_vec_hint<T>(rowj + c, tv[b][g], pol_keep);
                    }
                }
                T pv[ROWB];
#pragma unroll
                for (int b = 0; b < ROWB; b++) {
                    pv[b] = 0;
#pragma unroll
                    for (int g = 0; g < NG; g++)
#pragma unroll
                        for (int e = 0; e < V; e++) pv[b] += tv[b][g][e] * mi[g][e];
                }
                const T tot = warp_sum_multi<ROWB>(pv, lane);
                const int j = j0 + multi_index<ROWB>(lane);
                if ((lane & (32 / ROWB - 1)) == 0 && j < kk) Lk[j] = tot;
            }
            __syncwarp();
            // l <- L^-1 l : column sweep, one broadcast + one FMA per step
            for (int j = 0; j < kk - 1; j++) {
                const T lj = Lk[j];
                for (int i = j + 1 + lane; i < kk; i += 32) Lk[i] -= L[loff(i) + j] * lj;
                __syncwarp();
            }
            // l <- D^-1 l ; d -= l' D l
            T acc = 0;
            for (int i = lane; i < kk; i += 32) {
                const T t = Lk[i];
                const T q = t / D[i];
                Lk[i] = q;
                acc += t * q;
            }
            d -= warp_sum(acc);
            if (d < st->sing_tol || kk >= n) { // ns_active == 0 on this path (soft constraints not in kernel yet)
                sing = kk;
                d = 0;
            }
        }
        if (lane == 0) { D[kk] = d; WS[kk] = add; lam[kk] = lamval; }
        k = kk + 1;
        __syncwarp();
    }

    // ---- a3: LDL' row/column deletion + Gill-Golub-Murray-Saunders C1 update (factorization.c:112-151)
    //      + bookkeeping of daqp_remove_constraint (auxiliary.c:3-22). Returns 1 if the factor became singular.
    __device__ int raw_remove(int r) {
        n_remove++;
        const int kk = k;
        if (lane == 0) sense[WS[r]] &= ~B_ACTIVE;
        if (r != kk - 1) {
            const int nu = kk - r - 1;
            for (int t = lane; t < nu; t += 32) zl[r + t] = L[loff(r + 1 + t) + r]; // removed column
            __syncwarp();
            // compaction: new row i-1 <- old row i without column r (source and destination never overlap)
            for (int i = r + 1; i < kk; i++) {
                const T* src = L + loff(i);
                T* dst = L + loff(i - 1);
                for (int j = lane; j < i; j += 32)
                    if (j != r) dst[j - (j > r)] = src[j];
                __syncwarp();
            }
            T alpha = D[r];
            for (int t = 0; t < nu; t++) {
                const int c = r + t; // new index of this pivot (old index c+1)
                const T p = zl[c];
                const T Dold = D[c + 1];
                const T dbar = Dold + alpha * p * p;
                const T beta = p * alpha / dbar;
                alpha = Dold * alpha / dbar;
                __syncwarp(); // everyone has read D[c+1] / zl[c] before they are overwritten below
                if (lane == 0) D[c] = dbar;
                for (int s = t + 1 + lane; s < nu; s += 32) {
                    T* Lrc = L + loff(r + s) + c;
                    const T lv = *Lrc;
                    const T qs = zl[r + s] - p * lv;
                    zl[r + s] = qs;
                    *Lrc = lv + beta * qs;
                }
                __syncwarp();
            }
        }
        k = kk - 1;
        for (int base = r; base < k; base += 32) { // shift WS / lam down by one
            const int i = base + lane;
            int wv = 0; T lv = 0;
            if (i < k) { wv = WS[i + 1]; lv = lam[i + 1]; }
            __syncwarp();
            if (i < k) { WS[i] = wv; lam[i] = lv; }
        }
        __syncwarp();
        if (r < reuse) reuse = r;
        if (k > 0 && D[k - 1] < st->sing_tol) {
            sing = k - 1;
            __syncwarp();
            if (lane == 0) D[k - 1] = 0;
            __syncwarp();
            return 1;
        }
        return 0;
    }

    // ---- a10: daqp_pivot_last (auxiliary.c:379-396), recursion unrolled onto an explicit stack:
    // each pending "re-add the removed row once the nested removal returns" is one entry.
    __device__ void pivot_last() {
        int depth = 0;
        for (;;) {
            const int r = k - 2;
            if (k > 1) {
                const T Dr = D[r], Dl = D[k - 1];
                if (Dr < st->pivot_tol && Dr < Dl) {
                    if (lane == 0) { pst_id[depth] = WS[r]; pst_lam[depth] = lam[r]; }
                    __syncwarp();
                    depth++;
                    if (!raw_remove(r)) continue;
                }
            }
            for (;;) {
                if (depth == 0) return;
                depth--;
                if (sing != EMPTY_IND) continue;
                raw_add(pst_id[depth], pst_lam[depth]);
                break;
            }
        }
    }
    __device__ __forceinline__ void add_constraint(int add, T lamval) { raw_add(add, lamval); pivot_last(); }
    __device__ __forceinline__ void remove_constraint(int r) { if (!raw_remove(r)) pivot_last(); }

    // ---- a5: constrained stationary point L D L' lam* = -d_k (auxiliary.c:314-354), forward solve resumes at reuse
    __device__ void compute_csp() {
        n_csp++;
        const int kk = k, r = reuse;
        for (int i = r + lane; i < kk; i += 32) {
            const int id = WS[i];
            xl[i] = (sense[id] & B_LOWER) ? -dl[id] : -du[id];
        }
        __syncwarp();
        if (kk - r <= 3) {
            // few new rows (the common case after an add): one warp-wide dot product per row
            for (int i = r; i < kk; i++) {
                const T* Li = L + loff(i);
                T acc = 0;
                for (int j = lane; j < i; j += 32) acc += Li[j] * xl[j];
                acc = warp_sum(acc);
                if (lane == 0) xl[i] -= acc;
                __syncwarp();
            }
        } else {
            // rows >= r first absorb the already-solved prefix (independent per row), then a column sweep
            if (r > 0) {
                for (int i = r + lane; i < kk; i += 32) {
                    const T* Li = L + loff(i);
                    T s = xl[i];
                    for (int j = 0; j < r; j++) s -= Li[j] * xl[j];
                    xl[i] = s;
                }
                __syncwarp();
            }
            for (int j = r; j < kk - 1; j++) {
                const T xj = xl[j];
                for (int i = j + 1 + lane; i < kk; i += 32) xl[i] -= L[loff(i) + j] * xj;
                __syncwarp();
            }
        }
        for (int i = r + lane; i < kk; i += 32) zl[i] = xl[i] / D[i];
        __syncwarp();
        for (int i = lane; i < kk; i += 32) lams[i] = zl[i];
        __syncwarp();
        for (int j = kk - 1; j > 0; j--) { // L' lam* = z : row j of L is contiguous -> conflict-free
            const T lj = lams[j];
            const T* Lj = L + loff(j);
            for (int i = lane; i < j; i += 32) lams[i] -= Lj[i] * lj;
            __syncwarp();
        }
        reuse = kk;
    }

    // ---- a6: dual ratio test, step, removal (auxiliary.c:277-311)
    __device__ int remove_blocking() {
        T best = (T)1e30;
        int key = INT_MAX;
        const T dual_tol = st->dual_tol;
        for (int i = lane; i < k; i += 32) {
            const int sb = sense[WS[i]];
            if (sb & B_IMMUTABLE) continue;
            const T ls = lams[i];
            if (sb & B_LOWER) { if (ls < dual_tol) continue; }
            else if (ls > -dual_tol) continue;
            const T l = lam[i];
            const T ac = (sing == EMPTY_IND) ? -l / (ls - l) : -l / ls;
            if (ac < best) { best = ac; key = i; }
        }
        warp_argmin(best, key);
        if (key == INT_MAX) return 0;
        if (sing == EMPTY_IND) { for (int i = lane; i < k; i += 32) lam[i] += best * (lams[i] - lam[i]); }
        else { for (int i = lane; i < k; i += 32) lam[i] += best * lams[i]; }
        __syncwarp();
        sing = EMPTY_IND;
        remove_constraint(key);
        return 1;
    }

    // ---- a7: u = -Mk' lam*, fval = |u|^2 (auxiliary.c:46-88)
    __device__ void compute_primal() {
        T acc[NG][V];
#pragma unroll
        for (int g = 0; g < NG; g++)
#pragma unroll
            for (int e = 0; e < V; e++) acc[g][e] = 0;
        for (int i0 = 0; i0 < k; i0 += ROWB) { // ROWB rows per batch: all loads first, then the FMAs in index order
            T tv[ROWB][NG][V];
            T li[ROWB];
#pragma unroll
            for (int b = 0; b < ROWB; b++) {
                const int i = i0 + b;
                const T* row = Mr + (size_t)WS[min(i, k - 1)] * ldn;
                li[b] = (i < k) ? lams[i] : (T)0;
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    const int c = V * (lane + 32 * g);
#pragma unroll
                    for (int e = 0; e < V; e++) tv[b][g][e] = 0;
                    if (c < ldn && i < k) ldg_vec_hint<T>(row + c, tv[b][g], pol_keep);
                }
            }
#pragma unroll
            for (int b = 0; b < ROWB; b++)
#pragma unroll
                for (int g = 0; g < NG; g++)
#pragma unroll
                    for (int e = 0; e < V; e++) acc[g][e] -= tv[b][g][e] * li[b];
        }
        T part = 0;
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const int c = V * (lane + 32 * g);
            if (c < ldn) {
#pragma unroll
                for (int e = 0; e < V; e++) { u[c + e] = acc[g][e]; part += acc[g][e] * acc[g][e]; }
            }
        }
        soft_slack = 0;
        fval = warp_sum(part);
        __syncwarp();
    }

    // ---- a8: Mu = M u for all rows, most-violated inactive row (auxiliary.c:89-198). Returns 1 if a row was added.
    __device__ int add_infeasible() {
        n_scan++;
        constexpr int SG = 4;            // row groups per pass: up to SG 128-bit loads per lane and column
        constexpr int GR = 32 * V;       // rows per group
        T best = 0;
        int key = INT_MAX;
        const T ep = -st->primal_tol;
        for (int base = 0; base < m; base += SG * GR) {
            const int r0 = base + V * lane;
            const int ngrp = min(SG, (ldm - base + GR - 1) / GR); // warp-uniform
            T acc[SG][V];
#pragma unroll
            for (int g = 0; g < SG; g++)
#pragma unroll
                for (int e = 0; e < V; e++) acc[g][e] = 0;
            const T* col = Mt + r0;
#pragma unroll 4
            for (int c = 0; c < n; c++) {
                const T uc = u[c];
                T t[SG][V];
#pragma unroll
                for (int g = 0; g < SG; g++) {
#pragma unroll
                    for (int e = 0; e < V; e++) t[g][e] = 0;
                    if (g < ngrp && r0 + g * GR < ldm) ldg_vec_hint<T>(col + g * GR, t[g], pol_stream);
                }
#pragma unroll
                for (int g = 0; g < SG; g++)
#pragma unroll
                    for (int e = 0; e < V; e++) acc[g][e] += t[g][e] * uc;
                col += ldm;
            }
#pragma unroll
            for (int g = 0; g < SG; g++) {
                const int rr = r0 + g * GR;
                if (g < ngrp && rr < m) { // ldm is a multiple of V: the whole vector is addressable
                    T bu[V], bl[V], bs[V];
                    ldg_vec<T>(du + rr, bu);
                    ldg_vec<T>(dl + rr, bl);
                    ldg_vec<T>(sc + rr, bs);
#pragma unroll
                    for (int e = 0; e < V; e++) {
                        const int row = rr + e;
                        if (row >= m) continue;
                        if (sense[row] & (B_ACTIVE + B_IMMUTABLE)) continue;
                        const T mu = acc[g][e];
                        const T bound = ep * bs[e];
                        T cand = bu[e] - mu;
                        if (cand < best && cand < bound) { best = cand; key = 2 * row; }
                        else {
                            cand = mu - bl[e];
                            if (cand < best && cand < bound) { best = cand; key = 2 * row + 1; }
                        }
                    }
                }
            }
        }
        warp_argmin(best, key);
        if (key == INT_MAX) return 0;
        const int add = key >> 1, lower = key & 1;
        if (lane == 0) { if (lower) sense[add] |= B_LOWER; else sense[add] &= ~B_LOWER; }
        T* t = lam; lam = lams; lams = t; // lam <- lam* (auxiliary.c:159-160)
        __syncwarp();
        add_constraint(add, lower ? (T)-1 : (T)1);
        return 1;
    }

    // ---- a9: singular direction (auxiliary.c:357-376)
    __device__ void singular_direction() {
        const int s = sing;
        const T* Ls = L + loff(s);
        for (int i = lane; i < s; i += 32) lams[i] = -Ls[i];
        __syncwarp();
        for (int j = s - 1; j > 0; j--) {
            const T lj = lams[j];
            const T* Lj = L + loff(j);
            for (int i = lane; i < j; i += 32) lams[i] -= Lj[i] * lj;
            __syncwarp();
        }
        const bool flip = sense[WS[s]] & B_LOWER;
        for (int i = lane; i <= s; i += 32) {
            T val = (i == s) ? (T)1 : lams[i];
            lams[i] = flip ? -val : val;
        }
        __syncwarp();
    }

    // ---- a13: warm start / equality activation (auxiliary.c:399-479)
    __device__ int activate_constraints() {
        for (int i = 0; i < m; i++) {
            const int sb = sense[i];
            if (sb & B_ACTIVE) add_constraint(i, (sb & B_LOWER) ? (T)-1 : (T)1);
            if (sing != EMPTY_IND) {
                const int last = WS[k - 1];
                if (sense[last] & B_IMMUTABLE) {
                    singular_direction();
                    T resid = 0, scale = 0;
                    for (int j = lane; j < k; j += 32) {
                        const int id = WS[j];
                        const T b = (sense[id] & B_LOWER) ? dl[id] : du[id];
                        const T term = lams[j] * b;
                        resid += term;
                        scale += term < 0 ? -term : term;
                    }
                    resid = warp_sum(resid);
                    scale = (T)1 + warp_sum(scale);
                    if (lane == 0) sense[last] &= ~B_ACTIVE;
                    k--;
                    sing = EMPTY_IND;
                    if (reuse > k) reuse = k;
                    __syncwarp();
                    if (resid <= st->primal_tol * scale && resid >= -st->primal_tol * scale) continue;
                    return EXIT_OVERDETERMINED_INITIAL;
                }
                int flag = 1;
                for (int j = i + lane; j < m; j += 32) {
                    const int s2 = sense[j];
                    if (s2 & B_ACTIVE) {
                        if (s2 & B_IMMUTABLE) flag = EXIT_OVERDETERMINED_INITIAL;
                        else sense[j] = s2 & ~B_ACTIVE;
                    }
                }
                flag = __shfl_sync(FULL, __reduce_min_sync(FULL, flag), 0);
                k--;
                sing = EMPTY_IND;
                __syncwarp();
                return flag;
            }
        }
        return 1;
    }

    // ---- a12: one step of iterative refinement on the active rows (auxiliary.c:498-593)
    __device__ void refine_active() {
        reuse = 0;
        const int kk = k;
        for (int i = 0; i < kk; i++) {
            const int id = WS[i];
            const T* row = Mr + (size_t)id * ldn;
            T part = 0;
#pragma unroll
            for (int g = 0; g < NG; g++) {
                const int c = V * (lane + 32 * g);
                if (c < ldn) {
                    T t[V];
                    ldg_vec<T>(row + c, t);
#pragma unroll
                    for (int e = 0; e < V; e++) part += t[e] * u[c + e];
                }
            }
            part = warp_sum(part);
            const T dd = (sense[id] & B_LOWER) ? dl[id] : du[id];
            if (lane == 0) xl[i] = part - dd;
        }
        __syncwarp();
        for (int j = 0; j < kk - 1; j++) {
            const T xj = xl[j];
            for (int i = j + 1 + lane; i < kk; i += 32) xl[i] -= L[loff(i) + j] * xj;
            __syncwarp();
        }
        for (int i = lane; i < kk; i += 32) { const T z = xl[i] / D[i]; zl[i] = z; xl[i] = z; }
        __syncwarp();
        for (int j = kk - 1; j > 0; j--) {
            const T xj = xl[j];
            const T* Lj = L + loff(j);
            for (int i = lane; i < j; i += 32) xl[i] -= Lj[i] * xj;
            __syncwarp();
        }
        for (int i = lane; i < kk; i += 32) lams[i] += xl[i];
        T acc[NG][V];
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const int c = V * (lane + 32 * g);
#pragma unroll
            for (int e = 0; e < V; e++) acc[g][e] = (c < ldn) ? u[c + e] : (T)0;
        }
        for (int i = 0; i < kk; i++) {
            const T* row = Mr + (size_t)WS[i] * ldn;
            const T dlam = xl[i];
#pragma unroll
            for (int g = 0; g < NG; g++) {
                const int c = V * (lane + 32 * g);
                if (c < ldn) {
                    T t[V];
                    ldg_vec<T>(row + c, t);
#pragma unroll
                    for (int e = 0; e < V; e++) acc[g][e] -= t[e] * dlam;
                }
            }
        }
        __syncwarp();
        T part = 0;
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const int c = V * (lane + 32 * g);
            if (c < ldn) {
#pragma unroll
                for (int e = 0; e < V; e++) { u[c + e] = acc[g][e]; part += acc[g][e] * acc[g][e]; }
            }
        }
        fval = soft_slack + warp_sum(part);
        __syncwarp();
    }

    // ---- a11: the daqp_ldp state machine (daqp.c:6-108). Returns the exit flag, iterations in *iters.
    __device__ int solve(int* iters) {
        int exitflag = EXIT_ITERLIMIT, iter;
        int tried_repair = 0, cycle_counter = 0;
        T best_fval = -1;
        const T fval_bound = 2 * st->fval_bound;
        for (iter = 1; iter < st->iter_limit; ++iter) {
            if (sing == EMPTY_IND) {
                compute_csp();
                if (!remove_blocking()) {
                    compute_primal();
                    if (fval > fval_bound) { exitflag = EXIT_INFEASIBLE; break; }
                    if (!add_infeasible()) {
                        T min_D = (T)1e30;
                        for (int i = lane; i < k; i += 32) min_D = fmin(min_D, D[i]);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) min_D = fmin(min_D, __shfl_xor_sync(FULL, min_D, o));
                        if (k > 2 && tried_repair != 1 && min_D < st->refactor_tol) {
                            tried_repair = 1;
                            for (int i = lane; i < k; i += 32) {
                                const int id = WS[i];
                                if (lam[i] >= 0) sense[id] &= ~B_LOWER; else sense[id] |= B_LOWER;
                            }
                            __syncwarp();
                            reset();
                            activate_constraints();
                            continue;
                        }
                        if (k > 0 && min_D < st->pivot_tol) {
                            refine_active();
                            if (add_infeasible()) continue;
                        }
                        exitflag = (soft_slack > st->primal_tol) ? EXIT_SOFT_OPTIMAL : EXIT_OPTIMAL;
                        break;
                    }
                    if (fval - best_fval < st->progress_tol) {
                        if (cycle_counter++ > st->cycle_tol) {
                            if (tried_repair == 1) { exitflag = EXIT_CYCLE; break; }
                            tried_repair = 1;
                            reset();
                            activate_constraints();
                            cycle_counter = 0;
                            best_fval = -1;
                        }
                    } else {
                        best_fval = fval;
                        cycle_counter = 0;
                    }
                }
            } else {
                singular_direction();
                if (!remove_blocking()) { exitflag = EXIT_INFEASIBLE; break; }
            }
        }
        *iters = iter;
        return exitflag;
    }
};

// One warp per problem, persistent CTAs pulling problem indices from an atomic queue (iteration counts diverge).
template <typename T, int NG>
__global__ void __launch_bounds__(512, 1) ldp_solve_kernel(const LdpArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int V = VecOf<T>::N;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * (blockDim.x >> 5) + wib;
    const size_t per_warp = ldp_smem_per_warp<T>(a.n, a.m, a.cap);
    unsigned char* base = smem_raw + per_warp * wib;

    Warp<T, NG> w;
    w.lane = lane;
    w.pol_keep = policy_evict_last();
    w.pol_stream = policy_evict_first();
    w.n = a.n; w.m = a.m; w.ldm = a.ldm; w.ldn = a.ldn; w.cap = a.cap;
    w.st = &a.st;
    T* fp = reinterpret_cast<T*>(base);
    w.L = fp; fp += loff(a.cap);
    w.D = fp; fp += a.cap;
    T* lamA = fp; fp += a.cap;
    T* lamB = fp; fp += a.cap;
    w.xl = fp; fp += a.cap;
    w.zl = fp; fp += a.cap;
    w.u = fp; fp += round_up(a.n, V);
    w.WS = reinterpret_cast<int*>(fp);
    w.sense = reinterpret_cast<unsigned char*>(w.WS + a.cap);
    w.pst_id = a.pst_id + (size_t)gw * a.cap;
    w.pst_lam = a.pst_lam + (size_t)gw * a.cap;

    const int ntri = a.n * (a.n + 1) / 2;
    for (;;) {
        int p = 0;
        if (lane == 0) p = atomicAdd(a.work_counter, 1);
        p = __shfl_sync(FULL, p, 0);
        if (p >= a.P) break;
        const int sflag = a.setup_flag[p];
        if (sflag != SETUP_SOLVE && sflag != SETUP_SOLVE_ACTIVATE) { // finished by the setup kernel
            if (a.nact_out && lane == 0) a.nact_out[p] = 0;
            if (a.counts_out && lane < 4) a.counts_out[4 * (size_t)p + lane] = 0;
            if (a.sense_out)
                for (int i = lane; i < a.m; i += 32) a.sense_out[(size_t)p * a.ldm + i] = a.sense[(size_t)p * a.ldm + i];
            continue;
        }

        w.Mt = a.Mt + (size_t)p * a.n * a.ldm;
        w.Mr = a.Mr + (size_t)p * a.m * a.ldn;
        w.du = a.dupper + (size_t)p * a.ldm;
        w.dl = a.dlower + (size_t)p * a.ldm;
        w.sc = a.scaling + (size_t)p * a.ldm;
        w.lam = lamA; w.lams = lamB;
        w.fval = 0; w.soft_slack = 0;
        w.n_scan = w.n_add = w.n_remove = w.n_csp = 0;
        const unsigned char* sin = a.sense + (size_t)p * a.ldm;
        for (int i = lane; i < a.m; i += 32) w.sense[i] = sin[i];
        for (int i = lane; i < round_up(a.n, V); i += 32) w.u[i] = 0;
        __syncwarp();
        w.reset();

        int exitflag = 1, iters = 0;
        if (sflag == SETUP_SOLVE_ACTIVATE) exitflag = w.activate_constraints(); // utils.c:199-211
        if (exitflag < 0) {
            // setup failure (api.c:69-72): flag only, x untouched
            if (lane == 0) { a.exitflag[p] = exitflag; a.iter[p] = 0; }
        } else {
            exitflag = w.solve(&iters);
            // ---- a15/a16: ldp2qp_solution (daqp.c:111-139) + daqp_extract_result (api.c:455-495)
            const T* vv = a.v ? a.v + (size_t)p * a.n : nullptr;
            T* xo = a.x + (size_t)p * a.n;
            T vnorm = 0;
            if (vv) { for (int i = lane; i < a.n; i += 32) { const T t = vv[i]; vnorm += t * t; } vnorm = warp_sum(vnorm); }
            if (exitflag > 0) {
                const T* Ri = a.Rinv + (size_t)p * ntri;
                if (vv) for (int i = lane; i < a.n; i += 32) w.u[i] -= vv[i];
                __syncwarp();
                for (int i = 0; i < a.n; i++) { // x_i = sum_{j>=i} Rinv[i][j] (u-v)_j ; rows are independent
                    const T* row = Ri + roff(i, a.n);
                    T acc = 0;
                    for (int j = i + lane; j < a.n; j += 32) acc += row[j] * w.u[j];
                    acc = warp_sum(acc);
                    if (i < a.ms) acc /= w.sc[i];
                    if (lane == 0) xo[i] = acc;
                }
            } else {
                for (int i = lane; i < a.n; i += 32) xo[i] = w.u[i];
            }
            if (a.lam) {
                T* lo = a.lam + (size_t)p * a.m;
                for (int i = lane; i < a.m; i += 32) lo[i] = 0;
                __syncwarp();
                for (int i = lane; i < w.k; i += 32) {
                    const int id = w.WS[i];
                    lo[id] = (exitflag > 0) ? w.lams[i] * w.sc[id] : w.lams[i];
                }
            }
            if (lane == 0) {
                if (vv) a.fval[p] = (T)0.5 * (w.fval - vnorm);
                a.exitflag[p] = exitflag;
                a.iter[p] = iters;
            }
        }
        if (a.nact_out && lane == 0) a.nact_out[p] = w.k;
        if (a.ws_out) for (int i = lane; i < w.k; i += 32) a.ws_out[(size_t)p * a.cap + i] = w.WS[i];
        if (a.sense_out) for (int i = lane; i < a.m; i += 32) a.sense_out[(size_t)p * a.ldm + i] = w.sense[i];
        if (a.counts_out && lane == 0) {
            int* c = a.counts_out + 4 * (size_t)p;
            c[0] = w.n_scan; c[1] = w.n_add; c[2] = w.n_remove; c[3] = w.n_csp;
        }
        __syncwarp();
    }
}

} // namespace dq
