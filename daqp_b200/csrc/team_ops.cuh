// daqp_b200/csrc/team_ops.cuh -- the phases of the active-set iteration that a TEAM of TW warps (one CTA) executes
// together on ONE problem (team mode of ldp_solve_kernel, n > 64).
//
// Why a team: at n = 120 the packed LDL' factor is 58 KB, three problems fit an SM, and with a warp per problem the SM
// runs three warps. With four warps per problem the SM holds twelve, the feasibility scan, the row passes and the dot
// products are split four ways, and the triangular recurrences run blocked by 32 rows (the warp that owns a diagonal
// block finishes it with shuffle-broadcast pivots, the other warps apply the finished block to their rows).
//
// Structure: warp 0 of the CTA (the leader) runs the same state machine as the single-warp kernel (Warp<...>::step) and
// posts each heavy phase as a command in the CTA's mailbox (TeamBox, first bytes of shared memory); warps 1 .. TW-1 sit
// in a command loop. Both sides then call the SAME function below. These functions are deliberately free, __noinline__
// and argument-light: one copy of each in the kernel (the instruction cache is shared by twelve warps in different
// phases), and every shared-memory pointer is derived from the `extern __shared__` symbol inside the function, so that
// the compiler emits LDS / STS (a pointer that arrives through a struct member or an argument is generic to it, and the
// first version of this file ran on LD.E / ST.E with the solver state in local memory: 3.5x slower).
//
// Thread tid = 32 wid + lane owns row tid of the factor and element tid of the vectors. Every recurrence keeps the
// per-element operation order of the single-warp code (and so of the reference, factorization.c / auxiliary.c): only
// WHICH thread executes an update changes. A named barrier (id 1, 32 TW threads) separates the stages.
#pragma once
#include "common.cuh"
#include <limits.h>

namespace dq {

struct TeamBox {
    int cmd, a0, a1, p;      // command of the leader and its two arguments; problem index
    double alpha, fval;      // C1 recurrence carried between diagonal blocks of a removal; |u|^2 for the screening bound
    double rv[8], rv2[8];    // per-warp partial results (values)
    int rk[8], rs[8];        // per-warp partial results (keys / flags)
    // layout: byte offsets from the start of the CTA's shared memory (written once per launch by the leader)
    int oL, oD, olamA, olamB, oWS, osense, ou, ou32, otmp, opv, obv, oside;
    int cap, n, m, ldm, ldn, tune;
    double primal_tol, sing_tol, pivot_tol;
    // global arrays of the current problem (written by the leader when it takes a problem off the queue)
    const void *Mr, *Mt, *Mt32, *du, *dl, *sc;
};
constexpr int TEAM_BOX_BYTES = 384;
static_assert(sizeof(TeamBox) <= TEAM_BOX_BYTES, "TeamBox grew past its slot");
// Team widths: four warps per problem for 64 < n + 1 <= 128 (three problems fit an SM: twelve warps), two warps per
// problem for 32 < n + 1 <= 64 (the per-problem block without a staging arena is 14 KB at n = 50: twelve to fourteen
// problems fit, 24 - 28 warps instead of the 12 a warp per problem gives -- the single-warp kernel is bound by dependent
// latency at three warps per scheduler, not by issue slots).
constexpr int TEAM_WARPS = 4;
__host__ __device__ constexpr int team_max_ctas(int tw) { return tw == 2 ? 12 : 3; } // resident teams per SM (__launch_bounds__)
enum { TC_EXIT = 0, TC_FWD, TC_BWD, TC_REMOVE, TC_DOTS, TC_PRIMAL, TC_SCAN32, TC_SCAN64, TC_GRAM, TC_LDL };
constexpr int TEAM_GRAM_MIN = 8; // warm starts with fewer rows than this are activated one by one

#define TEAM_SMEM                                                   \
    extern __shared__ __align__(16) unsigned char smem_raw[];       \
    TeamBox* const box = reinterpret_cast<TeamBox*>(smem_raw)

// (bar.sync is an ALIGNED barrier: every lane of a warp must execute it together. The phases end in lane-predicated
// stores -- `if (lane == 0) box->... = ...`, `if (mine) tmp[i] = x` -- after which the hardware does not promise
// reconvergence, so the warp is converged explicitly first; compute-sanitizer's synccheck flags the barrier otherwise.)
// Leader and helpers reach the barrier from DIFFERENT instructions (the leader inside a phase wrapper, the helpers in
// their command loop). PTX identifies a barrier by its number, not by the instruction, so that is well defined; synccheck,
// however, only accepts a barrier all threads of the block reach at the same address. -DDAQP_B200_SYNCCHECK makes the
// barrier one out-of-line instruction for that tool (0 errors: profiles/sanitizer_r02.txt); the inlined form is 2-4 %
// faster on C4 and is what ships.
#ifdef DAQP_B200_SYNCCHECK
#define TEAM_BAR_INLINE __noinline__
#else
#define TEAM_BAR_INLINE __forceinline__
#endif
template <int TW> __device__ TEAM_BAR_INLINE void team_bar() {
    __syncwarp();
    asm volatile("bar.sync 1, %0;" ::"n"(32 * TW) : "memory");
}

// loads that stay in program order (asm volatile): a batch of them is issued back to back before the first use
__device__ __forceinline__ double ldg_ordered(const double* p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_ordered(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
template <bool LAST> // L2 evict_last (the screening copy is re-read every scan) or evict_first (pure stream)
__device__ __forceinline__ void ldg128_ordered(const void* p, float (&o)[4]) {
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "l"(p), "l"(LAST ? 0x14F0000000000000ull : 0x12F0000000000000ull));
}

// tmp <- L^-1 tmp on rows [rlo, len) (rows < rlo hold the solution already). Blocked by 32: the warp that owns the rows
// of block b finishes them with shuffle-broadcast pivots, publishes them, and the warps below apply the block's pivots to
// their own rows -- ascending pivots per row, the order of factorization.c:86-92 / auxiliary.c:334-337.
template <typename T, int TW>
__device__ __noinline__ void team_forward(int lane, int wid, int rlo, int len) {
    TEAM_SMEM;
    T* const tmp = reinterpret_cast<T*>(smem_raw + box->otmp);
    const int i = 32 * wid + lane;
    const bool mine = i >= rlo && i < len;
    T x = (i < len) ? tmp[i] : (T)0;
    const T* Li = reinterpret_cast<const T*>(smem_raw + box->oL) + loff(min(i, box->cap - 1));
    const int nb = (len + 30) >> 5; // blocks of the pivots 0 .. len-2
    for (int b = 0; b < nb; b++) {
        const int j0 = 32 * b, j1 = min(j0 + 32, len - 1);
        if (j0 + 32 > rlo) { // the block holds unsolved rows (uniform)
            if (wid == b) {
#pragma unroll 4
                for (int j = j0; j < j1; j++) {
                    const T lij = (mine && i > j) ? Li[j] : (T)0;
                    const T xj = __shfl_sync(FULL, x, j - j0);
                    if (mine && i > j) x -= lij * xj;
                }
                if (mine) tmp[i] = x;
            }
            team_bar<TW>();
        }
        if (wid > b && mine) {
#pragma unroll 8
            for (int j = j0; j < j1; j++) x -= Li[j] * tmp[j];
        }
    }
    if (mine && wid >= nb) tmp[i] = x; // rows below the last pivot block (the others were stored by their own block)
}

// tmp <- L^-T tmp ; descending pivots (auxiliary.c:343-352, 363-370)
template <typename T, int TW>
__device__ __noinline__ void team_backward(int lane, int wid, int len) {
    TEAM_SMEM;
    T* const tmp = reinterpret_cast<T*>(smem_raw + box->otmp);
    const T* const Lp = reinterpret_cast<const T*>(smem_raw + box->oL);
    const int i = 32 * wid + lane;
    T x = (i < len) ? tmp[i] : (T)0;
    for (int b = (len - 1) >> 5; b >= 0; b--) {
        const int j1 = min(32 * b + 31, len - 1), j0 = max(32 * b, 1); // pivots j1 .. j0 of this block
        if (wid == b) {
            const T* Lj = Lp + loff(j1) + i;
#pragma unroll 4
            for (int j = j1; j >= j0; j--) {
                const T lji = (i < j) ? Lj[0] : (T)0;
                const T xj = __shfl_sync(FULL, x, j - 32 * b);
                if (i < j) x -= lji * xj;
                Lj -= j - 1; // loff(j-1) = loff(j) - (j-1)
            }
            if (i < len) tmp[i] = x;
        }
        if (b > 0) {
            team_bar<TW>();
            if (wid < b) {
                const T* Lj = Lp + loff(j1) + i;
#pragma unroll 8
                for (int j = j1; j >= j0; j--) { x -= Lj[0] * tmp[j]; Lj -= j - 1; }
            }
        }
    }
}

// a3 for a team: delete row / column r of the factor of size kk (factorization.c:112-151). Thread s owns trailing row s
// (old row r+1+s), walks along it and writes every element straight to its compacted position (one row up, one column
// left). The pivots of block b are finished by warp b (Gill-Golub-Murray-Saunders C1 recurrences in the reference's
// order, lock-step), published (p_t, beta_t), and applied by the warps below. The first thread of a warp writes into the
// row of the LAST thread of the warp above, which runs unsynchronised: its 32 values of a stage are parked in a side
// buffer and put in place after the next barrier, when that row's reads of the stage are over.
template <typename T, int TW>
__device__ __noinline__ void team_remove(int lane, int wid, int r, int kk) {
    TEAM_SMEM;
    T* const Lp = reinterpret_cast<T*>(smem_raw + box->oL);
    T* const Dp = reinterpret_cast<T*>(smem_raw + box->oD);
    T* const pv = reinterpret_cast<T*>(smem_raw + box->opv);
    T* const bv = reinterpret_cast<T*>(smem_raw + box->obv);
    const int nu = kk - r - 1, s = 32 * wid + lane;
    const bool act = s < nu;
    const int io = min(r + 1 + s, box->cap - 1);
    const T* src = Lp + loff(io) + r + 1; // old element (s, t) = src[t]
    T* dst = Lp + loff(io - 1) + r;       // its compacted position = dst[t]
    T w = act ? src[-1] : (T)0;           // removed column
    // columns left of the removed one: row i moves up by one; thread j moves column j of every row
    if (s < r) {
        T* f = Lp + loff(r + 1) + s;
        for (int i = r + 1; i < kk; i++) { f[-(i - 1)] = f[0]; f += i; }
    }
    team_bar<TW>(); // every removed-column element is in a register before its slot is overwritten
    const bool boundary = wid > 0 && lane == 0;
    T* side = reinterpret_cast<T*>(smem_raw + box->oside) + 32 * (wid > 0 ? wid - 1 : 0);
    int pend0 = -1, pend1 = 0; // parked stage [pend0, pend1)
    const int nblk = (nu + 31) >> 5;
    for (int b = 0; b < nblk; b++) {
        const int t0 = 32 * b, t1 = min(t0 + 32, nu);
        if (wid == b) {
            T alpha = (b == 0) ? Dp[r] : (T)box->alpha;
            __syncwarp();
            for (int t = t0; t < t1; t++) {
                const T pvt = __shfl_sync(FULL, w, t - t0);
                const T Dold = Dp[r + 1 + t];
                const T dbar = Dold + alpha * pvt * pvt;
                const T rdb = frcp(dbar); // one reciprocal (= 1 / dbar, correctly rounded) for the two quotients of the reference
                const T beta = pvt * alpha * rdb;
                alpha = Dold * alpha * rdb;
                if (lane == 0) Dp[r + t] = dbar; // its old value was consumed one step earlier (as alpha for t = 0)
                if (lane == t - t0) { pv[t] = pvt; bv[t] = beta; }
                if (act && t < s) {
                    const T lv = src[t];
                    const T qs = w - pvt * lv;
                    w = qs;
                    dst[t] = lv + beta * qs;
                }
                __syncwarp();
            }
            if (lane == 0) box->alpha = (double)alpha;
        }
        team_bar<TW>();
        if (boundary && pend0 >= 0) { for (int t = pend0; t < pend1; t++) dst[t] = side[t - pend0]; pend0 = -1; }
        if (wid > b) { // whole warp: uniform trip count, lanes beyond the trailing rows idle
            for (int c0 = t0; c0 < t1; c0 += 8) {
                T lv[8];
#pragma unroll
                for (int e = 0; e < 8; e++) lv[e] = (act && c0 + e < t1) ? src[c0 + e] : (T)0;
                __syncwarp(); // the lane above has read these columns of its row before this lane overwrites them
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int t = c0 + e;
                    if (act && t < t1) {
                        const T qs = w - pv[t] * lv[e];
                        w = qs;
                        const T out = lv[e] + bv[t] * qs;
                        if (boundary) side[t - t0] = out; else dst[t] = out;
                    }
                }
            }
            if (boundary && act) { pend0 = t0; pend1 = t1; }
        }
    }
    team_bar<TW>();
    if (boundary && pend0 >= 0) for (int t = pend0; t < pend1; t++) dst[t] = side[t - pend0];
}

// l_j = M_{WS[j]} . m_add for j < kk, into row kk of the factor (factorization.c:59-84): the rows are dealt to the warps
// in batches of DB, every batch one transposed butterfly instead of DB full reductions. NG = 128-bit column groups.
template <typename T, int TW, int NG>
__device__ __noinline__ void team_dots(int lane, int wid, int add, int kk) {
    TEAM_SMEM;
    constexpr int V = VecOf<T>::N, DB = 8;
    const int ldn = box->ldn;
    const T* M = reinterpret_cast<const T*>(box->Mr) + V * lane; // this lane's column slice of row 0
    T mi[NG][V];
    bool okg[NG];
#pragma unroll
    for (int g = 0; g < NG; g++) {
        okg[g] = V * (lane + 32 * g) < ldn;
#pragma unroll
        for (int e = 0; e < V; e++) mi[g][e] = 0;
        if (okg[g]) ldg_vec<T>(M + (size_t)(unsigned)(add * ldn) + 32 * V * g, mi[g]);
    }
    const int* ws = reinterpret_cast<const int*>(smem_raw + box->oWS);
    T* Lk = reinterpret_cast<T*>(smem_raw + box->oL) + loff(kk);
    for (int j0 = DB * wid; j0 < kk; j0 += DB * TW) {
        T t[DB][NG][V];
#pragma unroll
        for (int rr = 0; rr < DB; rr++) {
            const T* row = M + (size_t)(unsigned)(ws[min(j0 + rr, kk - 1)] * ldn);
#pragma unroll
            for (int g = 0; g < NG; g++) {
#pragma unroll
                for (int e = 0; e < V; e++) t[rr][g][e] = 0;
                if (okg[g]) ldg_vec<T>(row + 32 * V * g, t[rr][g]);
            }
        }
        T pj[DB];
#pragma unroll
        for (int rr = 0; rr < DB; rr++) {
            pj[rr] = 0;
#pragma unroll
            for (int g = 0; g < NG; g++)
#pragma unroll
                for (int e = 0; e < V; e++) pj[rr] += t[rr][g][e] * mi[g][e];
        }
        const T total = warp_sum_multi<DB>(pj, lane);
        const int jr = j0 + multi_index<DB>(lane);
        if ((lane & (32 / DB - 1)) == 0 && jr < kk) Lk[jr] = total;
    }
}

// a13 for a team, first half: ALL the dot products K successive daqp_update_LDL_add calls would compute
// (factorization.c:40-84) in one pass -- g_ij = row(WS[i]) . row(WS[j]) into the strict lower triangle of L, g_ii into
// D. A warp takes four rows i at a time (register-resident) and streams the rows j <= i past them four at a time: every
// row fetched from L2 serves sixteen products, and no product waits for a factor update. Row blocks are dealt to the
// warps from the bottom (longest first).
template <typename T, int TW, int NG>
__device__ __noinline__ void team_gram(int lane, int wid, int K) {
    TEAM_SMEM;
    constexpr int V = VecOf<T>::N, IB = 4, JB = 4;
    const int ldn = box->ldn;
    const T* M = reinterpret_cast<const T*>(box->Mr) + V * lane; // this lane's column slice of row 0
    const int* ws = reinterpret_cast<const int*>(smem_raw + box->oWS);
    T* Lp = reinterpret_cast<T*>(smem_raw + box->oL);
    T* Dp = reinterpret_cast<T*>(smem_raw + box->oD);
    bool okg[NG];
#pragma unroll
    for (int g = 0; g < NG; g++) okg[g] = V * (lane + 32 * g) < ldn;
    const int nblk = (K + IB - 1) / IB;
    for (int b = nblk - 1 - wid; b >= 0; b -= TW) {
        const int i0 = IB * b, jend = min(i0 + IB, K);
        T mi[IB][NG][V];
#pragma unroll
        for (int r = 0; r < IB; r++) {
            const T* row = M + (size_t)(unsigned)(ws[min(i0 + r, K - 1)] * ldn);
#pragma unroll
            for (int g = 0; g < NG; g++) {
#pragma unroll
                for (int e = 0; e < V; e++) mi[r][g][e] = 0;
                if (okg[g]) ldg_vec<T>(row + 32 * V * g, mi[r][g]);
            }
        }
        for (int j0 = 0; j0 < jend; j0 += JB) {
            T t[JB][NG][V];
#pragma unroll
            for (int rr = 0; rr < JB; rr++) {
                const T* row = M + (size_t)(unsigned)(ws[min(j0 + rr, K - 1)] * ldn);
#pragma unroll
                for (int g = 0; g < NG; g++) {
#pragma unroll
                    for (int e = 0; e < V; e++) t[rr][g][e] = 0;
                    if (okg[g]) ldg_vec<T>(row + 32 * V * g, t[rr][g]);
                }
            }
            T pj[IB * JB];
#pragma unroll
            for (int r = 0; r < IB; r++)
#pragma unroll
                for (int rr = 0; rr < JB; rr++) {
                    T acc = 0;
#pragma unroll
                    for (int g = 0; g < NG; g++)
#pragma unroll
                        for (int e = 0; e < V; e++) acc += t[rr][g][e] * mi[r][g][e];
                    pj[r * JB + rr] = acc;
                }
            const T total = warp_sum_multi<IB * JB>(pj, lane);
            const int idx = multi_index<IB * JB>(lane), i = i0 + idx / JB, j = j0 + idx % JB;
            if ((lane & (32 / (IB * JB) - 1)) == 0 && i < K && j <= i) {
                if (j == i) Dp[i] = total; else Lp[loff(i) + j] = total;
            }
        }
    }
}

// a13 for a team, second half: the LDL' factor of the Gram matrix team_gram left in L / D, in place and right-looking --
// step j: thread i scales the entry of row i (l_ij = x_ij / d_j, factorization.c:96-99) and publishes l_ij and x_ij; after
// the barrier the rows below are dealt to the warps, which subtract l_tj x_rj from the entries t = j+1 .. r-1 of their
// rows: the products and their (ascending pivot) order are those of the forward substitutions of K successive daqp_update_LDL_add calls (factorization.c:86-92), and
// d_i = g_ii - sum_j x_ij l_ij (:100-103). The pass stops at the first row the one-by-one path treats specially -- a
// singular pivot (factorization.c:106-110) or a daqp_pivot_last swap (auxiliary.c:379-396) -- and reports its index in
// box->rk[0] (K = none); the leader then redoes the activation row by row.
template <typename T, int TW>
__device__ __noinline__ void team_ldl(int lane, int wid, int K) {
    TEAM_SMEM;
    T* const col = reinterpret_cast<T*>(smem_raw + box->otmp);  // l_ij of the step, by row
    T* const xcol = reinterpret_cast<T*>(smem_raw + box->opv);  // x_ij of the step, by row
    T* const Lp = reinterpret_cast<T*>(smem_raw + box->oL);
    T* Dp = reinterpret_cast<T*>(smem_raw + box->oD);
    volatile int* stop = box->rk;
    const int i = 32 * wid + lane;
    T* Li = Lp + loff(min(i, box->cap - 1));
    const T sing_tol = (T)box->sing_tol, pivot_tol = (T)box->pivot_tol;
    if (i == 0 && Dp[0] < sing_tol) stop[0] = 0;
    team_bar<TW>();
    T acc = 0;
    for (int j = 0; j + 1 < K; j++) {
        if (stop[0] <= j) break; // (written before the barrier that ended the previous step: the whole team sees it)
        // scale column j: thread = row
        const bool below = i > j && i < K;
        const T dj = Dp[j];
        const T x = below ? Li[j] : (T)0;
        const T l = fdiv(x, dj); // unconditional call: no divergence around the division
        if (below) {
            acc += x * l;
            Li[j] = l;
            col[i] = l;
            xcol[i] = x;
            if (i == j + 1) {
                const T d = Dp[i] - acc;
                Dp[i] = d;
                if (d < sing_tol || (dj < pivot_tol && dj < d)) stop[0] = i;
            }
        }
        team_bar<TW>();
        // trailing update L[r][t] -= l_tj x_rj for j < t < r: a WARP per row, lanes across its columns (contiguous, no bank
        // conflicts, and the rows -- of very different lengths -- are dealt round-robin; a thread per row left the team
        // waiting for the thread with the longest one: 3.6 k cycles per step, two thirds of a C4 activation)
        // (four rows per trip: their chains are independent, and with three teams on an SM there is little else to hide
        // the shared-memory round trips behind)
        for (int r0 = j + 2 + wid; r0 < K; r0 += 4 * TW) {
            T xr[4];
            T* Lr[4];
            int rr[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                rr[q] = r0 + q * TW;
                const int rc = min(rr[q], K - 1);
                xr[q] = xcol[rc];
                Lr[q] = Lp + loff(rc);
                if (rr[q] >= K) rr[q] = 0; // no columns: j + 1 + lane >= 1 > 0
            }
            for (int t = j + 1 + lane; t < rr[3] || t < rr[2] || t < rr[1] || t < rr[0]; t += 32) {
                const T ct = col[min(t, K - 1)];
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (t < rr[q]) Lr[q][t] -= ct * xr[q];
            }
        }
        team_bar<TW>();
    }
}

// a7 for a team: u = -sum_i lam*_i row(WS[i]); thread c owns column c and adds the rows in index order (the order of
// auxiliary.c:54-68), UNR rows in flight per thread. Per-warp partial |u|^2 goes to the box.
template <typename T, int TW>
__device__ __noinline__ void team_primal(int lane, int wid, int kk, int lsw) {
    TEAM_SMEM;
    constexpr int UNR = 16;
    const int c = 32 * wid + lane, ldn = box->ldn;
    const bool have = c < ldn;
    const T* ls = reinterpret_cast<const T*>(smem_raw + (lsw ? box->olamA : box->olamB));
    const int* ws = reinterpret_cast<const int*>(smem_raw + box->oWS);
    const T* col = reinterpret_cast<const T*>(box->Mr) + (have ? c : 0); // idle lanes re-read column 0
    T acc = 0;
    for (int i0 = 0; i0 < kk; i0 += UNR) {
        T v[UNR];
#pragma unroll
        for (int e = 0; e < UNR; e++) v[e] = ldg_ordered(col + (size_t)(unsigned)(ws[min(i0 + e, kk - 1)] * ldn));
#pragma unroll
        for (int e = 0; e < UNR; e++) if (i0 + e < kk) acc -= v[e] * ls[i0 + e];
    }
    if (have) {
        reinterpret_cast<T*>(smem_raw + box->ou)[c] = acc;
        reinterpret_cast<float*>(smem_raw + box->ou32)[c] = (float)acc;
    }
    const T part = warp_sum(have ? acc * acc : (T)0);
    if (lane == 0) box->rv[wid] = (double)part;
}

// a8 for a team, fp32 screening (see Warp::scan_screen for the error bound and the decision rule): the 32-row groups
// are dealt round-robin to the warps (group g = wid + TW r, row = 32 g + lane); quads of columns stream through PF
// register buffers per lane. Per-warp result -> box: rv = best, rk = key (INT_MAX: none), rv2 = runner-up, rs = sure.
template <int TW, int NR, int PF = 3, bool LAST = true>
__device__ __forceinline__ void team_scan32_n(int lane, int wid) {
    TEAM_SMEM;
    const int m = box->m, n = box->n;
    int rowi[NR];
    bool own[NR];
    float acc[NR];
    double bu[NR], bl[NR], bs[NR];
    const double* du = reinterpret_cast<const double*>(box->du);
    const double* dl = reinterpret_cast<const double*>(box->dl);
    const double* sc = reinterpret_cast<const double*>(box->sc);
#pragma unroll
    for (int r = 0; r < NR; r++) {
        rowi[r] = 32 * (wid + TW * r) + lane;
        own[r] = rowi[r] < m;
        rowi[r] = min(rowi[r], m - 1); // lanes beyond the last row re-read it (their result is discarded)
        acc[r] = 0.f;
        bu[r] = __ldg(du + rowi[r]); bl[r] = __ldg(dl + rowi[r]); bs[r] = __ldg(sc + rowi[r]);
    }
    const unsigned slab = (unsigned)m * 16u; // bytes of one quad of columns
    const char* src = reinterpret_cast<const char*>(box->Mt32);
    const int nq = (n + 3) >> 2;
    float buf[PF][NR][4];
#pragma unroll
    for (int i = 0; i < PF; i++)
#pragma unroll
        for (int r = 0; r < NR; r++) ldg128_ordered<LAST>(src + (size_t)min(i, nq - 1) * slab + 16u * rowi[r], buf[i][r]);
    const float* u32 = reinterpret_cast<const float*>(smem_raw + box->ou32);
    for (int q0 = 0; q0 < nq; q0 += PF) {
#pragma unroll
        for (int i = 0; i < PF; i++) {
            const int q = q0 + i;
            if (q < nq) {
                const float4 uq = *reinterpret_cast<const float4*>(u32 + 4 * q);
                const char* nx = src + (size_t)min(q + PF, nq - 1) * slab; // past the end: re-read the last quad (never used)
#pragma unroll
                for (int r = 0; r < NR; r++) {
                    acc[r] += buf[i][r][0] * uq.x; acc[r] += buf[i][r][1] * uq.y;
                    acc[r] += buf[i][r][2] * uq.z; acc[r] += buf[i][r][3] * uq.w;
                    ldg128_ordered<LAST>(nx + 16u * rowi[r], buf[i][r]);
                }
            }
        }
    }
    const double unorm = (double)sqrtf((float)box->fval) * 1.0001 + 1e-22;
    const double delta = 1.01 * (double)(n + 3) * 5.9604644775390625e-8 * unorm;
    const double ep = -box->primal_tol;
    const unsigned char* se = smem_raw + box->osense;
    double best = 1e300, second = 1e300;
    int key = INT_MAX;
    bool best_sure = false;
#pragma unroll
    for (int r = 0; r < NR; r++) { // ascending rows per lane: strict '<' keeps the first of equal candidates
        const double mu = (double)acc[r];
        const double cu = bu[r] - mu, cl = mu - bl[r];
        const bool lower = cl < cu;
        const double cand = lower ? cl : cu;
        const double bound = ep * bs[r];
        const bool possible = own[r] && !(se[rowi[r]] & (B_ACTIVE + B_IMMUTABLE)) && cand - delta < bound;
        const bool nb = possible && cand < best;
        second = nb ? best : ((possible && cand < second) ? cand : second);
        best_sure = nb ? (cand + delta < bound) : best_sure;
        key = nb ? 2 * rowi[r] + (int)lower : key;
        best = nb ? cand : best;
    }
    double wbest = best;
    int wkey = key;
    warp_argmin(wbest, wkey);
    wkey = __shfl_sync(FULL, wkey, 0);
    wbest = __shfl_sync(FULL, wbest, 0);
    double other = (key == wkey) ? second : best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) other = fmin(other, __shfl_xor_sync(FULL, other, o));
    const bool sure = __any_sync(FULL, key == wkey && best_sure);
    if (lane == 0) { box->rv[wid] = wbest; box->rk[wid] = wkey; box->rv2[wid] = other; box->rs[wid] = sure ? 1 : 0; }
}
template <int TW>
__device__ __noinline__ void team_scan32(int lane, int wid) {
    TEAM_SMEM;
    switch ((((box->m + 31) >> 5) + TW - 1) / TW) {
        case 1: team_scan32_n<TW, 1>(lane, wid); break;
        case 2: team_scan32_n<TW, 2>(lane, wid); break;
        case 3: team_scan32_n<TW, 3>(lane, wid); break;
        case 4: // (the C4 shape; experiment knobs: tune & 8 = stream policy for the screening copy, tune & 64 = deeper pipeline)
            if (box->tune & 8) team_scan32_n<TW, 4, 3, false>(lane, wid);
            else if (box->tune & 64) team_scan32_n<TW, 4, 5, true>(lane, wid);
            else team_scan32_n<TW, 4>(lane, wid);
            break;
        case 5: team_scan32_n<TW, 5>(lane, wid); break;
        default: team_scan32_n<TW, 6>(lane, wid); break; // m <= 768 with four warps (the host enables the screening up to there)
    }
}
constexpr int TEAM_SCREEN_MAX_GROUPS = 6;

// exact scan in T (auxiliary.c:89-152): groups of 32 V rows of the column-major matrix, dealt round-robin to the warps;
// UN columns in flight per lane. Per-warp result -> box: rv = most negative slack, rk = 2 row + lower (INT_MAX: none).
template <typename T, int TW>
__device__ __noinline__ void team_scan64(int lane, int wid) {
    TEAM_SMEM;
    constexpr int V = VecOf<T>::N, GR = 32 * V, UN = 8;
    const int m = box->m, n = box->n, ldm = box->ldm;
    const T* u = reinterpret_cast<const T*>(smem_raw + box->ou); // zero-padded behind column n-1
    const unsigned char* se = smem_raw + box->osense;
    const T* du = reinterpret_cast<const T*>(box->du);
    const T* dl = reinterpret_cast<const T*>(box->dl);
    const T* sc = reinterpret_cast<const T*>(box->sc);
    const T ep = -(T)box->primal_tol;
    T best = 0;
    int key = INT_MAX;
    for (int base = wid * GR; base < m; base += TW * GR) {
        const int r0 = base + V * lane;
        const bool own = r0 < ldm; // ldm is a multiple of V: the whole vector is addressable
        const T* col = reinterpret_cast<const T*>(box->Mt) + (own ? r0 : 0);
        T acc[V];
#pragma unroll
        for (int e = 0; e < V; e++) acc[e] = 0;
        for (int c0 = 0; c0 < n; c0 += UN) {
            T buf[UN][V];
#pragma unroll
            for (int i = 0; i < UN; i++) ldg_vec_hint_ordered<T>(col + (size_t)min(c0 + i, n - 1) * ldm, buf[i], policy_evict_first());
#pragma unroll
            for (int i = 0; i < UN; i++) {
                const T uc = u[c0 + i]; // zero for c0 + i >= n (U_PAD)
#pragma unroll
                for (int e = 0; e < V; e++) acc[e] += buf[i][e] * uc;
            }
        }
        if (own && r0 < m) {
            T bu[V], bl[V], bs[V];
            ldg_vec<T>(du + r0, bu); ldg_vec<T>(dl + r0, bl); ldg_vec<T>(sc + r0, bs);
#pragma unroll
            for (int e = 0; e < V; e++) {
                const int row = r0 + e;
                if (row >= m) continue;
                if (se[row] & (B_ACTIVE + B_IMMUTABLE)) continue;
                const T mu = acc[e];
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
    warp_argmin(best, key);
    if (lane == 0) { box->rv[wid] = (double)best; box->rk[wid] = key; }
}

// what a helper warp does for one command; the leader calls the same functions with wid = 0
template <typename T, int TW, int NG>
__device__ __forceinline__ void team_dispatch(int cmd, int lane, int wid, int a0, int a1) {
    switch (cmd) {
        case TC_FWD: team_forward<T, TW>(lane, wid, a0, a1); break;
        case TC_BWD: team_backward<T, TW>(lane, wid, a0); break;
        case TC_REMOVE: team_remove<T, TW>(lane, wid, a0, a1); break;
        case TC_DOTS: team_dots<T, TW, NG>(lane, wid, a0, a1); break;
        case TC_PRIMAL: team_primal<T, TW>(lane, wid, a0, a1); break;
        case TC_GRAM: team_gram<T, TW, NG>(lane, wid, a0); break;
        case TC_LDL: team_ldl<T, TW>(lane, wid, a0); break;
        case TC_SCAN32: if constexpr (sizeof(T) == 8) team_scan32<TW>(lane, wid); break;
        default: team_scan64<T, TW>(lane, wid); break;
    }
}

} // namespace dq
