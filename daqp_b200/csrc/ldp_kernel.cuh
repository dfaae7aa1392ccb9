// daqp_b200/csrc/ldp_kernel.cuh -- the hot path: batched dual active-set LDP solve, one warp per problem.
//
// Covers SURVEY.md §8(a) rows a2..a16: LDL' rank-1 add / remove, CSP triangular solves, dual ratio test, primal
// update, feasibility scan, singular direction, pivoting, warm-start activation, refinement, the daqp_ldp state
// machine, and the LDP->QP back-transform + result extraction.
//
// Memory plan per problem
//   shared (warp-private, lives for the whole solve): packed L, D, lam, lam*, xldl, zldl, active bounds, u, WS, sense,
//                                             and a small cp.async staging arena (scan ring / active-row chunks)
//   global, streamed every feasibility scan : Mt32 = fp32 copy of the constraint matrix in quad layout (screening);
//                                             Mt   = fp64 column-major [n][ldm] for the exact scan when the
//                                                    screening cannot name the winner
//   global, gathered per add / primal update: Mr   = same matrix, fp64 row-major [m][ldn] (128-bit coalesced rows)
//   global, once                            : dupper, dlower, scaling, Rinv (packed), v
// All matrix copies are written by the setup kernel; rows 0..ms-1 are the normalised rows of R^-1 (simple bounds,
// zero-filled below the diagonal), rows ms..m-1 the normalised rows of A R^-1.
//
// Three constraints shaped this file (all measured, see profiles/ and DESIGN.md §6):
//  * Convergence proofs. ptxas guards every warp collective with a divergence check and a slow path unless it can
//    PROVE the warp is converged, and one unproven spot taints the rest of the enclosing loop. Hence: structured
//    control flow (a problem loop around an iteration loop, no `continue`, one back edge per loop), an explicit
//    __syncwarp() before a back edge that follows a lane-divergent `if`, uniform-trip LANE_LOOPs, and uni() around
//    every branch condition that comes from memory or from a butterfly reduction.
//  * Instruction cache. A dozen warps of one SM sit in a dozen different phases, so the HOT instruction footprint
//    must stay near the SM's ~32 KB instruction cache (hit rate 94% / 82% / 75% at 29 / 39 / 45 KB). Heavy routines
//    have ONE call site on the hot path, the row-chunk issue code is one shared __noinline__ function, pipelined loops
//    are rotated so that their prologue is the first trip, sweeps are rolled, and the features the headline path does
//    not need (soft constraints, workspace state) live in a separate instantiation (EXT).
//  * Shared memory sets the occupancy (the packed factor alone is 10 KB at n = 50): the staging arena is sized so
//    that twelve problems stay resident per SM.
#pragma once
#include "common.cuh"
#include "team_ops.cuh"
#include <limits.h>
#include <algorithm>

namespace dq {

constexpr int U_PAD = 16; // zero entries behind u so that the fp64 scan pipeline can run past column n-1
constexpr int RB = 6;     // active rows per cp.async chunk of the row passes (two chunks in flight); measured on C3:
                          // 8 -> 112.0 ms (11 problems per SM), 6 -> 107.7 ms (12 per SM), 4 -> 115.8 ms
constexpr int QRING = 2;  // depth of the screening scan's ring (quads of columns in flight); 3 buys nothing (the scan
                          // is not latency-bound) and costs the twelfth resident problem, 4 -> 125 ms, 1 -> 119 ms
constexpr int SCR_PITCH = 34; // doubles per parked row of partial dot products (16-byte aligned rows)
constexpr int TW_REGS = 0;    // Warp<..., TW = TW_REGS>: one warp per problem, streaming phases staged through registers (no arena)

// Warp-uniform values are made PROVABLY uniform for the compiler by reading them from lane 0: loops and branches on
// them then compile to plain uniform control flow (no BSSY/BSYNC reconvergence bookkeeping, no BRA.DIV in front of
// every shuffle). Measured on the SASS of the sweeps: 20 -> 10 instructions per pivot step.
// Lane-strided loop with a UNIFORM trip count (the bound must be uniform): a loop whose trip count differs per lane
// makes ptxas treat everything downstream as possibly diverged.
#define LANE_LOOP(i, lo, hi) _Pragma("unroll 1") for (int i##_base = (lo); i##_base < (hi); i##_base += 32) if (const int i = i##_base + lane; i < (hi))
__device__ __forceinline__ int uni(int v) { return __reduce_max_sync(FULL, v); }       // REDUX: lands in a uniform register
__device__ __forceinline__ bool uni(bool v) { return __any_sync(FULL, v); }               // VOTEU: uniform predicate
__device__ __forceinline__ double uni(double v) { return __shfl_sync(FULL, v, 0); }
__device__ __forceinline__ float uni(float v) { return __shfl_sync(FULL, v, 0); }

template <typename T>
struct LdpArgs {
    int P, n, m, ms, ldm, ldn, cap;
    // shared-memory layout of one warp, in elements of T from the warp's base (host-computed, see ldp_layout)
    int oD, olamA, olamB, oxl, ozl, odact, ou, oWS /* in ints */, osense /* in bytes */, ocnt /* in ints */;
    int ou32; /* float copy of u for the screening scan, in floats from the warp's base */
    int oarena; /* byte offset of the cp.async staging arena: scan ring, or two chunks of RB active rows */
    unsigned rowbuf; /* bytes of one chunk buffer of the row passes */
    unsigned per_warp_bytes;
    // per-problem byte strides of the global arrays
    unsigned sMt, sMr, sVec, sRinv, sv, sMt32;
    const T* Mt;              // [P][n][ldm]
    const T* Mr;              // [P][m][ldn]
    const float* Mt32;        // [P][ceil(n/4)][m][4] fp32 copy of M in quad layout for the screening scan (nullptr: scan in T)
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
    T* soft_slack;            // [P] or nullptr
    // persistent-workspace mode (daqp_b200_workspace_*): the warp's shared-memory block (factor, multipliers, working
    // set, sense) is saved at exit and restored at the start of the next solve, like the reference's DAQPWorkspace keeps
    // its LDL' factor and working set between daqp_solve calls (src/api.c:214-260)
    char* state;              // [P][state_stride] or nullptr
    unsigned state_stride;    // bytes per problem: oarena rounded up to 16, plus 16 for {k, lsw, valid}
    int state_load, state_save;
    int ns_max;               // most soft constraints (sense & 8) any problem of the batch carries; cap = n + ns_max + 1
    // shared-matrix mode (EXT instantiation only): `grp` consecutive problems are LDPs over ONE constraint matrix --
    // Mt / Mr / Mt32 / scaling / Rinv are indexed by p / grp; bounds, sense, v and the saved state stay per problem. Used
    // by daqp_b200_minrep_* (the m LDPs of a polyhedron, reference daqp_minrep_work, src/utils.c:808-835) and by shared
    // workspaces (one H / A, many f / b: daqp_b200_workspace_setup_shared). The group then streams one matrix out of L2.
    int grp;                  // 0 or 1: every problem owns its matrices
    int* exitflag;            // [P]
    int* iter;                // [P]
    int* ws_out;              // [P][cap] or nullptr : final working set (factor order)
    int* nact_out;            // [P] or nullptr
    int aux;                  // the launch needs the decision log or the time limit: instantiations that carry them (AUX below)
    int* trace_out;           // [P][1 + 2 trace_cap] or nullptr: count, then (code, value) per working-set decision -- 1 add
    int trace_cap;            //   (2 row + lower), 2 remove (row), 3 refactor, 4 refine, 5 cycle repair, 7 exit (flag). Debugging aid.
    int* counts_out;          // [P][8] or nullptr  : scans, adds, removes, csp solves, pivot swaps, refinements, refactors, cycle repairs
    unsigned char* sense_out; // [P][ldm] or nullptr: final sense bits
    int* work_counter;        // dynamic problem queue
    int tune;                 // experiment knob (see daqp_b200.cu)
    // team mode (n > 64): one CTA of `team` warps per problem; the CTA's shared memory starts with a TeamBox and the
    // team's scratch vectors (team_prefix bytes), then the per-problem block laid out as for a single warp (no arena)
    int team;                 // 0 = one warp per problem
    unsigned team_prefix, otmp, opv, obv, oside; // byte offsets from the start of the CTA's shared memory
    int* pst_id;              // [total warps][cap] pivot stack (rare path)
    T* pst_lam;               // [total warps][cap]
    DevSettings<T> st;
};

// Fills the layout fields of LdpArgs; returns the bytes of shared memory one warp needs.
template <typename T>
inline size_t ldp_layout(LdpArgs<T>& a, int team = 0) {
    const int V = VecOf<T>::N, cap = a.cap;
    a.team = team; // (-1: one warp per problem without the staging arena, the register-staged kernel)
    a.team_prefix = 0;
    if (team > 1) { // box | tmp[32 team] | pv[32 team] | bv[32 team] | side[team-1][32]
        const unsigned nt = 32u * team, w = (unsigned)sizeof(T);
        a.otmp = TEAM_BOX_BYTES; a.opv = a.otmp + nt * w; a.obv = a.opv + nt * w; a.oside = a.obv + nt * w;
        a.team_prefix = (a.oside + (unsigned)(team - 1) * 32u * w + 15u) / 16u * 16u;
    }
    int o = loff(cap);
    a.oD = o; o += cap;
    a.olamA = o; o += cap;
    a.olamB = o; o += cap;
    a.oxl = o; o += cap;
    a.ozl = o; o += cap;
    a.odact = o; o += cap;
    a.ou = o; o += round_up(a.n, V) + U_PAD;
    size_t bytes = (size_t)o * sizeof(T);
    bytes = (bytes + 15) / 16 * 16; // WS is read four indices at a time; padded so that a whole chunk is addressable
    a.oWS = (int)(bytes / sizeof(int)); bytes += (size_t)round_up(cap, RB) * sizeof(int);
    a.ocnt = (int)(bytes / sizeof(int)); bytes += 8 * sizeof(int);
    a.osense = (int)bytes; bytes += (size_t)round_up(a.m, 4);
    bytes = (bytes + 15) / 16 * 16; // the screening scan reads u32 with 128-bit loads
    a.ou32 = (int)(bytes / sizeof(float)); bytes += (size_t)(round_up(a.n, 4) + U_PAD) * sizeof(float);
    bytes = (bytes + 15) / 16 * 16;
    // staging arena: QRING quads of the fp32 matrix (screening scan), or two chunk buffers of RB rows of the row-major
    // matrix (the buffer of a consumed chunk doubles as the scratch of the dot-product reduction)
    a.oarena = (int)bytes;
    const size_t slab = (size_t)a.m * 16, rowb = (size_t)a.ldn * sizeof(T);
    a.rowbuf = (unsigned)std::max<size_t>(RB * rowb, (size_t)RB * SCR_PITCH * sizeof(T));
    if (team == 0 || team == 1) bytes += std::max<size_t>(QRING * (a.m <= 256 ? slab : (size_t)2048), 2 * (size_t)a.rowbuf); // (m > 256: blocks of 128 rows)
    bytes = (bytes + 15) / 16 * 16; // (a team stages through registers: no arena)
    a.per_warp_bytes = (unsigned)bytes;
    a.sMt = (unsigned)((size_t)a.n * a.ldm * sizeof(T));
    a.sMr = (unsigned)((size_t)a.m * a.ldn * sizeof(T));
    a.sVec = (unsigned)((size_t)a.ldm * sizeof(T));
    a.sRinv = (unsigned)((size_t)a.n * (a.n + 1) / 2 * sizeof(T));
    a.sv = (unsigned)((size_t)a.n * sizeof(T));
    a.sMt32 = (unsigned)((size_t)((a.n + 3) / 4) * a.m * 16);
    return bytes + a.team_prefix;
}

// 128-bit global load, predicated inside the instruction (no branch): out keeps its value when pred is false.
template <typename T> __device__ __forceinline__ void ldg_vec_pred(const void* p, T (&out)[VecOf<T>::N], uint64_t pol, int pred);
template <> __device__ __forceinline__ void ldg_vec_pred<double>(const void* p, double (&out)[2], uint64_t pol, int pred) {
    asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %4, 0;\n @q ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;\n}"
                 : "+d"(out[0]), "+d"(out[1]) : "l"(p), "l"(pol), "r"(pred));
}
template <> __device__ __forceinline__ void ldg_vec_pred<float>(const void* p, float (&out)[4], uint64_t pol, int pred) {
    asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %6, 0;\n @q ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;\n}"
                 : "+f"(out[0]), "+f"(out[1]), "+f"(out[2]), "+f"(out[3]) : "l"(p), "l"(pol), "r"(pred));
}

// One chunk of RB active rows, global -> shared (cp.async, one commit group). ONE copy of this code serves both row
// passes (not inlined: the instruction cache is the scarce resource, see the header). ids = &WS[base]; rows at or
// beyond nrows are not fetched (their ids are stale but never dereferenced); okmask bit g = this lane owns a 16-byte
// slice in column group g.
template <typename T, int NG>
__device__ __noinline__ void issue_rows_fn(const int* ids, int nrows, unsigned buf, const char* M, unsigned rstride, unsigned okmask) {
    int id[RB];
    if constexpr (RB % 4 == 0) { // chunk starts are multiples of RB: 16-byte aligned index quads
        const int4* wsv = reinterpret_cast<const int4*>(ids);
#pragma unroll
        for (int v4 = 0; v4 < RB / 4; v4++) { const int4 t = wsv[v4]; id[4 * v4] = t.x; id[4 * v4 + 1] = t.y; id[4 * v4 + 2] = t.z; id[4 * v4 + 3] = t.w; }
    } else {
#pragma unroll
        for (int r = 0; r < RB; r++) id[r] = ids[r];
    }
#pragma unroll
    for (int r = 0; r < RB; r++) {
        const char* src = M + (size_t)(unsigned)id[r] * rstride;
#pragma unroll
        for (int g = 0; g < NG; g++) cp_async16_if(((okmask >> g) & 1u) && r < nrows, buf + r * rstride + 512 * g, src + 512 * g);
    }
    cp_async_commit();
}

// EXT = the extended feature set (soft constraints, persistent-workspace state): a separate instantiation, so that the
// plain path -- the one the headline benchmark runs -- does not carry that code through its instruction cache.
// TW > 1 = team mode: TW warps (one CTA) share one problem. Warp 0 (the leader) runs the same state machine as below and
// posts the heavy phases -- the triangular sweeps, the C1 downdate, the dot products of an LDL add, the primal update,
// the feasibility scan -- as commands that all TW warps execute together; warps 1.. sit in helper_loop().
template <typename T, int NV, bool EXT, int TW = 1>
struct Warp {
    static_assert(!(EXT && TW > 1), "team mode covers the plain path");
    static_assert(32 * TW >= 32 * NV || TW <= 1, "a team holds one factor row per thread");
    static constexpr int V = VecOf<T>::N;
    static constexpr int NG = (NV + 1) / 2;                         // 128-bit column groups of a row: n <= 32 NV - 1
    static constexpr int ROWB = (NG == 1) ? 8 : (NG == 2 ? 4 : 2);  // active rows fetched per batch
    enum { OP_ADD = 0, OP_REMOVE = 1 };

    const LdpArgs<T>& a; // kernel parameters (constant bank): sizes, layout offsets, array bases
    T* S;                // this warp's shared memory
    int lane, p;         // lane id, problem index
    int pm;              // EXT: index of the problem's matrix set (p / grp in shared-matrix mode, else p)
    int k, reuse, sing;  // warp-uniform solver state (n_active, reuse_ind, sing_ind)
    int lsw;             // which of the two lambda buffers currently is `lam` (the reference swaps pointers)
    T fval, soft_slack;
    int* pst_id; T* pst_lam;
    // daqp_ldp loop state (daqp.c:7-10), kept across step() calls
    int iter, tried_repair, cycle_counter;
    bool do_activate;
    T best_fval;

    __device__ __forceinline__ Warp(const LdpArgs<T>& args) : a(args) {}

    // ---- array accessors: base + constant-bank offset
    __device__ __forceinline__ T* L() const { return S; }
    __device__ __forceinline__ T* D() const { return S + a.oD; }
    __device__ __forceinline__ T* lam() const { return S + (lsw ? a.olamB : a.olamA); }
    __device__ __forceinline__ T* lams() const { return S + (lsw ? a.olamA : a.olamB); }
    __device__ __forceinline__ T* xl() const { return S + a.oxl; }
    __device__ __forceinline__ T* zl() const { return S + a.ozl; }
    __device__ __forceinline__ T* dact() const { return S + a.odact; }
    __device__ __forceinline__ T* u() const { return S + a.ou; }
    __device__ __forceinline__ float* u32() const { return reinterpret_cast<float*>(S) + a.ou32; }
    __device__ __forceinline__ int* WS() const { return reinterpret_cast<int*>(S) + a.oWS; }
    __device__ __forceinline__ int* cnt() const { return reinterpret_cast<int*>(S) + a.ocnt; }
    __device__ __forceinline__ unsigned char* sense() const { return reinterpret_cast<unsigned char*>(S) + a.osense; }
    __device__ __forceinline__ int pmat() const { if constexpr (EXT) return pm; else return p; }
    __device__ __forceinline__ const char* Mt() const { return reinterpret_cast<const char*>(a.Mt) + (size_t)pmat() * a.sMt; }
    __device__ __forceinline__ const char* Mr() const { return reinterpret_cast<const char*>(a.Mr) + (size_t)pmat() * a.sMr; }
    __device__ __forceinline__ const T* du() const { return reinterpret_cast<const T*>(reinterpret_cast<const char*>(a.dupper) + (size_t)p * a.sVec); }
    __device__ __forceinline__ const T* dl() const { return reinterpret_cast<const T*>(reinterpret_cast<const char*>(a.dlower) + (size_t)p * a.sVec); }
    __device__ __forceinline__ const T* sc() const { return reinterpret_cast<const T*>(reinterpret_cast<const char*>(a.scaling) + (size_t)pmat() * a.sVec); }
#ifdef DAQP_B200_PHASE_CLOCKS
    __device__ __forceinline__ void count(int) {}
#else
    __device__ __forceinline__ void count(int which) { if (lane == 0) cnt()[which]++; }
#endif
    // decision log (only when the caller asked for it: the cursor lives in the log itself, no register or shared memory)
    // -DDAQP_B200_PHASE_CLOCKS (profiling build only): the eight path counters are replaced by the leader's clock cycles
    // / 16 per phase -- 0 CSP or singular direction, 1 ratio test, 2 primal, 3 scan, 4 LDL add, 5 LDL remove, 7 activation
#ifdef DAQP_B200_PHASE_CLOCKS
#define PHASE_T0 const long long ph_t0 = clock64()
#define PHASE_T1(w) do { if (lane == 0) cnt()[w] += (int)((clock64() - ph_t0) >> 4); } while (0)
#else
#define PHASE_T0 do { } while (0)
#define PHASE_T1(w) do { } while (0)
#endif
    // (the decision log and the time limit are compiled into the extended and the team instantiations only: the plain
    // warp-per-problem kernel is bound by its instruction-cache footprint, and even dead code between hot blocks costs)
    static constexpr bool AUX = EXT || TW > 1;
    __device__ __forceinline__ void trace(int code, int val) {
        if constexpr (!AUX) return;
        if (a.trace_out != nullptr && lane == 0) {
            int* tr = a.trace_out + (size_t)p * (1 + 2 * a.trace_cap);
            const int c = tr[0];
            if (c < a.trace_cap) { tr[1 + 2 * c] = code; tr[2 + 2 * c] = val; }
            tr[0] = c + 1;
        }
    }

    __device__ __forceinline__ void reset() { sing = EMPTY_IND; k = 0; reuse = 0; } // daqp.c:142-146

    // ---- triangular sweeps. The vector lives in NV registers per lane (element i = lane + 32 q in x[q]); one pivot
    // per step: shuffle-broadcast the pivot value, one predicated load + FMA per register segment. No shared-memory
    // round trip and no __syncwarp inside the recurrence. Written as plain predicated C++ on PROVABLY uniform bounds
    // (see uni()): ptxas then emits, per pivot, 2 SHFL + per segment {ISETP, @p LDS.64, @p DFMA} and unrolls by four
    // (10 instructions per pivot for two segments, 7 for one).
    __device__ __forceinline__ void vload(T (&x)[NV], const T* src, int len) {
#pragma unroll
        for (int q = 0; q < NV; q++) { const int i = lane + 32 * q; x[q] = (i < len) ? src[i] : (T)0; }
    }
    __device__ __forceinline__ void vstore(const T (&x)[NV], T* dst, int lo, int len) {
#pragma unroll
        for (int q = 0; q < NV; q++) { const int i = lane + 32 * q; if (i >= lo && i < len) dst[i] = x[q]; }
    }
    // pivots [jbeg, jend) of register segment QP applied to row segments QP .. Q1-1
    template <int QP, int Q1>
    __device__ __forceinline__ void fwd_pivots(T (&x)[NV], const T* const (&row)[NV], const int (&lim)[NV], int jbeg, int jend) {
#pragma unroll 1 // (unroll 4: 120 ms, 2: 112 ms, 1: 109.5 ms -- the instruction cache decides, see the header)
        for (int j = jbeg; j < jend; j++) {
            const T xj = __shfl_sync(FULL, x[QP], j);
#pragma unroll
            for (int q = QP; q < Q1; q++)
                if (j < lim[q]) x[q] -= row[q][j] * xj;
        }
    }
    // x <- L^-1 x restricted to rows >= rlo (rows < rlo already hold the solution); ascending pivots, the order of
    // the reference's forward substitutions (factorization.c:86-92, auxiliary.c:334-337). rlo and len are uniform.
    __device__ __forceinline__ void forward_sweep(T (&x)[NV], int rlo, int len) {
        if constexpr (TW > 1) { // the vector goes through the team's scratch: thread i of the CTA owns element i
            vstore(x, team_tmp(), 0, len);
            team_run(TC_FWD, rlo, len);
            vload(x, team_tmp(), len);
        } else {
        const T* row[NV]; // &L[i][0] for this lane's rows
        int lim[NV];      // row index, or -1 for rows that must not be touched
#pragma unroll
        for (int q = 0; q < NV; q++) {
            const int i = lane + 32 * q;
            lim[q] = (i >= rlo && i < len) ? i : -1;
            row[q] = L() + loff(i);
        }
        fwd_from<0>(x, row, lim, len);
        }
    }
    template <int QP>
    __device__ __forceinline__ void fwd_from(T (&x)[NV], const T* const (&row)[NV], const int (&lim)[NV], int len) {
        const int jbeg = 32 * QP, jend = min(len - 1, 32 * QP + 32);
        if (jbeg >= jend) return;
        if constexpr (QP + 1 == NV) fwd_pivots<QP, NV>(x, row, lim, jbeg, jend);
        else {
            if (len <= 32 * QP + 32) fwd_pivots<QP, QP + 1>(x, row, lim, jbeg, jend); // no rows beyond this segment
            else { fwd_pivots<QP, NV>(x, row, lim, jbeg, jend); fwd_from<QP + 1>(x, row, lim, len); }
        }
    }
    // x <- L^-T x ; descending pivots (auxiliary.c:343-352, 363-370). Row j of the packed factor is contiguous.
    __device__ __forceinline__ void backward_sweep(T (&x)[NV], int len) {
        if constexpr (TW > 1) {
            vstore(x, team_tmp(), 0, len);
            team_run(TC_BWD, len, 0);
            vload(x, team_tmp(), len);
        } else
#pragma unroll
        for (int qp = NV - 1; qp >= 0; qp--) {
            const int jlo = max(1, 32 * qp), jhi = min(len - 1, 32 * qp + 31);
            // (running row pointer: with the offset recomputed from j, ptxas folded the counter update into a
            // lane-predicated move, declared the loop divergent and guarded every shuffle of the kernel behind it)
            const T* Lj = L() + loff(jhi) + lane;
#pragma unroll 1
            for (int j = jhi; j >= jlo; j--) {
                const T xj = __shfl_sync(FULL, x[qp], j);
#pragma unroll
                for (int q = 0; q <= qp; q++)
                    if (q < qp || lane + 32 * q < j) x[q] -= Lj[32 * q] * xj; // rows of lower segments are all above the pivot
                Lj -= j - 1; // loff(j-1) = loff(j) - (j-1)
            }
        }
    }

    // ---- active-row passes (LDL' add, primal update): rows WS[0..kk) of the row-major matrix stream through two
    // chunk buffers of RB rows (cp.async, one commit group per chunk, the next chunk in flight while this one is
    // consumed). Lane l copies, and later reads back, only its own 16-byte slices of a row, so the buffers need no
    // barrier: cp.async.wait_group is a per-thread wait. Everything that steers the loops is uniform, the chunk body
    // is fully unrolled with static slots: per row ~3 instructions to issue and ~5 to consume.
    // M = this lane's slice of row 0; rows >= kk are not fetched (their slots keep stale data that is never used).
    __device__ __forceinline__ void issue_rows(int base, int kk, unsigned buf, const char* M, const bool (&okg)[NG]) const {
        unsigned okmask = 0;
#pragma unroll
        for (int g = 0; g < NG; g++) okmask |= (unsigned)okg[g] << g;
        issue_rows_fn<T, NG>(WS() + base, kk - base, buf, M, a.ldn * (unsigned)sizeof(T), okmask);
    }

    // ---- a2: LDL' row append (factorization.c:21-111) + bookkeeping of daqp_add_constraint (auxiliary.c:27-41)
    __device__ __forceinline__ void raw_add(int add, T lamval) {
        count(1);
        const int sb = sense()[add];
        __syncwarp();
        if (lane == 0) sense()[add] = (unsigned char)(sb | B_ACTIVE);
        sing = EMPTY_IND;
        const T bound = (sb & B_LOWER) ? dl()[add] : du()[add]; // right-hand side of this row in every later CSP
        const char* M = Mr() + (size_t)(V * lane) * sizeof(T);   // this lane's column slice of row 0
        const unsigned rstride = a.ldn * (unsigned)sizeof(T);
        T mi[NG][V];
        bool okg[NG];
        T part = 0;
#pragma unroll
        for (int g = 0; g < NG; g++) {
            okg[g] = V * (lane + 32 * g) < a.ldn;
#pragma unroll
            for (int e = 0; e < V; e++) mi[g][e] = 0;
            if (okg[g]) ldg_vec<T>(reinterpret_cast<const T*>(M + (size_t)add * rstride) + 32 * V * g, mi[g]);
#pragma unroll
            for (int e = 0; e < V; e++) part += mi[g][e] * mi[g][e];
        }
        const int kk = uni(k);
        T* Lk = L() + loff(kk);
        const unsigned buf0 = smem_u32(S) + a.oarena;
        T d = warp_sum(part);
        int ns_active = 0; // soft constraints in the working set, the entering one included (factorization.c:48-52,60-62)
        if (EXT && a.ns_max > 0) {
            if (sb & B_SOFT) { d += a.st.rho_soft; ns_active = 1; }
            const unsigned char* se = sense();
            const int* wsp = WS();
            for (int base = 0; base < kk; base += 32) {
                const int i = base + lane;
                ns_active += __popc(__ballot_sync(FULL, i < kk && (se[wsp[min(i, kk - 1)]] & B_SOFT)));
            }
        }
        if (kk > 0) {
            if constexpr (TW > 1) team_run(TC_DOTS, add, kk);
            if constexpr (TW == TW_REGS) dots_regs(kk, mi, okg);
            // l_j = M_{WS[j]} . m_add: per-lane partial products of a chunk's rows are parked in the consumed buffer and
            // summed with a transposed read (lane = (row, quarter)) instead of RB full shuffle reductions.
            if constexpr (TW == 1) {
            issue_rows(0, kk, buf0 + 16 * lane, M, okg);
            for (int c0 = 0; c0 < kk; c0 += RB) {
                const unsigned cur = buf0 + ((c0 / RB) & 1) * a.rowbuf, nxt = buf0 + (((c0 / RB) & 1) ^ 1) * a.rowbuf;
                if (c0 + RB < kk) issue_rows(c0 + RB, kk, nxt + 16 * lane, M, okg); else cp_async_commit();
                cp_async_wait<1>();
                T pj[RB];
#pragma unroll
                for (int r = 0; r < RB; r++) {
                    pj[r] = 0;
#pragma unroll
                    for (int g = 0; g < NG; g++) {
                        if (okg[g]) { // a lane without columns in this group has no slice in the slot
                            T t[V];
                            lds_vec<T>(cur + 16 * lane + r * rstride + 512 * g, t);
#pragma unroll
                            for (int e = 0; e < V; e++) pj[r] += t[e] * mi[g][e];
                        }
                    }
                }
                __syncwarp(); // every lane has read its slices: the consumed buffer becomes the reduction scratch
                T* scr = reinterpret_cast<T*>(reinterpret_cast<char*>(S) + a.oarena + ((c0 / RB) & 1) * a.rowbuf);
#pragma unroll
                for (int r = 0; r < RB; r++) scr[r * SCR_PITCH + lane] = pj[r];
                __syncwarp();
                const int r8 = lane >> 2, q4 = lane & 3;
                const T* pr = scr + min(r8, RB - 1) * SCR_PITCH + 8 * q4; // lanes beyond the chunk's rows re-read the last one
                T sum = ((pr[0] + pr[1]) + (pr[2] + pr[3])) + ((pr[4] + pr[5]) + (pr[6] + pr[7]));
                sum += __shfl_xor_sync(FULL, sum, 1);
                sum += __shfl_xor_sync(FULL, sum, 2);
                const int jr = c0 + r8;
                if (q4 == 0 && r8 < RB && jr < kk) Lk[jr] = sum;
                __syncwarp(); // scratch reads are done before the buffer is refilled, results visible to the sweep
            }
            }
            // l <- L^-1 l in registers, then l <- D^-1 l ; d -= l' D l
            T lv[NV];
            vload(lv, Lk, kk);
            forward_sweep(lv, 0, kk);
            const T* Dp = D();
            T acc = 0;
#pragma unroll
            for (int q = 0; q < NV; q++) {
                if (q > 0 && 32 * q >= kk) break; // uniform: no rows in this register segment
                const int i = lane + 32 * q;
                const T t = lv[q];
                const T qd = fdiv(t, i < kk ? Dp[i] : (T)1); // unconditional call: no divergence around the division
                if (i < kk) { Lk[i] = qd; acc += t * qd; }
            }
            d -= warp_sum(acc);
            if (uni(d < a.st.sing_tol || kk >= a.n + ns_active)) {
                sing = kk;
                d = 0;
            }
        }
        if (lane == 0) { D()[kk] = d; WS()[kk] = add; lam()[kk] = lamval; dact()[kk] = bound; }
        k = kk + 1;
        __syncwarp();
    }

    // ---- a3: LDL' row/column deletion + Gill-Golub-Murray-Saunders C1 update (factorization.c:112-151)
    //      + bookkeeping of daqp_remove_constraint (auxiliary.c:3-22). Returns 1 if the factor became singular.
    // Lane = trailing row: the removed column (the reference's w = &zldl[rm_ind]) lives in registers, every lane walks
    // along its own row, and the update writes each element straight to its compacted position (one row up, one
    // column left), so only the columns left of the removed one need a copy pass. Same recurrences, same order.
    __device__ __forceinline__ int raw_remove(int r) {
        count(2);
        const int kk = uni(k);
        T* Lp = L();
        T* Dp = D();
        int* ws = WS();
        if (lane == 0) sense()[ws[r]] &= ~B_ACTIVE;
        if constexpr (TW > 1) {
            if (r != kk - 1) team_run(TC_REMOVE, r, kk);
        } else
        if (r != kk - 1) {
            const int nu = kk - r - 1; // trailing rows; trailing row s is old row r+1+s and becomes row r+s
            T w[NV];
            const T* src[NV];
            T* dst[NV];
            int srow[NV];
#pragma unroll
            for (int q = 0; q < NV; q++) {
                const int sidx = lane + 32 * q, io = min(r + 1 + sidx, a.cap - 1);
                srow[q] = sidx < nu ? sidx : -1;
                src[q] = Lp + loff(io) + r + 1;    // old element (s, t) = src[t]
                dst[q] = Lp + loff(io - 1) + r;    // its compacted position = dst[t]
                w[q] = sidx < nu ? src[q][-1] : (T)0; // removed column
            }
            __syncwarp();
            // columns left of the removed one: row i moves up by one (lane j handles column j of every row, so each
            // destination was read by the same lane one step earlier)
            if (r > 0) {
                const T* from = Lp + loff(r + 1);
                for (int i = r + 1; i < kk; i++) {
                    LANE_LOOP(j, 0, r) const_cast<T*>(from)[j - (i - 1)] = from[j];
                    from += i;
                }
            }
            __syncwarp();
            T alpha = Dp[r];
            __syncwarp(); // every lane holds D[r] before lane 0 overwrites it in the first step below
#pragma unroll
            for (int qp = 0; qp < NV; qp++) { // pivot = trailing row t, held in register segment qp
                const int tend = min(nu, 32 * qp + 32);
                for (int t = 32 * qp; t < tend; t++) {
                    const T pvt = __shfl_sync(FULL, w[qp], t);
                    const T Dold = Dp[r + 1 + t];
                    const T dbar = Dold + alpha * pvt * pvt;
                    const T rdb = frcp(dbar); // one reciprocal (= 1 / dbar, correctly rounded) for the two quotients of the reference
                    const T beta = pvt * alpha * rdb;
                    alpha = Dold * alpha * rdb;
                    if (lane == 0) Dp[r + t] = dbar; // its old value was consumed one step earlier (as alpha for t = 0)
#pragma unroll
                    for (int q = qp; q < NV; q++) {
                        if (t < srow[q]) {
                            const T lv = src[q][t];
                            const T qs = w[q] - pvt * lv;
                            w[q] = qs;
                            dst[q][t] = lv + beta * qs;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        k = kk - 1;
        T* lm = lam();
        T* da = dact();
        for (int base = r; base < kk - 1; base += 32) { // shift WS / lam / active bounds down by one
            const int i = base + lane;
            int wv = 0; T lv = 0, dv = 0;
            if (i < kk - 1) { wv = ws[i + 1]; lv = lm[i + 1]; dv = da[i + 1]; }
            __syncwarp();
            if (i < kk - 1) { ws[i] = wv; lm[i] = lv; da[i] = dv; }
        }
        __syncwarp();
        if (r < reuse) reuse = r;
        if (kk > 1 && uni(Dp[kk - 2] < a.st.sing_tol)) {
            sing = kk - 2;
            __syncwarp();
            if (lane == 0) Dp[kk - 2] = 0;
            __syncwarp();
            return 1;
        }
        return 0;
    }

    // ---- daqp_add_constraint / daqp_remove_constraint (auxiliary.c:3-44) including the pivoting they trigger
    // (daqp_pivot_last, auxiliary.c:379-396). The reference recurses (pivot_last -> remove_constraint -> pivot_last
    // ... -> add_constraint -> pivot_last); here ONE loop executes the requested operation and every operation the
    // pivoting rule asks for, so raw_add / raw_remove are instantiated once. A pending "re-add the removed row once
    // the nested removal returns" is one entry of an explicit stack (global scratch: rare path).
    __device__ __forceinline__ void modify(int op, int arg, T lamval) {
        // (structured control flow on purpose: one back edge per loop, no `continue` -- anything else makes ptxas
        // give up on convergence analysis and guard every later shuffle with a divergence check)
        int depth = 0;
        for (;;) {
            bool pivot_check = true;
            if (op == OP_ADD) raw_add(arg, lamval);
            else if (uni(raw_remove(arg))) pivot_check = false; // removal made the factor singular: no pivoting
            bool next = false;
            const int kq = uni(k);
            if (pivot_check && kq > 1) {
                const int r = kq - 2;
                const T Dr = D()[r], Dl = D()[kq - 1];
                if (uni(Dr < a.st.pivot_tol && Dr < Dl)) {
                    count(4);
                    if (lane == 0) { pst_id[depth] = WS()[r]; pst_lam[depth] = lam()[r]; }
                    __syncwarp();
                    depth++;
                    op = OP_REMOVE; arg = r;
                    next = true;
                }
            }
            // the innermost pivot_last / remove_constraint returned: unwind pending re-adds
            while (!next && depth > 0) {
                depth--;
                if (uni(sing == EMPTY_IND)) { // else auxiliary.c:392: abort this frame, keep unwinding
                    op = OP_ADD; arg = uni(pst_id[depth]); lamval = uni(pst_lam[depth]);
                    next = true;
                }
            }
            if (!next) return;
        }
    }

    // ---- a5: constrained stationary point L D L' lam* = -d_k (auxiliary.c:314-354), forward solve resumes at reuse
    __device__ __forceinline__ void compute_csp() {
        count(3);
        const int kk = uni(k), r = uni(reuse);
        T* xp = xl();
        const T* da = dact();
        T xv[NV];
        if (kk - r <= 2) {
            // one or two new rows (the common case after an add): one warp-wide dot product per row
            for (int i = r; i < kk; i++) {
                const T* Li = L() + loff(i);
                T acc = 0;
#pragma unroll
                for (int q = 0; q < NV; q++) { const int j = lane + 32 * q; if (j < i) acc += Li[j] * xp[j]; }
                acc = warp_sum(acc);
                if (lane == 0) xp[i] = -da[i] - acc;
                __syncwarp();
            }
            vload(xv, xp, kk);
        } else {
#pragma unroll
            for (int q = 0; q < NV; q++) {
                const int i = lane + 32 * q;
                xv[q] = (i < kk) ? (i >= r ? -da[i] : xp[i]) : (T)0;
            }
            forward_sweep(xv, r, kk);
            vstore(xv, xp, r, kk);
        }
        // z = D^-1 x (kept in zldl for rows >= r as the reference does), then lam* <- L^-T z
        T* z = zl();
        const T* Dp = D();
#pragma unroll
        for (int q = 0; q < NV; q++) {
            if (q > 0 && 32 * q >= kk) { xv[q] = 0; continue; } // uniform: no rows in this register segment
            const int i = lane + 32 * q;
            const bool fresh = i >= r && i < kk;
            const T zi = fdiv(xv[q], fresh ? Dp[i] : (T)1); // unconditional call: no divergence around the division
            if (fresh) z[i] = zi;
            xv[q] = (i < kk) ? (fresh ? zi : z[i]) : (T)0;
        }
        backward_sweep(xv, kk);
        vstore(xv, lams(), 0, kk);
        __syncwarp();
        reuse = kk;
    }

    // ---- a9: singular direction (auxiliary.c:357-376)
    __device__ __forceinline__ void singular_direction() {
        const int s = sing;
        const T* Ls = L() + loff(s);
        T pv[NV];
#pragma unroll
        for (int q = 0; q < NV; q++) { const int i = lane + 32 * q; pv[q] = (i < s) ? -Ls[i] : (T)0; }
        backward_sweep(pv, s);
        const bool flip = sense()[WS()[s]] & B_LOWER;
        T* ls = lams();
#pragma unroll
        for (int q = 0; q < NV; q++) {
            const int i = lane + 32 * q;
            if (i <= s) { const T val = (i == s) ? (T)1 : pv[q]; ls[i] = flip ? -val : val; }
        }
        __syncwarp();
    }

    // ---- a6: dual ratio test + step (auxiliary.c:277-306). Returns the blocking position or -1.
    // Straight-line over the register segments: the division is executed by every lane (masked lanes divide by one),
    // so there is no divergent region and no loop with a per-lane trip count.
    __device__ __forceinline__ int find_blocking() {
        T best = (T)1e30;
        int key = INT_MAX;
        const T dual_tol = a.st.dual_tol;
        T* lm = lam();
        const T* ls = lams();
        const int* ws = WS();
        const unsigned char* se = sense();
        const int kk = uni(k);
        const bool nonsing = uni(sing == EMPTY_IND);
        T lv[NV], sv[NV];
#pragma unroll
        for (int q = 0; q < NV; q++) {
            lv[q] = 0; sv[q] = 0;
            if (q > 0 && 32 * q >= kk) continue; // uniform: no rows in this register segment
            const int i = lane + 32 * q;
            const bool valid = i < kk;
            const int sb = valid ? se[ws[i]] : B_IMMUTABLE;
            const T s = valid ? ls[i] : (T)0, l = valid ? lm[i] : (T)0;
            lv[q] = l; sv[q] = s;
            const bool cand = !(sb & B_IMMUTABLE) && ((sb & B_LOWER) ? !(s < dual_tol) : !(s > -dual_tol));
            const T ac = fdiv(-l, cand ? (nonsing ? s - l : s) : (T)1);
            if (cand && ac < best) { best = ac; key = i; }
        }
        warp_argmin(best, key);
        key = uni(key);
        if (key == INT_MAX) return -1;
        best = uni(best);
#pragma unroll
        for (int q = 0; q < NV; q++) {
            const int i = lane + 32 * q;
            if (i < kk) lm[i] = nonsing ? lv[q] + best * (sv[q] - lv[q]) : lv[q] + best * sv[q];
        }
        __syncwarp();
        sing = EMPTY_IND;
        return key;
    }

    // ---- a7: u = -Mk' lam*, fval = |u|^2 (auxiliary.c:46-88)
    __device__ __forceinline__ void compute_primal() {
        if constexpr (TW > 1) {
            const int kq = uni(k);
            team_run(TC_PRIMAL, kq, lsw);
            T s = 0;
#pragma unroll
            for (int w2 = 0; w2 < TW; w2++) s += (T)tbox()->rv[w2]; // fixed order: every run sums the same way
            fval = uni(s);
            __syncwarp();
            return;
        }
        if constexpr (TW == TW_REGS) { primal_regs(); return; }
        const char* M = Mr() + (size_t)(V * lane) * sizeof(T);
        const unsigned rstride = a.ldn * (unsigned)sizeof(T);
        const T* ls = lams();
        T acc[NG][V];
        bool okg[NG];
#pragma unroll
        for (int g = 0; g < NG; g++) {
            okg[g] = V * (lane + 32 * g) < a.ldn;
#pragma unroll
            for (int e = 0; e < V; e++) acc[g][e] = 0;
        }
        // FMAs in index order (the order of the reference's accumulation, auxiliary.c:54-68)
        const int kk = uni(k);
        const unsigned buf0 = smem_u32(S) + a.oarena + 16 * lane;
        // rotated loop: trip c issues chunk c+1 (or commits an empty group) and consumes chunk c, so the issue code exists once
        for (int c0 = -RB; c0 < kk; c0 += RB) {
            const unsigned cur = buf0 + ((c0 / RB) & 1) * a.rowbuf, nxt = buf0 + (((c0 / RB) & 1) ^ 1) * a.rowbuf;
            if (c0 + RB < kk) issue_rows(c0 + RB, kk, nxt, M, okg); else cp_async_commit();
            cp_async_wait<1>();
#pragma unroll
            for (int r = 0; r < RB; r++) {
                if (c0 >= 0 && c0 + r < kk) {
                    const T li = ls[c0 + r];
#pragma unroll
                    for (int g = 0; g < NG; g++) {
                        if (okg[g]) {
                            T t[V];
                            lds_vec<T>(cur + r * rstride + 512 * g, t);
#pragma unroll
                            for (int e = 0; e < V; e++) acc[g][e] -= t[e] * li;
                        }
                    }
                }
            }
        }
        T* up = u();
        T part = 0;
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const int c = V * (lane + 32 * g);
            if (c < a.ldn) {
#pragma unroll
                for (int e = 0; e < V; e++) { up[c + e] = acc[g][e]; u32()[c + e] = (float)acc[g][e]; part += acc[g][e] * acc[g][e]; }
            }
        }
        T slack = 0; // soft_slack = rho_soft * sum over soft active rows of lam*^2 (auxiliary.c:69-84)
        if (EXT && a.ns_max > 0) {
            const unsigned char* se = sense();
            const int* wsp = WS();
            T sp = 0;
            LANE_LOOP(i, 0, kk) { if (se[wsp[i]] & B_SOFT) { const T li = ls[i]; sp += li * li; } }
            slack = uni(warp_sum(sp)) * a.st.rho_soft;
        }
        if constexpr (EXT) { soft_slack = slack; fval = slack + uni(warp_sum(part)); }
        else fval = uni(warp_sum(part));
        __syncwarp();
    }

    // rows [base, base + SG*32*V): Mu_r = M_r . u with one accumulator per owned row, then the candidate test of
    // auxiliary.c:110-122,136-149 in ascending row order per lane.
    // Software pipeline over columns: a ring of UN column buffers; a slot is consumed (FMAs) and immediately refilled
    // with the column UN ahead, so SG*UN 128-bit loads stay in flight for the whole scan. The column cursor saturates
    // at the last column (those re-reads hit L1/L2 and are multiplied by the zero padding behind u), so the loop has
    // neither a prologue special case nor a remainder loop.
    template <int SG, int UN>
    __device__ __forceinline__ void scan_rows(int base, T& best, int& key) {
        constexpr int GR = 32 * V;
        const uint64_t pol = policy_evict_first();
        const int r0 = base + V * lane;
        T acc[SG][V];
        bool own[SG];
#pragma unroll
        for (int g = 0; g < SG; g++) {
            own[g] = r0 + g * GR < a.ldm;
#pragma unroll
            for (int e = 0; e < V; e++) acc[g][e] = 0;
        }
        T buf[UN][SG][V];
        const char* col0 = Mt() + (size_t)r0 * sizeof(T);
        const unsigned cstride = a.sVec, clast = (unsigned)(a.n - 1) * a.sVec;
        unsigned coff = 0; // byte offset of the column the next load fetches
#pragma unroll
        for (int i = 0; i < UN; i++) {
#pragma unroll
            for (int g = 0; g < SG; g++) {
#pragma unroll
                for (int e = 0; e < V; e++) buf[i][g][e] = 0;
                if (own[g]) ldg_vec_hint_ordered<T>(reinterpret_cast<const T*>(col0 + coff) + g * GR, buf[i][g], pol);
            }
            coff = min(coff + cstride, clast);
        }
        const T* up = u();
        for (int c = 0; c < a.n; c += UN) {
#pragma unroll
            for (int i = 0; i < UN; i++) {
                const T uc = up[c + i]; // zero for c + i >= n
                const T* cp = reinterpret_cast<const T*>(col0 + coff);
#pragma unroll
                for (int g = 0; g < SG; g++) {
#pragma unroll
                    for (int e = 0; e < V; e++) acc[g][e] += buf[i][g][e] * uc;
                    if (own[g]) ldg_vec_hint_ordered<T>(cp + g * GR, buf[i][g], pol);
                }
                coff = min(coff + cstride, clast);
            }
        }
        const T ep = -a.st.primal_tol;
        const uint64_t polk = policy_evict_last();
        const unsigned char* se = sense();
#pragma unroll
        for (int g = 0; g < SG; g++) {
            const int rr = r0 + g * GR;
            if (rr < a.m) { // ldm is a multiple of V: the whole vector is addressable
                T bu[V], bl[V], bs[V];
                ldg_vec_hint<T>(du() + rr, bu, polk);
                ldg_vec_hint<T>(dl() + rr, bl, polk);
                ldg_vec_hint<T>(sc() + rr, bs, polk);
#pragma unroll
                for (int e = 0; e < V; e++) {
                    const int row = rr + e;
                    if (row >= a.m) continue;
                    if (se[row] & (B_ACTIVE + B_IMMUTABLE)) continue;
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

    // ---- fp32 SCREENING of the feasibility scan (T = double only).
    // The scan only has to name the most violated row, not its value. A float copy of the matrix (half the bytes, and
    // small enough that the resident problems' copies fit L2) gives every slack with a rigorous error bound
    //   |s32 - s64| <= delta = 1.01 (n+3) 2^-24 |u|_2        (rows are normalised to unit 2-norm; Cauchy-Schwarz)
    // so: a row whose s32 - delta >= threshold cannot be a candidate; if exactly one candidate lies within 2 delta of
    // the smallest s32 and it is below the threshold by more than delta, it IS the fp64 argmin. Anything ambiguous
    // (near-ties, rows within delta of the threshold that could win) returns -2 and the caller runs the fp64 scan, so
    // the selected row is always the one the fp64 scan selects.
    template <int NR>
    __device__ __forceinline__ int scan_screen() {
        // Quad layout (setup kernel): element (row r, column c) of the float copy sits at ((c/4) m + r) * 4 + c%4, so
        // one 128-bit word is ONE row x FOUR columns and lane l owns rows l, l+32, ... (NR = ceil(m/32) of them).
        // Per quad of columns: one broadcast 128-bit read of u, and per owned row one 128-bit read + four FMAs into
        // that row's accumulator. Quads stream through the lane-private cp.async ring, QRING in flight.
        const bool own_last = lane + 32 * (NR - 1) < a.m; // rows of the groups before the last always exist
        float acc[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) acc[r] = 0.f;
        // the rows' bounds are needed only after the products: fetch them first so that their latency hides behind the scan
        double bu[NR], bl[NR], bs[NR];
        {
            const double* dup = reinterpret_cast<const double*>(du()) + lane;
            const double* dlp = reinterpret_cast<const double*>(dl()) + lane;
            const double* scp = reinterpret_cast<const double*>(sc()) + lane;
#pragma unroll
            for (int r = 0; r < NR; r++) {
                bu[r] = bl[r] = bs[r] = 0;
                if (r < NR - 1 || own_last) { bu[r] = __ldg(dup + 32 * r); bl[r] = __ldg(dlp + 32 * r); bs[r] = __ldg(scp + 32 * r); }
            }
        }
        const unsigned slab = (unsigned)a.m * 16u; // bytes of one quad of columns
        const char* src = reinterpret_cast<const char*>(a.Mt32) + (size_t)pmat() * a.sMt32 + 16 * lane;
        const unsigned ring0 = smem_u32(S) + a.oarena + 16 * lane;
        const int nq = (a.n + 3) >> 2;
        // Ring slots are static (the quad loop is unrolled QRING times) and the loop is rotated: the trip that consumes
        // quad q also issues quad q + QRING into the slot it has just read, so the first trip (q < 0) is the prologue and
        // the issue code exists once. Every step commits one group, possibly empty.
        const unsigned ub = smem_u32(u32());
        for (int q0 = -QRING; q0 < nq; q0 += QRING) {
#pragma unroll
            for (int sl = 0; sl < QRING; sl++) {
                const int q = q0 + sl;
                if (q >= 0 && q < nq) {
                    cp_async_wait<QRING - 1>();
                    float uq[4];
                    lds_vec<float>(ub + 16 * q, uq);
#pragma unroll
                    for (int r = 0; r < NR; r++) {
                        if (r < NR - 1 || own_last) {
                            float t[4];
                            lds_vec<float>(ring0 + sl * slab + 512 * r, t);
#pragma unroll
                            for (int e = 0; e < 4; e++) acc[r] += t[e] * uq[e];
                        }
                    }
                }
                if (q + QRING < nq) {
#pragma unroll
                    for (int r = 0; r < NR; r++) {
                        if (r < NR - 1) cp_async16(ring0 + sl * slab + 512 * r, src + 512 * r);
                        else cp_async16_if(own_last, ring0 + sl * slab + 512 * r, src + 512 * r);
                    }
                    src += slab;
                }
                cp_async_commit();
            }
        }
        cp_async_wait<0>();
        // candidates in double from the float products; track the best and the runner-up among "possible" candidates
        // |u|_2 from a float square root, rounded up (the bound only has to be an upper bound; + covers underflow)
        const double unorm = (double)sqrtf((float)fval) * 1.0001 + 1e-22;
        const double delta = 1.01 * (double)(a.n + 3) * 5.9604644775390625e-8 * unorm;
        const double ep = -(double)a.st.primal_tol;
        const unsigned char* se = sense();
        double best = 1e300, second = 1e300;
        int key = INT_MAX;
        bool best_sure = false;
#pragma unroll
        for (int r = 0; r < NR; r++) {
            const int row = lane + 32 * r;
            // at most one side of a row can be violated beyond the tolerance (check_bounds guarantees
            // bupper >= blower - tol), so the row's candidate is its more violated side
            const double mu = (double)acc[r];
            const double cu = bu[r] - mu, cl = mu - bl[r];
            const bool lower = cl < cu;
            const double cand = lower ? cl : cu;
            const double bound = ep * bs[r];
            const bool possible = (r < NR - 1 || own_last) && !(se[row] & (B_ACTIVE + B_IMMUTABLE)) && cand - delta < bound;
            const bool nb = possible && cand < best;
            second = nb ? best : ((possible && cand < second) ? cand : second);
            best_sure = nb ? (cand + delta < bound) : best_sure;
            key = nb ? 2 * row + (int)lower : key;
            best = nb ? cand : best;
        }
        // warp: argmin of (best, key); runner-up = smallest value that is not the winner
        double wbest = best;
        int wkey = key;
        warp_argmin(wbest, wkey);
        wkey = uni(wkey);
        if (wkey == INT_MAX) return -1; // no row can be violated beyond the tolerance: certain
        wbest = uni(wbest);
        double other = (key == wkey) ? second : best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) other = fmin(other, __shfl_xor_sync(FULL, other, o));
        const bool sure = __any_sync(FULL, key == wkey && best_sure);
        if (sure && uni(other - wbest > 2.0 * delta)) return wkey;
        return -2; // ambiguous: decide in fp64
    }

    // The same screening for m > 256: the rows go through the ring in BLOCKS of 128 (four row groups per lane, 2 KB per ring
    // slot -- the arena a single warp has anyway), the per-lane candidate record is carried across the blocks and the warp
    // decides once at the end. Same error bound, same decision rule, same fallback.
    __device__ __forceinline__ int scan_screen_blocks() {
        constexpr int NR = 4, BR = 32 * NR;
        const unsigned slab = (unsigned)a.m * 16u; // bytes of one quad of columns in global memory
        const int nq = (a.n + 3) >> 2;
        const unsigned ub = smem_u32(u32());
        const double unorm = (double)sqrtf((float)fval) * 1.0001 + 1e-22;
        const double delta = 1.01 * (double)(a.n + 3) * 5.9604644775390625e-8 * unorm;
        const double ep = -(double)a.st.primal_tol;
        const unsigned char* se = sense();
        double best = 1e300, second = 1e300;
        int key = INT_MAX;
        bool best_sure = false;
        for (int row0 = 0; row0 < a.m; row0 += BR) {
            float acc[NR];
            bool own[NR];
            double bu[NR], bl[NR], bs[NR];
#pragma unroll
            for (int r = 0; r < NR; r++) {
                const int row = row0 + lane + 32 * r;
                acc[r] = 0.f;
                own[r] = row < a.m;
                bu[r] = bl[r] = bs[r] = 0;
                if (own[r]) {
                    bu[r] = __ldg(reinterpret_cast<const double*>(du()) + row);
                    bl[r] = __ldg(reinterpret_cast<const double*>(dl()) + row);
                    bs[r] = __ldg(reinterpret_cast<const double*>(sc()) + row);
                }
            }
            const char* src = reinterpret_cast<const char*>(a.Mt32) + (size_t)pmat() * a.sMt32 + 16 * (size_t)(row0 + lane);
            const unsigned ring0 = smem_u32(S) + a.oarena + 16 * lane;
            for (int q0 = -QRING; q0 < nq; q0 += QRING) {
#pragma unroll
                for (int sl = 0; sl < QRING; sl++) {
                    const int q = q0 + sl;
                    if (q >= 0 && q < nq) {
                        cp_async_wait<QRING - 1>();
                        float uq[4];
                        lds_vec<float>(ub + 16 * q, uq);
#pragma unroll
                        for (int r = 0; r < NR; r++) {
                            if (own[r]) {
                                float t[4];
                                lds_vec<float>(ring0 + sl * (16u * BR) + 512 * r, t);
#pragma unroll
                                for (int e = 0; e < 4; e++) acc[r] += t[e] * uq[e];
                            }
                        }
                    }
                    if (q + QRING < nq) {
#pragma unroll
                        for (int r = 0; r < NR; r++) cp_async16_if(own[r], ring0 + sl * (16u * BR) + 512 * r, src + 512 * r);
                        src += slab;
                    }
                    cp_async_commit();
                }
            }
            cp_async_wait<0>();
#pragma unroll
            for (int r = 0; r < NR; r++) { // ascending rows per lane across the blocks: strict '<' keeps the first of equal candidates
                const int row = row0 + lane + 32 * r;
                const double mu = (double)acc[r];
                const double cu = bu[r] - mu, cl = mu - bl[r];
                const bool lower = cl < cu;
                const double cand = lower ? cl : cu;
                const double bound = ep * bs[r];
                const bool possible = own[r] && !(se[min(row, a.m - 1)] & (B_ACTIVE + B_IMMUTABLE)) && cand - delta < bound;
                const bool nb = possible && cand < best;
                second = nb ? best : ((possible && cand < second) ? cand : second);
                best_sure = nb ? (cand + delta < bound) : best_sure;
                key = nb ? 2 * row + (int)lower : key;
                best = nb ? cand : best;
            }
            __syncwarp(); // the ring is refilled by the next block
        }
        double wbest = best;
        int wkey = key;
        warp_argmin(wbest, wkey);
        wkey = uni(wkey);
        if (wkey == INT_MAX) return -1;
        wbest = uni(wbest);
        double other = (key == wkey) ? second : best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) other = fmin(other, __shfl_xor_sync(FULL, other, o));
        const bool sure = __any_sync(FULL, key == wkey && best_sure);
        if (sure && uni(other - wbest > 2.0 * delta)) return wkey;
        return -2;
    }

    // ---- a8: Mu = M u for all rows, most-violated inactive row (auxiliary.c:89-152).
    // Returns 2*row + (1 if violated at the lower bound), or -1 when every inactive row is feasible.
    __device__ __forceinline__ int scan_infeasible() {
        count(0);
        if constexpr (TW > 1) return team_scan_leader();
        if constexpr (sizeof(T) == 8 && TW == TW_REGS) {
            if (a.Mt32 != nullptr) {
                int r;
                switch ((a.m + 31) >> 5) {
                    case 1: r = scan_screen_regs<1>(); break;
                    case 2: r = scan_screen_regs<2>(); break;
                    case 3: r = scan_screen_regs<3>(); break;
                    case 4: r = scan_screen_regs<4>(); break;
                    case 5: r = scan_screen_regs<5>(); break;
                    case 6: r = scan_screen_regs<6>(); break;
                    case 7: r = scan_screen_regs<7>(); break;
                    default: r = scan_screen_regs<8>(); break;
                }
                if (r != -2) return r;
            }
        } else
        if constexpr (sizeof(T) == 8) {
            if (a.Mt32 != nullptr) { // screening in fp32 (the host enables it up to m = 768; beyond 256 rows in blocks of 128)
                int r;
                // (the block form is compiled into the EXTENDED instantiation only -- the host sends m > 256 there: inlined
                // into the plain kernel it cost the C3 headline 2 % in instruction-cache misses without ever running)
                bool blocks = false;
                if constexpr (EXT) blocks = a.m > 256;
                if (blocks) r = scan_screen_blocks();
                else switch ((a.m + 31) >> 5) {
                    case 1: r = scan_screen<1>(); break;
                    case 2: r = scan_screen<2>(); break;
                    case 3: r = scan_screen<3>(); break;
                    case 4: r = scan_screen<4>(); break;
                    case 5: r = scan_screen<5>(); break;
                    case 6: r = scan_screen<6>(); break;
                    case 7: r = scan_screen<7>(); break;
                    default: r = scan_screen<8>(); break;
                }
                if (r != -2) return r;
            }
        }
        constexpr int GR = 32 * V; // rows per group (one 128-bit load per lane)
        T best = 0;
        int key = INT_MAX;
        // widest pipeline the register file allows for this row count: SG loads per column, UN columns in flight
        if (a.ldm <= GR) scan_rows<1, 12>(0, best, key);
        else if (a.ldm <= 2 * GR) scan_rows<2, 6>(0, best, key);
        else if (a.ldm <= 3 * GR) scan_rows<3, 5>(0, best, key);
        else for (int base = 0; base < a.m; base += 4 * GR) scan_rows<4, 4>(base, best, key);
        warp_argmin(best, key);
        key = uni(key);
        return key == INT_MAX ? -1 : key;
    }

    // =====================================================================================================================
    // Register-staged variants (TW == TW_REGS): the same three streaming phases WITHOUT the cp.async staging arena. At
    // n = 50 the arena is 4.8 KB of a warp's 18.6 KB: without it sixteen problems fit an SM instead of twelve, and the
    // kernel is bound by dependent latency at three warps per scheduler, not by any unit's throughput (ncu: issue slots
    // 42 %, shared-memory pipe 51 %, L2 25 %, DRAM 3 % of peak). Loads go global/L2 -> registers in program order
    // (asm volatile), a batch in flight while the previous one is consumed.
    // =====================================================================================================================
    // l_j = M_{WS[j]} . m_add for j < kk into row kk of the factor: batches of DB rows, one transposed butterfly per batch
    __device__ __forceinline__ void dots_regs(int kk, const T (&mi)[NG][V], const bool (&okg)[NG]) {
        constexpr int DB = 8;
        const T* M = reinterpret_cast<const T*>(Mr()) + V * lane;
        const unsigned ldn = a.ldn;
        const int* ws = WS();
        T* Lk = L() + loff(kk);
        const uint64_t pol = policy_evict_last();
#pragma unroll 1
        for (int j0 = 0; j0 < kk; j0 += DB) {
            T t[DB][NG][V];
#pragma unroll
            for (int rr = 0; rr < DB; rr++) {
                const T* row = M + (size_t)((unsigned)ws[min(j0 + rr, kk - 1)] * ldn);
#pragma unroll
                for (int g = 0; g < NG; g++) {
#pragma unroll
                    for (int e = 0; e < V; e++) t[rr][g][e] = 0;
                    if (okg[g]) ldg_vec_hint_ordered<T>(row + 32 * V * g, t[rr][g], pol);
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
        __syncwarp();
    }

    // u = -Mk' lam*, fval = |u|^2 : FMAs in index order (auxiliary.c:54-68), UNR rows in flight
    __device__ __forceinline__ void primal_regs() {
        constexpr int UNR = 8;
        const T* M = reinterpret_cast<const T*>(Mr()) + V * lane;
        const unsigned ldn = a.ldn;
        const T* ls = lams();
        const int* ws = WS();
        const int kk = uni(k);
        const uint64_t pol = policy_evict_last();
        T acc[NG][V];
        bool okg[NG];
#pragma unroll
        for (int g = 0; g < NG; g++) {
            okg[g] = V * (lane + 32 * g) < a.ldn;
#pragma unroll
            for (int e = 0; e < V; e++) acc[g][e] = 0;
        }
#pragma unroll 1
        for (int i0 = 0; i0 < kk; i0 += UNR) {
            T t[UNR][NG][V];
#pragma unroll
            for (int rr = 0; rr < UNR; rr++) {
                const T* row = M + (size_t)((unsigned)ws[min(i0 + rr, kk - 1)] * ldn);
#pragma unroll
                for (int g = 0; g < NG; g++) {
#pragma unroll
                    for (int e = 0; e < V; e++) t[rr][g][e] = 0;
                    if (okg[g]) ldg_vec_hint_ordered<T>(row + 32 * V * g, t[rr][g], pol);
                }
            }
#pragma unroll
            for (int rr = 0; rr < UNR; rr++) {
                const T li = (i0 + rr < kk) ? ls[i0 + rr] : (T)0; // rows past the end re-read the last one, times zero
#pragma unroll
                for (int g = 0; g < NG; g++)
#pragma unroll
                    for (int e = 0; e < V; e++) acc[g][e] -= t[rr][g][e] * li;
            }
        }
        T* up = u();
        T part = 0;
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const int c = V * (lane + 32 * g);
            if (c < a.ldn) {
#pragma unroll
                for (int e = 0; e < V; e++) { up[c + e] = acc[g][e]; u32()[c + e] = (float)acc[g][e]; part += acc[g][e] * acc[g][e]; }
            }
        }
        fval = uni(warp_sum(part));
        __syncwarp();
    }

    // fp32 screening scan (see scan_screen for the error bound and the decision rule), quads of columns through two
    // register buffers per owned row instead of the shared-memory ring
    template <int NR>
    __device__ __forceinline__ int scan_screen_regs() {
        const bool own_last = lane + 32 * (NR - 1) < a.m;
        float acc[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) acc[r] = 0.f;
        const unsigned slab = (unsigned)a.m * 16u;
        const char* src = reinterpret_cast<const char*>(a.Mt32) + (size_t)pmat() * a.sMt32 + 16 * lane;
        const int nq = (a.n + 3) >> 2;
        const uint64_t pol = policy_evict_last();
        const unsigned ub = smem_u32(u32());
        float b0[NR][4], b1[NR][4];
#pragma unroll
        for (int r = 0; r < NR; r++) {
#pragma unroll
            for (int e = 0; e < 4; e++) { b0[r][e] = 0.f; b1[r][e] = 0.f; }
            ldg_vec_pred<float>(src + 512 * r, b0[r], pol, r < NR - 1 || own_last);
        }
#pragma unroll 1
        for (int q = 0; q < nq; q += 2) { // two quads per trip: static buffer names, the next quad in flight while this one is used
            const bool more1 = q + 1 < nq, more2 = q + 2 < nq;
            const char* s1 = src + (size_t)min(q + 1, nq - 1) * slab;
#pragma unroll
            for (int r = 0; r < NR; r++) ldg_vec_pred<float>(s1 + 512 * r, b1[r], pol, more1 && (r < NR - 1 || own_last));
            float uq[4];
            lds_vec<float>(ub + 16 * q, uq);
#pragma unroll
            for (int r = 0; r < NR; r++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[r] += b0[r][e] * uq[e];
            const char* s2 = src + (size_t)min(q + 2, nq - 1) * slab;
#pragma unroll
            for (int r = 0; r < NR; r++) ldg_vec_pred<float>(s2 + 512 * r, b0[r], pol, more2 && (r < NR - 1 || own_last));
            if (more1) {
                lds_vec<float>(ub + 16 * (q + 1), uq);
#pragma unroll
                for (int r = 0; r < NR; r++)
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[r] += b1[r][e] * uq[e];
            }
        }
        double bu[NR], bl[NR], bs[NR]; // (loaded after the products here: registers are the scarce resource)
        {
            const double* dup = reinterpret_cast<const double*>(du()) + lane;
            const double* dlp = reinterpret_cast<const double*>(dl()) + lane;
            const double* scp = reinterpret_cast<const double*>(sc()) + lane;
#pragma unroll
            for (int r = 0; r < NR; r++) {
                bu[r] = bl[r] = bs[r] = 0;
                if (r < NR - 1 || own_last) { bu[r] = __ldg(dup + 32 * r); bl[r] = __ldg(dlp + 32 * r); bs[r] = __ldg(scp + 32 * r); }
            }
        }
        const double unorm = (double)sqrtf((float)fval) * 1.0001 + 1e-22;
        const double delta = 1.01 * (double)(a.n + 3) * 5.9604644775390625e-8 * unorm;
        const double ep = -(double)a.st.primal_tol;
        const unsigned char* se = sense();
        double best = 1e300, second = 1e300;
        int key = INT_MAX;
        bool best_sure = false;
#pragma unroll
        for (int r = 0; r < NR; r++) {
            const int row = lane + 32 * r;
            const double mu = (double)acc[r];
            const double cu = bu[r] - mu, cl = mu - bl[r];
            const bool lower = cl < cu;
            const double cand = lower ? cl : cu;
            const double bound = ep * bs[r];
            const bool possible = (r < NR - 1 || own_last) && !(se[row] & (B_ACTIVE + B_IMMUTABLE)) && cand - delta < bound;
            const bool nb = possible && cand < best;
            second = nb ? best : ((possible && cand < second) ? cand : second);
            best_sure = nb ? (cand + delta < bound) : best_sure;
            key = nb ? 2 * row + (int)lower : key;
            best = nb ? cand : best;
        }
        double wbest = best;
        int wkey = key;
        warp_argmin(wbest, wkey);
        wkey = uni(wkey);
        if (wkey == INT_MAX) return -1;
        wbest = uni(wbest);
        double other = (key == wkey) ? second : best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) other = fmin(other, __shfl_xor_sync(FULL, other, o));
        const bool sure = __any_sync(FULL, key == wkey && best_sure);
        if (sure && uni(other - wbest > 2.0 * delta)) return wkey;
        return -2;
    }

    // ---- team mode (TW > 1), leader side: the heavy phases live in team_ops.cuh; the leader posts a command, runs its own
    // share of the phase (wid = 0) and meets the helpers at the closing barrier.
    static __device__ __forceinline__ TeamBox* tbox() {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        return reinterpret_cast<TeamBox*>(smem_raw);
    }
    __device__ __forceinline__ T* team_tmp() const {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        return reinterpret_cast<T*>(smem_raw + a.otmp);
    }
    __device__ __forceinline__ void team_run(int cmd, int a0, int a1) {
        if (lane == 0) { TeamBox* b = tbox(); b->cmd = cmd; b->a0 = a0; b->a1 = a1; }
        team_bar<TW>();
        team_dispatch<T, TW, NG>(cmd, lane, 0, a0, a1);
        team_bar<TW>();
    }
    __device__ __forceinline__ void team_exit() {
        if (lane == 0) tbox()->cmd = TC_EXIT;
        team_bar<TW>();
    }
    // once per launch: where the per-problem block lives (byte offsets from the start of the CTA's shared memory)
    __device__ __forceinline__ void team_publish_layout() {
        if (lane == 0) {
            TeamBox* b = tbox();
            const int w = (int)sizeof(T), o = (int)a.team_prefix;
            b->oL = o; b->oD = o + a.oD * w; b->olamA = o + a.olamA * w; b->olamB = o + a.olamB * w;
            b->oWS = o + a.oWS * 4; b->osense = o + a.osense; b->ou = o + a.ou * w; b->ou32 = o + a.ou32 * 4;
            b->otmp = (int)a.otmp; b->opv = (int)a.opv; b->obv = (int)a.obv; b->oside = (int)a.oside;
            b->cap = a.cap; b->n = a.n; b->m = a.m; b->ldm = a.ldm; b->ldn = a.ldn; b->tune = a.tune;
            b->primal_tol = (double)a.st.primal_tol; b->sing_tol = (double)a.st.sing_tol; b->pivot_tol = (double)a.st.pivot_tol;
        }
    }
    // once per problem: its global arrays (the helpers see them with the first command of the problem)
    __device__ __forceinline__ void team_publish_problem() {
        if (lane == 0) {
            TeamBox* b = tbox();
            b->p = p;
            b->Mr = Mr(); b->Mt = Mt(); b->du = du(); b->dl = dl(); b->sc = sc();
            b->Mt32 = a.Mt32 ? reinterpret_cast<const char*>(a.Mt32) + (size_t)p * a.sMt32 : nullptr;
        }
    }
    // leader: run the screening (then, if it cannot name the row, the exact scan) on the whole team and combine
    __device__ __forceinline__ int team_scan_leader() {
        TeamBox* box = tbox();
        if constexpr (sizeof(T) == 8) {
            if (a.Mt32 != nullptr) {
                if (lane == 0) box->fval = (double)fval;
                team_run(TC_SCAN32, 0, 0);
                double wb = 1e300, other = 1e300;
                int wk = INT_MAX, win = 0;
#pragma unroll
                for (int w2 = 0; w2 < TW; w2++) {
                    const double b = box->rv[w2];
                    const int kx = box->rk[w2];
                    if (kx != INT_MAX && (b < wb || (b == wb && kx < wk))) { wb = b; wk = kx; win = w2; }
                }
                wk = uni(wk);
                if (wk == INT_MAX) { __syncwarp(); return -1; }
#pragma unroll
                for (int w2 = 0; w2 < TW; w2++) // runner-up: the winner warp's own, or another warp's best (1e300: none)
                    other = fmin(other, (w2 == win) ? box->rv2[w2] : box->rv[w2]);
                const double unorm = (double)sqrtf((float)fval) * 1.0001 + 1e-22;
                const double delta = 1.01 * (double)(a.n + 3) * 5.9604644775390625e-8 * unorm;
                const bool ok = uni(box->rs[win] != 0 && other - wb > 2.0 * delta);
                __syncwarp();
                if (ok) return wk;
            }
        }
#if defined(DAQP_B200_PHASE_CLOCKS) && DAQP_B200_PHASE_CLOCKS != 2
        if (lane == 0) cnt()[6]++; // exact scans (the screening was absent or could not name the row)
#endif
        team_run(TC_SCAN64, 0, 0);
        T wb = 0;
        int wk = INT_MAX;
#pragma unroll
        for (int w2 = 0; w2 < TW; w2++) {
            const T b = (T)box->rv[w2];
            const int kx = box->rk[w2];
            if (kx != INT_MAX && (wk == INT_MAX || b < wb || (b == wb && kx < wk))) { wb = b; wk = kx; }
        }
        wk = uni(wk);
        __syncwarp();
        return wk == INT_MAX ? -1 : wk;
    }

    // ---- a13: warm start / equality activation (auxiliary.c:399-479)
    // ---- a13 at once, single-warp form (see activate_constraints and team_ops.cuh: team_gram / team_ldl)
    // g_ij = row(WS[i]) . row(WS[j]) for all j <= i < K: two rows i in registers, the rows j streamed past them four at a time
    __device__ __forceinline__ void gram_rows(int K) {
        constexpr int IB = 2, JB = 4;
        const T* M = reinterpret_cast<const T*>(Mr()) + V * lane;
        const int* ws = WS();
        T* Lp = L();
        T* Dp = D();
        bool okg[NG];
#pragma unroll
        for (int g = 0; g < NG; g++) okg[g] = V * (lane + 32 * g) < a.ldn;
        for (int i0 = 0; i0 < K; i0 += IB) {
            const int jend = min(i0 + IB, K);
            T mi[IB][NG][V];
#pragma unroll
            for (int r = 0; r < IB; r++) {
                const T* row = M + (size_t)(unsigned)(ws[min(i0 + r, K - 1)] * a.ldn);
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
                    const T* row = M + (size_t)(unsigned)(ws[min(j0 + rr, K - 1)] * a.ldn);
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
        __syncwarp();
    }
    // in-place right-looking LDL' of that Gram matrix; lane owns rows lane + 32 q. Returns K, or the index of the first row
    // the row-by-row path treats specially (singular pivot, pivot swap). The published columns alternate between xldl and
    // zldl, which hold nothing before the first CSP.
    __device__ __forceinline__ int ldl_rows(int K) {
        T* Lp = L();
        T* Dp = D();
        T acc[NV];
#pragma unroll
        for (int q = 0; q < NV; q++) acc[q] = 0;
        int stop = uni((int)(Dp[0] < a.st.sing_tol)) ? 0 : K;
        for (int j = 0; j + 1 < K && stop > j; j++) {
            T* col = (j & 1) ? zl() : xl();
            const T dj = Dp[j];
            T x[NV];
            int bad = 0;
#pragma unroll
            for (int q = 0; q < NV; q++) {
                x[q] = 0;
                if (q > 0 && 32 * q >= K) break; // uniform: no rows in this register segment
                const int i = lane + 32 * q;
                const bool below = i > j && i < K;
                if (below) x[q] = Lp[loff(i) + j];
                const T l = fdiv(x[q], dj); // unconditional call: no divergence around the division
                if (below) {
                    acc[q] += x[q] * l;
                    Lp[loff(i) + j] = l;
                    col[i] = l;
                    if (i == j + 1) {
                        const T d = Dp[i] - acc[q];
                        Dp[i] = d;
                        bad = d < a.st.sing_tol || (dj < a.st.pivot_tol && dj < d);
                    }
                }
            }
            if (uni(bad)) stop = j + 1; // (uni = warp-wide max; it also orders the column before the reads below)
            __syncwarp();
#pragma unroll
            for (int q = 0; q < NV; q++) {
                if (q > 0 && 32 * q >= K) break;
                const int i = lane + 32 * q;
                if (i > j + 1 && i < K) {
                    T* Li = Lp + loff(i);
                    const T xq = x[q];
#pragma unroll 4
                    for (int t = j + 1; t < i; t++) Li[t] -= col[t] * xq;
                }
            }
        }
        __syncwarp();
        return stop;
    }

    __device__ __forceinline__ int activate_constraints() {
        unsigned char* se = sense();
        // A whole warm start is activated at once: one pass for every dot product the K row-by-row updates would compute,
        // one right-looking pass for the K forward substitutions -- the same products in the same order, without
        // K x (row fetch + dependent sweep). Anything the row-by-row path treats specially (singular pivot, pivot swap,
        // more rows than dimensions, soft rows) sends the activation back to it. A team splits both passes over its warps
        // (team_gram / team_ldl); a single warp runs them itself (gram_rows / ldl_rows; fp64, n <= 127).
        constexpr bool WARP_FAST = TW <= 1 && sizeof(T) == 8 && NG <= 2 && NV <= 4;
        if constexpr (TW > 1 || WARP_FAST) {
            if (uni(k == 0) && !(EXT && a.ns_max > 0)) {
                int K = 0;
                int* wsp = WS();
                for (int base = 0; base < a.m; base += 32) {
                    const int i = base + lane;
                    const bool on = i < a.m && (se[i] & B_ACTIVE);
                    const unsigned msk = __ballot_sync(FULL, on);
                    const int pos = K + __popc(msk & ((1u << lane) - 1u));
                    if (on && pos < a.cap) wsp[pos] = i;
                    K += __popc(msk);
                }
                K = uni(K);
                __syncwarp();
                if (K >= TEAM_GRAM_MIN && K <= a.n) {
                    int done;
                    if constexpr (TW > 1) {
                        if (lane == 0) tbox()->rk[0] = K;
#if defined(DAQP_B200_PHASE_CLOCKS) && DAQP_B200_PHASE_CLOCKS == 2 // (slot 6: cycles of the Gram pass, slot 5: of the LDL' pass)
                        { PHASE_T0; team_run(TC_GRAM, K, 0); PHASE_T1(6); }
                        { PHASE_T0; team_run(TC_LDL, K, 0); PHASE_T1(5); }
#else
                        team_run(TC_GRAM, K, 0);
                        team_run(TC_LDL, K, 0);
#endif
                        done = uni((int)reinterpret_cast<volatile TeamBox*>(tbox())->rk[0]);
                    } else {
                        gram_rows(K);
                        done = ldl_rows(K);
                    }
                    if (done >= K) {
                        LANE_LOOP(j, 0, K) {
                            const int id = wsp[j], sb = se[id];
                            lam()[j] = (sb & B_LOWER) ? (T)-1 : (T)1;
                            dact()[j] = (sb & B_LOWER) ? dl()[id] : du()[id];
                        }
                        if (lane == 0) cnt()[1] += K;
                        k = K;
                        sing = EMPTY_IND;
                        __syncwarp();
                        return 1;
                    }
                }
            }
        }
        for (int i = 0; i < a.m; i++) {
            const int sb = uni((int)se[i]);
            if (sb & B_ACTIVE) modify(OP_ADD, i, (sb & B_LOWER) ? (T)-1 : (T)1);
            if (uni(sing != EMPTY_IND)) {
                const int last = uni(WS()[k - 1]);
                if (uni((int)se[last]) & B_IMMUTABLE) {
                    singular_direction();
                    T resid = 0, scale = 0;
                    const int kq = uni(k);
                    LANE_LOOP(j, 0, kq) {
                        const T term = lams()[j] * dact()[j];
                        resid += term;
                        scale += term < 0 ? -term : term;
                    }
                    resid = uni(warp_sum(resid));
                    scale = (T)1 + uni(warp_sum(scale));
                    if (lane == 0) se[last] &= ~B_ACTIVE;
                    k--;
                    sing = EMPTY_IND;
                    if (reuse > k) reuse = k;
                    __syncwarp();
                    if (!uni(resid <= a.st.primal_tol * scale && resid >= -a.st.primal_tol * scale))
                        return EXIT_OVERDETERMINED_INITIAL;
                } else {
                int flag = 1;
                LANE_LOOP(j, i, a.m) {
                    const int s2 = se[j];
                    if (s2 & B_ACTIVE) {
                        if (s2 & B_IMMUTABLE) flag = EXIT_OVERDETERMINED_INITIAL;
                        else se[j] = s2 & ~B_ACTIVE;
                    }
                }
                flag = uni(__reduce_min_sync(FULL, flag));
                k--;
                sing = EMPTY_IND;
                __syncwarp();
                return flag;
                }
            }
        }
        return 1;
    }

    // ---- a12: one step of iterative refinement on the active rows (auxiliary.c:498-593)
    __device__ __forceinline__ void refine_active() {
        reuse = 0;
        const int kk = uni(k);
        const char* M = Mr() + (size_t)(V * lane) * sizeof(T);
        const unsigned rstride = a.ldn * (unsigned)sizeof(T);
        T* xp = xl();
        T* up = u();
        for (int i = 0; i < kk; i++) {
            const T* row = reinterpret_cast<const T*>(M + (size_t)WS()[i] * rstride);
            T part = 0;
#pragma unroll
            for (int g = 0; g < NG; g++) {
                const int c = V * (lane + 32 * g);
                if (c < a.ldn) {
                    T t[V];
                    ldg_vec<T>(row + 32 * V * g, t);
#pragma unroll
                    for (int e = 0; e < V; e++) part += t[e] * up[c + e];
                }
            }
            part = warp_sum(part);
            if (lane == 0) {
                T res = part - dact()[i];
                if (EXT && a.ns_max > 0 && (sense()[WS()[i]] & B_SOFT)) res -= a.st.rho_soft * lams()[i]; // auxiliary.c:534-535
                xp[i] = res;
            }
        }
        __syncwarp();
        {
            T rv[NV];
            vload(rv, xp, kk);
            forward_sweep(rv, 0, kk);
#pragma unroll
            for (int q = 0; q < NV; q++) {
                const int i = lane + 32 * q;
                if (i < kk) { rv[q] = rv[q] / D()[i]; zl()[i] = rv[q]; }
            }
            backward_sweep(rv, kk);
            vstore(rv, xp, 0, kk);
            __syncwarp();
        }
        LANE_LOOP(i, 0, kk) lams()[i] += xp[i];
        T acc[NG][V];
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const int c = V * (lane + 32 * g);
#pragma unroll
            for (int e = 0; e < V; e++) acc[g][e] = (c < a.ldn) ? up[c + e] : (T)0;
        }
        for (int i = 0; i < kk; i++) {
            const T* row = reinterpret_cast<const T*>(M + (size_t)WS()[i] * rstride);
            const T dlam = xp[i];
#pragma unroll
            for (int g = 0; g < NG; g++) {
                if (V * (lane + 32 * g) < a.ldn) {
                    T t[V];
                    ldg_vec<T>(row + 32 * V * g, t);
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
            if (c < a.ldn) {
#pragma unroll
                for (int e = 0; e < V; e++) { up[c + e] = acc[g][e]; u32()[c + e] = (float)acc[g][e]; part += acc[g][e] * acc[g][e]; }
            }
        }
        if constexpr (EXT) fval = soft_slack + uni(warp_sum(part)); // auxiliary.c:589-593: the slack term is not recomputed
        else fval = uni(warp_sum(part));
        __syncwarp();
    }

    // ---- a11: the daqp_ldp state machine (daqp.c:6-108), preceded by the activation daqp_update_ldp runs for a warm
    // start (utils.c:199-211). Same decisions in the same order as the reference, but RESUMABLE: step() executes
    // exactly one pass of the reference's for-loop and returns, so that the kernel can line all warps of a CTA up at
    // the top of every iteration (see ldp_solve_kernel). Control flow is arranged so that the direction solve, the
    // ratio test, the scan and the working-set modification each appear once.
    static constexpr int RUNNING = 0x7fffffff;
    __device__ __forceinline__ void begin(bool activate_first) {
        iter = 0; tried_repair = 0; cycle_counter = 0; best_fval = -1; do_activate = activate_first;
    }
    // Returns RUNNING, or the exit flag when the solve is over (iter holds the iteration count; iter == 0 means the
    // initial activation failed, which the reference reports as a setup failure).
    __device__ __forceinline__ int step() {
        const T fval_bound = 2 * a.st.fval_bound;
        if (uni(do_activate)) { // end of the previous iteration's refactor / cycle repair, or the warm start
            do_activate = false;
            if (iter > 0) reset();
            int aflag;
            { PHASE_T0; aflag = uni(activate_constraints()); PHASE_T1(7); }
            if (iter == 0 && aflag < 0) return aflag;
        }
        iter = uni(iter + 1);
        if (iter >= a.st.iter_limit) return EXIT_ITERLIMIT; // for(iter=1; iter < iter_limit; ++iter)
        if constexpr (TW > 1) { // experiment knobs: pull this iteration's streams into L2 while the sweeps run
            if ((a.tune & 16) && a.Mt32) {
                const char* q = reinterpret_cast<const char*>(a.Mt32) + (size_t)p * a.sMt32;
                for (unsigned off = 32768u * lane; off < a.sMt32; off += 32u * 32768u)
                    bulk_prefetch_l2(q + off, min(32768u, a.sMt32 - off) & ~15u);
            }
            if (a.tune & 32) {
                const unsigned rb = a.ldn * (unsigned)sizeof(T);
                const int kq = uni(k);
                LANE_LOOP(i, 0, kq) bulk_prefetch_l2(Mr() + (size_t)(unsigned)WS()[i] * rb, rb);
            }
        }
        const bool was_singular = uni(sing != EMPTY_IND);
        { PHASE_T0; if (!was_singular) compute_csp(); else singular_direction(); PHASE_T1(0); }
        int op = OP_REMOVE, arg;
        { PHASE_T0; arg = find_blocking(); PHASE_T1(1); }
        T lamval = 0;
        bool refined = false;
        if (arg < 0) { // no blocking constraint: dual feasible (or, in singular mode, primal infeasible)
            if (was_singular) return EXIT_INFEASIBLE; // daqp.c:88-93
            { PHASE_T0; compute_primal(); PHASE_T1(2); }
            if (fval > fval_bound) return EXIT_INFEASIBLE;
            bool again = true;
            while (again) { // at most two passes: the scan is repeated once after a refinement (daqp.c:52-56)
                again = false;
                int key;
                { PHASE_T0; key = uni(scan_infeasible()); PHASE_T1(3); }
                if (key >= 0) {
                    arg = key >> 1;
                    const int lower = key & 1;
                    if (lane == 0) { if (lower) sense()[arg] |= B_LOWER; else sense()[arg] &= ~B_LOWER; }
                    lsw ^= 1; // lam <- lam* : the reference swaps the two pointers (auxiliary.c:159-160)
                    __syncwarp();
                    op = OP_ADD;
                    lamval = lower ? (T)-1 : (T)1;
                } else {
                    // primal feasible: KKT point unless the factor is ill-conditioned (daqp.c:28-63)
                    T min_D = (T)1e30;
                    const int kq = uni(k);
                    LANE_LOOP(i, 0, kq) min_D = fmin(min_D, D()[i]);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) min_D = fmin(min_D, __shfl_xor_sync(FULL, min_D, o));
                    min_D = uni(min_D);
                    if (kq > 2 && uni(tried_repair != 1) && min_D < a.st.refactor_tol) {
                        tried_repair = 1;
                        count(6);
                        trace(3, kq);
                        LANE_LOOP(i, 0, kq) {
                            const int id = WS()[i];
                            if (lam()[i] >= 0) sense()[id] &= ~B_LOWER; else sense()[id] |= B_LOWER;
                        }
                        __syncwarp();
                        do_activate = true; // refactor: handled at the top of the next step
                        return RUNNING;
                    }
                    if (!refined && kq > 0 && min_D < a.st.pivot_tol) {
                        count(5);
                        trace(4, kq);
                        refine_active();
                        refined = true;
                        again = true;
                    } else return (EXT && uni(soft_slack > a.st.primal_tol)) ? EXIT_SOFT_OPTIMAL : EXIT_OPTIMAL; // daqp.c:59-62
                }
            }
        }
        if (AUX && a.trace_out != nullptr) trace(op == OP_ADD ? 1 : 2, op == OP_ADD ? 2 * arg + (lamval < 0 ? 1 : 0) : uni(WS()[arg]));
        { PHASE_T0; modify(op, arg, lamval); PHASE_T1(op == OP_ADD ? 4 : 5); } // the ONE place where the working set changes inside the loop
        if (op == OP_ADD && !refined) { // cycle guard, daqp.c:67-85 (skipped on the refine path, daqp.c:54-55)
            if (AUX && a.trace_out != nullptr) { // the objective the guard compares, bit for bit (8 = high word, 9 = low word)
                const long long fb = __double_as_longlong((double)fval);
                trace(8, (int)(fb >> 32)); trace(9, (int)(fb & 0xffffffffll));
            }
            if (uni(fval - best_fval < a.st.progress_tol)) {
                if (uni(cycle_counter++ > a.st.cycle_tol)) {
                    if (uni(tried_repair == 1)) return EXIT_CYCLE;
                    tried_repair = 1;
                    count(7);
                    trace(5, 0);
                    do_activate = true;
                    cycle_counter = 0;
                    best_fval = -1;
                }
            } else {
                best_fval = fval;
                cycle_counter = 0;
            }
        }
        return RUNNING;
    }
};

// One warp per problem, persistent CTAs pulling problem indices from an atomic queue (iteration counts diverge).
//
// The warps of a CTA never synchronise with each other (TW <= 1): every warp owns its problem for the whole solve and
// sits in its own phase of the iteration, which is why the hot instruction footprint, not the kernel's size, decides the
// instruction-cache hit rate (see the header). (An earlier version lined the warps up at one block barrier per iteration
// to share instruction-cache lines; keeping the hot code small did better.)
//
// TW > 1 (team mode, n > 64): one CTA of TW warps per problem -- at n = 120 the packed factor is 58 KB, so only three
// problems fit an SM and a warp each would leave the SM at three warps. Warp 0 runs the loop below, warps 1.. serve it.
template <typename T, int NV, bool EXT, int TW = 1>
__global__ void __launch_bounds__(TW > 1 ? 32 * TW : 512, TW > 1 ? team_max_ctas(TW) : 1) ldp_solve_kernel(const __grid_constant__ LdpArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = uni((int)(threadIdx.x >> 5)); // uniform: so are all shared-memory bases
    const int gw = TW > 1 ? (int)blockIdx.x : (int)(blockIdx.x * (blockDim.x >> 5) + wib);

    Warp<T, NV, EXT, TW> w(a);
    w.lane = lane;
    w.S = reinterpret_cast<T*>(smem_raw + (TW > 1 ? (size_t)a.team_prefix : (size_t)a.per_warp_bytes * wib));
    w.pst_id = a.pst_id + (size_t)gw * a.cap;
    w.pst_lam = a.pst_lam + (size_t)gw * a.cap;
    if constexpr (TW > 1) {
        if (wib != 0) { // helper warp: wait for a command, run this warp's share, meet the leader at the closing barrier
            const volatile TeamBox* b = reinterpret_cast<const volatile TeamBox*>(smem_raw);
            for (;;) {
                team_bar<TW>();
                const int cmd = b->cmd, a0 = b->a0, a1 = b->a1;
                if (cmd == TC_EXIT) return;
                team_dispatch<T, TW, (NV + 1) / 2>(cmd, lane, wib, a0, a1);
                team_bar<TW>();
            }
        }
        w.team_publish_layout();
    }

    // Structured on purpose (see modify()): a problem loop around an iteration loop, no loop-carried control flags.
    for (;;) {
        // ---- take the next problem off the queue (problems finished by the setup kernel are only passed through)
        int pq = 0;
        if (lane == 0) pq = atomicAdd(a.work_counter, 1);
        pq = uni(pq); // lanes other than 0 hold 0 and queue indices are non-negative
        if (pq >= a.P) break;
        w.p = pq;
        if constexpr (TW > 1) w.team_publish_problem(); // the helpers read it with the next command
        if constexpr (EXT) w.pm = a.grp > 1 ? pq / a.grp : pq;
        const int sflag = uni(a.setup_flag[pq]); // loaded values are divergent in ptxas' eyes until proven otherwise
        if (sflag != SETUP_SOLVE && sflag != SETUP_SOLVE_ACTIVATE) {
            if (a.nact_out && lane == 0) a.nact_out[pq] = 0;
            if (a.counts_out && lane < 8) a.counts_out[8 * (size_t)pq + lane] = 0;
            if (a.sense_out)
                LANE_LOOP(i, 0, a.m) a.sense_out[(size_t)pq * a.ldm + i] = a.sense[(size_t)pq * a.ldm + i];
            __syncwarp(); // reconverge before the back edge: a lane-divergent tail would leave the loop header diverged
        } else {
        {
            w.lsw = 0;
            w.fval = 0;
            if constexpr (EXT) w.soft_slack = 0;
            const unsigned char* sin = a.sense + (size_t)pq * a.ldm;
            unsigned char* se = w.sense();
            LANE_LOOP(i, 0, a.m) se[i] = sin[i];
            T* up = w.u();
            LANE_LOOP(i, 0, round_up(a.n, VecOf<T>::N) + U_PAD) up[i] = 0;
            if (lane < 8) w.cnt()[lane] = 0;
            float* u32p = w.u32();
            LANE_LOOP(i, 0, round_up(a.n, 4) + U_PAD) u32p[i] = 0.f;
            __syncwarp();
            w.reset();
            bool activate = sflag == SETUP_SOLVE_ACTIVATE;
            if (EXT && a.state && a.state_load) {
                // warm: continue from the previous solve's factor and working set (the reference's daqp_solve on a kept
                // workspace). The right-hand sides changed (daqp_update_d sets reuse_ind = 0, utils.c:506), so the active
                // bounds are re-read; a constraint that the new bounds turned into an equality (check_bounds,
                // utils.c:546-567) forces the reference's reset + re-activation (utils.c:199-211).
                const char* blob = a.state + (size_t)pq * a.state_stride;
                const int* meta = reinterpret_cast<const int*>(blob + a.state_stride - 16);
                if (uni(meta[2]) == 1) {
                    const int4* src4 = reinterpret_cast<const int4*>(blob);
                    int4* dst4 = reinterpret_cast<int4*>(w.S);
                    LANE_LOOP(i, 0, (int)((a.state_stride - 16) / 16)) dst4[i] = src4[i];
                    __syncwarp();
                    w.k = uni(meta[0]);
                    w.lsw = uni(meta[1]);
                    int fresh = 0; // marks the new bounds add to the kept sense bits
                    LANE_LOOP(i, 0, a.m) {
                        const int now = sin[i], kept = se[i];
                        if ((now & ~kept) & (B_ACTIVE | B_IMMUTABLE)) { fresh = 1; se[i] = (unsigned char)(kept | now); }
                    }
                    __syncwarp();
                    if (uni(fresh != 0)) { w.reset(); activate = true; }
                    else {
                        activate = false;
                        T* da = w.dact();
                        const int* wsp = w.WS();
                        LANE_LOOP(i, 0, w.k) {
                            const int id = wsp[i];
                            da[i] = (se[id] & B_LOWER) ? w.dl()[id] : w.du()[id];
                        }
                        LANE_LOOP(i, 0, round_up(a.n, VecOf<T>::N) + U_PAD) up[i] = 0;
                        LANE_LOOP(i, 0, round_up(a.n, 4) + U_PAD) u32p[i] = 0.f;
                        if (lane < 8) w.cnt()[lane] = 0;
                        __syncwarp();
                    }
                }
            }
            w.begin(activate);
        }
        if (Warp<T, NV, EXT, TW>::AUX && a.trace_out != nullptr && lane == 0) a.trace_out[(size_t)pq * (1 + 2 * a.trace_cap)] = 0;
        __syncwarp();
        int exitflag;
        // settings->time_limit: the clock starts with the problem's solve (api.c:15) and is read every 32nd iteration
        // (daqp.c:95-103); the iteration count reported is the one the check ran in. (ONE call site of step(): a second
        // inlined copy of the state machine costs a third of the kernel's speed in instruction-cache misses.)
        constexpr bool AUX = Warp<T, NV, EXT, TW>::AUX;
        const long long t0 = (AUX && a.st.time_limit_ns != 0) ? global_timer_ns() : 0;
        do {
            exitflag = uni(w.step());
            if constexpr (AUX) {
                if (a.st.time_limit_ns != 0 && exitflag == Warp<T, NV, EXT, TW>::RUNNING && (w.iter & 31) == 0 && w.iter > 0 &&
                    uni((int)(global_timer_ns() - t0 > a.st.time_limit_ns)))
                    exitflag = EXIT_TIMELIMIT;
            }
        } while (exitflag == Warp<T, NV, EXT, TW>::RUNNING);
        w.trace(7, exitflag);
        {
        const int p = w.p, kfin = uni(w.k);
        if (uni(w.iter == 0)) {
            // the initial activation failed (utils.c:209-210 -> api.c:69-72): flag only, x untouched
            if (lane == 0) { a.exitflag[p] = exitflag; a.iter[p] = 0; }
        } else {
            // ---- a15/a16: ldp2qp_solution (daqp.c:111-139) + daqp_extract_result (api.c:455-495)
            const T* vv = a.v ? reinterpret_cast<const T*>(reinterpret_cast<const char*>(a.v) + (size_t)p * a.sv) : nullptr;
            T* xo = a.x + (size_t)p * a.n;
            T* up = w.u();
            T vnorm = 0;
            if (vv) { LANE_LOOP(i, 0, a.n) { const T t = vv[i]; vnorm += t * t; } vnorm = warp_sum(vnorm); }
            if (exitflag > 0) {
                const T* Ri = reinterpret_cast<const T*>(reinterpret_cast<const char*>(a.Rinv) + (size_t)w.pmat() * a.sRinv);
                if (vv) LANE_LOOP(i, 0, a.n) up[i] -= vv[i];
                __syncwarp();
                for (int i = 0; i < a.n; i++) { // x_i = sum_{j>=i} Rinv[i][j] (u-v)_j ; rows are independent
                    const T* row = Ri + roff(i, a.n);
                    T acc = 0;
                    LANE_LOOP(j, i, a.n) acc += row[j] * up[j];
                    acc = warp_sum(acc);
                    if (i < a.ms) acc /= w.sc()[i];
                    if (lane == 0) xo[i] = acc;
                }
            } else {
                LANE_LOOP(i, 0, a.n) xo[i] = up[i];
            }
            if (a.lam) {
                T* lo = a.lam + (size_t)p * a.m;
                LANE_LOOP(i, 0, a.m) lo[i] = 0;
                __syncwarp();
                LANE_LOOP(i, 0, kfin) {
                    const int id = w.WS()[i];
                    lo[id] = (exitflag > 0) ? w.lams()[i] * w.sc()[id] : w.lams()[i];
                }
            }
            if (lane == 0) {
                if (vv) a.fval[p] = (T)0.5 * (w.fval - vnorm);
                if (a.soft_slack) a.soft_slack[p] = EXT ? w.soft_slack : (T)0;
                a.exitflag[p] = exitflag;
                a.iter[p] = w.iter;
            }
        }
        if (a.nact_out && lane == 0) a.nact_out[p] = w.k;
        if (a.ws_out) LANE_LOOP(i, 0, kfin) a.ws_out[(size_t)p * a.cap + i] = w.WS()[i];
        if (a.sense_out) LANE_LOOP(i, 0, a.m) a.sense_out[(size_t)p * a.ldm + i] = w.sense()[i];
        if (a.counts_out && lane < 8) a.counts_out[8 * (size_t)p + lane] = w.cnt()[lane];
        if (EXT && a.state && a.state_save) { // keep factor, multipliers, working set and sense for the next warm solve
            char* blob = a.state + (size_t)p * a.state_stride;
            int4* dst4 = reinterpret_cast<int4*>(blob);
            const int4* src4 = reinterpret_cast<const int4*>(w.S);
            __syncwarp();
            LANE_LOOP(i, 0, (int)((a.state_stride - 16) / 16)) dst4[i] = src4[i];
            if (lane == 0) {
                int* meta = reinterpret_cast<int*>(blob + a.state_stride - 16);
                meta[0] = w.k; meta[1] = w.lsw; meta[2] = (exitflag > 0 && w.sing == EMPTY_IND) ? 1 : 0; meta[3] = 0;
            }
        }
        __syncwarp();
        }
        }
    }
    if constexpr (TW > 1) w.team_exit(); // releases the helpers
}

} // namespace dq
