/* oracle/daqp_oracle.c -- TEST INFRASTRUCTURE ONLY (see daqp_oracle.h).
 *
 * Scalar CPU restatement of the reference's dual active-set QP path. Written from the behaviour of the reference
 * (file:line cited on every function); data structures are this file's own (square L instead of packed rows,
 * an explicit stack instead of the recursive pivoting), arithmetic expression order follows the reference so that
 * a build without FP reassociation is bit-identical to oracle/_ref/libdaqp_ref_strict.so.
 */
#include "daqp_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>

typedef orc_real real;

#define EMPTY_IND (-1)
#define ORC_INF ((real)1e30)

/* sense bits (include/constants.h:64-96) */
#define B_ACTIVE 1
#define B_LOWER 2
#define B_IMMUTABLE 4
#define B_SOFT 8
#define B_BINARY 16

/* exit flags (include/constants.h:42-51) */
#define EXIT_SOFT_OPTIMAL 2
#define EXIT_OPTIMAL 1
#define EXIT_INFEASIBLE (-1)
#define EXIT_CYCLE (-2)
#define EXIT_ITERLIMIT (-4)
#define EXIT_NONCONVEX (-5)
#define EXIT_OVERDETERMINED_INITIAL (-6)
#define EXIT_UNSUPPORTED (-8)

typedef struct {
    int n, m, ms, cap;
    const OrcSettings *st;
    /* LDP data */
    real *M;       /* (m-ms) x n, rows normalised                       */
    real *Rinv;    /* packed upper triangle by rows, or NULL (diagonal) */
    real *RinvD;   /* diagonal of R^-1 when H is diagonal, else NULL    */
    real *v;       /* R^-T f, or NULL when f == NULL                    */
    real *dupper, *dlower, *scaling, *Mu;
    int *sense;
    /* iterates */
    real *u, *xunc;
    real *lam, *lam_star, *D, *xl, *zl;
    real *L;       /* cap x cap, row i holds L[i][0..i-1]; unit diagonal implicit */
    int *WS;
    int k, reuse, sing, iterations;
    int unconstrained_optimal;
    real fval, soft_slack;
    /* pivoting stack (replaces the recursion of auxiliary.c:379-396) */
    int *pstack_id; real *pstack_lam;
    int n_scan, n_add, n_remove, n_csp;
    int n_pivot, n_refine, n_refactor, n_cycle; /* rare paths: pivot_last swaps, refinements, refactor-on-exit, cycle repairs */
    int *oplog; int oplog_cap, n_log;           /* decision log (debugging aid) */
} Ldp;

static void oplog(Ldp *w, int code, int val) {
    if (w->oplog && w->n_log < w->oplog_cap) { w->oplog[2 * w->n_log] = code; w->oplog[2 * w->n_log + 1] = val; }
    w->n_log++;
}

#define Lij(w, i, j) ((w)->L[(size_t)(i) * (w)->cap + (j)])

/* constants.h:39 -- offset such that (Rinv + roff(i,n))[j] is element (i,j), j>=i, of the packed upper triangle */
static int roff(int i, int n) { return ((2 * n - i - 1) * i) / 2; }

/* factorization.c:4-15: four partial sums, combined (s0+s1)+(s2+s3) */
static real dot4(const real *a, const real *b, int len) {
    real s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int i = 0;
    for (; i + 3 < len; i += 4) {
        s0 += a[i] * b[i];
        s1 += a[i + 1] * b[i + 1];
        s2 += a[i + 2] * b[i + 2];
        s3 += a[i + 3] * b[i + 3];
    }
    for (; i < len; i++) s0 += a[i] * b[i];
    return (s0 + s1) + (s2 + s3);
}

/* factorization.h:13-17: plain sequential dot */
static real dot1(const real *a, const real *b, int len) {
    real s = 0;
    for (int i = 0; i < len; i++) s += a[i] * b[i];
    return s;
}

/* Row of the normalised constraint matrix for constraint id. Returns pointer p such that p[j] is column j,
 * valid for j >= *first. NULL means the unit vector e_id (identity / diagonal Hessian, simple bound). */
static const real *con_row(const Ldp *w, int id, int *first) {
    if (id < w->ms) {
        *first = id;
        return w->Rinv ? w->Rinv + roff(id, w->n) : NULL;
    }
    *first = 0;
    return w->M + (size_t)w->n * (id - w->ms);
}

/* ---- LDL' row append: factorization.c:21-111 ------------------------------------------------------------ */
static void ldl_append(Ldp *w, int add) {
    const int k = w->k, n = w->n;
    int c0, ns_active = 0;
    w->sing = EMPTY_IND;
    w->n_add++;
    const real *mi = con_row(w, add, &c0);
    real d = mi ? dot4(mi + c0, mi + c0, n - c0) : (real)1;
    if (w->sense[add] & B_SOFT) { d += w->st->rho_soft; ns_active++; }
    w->D[k] = d;
    if (k == 0) return;

    real *l = &Lij(w, k, 0);
    for (int i = 0; i < k; i++) {
        int id = w->WS[i], r0, j;
        if (w->sense[id] & B_SOFT) ns_active++;
        const real *mk = con_row(w, id, &r0);
        j = (id < w->ms) ? (c0 > id ? c0 : id) : c0;
        if (mk == NULL) l[i] = mi ? mi[j] : (real)0;
        else if (mi == NULL) l[i] = mk[j];
        else l[i] = dot4(mk + j, mi + j, n - j);
    }
    for (int i = 0; i < k; i++) { /* l <- L^-1 l */
        real s = l[i];
        for (int j = 0; j < i; j++) s -= Lij(w, i, j) * l[j];
        l[i] = s;
    }
    real s = w->D[k];
    for (int i = 0; i < k; i++) { /* l <- D^-1 l ; d -= l'Dl */
        real t = l[i];
        l[i] /= w->D[i];
        s -= t * l[i];
    }
    w->D[k] = s;
    if (w->D[k] < w->st->sing_tol || k >= n + ns_active) {
        w->sing = k;
        w->D[k] = 0;
    }
}

/* ---- LDL' row/column deletion + rank-one update (Gill et al. 1974, C1): factorization.c:112-151 --------- */
static void ldl_delete(Ldp *w, int r) {
    const int k = w->k;
    w->n_remove++;
    if (k == r + 1) return;
    const int nu = k - r - 1;
    real *q = &w->zl[r]; /* reference reuses zldl[r..] as scratch (factorization.c:116) */
    for (int t = 0; t < nu; t++) q[t] = Lij(w, r + 1 + t, r);
    for (int i = r + 1; i < k; i++) {
        for (int j = 0; j < r; j++) Lij(w, i - 1, j) = Lij(w, i, j);
        for (int j = r + 1; j < i; j++) Lij(w, i - 1, j - 1) = Lij(w, i, j);
    }
    real alpha = w->D[r];
    for (int t = 0; t < nu; t++) {
        const int io = r + 1 + t; /* old index of this pivot; new index c = r+t */
        const int c = r + t;
        real p = q[t];
        real dbar = w->D[io] + alpha * p * p;
        w->D[io - 1] = dbar;
#ifdef ORC_GPU_ARITH /* the CUDA kernels' form: one correctly rounded reciprocal for both quotients (fixture triage only) */
        const real rdb = (real)1 / dbar;
        real beta = p * alpha * rdb;
        alpha = w->D[io] * alpha * rdb;
#else
        real beta = p * alpha / dbar;
        alpha = w->D[io] * alpha / dbar;
#endif
        for (int s = t + 1; s < nu; s++) {
            q[s] -= p * Lij(w, r + s, c);
            Lij(w, r + s, c) += beta * q[s];
        }
    }
}

static void pivot_last(Ldp *w);

/* raw bookkeeping halves of auxiliary.c:3-44 (without the trailing pivot_last calls) */
static void raw_add(Ldp *w, int add, real lam) {
    w->sense[add] |= B_ACTIVE;
    ldl_append(w, add);
    w->WS[w->k] = add;
    w->lam[w->k] = lam;
    w->k++;
}
/* returns 1 if the removal made the factor singular (auxiliary.c:19-22) */
static int raw_remove(Ldp *w, int r) {
    w->sense[w->WS[r]] &= ~B_ACTIVE;
    ldl_delete(w, r);
    w->k--;
    for (int i = r; i < w->k; i++) {
        w->WS[i] = w->WS[i + 1];
        w->lam[i] = w->lam[i + 1];
    }
    if (r < w->reuse) w->reuse = r;
    if (w->k > 0 && w->D[w->k - 1] < w->st->sing_tol) {
        w->sing = w->k - 1;
        w->D[w->k - 1] = 0;
        return 1;
    }
    return 0;
}

/* auxiliary.c:27-44 */
static void add_constraint(Ldp *w, int add, real lam) {
    raw_add(w, add, lam);
    pivot_last(w);
}
/* auxiliary.c:3-26 */
static void remove_constraint(Ldp *w, int r) {
    if (!raw_remove(w, r)) pivot_last(w);
}

/* auxiliary.c:379-396. The reference recurses (pivot_last -> remove_constraint -> pivot_last ... ->
 * add_constraint -> pivot_last). Unrolled here: every pending "re-add after the nested removal returns"
 * is one stack entry; add_constraint's trailing pivot_last is a tail call. */
static void pivot_last(Ldp *w) {
    int depth = 0;
    for (;;) {
        const int r = w->k - 2;
        if (w->k > 1 && w->D[r] < w->st->pivot_tol && w->D[r] < w->D[w->k - 1]) {
            w->pstack_id[depth] = w->WS[r];
            w->pstack_lam[depth] = w->lam[r];
            depth++;
            w->n_pivot++;
            if (!raw_remove(w, r)) continue; /* nested pivot_last inside remove_constraint */
        }
        /* a pivot_last invocation (or a remove_constraint that turned singular) returns here */
        for (;;) {
            if (depth == 0) return;
            depth--;
            if (w->sing != EMPTY_IND) continue; /* auxiliary.c:392: abort this frame, keep unwinding */
            raw_add(w, w->pstack_id[depth], w->pstack_lam[depth]);
            break; /* tail call: pivot_last again */
        }
    }
}

/* ---- constrained stationary point: auxiliary.c:314-354 -------------------------------------------------- */
static void compute_csp(Ldp *w) {
    const int k = w->k;
    w->n_csp++;
    for (int i = w->reuse; i < k; i++) {
        int id = w->WS[i];
        real s = (w->sense[id] & B_LOWER) ? -w->dlower[id] : -w->dupper[id];
        for (int j = 0; j < i; j++) s -= Lij(w, i, j) * w->xl[j];
        w->xl[i] = s;
    }
    for (int i = w->reuse; i < k; i++) w->zl[i] = w->xl[i] / w->D[i];
    for (int i = k - 1; i >= 0; i--) {
        real s = w->zl[i];
        for (int j = k - 1; j > i; j--) s -= w->lam_star[j] * Lij(w, j, i);
        w->lam_star[i] = s;
    }
    w->reuse = k;
}

/* ---- dual ratio test + removal: auxiliary.c:277-311 ----------------------------------------------------- */
static int remove_blocking(Ldp *w) {
    int rm = EMPTY_IND;
    real alpha = ORC_INF;
    const real dual_tol = w->st->dual_tol;
    for (int i = 0; i < w->k; i++) {
        int id = w->WS[i];
        real ac;
        if (w->sense[id] & B_IMMUTABLE) continue;
        if (w->sense[id] & B_LOWER) {
            if (w->lam_star[i] < dual_tol) continue;
        } else if (w->lam_star[i] > -dual_tol) continue;
        if (w->sing == EMPTY_IND) ac = -w->lam[i] / (w->lam_star[i] - w->lam[i]);
        else ac = -w->lam[i] / w->lam_star[i];
        if (ac < alpha) { alpha = ac; rm = i; }
    }
    if (rm == EMPTY_IND) return 0;
    if (w->sing == EMPTY_IND)
        for (int i = 0; i < w->k; i++) w->lam[i] += alpha * (w->lam_star[i] - w->lam[i]);
    else
        for (int i = 0; i < w->k; i++) w->lam[i] += alpha * w->lam_star[i];
    w->sing = EMPTY_IND;
    oplog(w, 2, w->WS[rm]);
    remove_constraint(w, rm);
    return 1;
}

/* ---- primal iterate u = -Mk' lam*, fval: auxiliary.c:46-88 ---------------------------------------------- */
static void compute_primal(Ldp *w) {
    const int n = w->n;
    real fv = 0;
    for (int j = 0; j < n; j++) w->u[j] = 0;
    for (int i = 0; i < w->k; i++) {
        int id = w->WS[i], c0;
        const real li = w->lam_star[i];
        const real *row = con_row(w, id, &c0);
        if (row == NULL) w->u[id] -= li;
        else for (int j = c0; j < n; j++) w->u[j] -= row[j] * li;
        if (w->sense[id] & B_SOFT) fv += li * li;
    }
    fv = fv * w->st->rho_soft;
    w->soft_slack = fv;
#ifdef ORC_GPU_ARITH /* the solve kernel's order: lane L owns columns 2L, 2L+1 (+64 per group), xor-butterfly over the warp */
    {
        real part[32];
        for (int L = 0; L < 32; L++) {
            part[L] = 0;
            for (int c = 2 * L; c < n; c += 64)
                for (int e = 0; e < 2 && c + e < n; e++) part[L] += w->u[c + e] * w->u[c + e];
        }
        for (int o = 16; o > 0; o >>= 1) {
            real nx[32];
            for (int L = 0; L < 32; L++) nx[L] = part[L] + part[L ^ o];
            memcpy(part, nx, sizeof(part));
        }
        fv += part[0];
    }
#else
    for (int j = 0; j < n; j++) fv += w->u[j] * w->u[j];
#endif
    w->fval = fv;
}

/* ---- feasibility scan: auxiliary.c:89-198 --------------------------------------------------------------- */
static int add_infeasible(Ldp *w) {
    const int n = w->n;
    const real ep = -w->st->primal_tol;
    real min_val = 0;
    int add = EMPTY_IND, isupper = 0;
    w->n_scan++;
    for (int j = 0; j < w->m; j++) {
        real mu, bound, cand;
        if (w->sense[j] & (B_ACTIVE + B_IMMUTABLE)) continue;
        if (j < w->ms) mu = w->Rinv ? dot1(w->Rinv + roff(j, n) + j, w->u + j, n - j) : w->u[j];
        else mu = dot1(w->M + (size_t)n * (j - w->ms), w->u, n);
        bound = ep * w->scaling[j];
        cand = w->dupper[j] - mu;
        if (cand < min_val && cand < bound) { add = j; isupper = 1; min_val = cand; }
        else {
            cand = mu - w->dlower[j];
            if (cand < min_val && cand < bound) { add = j; isupper = 0; min_val = cand; }
        }
    }
    if (add == EMPTY_IND) return 0;
    if (isupper) w->sense[add] &= ~B_LOWER; else w->sense[add] |= B_LOWER;
    real *t = w->lam; w->lam = w->lam_star; w->lam_star = t; /* lam <- lam* (auxiliary.c:159-160) */
    oplog(w, 1, 2 * add + (isupper ? 0 : 1));
    add_constraint(w, add, isupper ? (real)1 : (real)-1);
    return 1;
}

/* ---- singular direction: auxiliary.c:357-376 ------------------------------------------------------------ */
static void singular_direction(Ldp *w) {
    const int s = w->sing;
    for (int i = s - 1; i >= 0; i--) {
        real p = -Lij(w, s, i);
        for (int j = s - 1; j > i; j--) p -= w->lam_star[j] * Lij(w, j, i);
        w->lam_star[i] = p;
    }
    w->lam_star[s] = 1;
    if (w->sense[w->WS[s]] & B_LOWER)
        for (int i = 0; i <= s; i++) w->lam_star[i] = -w->lam_star[i];
}

static void reset_ws(Ldp *w) { w->sing = EMPTY_IND; w->k = 0; w->reuse = 0; } /* daqp.c:142-146 */

/* ---- warm start / equalities: auxiliary.c:399-479 ------------------------------------------------------- */
static int activate_constraints(Ldp *w) {
    for (int i = 0; i < w->m; i++) {
        if (w->sense[i] & B_ACTIVE)
            add_constraint(w, i, (w->sense[i] & B_LOWER) ? (real)-1.0 : (real)1.0);
        if (w->sing != EMPTY_IND) {
            int last = w->WS[w->k - 1];
            if (w->sense[last] & B_IMMUTABLE) {
                real resid = 0, scale = 1;
                singular_direction(w);
                for (int j = 0; j < w->k; j++) {
                    int id = w->WS[j];
                    real b = (w->sense[id] & B_LOWER) ? w->dlower[id] : w->dupper[id];
                    real term = w->lam_star[j] * b;
                    resid += term;
                    scale += term < 0 ? -term : term;
                }
                w->sense[last] &= ~B_ACTIVE;
                w->k--;
                w->sing = EMPTY_IND;
                if (w->reuse > w->k) w->reuse = w->k;
                if (resid <= w->st->primal_tol * scale && resid >= -w->st->primal_tol * scale) continue;
                return EXIT_OVERDETERMINED_INITIAL;
            }
            int flag = 1;
            for (; i < w->m; i++) {
                if (w->sense[i] & B_ACTIVE) {
                    if (w->sense[i] & B_IMMUTABLE) flag = EXIT_OVERDETERMINED_INITIAL;
                    else w->sense[i] &= ~B_ACTIVE;
                }
            }
            w->k--;
            w->sing = EMPTY_IND;
            return flag;
        }
    }
    return 1;
}

/* ---- one step of iterative refinement: auxiliary.c:498-593 ---------------------------------------------- */
static void refine_active(Ldp *w) {
    const int n = w->n, k = w->k;
    w->reuse = 0;
    for (int i = 0; i < k; i++) {
        int id = w->WS[i], c0;
        const real *row = con_row(w, id, &c0);
        real mu = 0;
        if (row == NULL) mu = w->u[id];
        else for (int j = c0; j < n; j++) mu += row[j] * w->u[j];
        real d = (w->sense[id] & B_LOWER) ? w->dlower[id] : w->dupper[id];
        w->xl[i] = mu - d;
        if (w->sense[id] & B_SOFT) w->xl[i] -= w->st->rho_soft * w->lam_star[i];
    }
    for (int i = 0; i < k; i++) {
        real s = w->xl[i];
        for (int j = 0; j < i; j++) s -= Lij(w, i, j) * w->xl[j];
        w->xl[i] = s;
    }
    for (int i = 0; i < k; i++) w->zl[i] = w->xl[i] / w->D[i];
    for (int i = k - 1; i >= 0; i--) {
        real s = w->zl[i];
        for (int j = k - 1; j > i; j--) s -= w->xl[j] * Lij(w, j, i);
        w->xl[i] = s;
    }
    for (int i = 0; i < k; i++) w->lam_star[i] += w->xl[i];
    for (int i = 0; i < k; i++) {
        int id = w->WS[i], c0;
        const real dl = w->xl[i];
        const real *row = con_row(w, id, &c0);
        if (row == NULL) w->u[id] -= dl;
        else for (int j = c0; j < n; j++) w->u[j] -= row[j] * dl;
    }
    real fv = w->soft_slack;
    for (int j = 0; j < n; j++) fv += w->u[j] * w->u[j];
    w->fval = fv;
}

/* ---- the dual active-set loop: daqp.c:6-108 ------------------------------------------------------------- */
static int ldp_solve(Ldp *w) {
    int exitflag = EXIT_ITERLIMIT, iter;
    int tried_repair = 0, cycle_counter = 0;
    real best_fval = -1;
    const real fval_bound = 2 * w->st->fval_bound;

    for (iter = 1; iter < w->st->iter_limit; ++iter) {
        if (w->sing == EMPTY_IND) {
            compute_csp(w);
            if (!remove_blocking(w)) {
                compute_primal(w);
                if (w->fval > fval_bound) { exitflag = EXIT_INFEASIBLE; break; }
                if (!add_infeasible(w)) {
                    real min_D = w->D[0]; /* reference reads D[0] even when k==0; the value is then unused */
                    for (int i = 1; i < w->k; i++) if (w->D[i] < min_D) min_D = w->D[i];
                    if (w->k > 2 && tried_repair != 1 && min_D < w->st->refactor_tol) {
                        tried_repair = 1;
                        w->n_refactor++;
                        oplog(w, 3, w->k);
                        for (int i = 0; i < w->k; i++) {
                            if (w->lam[i] >= 0) w->sense[w->WS[i]] &= ~B_LOWER;
                            else w->sense[w->WS[i]] |= B_LOWER;
                        }
                        reset_ws(w);
                        activate_constraints(w);
                        continue;
                    }
                    if (w->k > 0 && min_D < w->st->pivot_tol) {
                        w->n_refine++;
                        oplog(w, 4, w->k);
                        refine_active(w);
                        if (add_infeasible(w)) continue;
                    }
                    exitflag = (w->soft_slack > w->st->primal_tol) ? EXIT_SOFT_OPTIMAL : EXIT_OPTIMAL;
                    break;
                }
                if (w->oplog) {
                    double fd = (double)w->fval; long long fb;
                    memcpy(&fb, &fd, sizeof(fb));
                    oplog(w, 8, (int)(fb >> 32)); oplog(w, 9, (int)(fb & 0xffffffffll));
                }
                if (w->fval - best_fval < w->st->progress_tol) {
                    if (cycle_counter++ > w->st->cycle_tol) {
                        if (tried_repair == 1) { exitflag = EXIT_CYCLE; break; }
                        tried_repair = 1;
                        w->n_cycle++;
                        oplog(w, 5, 0);
                        reset_ws(w);
                        activate_constraints(w);
                        cycle_counter = 0;
                        best_fval = -1;
                    }
                } else {
                    best_fval = w->fval;
                    cycle_counter = 0;
                }
            }
        } else {
            singular_direction(w);
            if (!remove_blocking(w)) { exitflag = EXIT_INFEASIBLE; break; }
        }
    }
    w->iterations = iter;
    oplog(w, 7, exitflag);
    return exitflag;
}

/* ---- setup pieces: utils.c ------------------------------------------------------------------------------ */

/* utils.c:13-20 */
static real prox_eps_scaled(const OrcSettings *st, real hscale) {
    real eps = st->eps_prox;
    if (eps < 0.0) eps = -eps;
    real floor_ = (real)sqrt(st->zero_tol) * hscale;
    if (eps > 0.0 && eps < floor_) eps = floor_;
    return eps;
}

/* utils.c:223-391 restricted to unfactored H. Returns 1 ok, <0 exit flag. Sets *n_prox when the reference would
 * regularise (then the result is out of scope for this path). */
static int factor_hessian(Ldp *w, const real *H, int *n_prox) {
    const int n = w->n;
    const OrcSettings *st = w->st;
    const real zero_tol = st->zero_tol;
    real hscale = 0;
    *n_prox = 0;
    if (st->eps_prox > 0.0) { *n_prox = n; return 1; } /* forced proximal mode: out of scope */

    int is_diag = 1;
    for (int i = 0; i < n && is_diag; i++) { /* utils.c:246-255: strict upper triangle only */
        real ad = H[i * n + i];
        if (ad < 0) ad = -ad;
        if (ad > hscale) hscale = ad;
        for (int j = i + 1; j < n; j++)
            if (H[i * n + j] > zero_tol || H[i * n + j] < -zero_tol) { is_diag = 0; break; }
    }
    if (is_diag) { /* utils.c:284-312 */
        real factor_tol = zero_tol;
        if (hscale > 0) factor_tol = zero_tol * hscale;
        w->RinvD = w->Rinv; w->Rinv = NULL;
        for (int i = 0; i < n; i++) {
            real Hi = H[i * n + i];
            if (Hi <= factor_tol) { (*n_prox)++; Hi += prox_eps_scaled(st, hscale); }
            if (Hi <= zero_tol) return EXIT_NONCONVEX;
            Hi = (real)sqrt(Hi);
            w->RinvD[i] = 1 / Hi;
            if (i < w->ms) w->scaling[i] = Hi;
        }
        return 1;
    }
    real *R = w->Rinv;
    for (int i = 0, p = 0; i < n; i++) { /* utils.c:319-323 */
        R[p++] = H[i * n + i];
        for (int j = i + 1; j < n; j++) R[p++] = (real)0.5 * (H[i * n + j] + H[j * n + i]);
    }
    real min_piv = ORC_INF, max_piv = 0;
    for (int i = 0; i < n; i++) { /* utils.c:337-352: upper Cholesky by rows, 1/r_ii on the diagonal */
        real *Ri = R + roff(i, n);
        real di = Ri[i];
        for (int k = 0; k < i; k++) { const real *Rk = R + roff(k, n); di -= Rk[i] * Rk[i]; }
        if (di <= zero_tol) goto singular;
        if (di < min_piv) min_piv = di;
        if (di > max_piv) max_piv = di;
        di = 1 / (real)sqrt(di);
        for (int j = i + 1; j < n; j++) {
            for (int k = 0; k < i; k++) { const real *Rk = R + roff(k, n); Ri[j] -= Rk[i] * Rk[j]; }
            Ri[j] *= di;
        }
        Ri[i] = di;
    }
    if (min_piv <= zero_tol * max_piv) goto singular;
    for (int k = 0; k < n; k++) { /* utils.c:380-389: R -> R^-1 in place */
        real *Rk = R + roff(k, n);
        for (int j = k + 1; j < n; j++) Rk[j] *= -Rk[k];
        for (int i = k + 1; i < n; i++) {
            const real *Ri = R + roff(i, n);
            Rk[i] *= Ri[i];
            for (int j = i + 1; j < n; j++) Rk[j] -= Ri[j] * Rk[i];
        }
    }
    return 1;
singular: /* utils.c:356-377: the reference shifts H and hands over to the proximal-point driver */
    if (prox_eps_scaled(st, hscale) <= 0) return EXIT_NONCONVEX;
    *n_prox = n;
    return 1;
}

/* utils.c:434-472 (mask has UPDATE_Rinv) followed by utils.c:586-613 */
static int form_M(Ldp *w, const OrcProblem *qp) {
    const int n = w->n, mA = w->m - w->ms;
    if (w->Rinv != NULL) {
        for (int r = 0; r < mA; r++) {
            const real *a = qp->A + (size_t)r * n;
            real *mr = w->M + (size_t)r * n;
            /* M[r][c] = sum_{i<=c} A[r][i] Rinv[i][c], accumulated from i=c down to 0 */
            for (int i = n - 1; i >= 0; i--) {
                const real *Ri = w->Rinv + roff(i, n);
                for (int c = n - 1; c > i; c--) mr[c] += Ri[c] * a[i];
                mr[i] = Ri[i] * a[i];
            }
        }
    } else {
        for (int r = 0; r < mA; r++)
            for (int c = 0; c < n; c++)
                w->M[(size_t)r * n + c] = qp->A[(size_t)r * n + c] * w->RinvD[c];
    }
    reset_ws(w);
    for (int i = w->ms; i < w->m; i++) { /* normalise rows */
        real *mr = w->M + (size_t)(i - w->ms) * n;
        real s = 0;
        for (int j = 0; j < n; j++) s += mr[j] * mr[j];
        if (s < w->st->zero_tol) {
            w->scaling[i] = 1.0;
            if (qp->bupper[i] < -w->st->zero_tol || qp->blower[i] > w->st->zero_tol)
                if (!(w->sense[i] & B_IMMUTABLE) && !(w->sense[i] & B_SOFT)) return EXIT_INFEASIBLE;
            w->sense[i] = B_IMMUTABLE;
            continue;
        }
        s = 1 / (real)sqrt(s);
        w->scaling[i] = s;
        for (int j = 0; j < n; j++) mr[j] *= s;
    }
    return 0;
}

static void ldp_free(Ldp *w) {
    free(w->M); free(w->Rinv ? w->Rinv : w->RinvD); free(w->v); free(w->dupper); free(w->dlower);
    free(w->scaling); free(w->Mu); free(w->sense); free(w->u); free(w->xunc); free(w->lam); free(w->lam_star);
    free(w->D); free(w->xl); free(w->zl); free(w->L); free(w->WS); free(w->pstack_id); free(w->pstack_lam);
}

void orc_default_settings(OrcSettings *s) { /* api.c:505-527, constants.h:15-29 */
    s->primal_tol = 1e-6; s->dual_tol = 1e-12; s->zero_tol = 1e-11; s->pivot_tol = 1e-6;
    s->progress_tol = 1e-14; s->cycle_tol = 10; s->iter_limit = 10000; s->fval_bound = ORC_INF;
    s->eps_prox = (real)-1e-6; s->eta_prox = -1.0; s->rho_soft = 1e-6; s->rel_subopt = 0; s->abs_subopt = 0;
    s->sing_tol = (real)3.7e-11; s->refactor_tol = 1e-9; s->time_limit = 0;
}

/* api.c:62-79 + setup_daqp_main api.c:93-160 + daqp_update_ldp utils.c:58-221 + daqp_solve api.c:8-59 +
 * daqp_extract_result api.c:455-495 */
void orc_quadprog(OrcResult *res, const OrcProblem *qp, const OrcSettings *settings, OrcTrace *trace) {
    OrcSettings defaults;
    Ldp W, *w = &W;
    const int n = qp->n, m = qp->m, ms = qp->ms, mA = m - ms;
    int ns = 0, nb = 0, flag, do_activate = 0, n_prox = 0, unc = 0;
    memset(w, 0, sizeof(W));
    if (settings == NULL) { orc_default_settings(&defaults); settings = &defaults; }
    res->setup_time = 0; res->solve_time = 0;
    if (trace) {
        trace->n_active = 0; trace->n_scan = trace->n_add = trace->n_remove = trace->n_csp = 0;
        trace->n_pivot = trace->n_refine = trace->n_refactor = trace->n_cycle = 0;
        trace->n_log = 0;
        w->oplog = trace->oplog; w->oplog_cap = trace->oplog_cap;
    }

    if (qp->sense != NULL)
        for (int i = 0; i < m; i++) { if (qp->sense[i] & B_SOFT) ns++; if (qp->sense[i] & B_BINARY) nb++; }
    if (qp->H == NULL && qp->f == NULL && nb == 0 && qp->nh <= 1 && qp->problem_type == 0 && n > 0) {
        /* the LDP min |x|^2 itself (utils.c:103-110 leaves Rinv == NULL, M = A): the identity Hessian walks the same
         * iterates bit for bit (unit Cholesky factor, products with 1.0) -- checked against the reference on the
         * ldp_* fixtures, tests/test_oracle.py */
        OrcProblem q2 = *qp;
        orc_real *eye = (orc_real *)calloc((size_t)n * n, sizeof(orc_real));
        for (int i = 0; i < n; i++) eye[(size_t)i * n + i] = 1;
        q2.H = eye;
        orc_quadprog(res, &q2, settings, trace);
        free(eye);
        return;
    }
    if (nb > 0 || qp->nh > 1 || qp->problem_type != 0 || qp->H == NULL) { res->exitflag = EXIT_UNSUPPORTED; return; }

    w->n = n; w->m = m; w->ms = ms; w->cap = n + ns + 1; w->st = settings;
    /* zeroed: after an exit inside the add branch (EXIT_CYCLE) the reference reports the entry of the entering row from
     * the multiplier buffer it has not written yet (api.c:463-466 reads lam_star after the swap of auxiliary.c:159-160) */
    w->lam = calloc(w->cap, sizeof(real)); w->lam_star = calloc(w->cap, sizeof(real));
    w->D = malloc(sizeof(real) * w->cap); w->xl = malloc(sizeof(real) * w->cap); w->zl = malloc(sizeof(real) * w->cap);
    w->L = malloc(sizeof(real) * (size_t)w->cap * w->cap); w->WS = malloc(sizeof(int) * w->cap);
    w->pstack_id = malloc(sizeof(int) * w->cap); w->pstack_lam = malloc(sizeof(real) * w->cap);
    w->u = calloc(n > 0 ? n : 1, sizeof(real)); w->xunc = malloc(sizeof(real) * (n > 0 ? n : 1));
    w->scaling = malloc(sizeof(real) * (m > 0 ? m : 1));
    for (int i = 0; i < ms; i++) w->scaling[i] = 1;
    w->M = calloc((size_t)(mA > 0 ? mA : 1) * n, sizeof(real));
    w->Mu = malloc(sizeof(real) * (mA > 0 ? mA : 1));
    w->dupper = malloc(sizeof(real) * (m > 0 ? m : 1)); w->dlower = malloc(sizeof(real) * (m > 0 ? m : 1));
    w->sense = malloc(sizeof(int) * (m > 0 ? m : 1));
    w->Rinv = malloc(sizeof(real) * ((size_t)n * (n + 1) / 2 + 1));
    w->v = qp->f ? calloc(n > 0 ? n : 1, sizeof(real)) : NULL;
    reset_ws(w);

    /* --- daqp_update_ldp with mask = Rinv|M|v|d|sense|unconstrained|eliminate (api.c:165-180, 67-68) --- */
    flag = 0;
    if (qp->sense == NULL) for (int i = 0; i < m; i++) w->sense[i] = 0;
    else { for (int i = 0; i < m; i++) w->sense[i] = qp->sense[i]; do_activate = 1; }
    for (int i = 0; i < m; i++) { /* utils.c:546-567 */
        if (w->sense[i] & B_IMMUTABLE) continue;
        real diff = qp->bupper[i] - qp->blower[i];
        if (diff < -settings->primal_tol) { flag = EXIT_INFEASIBLE; goto setup_failed; }
        else if (diff < settings->zero_tol && !(w->sense[i] & B_SOFT)) {
            w->sense[i] |= B_ACTIVE + B_IMMUTABLE;
            do_activate = 1;
        }
    }
    flag = factor_hessian(w, qp->H, &n_prox);
    if (flag < 0) goto setup_failed;
    if (n_prox > 0) { flag = EXIT_UNSUPPORTED; goto setup_failed; } /* reference: proximal-point driver */

    if (w->v != NULL) { /* utils.c:474-497 with UPDATE_Rinv: v = Rinv' f, accumulated from the last row up */
        if (w->Rinv == NULL) for (int i = 0; i < n; i++) w->v[i] = qp->f[i] * w->RinvD[i];
        else for (int j = n - 1; j >= 0; j--) {
            const real *Rj = w->Rinv + roff(j, n);
            for (int i = n - 1; i > j; i--) w->v[i] += Rj[i] * qp->f[j];
            w->v[j] = Rj[j] * qp->f[j];
        }
    }
    /* utils.c:618-687: try the unconstrained optimum (only when nothing is pre-activated / immutable) */
    unc = 1;
    for (int i = 0; i < m; i++) if (w->sense[i] & (B_ACTIVE + B_IMMUTABLE)) { unc = 0; break; }
    if (unc) {
        int feasible = 1;
        real *x = w->xunc;
        if (w->v != NULL) {
            if (w->Rinv != NULL)
                for (int i = 0; i < n; i++) x[i] = -dot1(w->Rinv + roff(i, n) + i, w->v + i, n - i);
            else for (int i = 0; i < n; i++) x[i] = -w->RinvD[i] * w->v[i];
        } else for (int i = 0; i < n; i++) x[i] = 0.0;
        for (int i = 0; i < ms; i++) {
            w->dupper[i] = qp->bupper[i] - x[i];
            w->dlower[i] = qp->blower[i] - x[i];
            if (w->dupper[i] < -settings->primal_tol || w->dlower[i] > settings->primal_tol) feasible = 0;
        }
        for (int i = ms; i < m; i++) {
            real s = dot1(qp->A + (size_t)(i - ms) * n, x, n);
            w->dupper[i] = qp->bupper[i] - s;
            w->dlower[i] = qp->blower[i] - s;
            if (w->dupper[i] < -settings->primal_tol || w->dlower[i] > settings->primal_tol) feasible = 0;
        }
        if (feasible) { reset_ws(w); w->unconstrained_optimal = 1; goto solve; }
    }
    flag = form_M(w, qp);
    if (flag < 0) goto setup_failed;
    if (w->Rinv != NULL) /* utils.c:569-585 */
        for (int i = 0; i < ms; i++) {
            real *Ri = w->Rinv + roff(i, n);
            real s = 0;
            for (int j = i; j < n; j++) s += Ri[j] * Ri[j];
            s = 1 / (real)sqrt(s);
            w->scaling[i] = s;
            for (int j = i; j < n; j++) Ri[j] *= s;
        }
    if (unc) { /* utils.c:151-158 */
        for (int i = 0; i < m; i++) { w->dupper[i] *= w->scaling[i]; w->dlower[i] *= w->scaling[i]; }
        w->reuse = 0;
    } else { /* utils.c:499-544 */
        w->reuse = 0;
        for (int i = 0; i < m; i++) {
            w->dupper[i] = qp->bupper[i] * w->scaling[i];
            w->dlower[i] = qp->blower[i] * w->scaling[i];
        }
        if (w->v != NULL) {
            for (int i = 0; i < ms; i++) {
                real s = w->Rinv ? dot1(w->Rinv + roff(i, n) + i, w->v + i, n - i) : w->v[i];
                w->dupper[i] += s; w->dlower[i] += s;
            }
            for (int i = ms; i < m; i++) {
                real s = dot1(w->M + (size_t)(i - ms) * n, w->v, n);
                w->dupper[i] += s; w->dlower[i] += s;
            }
        }
    }
    if (do_activate) { /* utils.c:199-211 */
        reset_ws(w);
        flag = activate_constraints(w);
        if (flag < 0) goto setup_failed;
    }

solve: /* api.c:8-59 */
    if (!w->unconstrained_optimal) {
        res->exitflag = ldp_solve(w);
        if (res->exitflag > 0) { /* daqp.c:111-139 */
            real *x = w->u;
            if (w->v != NULL) for (int i = 0; i < n; i++) x[i] = w->u[i] - w->v[i];
            if (w->Rinv != NULL) {
                for (int i = 0; i < n; i++) {
                    const real *Ri = w->Rinv + roff(i, n);
                    x[i] *= Ri[i];
                    for (int j = i + 1; j < n; j++) x[i] += Ri[j] * x[j];
                }
                for (int i = 0; i < ms; i++) x[i] /= w->scaling[i];
            } else for (int i = 0; i < n; i++) x[i] *= w->RinvD[i];
            for (int i = 0; i < w->k; i++) w->lam_star[i] *= w->scaling[w->WS[i]];
        }
    } else {
        w->iterations = 1; w->fval = 0; w->soft_slack = 0;
        res->exitflag = EXIT_OPTIMAL;
    }
    { /* api.c:455-495 */
        const real *x = w->unconstrained_optimal ? w->xunc : w->u;
        for (int i = 0; i < n; i++) res->x[i] = x[i];
        if (res->lam != NULL) {
            for (int i = 0; i < m; i++) res->lam[i] = 0;
            for (int i = 0; i < w->k; i++) res->lam[w->WS[i]] = w->lam_star[i];
        }
        if (w->v != NULL) {
            res->fval = w->fval;
            for (int i = 0; i < n; i++) res->fval -= w->v[i] * w->v[i];
            res->fval *= 0.5;
        }
        res->soft_slack = w->soft_slack;
        res->iter = w->iterations;
        res->nodes = 1;
    }
    if (trace) {
        trace->n_active = w->k;
        if (trace->ws) for (int i = 0; i < w->k; i++) trace->ws[i] = w->WS[i];
        if (trace->sense_out) for (int i = 0; i < m; i++) trace->sense_out[i] = w->sense[i];
        trace->n_scan = w->n_scan; trace->n_add = w->n_add; trace->n_remove = w->n_remove; trace->n_csp = w->n_csp;
        trace->n_pivot = w->n_pivot; trace->n_refine = w->n_refine; trace->n_refactor = w->n_refactor; trace->n_cycle = w->n_cycle;
        trace->n_log = w->n_log;
    }
    ldp_free(w);
    return;

setup_failed: /* api.c:69-72: exit flag only, res->x untouched */
    res->exitflag = flag;
    ldp_free(w);
}

/* ---- warm-start initialisers: api.c:577-631 ----------------------------------------------------------------------- */
/* api.c:579-617: tol = 1e-9; simple bounds compare x[i], general rows the left-to-right sum of factorization.h:13-17 */
void orc_primal_init_active(OrcProblem *qp, const orc_real *x) {
    const real tol = (real)1e-9;
    for (int i = 0; i < qp->m; i++) {
        real ax, slack;
        if (qp->sense[i] & B_IMMUTABLE) continue;
        ax = i < qp->ms ? x[i] : dot1(x, qp->A + (size_t)(i - qp->ms) * qp->n, qp->n);
        slack = ax - qp->bupper[i];
        if (slack < tol && slack > -tol) { qp->sense[i] |= B_ACTIVE; qp->sense[i] &= ~B_LOWER; }
        else {
            slack = ax - qp->blower[i];
            if (slack < tol && slack > -tol) qp->sense[i] |= B_ACTIVE + B_LOWER;
        }
    }
}

/* api.c:620-631: tol = 1e-12 */
void orc_dual_init_active(OrcProblem *qp, const orc_real *lam) {
    const real tol = (real)1e-12;
    for (int i = 0; i < qp->m; i++) {
        if (qp->sense[i] & B_IMMUTABLE) continue;
        if (lam[i] > tol) { qp->sense[i] |= B_ACTIVE; qp->sense[i] &= ~B_LOWER; }
        else if (lam[i] < -tol) qp->sense[i] |= B_ACTIVE + B_LOWER;
    }
}

/* ---- minimal representation: daqp_minrep (api.c:531-556) + daqp_minrep_work (utils.c:808-835) ------------- */
/* The reference wraps the polyhedron in a bare workspace: M = A as given (no normalisation, scaling == NULL so the
 * violation threshold is -primal_tol itself, auxiliary.c:110,136), Rinv == NULL (simple bounds are unit rows),
 * v == NULL, dlower = -inf, sense = 0, default settings. scaling == NULL is restated as scaling = 1 (ep * 1 == ep). */
static void minrep_open(Ldp *w, OrcSettings *st, const real *A, const real *b, int n, int m, int ms) {
    const int mA = m - ms;
    memset(w, 0, sizeof(*w));
    orc_default_settings(st);
    w->n = n; w->m = m; w->ms = ms; w->cap = n + 1; w->st = st;
    /* zeroed: after an exit inside the add branch (EXIT_CYCLE) the reference reports the entry of the entering row from
     * the multiplier buffer it has not written yet (api.c:463-466 reads lam_star after the swap of auxiliary.c:159-160) */
    w->lam = calloc(w->cap, sizeof(real)); w->lam_star = calloc(w->cap, sizeof(real));
    w->D = malloc(sizeof(real) * w->cap); w->xl = malloc(sizeof(real) * w->cap); w->zl = malloc(sizeof(real) * w->cap);
    w->L = malloc(sizeof(real) * (size_t)w->cap * w->cap); w->WS = malloc(sizeof(int) * w->cap);
    w->pstack_id = malloc(sizeof(int) * w->cap); w->pstack_lam = malloc(sizeof(real) * w->cap);
    w->u = calloc(n, sizeof(real)); w->xunc = malloc(sizeof(real) * n);
    w->scaling = malloc(sizeof(real) * m);
    w->M = malloc(sizeof(real) * (size_t)(mA > 0 ? mA : 1) * n);
    w->Mu = malloc(sizeof(real) * (mA > 0 ? mA : 1));
    w->dupper = malloc(sizeof(real) * m); w->dlower = malloc(sizeof(real) * m);
    w->sense = malloc(sizeof(int) * m);
    memcpy(w->M, A, sizeof(real) * (size_t)mA * n);
    for (int i = 0; i < m; i++) { w->scaling[i] = 1; w->dupper[i] = b[i]; w->dlower[i] = -ORC_INF; w->sense[i] = 0; }
    reset_ws(w);
}

/* One pass of the loop body of daqp_minrep_work (utils.c:816-822): constraint i as an active equality, then daqp_ldp */
static int minrep_probe(Ldp *w, int i) {
    reset_ws(w);
    w->sense[i] = B_ACTIVE + B_IMMUTABLE;
    add_constraint(w, i, (real)1);
    return ldp_solve(w);
}

/* auxiliary.c:482-488 */
static void deactivate_all(Ldp *w) {
    for (int j = 0; j < w->k; j++)
        if (!(w->sense[w->WS[j]] & B_IMMUTABLE)) w->sense[w->WS[j]] &= ~B_ACTIVE;
}

/* Reference order: one constraint after the other; a redundant constraint stays IMMUTABLE (ignored from then on), a
 * constraint active at an optimum is marked non-redundant and never probed (utils.c:808-835). */
void orc_minrep(int *is_redundant, const orc_real *A, const orc_real *b, int n, int m, int ms) {
    Ldp W, *w = &W;
    OrcSettings st;
    if (m <= 0) return;
    minrep_open(w, &st, A, b, n, m, ms);
    for (int i = 0; i < m; i++) is_redundant[i] = -1;
    for (int i = 0; i < m; i++) {
        if (is_redundant[i] != -1 || (w->sense[i] & B_IMMUTABLE)) continue;
        int flag = minrep_probe(w, i);
        if (flag == EXIT_INFEASIBLE) {
            is_redundant[i] = 1;
            w->sense[i] &= ~B_ACTIVE;
        } else {
            is_redundant[i] = 0;
            w->sense[i] &= ~B_IMMUTABLE;
            if (flag == EXIT_OPTIMAL)
                for (int j = 0; j < w->k; j++) is_redundant[w->WS[j]] = 0;
        }
        deactivate_all(w);
    }
    ldp_free(w);
}

/* The same m probes, each against the FULL polyhedron (no constraint dropped, none skipped): what the batched GPU
 * path computes. Returns exit flag and iteration count of every LDP; is_redundant[i] = (flag == INFEASIBLE). */
void orc_minrep_independent(int *is_redundant, int *exitflag, int *iter, const orc_real *A, const orc_real *b,
                            int n, int m, int ms) {
    Ldp W, *w = &W;
    OrcSettings st;
    if (m <= 0) return;
    minrep_open(w, &st, A, b, n, m, ms);
    for (int i = 0; i < m; i++) {
        for (int j = 0; j < m; j++) w->sense[j] = 0;
        int flag = minrep_probe(w, i);
        is_redundant[i] = flag == EXIT_INFEASIBLE ? 1 : 0;
        if (exitflag) exitflag[i] = flag;
        if (iter) iter[i] = w->iterations;
    }
    ldp_free(w);
}

typedef struct {
    int N, n, m, ms;
    const orc_real *H, *f, *A, *bupper, *blower;
    const int *sense;
    const OrcSettings *settings;
    orc_real *x, *lam, *fval;
    int *exitflag, *iter, *trace_counts;
    int next; /* shared chunk counter */
} PackedJob;

static void packed_one(PackedJob *J, int p) {
    const int n = J->n, m = J->m, ms = J->ms;
    const size_t mA = (size_t)(m - ms);
    OrcProblem qp;
    OrcResult r;
    OrcTrace tr;
    memset(&tr, 0, sizeof(tr));
    qp.n = n; qp.m = m; qp.ms = ms;
    qp.H = (orc_real *)J->H + (size_t)p * n * n;
    qp.f = J->f ? (orc_real *)J->f + (size_t)p * n : NULL;
    qp.A = (orc_real *)J->A + (size_t)p * mA * n;
    qp.bupper = (orc_real *)J->bupper + (size_t)p * m;
    qp.blower = (orc_real *)J->blower + (size_t)p * m;
    qp.sense = J->sense ? (int *)J->sense + (size_t)p * m : NULL;
    qp.break_points = NULL; qp.nh = 0; qp.problem_type = 0;
    r.x = J->x + (size_t)p * n;
    r.lam = J->lam ? J->lam + (size_t)p * m : NULL;
    r.fval = 0; r.soft_slack = 0; r.iter = 0; r.nodes = 0;
    orc_quadprog(&r, &qp, J->settings, J->trace_counts ? &tr : NULL);
    if (J->fval) J->fval[p] = r.fval;
    J->exitflag[p] = r.exitflag;
    if (J->iter) J->iter[p] = r.iter;
    if (J->trace_counts) {
        int *tc = J->trace_counts + 8 * (size_t)p;
        tc[0] = tr.n_scan; tc[1] = tr.n_add; tc[2] = tr.n_remove; tc[3] = tr.n_csp;
        tc[4] = tr.n_pivot; tc[5] = tr.n_refine; tc[6] = tr.n_refactor; tc[7] = tr.n_cycle;
    }
}

static void *packed_worker(void *arg) {
    PackedJob *J = (PackedJob *)arg;
    for (;;) {
        int start = __atomic_fetch_add(&J->next, 16, __ATOMIC_RELAXED);
        if (start >= J->N) break;
        int end = start + 16 < J->N ? start + 16 : J->N;
        for (int p = start; p < end; p++) packed_one(J, p);
    }
    return NULL;
}

double orc_solve_packed(int N, int n, int m, int ms,
                        const orc_real *H, const orc_real *f, const orc_real *A,
                        const orc_real *bupper, const orc_real *blower, const int *sense,
                        const OrcSettings *settings,
                        orc_real *x, orc_real *lam, orc_real *fval, int *exitflag, int *iter,
                        int *trace_counts, int nthreads) {
    struct timespec t0, t1;
    PackedJob J = {N, n, m, ms, H, f, A, bupper, blower, sense, settings, x, lam, fval, exitflag, iter, trace_counts, 0};
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (nthreads <= 1) {
        for (int p = 0; p < N; p++) packed_one(&J, p);
    } else {
        pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
        for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, packed_worker, &J);
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
        free(th);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
