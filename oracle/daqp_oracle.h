/* oracle/daqp_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (own code, scalar C) of the reference's dual active-set hot path:
 *   daqp_quadprog -> QP->LDP transform -> daqp_ldp loop -> ldp2qp_solution -> daqp_extract_result
 * (reference: src/api.c:62-79, src/utils.c:58-687, src/daqp.c:6-139, src/auxiliary.c, src/factorization.c).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product (daqp_b200/) never links or calls it.
 *
 * Parity pinning: tests/test_oracle.py checks this file against (1) the reference's known-answer tests,
 * (2) the reference itself compiled from /root/reference into oracle/_ref (bit-exact against the
 * -O2 -ffp-contract=off build), (3) committed golden vectors under tests/golden/.
 *
 * Compile with -DORC_SINGLE for the c_float=float variant (reference: include/types.h:8-12).
 */
#ifndef DAQP_ORACLE_H
#define DAQP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#ifdef ORC_SINGLE
typedef float orc_real;
#else
typedef double orc_real;
#endif

/* Same field order / layout as DAQPProblem (include/types.h:32-49). */
typedef struct {
    int n, m, ms;
    orc_real *H, *f, *A, *bupper, *blower;
    int *sense;
    int *break_points;
    int nh;
    int problem_type;
} OrcProblem;

/* Same layout as DAQPSettings (include/types.h:52-74). */
typedef struct {
    orc_real primal_tol, dual_tol, zero_tol, pivot_tol, progress_tol;
    int cycle_tol, iter_limit;
    orc_real fval_bound;
    orc_real eps_prox, eta_prox;
    orc_real rho_soft;
    orc_real rel_subopt, abs_subopt;
    orc_real sing_tol, refactor_tol, time_limit;
} OrcSettings;

/* Same layout as DAQPResult (include/api.h:15-27). */
typedef struct {
    orc_real *x, *lam;
    orc_real fval, soft_slack;
    int exitflag, iter, nodes;
    orc_real solve_time, setup_time;
} OrcResult;

/* Extra observability the reference only exposes through its workspace: final working set + op counts. */
typedef struct {
    int n_active;     /* final |WS| */
    int *ws;          /* caller buffer, >= n+ns+1 ints (may be NULL) */
    int *sense_out;   /* caller buffer, m ints: final sense bits (may be NULL) */
    int n_scan;       /* calls of the feasibility scan (daqp_add_infeasible)        */
    int n_add;        /* LDL row appends (daqp_update_LDL_add)                        */
    int n_remove;     /* LDL row deletions (daqp_update_LDL_remove incl. last-row)    */
    int n_csp;        /* CSP solves                                                    */
    /* rare control paths of daqp_ldp (so that a test can assert they were taken) */
    int n_pivot;      /* daqp_pivot_last swaps (auxiliary.c:379-396)                   */
    int n_refine;     /* daqp_refine_active calls (daqp.c:52-56)                       */
    int n_refactor;   /* refactor-on-exit repairs (daqp.c:33-46)                       */
    int n_cycle;      /* cycle-guard repairs (daqp.c:67-81)                            */
    /* decision log (debugging aid, same codes as DAQPB200Diag.trace): (code, value) pairs -- 1 add (2 row + lower),
     * 2 remove (row), 3 refactor, 4 refine, 5 cycle repair, 7 exit (flag) */
    int *oplog;       /* caller buffer of 2 * oplog_cap ints, or NULL                  */
    int oplog_cap;
    int n_log;        /* decisions made (may exceed oplog_cap)                         */
} OrcTrace;

void orc_default_settings(OrcSettings *s);

/* Drop-in equivalent of daqp_quadprog() restricted to the hot-path scope (dense H>0, no binaries/hierarchy/
 * AVI/prox). Out-of-scope inputs return exitflag -8. trace may be NULL. */
void orc_quadprog(OrcResult *res, const OrcProblem *qp, const OrcSettings *settings, OrcTrace *trace);

/* Loop orc_quadprog over a strided homogeneous batch (same layout as daqp_b200_solve_packed); used by bench.py's
 * cpu_baseline leg so the timed loop has no Python in it. nthreads>1 uses pthreads (dynamic chunks of 16 problems).
 * sense may be NULL. x:[N][n] lam:[N][m] fval:[N] exitflag:[N] iter:[N]. Returns wall seconds. */
double orc_solve_packed(int N, int n, int m, int ms,
                        const orc_real *H, const orc_real *f, const orc_real *A,
                        const orc_real *bupper, const orc_real *blower, const int *sense,
                        const OrcSettings *settings,
                        orc_real *x, orc_real *lam, orc_real *fval, int *exitflag, int *iter,
                        int *trace_counts /* [N][8] scan,add,remove,csp,pivot,refine,refactor,cycle or NULL */,
                        int nthreads);

/* daqp_primal_init_active / daqp_dual_init_active (api.c:577-631): warm-start bits of qp->sense from an iterate. */
void orc_primal_init_active(OrcProblem *qp, const orc_real *x);
void orc_dual_init_active(OrcProblem *qp, const orc_real *lam);

/* daqp_minrep (api.c:531-556, utils.c:808-835): is_redundant[m] for {x : [I(ms); A] x <= b}, constraints probed one
 * after the other exactly like the reference. */
void orc_minrep(int *is_redundant, const orc_real *A, const orc_real *b, int n, int m, int ms);
/* The same probes run independently of each other (nothing dropped, nothing skipped): the batched GPU semantics.
 * exitflag / iter may be NULL. */
void orc_minrep_independent(int *is_redundant, int *exitflag, int *iter, const orc_real *A, const orc_real *b,
                            int n, int m, int ms);

#ifdef __cplusplus
}
#endif
#endif
