/* oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY.
 * Thin pthread loop around the UNMODIFIED reference's daqp_quadprog() (linked from oracle/_ref/libdaqp_ref.so) over
 * a packed homogeneous batch, so bench.py's CPU legs time the reference itself with no Python in the loop.
 * Own code; the only reference symbol used is daqp_quadprog (include/api.h:30). Struct layouts: daqp_oracle.h. */
#include "daqp_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <time.h>

extern void daqp_quadprog(OrcResult *res, OrcProblem *qp, OrcSettings *settings);

typedef struct {
    int N, n, m, ms;
    const orc_real *H, *f, *A, *bupper, *blower;
    const int *sense;
    OrcSettings *settings;
    orc_real *x, *lam, *fval;
    int *exitflag, *iter;
    int next;
} Job;

static void one(Job *J, int p) {
    const int n = J->n, m = J->m, ms = J->ms;
    const size_t mA = (size_t)(m - ms);
    OrcProblem qp;
    OrcResult r;
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
    r.fval = 0; r.soft_slack = 0; r.exitflag = 0; r.iter = 0; r.nodes = 0;
    daqp_quadprog(&r, &qp, J->settings);
    if (J->fval) J->fval[p] = r.fval;
    J->exitflag[p] = r.exitflag;
    if (J->iter) J->iter[p] = r.iter;
}

static void *worker(void *arg) {
    Job *J = (Job *)arg;
    for (;;) {
        int start = __atomic_fetch_add(&J->next, 16, __ATOMIC_RELAXED);
        if (start >= J->N) break;
        int end = start + 16 < J->N ? start + 16 : J->N;
        for (int p = start; p < end; p++) one(J, p);
    }
    return NULL;
}

/* Returns wall seconds (CLOCK_MONOTONIC around the whole loop, as BASELINE.md §3.3). */
double ref_solve_packed(int N, int n, int m, int ms, const orc_real *H, const orc_real *f, const orc_real *A,
                        const orc_real *bupper, const orc_real *blower, const int *sense, OrcSettings *settings,
                        orc_real *x, orc_real *lam, orc_real *fval, int *exitflag, int *iter, int nthreads) {
    struct timespec t0, t1;
    Job J = {N, n, m, ms, H, f, A, bupper, blower, sense, settings, x, lam, fval, exitflag, iter, 0};
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (nthreads <= 1) {
        for (int p = 0; p < N; p++) one(&J, p);
    } else {
        pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
        for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, worker, &J);
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
        free(th);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
