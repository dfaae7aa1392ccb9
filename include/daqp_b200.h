/* include/daqp_b200.h -- C ABI of the B200-native batched dual active-set QP engine.
 *
 * Drop-in boundary. The three structs below have the SAME field order and layout as the reference
 * (DAQPProblem: reference include/types.h:14-50, DAQPSettings: include/types.h:52-74, DAQPResult:
 * include/api.h:15-27), so the reference's language interfaces (Cython daqp.pxd:50-88, Julia types.jl:46-177,
 * Eigen daqp.cpp:111-112, MATLAB daqpmex.c) bind to this library unchanged for the calls listed here.
 *
 *   daqp_quadprog()          replaces reference include/api.h:30 / src/api.c:62-79 (runs a batch of one on the GPU)
 *   daqp_default_settings()  replaces reference include/api.h:52 / src/api.c:505-527
 *   daqp_quadprog_batch()    NEW: N independent daqp_quadprog() calls in one launch (the reference has no batch API)
 *   daqp_b200_solve_packed() NEW: same for a homogeneous batch in strided host arrays (no per-problem pointers)
 *   daqp_b200_solve_device() NEW: same with device-resident arrays, asynchronous on a caller stream
 *   daqp_minrep()            replaces reference include/api.h:54 / src/api.c:531-556 (its m LDPs run concurrently)
 *   daqp_b200_minrep_batch() NEW: P polyhedra per call
 *
 * Exit flags are the reference's (include/constants.h:42-51). Problems outside the hot-path scope (binary
 * constraints, hierarchies, AVI, H == NULL, singular H that needs the proximal-point driver) return
 * DAQP_EXIT_UNSUPPORTED (-8); there is no CPU fallback inside this library.
 *
 * Arithmetic is fp64 (c_float = double), like the reference's default build; the *_f32 entry points run the same
 * kernels in fp32.
 */
#ifndef DAQP_B200_H
#define DAQP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double c_float;

#define DAQP_INF ((c_float)1e30)

/* exit flags -- reference include/constants.h:42-51 */
#define DAQP_EXIT_SOFT_OPTIMAL 2
#define DAQP_EXIT_OPTIMAL 1
#define DAQP_EXIT_INFEASIBLE -1
#define DAQP_EXIT_CYCLE -2
#define DAQP_EXIT_UNBOUNDED -3
#define DAQP_EXIT_ITERLIMIT -4
#define DAQP_EXIT_NONCONVEX -5
#define DAQP_EXIT_OVERDETERMINED_INITIAL -6
#define DAQP_EXIT_TIMELIMIT -7
#define DAQP_EXIT_UNSUPPORTED -8

/* constraint sense bits -- reference include/constants.h:64-96 */
#define DAQP_ACTIVE 1
#define DAQP_LOWER 2
#define DAQP_IMMUTABLE 4
#define DAQP_SOFT 8
#define DAQP_BINARY 16

/* min 0.5 x'Hx + f'x  s.t.  blower <= [I(ms); A] x <= bupper  -- reference include/types.h:14-50 */
typedef struct {
    int n;  /* number of variables */
    int m;  /* number of constraints, simple bounds first */
    int ms; /* number of simple bounds */
    c_float* H; /* n x n row-major */
    c_float* f; /* n, may be NULL */
    c_float* A; /* (m-ms) x n row-major */
    c_float* bupper; /* m */
    c_float* blower; /* m */
    int* sense; /* m bit flags, may be NULL */
    int* break_points; /* hierarchical QP: unsupported here */
    int nh;
    int problem_type; /* 0 QP; 1 AVI / 2 prefactored H: unsupported here */
} DAQPProblem;

/* reference include/types.h:52-74 */
typedef struct {
    c_float primal_tol;
    c_float dual_tol;
    c_float zero_tol;
    c_float pivot_tol;
    c_float progress_tol;
    int cycle_tol;
    int iter_limit;
    c_float fval_bound;
    c_float eps_prox;
    c_float eta_prox;
    c_float rho_soft;
    c_float rel_subopt;
    c_float abs_subopt;
    c_float sing_tol;
    c_float refactor_tol;
    c_float time_limit; /* seconds per problem, 0 = none: the clock starts when a warp takes the problem and is read every
                           32nd iteration (daqp.c:95-103) -> DAQP_EXIT_TIMELIMIT; B&B also checks it once per wave.
                           fp64 entries only (the fp32 entries ignore it) */
} DAQPSettings;

/* reference include/api.h:15-27 */
typedef struct {
    c_float* x;   /* n, caller-owned */
    c_float* lam; /* m, caller-owned, may be NULL */
    c_float fval;
    c_float soft_slack;
    int exitflag;
    int iter;
    int nodes;
    c_float solve_time; /* seconds: device time of the batch this problem was part of */
    c_float setup_time;
} DAQPResult;

/* update masks of daqp_update_ldp -- reference include/constants.h:54-61 */
#define DAQP_UPDATE_Rinv 1
#define DAQP_UPDATE_M 2
#define DAQP_UPDATE_v 4
#define DAQP_UPDATE_d 8
#define DAQP_UPDATE_sense 16
#define DAQP_UPDATE_hierarchy 32
#define DAQP_UPDATE_unconstrained 64
#define DAQP_UPDATE_eliminate 128

/* reference include/types.h:187-264 (SOFT_WEIGHTS off, the default build): same field order, sizes and offsets, because
 * the reference's interfaces allocate this struct themselves (Cython daqp.pyx:264, Eigen daqp.cpp:148) and read fields
 * by name or by offset (Julia api.jl:444-457). The LDP and the LDL' factor of a workspace live on the GPU; of the
 * pointer members this library fills the ones interfaces read back -- x, lam, lam_star, WS, sense, scaling (host
 * copies refreshed by daqp_solve) -- and leaves the rest NULL. The bnb / avi / eq slots hold this library's own
 * bookkeeping (those drivers are out of scope here). */
typedef struct {
    DAQPProblem* qp;
    int n, m, ms;
    c_float *M, *dupper, *dlower, *Rinv, *v;
    int* sense;
    c_float *scaling, *RinvD;
    c_float *x, *xold, *lam, *lam_star, *u;
    c_float fval;
    c_float *L, *D, *xldl, *zldl;
    int reuse_ind;
    int* WS;
    int n_active;
    int iterations;
    int sing_ind;
    int* prox_mask;
    int n_prox;
    c_float soft_slack;
    DAQPSettings* settings;
    void* bnb;          /* reference: DAQPBnB*            */
    int nh;
    int* break_points;
    void* avi;          /* reference: DAQPAVI*;    here: host-side bookkeeping of the workspace          */
    void* eq;           /* reference: DAQPEqElim*; here: the device-resident batch-of-one workspace     */
    void* timer;
    c_float* Mu;
} DAQPWorkspace;

/* ---- drop-in entry points ------------------------------------------------------------------------------- */

/* reference include/api.h:30 (src/api.c:62-79). settings == NULL selects the defaults. */
void daqp_quadprog(DAQPResult* res, DAQPProblem* qp, DAQPSettings* settings);

/* reference include/api.h:52 (src/api.c:505-527) */
void daqp_default_settings(DAQPSettings* settings);

/* ---- drop-in workspace flow: what daqp.Model / the Eigen class DAQP / Julia's Model call ----------------------
 * reference include/api.h:33-49, src/api.c:8-59,84-160,244-275,399-431, include/utils.h:11, src/utils.c:58-221.
 * A workspace is a batch of one on the GPU: setup_daqp runs the QP -> LDP transform and keeps the LDP on the device,
 * daqp_update_ldp(DAQP_UPDATE_v / _d) recomputes v and d there, daqp_solve continues from the factor and working set of
 * the previous solve. Masks that change the matrices or the sense bits (Rinv, M, sense) redo the transform. With
 * DAQP_UPDATE_unconstrained in init_mask (the mask daqp_quadprog itself uses, src/api.c:67-68) setup_daqp_main + daqp_solve
 * together ARE one daqp_quadprog call and run as one. Return values and the ownership of `settings` follow the
 * reference: 1 on success, the negative exit flag on failure (then the workspace is already freed, and a settings
 * struct the caller installed is left to the caller). */
int setup_daqp(DAQPProblem* qp, DAQPWorkspace* work, c_float* setup_time);
int setup_daqp_main(DAQPProblem* qp, DAQPWorkspace* work, c_float* setup_time, int init_mask);
void daqp_solve(DAQPResult* res, DAQPWorkspace* work);
int daqp_update_ldp(const int mask, DAQPWorkspace* work, DAQPProblem* qp);
void allocate_daqp_settings(DAQPWorkspace* work);
void free_daqp_workspace(DAQPWorkspace* work);
void free_daqp_ldp(DAQPWorkspace* work);
void daqp_set_primal_start(DAQPWorkspace* work, c_float* x);
/* hand-filled workspaces (Julia: api.jl:428-459) and the pieces of daqp_solve the interfaces call one by one:
 * allocate_daqp_workspace (src/api.c:295-340: host iterates only -- the factor lives on the device), daqp_ldp
 * (src/daqp.c:6-108: on a workspace from setup_daqp the solve of daqp_solve; on a hand-filled one -- M, dupper, dlower,
 * sense, m, ms set by the caller, Rinv == NULL -- a batch of one through daqp_b200_ldp_batch; work->u, lam_star, WS,
 * n_active, fval, iterations, sense are filled in), reset_daqp_workspace + daqp_deactivate_constraints (src/daqp.c:142-146,
 * src/auxiliary.c:482-488: the next solve starts cold), daqp_extract_result (src/api.c:455-495). */
void allocate_daqp_workspace(DAQPWorkspace* work, int n, int ns);
int daqp_ldp(DAQPWorkspace* work);
void reset_daqp_workspace(DAQPWorkspace* work);
void daqp_deactivate_constraints(DAQPWorkspace* work);
void daqp_extract_result(DAQPResult* res, DAQPWorkspace* work);
/* reference include/api.h:55 (src/api.c:562-574): index of the first constraint x violates by more than tol, or m */
int daqp_first_violating(c_float* x, c_float* A, c_float* bu, c_float* bl, int n, int m, int ms, c_float tol);

/* ---- batch entry points (new) --------------------------------------------------------------------------- */

/* N independent problems, identical in effect to N calls of daqp_quadprog(&res[i], &qps[i], settings).
 * Problems may differ in (n, m, ms): they are grouped by shape and the groups run side by side on a few lanes
 * (sub-engines with their own streams and pinned staging), largest estimated cost first. Returns 0, or a negative
 * CUDA-side error code (then no result field is valid). */
int daqp_quadprog_batch(int N, DAQPProblem* qps, DAQPResult* res, DAQPSettings* settings);

/* The same for callers of the reference's SINGLE-PRECISION build (-DDAQP_SINGLE_PRECISION, include/types.h:8-12): the
 * two structs with c_float = float (settings stay the double struct and are rounded to float like the reference's
 * constants). Arithmetic is fp32 end to end; plain inequality / equality / warm-start path (no soft constraints). */
typedef struct {
    int n, m, ms;
    float *H, *f, *A, *bupper, *blower;
    int* sense;
    int* break_points;
    int nh;
    int problem_type;
} DAQPProblemF32;
typedef struct {
    float *x, *lam;
    float fval, soft_slack;
    int exitflag, iter, nodes;
    float solve_time, setup_time;
} DAQPResultF32;
int daqp_quadprog_batch_f32(int N, DAQPProblemF32* qps, DAQPResultF32* res, DAQPSettings* settings);

typedef struct DAQPB200Handle DAQPB200Handle;

/* Engine bound to one CUDA device: owns streams and device scratch that is reused across calls.
 * device < 0 selects the current device. Returns 0 on success. */
int daqp_b200_create(DAQPB200Handle** out, int device);
void daqp_b200_destroy(DAQPB200Handle* h);

/* Optional per-problem diagnostics of the last solve (device pointers for solve_device, host for solve_packed).
 * Any member may be NULL. */
typedef struct {
    int* n_active;  /* [N]          final size of the working set                         */
    int* ws;        /* [N][n+ns+1]  final working set in factor order; ns = the largest number of soft
                                    constraints (sense & DAQP_SOFT) any problem of the batch carries, 0 without sense */
    int* counts;    /* [N][8]       feasibility scans, LDL adds, LDL removes, CSP solves, then the rare paths of
                                    daqp_ldp: pivot swaps (auxiliary.c:379-396), refinements (daqp.c:52-56),
                                    refactor-on-exit repairs (daqp.c:33-46), cycle-guard repairs (daqp.c:67-81) */
    unsigned char* sense; /* [N][ldm], ldm = m rounded up to 4: final sense bits          */
    c_float* soft_slack;  /* [N]    DAQPResult.soft_slack (reference src/api.c:469)       */
    int* trace;     /* [N][1 + 2 trace_cap] decision log, a debugging aid (device pointer for solve_device only; the
                       host entries ignore it): entry 0 = number of decisions, then (code, value) pairs in order --
                       1 add (2 row + lower), 2 remove (row), 3 refactor, 4 refine, 5 cycle repair, 7 exit (flag), 8 / 9 high and
                       low word of the objective the cycle guard compares after an add (daqp.c:67)  */
    int trace_cap;  /* decisions the log has room for (0 without a log)                   */
} DAQPB200Diag;

/* Homogeneous batch, strided HOST arrays: H[N][n][n], f[N][n] (or NULL), A[N][m-ms][n], bupper/blower[N][m],
 * sense[N][m] (or NULL); outputs x[N][n], lam[N][m] (or NULL), fval[N], exitflag[N], iter[N].
 * Host<->device copies are pipelined with the solve in chunks; pinned host memory makes them asynchronous.
 * h == NULL uses a process-wide default engine on the current device. Blocks until the results are in place. */
int daqp_b200_solve_packed(DAQPB200Handle* h, int N, int n, int m, int ms,
                           const c_float* H, const c_float* f, const c_float* A,
                           const c_float* bupper, const c_float* blower, const int* sense,
                           const DAQPSettings* settings,
                           c_float* x, c_float* lam, c_float* fval, int* exitflag, int* iter,
                           const DAQPB200Diag* diag);

/* The same batch on SEVERAL GPUs of this process: the batch is cut into ndev contiguous blocks, one host thread and one
 * engine per device, no data exchanged between devices (problems are independent). devices == NULL selects
 * 0 .. ndev-1; ndev <= 0 selects every visible device. seconds_per_device ([ndev], may be NULL) receives the wall time
 * each device's block took, copies included. Pinned host arrays (daqp_b200_pin) keep the copies of the devices apart. */
int daqp_b200_solve_packed_multi(int ndev, const int* devices, int N, int n, int m, int ms,
                                 const c_float* H, const c_float* f, const c_float* A,
                                 const c_float* bupper, const c_float* blower, const int* sense,
                                 const DAQPSettings* settings,
                                 c_float* x, c_float* lam, c_float* fval, int* exitflag, int* iter,
                                 double* seconds_per_device);
/* Page-lock / release a caller array (cudaHostRegister / cudaHostUnregister) for callers without a CUDA toolchain. */
int daqp_b200_pin(void* ptr, size_t bytes);
int daqp_b200_unpin(void* ptr);

/* Same with DEVICE arrays; enqueues on `stream` (a cudaStream_t; NULL = the engine's own stream) and returns
 * without synchronising. */
int daqp_b200_solve_device(DAQPB200Handle* h, int N, int n, int m, int ms,
                           const c_float* dH, const c_float* df, const c_float* dA,
                           const c_float* dbupper, const c_float* dblower, const int* dsense,
                           const DAQPSettings* settings,
                           c_float* dx, c_float* dlam, c_float* dfval, int* dexitflag, int* diter,
                           const DAQPB200Diag* diag, void* stream);

/* fp32 arithmetic end to end (c_float = float): the batched form of the reference built with -DDAQP_SINGLE_PRECISION
 * (include/types.h:8-12). Same array shapes in float; settings stay the double struct and are rounded to float like
 * the reference's constants. Plain inequality / equality / warm-start path only (no soft constraints), n <= 158;
 * diag->soft_slack must be NULL. */
int daqp_b200_solve_packed_f32(DAQPB200Handle* h, int N, int n, int m, int ms,
                               const float* H, const float* f, const float* A,
                               const float* bupper, const float* blower, const int* sense,
                               const DAQPSettings* settings,
                               float* x, float* lam, float* fval, int* exitflag, int* iter,
                               const DAQPB200Diag* diag);
int daqp_b200_solve_device_f32(DAQPB200Handle* h, int N, int n, int m, int ms,
                               const float* dH, const float* df, const float* dA,
                               const float* dbupper, const float* dblower, const int* dsense,
                               const DAQPSettings* settings,
                               float* dx, float* dlam, float* dfval, int* dexitflag, int* diter,
                               const DAQPB200Diag* diag, void* stream);

/* ---- persistent batch workspace (new): setup once, update(f, b) + solve many ------------------------------
 * The batched counterpart of the reference's workspace flow -- setup_daqp() once, then daqp_update_ldp(DAQP_UPDATE_v +
 * DAQP_UPDATE_d) and daqp_solve() per step (include/api.h:33-37, src/api.c:88-160,214-260, src/utils.c:58-221,
 * docs/docs/c.md:44-77): the Cholesky factor, R^-1, M and the row scaling stay on the device; with warm != 0 every
 * problem continues from the LDL' factor, multipliers and working set its previous solve ended with, exactly like
 * daqp_solve() on a kept DAQPWorkspace. All arrays are HOST arrays shaped as in daqp_b200_solve_packed; H and A are
 * read by the setup only. */
typedef struct DAQPB200Workspace DAQPB200Workspace;
int daqp_b200_workspace_setup(DAQPB200Handle* h, int N, int n, int m, int ms,
                              const c_float* H, const c_float* f, const c_float* A,
                              const c_float* bupper, const c_float* blower, const int* sense,
                              const DAQPSettings* settings, DAQPB200Workspace** out);
/* Shared workspace: G matrix sets -- H[G][n][n], A[G][m-ms][n], sense[G][m] (or NULL) -- and K problems per set that
 * differ only in f and the bounds (one MPC controller evaluated for many states; the reference would run setup_daqp(H_g,
 * f_p, A_g, b_p) + daqp_solve per problem, src/api.c:88-160). The QP -> LDP transform runs ONCE per set and the K
 * problems of a set stream the same device copy of R^-1 / M out of L2. Problem p = g * K + k everywhere below:
 * daqp_b200_workspace_update must give f[G*K][n], bupper[G*K][m] and blower[G*K][m] before the first solve; update / solve
 * / free (and the _device variants) are the calls of the plain workspace with N = G * K. */
int daqp_b200_workspace_setup_shared(DAQPB200Handle* h, int G, int K, int n, int m, int ms,
                                     const c_float* H, const c_float* A, const int* sense,
                                     const DAQPSettings* settings, DAQPB200Workspace** out);
/* New linear term and / or bounds (NULL keeps the current one). Asynchronous; the next solve is ordered after it. */
int daqp_b200_workspace_update(DAQPB200Workspace* w, const c_float* f, const c_float* bupper, const c_float* blower);
/* Active-set loop + result extraction. warm = 0: start from the sense bits; warm != 0: continue from the previous solve.
 * Blocks until the results are in the host arrays. diag may be NULL (ws rows hold n + ns + 1 entries). */
int daqp_b200_workspace_solve(DAQPB200Workspace* w, int warm, c_float* x, c_float* lam, c_float* fval,
                              int* exitflag, int* iter, const DAQPB200Diag* diag);
/* The same two calls for closed loops that live on the GPU: DEVICE arrays, asynchronous on `stream` (a cudaStream_t;
 * NULL = the engine's own stream), no host copies, no synchronisation. Use ONE stream for a workspace. */
int daqp_b200_workspace_update_device(DAQPB200Workspace* w, const c_float* df, const c_float* dbupper,
                                      const c_float* dblower, void* stream);
int daqp_b200_workspace_solve_device(DAQPB200Workspace* w, int warm, c_float* dx, c_float* dlam, c_float* dfval,
                                     int* dexitflag, int* diter, void* stream);
void daqp_b200_workspace_free(DAQPB200Workspace* w);
/* Exit flags the setup (or the last update) raised per problem, 0 where none: infeasible bounds (-1), non-convex or
 * singular Hessian (-5), out-of-scope input (-8). HOST array [N]; synchronises the workspace's stream. */
int daqp_b200_workspace_flags(DAQPB200Workspace* w, int* exitflag);

/* ---- warm-start initialisers (the callers on the input side of the path) -----------------------------------------
 * reference include/api.h:57-58 (src/api.c:577-631): set the ACTIVE / LOWER bits of qp->sense from a primal iterate
 * (rows with |a_i'x - bound_i| < 1e-9) or from a dual iterate (|lam_i| > 1e-12); IMMUTABLE rows are left alone. The
 * next solve activates those rows first. Drop-in signatures; one problem, qp->sense updated in place. */
void daqp_primal_init_active(DAQPProblem* qp, c_float* x);
void daqp_dual_init_active(DAQPProblem* qp, c_float* lam);

/* NEW: the same for N problems of one shape. Exactly one of x ([N][n]) and lam ([N][m]) is given (x wins when both
 * are); A / bupper / blower are read only with x. sense ([N][m] ints) is updated in place. HOST arrays, blocking. */
int daqp_b200_init_active(DAQPB200Handle* h, int N, int n, int m, int ms, const c_float* x, const c_float* lam,
                          const c_float* A, const c_float* bupper, const c_float* blower, int* sense);
/* DEVICE arrays, asynchronous on `stream`: lets a closed loop derive the next step's warm start from the previous
 * step's x or lam without leaving the GPU (solve_device -> init_active_device -> solve_device). */
int daqp_b200_init_active_device(DAQPB200Handle* h, int N, int n, int m, int ms, const c_float* dx, const c_float* dlam,
                                 const c_float* dA, const c_float* dbupper, const c_float* dblower, int* dsense,
                                 void* stream);

/* NEW: daqp_first_violating (reference include/api.h:55) for N points x[N][n] against one polyhedron: first[p] = index of
 * the first constraint point p violates by more than tol, or m. HOST arrays, blocking. */
int daqp_b200_first_violating_batch(DAQPB200Handle* h, int N, int n, int m, int ms, const c_float* x, const c_float* A,
                                    const c_float* bupper, const c_float* blower, c_float tol, int* first);

/* ---- binary constraints: branch and bound with batched node relaxations (batched LDP consumer) -------------------------
 * reference src/bnb.c:23-128 (what daqp_quadprog runs when a constraint carries DAQP_BINARY: it must hold with equality
 * at its lower or at its upper bound). The tree search stays on the host; every wave of open nodes (at most wave_width,
 * 0 = 256) is ONE launch of the solve kernel in shared-matrix mode -- the QP -> LDP transform runs once, a node is its
 * sense bytes (fixed binaries ACTIVE + IMMUTABLE, the parent's working set as warm start), the incumbent prunes through
 * settings.fval_bound exactly as in the reference. Same optimum, exit flag and fval as the reference; res->nodes and
 * res->iter count this walk (waves of the deepest open nodes, not strict depth first); res->lam holds the incumbent's
 * multipliers. daqp_quadprog / daqp_quadprog_batch route problems with binary constraints here. */
int daqp_b200_bnb(DAQPB200Handle* h, const DAQPProblem* qp, const DAQPSettings* settings, DAQPResult* res, int wave_width);

/* ---- raw LDPs (batched LDP consumer) -----------------------------------------------------------------------------------
 * P problems  min |u|^2  s.t.  blower <= [I(ms); A] u <= bupper  solved as they stand: M = A as given (NO normalisation,
 * scaling == NULL), Rinv == NULL, v == NULL -- what the reference's daqp_ldp (include/daqp.h, src/daqp.c:6-108) runs on a
 * workspace whose LDP fields the caller filled in himself, the way Julia's polyhedral tools test feasibility
 * (interfaces/daqp-julia/src/api.jl:428-459; settings->fval_bound = their max_radius). A[P][m-ms][n], bupper / blower
 * [P][m], sense[P][m] or NULL (ACTIVE rows are activated first; SOFT / BINARY bits: -8). Outputs: u[P][n], lam[P][m]
 * (scattered like DAQPResult.lam), fval[P] = |u|^2 / 2, exitflag[P], iter[P]; any of u / lam / fval / iter may be NULL.
 * diag (optional): n_active[P], ws[P][n+1] in factor order, sense[P][ldm]. */
int daqp_b200_ldp_batch(DAQPB200Handle* h, int P, int n, int m, int ms, const c_float* A, const c_float* bupper,
                        const c_float* blower, const int* sense, const DAQPSettings* settings, c_float* u, c_float* lam,
                        c_float* fval, int* exitflag, int* iter, const DAQPB200Diag* diag);
/* the same on device arrays, asynchronous on `stream` (NULL = the engine's own) */
int daqp_b200_ldp_device(DAQPB200Handle* h, int P, int n, int m, int ms, const c_float* dA, const c_float* dbupper,
                         const c_float* dblower, const int* dsense, const DAQPSettings* settings, c_float* du, c_float* dlam,
                         c_float* dfval, int* dexitflag, int* diter, void* stream);

/* ---- minimal representation of polyhedra (batched LDP consumer) ---------------------------------------------
 * reference include/api.h:54 (src/api.c:531-556, src/utils.c:808-835): is_redundant[i] = 1 iff constraint i of
 * {x : [I(ms); A] x <= b} is redundant (the LDP with row i turned into an active equality is infeasible), else 0.
 * Drop-in signature; the m LDPs run concurrently on the GPU instead of one after the other. On a CUDA-side failure
 * every entry is set to -1. */
void daqp_minrep(int* is_redundant, c_float* A, c_float* b, int n, int m, int ms);

/* NEW: P polyhedra of one shape in one call -- A[P][m-ms][n], b[P][m] -> is_redundant[P][m]; all P*m LDPs are solved
 * concurrently (one warp each), the m LDPs of a polyhedron sharing one device copy of its matrix. exitflag / iter
 * ([P][m], may be NULL) return the exit flag and iteration count of every LDP (what daqp_ldp returns for it in the
 * reference when no earlier constraint was found redundant). HOST arrays; blocks until the results are in place.
 * An EMPTY polyhedron (every probe infeasible) is answered like the reference answers it -- constraints dropped from the
 * front until the rest is non-empty -- by re-running it with those constraints taken out. */
int daqp_b200_minrep_batch(DAQPB200Handle* h, int P, int n, int m, int ms, const c_float* A, const c_float* b,
                           const DAQPSettings* settings, int* is_redundant, int* exitflag, int* iter);
/* Same with DEVICE arrays, asynchronous on `stream` (a cudaStream_t; NULL = the engine's own stream). One round, no
 * host synchronisation: an empty polyhedron comes back with every entry 1 (the caller can see that and decide). */
int daqp_b200_minrep_device(DAQPB200Handle* h, int P, int n, int m, int ms, const c_float* dA, const c_float* db,
                            const DAQPSettings* settings, int* dis_redundant, int* dexitflag, int* diter, void* stream);

/* Device-time accounting of the engine since the last reset (CUDA events on the launching stream). */
typedef struct {
    int setup_launches; /* qp_setup_kernel launches   */
    int solve_launches; /* ldp_solve_kernel launches  */
    double setup_ms;    /* summed device time of the setup kernels */
    double solve_ms;    /* summed device time of the solve kernels */
    int warps_per_sm;   /* resident problems per SM in the last solve launch */
    long long scratch_bytes;
} DAQPB200Stats;
/* Synchronises the engine's streams, then reports and optionally clears the counters. */
int daqp_b200_get_stats(DAQPB200Handle* h, DAQPB200Stats* out, int reset);

/* Upper bound on device scratch per engine (bytes); batches that need more are processed in chunks. */
void daqp_b200_set_scratch_limit(DAQPB200Handle* h, long long bytes);

const char* daqp_b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* DAQP_B200_H */
