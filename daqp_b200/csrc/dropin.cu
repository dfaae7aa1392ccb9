// daqp_b200/csrc/dropin.cu -- the reference's WORKSPACE entry points with their own signatures (include/api.h:33-49,
// include/utils.h:11), as batch-of-one wrappers over the batched engine, so that the reference's interfaces bind to
// this library unchanged: Cython daqp.Model (interfaces/daqp-python/daqp.pxd:50-88, daqp.pyx:8-57,264-572), the Eigen
// class DAQP (interfaces/daqp-eigen/daqp.cpp:141-322), Julia's Model (interfaces/daqp-julia/src/api.jl:202-292).
//
//   setup_daqp / setup_daqp_main   src/api.c:84-160      allocate_daqp_settings   src/api.c:277-282
//   daqp_solve                     src/api.c:8-59        free_daqp_workspace      src/api.c:399-426
//   daqp_update_ldp                src/utils.c:58-221    free_daqp_ldp            src/api.c:244-275
//   daqp_set_primal_start          src/api.c:636-641     daqp_first_violating     src/api.c:562-574
//
// The caller owns the DAQPWorkspace struct (the interfaces calloc it); this file fills the fields interfaces read back
// (n, m, ms, qp, settings, x, lam, lam_star, WS, n_active, iterations, fval, soft_slack, sense, scaling-free) from the
// GPU results. There is no CPU solve path: everything numerical happens in the kernels behind daqp_b200_*.
#include "../../include/daqp_b200.h"

#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

// host-side bookkeeping of one workspace (DAQPWorkspace::avi slot)
struct DropinHost {
    int raw = 0;             // allocate_daqp_workspace: the caller fills M / dupper / dlower / sense himself (daqp_ldp)
    int cold_next = 0;       // reset_daqp_workspace was called: the next solve starts from an empty working set
    int one_shot = 0;        // init_mask carried DAQP_UPDATE_unconstrained: setup + solve are one daqp_quadprog call
    int ns = 0;              // soft constraints of the problem (sizes WS / lam like src/api.c:296-313)
    int have_result = 0;     // one-shot: the result computed at setup time is waiting for daqp_solve
    c_float fval = 0, soft_slack = 0, solve_time = 0;
    int exitflag = 0, iter = 0;
    std::vector<c_float> x, lam;   // results of the last solve, [n] and [m]
    std::vector<int> ws;           // working set in factor order
    std::vector<unsigned char> sense8;
    std::vector<c_float> eye;      // H == NULL, f == NULL (the LDP min |x|^2 itself): identity Hessian handed to the engine
};

DropinHost* host_of(DAQPWorkspace* w) { return static_cast<DropinHost*>(w->avi); }
DAQPB200Workspace* dev_of(DAQPWorkspace* w) { return static_cast<DAQPB200Workspace*>(w->eq); }

// the Hessian the engine gets: the caller's, or the identity for a pure LDP (same iterates as the reference's Rinv == NULL path)
const c_float* hessian_of(DropinHost* hs, const DAQPProblem* qp) {
    if (qp->H) return qp->H;
    const size_t n = (size_t)qp->n;
    if (hs->eye.size() != n * n) {
        hs->eye.assign(n * n, 0);
        for (size_t d = 0; d < n; d++) hs->eye[d * n + d] = 1;
    }
    return hs->eye.data();
}

bool in_scope(const DAQPProblem* qp) {
    if (!qp || (qp->H == nullptr && qp->f != nullptr) || qp->nh > 1 || qp->problem_type != 0 || qp->n < 1 || qp->m < qp->ms || qp->ms > qp->n) return false;
    if (qp->m > qp->ms && qp->A == nullptr) return false;
    if (qp->m > 0 && (qp->bupper == nullptr || qp->blower == nullptr)) return false;
    if (qp->sense)
        for (int i = 0; i < qp->m; i++) if (qp->sense[i] & DAQP_BINARY) return false;
    return true;
}

// host copies of the iterates the interfaces read (src/api.c:296-340 allocates the same arrays, n + ns + 1 long)
void alloc_iterates(DAQPWorkspace* w, int n, int m, int ns) {
    const int cap = n + ns + 1;
    w->x = static_cast<c_float*>(calloc((size_t)(n > 0 ? n : 1), sizeof(c_float)));
    w->xold = nullptr;
    w->u = w->x; // the reference aliases them too (src/api.c:316)
    w->lam = static_cast<c_float*>(calloc((size_t)cap, sizeof(c_float)));
    w->lam_star = static_cast<c_float*>(calloc((size_t)cap, sizeof(c_float)));
    w->WS = static_cast<int*>(calloc((size_t)cap, sizeof(int)));
    w->sense = static_cast<int*>(calloc((size_t)(m > 0 ? m : 1), sizeof(int)));
    w->n_active = 0; w->iterations = 0; w->sing_ind = -1; w->reuse_ind = 0; w->fval = 0; w->soft_slack = 0; w->n_prox = 0;
}

// one batched solve of one problem through the array-of-struct entry (what daqp_quadprog runs), keeping the diagnostics
int run_one_shot(DAQPWorkspace* w, DropinHost* hs) {
    DAQPProblem* qp = w->qp;
    const int n = qp->n, m = qp->m, cap = n + hs->ns + 1, ldm = (m + 3) / 4 * 4;
    hs->x.assign((size_t)n, 0); hs->lam.assign((size_t)(m > 0 ? m : 1), 0); hs->ws.assign((size_t)cap, 0);
    hs->sense8.assign((size_t)(ldm > 0 ? ldm : 4), 0);
    int nact = 0;
    DAQPB200Diag dg{};
    dg.n_active = &nact; dg.ws = hs->ws.data(); dg.sense = hs->sense8.data(); dg.soft_slack = &hs->soft_slack;
    DAQPB200Stats before{}, after{};
    daqp_b200_get_stats(nullptr, &before, 0);
    hs->fval = 0;
    int rc = daqp_b200_solve_packed(nullptr, 1, n, m, qp->ms, hessian_of(hs, qp), qp->f, qp->A, qp->bupper, qp->blower, qp->sense,
                                    w->settings, hs->x.data(), hs->lam.data(), &hs->fval, &hs->exitflag, &hs->iter, &dg);
    if (rc) return DAQP_EXIT_UNSUPPORTED;
    daqp_b200_get_stats(nullptr, &after, 0);
    hs->solve_time = 1e-3 * (after.solve_ms - before.solve_ms);
    w->n_active = nact;
    hs->have_result = 1;
    return hs->exitflag;
}

void publish(DAQPResult* res, DAQPWorkspace* w, DropinHost* hs, int nact) {
    const int n = w->n, m = w->m;
    const bool solved = hs->iter > 0; // setup failures leave x untouched (src/api.c:69-72)
    res->exitflag = hs->exitflag;
    res->iter = hs->iter; res->nodes = 1; res->solve_time = hs->solve_time;
    w->iterations = hs->iter; w->n_active = nact;
    if (!solved) return;
    for (int i = 0; i < n; i++) { res->x[i] = hs->x[i]; w->x[i] = hs->x[i]; }
    if (res->lam) for (int i = 0; i < m; i++) res->lam[i] = hs->lam[i];
    for (int i = 0; i < nact; i++) { w->WS[i] = hs->ws[i]; w->lam_star[i] = hs->lam[hs->ws[i]]; w->lam[i] = w->lam_star[i]; }
    for (int i = 0; i < m; i++) w->sense[i] = hs->sense8[i];
    if (w->qp && w->qp->f) res->fval = hs->fval;
    res->soft_slack = hs->soft_slack;
    w->fval = hs->fval; w->soft_slack = hs->soft_slack;
}

} // namespace

extern "C" void allocate_daqp_settings(DAQPWorkspace* work) { // src/api.c:277-282
    if (work->settings == nullptr) {
        work->settings = static_cast<DAQPSettings*>(malloc(sizeof(DAQPSettings)));
        daqp_default_settings(work->settings);
    }
}

extern "C" void free_daqp_ldp(DAQPWorkspace* work) { // src/api.c:244-275: the LDP of a workspace lives on the device
    if (work->eq) { daqp_b200_workspace_free(dev_of(work)); work->eq = nullptr; }
    if (work->sense) { free(work->sense); work->sense = nullptr; }
}

extern "C" void free_daqp_workspace(DAQPWorkspace* work) { // src/api.c:399-426
    if (work->lam != nullptr) {
        free(work->lam); free(work->lam_star); free(work->WS); free(work->x);
        work->lam = nullptr; work->lam_star = nullptr; work->WS = nullptr; work->x = nullptr; work->u = nullptr;
    }
    if (work->settings != nullptr) { free(work->settings); work->settings = nullptr; }
    if (work->avi) { delete host_of(work); work->avi = nullptr; }
}

extern "C" int setup_daqp_main(DAQPProblem* qp, DAQPWorkspace* work, c_float* setup_time, int init_mask) {
    if (setup_time) *setup_time = 0;
    const bool own_settings = work->settings == nullptr;
    auto fail_with = [&](int flag) { // src/api.c:137-149: a settings struct the caller installed is not freed here
        if (!own_settings) work->settings = nullptr;
        free_daqp_ldp(work);
        free_daqp_workspace(work);
        return flag;
    };
    if (own_settings) allocate_daqp_settings(work);
    if (!in_scope(qp)) return fail_with(DAQP_EXIT_UNSUPPORTED);
    int ns = 0;
    if (qp->sense)
        for (int i = 0; i < qp->m; i++) ns += (qp->sense[i] & DAQP_SOFT) ? 1 : 0;
    if (work->eq || work->avi || work->lam) { // set up again on a live workspace: start clean, keep the settings
        DAQPSettings* keep = work->settings;
        work->settings = nullptr;
        free_daqp_ldp(work); free_daqp_workspace(work);
        work->settings = keep;
    }
    work->qp = qp; work->n = qp->n; work->m = qp->m; work->ms = qp->ms; work->nh = 1; work->break_points = nullptr;
    alloc_iterates(work, qp->n, qp->m, ns);
    DropinHost* hs = new DropinHost();
    hs->ns = ns;
    hs->one_shot = (init_mask & DAQP_UPDATE_unconstrained) ? 1 : 0;
    work->avi = hs;
    DAQPB200Stats before{}, after{};
    daqp_b200_get_stats(nullptr, &before, 0);
    if (hs->one_shot) {
        // setup_daqp_main(mask of daqp_quadprog) + daqp_solve IS daqp_quadprog (src/api.c:62-79): run it as one call now and
        // hand the result out in daqp_solve; failures of the transform surface here, like in the reference
        const int flag = run_one_shot(work, hs);
        if (flag < 0 && hs->iter == 0) return fail_with(flag);
    } else {
        DAQPB200Workspace* dw = nullptr;
        if (daqp_b200_workspace_setup(nullptr, 1, qp->n, qp->m, qp->ms, hessian_of(hs, qp), qp->f, qp->A, qp->bupper, qp->blower, qp->sense,
                                      work->settings, &dw) != 0)
            return fail_with(DAQP_EXIT_UNSUPPORTED);
        work->eq = dw;
        int flag = 0;
        if (daqp_b200_workspace_flags(dw, &flag) != 0) return fail_with(DAQP_EXIT_UNSUPPORTED);
        if (flag < 0) return fail_with(flag);
    }
    daqp_b200_get_stats(nullptr, &after, 0);
    if (setup_time) *setup_time = 1e-3 * (after.setup_ms - before.setup_ms);
    return 1;
}

extern "C" int setup_daqp(DAQPProblem* qp, DAQPWorkspace* work, c_float* setup_time) { // src/api.c:84-86
    return setup_daqp_main(qp, work, setup_time, 0);
}

extern "C" void daqp_solve(DAQPResult* res, DAQPWorkspace* work) { // src/api.c:8-59
    DropinHost* hs = host_of(work);
    if (!hs) { res->exitflag = DAQP_EXIT_UNSUPPORTED; return; }
    if (hs->one_shot) {
        if (!hs->have_result && run_one_shot(work, hs) == DAQP_EXIT_UNSUPPORTED && hs->iter == 0) { res->exitflag = DAQP_EXIT_UNSUPPORTED; return; }
        hs->have_result = 0; // a second daqp_solve on the same workspace solves again
        publish(res, work, hs, work->n_active);
        return;
    }
    DAQPB200Workspace* dw = dev_of(work);
    if (!dw) { res->exitflag = DAQP_EXIT_UNSUPPORTED; return; }
    const int n = work->n, m = work->m, cap = n + hs->ns + 1, ldm = (m + 3) / 4 * 4;
    hs->x.assign((size_t)n, 0); hs->lam.assign((size_t)(m > 0 ? m : 1), 0); hs->ws.assign((size_t)cap, 0);
    hs->sense8.assign((size_t)(ldm > 0 ? ldm : 4), 0);
    int nact = 0;
    DAQPB200Diag dg{};
    dg.n_active = &nact; dg.ws = hs->ws.data(); dg.sense = hs->sense8.data(); dg.soft_slack = &hs->soft_slack;
    DAQPB200Stats before{}, after{};
    daqp_b200_get_stats(nullptr, &before, 0);
    hs->fval = 0;
    // warm: continue from the factor and working set the previous solve left (what daqp_solve does on a kept workspace)
    const int warm = hs->cold_next ? 0 : 1;
    hs->cold_next = 0;
    if (daqp_b200_workspace_solve(dw, warm, hs->x.data(), hs->lam.data(), &hs->fval, &hs->exitflag, &hs->iter, &dg) != 0) {
        res->exitflag = DAQP_EXIT_UNSUPPORTED;
        return;
    }
    daqp_b200_get_stats(nullptr, &after, 0);
    hs->solve_time = 1e-3 * (after.solve_ms - before.solve_ms);
    publish(res, work, hs, nact);
}

extern "C" int daqp_update_ldp(const int mask, DAQPWorkspace* work, DAQPProblem* qp) { // src/utils.c:58-221
    DropinHost* hs = host_of(work);
    if (!hs || !in_scope(qp)) return DAQP_EXIT_UNSUPPORTED;
    work->qp = qp;
    if (hs->one_shot) { hs->have_result = 0; return 0; } // the next daqp_solve runs the whole call on the new data
    DAQPB200Workspace* dw = dev_of(work);
    const bool same_shape = qp->n == work->n && qp->m == work->m && qp->ms == work->ms;
    if (!dw || !same_shape || (mask & (DAQP_UPDATE_Rinv | DAQP_UPDATE_M | DAQP_UPDATE_sense))) {
        // new matrices or sense bits: redo the transform (the reference resets the working set as well, utils.c:199-211,470)
        if (dw) { daqp_b200_workspace_free(dw); work->eq = nullptr; }
        int ns = 0;
        if (qp->sense)
            for (int i = 0; i < qp->m; i++) ns += (qp->sense[i] & DAQP_SOFT) ? 1 : 0;
        if (!same_shape || ns != hs->ns) { // the host iterates follow the problem's size
            free(work->lam); free(work->lam_star); free(work->WS); free(work->x); free(work->sense);
            work->n = qp->n; work->m = qp->m; work->ms = qp->ms;
            alloc_iterates(work, qp->n, qp->m, ns);
            hs->ns = ns;
        }
        if (daqp_b200_workspace_setup(nullptr, 1, qp->n, qp->m, qp->ms, hessian_of(hs, qp), qp->f, qp->A, qp->bupper, qp->blower, qp->sense,
                                      work->settings, &dw) != 0)
            return DAQP_EXIT_UNSUPPORTED;
        work->eq = dw;
    } else if (mask & (DAQP_UPDATE_v | DAQP_UPDATE_d)) {
        if (daqp_b200_workspace_update(dw, (mask & DAQP_UPDATE_v) ? qp->f : nullptr, (mask & DAQP_UPDATE_d) ? qp->bupper : nullptr,
                                       (mask & DAQP_UPDATE_d) ? qp->blower : nullptr) != 0)
            return DAQP_EXIT_UNSUPPORTED;
    }
    int flag = 0;
    if (daqp_b200_workspace_flags(dw, &flag) != 0) return DAQP_EXIT_UNSUPPORTED;
    return flag < 0 ? flag : 0;
}

// ---- hand-filled workspaces and the pieces of daqp_solve -------------------------------------------------------------
extern "C" void reset_daqp_workspace(DAQPWorkspace* work) { // src/daqp.c:142-146
    work->sing_ind = -1; work->n_active = 0; work->reuse_ind = 0;
    if (DropinHost* hs = host_of(work)) hs->cold_next = 1; // the device keeps the factor: tell the next solve not to use it
}

extern "C" void daqp_deactivate_constraints(DAQPWorkspace* work) { // src/auxiliary.c:482-488
    if (!work->sense || !work->WS) return;
    for (int i = 0; i < work->n_active; i++) {
        const int id = work->WS[i];
        if (id >= 0 && id < work->m && !(work->sense[id] & DAQP_IMMUTABLE)) work->sense[id] &= ~DAQP_ACTIVE;
    }
}

extern "C" void allocate_daqp_workspace(DAQPWorkspace* work, int n, int ns) { // src/api.c:295-340
    const int cap = n + ns + 1;
    work->n = n;
    work->Rinv = nullptr; work->RinvD = nullptr; work->v = nullptr; work->scaling = nullptr; work->Mu = nullptr;
    work->lam = static_cast<c_float*>(calloc((size_t)cap, sizeof(c_float)));
    work->lam_star = static_cast<c_float*>(calloc((size_t)cap, sizeof(c_float)));
    work->WS = static_cast<int*>(calloc((size_t)cap, sizeof(int)));
    work->x = static_cast<c_float*>(calloc((size_t)(n > 0 ? n : 1), sizeof(c_float)));
    work->u = work->x; work->xold = nullptr;
    work->L = nullptr; work->D = nullptr; work->xldl = nullptr; work->zldl = nullptr; // the factor lives on the device
    work->prox_mask = nullptr; work->n_prox = 0;
    work->bnb = nullptr; work->nh = 1; work->break_points = nullptr; work->eq = nullptr; work->timer = nullptr;
    work->fval = 0; work->soft_slack = 0; work->iterations = 0;
    DropinHost* hs = new DropinHost();
    hs->raw = 1; hs->ns = ns;
    work->avi = hs;
    reset_daqp_workspace(work);
    hs->cold_next = 0;
}

extern "C" int daqp_ldp(DAQPWorkspace* work) { // src/daqp.c:6-108
    DropinHost* hs = host_of(work);
    if (!hs) return DAQP_EXIT_UNSUPPORTED;
    if (!hs->raw) { // a workspace from setup_daqp: the solve inside daqp_solve (work->x then holds the QP-space solution)
        std::vector<c_float> x((size_t)(work->n > 0 ? work->n : 1)), lam((size_t)(work->m > 0 ? work->m : 1));
        DAQPResult tmp{};
        tmp.x = x.data(); tmp.lam = lam.data();
        daqp_solve(&tmp, work);
        return tmp.exitflag;
    }
    const int n = work->n, m = work->m, ms = work->ms, cap = n + 1, ldm = (m + 3) / 4 * 4;
    if (n < 1 || m < 1 || ms < 0 || ms > n || m < ms || work->Rinv || work->RinvD || hs->ns > 0 || (m > ms && !work->M) || !work->dupper ||
        !work->dlower)
        return DAQP_EXIT_UNSUPPORTED;
    hs->lam.assign((size_t)m, 0); hs->ws.assign((size_t)cap, 0); hs->sense8.assign((size_t)ldm, 0);
    int nact = 0, flag = DAQP_EXIT_UNSUPPORTED, iter = 0;
    c_float fv = 0;
    DAQPB200Diag dg{};
    dg.n_active = &nact; dg.ws = hs->ws.data(); dg.sense = hs->sense8.data();
    if (daqp_b200_ldp_batch(nullptr, 1, n, m, ms, work->M, work->dupper, work->dlower, work->sense, work->settings, work->u,
                            hs->lam.data(), &fv, &flag, &iter, &dg) != 0)
        return DAQP_EXIT_UNSUPPORTED;
    work->iterations = iter; work->n_active = nact; work->sing_ind = -1; work->reuse_ind = nact;
    for (int i = 0; i < nact && i < cap; i++) { work->WS[i] = hs->ws[i]; work->lam_star[i] = hs->lam[hs->ws[i]]; work->lam[i] = work->lam_star[i]; }
    if (work->sense) for (int i = 0; i < m; i++) work->sense[i] = hs->sense8[i];
    if (flag > 0) work->fval = 2 * fv; // |u|^2 (daqp.c keeps the squared norm, auxiliary.c:85-87)
    hs->exitflag = flag; hs->iter = iter; hs->fval = 0; hs->soft_slack = 0;
    return flag;
}

extern "C" void daqp_extract_result(DAQPResult* res, DAQPWorkspace* work) { // src/api.c:455-495
    DropinHost* hs = host_of(work);
    for (int i = 0; i < work->n; i++) res->x[i] = work->x[i];
    if (res->lam) {
        for (int i = 0; i < work->m; i++) res->lam[i] = 0;
        for (int i = 0; i < work->n_active; i++) res->lam[work->WS[i]] = work->lam_star[i];
    }
    if (hs && !hs->raw && work->qp && work->qp->f) res->fval = hs->fval; // (no linear term: fval is left as it is)
    res->soft_slack = work->soft_slack;
    res->iter = work->iterations;
    res->nodes = 1;
}

extern "C" void daqp_set_primal_start(DAQPWorkspace* work, c_float* x) { // src/api.c:636-641
    if (work->x && x) for (int i = 0; i < work->n; i++) work->x[i] = x[i];
}

// src/api.c:562-574: a batch of one through the batched kernel (daqp_b200_first_violating_batch, warmstart_kernel.cuh).
// Like the reference it has no error channel: on a CUDA-side failure it returns m ("nothing violated") and
// daqp_b200_last_error() says why.
extern "C" int daqp_first_violating(c_float* x, c_float* A, c_float* bu, c_float* bl, int n, int m, int ms, c_float tol) {
    int first = m;
    if (daqp_b200_first_violating_batch(nullptr, 1, n, m, ms, x, A, bu, bl, tol, &first) != 0) return m;
    return first;
}
