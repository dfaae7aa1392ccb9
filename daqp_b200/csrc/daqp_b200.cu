// daqp_b200/csrc/daqp_b200.cu -- host side of the C ABI declared in include/daqp_b200.h.
//
// Mirrors the reference's entry sequence for the hot path (src/api.c:62-79: setup_daqp_main -> daqp_solve ->
// daqp_extract_result) as two kernel launches per chunk of problems:
//   qp_setup_kernel  (QP -> LDP, setup_kernel.cuh)   then   ldp_solve_kernel (active-set loop + extraction).
// There is no CPU solve path in this library: without a CUDA device every entry point reports an error.
#include "../../include/daqp_b200.h"
#include "ldp_kernel.cuh"
#include "setup_kernel.cuh"
#include "update_kernel.cuh"
#include "minrep_kernel.cuh"
#include "warmstart_kernel.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

using namespace dq;

static thread_local std::string g_last_error;
static int fail(const char* what, cudaError_t e, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "daqp_b200: %s failed at line %d: %s", what, line, cudaGetErrorString(e));
    g_last_error = buf;
    return -100 - (int)e;
}
#define CK(call)                                                       \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) return fail(#call, e_, __LINE__);       \
    } while (0)

extern "C" const char* daqp_b200_last_error(void) { return g_last_error.c_str(); }

struct EventTriple { cudaEvent_t e0, e1, e2; };

struct DAQPB200Handle {
    int device = 0, num_sms = 0;
    size_t smem_optin = 0;
    cudaStream_t compute = nullptr, copy_in = nullptr, copy_out = nullptr;
    char* arena = nullptr;
    size_t arena_bytes = 0;
    char* stage = nullptr;
    size_t stage_bytes = 0;
    long long scratch_limit = 24ll << 30;
    std::vector<EventTriple> pending, free_events;
    DAQPB200Stats stats{};
    std::mutex mu;
    // The scratch arena is reused from offset 0 by every call. Calls may arrive on different streams: the last user records
    // `arena_ev`, the next one's stream waits on it before it touches the arena.
    cudaEvent_t arena_ev = nullptr;
    cudaStream_t arena_stream = nullptr;
    bool arena_used = false;
    char* pin = nullptr;          // pinned host mirror of `stage` for small batches (one copy in, one copy out)
    size_t pin_bytes = 0;
    int live_workspaces = 0;      // persistent workspaces that still point at this handle
};

// order this call after the previous user of the arena (no-op on the same stream)
static int arena_acquire(DAQPB200Handle* h, cudaStream_t s) {
    if (h->arena_used && h->arena_stream != s) {
        cudaError_t e = cudaStreamWaitEvent(s, h->arena_ev, 0);
        if (e != cudaSuccess) { g_last_error = "daqp_b200: cudaStreamWaitEvent on the arena event failed"; return -100 - (int)e; }
    }
    return 0;
}
static int arena_release(DAQPB200Handle* h, cudaStream_t s) {
    cudaError_t e = cudaEventRecord(h->arena_ev, s);
    if (e != cudaSuccess) { g_last_error = "daqp_b200: cudaEventRecord on the arena event failed"; return -100 - (int)e; }
    h->arena_used = true; h->arena_stream = s;
    return 0;
}

extern "C" void daqp_default_settings(DAQPSettings* s) { // reference src/api.c:505-527, include/constants.h:15-29
    s->primal_tol = 1e-6; s->dual_tol = 1e-12; s->zero_tol = 1e-11; s->pivot_tol = 1e-6;
    s->progress_tol = 1e-14; s->cycle_tol = 10; s->iter_limit = 10000; s->fval_bound = DAQP_INF;
    s->eps_prox = -1e-6; s->eta_prox = -1.0; s->rho_soft = 1e-6; s->rel_subopt = 0; s->abs_subopt = 0;
    s->sing_tol = 3.7e-11; s->refactor_tol = 1e-9; s->time_limit = 0;
}

template <typename T>
static DevSettings<T> to_dev_settings(const DAQPSettings* s) {
    DAQPSettings d;
    if (!s) { daqp_default_settings(&d); s = &d; }
    DevSettings<T> o;
    o.primal_tol = (T)s->primal_tol; o.dual_tol = (T)s->dual_tol; o.zero_tol = (T)s->zero_tol;
    o.pivot_tol = (T)s->pivot_tol; o.progress_tol = (T)s->progress_tol; o.fval_bound = (T)s->fval_bound;
    o.rho_soft = (T)s->rho_soft; o.sing_tol = (T)s->sing_tol; o.refactor_tol = (T)s->refactor_tol;
    o.eps_prox = (T)s->eps_prox; o.cycle_tol = s->cycle_tol; o.iter_limit = s->iter_limit;
    o.time_limit_ns = s->time_limit > 0 ? (long long)(s->time_limit * 1e9) + 1 : 0;
    return o;
}

extern "C" int daqp_b200_create(DAQPB200Handle** out, int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_last_error = "daqp_b200: no CUDA device available (this library has no CPU path)";
        return -1;
    }
    if (device < 0) CK(cudaGetDevice(&device));
    CK(cudaSetDevice(device));
    DAQPB200Handle* h = new DAQPB200Handle();
    h->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    CK(cudaStreamCreateWithFlags(&h->compute, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->arena_ev, cudaEventDisableTiming));
    const char* lim = getenv("DAQP_B200_SCRATCH_GB");
    if (lim) h->scratch_limit = (long long)(atof(lim) * (double)(1ll << 30));
    *out = h;
    return 0;
}

extern "C" void daqp_b200_destroy(DAQPB200Handle* h) {
    if (!h) return;
    {
        std::lock_guard<std::mutex> lk(h->mu);
        if (h->live_workspaces > 0) { // workspaces dereference the handle: free them first
            g_last_error = "daqp_b200: destroy refused, persistent workspaces of this engine are still alive";
            return;
        }
    }
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->arena_ev) cudaEventDestroy(h->arena_ev);
    if (h->pin) cudaFreeHost(h->pin);
    for (auto& t : h->pending) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); cudaEventDestroy(t.e2); }
    for (auto& t : h->free_events) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); cudaEventDestroy(t.e2); }
    if (h->arena) cudaFree(h->arena);
    if (h->stage) cudaFree(h->stage);
    cudaStreamDestroy(h->compute); cudaStreamDestroy(h->copy_in); cudaStreamDestroy(h->copy_out);
    delete h;
}

extern "C" void daqp_b200_set_scratch_limit(DAQPB200Handle* h, long long bytes) { if (h) h->scratch_limit = bytes; }

static std::mutex g_default_mu;
static std::map<int, DAQPB200Handle*> g_default;
static int default_handle(DAQPB200Handle** out) {
    std::lock_guard<std::mutex> lk(g_default_mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        g_last_error = "daqp_b200: no CUDA device available (this library has no CPU path)";
        return -1;
    }
    auto it = g_default.find(dev);
    if (it == g_default.end()) {
        DAQPB200Handle* h = nullptr;
        int rc = daqp_b200_create(&h, dev);
        if (rc) return rc;
        g_default[dev] = h;
        *out = h;
    } else *out = it->second;
    return 0;
}

static int drain_events(DAQPB200Handle* h) {
    for (auto& t : h->pending) {
        CK(cudaEventSynchronize(t.e2));
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, t.e0, t.e1));
        CK(cudaEventElapsedTime(&b, t.e1, t.e2));
        h->stats.setup_ms += a;
        h->stats.solve_ms += b;
        h->free_events.push_back(t);
    }
    h->pending.clear();
    return 0;
}

extern "C" int daqp_b200_get_stats(DAQPB200Handle* h, DAQPB200Stats* out, int reset) {
    if (!h) { int rc = default_handle(&h); if (rc) return rc; }
    std::lock_guard<std::mutex> lk(h->mu);
    CK(cudaSetDevice(h->device));
    int rc = drain_events(h);
    if (rc) return rc;
    h->stats.scratch_bytes = (long long)h->arena_bytes;
    if (out) *out = h->stats;
    if (reset) { int w = h->stats.warps_per_sm; h->stats = DAQPB200Stats{}; h->stats.warps_per_sm = w; }
    return 0;
}

static int get_events(DAQPB200Handle* h, EventTriple* t) {
    if (!h->free_events.empty()) { *t = h->free_events.back(); h->free_events.pop_back(); return 0; }
    if (h->pending.size() > 4096) { int rc = drain_events(h); if (rc) return rc; return get_events(h, t); }
    CK(cudaEventCreate(&t->e0)); CK(cudaEventCreate(&t->e1)); CK(cudaEventCreate(&t->e2));
    return 0;
}

static int ensure(char** buf, size_t* have, size_t need) {
    if (*have >= need) return 0;
    if (*buf) { CK(cudaDeviceSynchronize()); CK(cudaFree(*buf)); *buf = nullptr; *have = 0; }
    CK(cudaMalloc((void**)buf, need));
    *have = need;
    return 0;
}

struct Carver {
    char* p; size_t off = 0;
    explicit Carver(char* base) : p(base) {}
    template <typename U> U* take(size_t count) {
        off = (off + 255) / 256 * 256;
        U* r = reinterpret_cast<U*>(p + off);
        off += count * sizeof(U);
        return r;
    }
};

template <typename T>
static size_t scratch_per_problem(int n, int m, int ldm, int ldn) {
    size_t e = (size_t)n * ldm + (size_t)m * ldn + 3 * (size_t)ldm + (size_t)n * (n + 1) / 2 + n;
    return e * sizeof(T) + (size_t)((n + 3) / 4) * m * 16 + ldm + sizeof(int) + 64 // + fp32 quad copy + sense + flag + slack
           + (size_t)(n + 1) * sizeof(T) + 4 * sizeof(int);                           // + hand-over of the split setup
}

template <typename T, int NV, bool EXT>
static cudaError_t launch_solve_x(const LdpArgs<T>& a, int grid, int block, size_t smem, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(ldp_solve_kernel<T, NV, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin_smem());
    if (e != cudaSuccess) return e;
    ldp_solve_kernel<T, NV, EXT><<<grid, block, smem, s>>>(a);
    return cudaGetLastError();
}
// soft constraints or a persistent workspace select the extended instantiation
template <typename T, int NV>
static cudaError_t launch_solve(const LdpArgs<T>& a, int grid, int block, size_t smem, cudaStream_t s) {
    if constexpr (sizeof(T) == 4) { // fp32: plain path only, n <= 158 (five register segments)
        if (a.ns_max > 0 || a.state || a.grp > 1 || NV > 5) return cudaErrorNotSupported;
        if constexpr (NV <= 5) return launch_solve_x<T, NV, false>(a, grid, block, smem, s);
        else return cudaErrorNotSupported;
    } else {
        return (a.ns_max > 0 || a.state || a.grp > 1 || a.aux) ? launch_solve_x<T, NV, true>(a, grid, block, smem, s)
                                         : launch_solve_x<T, NV, false>(a, grid, block, smem, s);
    }
}
// team mode (n > 64, plain fp64 path): instantiated in team_launch.cu (separate translation unit: compiled in parallel)
cudaError_t daqp_b200_launch_solve_team(const LdpArgs<double>& a, int nv, int grid, size_t smem, cudaStream_t s);
cudaError_t daqp_b200_launch_solve_regs(const LdpArgs<double>& a, int nv, int grid, int block, size_t smem, cudaStream_t s);

// split QP -> LDP transform (fp64, n <= 127): factor kernel + tensor-core product kernel, instantiated in setup2_launch.cu
cudaError_t daqp_b200_launch_setup_split(const SetupArgs<double>& a, int num_sms, size_t smem_optin, cudaStream_t s);

template <typename T, int NGS>
static cudaError_t launch_setup(const SetupArgs<T>& a, int grid, int block, size_t smem, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(qp_setup_kernel<T, NGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin_smem());
    if (e != cudaSuccess) return e;
    qp_setup_kernel<T, NGS><<<grid, block, smem, s>>>(a);
    return cudaGetLastError();
}

// Largest number of soft constraints (sense & 8) any problem of the batch carries: it sizes the factor storage
// (the reference allocates n + ns + 1 rows, src/api.c:305-313).
// One warp per problem, lanes across its m entries: coalesced (a thread per problem read the batch with a stride of m ints
// and took 23 ms per 100 k C3 problems -- a fifth of a warm-started step; this form takes 0.03 ms).
__global__ void max_soft_kernel(const int* sense, int N, int m, int* out) {
    const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
    int best = 0;
    for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < N; p += warps) {
        const int* sp = sense + (size_t)p * m;
        int c = 0;
        for (int i = lane; i < m; i += 32) c += (sp[i] & B_SOFT) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        best = max(best, c);
    }
    if (lane == 0 && best > 0) atomicMax(out, best);
}

// Device arrays of a persistent workspace (daqp_b200_workspace_*): the LDP of a batch kept across solves.
template <typename T>
struct Persist {
    T *Mt = nullptr, *Mr = nullptr, *du = nullptr, *dl = nullptr, *sc = nullptr, *Ri = nullptr, *vv = nullptr;
    float* Mt32 = nullptr;
    unsigned char *sense8 = nullptr, *sense_static = nullptr;
    int* sflag = nullptr;
    char* state = nullptr;
    unsigned state_stride = 0;
    int phase = 0;      // 1 = QP -> LDP only (no shortcut), 2 = solve only
    int state_load = 0; // phase 2: continue from the saved factor / working set
    int grp = 0;        // shared workspace: problems per matrix set (the matrix arrays hold N / grp sets); 0 = own matrices
};

// Device-resident batch: the core of every entry point.
template <typename T>
static int solve_device_impl(DAQPB200Handle* h, int N, int n, int m, int ms, const T* dH, const T* df, const T* dA,
                             const T* dbu, const T* dbl, const int* dsense, const DAQPSettings* settings, T* dx,
                             T* dlam, T* dfval, int* dflag, int* diter, const DAQPB200Diag* diag, cudaStream_t stream,
                             int ns_max = -1, Persist<T>* ps = nullptr) {
    if (N <= 0) return 0;
    if (n < 1 || m < ms || ms < 0 || ms > n) { g_last_error = "daqp_b200: invalid problem dimensions"; return -2; }
    constexpr int V = VecOf<T>::N;
    if (!dsense) ns_max = 0;
    if (ns_max < 0) { // sense lives on the device: count there (one small kernel and a 4-byte read-back)
        int* dcount = nullptr;
        int rc0 = ensure(&h->arena, &h->arena_bytes, 1 << 20);
        if (rc0) return rc0;
        rc0 = arena_acquire(h, stream);
        if (rc0) return rc0;
        dcount = reinterpret_cast<int*>(h->arena);
        CK(cudaMemsetAsync(dcount, 0, sizeof(int), stream));
        max_soft_kernel<<<std::min(8 * h->num_sms, (N + 7) / 8), 256, 0, stream>>>(dsense, N, m, dcount);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&ns_max, dcount, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }
    const int ldm = round_up(std::max(m, 1), 4), ldn = round_up(n, V), cap = n + ns_max + 1, mA = m - ms;
    const int nv = (cap + 31) / 32, ngs = (n + 31) / 32; // register segments of a length-cap vector / of a row
    if (nv > 8 || ngs > 8) { g_last_error = "daqp_b200: n + (soft constraints) > 255 is not supported"; return -2; }

    LdpArgs<T> la;
    memset(&la, 0, sizeof(la));
    la.n = n; la.m = m; la.ms = ms; la.ldm = ldm; la.ldn = ldn; la.cap = cap;
    // n > 64: a warp per problem leaves the SM nearly empty (three problems fit at n = 120) -- a team of four warps per
    // problem instead (plain fp64 path; soft constraints / workspaces / shared matrices stay on the warp kernel)
    // Where the switch sits (measured per shape, scripts/bench_warm.py, 12 500 problems, cold solve ms team / warp): a warp per
    // problem keeps its fp32 screening ring up to m = 256 and wins there until n ~ 100 (n = 66: 70 / 31, 72: 81 / 39, 80: 101 /
    // 65, 96: 131 / 110, 112: 157 / 171); beyond m = 256, where the warp streams the copy through its ring in row blocks, the
    // team wins from n ~ 88 on (n = 80, m = 320: 125 / 101; 88, 400: 169 / 180; 96, 288: 149 / 152; 104, 312: 175 / 227; C4: 3x).
    int team = 0;
    const bool team_ok = sizeof(T) == 8 && ns_max == 0 && !ps && nv >= 3 && cap <= 128;
    if (team_ok && (m > 256 ? n >= 88 : n >= 104)) team = 4;
    if (const char* tenv = getenv("DAQP_B200_TEAM")) { // experiment knob: 0 = a warp per problem everywhere; 4 = a team wherever it
        const int tv = atoi(tenv);                        // can run; 2 = two-warp teams for 32 < n + 1 <= 64 as well (measured on C3:
        if (tv == 0) team = 0;                            // 220 ms against 105 ms for a warp per problem -- the iteration is bound by
        if (tv == 4 && team_ok) team = 4;                 // L2 round trips either way, and the fork-join adds barriers to each)
        if (tv == 2 && sizeof(T) == 8 && ns_max == 0 && !ps && nv == 2) team = 2;
    }
    // experiment knob DAQP_B200_REGSTAGE=1 (n <= 63, plain fp64 path): the register-staged warp kernel -- no staging arena in
    // shared memory, 16 problems per SM at n = 50 instead of 12. Measured on C3: 104.9 ms against 104.65 ms for the
    // cp.async-staged kernel with 12 -- a third more resident warps buys nothing, so it stays opt-in.
    // the decision log and settings.time_limit live in the extended and the team instantiations of the solve kernel
    const DevSettings<T> st = to_dev_settings<T>(settings);
    // (... and so does the screening scan in row blocks for m > 256: see scan_infeasible)
    const bool aux = sizeof(T) == 8 && ((diag && diag->trace && diag->trace_cap > 0) || st.time_limit_ns != 0 || (m > 256 && m <= 768));
    la.aux = aux ? 1 : 0;
    bool regstage = false;
    if (const char* renv = getenv("DAQP_B200_REGSTAGE")) regstage = atoi(renv) != 0 && sizeof(T) == 8 && ns_max == 0 && !ps && team == 0 && nv <= 2 && !aux;
    const size_t smem_solve_w = ldp_layout<T>(la, regstage ? -1 : team), smem_setup_w = setup_smem_per_warp<T>(n);
    const size_t budget = h->smem_optin;
    int w_solve = (int)std::min<size_t>(16, budget / smem_solve_w), w_setup = (int)std::min<size_t>(16, budget / smem_setup_w);
    if (team) w_solve = (int)std::min<size_t>(team_max_ctas(team), (budget + 1024) / (smem_solve_w + 1024)); // CTAs (= problems) per SM
    if (w_solve < 1 || w_setup < 1) { g_last_error = "daqp_b200: problem too large for shared memory"; return -2; }
    if (const char* wenv = getenv("DAQP_B200_WARPS")) w_solve = std::max(1, std::min(w_solve, atoi(wenv))); // tuning knob
    h->stats.warps_per_sm = w_solve;

    const size_t per = scratch_per_problem<T>(n, m, ldm, ldn);
    int chunk = (int)std::min<long long>(N, std::max<long long>(1, (h->scratch_limit - (8 << 20)) / (long long)per));
    if (ps) chunk = N; // a workspace owns its LDP arrays: one launch over the whole batch
    const int grid_max = h->num_sms;
    const size_t pst = (size_t)grid_max * 16 * cap * (sizeof(int) + sizeof(T)) + 4096;
    int rc = ensure(&h->arena, &h->arena_bytes, (ps ? 0 : (size_t)chunk * per) + pst + (1 << 20));
    if (rc) return rc;

    rc = arena_acquire(h, stream);
    if (rc) return rc;
    int tune = 0; // experiment knob: 1 = bulk L2 prefetch before the scan, 2 = no fp32 screening, 4 = stream policy for Mt32
    if (const char* tenv = getenv("DAQP_B200_TUNE")) tune = atoi(tenv);
    const bool screening = sizeof(T) == 8 && !(tune & 2) && (team ? (m + 31) / 32 <= TEAM_SCREEN_MAX_GROUPS * team : m <= (regstage ? 256 : 768));
    for (int p0 = 0; p0 < N; p0 += chunk) {
        const int P = std::min(chunk, N - p0);
        Carver cv(h->arena);
        int* counters = cv.take<int>(64);
        T* Mt = ps ? ps->Mt : cv.take<T>((size_t)P * n * ldm);
        T* Mr = ps ? ps->Mr : cv.take<T>((size_t)P * m * ldn);
        float* Mt32 = !screening ? nullptr : ps ? ps->Mt32 : cv.take<float>((size_t)P * ((n + 3) / 4) * m * 4);
        T* du = ps ? ps->du : cv.take<T>((size_t)P * ldm);
        T* dl = ps ? ps->dl : cv.take<T>((size_t)P * ldm);
        T* sc = ps ? ps->sc : cv.take<T>((size_t)P * ldm);
        T* Ri = ps ? ps->Ri : cv.take<T>((size_t)P * n * (n + 1) / 2);
        T* vv = ps ? ps->vv : cv.take<T>((size_t)P * n);
        unsigned char* sense8 = ps ? ps->sense8 : cv.take<unsigned char>((size_t)P * ldm);
        int* sflag = ps ? ps->sflag : cv.take<int>(P);
        T* xu_s = cv.take<T>((size_t)P * n);
        T* vnorm_s = cv.take<T>(P);
        int* info_s = cv.take<int>((size_t)P * 4);
        int* pst_id = cv.take<int>((size_t)grid_max * 16 * cap);
        T* pst_lam = cv.take<T>((size_t)grid_max * 16 * cap);

        EventTriple ev;
        rc = get_events(h, &ev);
        if (rc) return rc;
        CK(cudaMemsetAsync(counters, 0, 64 * sizeof(int), stream));
        CK(cudaEventRecord(ev.e0, stream));

        SetupArgs<T> sa;
        sa.P = P; sa.n = n; sa.m = m; sa.ms = ms; sa.ldm = ldm; sa.ldn = ldn;
        sa.H = dH + (size_t)p0 * n * n; sa.f = df ? df + (size_t)p0 * n : nullptr; sa.A = dA + (size_t)p0 * mA * n;
        sa.bupper = dbu + (size_t)p0 * m; sa.blower = dbl + (size_t)p0 * m;
        sa.sense_in = dsense ? dsense + (size_t)p0 * m : nullptr;
        sa.Mt = Mt; sa.Mr = Mr; sa.Mt32 = Mt32; sa.dupper = du; sa.dlower = dl; sa.scaling = sc; sa.Rinv = Ri; sa.v = vv;
        sa.sense = sense8; sa.setup_flag = sflag;
        sa.x = dx + (size_t)p0 * n; sa.lam = dlam ? dlam + (size_t)p0 * m : nullptr; sa.fval = dfval + p0;
        sa.exitflag = dflag + p0; sa.iter = diter + p0;
        sa.work_counter = counters; sa.st = st;
        sa.soft_slack = nullptr; sa.ns_max = ns_max;
        if constexpr (sizeof(T) == sizeof(c_float)) sa.soft_slack = (diag && diag->soft_slack) ? diag->soft_slack + p0 : nullptr;
        sa.no_shortcut = ps ? 1 : 0; sa.sense_static = ps ? ps->sense_static : nullptr;
        sa.xu = xu_s; sa.vnorm = vnorm_s; sa.info = info_s;
        bool split = sizeof(T) == 8 && n <= 127; // factor kernel + tensor-core product kernel
        if (const char* senv = getenv("DAQP_B200_SETUP")) { if (!strcmp(senv, "fused")) split = false; }
        if ((!ps || ps->phase == 1) && split) {
            cudaError_t e = cudaErrorNotSupported;
            if constexpr (sizeof(T) == 8) e = daqp_b200_launch_setup_split(sa, h->num_sms, h->smem_optin, stream);
            if (e != cudaSuccess) return fail("split setup launch", e, __LINE__);
            h->stats.setup_launches += 2;
        } else if (!ps || ps->phase == 1) {
            const int grid = std::min(grid_max, (P + w_setup - 1) / w_setup);
            const size_t smem = smem_setup_w * w_setup;
            cudaError_t e = cudaErrorInvalidValue;
            switch (ngs) {
#ifndef DAQP_B200_FAST_BUILD /* experiment builds instantiate the C3 shape only */
                case 1: e = launch_setup<T, 1>(sa, grid, 32 * w_setup, smem, stream); break;
#endif
                case 2: e = launch_setup<T, 2>(sa, grid, 32 * w_setup, smem, stream); break;
#ifndef DAQP_B200_FAST_BUILD
                case 3: e = launch_setup<T, 3>(sa, grid, 32 * w_setup, smem, stream); break;
                case 4: e = launch_setup<T, 4>(sa, grid, 32 * w_setup, smem, stream); break;
                default: e = launch_setup<T, 8>(sa, grid, 32 * w_setup, smem, stream); break;
#endif
            }
            if (e != cudaSuccess) return fail("qp_setup_kernel launch", e, __LINE__);
            h->stats.setup_launches++;
        }
        CK(cudaEventRecord(ev.e1, stream));

        la.P = P;
        la.Mt = Mt; la.Mr = Mr; la.Mt32 = Mt32; la.dupper = du; la.dlower = dl; la.scaling = sc; la.Rinv = Ri;
        la.v = df ? vv : nullptr; la.sense = sense8; la.setup_flag = sflag;
        la.x = sa.x; la.lam = sa.lam; la.fval = sa.fval; la.exitflag = sa.exitflag; la.iter = sa.iter;
        la.ws_out = (diag && diag->ws) ? diag->ws + (size_t)p0 * cap : nullptr;
        la.nact_out = (diag && diag->n_active) ? diag->n_active + p0 : nullptr;
        la.counts_out = (diag && diag->counts) ? diag->counts + 8 * (size_t)p0 : nullptr;
        la.sense_out = (diag && diag->sense) ? diag->sense + (size_t)p0 * ldm : nullptr;
        la.trace_cap = (diag && diag->trace) ? diag->trace_cap : 0;
        la.trace_out = (diag && diag->trace && diag->trace_cap > 0) ? diag->trace + (size_t)p0 * (1 + 2 * diag->trace_cap) : nullptr;
        la.work_counter = counters + 32; la.pst_id = pst_id; la.pst_lam = pst_lam; la.st = st;
        la.soft_slack = sa.soft_slack; la.ns_max = ns_max;
        la.tune = tune;
        if (ps) { la.state = ps->state; la.state_stride = ps->state_stride; la.state_load = ps->state_load; la.state_save = 1; la.grp = ps->grp; }
        if (team) {
            cudaError_t e = cudaErrorNotSupported;
            if constexpr (sizeof(T) == 8)
                e = daqp_b200_launch_solve_team(la, nv, std::min(grid_max * w_solve, P), smem_solve_w, stream);
            if (e != cudaSuccess) return fail("ldp_solve_kernel (team) launch", e, __LINE__);
            h->stats.solve_launches++;
        } else if (regstage) {
            cudaError_t e = cudaErrorNotSupported;
            if constexpr (sizeof(T) == 8)
                e = daqp_b200_launch_solve_regs(la, nv, std::min(grid_max, (P + w_solve - 1) / w_solve), 32 * w_solve, smem_solve_w * w_solve, stream);
            if (e != cudaSuccess) return fail("ldp_solve_kernel (register-staged) launch", e, __LINE__);
            h->stats.solve_launches++;
        } else if (!ps || ps->phase == 2) {
            const int grid = std::min(grid_max, (P + w_solve - 1) / w_solve);
            const size_t smem = smem_solve_w * w_solve;
            cudaError_t e = cudaErrorInvalidValue;
            switch (nv) {
#ifndef DAQP_B200_FAST_BUILD
                case 1: e = launch_solve<T, 1>(la, grid, 32 * w_solve, smem, stream); break;
#endif
                case 2: e = launch_solve<T, 2>(la, grid, 32 * w_solve, smem, stream); break;
#ifndef DAQP_B200_FAST_BUILD
                case 3: e = launch_solve<T, 3>(la, grid, 32 * w_solve, smem, stream); break;
                case 4: e = launch_solve<T, 4>(la, grid, 32 * w_solve, smem, stream); break;
                case 5: e = launch_solve<T, 5>(la, grid, 32 * w_solve, smem, stream); break;
                case 6: e = launch_solve<T, 6>(la, grid, 32 * w_solve, smem, stream); break;
                case 7: e = launch_solve<T, 7>(la, grid, 32 * w_solve, smem, stream); break;
                default: e = launch_solve<T, 8>(la, grid, 32 * w_solve, smem, stream); break;
#endif
            }
            if (e != cudaSuccess) return fail("ldp_solve_kernel launch", e, __LINE__);
            h->stats.solve_launches++;
        }
        CK(cudaEventRecord(ev.e2, stream));
        h->pending.push_back(ev);
    }
    return arena_release(h, stream);
}

extern "C" int daqp_b200_solve_device(DAQPB200Handle* h, int N, int n, int m, int ms, const c_float* dH,
                                      const c_float* df, const c_float* dA, const c_float* dbupper,
                                      const c_float* dblower, const int* dsense, const DAQPSettings* settings,
                                      c_float* dx, c_float* dlam, c_float* dfval, int* dexitflag, int* diter,
                                      const DAQPB200Diag* diag, void* stream) {
    if (!h) { int rc = default_handle(&h); if (rc) return rc; }
    std::lock_guard<std::mutex> lk(h->mu);
    CK(cudaSetDevice(h->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : h->compute;
    return solve_device_impl<double>(h, N, n, m, ms, dH, df, dA, dbupper, dblower, dsense, settings, dx, dlam, dfval,
                                     dexitflag, diter, diag, s);
}

// Host arrays: chunks are copied in on one stream while the previous chunk is solved on another and the one before
// that is copied out on a third (double-buffered device staging).
template <typename T>
static int solve_packed_impl(DAQPB200Handle* h, int N, int n, int m, int ms, const T* H, const T* f, const T* A,
                             const T* bupper, const T* blower, const int* sense, const DAQPSettings* settings, T* x,
                             T* lam, T* fval, int* exitflag, int* iter, const DAQPB200Diag* diag) {
    if (!h) { int rc = default_handle(&h); if (rc) return rc; }
    std::lock_guard<std::mutex> lk(h->mu);
    CK(cudaSetDevice(h->device));
    if (N <= 0) return 0;
    if (n < 1 || m < ms || ms < 0 || ms > n) { g_last_error = "daqp_b200: invalid problem dimensions"; return -2; }
    int ns_max = 0; // most soft constraints per problem: sizes the factor (and the ws diagnostic rows)
    if (sense)
        for (int p = 0; p < N; p++) {
            int c = 0;
            for (int i = 0; i < m; i++) c += (sense[(size_t)p * m + i] & DAQP_SOFT) ? 1 : 0;
            ns_max = std::max(ns_max, c);
        }
    const int mA = m - ms, cap = n + ns_max + 1, ldm = round_up(std::max(m, 1), 4);
    // The pipeline is bound by the host link (8.3 GB per 100k C3 problems at ~55 GB/s): what is left to tune is the part
    // that cannot overlap -- the first chunk's copy-in and the last chunk's solve -- so chunks are small and the first
    // and last one smaller still. Measured on C3 (ms per 100k): 16384 -> 174, 8192 -> 163, 6144 with 1024 ends -> 160.
    int chunk = 6144, first_chunk = 1024;
    if (const char* c = getenv("DAQP_B200_HOST_CHUNK")) { chunk = std::max(1, atoi(c)); first_chunk = chunk; }
    if (const char* c = getenv("DAQP_B200_HOST_FIRST_CHUNK")) first_chunk = std::max(1, atoi(c));
    chunk = std::min(chunk, N);
    first_chunk = std::min(first_chunk, chunk);
    const size_t in_b = ((size_t)n * n + n + (size_t)mA * n + 2 * (size_t)m) * sizeof(T) + (size_t)m * sizeof(int);
    const size_t out_b = ((size_t)n + m + 2) * sizeof(T) + 2 * sizeof(int) + (size_t)(cap + 1 + 8) * sizeof(int) + ldm;
    const size_t per_buf = ((size_t)chunk * (in_b + out_b) + 17 * 256 + 255) / 256 * 256; // buffer bases stay 256-byte aligned
    // three staging buffers: the copy-in of chunk c+2 does not have to wait for the solve of chunk c
    constexpr int NB = 3;
    int rc = ensure(&h->stage, &h->stage_bytes, NB * per_buf);
    if (rc) return rc;

    // ---- small batches (a single daqp_quadprog call, a handful of problems): latency, not bandwidth. Everything goes
    // through ONE pinned mirror of the staging buffer: one copy in, the two kernels, one copy out, one synchronisation on
    // one stream (the chunk pipeline below costs nine event creations and three stream synchronisations per call).
    if ((size_t)N * (in_b + out_b) <= ((size_t)1 << 20)) {
        if (h->pin_bytes < per_buf) {
            if (h->pin) { CK(cudaFreeHost(h->pin)); h->pin = nullptr; h->pin_bytes = 0; }
            CK(cudaHostAlloc((void**)&h->pin, per_buf, cudaHostAllocDefault));
            h->pin_bytes = per_buf;
        }
        struct View { T *H, *f, *A, *bu, *bl; int* sense; T *x, *lam, *fval, *slack; int *flag, *iter, *nact, *ws, *counts; unsigned char* so; char* out0; };
        auto carve = [&](char* base) {
            View v; Carver cv(base);
            v.H = cv.take<T>((size_t)chunk * n * n); v.f = cv.take<T>((size_t)chunk * n);
            v.A = cv.take<T>((size_t)chunk * mA * n); v.bu = cv.take<T>((size_t)chunk * m);
            v.bl = cv.take<T>((size_t)chunk * m); v.sense = cv.take<int>((size_t)chunk * m);
            v.x = cv.take<T>((size_t)chunk * n); v.out0 = reinterpret_cast<char*>(v.x);
            v.lam = cv.take<T>((size_t)chunk * m);
            v.fval = cv.take<T>(chunk); v.slack = cv.take<T>(chunk); v.flag = cv.take<int>(chunk); v.iter = cv.take<int>(chunk);
            v.nact = cv.take<int>(chunk); v.ws = cv.take<int>((size_t)chunk * cap);
            v.counts = cv.take<int>((size_t)chunk * 8); v.so = cv.take<unsigned char>((size_t)chunk * ldm);
            return std::make_pair(v, cv.off);
        };
        auto hv = carve(h->pin).first;
        auto dvp = carve(h->stage);
        auto dv = dvp.first;
        const size_t in_bytes = (size_t)(hv.out0 - h->pin), all_bytes = dvp.second;
        memcpy(hv.H, H, (size_t)N * n * n * sizeof(T));
        if (f) memcpy(hv.f, f, (size_t)N * n * sizeof(T));
        if (mA > 0) memcpy(hv.A, A, (size_t)N * mA * n * sizeof(T));
        if (m > 0) {
            memcpy(hv.bu, bupper, (size_t)N * m * sizeof(T)); memcpy(hv.bl, blower, (size_t)N * m * sizeof(T));
            if (sense) memcpy(hv.sense, sense, (size_t)N * m * sizeof(int));
        }
        if (fval) memcpy(hv.fval, fval, (size_t)N * sizeof(T)); // untouched entries (no linear term) keep the caller's value
        cudaStream_t s1 = h->compute;
        CK(cudaMemcpyAsync(h->stage, h->pin, in_bytes, cudaMemcpyHostToDevice, s1));
        if (fval) CK(cudaMemcpyAsync(dv.fval, hv.fval, (size_t)N * sizeof(T), cudaMemcpyHostToDevice, s1));
        DAQPB200Diag dd{};
        if (diag) { dd.n_active = diag->n_active ? dv.nact : nullptr; dd.ws = diag->ws ? dv.ws : nullptr;
                    dd.counts = diag->counts ? dv.counts : nullptr; dd.sense = diag->sense ? dv.so : nullptr;
                    if constexpr (sizeof(T) == sizeof(c_float)) dd.soft_slack = diag->soft_slack ? dv.slack : nullptr; }
        rc = solve_device_impl<T>(h, N, n, m, ms, dv.H, f ? dv.f : nullptr, dv.A, dv.bu, dv.bl, sense ? dv.sense : nullptr, settings,
                                  dv.x, lam ? dv.lam : nullptr, dv.fval, dv.flag, dv.iter, diag ? &dd : nullptr, s1, ns_max);
        if (rc) { cudaStreamSynchronize(s1); return rc; }
        CK(cudaMemcpyAsync(hv.out0, dv.out0, all_bytes - in_bytes, cudaMemcpyDeviceToHost, s1));
        CK(cudaStreamSynchronize(s1));
        memcpy(x, hv.x, (size_t)N * n * sizeof(T));
        if (lam && m > 0) memcpy(lam, hv.lam, (size_t)N * m * sizeof(T));
        if (fval) memcpy(fval, hv.fval, (size_t)N * sizeof(T));
        memcpy(exitflag, hv.flag, (size_t)N * sizeof(int));
        if (iter) memcpy(iter, hv.iter, (size_t)N * sizeof(int));
        if (diag) {
            if (diag->n_active) memcpy(diag->n_active, hv.nact, (size_t)N * sizeof(int));
            if (diag->ws) memcpy(diag->ws, hv.ws, (size_t)N * cap * sizeof(int));
            if (diag->counts) memcpy(diag->counts, hv.counts, (size_t)N * 8 * sizeof(int));
            if (diag->sense) memcpy(diag->sense, hv.so, (size_t)N * ldm);
            if constexpr (sizeof(T) == sizeof(c_float))
                if (diag->soft_slack) memcpy(diag->soft_slack, hv.slack, (size_t)N * sizeof(T));
        }
        return 0;
    }

    struct Buf { T *H, *f, *A, *bu, *bl; int* sense; T *x, *lam, *fval, *slack; int *flag, *iter, *nact, *ws, *counts; unsigned char* so; };
    Buf b[NB];
    for (int i = 0; i < NB; i++) {
        Carver cv(h->stage + i * per_buf);
        b[i].H = cv.take<T>((size_t)chunk * n * n); b[i].f = cv.take<T>((size_t)chunk * n);
        b[i].A = cv.take<T>((size_t)chunk * mA * n); b[i].bu = cv.take<T>((size_t)chunk * m);
        b[i].bl = cv.take<T>((size_t)chunk * m); b[i].sense = cv.take<int>((size_t)chunk * m);
        b[i].x = cv.take<T>((size_t)chunk * n); b[i].lam = cv.take<T>((size_t)chunk * m);
        b[i].fval = cv.take<T>(chunk); b[i].slack = cv.take<T>(chunk); b[i].flag = cv.take<int>(chunk); b[i].iter = cv.take<int>(chunk);
        b[i].nact = cv.take<int>(chunk); b[i].ws = cv.take<int>((size_t)chunk * cap);
        b[i].counts = cv.take<int>((size_t)chunk * 8); b[i].so = cv.take<unsigned char>((size_t)chunk * ldm);
    }
    cudaEvent_t ev_in[NB], ev_done[NB], ev_out[NB];
    for (int i = 0; i < NB; i++) {
        CK(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
    }
    int result = 0, c = 0;
    // chunk schedule: small first chunk (its copy-in is exposed), small last chunk (its solve and copy-out are exposed)
    std::vector<int> sched;
    for (int left = N; left > 0;) {
        int P = std::min(sched.empty() ? first_chunk : chunk, left);
        if (left > first_chunk && left - P < first_chunk) P = left - first_chunk; // leave exactly one small chunk
        sched.push_back(P);
        left -= P;
    }
    int p0 = 0;
    // one chunk: enqueue copy-in, solve, copy-out. A failing call returns from the lambda only -- the streams are always
    // synchronised and the events destroyed below, so no copy into the caller's arrays is in flight when this function
    // returns with an error.
    auto enqueue_chunk = [&](int P, int p0, int c) -> int {
        const int s = c % NB;
        Buf& B = b[s];
        // input buffers are free once the solve that read them (two chunks ago) has finished
        if (c >= NB) CK(cudaStreamWaitEvent(h->copy_in, ev_done[s], 0));
        CK(cudaMemcpyAsync(B.H, H + (size_t)p0 * n * n, (size_t)P * n * n * sizeof(T), cudaMemcpyHostToDevice, h->copy_in));
        if (f) CK(cudaMemcpyAsync(B.f, f + (size_t)p0 * n, (size_t)P * n * sizeof(T), cudaMemcpyHostToDevice, h->copy_in));
        if (mA > 0) CK(cudaMemcpyAsync(B.A, A + (size_t)p0 * mA * n, (size_t)P * mA * n * sizeof(T), cudaMemcpyHostToDevice, h->copy_in));
        if (m > 0) {
            CK(cudaMemcpyAsync(B.bu, bupper + (size_t)p0 * m, (size_t)P * m * sizeof(T), cudaMemcpyHostToDevice, h->copy_in));
            CK(cudaMemcpyAsync(B.bl, blower + (size_t)p0 * m, (size_t)P * m * sizeof(T), cudaMemcpyHostToDevice, h->copy_in));
            if (sense) CK(cudaMemcpyAsync(B.sense, sense + (size_t)p0 * m, (size_t)P * m * sizeof(int), cudaMemcpyHostToDevice, h->copy_in));
        }
        CK(cudaEventRecord(ev_in[s], h->copy_in));
        CK(cudaStreamWaitEvent(h->compute, ev_in[s], 0));
        if (c >= NB) CK(cudaStreamWaitEvent(h->compute, ev_out[s], 0)); // output buffers drained
        DAQPB200Diag dd{};
        if (diag) { dd.n_active = diag->n_active ? B.nact : nullptr; dd.ws = diag->ws ? B.ws : nullptr;
                    dd.counts = diag->counts ? B.counts : nullptr; dd.sense = diag->sense ? B.so : nullptr;
                    if constexpr (sizeof(T) == sizeof(c_float)) dd.soft_slack = diag->soft_slack ? B.slack : nullptr; }
        const int r2 = solve_device_impl<T>(h, P, n, m, ms, B.H, f ? B.f : nullptr, B.A, B.bu, B.bl, sense ? B.sense : nullptr,
                                            settings, B.x, lam ? B.lam : nullptr, B.fval, B.flag, B.iter, diag ? &dd : nullptr,
                                            h->compute, ns_max);
        if (r2) return r2;
        CK(cudaEventRecord(ev_done[s], h->compute));
        CK(cudaStreamWaitEvent(h->copy_out, ev_done[s], 0));
        CK(cudaMemcpyAsync(x + (size_t)p0 * n, B.x, (size_t)P * n * sizeof(T), cudaMemcpyDeviceToHost, h->copy_out));
        if (lam && m > 0) CK(cudaMemcpyAsync(lam + (size_t)p0 * m, B.lam, (size_t)P * m * sizeof(T), cudaMemcpyDeviceToHost, h->copy_out));
        if (fval) CK(cudaMemcpyAsync(fval + p0, B.fval, (size_t)P * sizeof(T), cudaMemcpyDeviceToHost, h->copy_out));
        CK(cudaMemcpyAsync(exitflag + p0, B.flag, (size_t)P * sizeof(int), cudaMemcpyDeviceToHost, h->copy_out));
        if (iter) CK(cudaMemcpyAsync(iter + p0, B.iter, (size_t)P * sizeof(int), cudaMemcpyDeviceToHost, h->copy_out));
        if (diag) {
            if (diag->n_active) CK(cudaMemcpyAsync(diag->n_active + p0, B.nact, (size_t)P * sizeof(int), cudaMemcpyDeviceToHost, h->copy_out));
            if (diag->ws) CK(cudaMemcpyAsync(diag->ws + (size_t)p0 * cap, B.ws, (size_t)P * cap * sizeof(int), cudaMemcpyDeviceToHost, h->copy_out));
            if (diag->counts) CK(cudaMemcpyAsync(diag->counts + (size_t)p0 * 8, B.counts, (size_t)P * 8 * sizeof(int), cudaMemcpyDeviceToHost, h->copy_out));
            if (diag->sense) CK(cudaMemcpyAsync(diag->sense + (size_t)p0 * ldm, B.so, (size_t)P * ldm, cudaMemcpyDeviceToHost, h->copy_out));
            if constexpr (sizeof(T) == sizeof(c_float))
                if (diag->soft_slack) CK(cudaMemcpyAsync(diag->soft_slack + p0, B.slack, (size_t)P * sizeof(T), cudaMemcpyDeviceToHost, h->copy_out));
        }
        CK(cudaEventRecord(ev_out[s], h->copy_out));
        return 0;
    };
    for (size_t ci = 0; ci < sched.size() && result == 0; p0 += sched[ci], ci++, c++) result = enqueue_chunk(sched[ci], p0, c);
    cudaError_t e1 = cudaStreamSynchronize(h->copy_in), e2 = cudaStreamSynchronize(h->compute), e3 = cudaStreamSynchronize(h->copy_out);
    for (int i = 0; i < NB; i++) { cudaEventDestroy(ev_in[i]); cudaEventDestroy(ev_done[i]); cudaEventDestroy(ev_out[i]); }
    if (result) return result;
    if (e1 != cudaSuccess) return fail("copy_in stream", e1, __LINE__);
    if (e2 != cudaSuccess) return fail("compute stream", e2, __LINE__);
    if (e3 != cudaSuccess) return fail("copy_out stream", e3, __LINE__);
    return 0;
}

extern "C" int daqp_b200_solve_packed(DAQPB200Handle* h, int N, int n, int m, int ms, const c_float* H,
                                      const c_float* f, const c_float* A, const c_float* bupper, const c_float* blower,
                                      const int* sense, const DAQPSettings* settings, c_float* x, c_float* lam,
                                      c_float* fval, int* exitflag, int* iter, const DAQPB200Diag* diag) {
    return solve_packed_impl<c_float>(h, N, n, m, ms, H, f, A, bupper, blower, sense, settings, x, lam, fval, exitflag,
                                      iter, diag);
}

// fp32 arithmetic end to end: the batched form of the reference built with -DDAQP_SINGLE_PRECISION (include/types.h:8-12)
extern "C" int daqp_b200_solve_packed_f32(DAQPB200Handle* h, int N, int n, int m, int ms, const float* H, const float* f,
                                          const float* A, const float* bupper, const float* blower, const int* sense,
                                          const DAQPSettings* settings, float* x, float* lam, float* fval,
                                          int* exitflag, int* iter, const DAQPB200Diag* diag) {
    return solve_packed_impl<float>(h, N, n, m, ms, H, f, A, bupper, blower, sense, settings, x, lam, fval, exitflag,
                                    iter, diag);
}

extern "C" int daqp_b200_solve_device_f32(DAQPB200Handle* h, int N, int n, int m, int ms, const float* dH,
                                          const float* df, const float* dA, const float* dbupper, const float* dblower,
                                          const int* dsense, const DAQPSettings* settings, float* dx, float* dlam,
                                          float* dfval, int* dexitflag, int* diter, const DAQPB200Diag* diag,
                                          void* stream) {
    if (!h) { int rc = default_handle(&h); if (rc) return rc; }
    std::lock_guard<std::mutex> lk(h->mu);
    CK(cudaSetDevice(h->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : h->compute;
    return solve_device_impl<float>(h, N, n, m, ms, dH, df, dA, dbupper, dblower, dsense, settings, dx, dlam, dfval,
                                    dexitflag, diter, diag, s);
}

// ---- batched minimal representation (reference daqp_minrep: src/api.c:531-556, src/utils.c:808-835) ---------------
// Device arrays in, device array out: P polyhedra x m constraints = P m LDPs, solved concurrently by the solve kernel in
// shared-matrix mode (minrep_kernel.cuh). Chunked over polyhedra when the scratch limit asks for it.
// (`single` switches to ONE raw LDP per polyhedron with the caller's lower bounds and sense bits: daqp_b200_ldp_batch)
struct LdpSingle { const c_float* dbl; const int* dsense; c_float *dx, *dlam, *dfval; int *dnact, *dws; unsigned char* dsense_out; };
static int minrep_device_impl(DAQPB200Handle* h, int P, int n, int m, int ms, const c_float* dA, const c_float* db,
                              const DAQPSettings* settings, int* dred, int* dflag_out, int* diter_out,
                              cudaStream_t stream, const unsigned char* ddropped = nullptr, const LdpSingle* single = nullptr) {
    if (P <= 0 || m <= 0) return 0;
    if (n < 1 || m < ms || ms < 0 || ms > n) { g_last_error = "daqp_b200: invalid polyhedron dimensions"; return -2; }
    typedef c_float T;
    constexpr int V = VecOf<T>::N;
    const int ldm = round_up(m, 4), ldn = round_up(n, V), cap = n + 1, mA = m - ms, nv = (cap + 31) / 32;
    if (nv > 8) { g_last_error = "daqp_b200: n > 255 is not supported"; return -2; }
    LdpArgs<T> la;
    memset(&la, 0, sizeof(la));
    la.n = n; la.m = m; la.ms = ms; la.ldm = ldm; la.ldn = ldn; la.cap = cap;
    const size_t smem_w = ldp_layout<T>(la);
    int w_solve = (int)std::min<size_t>(16, h->smem_optin / smem_w);
    if (w_solve < 1) { g_last_error = "daqp_b200: polyhedron too large for shared memory"; return -2; }
    h->stats.warps_per_sm = w_solve;
    const size_t ntri = (size_t)n * (n + 1) / 2;
    const size_t per_poly = ((size_t)n * ldm + (size_t)m * ldn + ldm + ntri) * sizeof(T) + 4 * 256;
    const size_t per_ldp = ((size_t)2 * ldm + n + 1) * sizeof(T) + ldm + 3 * sizeof(int);
    const int K = single ? 1 : m; // LDPs per polyhedron
    const size_t per = per_poly + (size_t)K * per_ldp + (single ? (size_t)n * sizeof(T) : 0);
    int chunk = (int)std::min<long long>(P, std::max<long long>(1, (h->scratch_limit - (8 << 20)) / (long long)per));
    chunk = (int)std::min<long long>(chunk, (long long)(INT_MAX / 2) / m); // LDP indices are ints
    const int grid_max = h->num_sms;
    const size_t pst = (size_t)grid_max * 16 * cap * (sizeof(int) + sizeof(T)) + 4096;
    int rc = ensure(&h->arena, &h->arena_bytes, (size_t)chunk * per + pst + (1 << 20));
    if (rc) return rc;
    rc = arena_acquire(h, stream);
    if (rc) return rc;
    const DevSettings<T> st = to_dev_settings<T>(settings);
    for (int q0 = 0; q0 < P; q0 += chunk) {
        const int Q = std::min(chunk, P - q0);
        const size_t NL = (size_t)Q * K;
        Carver cv(h->arena);
        int* counters = cv.take<int>(64);
        MinrepArgs ma;
        ma.P = Q; ma.n = n; ma.m = m; ma.ms = ms; ma.ldm = ldm; ma.ldn = ldn;
        ma.A = dA + (size_t)q0 * mA * n; ma.b = db + (size_t)q0 * m;
        ma.dropped = ddropped ? ddropped + (size_t)q0 * m : nullptr;
        ma.Mt = cv.take<T>((size_t)Q * n * ldm); ma.Mr = cv.take<T>((size_t)Q * m * ldn);
        ma.scaling = cv.take<T>((size_t)Q * ldm); ma.Rinv = cv.take<T>((size_t)Q * ntri);
        ma.dupper = cv.take<T>(NL * ldm); ma.dlower = cv.take<T>(NL * ldm);
        ma.sense = cv.take<unsigned char>(NL * ldm);
        ma.setup_flag = cv.take<int>(NL);
        ma.exitflag = dflag_out ? dflag_out + (size_t)q0 * K : cv.take<int>(NL);
        ma.iter = diter_out ? diter_out + (size_t)q0 * K : cv.take<int>(NL);
        T* xs = (single && single->dx) ? single->dx + (size_t)q0 * n : cv.take<T>(NL * n);
        T* fv = (single && single->dfval) ? single->dfval + q0 : cv.take<T>(NL);
        T* vzero = single ? cv.take<T>((size_t)Q * n) : nullptr;
        int* pst_id = cv.take<int>((size_t)grid_max * 16 * cap);
        T* pst_lam = cv.take<T>((size_t)grid_max * 16 * cap);

        EventTriple ev;
        rc = get_events(h, &ev);
        if (rc) return rc;
        CK(cudaMemsetAsync(counters, 0, 64 * sizeof(int), stream));
        CK(cudaEventRecord(ev.e0, stream));
        if (single) ldp_prep_kernel<<<std::min(Q, 8 * grid_max), 256, 0, stream>>>(ma, single->dbl + (size_t)q0 * m,
                                                                                 single->dsense ? single->dsense + (size_t)q0 * m : nullptr, vzero);
        else minrep_prep_kernel<<<std::min(Q, 8 * grid_max), 256, 0, stream>>>(ma);
        CK(cudaGetLastError());
        h->stats.setup_launches++;
        CK(cudaEventRecord(ev.e1, stream));

        la.P = (int)NL; la.grp = K;
        la.Mt = ma.Mt; la.Mr = ma.Mr; la.Mt32 = nullptr; la.scaling = ma.scaling; la.Rinv = ma.Rinv; la.v = vzero;
        la.dupper = ma.dupper; la.dlower = ma.dlower; la.sense = ma.sense; la.setup_flag = ma.setup_flag;
        la.x = xs; la.lam = (single && single->dlam) ? single->dlam + (size_t)q0 * m : nullptr; la.fval = fv;
        la.exitflag = ma.exitflag; la.iter = ma.iter;
        if (single) {
            la.nact_out = single->dnact ? single->dnact + q0 : nullptr;
            la.ws_out = single->dws ? single->dws + (size_t)q0 * cap : nullptr;
            la.sense_out = single->dsense_out ? single->dsense_out + (size_t)q0 * ldm : nullptr;
        }
        la.work_counter = counters + 32; la.pst_id = pst_id; la.pst_lam = pst_lam; la.st = st;
        const int grid = (int)std::min<size_t>(grid_max, (NL + w_solve - 1) / w_solve);
        const size_t smem = smem_w * w_solve;
        cudaError_t e = cudaErrorInvalidValue;
        switch (nv) {
#ifndef DAQP_B200_FAST_BUILD
            case 1: e = launch_solve<T, 1>(la, grid, 32 * w_solve, smem, stream); break;
#endif
            case 2: e = launch_solve<T, 2>(la, grid, 32 * w_solve, smem, stream); break;
#ifndef DAQP_B200_FAST_BUILD
            case 3: e = launch_solve<T, 3>(la, grid, 32 * w_solve, smem, stream); break;
            case 4: e = launch_solve<T, 4>(la, grid, 32 * w_solve, smem, stream); break;
            case 5: e = launch_solve<T, 5>(la, grid, 32 * w_solve, smem, stream); break;
            case 6: e = launch_solve<T, 6>(la, grid, 32 * w_solve, smem, stream); break;
            case 7: e = launch_solve<T, 7>(la, grid, 32 * w_solve, smem, stream); break;
            default: e = launch_solve<T, 8>(la, grid, 32 * w_solve, smem, stream); break;
#endif
        }
        if (e != cudaSuccess) return fail("ldp_solve_kernel launch (minrep)", e, __LINE__);
        h->stats.solve_launches++;
        if (!single) {
            minrep_finish_kernel<<<(int)std::min<size_t>(1024, (NL + 255) / 256), 256, 0, stream>>>(ma.exitflag, dred + (size_t)q0 * m, NL);
            CK(cudaGetLastError());
        }
        CK(cudaEventRecord(ev.e2, stream));
        h->pending.push_back(ev);
        if (q0 + chunk < P) CK(cudaStreamSynchronize(stream)); // the next chunk reuses the scratch
    }
    return arena_release(h, stream);
}

extern "C" int daqp_b200_minrep_device(DAQPB200Handle* h, int P, int n, int m, int ms, const c_float* dA,
                                       const c_float* db, const DAQPSettings* settings, int* dis_redundant,
                                       int* dexitflag, int* diter, void* stream) {
    if (!h) { int rc = default_handle(&h); if (rc) return rc; }
    std::lock_guard<std::mutex> lk(h->mu);
    CK(cudaSetDevice(h->device));
    return minrep_device_impl(h, P, n, m, ms, dA, db, settings, dis_redundant, dexitflag, diter,
                              stream ? (cudaStream_t)stream : h->compute);
}

// ---- P raw LDPs in one call (batched form of daqp_ldp on hand-filled workspaces, api.jl:440-459) --------------------------
extern "C" int daqp_b200_ldp_device(DAQPB200Handle* h, int P, int n, int m, int ms, const c_float* dA, const c_float* dbupper,
                                    const c_float* dblower, const int* dsense, const DAQPSettings* settings, c_float* du,
                                    c_float* dlam, c_float* dfval, int* dexitflag, int* diter, void* stream) {
    if (!h) { int rc = default_handle(&h); if (rc) return rc; }
    std::lock_guard<std::mutex> lk(h->mu);
    CK(cudaSetDevice(h->device));
    LdpSingle sg{dblower, dsense, du, dlam, dfval, nullptr, nullptr, nullptr};
    return minrep_device_impl(h, P, n, m, ms, dA, dbupper, settings, nullptr, dexitflag, diter,
                              stream ? (cudaStream_t)stream : h->compute, nullptr, &sg);
}

extern "C" int daqp_b200_ldp_batch(DAQPB200Handle* h, int P, int n, int m, int ms, const c_float* A, const c_float* bupper,
                                   const c_float* blower, const int* sense, const DAQPSettings* settings, c_float* u, c_float* lam,
                                   c_float* fval, int* exitflag, int* iter, const DAQPB200Diag* diag) {
    if (!h) { int rc = default_handle(&h); if (rc) return rc; }
    std::lock_guard<std::mutex> lk(h->mu);
    CK(cudaSetDevice(h->device));
    if (P <= 0 || m <= 0) return 0;
    if (n < 1 || m < ms || ms < 0 || ms > n) { g_last_error = "daqp_b200: invalid LDP dimensions"; return -2; }
    const size_t mA = (size_t)(m - ms), nA = (size_t)P * mA * n, nb = (size_t)P * m, nx = (size_t)P * n;
    const int ldm = round_up(m, 4), cap = n + 1;
    int rc = ensure(&h->stage, &h->stage_bytes, (nA + 3 * nb + nx + P) * sizeof(c_float) + (nb + (size_t)P * (3 + cap)) * sizeof(int) +
                                                    (size_t)P * ldm + 16 * 256);
    if (rc) return rc;
    Carver cv(h->stage);
    int* dnact = cv.take<int>((size_t)P); int* dws = cv.take<int>((size_t)P * cap);
    unsigned char* dso = cv.take<unsigned char>((size_t)P * ldm);
    c_float* dA = cv.take<c_float>(nA); c_float* dbu = cv.take<c_float>(nb); c_float* dbl = cv.take<c_float>(nb);
    c_float* dlam = cv.take<c_float>(nb); c_float* dx = cv.take<c_float>(nx); c_float* dfv = cv.take<c_float>((size_t)P);
    int* dse = cv.take<int>(nb); int* dflag = cv.take<int>((size_t)P); int* dit = cv.take<int>((size_t)P);
    cudaStream_t s = h->compute;
    if (nA) CK(cudaMemcpyAsync(dA, A, nA * sizeof(c_float), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dbu, bupper, nb * sizeof(c_float), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dbl, blower, nb * sizeof(c_float), cudaMemcpyHostToDevice, s));
    if (sense) CK(cudaMemcpyAsync(dse, sense, nb * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(dfv, 0, (size_t)P * sizeof(c_float), s));
    LdpSingle sg{dbl, sense ? dse : nullptr, dx, dlam, dfv, dnact, dws, dso};
    rc = minrep_device_impl(h, P, n, m, ms, dA, dbu, settings, nullptr, dflag, dit, s, nullptr, &sg);
    if (rc) return rc;
    if (diag) { // working sets in factor order ([P][n + 1]), their sizes, final sense bytes ([P][ldm])
        if (diag->n_active) CK(cudaMemcpyAsync(diag->n_active, dnact, (size_t)P * sizeof(int), cudaMemcpyDeviceToHost, s));
        if (diag->ws) CK(cudaMemcpyAsync(diag->ws, dws, (size_t)P * cap * sizeof(int), cudaMemcpyDeviceToHost, s));
        if (diag->sense) CK(cudaMemcpyAsync(diag->sense, dso, (size_t)P * ldm, cudaMemcpyDeviceToHost, s));
    }
    if (u) CK(cudaMemcpyAsync(u, dx, nx * sizeof(c_float), cudaMemcpyDeviceToHost, s));
    if (lam) CK(cudaMemcpyAsync(lam, dlam, nb * sizeof(c_float), cudaMemcpyDeviceToHost, s));
    if (fval) CK(cudaMemcpyAsync(fval, dfv, (size_t)P * sizeof(c_float), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(exitflag, dflag, (size_t)P * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (iter) CK(cudaMemcpyAsync(iter, dit, (size_t)P * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

// One round on host arrays: stage, run, copy back. `dropped` ([P][m] bytes) may be NULL.
static int minrep_host_round(DAQPB200Handle* h, int P, int n, int m, int ms, const c_float* A, const c_float* b,
                             const unsigned char* dropped, const DAQPSettings* settings, int* is_redundant,
                             int* exitflag, int* iter) {
    const size_t mA = (size_t)(m - ms), nA = (size_t)P * mA * n, nb = (size_t)P * m;
    int rc = ensure(&h->stage, &h->stage_bytes, (nA + nb) * sizeof(c_float) + 3 * nb * sizeof(int) + nb + 8 * 256);
    if (rc) return rc;
    Carver cv(h->stage);
    c_float* dA = cv.take<c_float>(nA);
    c_float* db = cv.take<c_float>(nb);
    int* dred = cv.take<int>(nb);
    int* dflag = cv.take<int>(nb);
    int* dit = cv.take<int>(nb);
    unsigned char* ddr = cv.take<unsigned char>(nb);
    cudaStream_t s = h->compute;
    if (nA) CK(cudaMemcpyAsync(dA, A, nA * sizeof(c_float), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(db, b, nb * sizeof(c_float), cudaMemcpyHostToDevice, s));
    if (dropped) CK(cudaMemcpyAsync(ddr, dropped, nb, cudaMemcpyHostToDevice, s));
    rc = minrep_device_impl(h, P, n, m, ms, dA, db, settings, dred, dflag, dit, s, dropped ? ddr : nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(is_redundant, dred, nb * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(exitflag, dflag, nb * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(iter, dit, nb * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

extern "C" int daqp_b200_minrep_batch(DAQPB200Handle* h, int P, int n, int m, int ms, const c_float* A,
                                      const c_float* b, const DAQPSettings* settings, int* is_redundant,
                                      int* exitflag, int* iter) {
    if (!h) { int rc = default_handle(&h); if (rc) return rc; }
    std::lock_guard<std::mutex> lk(h->mu);
    CK(cudaSetDevice(h->device));
    if (P <= 0 || m <= 0) return 0;
    if (n < 1 || m < ms || ms < 0 || ms > n) { g_last_error = "daqp_b200: invalid polyhedron dimensions"; return -2; }
    const size_t mA = (size_t)(m - ms), nb = (size_t)P * m;
    std::vector<int> flag_own, iter_own;
    if (!exitflag) { flag_own.resize(nb); exitflag = flag_own.data(); }
    if (!iter) { iter_own.resize(nb); iter = iter_own.data(); }
    int rc = minrep_host_round(h, P, n, m, ms, A, b, nullptr, settings, is_redundant, exitflag, iter);
    if (rc) return rc;
    // Empty polyhedra (every probe infeasible): the reference's answer is the one its probing ORDER produces -- constraints
    // are dropped from the front until the rest is non-empty (utils.c:814-824). Re-run those polyhedra with their first
    // remaining constraint dropped until a probe is feasible; everything else is final after the first round.
    std::vector<int> todo;
    std::vector<unsigned char> dropped; // [todo.size()][m]
    for (int q = 0; q < P; q++) {
        bool all = true;
        for (int i = 0; i < m && all; i++) all = is_redundant[(size_t)q * m + i] == 1;
        if (all) todo.push_back(q);
    }
    dropped.assign(todo.size() * (size_t)m, 0);
    std::vector<c_float> As, bs;
    std::vector<int> rs, fs, is;
    while (!todo.empty()) {
        // drop the first remaining constraint of every polyhedron still undecided
        std::vector<int> next;
        std::vector<unsigned char> ndrop;
        for (size_t t = 0; t < todo.size(); t++) {
            unsigned char* d = &dropped[t * m];
            int first = 0;
            while (first < m && d[first]) first++;
            if (first >= m - 1) continue; // one constraint (or none) left: it keeps the value of its last probe
            d[first] = 1;
            next.push_back(todo[t]);
            ndrop.insert(ndrop.end(), d, d + m);
        }
        todo.swap(next); dropped.swap(ndrop);
        if (todo.empty()) break;
        const size_t T = todo.size();
        As.resize(T * mA * n); bs.resize(T * m); rs.resize(T * m); fs.resize(T * m); is.resize(T * m);
        for (size_t t = 0; t < T; t++) {
            if (mA) memcpy(&As[t * mA * n], A + (size_t)todo[t] * mA * n, mA * n * sizeof(c_float));
            memcpy(&bs[t * m], b + (size_t)todo[t] * m, (size_t)m * sizeof(c_float));
        }
        rc = minrep_host_round(h, (int)T, n, m, ms, As.data(), bs.data(), dropped.data(), settings, rs.data(), fs.data(), is.data());
        if (rc) return rc;
        next.clear(); ndrop.clear();
        for (size_t t = 0; t < T; t++) {
            const size_t o = (size_t)todo[t] * m;
            bool all = true;
            for (int i = 0; i < m; i++) {
                is_redundant[o + i] = rs[t * m + i];
                if (!dropped[t * m + i]) { exitflag[o + i] = fs[t * m + i]; iter[o + i] = is[t * m + i]; }
                all = all && rs[t * m + i] == 1;
            }
            if (all) { next.push_back(todo[t]); ndrop.insert(ndrop.end(), &dropped[t * m], &dropped[t * m] + m); }
        }
        todo.swap(next); dropped.swap(ndrop);
    }
    return 0;
}

// Drop-in for the reference's daqp_minrep (include/api.h:54, src/api.c:531-556): one polyhedron, its m LDPs in one launch.
// Like the reference it has no way to report an error; on failure is_redundant is filled with -1 (the reference's
// "not decided" marker, src/utils.c:811-812) and daqp_b200_last_error() says why.
extern "C" void daqp_minrep(int* is_redundant, c_float* A, c_float* b, int n, int m, int ms) {
    if (daqp_b200_minrep_batch(nullptr, 1, n, m, ms, A, b, nullptr, is_redundant, nullptr, nullptr) != 0)
        for (int i = 0; i < m; i++) is_redundant[i] = -1;
}

// ---- warm-start initialisers (reference daqp_primal_init_active / daqp_dual_init_active, src/api.c:577-631) --------
static int init_active_launch(DAQPB200Handle* h, int N, int n, int m, int ms, const c_float* dx, const c_float* dlam,
                              const c_float* dA, const c_float* dbu, const c_float* dbl, int* dsense, cudaStream_t s) {
    if (N <= 0 || m <= 0) return 0;
    if (n < 1 || m < ms || ms < 0 || ms > n) { g_last_error = "daqp_b200: invalid problem dimensions"; return -2; }
    if (!dx && !dlam) { g_last_error = "daqp_b200: init_active needs a primal or a dual iterate"; return -2; }
    InitActiveArgs ia;
    ia.N = N; ia.n = n; ia.m = m; ia.ms = ms; ia.x = dx; ia.lam = dlam; ia.A = dA; ia.bupper = dbu; ia.blower = dbl;
    ia.sense = dsense;
    const size_t smem = ia_smem_per_warp(n) * IA_WARPS;
    if (smem > h->smem_optin) { g_last_error = "daqp_b200: n too large for the init_active tile"; return -2; }
    CK(cudaFuncSetAttribute(init_active_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin_smem()));
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, h->smem_optin / (smem + 1024)));
    init_active_kernel<<<std::min(h->num_sms * per_sm, (N + IA_WARPS - 1) / IA_WARPS), 32 * IA_WARPS, smem, s>>>(ia);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int daqp_b200_init_active_device(DAQPB200Handle* h, int N, int n, int m, int ms, const c_float* dx,
                                            const c_float* dlam, const c_float* dA, const c_float* dbupper,
                                            const c_float* dblower, int* dsense, void* stream) {
    if (!h) { int rc = default_handle(&h); if (rc) return rc; }
    std::lock_guard<std::mutex> lk(h->mu);
    CK(cudaSetDevice(h->device));
    return init_active_launch(h, N, n, m, ms, dx, dlam, dA, dbupper, dblower, dsense,
                              stream ? (cudaStream_t)stream : h->compute);
}

extern "C" int daqp_b200_init_active(DAQPB200Handle* h, int N, int n, int m, int ms, const c_float* x,
                                     const c_float* lam, const c_float* A, const c_float* bupper,
                                     const c_float* blower, int* sense) {
    if (!h) { int rc = default_handle(&h); if (rc) return rc; }
    std::lock_guard<std::mutex> lk(h->mu);
    CK(cudaSetDevice(h->device));
    if (N <= 0 || m <= 0) return 0;
    if (n < 1 || m < ms || ms < 0 || ms > n) { g_last_error = "daqp_b200: invalid problem dimensions"; return -2; }
    if (!x && !lam) { g_last_error = "daqp_b200: init_active needs a primal or a dual iterate"; return -2; }
    const size_t mA = (size_t)(m - ms), nA = x ? (size_t)N * mA * n : 0, nb = (size_t)N * m, nx = x ? (size_t)N * n : 0;
    int rc = ensure(&h->stage, &h->stage_bytes, (nA + 3 * nb + nx) * sizeof(c_float) + nb * sizeof(int) + 8 * 256);
    if (rc) return rc;
    Carver cv(h->stage);
    c_float* dA = cv.take<c_float>(nA); c_float* dx = cv.take<c_float>(nx); c_float* dl = cv.take<c_float>(nb);
    c_float* dbu = cv.take<c_float>(nb); c_float* dbl = cv.take<c_float>(nb); int* dse = cv.take<int>(nb);
    cudaStream_t s = h->compute;
    if (x) {
        if (nA) CK(cudaMemcpyAsync(dA, A, nA * sizeof(c_float), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(dx, x, nx * sizeof(c_float), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(dbu, bupper, nb * sizeof(c_float), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(dbl, blower, nb * sizeof(c_float), cudaMemcpyHostToDevice, s));
    } else {
        CK(cudaMemcpyAsync(dl, lam, nb * sizeof(c_float), cudaMemcpyHostToDevice, s));
    }
    CK(cudaMemcpyAsync(dse, sense, nb * sizeof(int), cudaMemcpyHostToDevice, s));
    rc = init_active_launch(h, N, n, m, ms, x ? dx : nullptr, x ? nullptr : dl, dA, dbu, dbl, dse, s);
    if (rc) return rc;
    CK(cudaMemcpyAsync(sense, dse, nb * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

// ---- daqp_first_violating, batched (reference src/api.c:562-574): N points against one polyhedron ------------------------
extern "C" int daqp_b200_first_violating_batch(DAQPB200Handle* h, int N, int n, int m, int ms, const c_float* x, const c_float* A,
                                               const c_float* bupper, const c_float* blower, c_float tol, int* first) {
    if (!h) { int rc = default_handle(&h); if (rc) return rc; }
    std::lock_guard<std::mutex> lk(h->mu);
    CK(cudaSetDevice(h->device));
    if (N <= 0) return 0;
    if (n < 1 || m < ms || ms < 0 || ms > n) { g_last_error = "daqp_b200: invalid polyhedron dimensions"; return -2; }
    const size_t mA = (size_t)(m - ms), nx = (size_t)N * n, nA = mA * n;
    int rc = ensure(&h->stage, &h->stage_bytes, (nx + nA + 2 * (size_t)m) * sizeof(c_float) + (size_t)N * sizeof(int) + 8 * 256);
    if (rc) return rc;
    Carver cv(h->stage);
    c_float* dx = cv.take<c_float>(nx); c_float* dA = cv.take<c_float>(nA);
    c_float* dbu = cv.take<c_float>(m); c_float* dbl = cv.take<c_float>(m); int* dout = cv.take<int>(N);
    cudaStream_t s = h->compute;
    CK(cudaMemcpyAsync(dx, x, nx * sizeof(c_float), cudaMemcpyHostToDevice, s));
    if (nA) CK(cudaMemcpyAsync(dA, A, nA * sizeof(c_float), cudaMemcpyHostToDevice, s));
    if (m) {
        CK(cudaMemcpyAsync(dbu, bupper, (size_t)m * sizeof(c_float), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(dbl, blower, (size_t)m * sizeof(c_float), cudaMemcpyHostToDevice, s));
    }
    const int warps = 8;
    first_violating_kernel<<<std::min(h->num_sms * 8, (N + warps - 1) / warps), 32 * warps, (size_t)warps * n * sizeof(c_float), s>>>(
        N, n, m, ms, dx, dA, dbu, dbl, tol, dout);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(first, dout, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

// Drop-ins (reference include/api.h:57-58): one problem, qp->sense updated in place. Like the reference they return
// nothing; without a CUDA device qp->sense is left untouched and daqp_b200_last_error() says why.
extern "C" void daqp_primal_init_active(DAQPProblem* qp, c_float* x) {
    if (!qp || !qp->sense || !x) return;
    daqp_b200_init_active(nullptr, 1, qp->n, qp->m, qp->ms, x, nullptr, qp->A, qp->bupper, qp->blower, qp->sense);
}
extern "C" void daqp_dual_init_active(DAQPProblem* qp, c_float* lam) {
    if (!qp || !qp->sense || !lam) return;
    daqp_b200_init_active(nullptr, 1, qp->n, qp->m, qp->ms, nullptr, lam, qp->A, qp->bupper, qp->blower, qp->sense);
}

// ---- persistent batch workspace: setup once, update(f, b) + solve many (reference setup_daqp / daqp_update_ldp /
// daqp_solve on a kept DAQPWorkspace: src/api.c:88-160, src/utils.c:58-221, docs/docs/c.md:44-77) ---------------
struct DAQPB200Workspace {
    DAQPB200Handle* h = nullptr;
    int N = 0, n = 0, m = 0, ms = 0, ldm = 0, ns_max = 0, cap = 0;
    bool has_f = false, has_sense = false;
    bool need_data = false; // shared workspace: f and both bounds have not been given yet
    DAQPSettings settings{};
    Persist<c_float> ps;
    std::vector<void*> owned; // every device allocation of the workspace
    bool counted = false;     // registered in the handle's live_workspaces
    c_float *d_f = nullptr, *d_bu = nullptr, *d_bl = nullptr;                           // inputs of update
    c_float *d_x = nullptr, *d_lam = nullptr, *d_fval = nullptr, *d_slack = nullptr;    // outputs of solve
    int *d_flag = nullptr, *d_iter = nullptr, *d_nact = nullptr, *d_ws = nullptr, *d_counts = nullptr;
    int* d_sense = nullptr; // user sense (device copy; only its non-NULL-ness matters after the setup)
    unsigned char* d_so = nullptr;
};

template <typename U> static int ws_alloc(DAQPB200Workspace* w, U** p, size_t count) {
    CK(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(U)));
    w->owned.push_back(*p);
    return 0;
}

extern "C" void daqp_b200_workspace_free(DAQPB200Workspace* w) {
    if (!w) return;
    cudaSetDevice(w->h->device);
    cudaDeviceSynchronize();
    for (void* p : w->owned) cudaFree(p);
    if (w->counted) { std::lock_guard<std::mutex> lk(w->h->mu); w->h->live_workspaces--; } // (a failed setup frees with mu held: not counted yet)
    delete w;
}

// K == 0: N problems, each with its own H / A / f / bounds (the plain workspace). K > 0: N matrix sets (H, A, sense), K
// problems per set whose f / bounds arrive with the first update (shared workspace).
static int workspace_setup_impl(DAQPB200Handle* h, int N, int K, int n, int m, int ms, const c_float* H,
                                const c_float* f, const c_float* A, const c_float* bupper,
                                const c_float* blower, const int* sense, const DAQPSettings* settings,
                                DAQPB200Workspace** out) {
    typedef c_float T;
    if (!h) { int rc = default_handle(&h); if (rc) return rc; }
    std::lock_guard<std::mutex> lk(h->mu);
    CK(cudaSetDevice(h->device));
    if (N <= 0 || K < 0 || n < 1 || m < ms || ms < 0 || ms > n || (K > 0 && (long long)N * K > INT_MAX / 2)) {
        g_last_error = "daqp_b200: invalid problem dimensions"; return -2;
    }
    const int G = N;           // matrix sets
    if (K > 0) N = G * K;      // problems
    DAQPB200Workspace* w = new DAQPB200Workspace();
    w->h = h; w->N = N; w->n = n; w->m = m; w->ms = ms; w->ldm = round_up(std::max(m, 1), 4);
    w->has_f = K > 0 || f != nullptr; w->has_sense = sense != nullptr; w->need_data = K > 0;
    if (settings) w->settings = *settings; else daqp_default_settings(&w->settings);
    if (sense)
        for (int p = 0; p < G; p++) {
            int c = 0;
            for (int i = 0; i < m; i++) c += (sense[(size_t)p * m + i] & DAQP_SOFT) ? 1 : 0;
            w->ns_max = std::max(w->ns_max, c);
        }
    w->cap = n + w->ns_max + 1;
    const int ldm = w->ldm, ldn = round_up(n, VecOf<T>::N), mA = m - ms;
    LdpArgs<T> la;
    memset(&la, 0, sizeof(la));
    la.n = n; la.m = m; la.ms = ms; la.ldm = ldm; la.ldn = ldn; la.cap = w->cap;
    ldp_layout<T>(la);
    Persist<T>& ps = w->ps;
    ps.state_stride = (unsigned)round_up(la.oarena, 16) + 16;
    ps.grp = K;
    int rc = 0;
    T *dH = nullptr, *dA = nullptr;
#define WS_TRY(call) do { rc = (call); if (rc) { daqp_b200_workspace_free(w); return rc; } } while (0)
    // the matrix arrays hold G sets, everything else N problems
    WS_TRY(ws_alloc(w, &ps.Mt, (size_t)G * n * ldm));
    WS_TRY(ws_alloc(w, &ps.Mr, (size_t)G * m * ldn));
    if (m <= 768) WS_TRY(ws_alloc(w, &ps.Mt32, (size_t)G * ((n + 3) / 4) * m * 4));
    WS_TRY(ws_alloc(w, &ps.du, (size_t)N * ldm)); WS_TRY(ws_alloc(w, &ps.dl, (size_t)N * ldm));
    WS_TRY(ws_alloc(w, &ps.sc, (size_t)G * ldm)); WS_TRY(ws_alloc(w, &ps.Ri, (size_t)G * n * (n + 1) / 2));
    WS_TRY(ws_alloc(w, &ps.vv, (size_t)N * n)); WS_TRY(ws_alloc(w, &ps.sense8, (size_t)N * ldm));
    WS_TRY(ws_alloc(w, &ps.sense_static, (size_t)G * ldm)); WS_TRY(ws_alloc(w, &ps.sflag, (size_t)N));
    WS_TRY(ws_alloc(w, &ps.state, (size_t)N * ps.state_stride));
    WS_TRY(ws_alloc(w, &w->d_f, (size_t)N * n)); WS_TRY(ws_alloc(w, &w->d_bu, (size_t)N * m)); WS_TRY(ws_alloc(w, &w->d_bl, (size_t)N * m));
    WS_TRY(ws_alloc(w, &w->d_x, (size_t)N * n)); WS_TRY(ws_alloc(w, &w->d_lam, (size_t)N * m));
    WS_TRY(ws_alloc(w, &w->d_fval, (size_t)N)); WS_TRY(ws_alloc(w, &w->d_slack, (size_t)N));
    WS_TRY(ws_alloc(w, &w->d_flag, (size_t)N)); WS_TRY(ws_alloc(w, &w->d_iter, (size_t)N));
    WS_TRY(ws_alloc(w, &w->d_nact, (size_t)N)); WS_TRY(ws_alloc(w, &w->d_ws, (size_t)N * w->cap));
    WS_TRY(ws_alloc(w, &w->d_counts, (size_t)N * 8)); WS_TRY(ws_alloc(w, &w->d_so, (size_t)N * ldm));
    if (sense) WS_TRY(ws_alloc(w, &w->d_sense, (size_t)G * m));
    WS_TRY(ws_alloc(w, &dH, (size_t)G * n * n)); WS_TRY(ws_alloc(w, &dA, (size_t)G * std::max(mA, 1) * n));
    cudaStream_t st = h->compute;
#define WS_CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { daqp_b200_workspace_free(w); return fail(#call, e_, __LINE__); } } while (0)
    WS_CK(cudaMemsetAsync(ps.state, 0, (size_t)N * ps.state_stride, st));
    WS_CK(cudaMemsetAsync(w->d_fval, 0, (size_t)N * sizeof(T), st));
    WS_CK(cudaMemsetAsync(w->d_flag, 0, (size_t)N * sizeof(int), st));
    WS_CK(cudaMemsetAsync(w->d_iter, 0, (size_t)N * sizeof(int), st));
    WS_CK(cudaMemcpyAsync(dH, H, (size_t)G * n * n * sizeof(T), cudaMemcpyHostToDevice, st));
    if (mA > 0) WS_CK(cudaMemcpyAsync(dA, A, (size_t)G * mA * n * sizeof(T), cudaMemcpyHostToDevice, st));
    if (m > 0 && sense) WS_CK(cudaMemcpyAsync(w->d_sense, sense, (size_t)G * m * sizeof(int), cudaMemcpyHostToDevice, st));
    if (K == 0) {
        if (f) WS_CK(cudaMemcpyAsync(w->d_f, f, (size_t)N * n * sizeof(T), cudaMemcpyHostToDevice, st));
        if (m > 0) {
            WS_CK(cudaMemcpyAsync(w->d_bu, bupper, (size_t)N * m * sizeof(T), cudaMemcpyHostToDevice, st));
            WS_CK(cudaMemcpyAsync(w->d_bl, blower, (size_t)N * m * sizeof(T), cudaMemcpyHostToDevice, st));
        }
        ps.phase = 1;
        WS_TRY(solve_device_impl<T>(h, N, n, m, ms, dH, f ? w->d_f : nullptr, dA, w->d_bu, w->d_bl, w->d_sense, &w->settings,
                                    w->d_x, w->d_lam, w->d_fval, w->d_flag, w->d_iter, nullptr, st, w->ns_max, &ps));
    } else {
        // Shared workspace: the QP -> LDP transform runs once per matrix set, with open bounds and no linear term; what it
        // writes per PROBLEM (d, v, sense, flags) goes to scratch -- the first update produces the real ones. Only the flags
        // the Hessian raises (non-convex, singular) are kept: they hold for every problem of the set.
        T *tbu = nullptr, *tbl = nullptr, *tdu = nullptr, *tdl = nullptr, *tv = nullptr, *tx = nullptr, *tlam = nullptr, *tfv = nullptr;
        unsigned char* tse = nullptr;
        int *tsf = nullptr, *tfl = nullptr, *tit = nullptr;
        WS_TRY(ws_alloc(w, &tbu, (size_t)G * m)); WS_TRY(ws_alloc(w, &tbl, (size_t)G * m));
        WS_TRY(ws_alloc(w, &tdu, (size_t)G * ldm)); WS_TRY(ws_alloc(w, &tdl, (size_t)G * ldm));
        WS_TRY(ws_alloc(w, &tv, (size_t)G * n)); WS_TRY(ws_alloc(w, &tx, (size_t)G * n));
        WS_TRY(ws_alloc(w, &tlam, (size_t)G * m)); WS_TRY(ws_alloc(w, &tfv, (size_t)G));
        WS_TRY(ws_alloc(w, &tse, (size_t)G * ldm)); WS_TRY(ws_alloc(w, &tsf, (size_t)G));
        WS_TRY(ws_alloc(w, &tfl, (size_t)G)); WS_TRY(ws_alloc(w, &tit, (size_t)G));
        std::vector<T> open_u((size_t)G * std::max(m, 1), (T)DAQP_INF), open_l((size_t)G * std::max(m, 1), (T)-DAQP_INF);
        if (m > 0) {
            WS_CK(cudaMemcpyAsync(tbu, open_u.data(), (size_t)G * m * sizeof(T), cudaMemcpyHostToDevice, st));
            WS_CK(cudaMemcpyAsync(tbl, open_l.data(), (size_t)G * m * sizeof(T), cudaMemcpyHostToDevice, st));
        }
        WS_CK(cudaMemsetAsync(tfl, 0, (size_t)G * sizeof(int), st));
        Persist<T> p1 = ps;
        p1.du = tdu; p1.dl = tdl; p1.vv = tv; p1.sense8 = tse; p1.sflag = tsf; p1.grp = 0; p1.phase = 1;
        WS_TRY(solve_device_impl<T>(h, G, n, m, ms, dH, nullptr, dA, tbu, tbl, w->d_sense, &w->settings, tx, tlam, tfv, tfl,
                                    tit, nullptr, st, w->ns_max, &p1));
        std::vector<int> gflag(G), pflag((size_t)N);
        WS_CK(cudaMemcpyAsync(gflag.data(), tfl, (size_t)G * sizeof(int), cudaMemcpyDeviceToHost, st));
        WS_CK(cudaStreamSynchronize(st));
        for (int g = 0; g < G; g++)
            for (int k = 0; k < K; k++)
                pflag[(size_t)g * K + k] = (gflag[g] == DAQP_EXIT_NONCONVEX || gflag[g] == DAQP_EXIT_UNSUPPORTED) ? gflag[g] : 0;
        WS_CK(cudaMemcpyAsync(w->d_flag, pflag.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, st));
        for (void* t : {(void*)tbu, (void*)tbl, (void*)tdu, (void*)tdl, (void*)tv, (void*)tx, (void*)tlam, (void*)tfv,
                        (void*)tse, (void*)tsf, (void*)tfl, (void*)tit}) {
            WS_CK(cudaStreamSynchronize(st));
            cudaFree(t);
            w->owned.erase(std::remove(w->owned.begin(), w->owned.end(), t), w->owned.end());
        }
    }
    WS_CK(cudaStreamSynchronize(st));
    // H and A are not needed again: the LDP is what the workspace keeps
    cudaFree(dH); cudaFree(dA);
    w->owned.erase(std::remove(w->owned.begin(), w->owned.end(), (void*)dH), w->owned.end());
    w->owned.erase(std::remove(w->owned.begin(), w->owned.end(), (void*)dA), w->owned.end());
#undef WS_TRY
#undef WS_CK
    w->counted = true;
    h->live_workspaces++; // (h->mu is held)
    *out = w;
    return 0;
}

extern "C" int daqp_b200_workspace_setup(DAQPB200Handle* h, int N, int n, int m, int ms, const c_float* H,
                                         const c_float* f, const c_float* A, const c_float* bupper,
                                         const c_float* blower, const int* sense, const DAQPSettings* settings,
                                         DAQPB200Workspace** out) {
    return workspace_setup_impl(h, N, 0, n, m, ms, H, f, A, bupper, blower, sense, settings, out);
}

extern "C" int daqp_b200_workspace_setup_shared(DAQPB200Handle* h, int G, int K, int n, int m, int ms, const c_float* H,
                                                const c_float* A, const int* sense, const DAQPSettings* settings,
                                                DAQPB200Workspace** out) {
    if (K < 1) { g_last_error = "daqp_b200: a shared workspace needs K >= 1 problems per matrix set"; return -2; }
    return workspace_setup_impl(h, G, K, n, m, ms, H, nullptr, A, nullptr, nullptr, sense, settings, out);
}

extern "C" int daqp_b200_workspace_flags(DAQPB200Workspace* w, int* exitflag) {
    if (!w) { g_last_error = "daqp_b200: null workspace"; return -2; }
    std::lock_guard<std::mutex> lk(w->h->mu);
    CK(cudaSetDevice(w->h->device));
    CK(cudaMemcpyAsync(exitflag, w->d_flag, (size_t)w->N * sizeof(int), cudaMemcpyDeviceToHost, w->h->compute));
    CK(cudaStreamSynchronize(w->h->compute));
    return 0;
}

// kind: host arrays are staged with cudaMemcpyHostToDevice, device arrays with DeviceToDevice (the workspace keeps its
// own copy of the current f / bounds either way: a later update may replace only some of them)
static int workspace_update_impl(DAQPB200Workspace* w, const c_float* f, const c_float* bupper, const c_float* blower,
                                 cudaMemcpyKind kind, cudaStream_t st) {
    typedef c_float T;
    DAQPB200Handle* h = w->h;
    const int N = w->N, n = w->n, m = w->m;
    if (f && !w->has_f) { g_last_error = "daqp_b200: the workspace was set up without a linear term"; return -2; }
    if (w->need_data) {
        if (!f || !bupper || !blower) {
            g_last_error = "daqp_b200: the first update of a shared workspace must give f, bupper and blower"; return -2;
        }
        w->need_data = false;
    }
    if (f) CK(cudaMemcpyAsync(w->d_f, f, (size_t)N * n * sizeof(T), kind, st));
    if (bupper) CK(cudaMemcpyAsync(w->d_bu, bupper, (size_t)N * m * sizeof(T), kind, st));
    if (blower) CK(cudaMemcpyAsync(w->d_bl, blower, (size_t)N * m * sizeof(T), kind, st));
    UpdateArgs<T> ua;
    ua.P = N; ua.n = n; ua.m = m; ua.ms = w->ms; ua.ldm = w->ldm; ua.grp = w->ps.grp;
    ua.f = f ? w->d_f : nullptr; ua.bupper = w->d_bu; ua.blower = w->d_bl;
    ua.Rinv = w->ps.Ri; ua.Mt = w->ps.Mt; ua.scaling = w->ps.sc; ua.sense_static = w->ps.sense_static;
    ua.v = w->ps.vv; ua.dupper = w->ps.du; ua.dlower = w->ps.dl; ua.sense = w->ps.sense8;
    ua.setup_flag = w->ps.sflag; ua.exitflag = w->d_flag; ua.iter = w->d_iter;
    ua.st = to_dev_settings<T>(&w->settings);
    const int warps = 8;
    const size_t smem = (size_t)warps * 2 * n * sizeof(T);
    ldp_update_kernel<T><<<std::min(h->num_sms * 4, (N + warps - 1) / warps), 32 * warps, smem, st>>>(ua);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int daqp_b200_workspace_update(DAQPB200Workspace* w, const c_float* f, const c_float* bupper,
                                          const c_float* blower) {
    if (!w) { g_last_error = "daqp_b200: null workspace"; return -2; }
    std::lock_guard<std::mutex> lk(w->h->mu);
    CK(cudaSetDevice(w->h->device));
    return workspace_update_impl(w, f, bupper, blower, cudaMemcpyHostToDevice, w->h->compute);
}

extern "C" int daqp_b200_workspace_update_device(DAQPB200Workspace* w, const c_float* df, const c_float* dbupper,
                                                 const c_float* dblower, void* stream) {
    if (!w) { g_last_error = "daqp_b200: null workspace"; return -2; }
    std::lock_guard<std::mutex> lk(w->h->mu);
    CK(cudaSetDevice(w->h->device));
    return workspace_update_impl(w, df, dbupper, dblower, cudaMemcpyDeviceToDevice,
                                 stream ? (cudaStream_t)stream : w->h->compute);
}

extern "C" int daqp_b200_workspace_solve(DAQPB200Workspace* w, int warm, c_float* x, c_float* lam, c_float* fval,
                                         int* exitflag, int* iter, const DAQPB200Diag* diag) {
    typedef c_float T;
    if (!w) { g_last_error = "daqp_b200: null workspace"; return -2; }
    DAQPB200Handle* h = w->h;
    std::lock_guard<std::mutex> lk(h->mu);
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->compute;
    const int N = w->N, n = w->n, m = w->m;
    if (w->need_data) { g_last_error = "daqp_b200: shared workspace has no f / bounds yet (call update first)"; return -2; }
    DAQPB200Diag dd{};
    dd.n_active = w->d_nact; dd.ws = w->d_ws; dd.counts = w->d_counts; dd.sense = w->d_so; dd.soft_slack = w->d_slack;
    w->ps.phase = 2; w->ps.state_load = warm ? 1 : 0;
    int rc = solve_device_impl<T>(h, N, n, m, w->ms, nullptr, w->has_f ? w->d_f : nullptr, nullptr, w->d_bu, w->d_bl,
                                  w->has_sense ? w->d_sense : nullptr, &w->settings, w->d_x, w->d_lam, w->d_fval,
                                  w->d_flag, w->d_iter, &dd, st, w->ns_max, &w->ps);
    if (rc) return rc;
    CK(cudaMemcpyAsync(x, w->d_x, (size_t)N * n * sizeof(T), cudaMemcpyDeviceToHost, st));
    if (lam && m > 0) CK(cudaMemcpyAsync(lam, w->d_lam, (size_t)N * m * sizeof(T), cudaMemcpyDeviceToHost, st));
    if (fval) CK(cudaMemcpyAsync(fval, w->d_fval, (size_t)N * sizeof(T), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(exitflag, w->d_flag, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (iter) CK(cudaMemcpyAsync(iter, w->d_iter, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (diag) {
        if (diag->n_active) CK(cudaMemcpyAsync(diag->n_active, w->d_nact, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (diag->ws) CK(cudaMemcpyAsync(diag->ws, w->d_ws, (size_t)N * w->cap * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (diag->counts) CK(cudaMemcpyAsync(diag->counts, w->d_counts, (size_t)N * 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (diag->sense) CK(cudaMemcpyAsync(diag->sense, w->d_so, (size_t)N * w->ldm, cudaMemcpyDeviceToHost, st));
        if (diag->soft_slack) CK(cudaMemcpyAsync(diag->soft_slack, w->d_slack, (size_t)N * sizeof(T), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

// Device-resident variant for closed loops that live on the GPU: results go straight to the caller's device arrays,
// nothing is copied to the host and nothing synchronises.
extern "C" int daqp_b200_workspace_solve_device(DAQPB200Workspace* w, int warm, c_float* dx, c_float* dlam,
                                                c_float* dfval, int* dexitflag, int* diter, void* stream) {
    typedef c_float T;
    if (!w) { g_last_error = "daqp_b200: null workspace"; return -2; }
    DAQPB200Handle* h = w->h;
    std::lock_guard<std::mutex> lk(h->mu);
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->compute;
    if (w->need_data) { g_last_error = "daqp_b200: shared workspace has no f / bounds yet (call update first)"; return -2; }
    // the update kernel records setup failures (infeasible bounds) in the workspace's own flag / iter arrays: carry them over
    CK(cudaMemcpyAsync(dexitflag, w->d_flag, (size_t)w->N * sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (diter) CK(cudaMemcpyAsync(diter, w->d_iter, (size_t)w->N * sizeof(int), cudaMemcpyDeviceToDevice, st));
    w->ps.phase = 2; w->ps.state_load = warm ? 1 : 0;
    int rc = solve_device_impl<T>(h, w->N, w->n, w->m, w->ms, nullptr, w->has_f ? w->d_f : nullptr, nullptr, w->d_bu, w->d_bl,
                                  w->has_sense ? w->d_sense : nullptr, &w->settings, dx, dlam, dfval ? dfval : w->d_fval,
                                  dexitflag, diter ? diter : w->d_iter, nullptr, st, w->ns_max, &w->ps);
    if (rc) return rc;
    // keep the workspace's flag array current: the next update consults it for setups the Hessian made permanent failures
    CK(cudaMemcpyAsync(w->d_flag, dexitflag, (size_t)w->N * sizeof(int), cudaMemcpyDeviceToDevice, st));
    return 0;
}

// ---- branch and bound over binary constraints: the node relaxations as a batched LDP consumer --------------------------
// reference src/bnb.c:23-128 (daqp_bnb, daqp_process_node, daqp_get_branch_id, daqp_spawn_children). The reference walks
// the tree depth first, one relaxation (= one daqp_ldp on the same workspace) at a time. Here the TREE stays on the host
// and every WAVE of open nodes is one launch of the solve kernel in shared-matrix mode: the QP -> LDP transform runs once
// for the problem, all nodes read the same device copy of the matrices, a node is nothing but its own sense bytes -- the
// binaries fixed on the way down as ACTIVE + IMMUTABLE (+ LOWER) rows, the parent's final working set as warm-start bits.
// Pruning is the reference's: the incumbent's objective goes into settings.fval_bound, a node whose objective passes it
// ends INFEASIBLE inside the kernel (daqp.c:19-23). The branching rule (first free binary in index order that is not
// within primal_tol of an endpoint, nearer endpoint first) is bnb.c:130-158 evaluated on x. What differs from the
// reference is the ORDER in which nodes are visited (waves of the deepest open nodes instead of strict depth first), so
// `nodes` and `iter` count a different walk to the same optimum.
__global__ void bnb_wave_kernel(int N, int ldm, const unsigned char* base, const unsigned char* node, const int* used,
                                unsigned char* sense, int* sflag) {
    const int p = blockIdx.x;
    if (p >= N) return;
    for (int r = threadIdx.x; r < ldm; r += blockDim.x)
        sense[(size_t)p * ldm + r] = base[(size_t)p * ldm + r] | node[(size_t)p * ldm + r];
    if (threadIdx.x == 0) sflag[p] = used[p] ? SETUP_SOLVE_ACTIVATE : EXIT_INFEASIBLE; // unused slots are passed through
}

extern "C" int daqp_b200_bnb(DAQPB200Handle* h, const DAQPProblem* qp, const DAQPSettings* settings_in, DAQPResult* res,
                             int wave_width) {
    if (!h) { int rc = default_handle(&h); if (rc) return rc; }
    if (!qp || !res || qp->H == nullptr || qp->nh > 1 || qp->problem_type != 0 || qp->n < 1 || qp->m < qp->ms || qp->ms > qp->n ||
        !qp->sense || (qp->m > qp->ms && qp->A == nullptr) || qp->bupper == nullptr || qp->blower == nullptr) {
        g_last_error = "daqp_b200: invalid problem for branch and bound"; return -2;
    }
    const int n = qp->n, m = qp->m, ms = qp->ms, ldm = round_up(std::max(m, 1), 4);
    const int W = std::max(1, wave_width > 0 ? wave_width : 256);
    DAQPSettings st;
    if (settings_in) st = *settings_in; else daqp_default_settings(&st);
    std::vector<int> bin_ids, sense_base((size_t)m);
    for (int i = 0; i < m; i++) {
        if (qp->sense[i] & DAQP_BINARY) bin_ids.push_back(i);
        sense_base[i] = qp->sense[i] & ~DAQP_BINARY; // the relaxation sees a binary as a two-sided inequality
    }
    res->nodes = 0; res->iter = 0; res->soft_slack = 0; res->setup_time = 0; res->solve_time = 0;
    if ((int)bin_ids.size() > n) { res->exitflag = DAQP_EXIT_OVERDETERMINED_INITIAL; return 0; } // api.c:226
    // ---- one transform, W node slots
    DAQPB200Workspace* w = nullptr;
    int rc = workspace_setup_impl(h, 1, W, n, m, ms, qp->H, nullptr, qp->A, nullptr, nullptr, sense_base.data(), &st, &w);
    if (rc) return rc;
    struct Guard { DAQPB200Workspace* w; ~Guard() { daqp_b200_workspace_free(w); } } guard{w};
    {
        std::vector<c_float> fK((size_t)W * n, 0), buK((size_t)W * m), blK((size_t)W * m);
        for (int k = 0; k < W; k++) {
            if (qp->f) memcpy(&fK[(size_t)k * n], qp->f, sizeof(c_float) * n);
            memcpy(&buK[(size_t)k * m], qp->bupper, sizeof(c_float) * m);
            memcpy(&blK[(size_t)k * m], qp->blower, sizeof(c_float) * m);
        }
        rc = daqp_b200_workspace_update(w, fK.data(), buK.data(), blK.data());
        if (rc) return rc;
    }
    std::vector<int> flag0((size_t)W);
    rc = daqp_b200_workspace_flags(w, flag0.data());
    if (rc) return rc;
    if (flag0[0] < 0) { res->exitflag = flag0[0]; return 0; } // the transform failed (non-convex, infeasible bounds, ...)
    // sense bytes after the update (bound-derived equalities included), |v|^2 for the objective bound, device scratch
    unsigned char *d_base = nullptr, *d_node = nullptr;
    int* d_used = nullptr;
    c_float vnorm = 0;
    {
        std::lock_guard<std::mutex> lk(h->mu);
        CK(cudaSetDevice(h->device));
        CK(cudaMalloc((void**)&d_base, (size_t)W * ldm)); w->owned.push_back(d_base);
        CK(cudaMalloc((void**)&d_node, (size_t)W * ldm)); w->owned.push_back(d_node);
        CK(cudaMalloc((void**)&d_used, (size_t)W * sizeof(int))); w->owned.push_back(d_used);
        CK(cudaMemcpyAsync(d_base, w->ps.sense8, (size_t)W * ldm, cudaMemcpyDeviceToDevice, h->compute));
        std::vector<c_float> v((size_t)n);
        CK(cudaMemcpyAsync(v.data(), w->ps.vv, sizeof(c_float) * n, cudaMemcpyDeviceToHost, h->compute));
        CK(cudaStreamSynchronize(h->compute));
        if (qp->f) for (int i = 0; i < n; i++) vnorm += v[i] * v[i];
    }
    struct Node { std::vector<int> fixed; std::vector<unsigned char> warm; }; // fixed: id or id | (1 << 16) for "at lower"
    std::vector<Node> open(1);
    const c_float eps_r = 1 / (1 + st.rel_subopt);
    c_float bound = (st.fval_bound - st.abs_subopt) * eps_r; // bnb.c:29-31
    bool have_inc = false;
    c_float best_internal = 0, best_slack = 0;
    std::vector<c_float> best_x((size_t)n), best_lam((size_t)std::max(m, 1));
    std::vector<c_float> x((size_t)W * n), lam((size_t)W * std::max(m, 1)), fval((size_t)W), slack((size_t)W);
    std::vector<int> flag((size_t)W), iter((size_t)W), used((size_t)W), nact((size_t)W);
    std::vector<unsigned char> node8((size_t)W * ldm), so((size_t)W * ldm);
    int last_fail = 0;
    const auto t0 = std::chrono::steady_clock::now();
    while (!open.empty()) {
        if (st.time_limit > 0 && // the limit also holds across the tree (bnb.c:52-60), checked once per wave
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > st.time_limit) {
            last_fail = DAQP_EXIT_TIMELIMIT;
            break;
        }
        const int cnt = (int)std::min<size_t>(W, open.size());
        std::vector<Node> wave(std::make_move_iterator(open.end() - cnt), std::make_move_iterator(open.end())); // the deepest nodes
        open.resize(open.size() - cnt);
        std::fill(node8.begin(), node8.end(), 0);
        for (int k = 0; k < W; k++) used[k] = k < cnt;
        for (int k = 0; k < cnt; k++) {
            unsigned char* s8 = &node8[(size_t)k * ldm];
            if (!wave[k].warm.empty()) memcpy(s8, wave[k].warm.data(), (size_t)m);
            for (int fx : wave[k].fixed) {
                const int id = fx & 0xffff;
                s8[id] = (unsigned char)((sense_base[id] & ~(DAQP_ACTIVE | DAQP_LOWER)) | DAQP_ACTIVE | DAQP_IMMUTABLE | ((fx >> 16) ? DAQP_LOWER : 0));
            }
        }
        {
            std::lock_guard<std::mutex> lk(h->mu);
            CK(cudaSetDevice(h->device));
            CK(cudaMemcpyAsync(d_node, node8.data(), (size_t)W * ldm, cudaMemcpyHostToDevice, h->compute));
            CK(cudaMemcpyAsync(d_used, used.data(), (size_t)W * sizeof(int), cudaMemcpyHostToDevice, h->compute));
            bnb_wave_kernel<<<W, 64, 0, h->compute>>>(W, ldm, d_base, d_node, d_used, w->ps.sense8, w->ps.sflag);
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(h->compute)); // (node8 / used are reused by the next wave)
            w->settings.fval_bound = bound;
        }
        DAQPB200Diag dg{};
        dg.n_active = nact.data(); dg.sense = so.data(); dg.soft_slack = slack.data();
        rc = daqp_b200_workspace_solve(w, 0, x.data(), lam.data(), fval.data(), flag.data(), iter.data(), &dg);
        if (rc) return rc;
        for (int k = 0; k < cnt; k++) {
            res->nodes++;
            res->iter += iter[k];
            const int fl = flag[k];
            if (fl == DAQP_EXIT_OVERDETERMINED_INITIAL && !wave[k].warm.empty()) {
                // a fixed binary clashed with the warm-start rows, which are only a guess: once more from the fixings alone
                Node cold; cold.fixed = wave[k].fixed;
                open.push_back(std::move(cold));
                continue;
            }
            if (fl == DAQP_EXIT_INFEASIBLE || fl == DAQP_EXIT_OVERDETERMINED_INITIAL) continue; // cut (dominated, or the fixings clash)
            if (fl < 0) { last_fail = fl; open.clear(); break; }                                   // inner solver failed (bnb.c:63)
            const c_float* xk = &x[(size_t)k * n];
            const unsigned char* sk = &so[(size_t)k * ldm];
            const c_float internal = 2 * fval[k] + vnorm; // work->fval of this node (api.c:471-477 inverted)
            if (have_inc && !(internal <= 2 * bound)) continue; // the bound moved while the wave was in flight
            // daqp_get_branch_id (bnb.c:130-158) on x: first free binary that is not within primal_tol of an endpoint
            int branch = -1;
            for (int id : bin_ids) {
                if (sk[id] & DAQP_ACTIVE) continue;
                c_float val = 0;
                if (id < ms) val = xk[id];
                else { const c_float* row = qp->A + (size_t)(id - ms) * n; for (int j = 0; j < n; j++) val += row[j] * xk[j]; }
                const c_float diff = 0.5 * (qp->bupper[id] + qp->blower[id]) - val;
                const c_float dist = 0.5 * (qp->bupper[id] - qp->blower[id]) - (diff < 0 ? -diff : diff);
                if (dist <= st.primal_tol) continue;
                branch = diff < 0 ? id : (id | (1 << 16)); // explore the endpoint nearest to the relaxation first
                break;
            }
            if (branch < 0) { // integer feasible: new incumbent (bnb.c:66-69)
                if (!have_inc || internal < best_internal) {
                    have_inc = true; best_internal = internal; best_slack = slack[k];
                    memcpy(best_x.data(), xk, sizeof(c_float) * n);
                    if (m > 0) memcpy(best_lam.data(), &lam[(size_t)k * m], sizeof(c_float) * m);
                    bound = (0.5 * internal - st.abs_subopt) * eps_r;
                }
                continue;
            }
            Node far, near; // bnb.c:160-176: the far endpoint is pushed first, the near one is processed first
            far.fixed = wave[k].fixed; far.fixed.push_back(branch ^ (1 << 16));
            near.fixed = wave[k].fixed; near.fixed.push_back(branch);
            far.warm.assign((size_t)m, 0);
            for (int i = 0; i < m; i++) // the parent's working set as warm-start bits (daqp_save_warmstart, bnb.c:209-221)
                if ((sk[i] & DAQP_ACTIVE) && !(sk[i] & DAQP_IMMUTABLE)) far.warm[i] = (unsigned char)((sense_base[i] & ~DAQP_LOWER) | DAQP_ACTIVE | (sk[i] & DAQP_LOWER));
            near.warm = far.warm;
            open.push_back(std::move(far));
            open.push_back(std::move(near));
        }
    }
    res->solve_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!have_inc) { res->exitflag = last_fail < 0 ? last_fail : DAQP_EXIT_INFEASIBLE; return 0; }
    memcpy(res->x, best_x.data(), sizeof(c_float) * n);
    if (res->lam && m > 0) memcpy(res->lam, best_lam.data(), sizeof(c_float) * m);
    if (qp->f) res->fval = 0.5 * (best_internal - vnorm);
    res->soft_slack = best_slack;
    res->exitflag = last_fail < DAQP_EXIT_INFEASIBLE ? last_fail : DAQP_EXIT_OPTIMAL; // bnb.c:84
    return 0;
}

// ---- one process, several GPUs: the batch is cut into contiguous blocks, one host thread + engine per device ---------
extern "C" int daqp_b200_solve_packed_multi(int ndev, const int* devices, int N, int n, int m, int ms, const c_float* H,
                                            const c_float* f, const c_float* A, const c_float* bupper, const c_float* blower,
                                            const int* sense, const DAQPSettings* settings, c_float* x, c_float* lam,
                                            c_float* fval, int* exitflag, int* iter, double* seconds_per_device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        g_last_error = "daqp_b200: no CUDA device available (this library has no CPU path)";
        return -1;
    }
    if (ndev <= 0) ndev = count;
    if (N <= 0) return 0;
    const int mA = m - ms;
    std::vector<int> rcs((size_t)ndev, 0);
    std::vector<std::string> errs((size_t)ndev);
    std::vector<std::thread> th;
    const int base = N / ndev, extra = N % ndev;
    int lo = 0;
    for (int r = 0; r < ndev; r++) {
        const int cnt = base + (r < extra ? 1 : 0), p0 = lo, dev = devices ? devices[r] : r % count;
        lo += cnt;
        th.emplace_back([=, &rcs, &errs]() {
            const auto t0 = std::chrono::steady_clock::now();
            int rc = 0;
            if (cnt > 0) {
                if (cudaSetDevice(dev) != cudaSuccess) rc = -1;
                DAQPB200Handle* h = nullptr;
                if (!rc) rc = default_handle(&h);
                if (!rc)
                    rc = daqp_b200_solve_packed(h, cnt, n, m, ms, H + (size_t)p0 * n * n, f ? f + (size_t)p0 * n : nullptr,
                                                A + (size_t)p0 * mA * n, bupper + (size_t)p0 * m, blower + (size_t)p0 * m,
                                                sense ? sense + (size_t)p0 * m : nullptr, settings, x + (size_t)p0 * n,
                                                lam ? lam + (size_t)p0 * m : nullptr, fval ? fval + p0 : nullptr, exitflag + p0,
                                                iter ? iter + p0 : nullptr, nullptr);
                if (rc) errs[r] = g_last_error; // (thread-local: carry it to the caller's thread)
            }
            rcs[r] = rc;
            if (seconds_per_device) seconds_per_device[r] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        });
    }
    for (auto& t : th) t.join();
    for (int r = 0; r < ndev; r++)
        if (rcs[r]) { g_last_error = errs[r]; return rcs[r]; }
    return 0;
}

// cudaHostRegister / cudaHostUnregister for callers without a CUDA toolchain: pinned arrays make the chunked copies of
// daqp_b200_solve_packed asynchronous (pageable memory is staged by the driver, at about half the rate)
extern "C" int daqp_b200_pin(void* ptr, size_t bytes) { CK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault)); return 0; }
extern "C" int daqp_b200_unpin(void* ptr) { CK(cudaHostUnregister(ptr)); return 0; }

// ---- array-of-struct batch (identical in effect to N daqp_quadprog calls) -----------------------------------------------
// Problems are grouped by shape; the groups are dealt, largest estimated cost first, to a few LANES (sub-engines with
// their own streams, scratch and pinned staging) that run side by side: while one lane's group is being solved, another
// lane packs its next group into pinned memory or copies results out. A group is packed by several host threads.
namespace {
struct Lane {
    DAQPB200Handle* h = nullptr;
    char* pin = nullptr;
    size_t pin_bytes = 0;
};
std::mutex g_lane_mu;
std::map<int, std::vector<Lane>> g_lanes;
constexpr int N_LANES = 3;

int get_lanes(std::vector<Lane>** out) {
    std::lock_guard<std::mutex> lk(g_lane_mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { g_last_error = "daqp_b200: no CUDA device available (this library has no CPU path)"; return -1; }
    auto& v = g_lanes[dev];
    if (v.empty()) {
        v.resize(N_LANES);
        int rc = default_handle(&v[0].h); // lane 0 is the process-wide default engine (single calls reuse its buffers)
        if (rc) { v.clear(); return rc; }
        for (int i = 1; i < N_LANES; i++) {
            rc = daqp_b200_create(&v[i].h, dev);
            if (rc) { v.clear(); return rc; }
        }
    }
    *out = &v;
    return 0;
}

template <typename F> void parallel_for(size_t count, size_t min_per_thread, F fn) {
    const size_t hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nt = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(hw, 8), count / std::max<size_t>(1, min_per_thread)));
    if (nt <= 1) { fn(0, count); return; }
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++) th.emplace_back([=]() { fn(count * t / nt, count * (t + 1) / nt); });
    for (auto& t : th) t.join();
}

struct ShapeKey { int n, m, ms; bool has_f, has_s; };

template <typename T> struct Aos;
template <> struct Aos<double> { typedef DAQPProblem Problem; typedef DAQPResult Result; };
template <> struct Aos<float> { typedef DAQPProblemF32 Problem; typedef DAQPResultF32 Result; };

// one shape group on one lane: pack into the lane's pinned buffer, solve, scatter the results to the caller's structs
template <typename T>
int run_group(Lane& L, const ShapeKey& k, const int* ids, size_t G, typename Aos<T>::Problem* qps, typename Aos<T>::Result* res,
              DAQPSettings* settings) {
    const int n = k.n, m = k.m, ms = k.ms, mA = m - ms;
    const size_t nH = G * n * n, nf = k.has_f ? G * n : 0, nA = G * (size_t)mA * n, nb = G * m, nx = G * n;
    const size_t bytes = (nH + nf + nA + 2 * nb + nx + nb + 2 * G) * sizeof(T) + (k.has_s ? nb : 0) * sizeof(int) + 2 * G * sizeof(int) + 16 * 64;
    cudaSetDevice(L.h->device);
    if (L.pin_bytes < bytes) {
        if (L.pin) { cudaFreeHost(L.pin); L.pin = nullptr; L.pin_bytes = 0; }
        if (cudaHostAlloc((void**)&L.pin, bytes, cudaHostAllocDefault) != cudaSuccess) { g_last_error = "daqp_b200: cudaHostAlloc failed"; return -3; }
        L.pin_bytes = bytes;
    }
    char* q = L.pin;
    auto take = [&](size_t count, size_t size) { char* r = q; q += (count * size + 63) / 64 * 64; return r; };
    T* H = (T*)take(nH, sizeof(T)); T* f = (T*)take(nf, sizeof(T));
    T* A = (T*)take(nA, sizeof(T)); T* bu = (T*)take(nb, sizeof(T));
    T* bl = (T*)take(nb, sizeof(T)); int* se = (int*)take(k.has_s ? nb : 0, sizeof(int));
    T* x = (T*)take(nx, sizeof(T)); T* lam = (T*)take(nb, sizeof(T));
    T* fv = (T*)take(G, sizeof(T)); T* slack = (T*)take(G, sizeof(T));
    int* flag = (int*)take(G, sizeof(int)); int* it = (int*)take(G, sizeof(int));
    parallel_for(G, 256, [&](size_t g0, size_t g1) {
        for (size_t g = g0; g < g1; g++) {
            const auto& p = qps[ids[g]];
            if (p.H) memcpy(H + g * n * n, p.H, sizeof(T) * n * n);
            else { // H == NULL, f == NULL: the LDP min |x|^2 itself -- the identity Hessian gives the reference's iterates bit for bit
                T* Hg = H + g * n * n;
                memset(Hg, 0, sizeof(T) * n * n);
                for (int d = 0; d < n; d++) Hg[(size_t)d * n + d] = (T)1;
            }
            if (k.has_f) memcpy(f + g * n, p.f, sizeof(T) * n);
            if (mA > 0) memcpy(A + g * (size_t)mA * n, p.A, sizeof(T) * mA * n);
            if (m > 0) { memcpy(bu + g * m, p.bupper, sizeof(T) * m); memcpy(bl + g * m, p.blower, sizeof(T) * m); }
            if (k.has_s) memcpy(se + g * m, p.sense, sizeof(int) * m);
            fv[g] = res[ids[g]].fval; // the reference leaves fval untouched when there is no linear term
            slack[g] = 0;
        }
    });
    DAQPB200Diag dg{};
    if constexpr (sizeof(T) == sizeof(c_float)) dg.soft_slack = slack;
    DAQPB200Stats before{}, after{};
    daqp_b200_get_stats(L.h, &before, 0);
    int rc = solve_packed_impl<T>(L.h, (int)G, n, m, ms, H, k.has_f ? f : nullptr, A, bu, bl, k.has_s ? se : nullptr, settings, x, lam,
                                  fv, flag, it, sizeof(T) == sizeof(c_float) ? &dg : nullptr);
    if (rc) return rc;
    daqp_b200_get_stats(L.h, &after, 0);
    const T t_setup = (T)(1e-3 * (after.setup_ms - before.setup_ms)), t_solve = (T)(1e-3 * (after.solve_ms - before.solve_ms));
    parallel_for(G, 1024, [&](size_t g0, size_t g1) {
        for (size_t g = g0; g < g1; g++) {
            auto& r = res[ids[g]];
            r.exitflag = flag[g];
            r.setup_time = t_setup; r.solve_time = t_solve;
            r.iter = it[g];
            if (it[g] <= 0) continue; // setup failures leave x untouched (api.c:69-72): exit flags raised before the solve
            memcpy(r.x, x + g * n, sizeof(T) * n);
            if (r.lam && m > 0) memcpy(r.lam, lam + g * m, sizeof(T) * m);
            if (k.has_f) r.fval = fv[g];
            r.soft_slack = slack[g];
        }
    });
    return 0;
}

template <typename T>
int quadprog_batch_impl(int N, typename Aos<T>::Problem* qps, typename Aos<T>::Result* res, DAQPSettings* settings) {
    if (N <= 0) return 0;
    typedef std::tuple<int, int, int, bool, bool> Key;
    std::map<Key, std::vector<int>> groups;
    for (int i = 0; i < N; i++) {
        const auto& q = qps[i];
        res[i].nodes = 1; res[i].soft_slack = 0; res[i].solve_time = 0; res[i].setup_time = 0;
        // (H == NULL with a linear term is an LP: the reference's proximal-point driver, out of scope)
        const bool bad = (q.H == nullptr && q.f != nullptr) || q.nh > 1 || q.problem_type != 0 || q.n < 1 || q.m < q.ms || q.ms > q.n ||
                         (q.m > q.ms && q.A == nullptr) || (q.m > 0 && (q.bupper == nullptr || q.blower == nullptr));
        if (bad) { res[i].exitflag = DAQP_EXIT_UNSUPPORTED; res[i].iter = 0; continue; }
        if constexpr (sizeof(T) == sizeof(c_float)) { // binary constraints: the tree search, its node relaxations batched
            bool binary = false;
            if (q.sense) for (int r = 0; r < q.m && !binary; r++) binary = (q.sense[r] & DAQP_BINARY) != 0;
            if (binary) {
                int rcb = daqp_b200_bnb(nullptr, &q, settings, &res[i], 0);
                if (rcb) return rcb;
                continue;
            }
        }
        groups[std::make_tuple(q.n, q.m, q.ms, q.f != nullptr, q.sense != nullptr)].push_back(i);
    }
    if (groups.empty()) return 0;
    std::vector<Lane>* lanes = nullptr;
    int rc = get_lanes(&lanes);
    if (rc) return rc;
    // largest estimated cost first (n^2 m per problem); a lane takes the next piece as soon as it is free. A
    // large group is cut into pieces so that a homogeneous batch -- one group -- still fills all lanes: while one lane's
    // piece is being copied and solved, the others pack theirs out of the caller's pageable structs.
    struct Piece { double cost; ShapeKey key; const int* ids; size_t count; };
    std::vector<Piece> order;
    for (auto& kv : groups) {
        const double n = std::get<0>(kv.first), m = std::get<1>(kv.first);
        const ShapeKey key{std::get<0>(kv.first), std::get<1>(kv.first), std::get<2>(kv.first), std::get<3>(kv.first), std::get<4>(kv.first)};
        const size_t G = kv.second.size();
        const size_t per_problem = (size_t)(n * n + n + std::max(m - std::get<2>(kv.first), 0.0) * n + 2 * m) * sizeof(T);
        // pieces of ~128 MB of input, at least 1024 problems; small groups stay whole
        size_t piece_mb = 128;
        if (const char* penv = getenv("DAQP_B200_AOS_PIECE_MB")) piece_mb = (size_t)std::max(1, atoi(penv)); // tuning knob
        const size_t piece = std::max<size_t>(1024, (piece_mb << 20) / std::max<size_t>(per_problem, 1));
        const size_t np = G > 2 * piece ? (G + piece - 1) / piece : 1;
        for (size_t q = 0; q < np; q++) {
            const size_t lo = G * q / np, hi = G * (q + 1) / np;
            order.push_back({n * n * std::max(m, 1.0) * (double)(hi - lo), key, kv.second.data() + lo, hi - lo});
        }
    }
    std::stable_sort(order.begin(), order.end(), [](const Piece& a, const Piece& b) { return a.cost > b.cost; });
    if (order.size() == 1) return run_group<T>((*lanes)[0], order[0].key, order[0].ids, order[0].count, qps, res, settings);
    const int nl = (int)std::min<size_t>(lanes->size(), order.size());
    std::vector<int> rcs((size_t)nl, 0);
    std::vector<std::string> errs((size_t)nl);
    std::vector<std::thread> th;
    std::atomic<size_t> next{0};
    for (int l = 0; l < nl; l++)
        th.emplace_back([&, l]() {
            for (;;) {
                const size_t gi = next.fetch_add(1);
                if (gi >= order.size() || rcs[l]) break;
                rcs[l] = run_group<T>((*lanes)[l], order[gi].key, order[gi].ids, order[gi].count, qps, res, settings);
                if (rcs[l]) errs[l] = g_last_error;
            }
        });
    for (auto& t : th) t.join();
    for (int l = 0; l < nl; l++)
        if (rcs[l]) { g_last_error = errs[l]; return rcs[l]; }
    return 0;
}
} // namespace

extern "C" int daqp_quadprog_batch(int N, DAQPProblem* qps, DAQPResult* res, DAQPSettings* settings) {
    return quadprog_batch_impl<double>(N, qps, res, settings);
}
// the same for the single-precision ABI (the reference built with -DDAQP_SINGLE_PRECISION, include/types.h:8-12)
extern "C" int daqp_quadprog_batch_f32(int N, DAQPProblemF32* qps, DAQPResultF32* res, DAQPSettings* settings) {
    return quadprog_batch_impl<float>(N, qps, res, settings);
}

extern "C" void daqp_quadprog(DAQPResult* res, DAQPProblem* qp, DAQPSettings* settings) {
    int rc = daqp_quadprog_batch(1, qp, res, settings);
    if (rc) {
        fprintf(stderr, "%s\n", g_last_error.c_str());
        res->exitflag = DAQP_EXIT_UNSUPPORTED;
    }
}
