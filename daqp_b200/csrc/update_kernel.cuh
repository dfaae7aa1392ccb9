// daqp_b200/csrc/update_kernel.cuh -- persistent-workspace path: new linear term and bounds on a kept LDP.
//
// Mirrors the reference's daqp_update_ldp(DAQP_UPDATE_v + DAQP_UPDATE_d) (src/utils.c:58-221) for a batch whose H and A
// have not changed since daqp_b200_workspace_setup:
//   check_bounds (utils.c:546-567) -> v = R^-T f with the column scaling of the normalised simple-bound rows
//   (utils.c:474-497) -> d = b * scaling + M v (utils.c:499-544).
// The Cholesky factor, R^-1, M (both layouts + the fp32 screening copy) and the row scaling stay on the device;
// the kernel streams the column-major matrix once (lane = row, 64-bit coalesced) and is bound by that read.
#pragma once
#include "common.cuh"

namespace dq {

template <typename T>
struct UpdateArgs {
    int P, n, m, ms, ldm;
    int grp;                           // shared workspace: Rinv / Mt / scaling / sense_static are those of set p / grp (0, 1: own)
    const T *f, *bupper, *blower;      // new data: [P][n] (nullptr keeps v), [P][m], [P][m]
    const T *Rinv, *Mt, *scaling;      // kept LDP: packed R^-1 (simple-bound rows normalised), Mt [P][n][ldm], scaling [P][ldm]
    const unsigned char* sense_static; // [P][ldm] user sense bits, zero rows IMMUTABLE, no equality marks
    T *v, *dupper, *dlower;            // [P][n], [P][ldm], [P][ldm]
    unsigned char* sense;              // [P][ldm] sense bits the next solve starts from
    int *setup_flag, *exitflag, *iter; // [P]
    DevSettings<T> st;
};

template <typename T>
__global__ void __launch_bounds__(512) ldp_update_kernel(const UpdateArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int n = a.n, m = a.m, ms = a.ms, ldm = a.ldm;
    T* fs = reinterpret_cast<T*>(smem_raw) + (size_t)wib * 2 * n; // f with the column scaling applied
    T* vv = fs + n;
    for (int p = blockIdx.x * wpb + wib; p < a.P; p += gridDim.x * wpb) {
        // a Hessian the setup rejected (non-convex, or singular: proximal driver needed) stays rejected: these two exit
        // flags are only ever raised by the setup (utils.c:356-377), new f / b cannot cure them
        const int prev = a.exitflag[p];
        if (prev == EXIT_NONCONVEX || prev == EXIT_UNSUPPORTED) continue;
        const int pm = a.grp > 1 ? p / a.grp : p;
        const T* sc = a.scaling + (size_t)pm * ldm;
        const T* bu = a.bupper + (size_t)p * m;
        const T* bl = a.blower + (size_t)p * m;
        const T* Ri = a.Rinv + (size_t)pm * n * (n + 1) / 2;
        T* vg = a.v + (size_t)p * n;
        // ---- v = R^-T f (utils.c:474-497): column j < ms of the normalised R^-1 carries 1 / scaling[j]
        if (a.f) {
            const T* f = a.f + (size_t)p * n;
            for (int i = lane; i < n; i += 32) fs[i] = i < ms ? f[i] / sc[i] : f[i];
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                T s = Ri[roff(i, n) + i] * fs[i];
                for (int j = i - 1; j >= 0; j--) s += Ri[roff(j, n) + i] * fs[j];
                vv[i] = s;
                vg[i] = s;
            }
        } else {
            for (int i = lane; i < n; i += 32) vv[i] = vg[i];
        }
        __syncwarp();
        // ---- check_bounds on the new bounds (utils.c:546-567) and d = b * scaling + M v (utils.c:499-544)
        int bad = 0, any_active = 0;
        const T* Mt = a.Mt + (size_t)pm * n * ldm;
        for (int r = lane; r < ldm; r += 32) {
            int s = 0;
            T du = 0, dl = 0;
            if (r < m) {
                s = a.sense_static[(size_t)pm * ldm + r];
                const T u_ = bu[r], l_ = bl[r];
                if (!(s & B_IMMUTABLE)) {
                    const T diff = u_ - l_;
                    if (diff < -a.st.primal_tol) bad = 1;
                    else if (diff < a.st.zero_tol && !(s & B_SOFT)) s |= B_ACTIVE + B_IMMUTABLE;
                }
                if (s & B_ACTIVE) any_active = 1;
                T sum = 0;
                for (int j = 0; j < n; j++) sum += Mt[(size_t)j * ldm + r] * vv[j]; // rows < ms: zeros left of the diagonal
                const T scl = sc[r];
                du = u_ * scl + sum;
                dl = l_ * scl + sum;
            }
            a.sense[(size_t)p * ldm + r] = (unsigned char)s;
            a.dupper[(size_t)p * ldm + r] = du;
            a.dlower[(size_t)p * ldm + r] = dl;
        }
        bad = __any_sync(FULL, bad);
        any_active = __any_sync(FULL, any_active);
        if (lane == 0) {
            if (bad) { a.setup_flag[p] = EXIT_INFEASIBLE; a.exitflag[p] = EXIT_INFEASIBLE; a.iter[p] = 0; }
            else a.setup_flag[p] = any_active ? SETUP_SOLVE_ACTIVATE : SETUP_SOLVE;
        }
        __syncwarp();
    }
}

} // namespace dq
