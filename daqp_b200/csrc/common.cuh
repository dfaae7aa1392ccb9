// daqp_b200/csrc/common.cuh -- shared device helpers for the batched dual active-set QP engine (sm_100a).
//
// Execution model: ONE WARP PER PROBLEM for the whole solve. All control flow is warp-uniform (every lane
// carries the same n_active / sing_ind / reuse_ind in registers); vectors and the LDL' factor live in shared
// memory private to the warp; the constraint matrix is streamed from HBM/L2 with 128-bit loads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dq {

constexpr int EMPTY_IND = -1;
constexpr unsigned FULL = 0xffffffffu;

// sense bits: reference include/constants.h:64-96
constexpr int B_ACTIVE = 1, B_LOWER = 2, B_IMMUTABLE = 4, B_SOFT = 8, B_BINARY = 16;

// exit flags: reference include/constants.h:42-51
constexpr int EXIT_SOFT_OPTIMAL = 2, EXIT_OPTIMAL = 1, EXIT_INFEASIBLE = -1, EXIT_CYCLE = -2, EXIT_UNBOUNDED = -3,
              EXIT_ITERLIMIT = -4, EXIT_NONCONVEX = -5, EXIT_OVERDETERMINED_INITIAL = -6, EXIT_TIMELIMIT = -7,
              EXIT_UNSUPPORTED = -8;

// setup kernel -> solve kernel hand-over (per problem)
constexpr int SETUP_SOLVE = 0;          // LDP formed, run the active-set loop
constexpr int SETUP_SOLVE_ACTIVATE = 1; // same, but sense carries pre-activated rows (warm start / equalities)
constexpr int SETUP_UNCONSTRAINED = 2;  // unconstrained optimum is feasible: x already final (utils.c:679-683)
// negative values are final exit flags raised by the setup (api.c:69-72)

template <typename T>
struct DevSettings { // the DAQPSettings fields the hot path reads (include/types.h:52-74)
    T primal_tol, dual_tol, zero_tol, pivot_tol, progress_tol, fval_bound, rho_soft, sing_tol, refactor_tol, eps_prox;
    int cycle_tol, iter_limit;
    long long time_limit_ns; // settings->time_limit in ns of %globaltimer; 0 = no limit (daqp.c:95-103)
};

__device__ __forceinline__ long long global_timer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <typename T> struct VecOf;
template <> struct VecOf<double> { using type = double2; static constexpr int N = 2; };
template <> struct VecOf<float> { using type = float4; static constexpr int N = 4; };

// 128-bit read-only global load
template <typename T> __device__ __forceinline__ void ldg_vec(const T* p, T (&out)[VecOf<T>::N]);
template <> __device__ __forceinline__ void ldg_vec<double>(const double* p, double (&out)[2]) {
    double2 v = __ldg(reinterpret_cast<const double2*>(p));
    out[0] = v.x; out[1] = v.y;
}
template <> __device__ __forceinline__ void ldg_vec<float>(const float* p, float (&out)[4]) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
// L2 eviction policies (createpolicy + ld.global.nc.L2::cache_hint): the streamed column-major matrix is marked
// evict_first so that it does not push the small, re-read set of active rows (marked evict_last) out of L2.
// The two policies are the fixed encodings createpolicy.fractional.L2::evict_{first,last}.b64 (fraction 1.0)
// produces. Spelling them as constants lets the compiler keep them in uniform registers; taking them from the
// createpolicy instruction put them in vector registers and cost two R2UR moves in front of every load (6.6 % of all
// executed instructions in the v7 profile).
__device__ __forceinline__ constexpr uint64_t policy_evict_first() { return 0x12F0000000000000ull; }
__device__ __forceinline__ constexpr uint64_t policy_evict_last() { return 0x14F0000000000000ull; }
template <typename T> __device__ __forceinline__ void ldg_vec_hint(const T* p, T (&out)[VecOf<T>::N], uint64_t pol);
template <> __device__ __forceinline__ void ldg_vec_hint<double>(const double* p, double (&out)[2], uint64_t pol) {
    asm("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(out[0]), "=d"(out[1]) : "l"(p), "l"(pol));
}
template <> __device__ __forceinline__ void ldg_vec_hint<float>(const float* p, float (&out)[4], uint64_t pol) {
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(out[0]), "=f"(out[1]), "=f"(out[2]), "=f"(out[3]) : "l"(p), "l"(pol));
}
// One-instruction bulk prefetch of a contiguous global range into L2 (TMA bulk prefetch, no registers, no smem).
// bytes must be a multiple of 16 and p 16-byte aligned. Issued by one lane.
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// Same loads, but "volatile": the compiler keeps them in program order, which is what pins the depth of a
// hand-written software pipeline (a slot is refilled right after it has been consumed).
template <typename T> __device__ __forceinline__ void ldg_vec_hint_ordered(const T* p, T (&out)[VecOf<T>::N], uint64_t pol);
template <> __device__ __forceinline__ void ldg_vec_hint_ordered<double>(const double* p, double (&out)[2], uint64_t pol) {
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(out[0]), "=d"(out[1]) : "l"(p), "l"(pol));
}
template <> __device__ __forceinline__ void ldg_vec_hint_ordered<float>(const float* p, float (&out)[4], uint64_t pol) {
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(out[0]), "=f"(out[1]), "=f"(out[2]), "=f"(out[3]) : "l"(p), "l"(pol));
}
template <typename T> __device__ __forceinline__ void stg_vec(T* p, const T (&in)[VecOf<T>::N]);
template <> __device__ __forceinline__ void stg_vec<double>(double* p, const double (&in)[2]) {
    *reinterpret_cast<double2*>(p) = make_double2(in[0], in[1]);
}
template <> __device__ __forceinline__ void stg_vec<float>(float* p, const float (&in)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
}

// value select that the optimiser cannot turn into an indexed access of a register array (which would send the
// array to local memory): opaque PTX selp
__device__ __forceinline__ double sel(bool c, double a, double b) {
    double r;
    asm("{\n .reg .pred q;\n setp.ne.b32 q, %3, 0;\n selp.f64 %0, %1, %2, q;\n}" : "=d"(r) : "d"(a), "d"(b), "r"((int)c));
    return r;
}
__device__ __forceinline__ float sel(bool c, float a, float b) {
    float r;
    asm("{\n .reg .pred q;\n setp.ne.b32 q, %3, 0;\n selp.f32 %0, %1, %2, q;\n}" : "=f"(r) : "f"(a), "f"(b), "r"((int)c));
    return r;
}

// ---- asynchronous global -> shared copies (cp.async / LDGSTS), 16 bytes per lane, lane-private ring slots ----------
// Every lane copies, and later reads back, only its own 16-byte chunk of a matrix row / column, so the ring needs no
// barrier at all: cp.async.wait_group is a per-thread wait. The data never occupies registers while in flight, which
// is what lets a rolled, 30-instruction loop keep 8 columns (or rows) in flight.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// src_bytes = 16 copies, src_bytes = 0 writes zeros without touching global memory
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src, int src_bytes, uint64_t pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(dst), "l"(src), "r"(src_bytes), "l"(pol) : "memory");
}
// plain 16-byte copy (no cache-hint operand: saves the uniform-register set-up in front of every group)
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// the same, predicated INSIDE the instruction: a row's copies stay in one basic block (a branch around each copy costs
// a reconvergence pair and makes ptxas re-emit its three filler instructions in front of every LDGSTS)
__device__ __forceinline__ void cp_async16_if(bool pred, unsigned dst, const void* src) {
    asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q cp.async.cg.shared.global [%0], [%1], 16;\n}" ::"r"(dst), "l"(src), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <typename T> __device__ __forceinline__ void lds_vec(unsigned addr, T (&out)[VecOf<T>::N]);
template <> __device__ __forceinline__ void lds_vec<double>(unsigned addr, double (&out)[2]) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(out[0]), "=d"(out[1]) : "r"(addr));
}
template <> __device__ __forceinline__ void lds_vec<float>(unsigned addr, float (&out)[4]) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(out[0]), "=f"(out[1]), "=f"(out[2]), "=f"(out[3]) : "r"(addr));
}

// x = (i < j) ? x - l * y : x   as ONE compare plus ONE predicated FMA (the C++ form costs a compare, an FMA and a
// two-instruction select per double)
__device__ __forceinline__ void fms_if_lt(double& x, double l, double y, int i, int j) {
    asm("{\n .reg .pred q;\n setp.lt.s32 q, %3, %4;\n @q fma.rn.f64 %0, %1, %2, %0;\n}" : "+d"(x) : "d"(-l), "d"(y), "r"(i), "r"(j));
}
__device__ __forceinline__ void fms_if_lt(float& x, float l, float y, int i, int j) {
    asm("{\n .reg .pred q;\n setp.lt.s32 q, %3, %4;\n @q fma.rn.f32 %0, %1, %2, %0;\n}" : "+f"(x) : "f"(-l), "f"(y), "r"(i), "r"(j));
}

// The reductions and the fp64 division are deliberately NOT inlined in the solve kernel: each expands to 20-30
// instructions and is used at a dozen sites; one shared copy keeps the hot code inside the instruction cache.
template <typename T> __device__ __noinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
static __device__ __noinline__ double fdiv(double a, double b) { return a / b; }
__device__ __forceinline__ float fdiv(float a, float b) { return a / b; }
// correctly rounded reciprocal: bit-identical to 1/x (IEEE division is correctly rounded too) at a third of the
// instructions of the general division, and short enough to inline
__device__ __forceinline__ double frcp(double x) { return __drcp_rn(x); }
__device__ __forceinline__ float frcp(float x) { return __frcp_rn(x); }

// Sum B per-lane values across the warp with a transposed butterfly: B/2 + B/4 + ... + 1 exchanges replace B full
// 5-step reductions. Returns the warp total of value number multi_index<B>(lane); the 32/B lanes that share the
// top log2(B) lane bits all return the same total.
template <int B, typename T> __device__ __forceinline__ T warp_sum_multi(T (&v)[B], int lane) {
    static_assert(B == 1 || B == 2 || B == 4 || B == 8 || B == 16, "B must be a power of two <= 16");
    int bit = 16;
#pragma unroll
    for (int half = B / 2; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = lane & bit;
#pragma unroll
        for (int i = 0; i < half; i++) {
            const T send = up ? v[i] : v[i + half];
            const T keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, bit);
        }
    }
    T r = v[0];
#pragma unroll
    for (int o = 16 / B; o >= 1; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
    return r;
}
// which of the B values a lane ends up holding after warp_sum_multi<B>
template <int B> __device__ __forceinline__ int multi_index(int lane) {
    int idx = 0, bit = 16;
#pragma unroll
    for (int half = B / 2; half >= 1; half >>= 1, bit >>= 1) idx += (lane & bit) ? half : 0;
    return idx;
}

// Lexicographic (value, key) minimum over the warp. Reproduces a sequential "strict <, first index wins" scan
// (reference auxiliary.c:112,118,139,145,293): the smallest value wins, ties go to the smallest key.
// Lanes without a candidate pass key = INT_MAX. NaN never enters because candidates are admitted with '<'.
template <typename T> __device__ __noinline__ void warp_argmin(T& val, int& key) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T ov = __shfl_xor_sync(FULL, val, o);
        int ok = __shfl_xor_sync(FULL, key, o);
        if (ov < val || (ov == val && ok < key)) { val = ov; key = ok; }
    }
}

template <typename T> __device__ __forceinline__ T rsqrt_exact(T x);
template <> __device__ __forceinline__ double rsqrt_exact<double>(double x) { return 1.0 / sqrt(x); }
template <> __device__ __forceinline__ float rsqrt_exact<float>(float x) { return 1.0f / sqrtf(x); }

__host__ __device__ __forceinline__ int round_up(int x, int a) { return (x + a - 1) / a * a; }

// Largest dynamic shared memory a kernel of the current device may opt in to. Every launcher raises its kernel's limit to
// THIS value, never to the launch's own size: the attribute is per kernel function, and two host threads (lanes of the
// array-of-struct batch) launching the same instantiation with different sizes would otherwise lower it under each other.
inline int max_optin_smem() {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return v;
}

// packed strict-lower storage of the unit-lower factor: row i holds L[i][0..i-1] at offset i(i-1)/2
__host__ __device__ __forceinline__ int loff(int i) { return (i * (i - 1)) >> 1; }
// packed upper-triangular (by rows) offset such that (R + roff(i,n))[j] is element (i,j), j >= i
// (reference DAQP_R_OFFSET, include/constants.h:39)
__host__ __device__ __forceinline__ int roff(int i, int n) { return ((2 * n - i - 1) * i) / 2; }

} // namespace dq
