// daqp_b200/csrc/setup2_launch.cu -- instantiations of the split QP -> LDP transform (setup2_kernel.cuh): its own
// translation unit so that it compiles next to daqp_b200.cu.
#include <algorithm>
#include "setup2_kernel.cuh"

using namespace dq;

template <int TW>
static cudaError_t launch_factor(const SetupArgs<double>& a, int num_sms, size_t smem_optin, cudaStream_t s) {
    const size_t per = factor_smem_per_team<double>(a.n);
    int grid, block;
    size_t smem;
    if (TW == 1) {
        const int w = (int)std::min<size_t>(FACTOR_MAX_WARPS, smem_optin / per);
        if (w < 1) return cudaErrorInvalidConfiguration;
        block = 32 * w; smem = per * w; grid = std::min(num_sms, (a.P + w - 1) / w);
    } else {
        const int ctas = (int)std::min<size_t>(4, (smem_optin + 1024) / (per + 1024));
        if (ctas < 1) return cudaErrorInvalidConfiguration;
        block = 32 * TW; smem = per; grid = std::min(num_sms * ctas, a.P);
    }
    cudaError_t e = cudaFuncSetAttribute(qp_factor_kernel<double, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin_smem());
    if (e != cudaSuccess) return e;
    qp_factor_kernel<double, TW><<<grid, block, smem, s>>>(a);
    return cudaGetLastError();
}

template <int NCT>
static cudaError_t launch_product(const SetupArgs<double>& a, int num_sms, size_t smem_optin, cudaStream_t s) {
    const size_t smem = product_smem(a.n);
    const int ctas = (int)std::min<size_t>(NCT <= 9 ? 4 : 3, (smem_optin + 1024) / (smem + 1024));
    if (ctas < 1) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(qp_product_kernel<NCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin_smem());
    if (e != cudaSuccess) return e;
    qp_product_kernel<NCT><<<std::min(num_sms * ctas, a.P), 128, smem, s>>>(a);
    return cudaGetLastError();
}

// fp64, n <= 127. work_counter[0] and [1] must be zero.
cudaError_t daqp_b200_launch_setup_split(const SetupArgs<double>& a, int num_sms, size_t smem_optin, cudaStream_t s) {
    cudaError_t e = a.n > 64 ? launch_factor<4>(a, num_sms, smem_optin, s) : launch_factor<1>(a, num_sms, smem_optin, s);
    if (e != cudaSuccess) return e;
    const int nct = (a.n + 8) >> 3;
    if (nct <= 2) return launch_product<2>(a, num_sms, smem_optin, s);
    if (nct <= 4) return launch_product<4>(a, num_sms, smem_optin, s);
    if (nct <= 7) return launch_product<7>(a, num_sms, smem_optin, s);
    if (nct <= 9) return launch_product<9>(a, num_sms, smem_optin, s);
    if (nct <= 12) return launch_product<12>(a, num_sms, smem_optin, s);
    if (nct <= 16) return launch_product<16>(a, num_sms, smem_optin, s);
    return cudaErrorNotSupported;
}
