// daqp_b200/csrc/team_launch.cu -- instantiations of the solve kernel's team mode (one CTA of four warps per problem,
// n > 64; see ldp_kernel.cuh). Its own translation unit so that it compiles next to daqp_b200.cu.
#include "ldp_kernel.cuh"

using namespace dq;

template <int NV>
static cudaError_t launch_team(const LdpArgs<double>& a, int grid, size_t smem, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(ldp_solve_kernel<double, NV, false, TEAM_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    ldp_solve_kernel<double, NV, false, TEAM_WARPS><<<grid, 32 * TEAM_WARPS, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t daqp_b200_launch_solve_team(const LdpArgs<double>& a, int nv, int grid, size_t smem, cudaStream_t s) {
    switch (nv) {
        case 3: return launch_team<3>(a, grid, smem, s);
        case 4: return launch_team<4>(a, grid, smem, s);
        default: return cudaErrorNotSupported;
    }
}
