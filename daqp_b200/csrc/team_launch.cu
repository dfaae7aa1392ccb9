// daqp_b200/csrc/team_launch.cu -- instantiations of the solve kernel's team mode (one CTA of four warps per problem,
// n > 64; see ldp_kernel.cuh). Its own translation unit so that it compiles next to daqp_b200.cu.
#include "ldp_kernel.cuh"

using namespace dq;

template <int NV, int TW>
static cudaError_t launch_team(const LdpArgs<double>& a, int grid, size_t smem, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(ldp_solve_kernel<double, NV, false, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin_smem());
    if (e != cudaSuccess) return e;
    ldp_solve_kernel<double, NV, false, TW><<<grid, 32 * TW, smem, s>>>(a);
    return cudaGetLastError();
}

// a.team = warps per problem (2: factor rows <= 64, 4: <= 128)
cudaError_t daqp_b200_launch_solve_team(const LdpArgs<double>& a, int nv, int grid, size_t smem, cudaStream_t s) {
    if (a.team == 2 && nv == 2) return launch_team<2, 2>(a, grid, smem, s);
    if (a.team == 4 && nv == 3) return launch_team<3, 4>(a, grid, smem, s);
    if (a.team == 4 && nv == 4) return launch_team<4, 4>(a, grid, smem, s);
    return cudaErrorNotSupported;
}

// one warp per problem, streaming phases staged through registers (no shared-memory arena: more problems per SM)
template <int NV>
static cudaError_t launch_regs(const LdpArgs<double>& a, int grid, int block, size_t smem, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(ldp_solve_kernel<double, NV, false, TW_REGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin_smem());
    if (e != cudaSuccess) return e;
    ldp_solve_kernel<double, NV, false, TW_REGS><<<grid, block, smem, s>>>(a);
    return cudaGetLastError();
}
cudaError_t daqp_b200_launch_solve_regs(const LdpArgs<double>& a, int nv, int grid, int block, size_t smem, cudaStream_t s) {
    if (nv == 1) return launch_regs<1>(a, grid, block, smem, s);
    if (nv == 2) return launch_regs<2>(a, grid, block, smem, s);
    return cudaErrorNotSupported;
}
