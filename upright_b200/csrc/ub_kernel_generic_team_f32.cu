// solve kernel instantiation: generic_team (RuntimeDims, float), a team of UB_TEAM_WARPS warps per instance
#include "ub_launch.cuh"
namespace ub {
cudaError_t launch_generic_team_f32(const DevProblem<float>& Ph, const DevProblem<float>* Pg, const Layout& L, const BatchArgs<float>& A,
                                  int tpc, int grid, size_t smem, cudaStream_t stream) {
    return launch_solve_kernel<float, RuntimeDims, UB_TEAM_WARPS>(Ph, Pg, L, A, tpc, grid, smem, stream);
}
}
