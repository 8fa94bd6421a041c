// solve kernel instantiation: generic_team (RuntimeDims, double), a team of UB_TEAM_WARPS warps per instance
#include "ub_launch.cuh"
namespace ub {
cudaError_t launch_generic_team_f64(const DevProblem<double>& Ph, const DevProblem<double>* Pg, const Layout& L, const BatchArgs<double>& A,
                                  int tpc, int grid, size_t smem, cudaStream_t stream) {
    return launch_solve_kernel<double, RuntimeDims, UB_TEAM_WARPS>(Ph, Pg, L, A, tpc, grid, smem, stream);
}
}
