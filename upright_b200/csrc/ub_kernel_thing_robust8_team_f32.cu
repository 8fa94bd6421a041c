// solve kernel instantiation: thing_robust8_team (StaticDims<9, 1, 32, 8>, float), a team of UB_TEAM_WARPS warps per instance
#include "ub_launch.cuh"
namespace ub {
cudaError_t launch_thing_robust8_team_f32(const DevProblem<float>& Ph, const DevProblem<float>* Pg, const Layout& L, const BatchArgs<float>& A,
                                int tpc, int grid, size_t smem, cudaStream_t stream) {
    return launch_solve_kernel<float, StaticDims<9, 1, 32, 8>, UB_TEAM_WARPS>(Ph, Pg, L, A, tpc, grid, smem, stream);
}
}
