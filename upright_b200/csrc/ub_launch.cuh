// Per-instantiation launchers of ub::solve_batch_kernel.  Every (precision, dimension set) lives in its own
// translation unit (ub_kernel_*.cu, generated from UB_DEFINE_LAUNCHER) so that the library builds in parallel.
#pragma once
#include <cuda_runtime.h>

#include "ub_solver.cuh"

namespace ub {

template <typename T>
using LaunchFn = cudaError_t (*)(const DevProblem<T>& host_copy, const DevProblem<T>* device_copy, const Layout&,
                                 const BatchArgs<T>&, int warps_per_cta, int grid, size_t smem, cudaStream_t);

template <typename T, typename D, int TW = 1>
cudaError_t launch_solve_kernel(const DevProblem<T>& Ph, const DevProblem<T>* Pg, const Layout& L, const BatchArgs<T>& A, int wpc,
                                int grid, size_t smem, cudaStream_t stream) {
    auto kernel = solve_batch_kernel<T, D, TW>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    kernel<<<grid, wpc * 32 * TW, smem, stream>>>(Ph, Pg, L, A, wpc);   // wpc = instance teams per CTA
    return cudaGetLastError();
}

#define UB_DECLARE_LAUNCHER(NAME)                                                                                       \
    cudaError_t launch_##NAME##_f32(const DevProblem<float>&, const DevProblem<float>*, const Layout&,                  \
                                    const BatchArgs<float>&, int, int, size_t, cudaStream_t);                           \
    cudaError_t launch_##NAME##_f64(const DevProblem<double>&, const DevProblem<double>*, const Layout&,                \
                                    const BatchArgs<double>&, int, int, size_t, cudaStream_t);
#define UB_DEFINE_LAUNCHER(NAME, T, SUFFIX, ...)                                                                        \
    cudaError_t launch_##NAME##_##SUFFIX(const DevProblem<T>& Ph, const DevProblem<T>* Pg, const Layout& L,             \
                                         const BatchArgs<T>& A, int wpc, int grid, size_t smem, cudaStream_t stream) {  \
        return launch_solve_kernel<T, __VA_ARGS__>(Ph, Pg, L, A, wpc, grid, smem, stream);                             \
    }

UB_DECLARE_LAUNCHER(generic)
UB_DECLARE_LAUNCHER(thing_1obj)    // cfg2: nq 9, nf 1, nc 4, nb 1
UB_DECLARE_LAUNCHER(thing_obs12)   // cfg4: cfg2 dims + 12 sphere pairs
UB_DECLARE_LAUNCHER(ur10_1obj)     // cfg1: nq 6, nf 1, nc 4, nb 1
UB_DECLARE_LAUNCHER(thing_arch)    // cfg3: nq 9, nf 3, nc 16, nb 3
UB_DECLARE_LAUNCHER(thing_robust8) // cfg5: nq 9, nf 1, nc 32, nb 8
// large stage matrices (cfg3, cfg5): a team of UB_TEAM_WARPS warps per instance
#define UB_TEAM_WARPS 4
UB_DECLARE_LAUNCHER(thing_arch_team)
UB_DECLARE_LAUNCHER(thing_robust8_team)
UB_DECLARE_LAUNCHER(generic_team)   // run-time dimensions, large stage matrices

}  // namespace ub
