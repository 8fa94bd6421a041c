// Per-instantiation launchers of ub::solve_batch_kernel.  Every (precision, dimension set) lives in its own
// translation unit (ub_kernel_*.cu, generated from UB_DEFINE_LAUNCHER) so that the library builds in parallel.
#pragma once
#include <cuda_runtime.h>

#include "ub_solver.cuh"

namespace ub {

// F = float: the product kernels (Riccati in fp32, iterate / residuals / linearisation / force block in fp64);
// F = double: the fp64 validation kernels
template <typename F>
using LaunchFn = cudaError_t (*)(const DevProblem<F>& host_copy, const DevProblem<F>* device_copy,
                                 const DevProblem<double>* device_copy_f64, const Layout&, const BatchArgs<F>&,
                                 int warps_per_cta, int grid, size_t smem, cudaStream_t);

template <typename F, typename D>
cudaError_t launch_solve_kernel(const DevProblem<F>& Ph, const DevProblem<F>* Pg, const DevProblem<double>* Pgr, const Layout& L,
                                const BatchArgs<F>& A, int wpc, int grid, size_t smem, cudaStream_t stream) {
    auto kernel = solve_batch_kernel<F, D>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    kernel<<<grid, wpc * 32, smem, stream>>>(Ph, Pg, Pgr, L, A, wpc);   // wpc = instances (warps) per CTA
    return cudaGetLastError();
}

#define UB_DECLARE_LAUNCHER(NAME)                                                                                       \
    cudaError_t launch_##NAME##_f32(const DevProblem<float>&, const DevProblem<float>*, const DevProblem<double>*,      \
                                    const Layout&, const BatchArgs<float>&, int, int, size_t, cudaStream_t);           \
    cudaError_t launch_##NAME##_f64(const DevProblem<double>&, const DevProblem<double>*, const DevProblem<double>*,    \
                                    const Layout&, const BatchArgs<double>&, int, int, size_t, cudaStream_t);
#define UB_DEFINE_LAUNCHER(NAME, T, SUFFIX, ...)                                                                        \
    cudaError_t launch_##NAME##_##SUFFIX(const DevProblem<T>& Ph, const DevProblem<T>* Pg, const DevProblem<double>* Pgr, \
                                         const Layout& L, const BatchArgs<T>& A, int wpc, int grid, size_t smem,       \
                                         cudaStream_t stream) {                                                        \
        return launch_solve_kernel<T, __VA_ARGS__>(Ph, Pg, Pgr, L, A, wpc, grid, smem, stream);                        \
    }

// the BASELINE shapes, dimensions compile-time
#define UB_DIMS_THING_1OBJ StaticDims<9, 1, 4, 1>                      /* cfg2: nq 9, nf 1, nc 4, nb 1 */
#define UB_DIMS_THING_OBS12 StaticDims<9, 1, 4, 1, 12>                 /* cfg4: cfg2 dims + 12 sphere pairs */
#define UB_DIMS_UR10_1OBJ StaticDims<6, 1, 4, 1>                       /* cfg1: nq 6, nf 1, nc 4, nb 1 */
#define UB_DIMS_THING_ARCH StaticDims<9, 3, 16, 3, 0, 20, true>        /* cfg3: nq 9, nf 3, nc 16, nb 3, bodies share contacts */
#define UB_DIMS_THING_ROBUST8 StaticDims<9, 1, 32, 8>                  /* cfg5: nq 9, nf 1, nc 32, nb 8 */
UB_DECLARE_LAUNCHER(thing_1obj)
UB_DECLARE_LAUNCHER(thing_obs12)
UB_DECLARE_LAUNCHER(ur10_1obj)
UB_DECLARE_LAUNCHER(thing_arch)
UB_DECLARE_LAUNCHER(thing_robust8)
// everything else: robot compile-time (the Riccati recursion depends on nq alone), the rest at run time
UB_DECLARE_LAUNCHER(generic9)
UB_DECLARE_LAUNCHER(generic6)

}  // namespace ub
