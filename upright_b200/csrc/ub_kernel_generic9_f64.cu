// solve kernel instantiation: generic9 (RuntimeDims<9>), F = double
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(generic9, double, f64, RuntimeDims<9>)
}
