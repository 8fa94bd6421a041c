// solve kernel instantiation: generic6 (RuntimeDims<6>), F = double
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(generic6, double, f64, RuntimeDims<6>)
}
