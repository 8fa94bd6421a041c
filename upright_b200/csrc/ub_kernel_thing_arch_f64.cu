// solve kernel instantiation: thing_arch (UB_DIMS_THING_ARCH), F = double
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(thing_arch, double, f64, UB_DIMS_THING_ARCH)
}
