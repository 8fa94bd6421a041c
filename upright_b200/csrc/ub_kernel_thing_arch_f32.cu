// solve kernel instantiation: thing_arch (UB_DIMS_THING_ARCH), F = float
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(thing_arch, float, f32, UB_DIMS_THING_ARCH)
}
