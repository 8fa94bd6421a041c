// solve kernel instantiation: thing_arch (StaticDims<9, 3, 16, 3>), double
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(thing_arch, double, f64, StaticDims<9, 3, 16, 3>)
}
