// solve kernel instantiation: thing_robust8 (StaticDims<9, 1, 32, 8>), double
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(thing_robust8, double, f64, StaticDims<9, 1, 32, 8>)
}
