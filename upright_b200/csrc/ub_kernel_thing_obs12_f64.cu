// solve kernel instantiation: thing_obs12 (StaticDims<9, 1, 4, 1, 12>), double
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(thing_obs12, double, f64, StaticDims<9, 1, 4, 1, 12>)
}
