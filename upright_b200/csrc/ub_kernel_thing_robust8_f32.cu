// solve kernel instantiation: thing_robust8 (StaticDims<9, 1, 32, 8>), float
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(thing_robust8, float, f32, StaticDims<9, 1, 32, 8>)
}
