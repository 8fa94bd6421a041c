// solve kernel instantiation: thing_arch (StaticDims<9, 3, 16, 3>), float
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(thing_arch, float, f32, StaticDims<9, 3, 16, 3>)
}
