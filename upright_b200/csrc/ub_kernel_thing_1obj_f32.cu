// solve kernel instantiation: thing_1obj (StaticDims<9, 1, 4, 1>), float
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(thing_1obj, float, f32, StaticDims<9, 1, 4, 1>)
}
