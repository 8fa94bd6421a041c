// solve kernel instantiation: ur10_1obj (StaticDims<6, 1, 4, 1>), float
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(ur10_1obj, float, f32, StaticDims<6, 1, 4, 1>)
}
