// solve kernel instantiation: generic6 (RuntimeDims<6>), F = float
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(generic6, float, f32, RuntimeDims<6>)
}
