// solve kernel instantiation: generic9 (RuntimeDims<9>), F = float
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(generic9, float, f32, RuntimeDims<9>)
}
