// solve kernel instantiation: generic (RuntimeDims), float
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(generic, float, f32, RuntimeDims)
}
