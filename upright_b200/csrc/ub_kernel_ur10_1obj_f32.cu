// solve kernel instantiation: ur10_1obj (UB_DIMS_UR10_1OBJ), F = float
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(ur10_1obj, float, f32, UB_DIMS_UR10_1OBJ)
}
