// solve kernel instantiation: thing_1obj (UB_DIMS_THING_1OBJ), F = float
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(thing_1obj, float, f32, UB_DIMS_THING_1OBJ)
}
