// solve kernel instantiation: thing_obs12 (UB_DIMS_THING_OBS12), F = float
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(thing_obs12, float, f32, UB_DIMS_THING_OBS12)
}
