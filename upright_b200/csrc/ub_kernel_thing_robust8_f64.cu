// solve kernel instantiation: thing_robust8 (UB_DIMS_THING_ROBUST8), F = double
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(thing_robust8, double, f64, UB_DIMS_THING_ROBUST8)
}
