// solve kernel instantiation: thing_1obj (StaticDims<9, 1, 4, 1>), double
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(thing_1obj, double, f64, StaticDims<9, 1, 4, 1>)
}
