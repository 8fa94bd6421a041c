// solve kernel instantiation: generic (RuntimeDims), double
#include "ub_launch.cuh"
namespace ub {
UB_DEFINE_LAUNCHER(generic, double, f64, RuntimeDims)
}
