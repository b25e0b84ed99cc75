// Kernel instantiations of one epilogue family (see layer_launch.h).
#include "layer_launch.h"

namespace qv2x {
int run_layer(const LayerLaunch& c, const RequantEpilogue<3, true, false>& e) { return run_layer_impl<3>(c, e); }
int run_layer(const LayerLaunch& c, const RequantEpilogue<3, false, false>& e) { return run_layer_impl<3>(c, e); }
}  // namespace qv2x
