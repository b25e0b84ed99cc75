// Kernel instantiations of one epilogue family (see layer_launch.h).
#include "layer_launch.h"

namespace qv2x {
int run_layer(const LayerLaunch& c, const FixedEpilogue<3, false, true, true>& e) { return run_layer_impl<3>(c, e); }
int run_layer(const LayerLaunch& c, const FixedEpilogue<3, false, false, true>& e) { return run_layer_impl<3>(c, e); }
}  // namespace qv2x
