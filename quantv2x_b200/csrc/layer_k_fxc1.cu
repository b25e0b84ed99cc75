// Kernel instantiations of one epilogue family (see layer_launch.h).
#include "layer_launch.h"

namespace qv2x {
int run_layer(const LayerLaunch& c, const FixedEpilogueC<1, true>& e) { return run_layer_impl<1>(c, e); }
int run_layer(const LayerLaunch& c, const FixedEpilogueC<1, false>& e) { return run_layer_impl<1>(c, e); }
}  // namespace qv2x
