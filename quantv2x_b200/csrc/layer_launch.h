// Launch context of one quantized layer and the per-epilogue launch entry points.  The igemm kernel is instantiated
// once per (tile shape, epilogue type); the epilogue families live in separate translation units (layer_k_*.cu) so that
// the build parallelises -- layer.cu only fills the epilogue structs and calls run_layer().
#pragma once
#include "epilogue_fixed.cuh"
#include "epilogue_requant.cuh"
#include "igemm_launch.cuh"

namespace qv2x {

struct LayerLaunch {
    int block_n, bk;
    bool halo;
    CUtensorMap tmA, tmB;
    IgemmGeom g;
    HaloPlan hp;
    cudaStream_t stream;
};

template <int G, class Epi>
inline int run_layer_impl(const LayerLaunch& c, const Epi& e) {
    if (c.halo) return dispatch_igemm_halo<G>(c.bk, c.tmA, c.tmB, c.g, e, c.hp, c.stream);
    return dispatch_igemm<G>(c.block_n, c.bk, c.tmA, c.tmB, c.g, e, c.stream);
}

#define QV2X_LAYER_EPILOGUES(X)                   \
    X(1, (FixedEpilogue<1, false, true, false>))  \
    X(1, (FixedEpilogue<1, false, false, false>)) \
    X(1, (FixedEpilogue<1, false, true, true>))   \
    X(1, (FixedEpilogue<1, false, false, true>))  \
    X(3, (FixedEpilogue<3, false, true, false>))  \
    X(3, (FixedEpilogue<3, false, false, false>)) \
    X(3, (FixedEpilogue<3, false, true, true>))   \
    X(3, (FixedEpilogue<3, false, false, true>))  \
    X(3, (FixedEpilogue<3, true, false, false>))  \
    X(3, (FixedEpilogue<3, true, false, true>))   \
    X(1, (FixedEpilogueC<1, true>))               \
    X(1, (FixedEpilogueC<1, false>))              \
    X(3, (FixedEpilogueC<3, true>))               \
    X(3, (FixedEpilogueC<3, false>))              \
    X(1, (RequantEpilogue<1, false, false>))      \
    X(1, (RequantEpilogue<1, false, true, true>)) \
    X(3, (RequantEpilogue<3, true, false>))       \
    X(3, (RequantEpilogue<3, false, false>))      \
    X(3, (RequantEpilogue<3, false, true>))

#define QV2X_UNPAREN(...) __VA_ARGS__
#define QV2X_DECL(G, E) int run_layer(const LayerLaunch& c, const QV2X_UNPAREN E& e);
QV2X_LAYER_EPILOGUES(QV2X_DECL)
#undef QV2X_DECL

}  // namespace qv2x
