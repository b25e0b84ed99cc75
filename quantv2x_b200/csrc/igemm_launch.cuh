// Host-side launch helpers for igemm.cuh shared by the layer / codebook translation units.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "host_common.h"
#include "igemm.cuh"

namespace qv2x {

// Largest magnitude representable by three signed base-256 digits, each in [-128, 127].
constexpr double kDigitMax = 127.0 * 65536.0 + 127.0 * 256.0 + 127.0;

inline void split_digits(long long m, int8_t d[3]) {
    // balanced base-256: m = d[0]*65536 + d[1]*256 + d[2], every digit in [-128, 127]
    long long lo = ((m + 128) & 255) - 128;
    m = (m - lo) / 256;
    long long mid = ((m + 128) & 255) - 128;
    m = (m - mid) / 256;
    d[0] = static_cast<int8_t>(m);
    d[1] = static_cast<int8_t>(mid);
    d[2] = static_cast<int8_t>(lo);
}

inline uint32_t make_idesc_i8(int block_n, bool b_signed) {
    uint32_t d = 0;
    d |= 2u << 4;                               // accumulator format: S32
    d |= 0u << 7;                               // A: unsigned 8-bit
    d |= (b_signed ? 1u : 0u) << 10;            // B: signed / unsigned 8-bit
    d |= static_cast<uint32_t>(block_n >> 3) << 17;
    d |= static_cast<uint32_t>(kTileM >> 4) << 24;
    return d;
}

inline void choose_tile_box(int ho, int wo, int* tw, int* th) {
    long long best = -1;
    for (int w = 128; w >= 8; w >>= 1) {
        const int h = 128 / w;
        const long long area = static_cast<long long>((wo + w - 1) / w) * w * ((ho + h - 1) / h) * h;
        if (best < 0 || area < best) {
            best = area;
            *tw = w;
            *th = h;
        }
    }
}

// Launch with programmatic stream serialization (see grid_dep_* in ptx.cuh); QV2X_PDL=0 restores plain launches.
template <class Kern, class... Args>
inline cudaError_t launch_pdl(Kern kern, int grid, int threads, int smem_bytes, cudaStream_t stream, Args... args) {
    static const int pdl = getenv("QV2X_PDL") ? atoi(getenv("QV2X_PDL")) : 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(static_cast<unsigned>(threads));
    cfg.dynamicSmemBytes = static_cast<size_t>(smem_bytes);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

template <int BLOCK_N, int BK, int G, class Epi, int TPS = 1>
static int launch_igemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const IgemmGeom& g, const Epi& epi,
                        cudaStream_t stream) {
    using Cfg = IgemmCfg<BLOCK_N, BK, Epi::kMaxStages, TPS>;
    auto kern = igemm_kernel<BLOCK_N, BK, G, Epi, TPS>;
    static bool attr_set = false;
    if (!attr_set) {
        QV2X_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        attr_set = true;
    }
    const int total = g.n_img * g.tiles_y * g.tiles_x * g.n_tiles;
    const int grid = std::min(total, num_sms());
    IgemmGeom gg = g;
    gg.debug = g_debug_flags;
    gg.trace = g_trace_ptr;
    gg.taps_h = g.taps / g.taps_w;
    gg.fd_ntiles = FastDiv(g.n_tiles);
    gg.fd_tx = FastDiv(g.tiles_x);
    gg.fd_ty = FastDiv(g.tiles_y);
    gg.tw_shift = 0;
    while ((1 << gg.tw_shift) < g.tw) ++gg.tw_shift;
    if ((1 << gg.tw_shift) != g.tw) return set_error(QV2X_ERR_INVALID, "tile width %d is not a power of two", g.tw);
    QV2X_CUDA_OK(launch_pdl(kern, grid, igemm_threads<Epi, BLOCK_N>(), Cfg::kSmemBytes, stream, tmA, tmB, gg, epi));
    g_launch_count.fetch_add(1);
    return 0;
}

// ---------------------------------------------------------------- HALO mainloop (3x3, stride 1)
// Ring depths and the weights-resident mode are run-time launch parameters of one kernel per (BLOCK_N, BK, G, Epi).
struct HaloPlan {
    int block_n, a_slots, b_stages, resident;
    long long cost;     // model: SM cycles of the slowest CTA
};

// Cost model in SM cycles (measured constants, profiles/r2_exp_mma_rate.log): an M=128 x N x K=32 int8 MMA takes
// max(N/2, (4096 + 32 N)/128) cycles (tensor pipe vs the 128 B/clk operand read port); the requant epilogue issues
// about 12 instructions per output over 4 schedulers; streamed weights arrive at about 60 B/clk per SM when all SMs
// pull from L2 together.
inline HaloPlan plan_halo(int bn, int bk, int groups, int cblocks, long long m_tiles, int n_total, bool allow_resident,
                          int force_resident) {
    HaloPlan hp{};
    hp.block_n = bn;
    const int a_slot = ((kHaloW * kHaloH * bk) + 1023) / 1024 * 1024;
    const int b_stage = bn * bk;
    const int budget = 227 * 1024 - 1024 - kBarrierBytes - kEpiSmemBytes - kHaloSmemBytes;
    const int kblocks = groups * cblocks * 9;
    const int n_tiles = n_total / bn;
    const long long tiles = m_tiles * n_tiles;
    const int sms = num_sms();
    bool resident = allow_resident && kblocks <= kHaloMaxBStages && (2 * a_slot + kblocks * b_stage) <= budget &&
                    sms >= n_tiles;
    if (force_resident == 0) resident = false;
    hp.resident = resident ? 1 : 0;
    if (resident) {
        hp.b_stages = kblocks;
        hp.a_slots = std::min(kHaloMaxASlots, (budget - kblocks * b_stage) / a_slot);
    } else {
        // streamed: windows are needed once per nine B stages and are prefetched independently, so two or three
        // slots suffice; the rest of the ring is B stages (latency cover of the weight stream)
        // (BLOCK_N <= 128: a ring of exactly nine stages, so that tap t always sits in stage t -- static descriptors)
        hp.a_slots = (bn >= 256) ? 2 : (bn == 128 ? 2 : 4);
        hp.b_stages = std::min(kHaloMaxBStages, (budget - hp.a_slots * a_slot) / b_stage);
        if (bn <= 128 && hp.b_stages >= 9) {
            hp.b_stages = 9;
            hp.a_slots = std::min(kHaloMaxASlots, (budget - 9 * b_stage) / a_slot);
        }
        if (hp.b_stages < 2) {
            hp.cost = -1;
            return hp;
        }
    }
    long long grid = std::min<long long>(tiles, sms);
    if (resident) grid = grid / n_tiles * n_tiles;
    const long long rounds = (tiles + grid - 1) / grid;
    const long long mma = static_cast<long long>(kblocks) * (bk / 32) * std::max(bn / 2, (4096 + 32 * bn) / 128);
    const long long epi = 12LL * bn * (groups > 1 ? 2 : 1);
    const long long stream = resident ? 0 : (static_cast<long long>(kblocks) * b_stage + groups * cblocks * a_slot) / 60;
    const long long per_tile = std::max(std::max(mma, epi), stream);
    // (resident weights: the one-time load overlaps the first window; measured slightly ahead of streaming whenever
    //  it fits and a CTA sees two or more tiles)
    hp.cost = rounds * per_tile + 3000 + (resident ? (rounds >= 2 ? -1 : static_cast<long long>(kblocks) * b_stage / 60) : 0);
    return hp;
}

template <int BLOCK_N, int BK, int G, class Epi, bool BRES, bool SRING>
static int launch_igemm_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const IgemmGeom& g, const Epi& epi,
                             const HaloPlan& hp, cudaStream_t stream) {
    using HCfg = HaloCfg<BLOCK_N, BK>;
    auto kern = igemm_kernel<BLOCK_N, BK, G, Epi, 1, true, BRES, SRING>;
    static int attr_bytes = 0;
    const int smem_bytes = HCfg::smem_bytes(hp.a_slots, hp.b_stages);
    if (smem_bytes > attr_bytes) {
        QV2X_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_bytes = smem_bytes;
    }
    IgemmGeom gg = g;
    gg.debug = g_debug_flags;
    gg.trace = g_trace_ptr;
    gg.taps_h = 3;
    gg.fd_ntiles = FastDiv(g.n_tiles);
    gg.fd_tx = FastDiv(g.tiles_x);
    gg.fd_ty = FastDiv(g.tiles_y);
    gg.tw_shift = 3;
    gg.a_slots = hp.a_slots;
    gg.b_stages = hp.b_stages;
    gg.b_resident = hp.resident;
    const int total = g.n_img * g.tiles_y * g.tiles_x * g.n_tiles;
    int grid = std::min(total, num_sms());
    if (hp.resident) grid = grid / g.n_tiles * g.n_tiles;
    QV2X_CUDA_OK(launch_pdl(kern, grid, igemm_threads<Epi, BLOCK_N>(), smem_bytes, stream, tmA, tmB, gg, epi));
    g_launch_count.fetch_add(1);
    return 0;
}

template <int G, class Epi>
int dispatch_igemm_halo(int bk, const CUtensorMap& tmA, const CUtensorMap& tmB, const IgemmGeom& g, const Epi& epi,
                        const HaloPlan& hp, cudaStream_t stream) {
#define QV2X_HCASE(BN, BKK)                                                                                   \
    if (hp.block_n == BN && bk == BKK) {                                                                      \
        if (hp.resident) return launch_igemm_halo<BN, BKK, G, Epi, true, false>(tmA, tmB, g, epi, hp, stream); \
        if (BN <= 128 && hp.b_stages == 9)                                                                    \
            return launch_igemm_halo<BN, BKK, G, Epi, false, (BN <= 128)>(tmA, tmB, g, epi, hp, stream);      \
        return launch_igemm_halo<BN, BKK, G, Epi, false, false>(tmA, tmB, g, epi, hp, stream);                \
    }
    if constexpr (G == 1) {
        QV2X_HCASE(256, 128)
    }
    QV2X_HCASE(128, 128)
    QV2X_HCASE(64, 128)
    QV2X_HCASE(128, 64)
    QV2X_HCASE(64, 64)
#undef QV2X_HCASE
    return set_error(QV2X_ERR_INVALID, "no HALO igemm instantiation for BLOCK_N=%d BK=%d G=%d", hp.block_n, bk, G);
}

template <int G, class Epi>
int dispatch_igemm(int block_n, int bk, const CUtensorMap& tmA, const CUtensorMap& tmB, const IgemmGeom& g,
                   const Epi& epi, cudaStream_t stream) {
#define QV2X_CASE(BN, BKK) \
    if (block_n == BN && bk == BKK) return launch_igemm<BN, BKK, G, Epi>(tmA, tmB, g, epi, stream);
    if constexpr (G == 1) {
        // 3x3 convs over 64-byte channel blocks: three taps per pipeline stage
        if (bk == 64 && g.taps == 9 && g.cblocks == 1) {
            if (block_n == 64) return launch_igemm<64, 64, G, Epi, 3>(tmA, tmB, g, epi, stream);
            if (block_n == 128) return launch_igemm<128, 64, G, Epi, 3>(tmA, tmB, g, epi, stream);
        }
        QV2X_CASE(256, 128)
        QV2X_CASE(256, 64)
    }
    QV2X_CASE(128, 128)
    QV2X_CASE(128, 64)
    QV2X_CASE(64, 128)
    QV2X_CASE(64, 64)
#undef QV2X_CASE
    return set_error(QV2X_ERR_INVALID, "no igemm instantiation for BLOCK_N=%d BK=%d G=%d", block_n, bk, G);
}


inline int make_weight_tmap(CUtensorMap* tm, const void* d_w, int rows, int k_total, int block_n, int bk) {
    const uint64_t dims[2] = {static_cast<uint64_t>(k_total), static_cast<uint64_t>(rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(k_total)};
    const uint32_t box[2] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(block_n)};
    const uint32_t es[2] = {1, 1};
    return encode_tmap_u8(tm, d_w, 2, dims, strides, box, es, bk);
}

// The tile's (16+2) x (8+2) pixel input window as ONE box of the HALO mainloop (element strides 1).
inline int make_halo_tmap(CUtensorMap* tm, const void* d_x, int n_img, int hi, int wi, int cstride, int bk) {
    const uint64_t dims[4] = {static_cast<uint64_t>(cstride), static_cast<uint64_t>(wi), static_cast<uint64_t>(hi),
                              static_cast<uint64_t>(n_img)};
    const uint64_t strides[3] = {static_cast<uint64_t>(cstride), static_cast<uint64_t>(cstride) * wi,
                                 static_cast<uint64_t>(cstride) * wi * hi};
    const uint32_t box[4] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(kHaloW), static_cast<uint32_t>(kHaloH), 1};
    const uint32_t es[4] = {1, 1, 1, 1};
    return encode_tmap_u8(tm, d_x, 4, dims, strides, box, es, bk);
}

inline int make_act_tmap(CUtensorMap* tm, const void* d_x, int n_img, int hi, int wi, int cstride, int tw, int th,
                  int stride, int bk) {
    const uint64_t dims[4] = {static_cast<uint64_t>(cstride), static_cast<uint64_t>(wi), static_cast<uint64_t>(hi),
                              static_cast<uint64_t>(n_img)};
    const uint64_t strides[3] = {static_cast<uint64_t>(cstride), static_cast<uint64_t>(cstride) * wi,
                                 static_cast<uint64_t>(cstride) * wi * hi};
    const uint32_t box[4] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(tw * stride),
                             static_cast<uint32_t>(th * stride), 1};
    const uint32_t es[4] = {1, static_cast<uint32_t>(stride), static_cast<uint32_t>(stride), 1};
    return encode_tmap_u8(tm, d_x, 4, dims, strides, box, es, bk);
}

}  // namespace qv2x
