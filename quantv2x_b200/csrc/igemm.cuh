// Persistent, warp-specialised int8 implicit-GEMM for sm_100a.
//
//   warp 0      : TMA producer   (A = activation tile via a 4-D tiled tensor map -> im2col for free,
//                                 zero OOB fill = conv zero padding; B = weight tile via a 2-D map)
//   warp 1      : MMA issuer     (tcgen05.mma kind::i8, M=128, N=BLOCK_N, K=32 per instruction,
//                                 int32 accumulators in a ring of TMEM slots)
//   warp 2      : TMEM allocator
//   warps 4..11 : epilogue       (tcgen05.ld -> registers -> fused epilogue functor -> global)
//
// One output tile = 128 output pixels (a tw x th box of one image) x BLOCK_N output columns.
// A tile reduces over G "groups"; every group owns one TMEM slot (its own int32 accumulator):
//   G = 1 : ordinary quantized conv (QuantModule over nn.Conv2d, reference quant_layer.py:391-410)
//   G = 3 : (a) the shrinker's first conv, whose input is the concat of three tensors with three
//               activation scales (reference base_bev_backbone.py:111-112), one group per scale;
//           (b) GEMMs whose real-valued weights are carried as three signed base-256 digits
//               (deblocks with per-input-channel scales, folded codebook distance GEMM).
// K is walked group-major, then tap-major, then BK-byte channel blocks.
#pragma once
#include <type_traits>

#include "ptx.cuh"

namespace qv2x {

constexpr int kMaxGroups = 3;
constexpr int kMaxSteps = 24;
constexpr int kTileM = 128;

// Exact division of a tile / column index by a run-time constant without the ~35-instruction division sequence:
// q = umulhi(n, ceil(2^32 / d)) is exact for n * d < 2^32 (tile counts and column counts are far below that).
struct FastDiv {
    uint32_t mul, d;
    __host__ __device__ FastDiv() : mul(0), d(1) {}
    __host__ explicit FastDiv(int dd) : mul(dd > 1 ? static_cast<uint32_t>((0x100000000ull + dd - 1) / dd) : 0u),
                                        d(static_cast<uint32_t>(dd)) {}
    __device__ __forceinline__ int div(int n) const {
        return d == 1 ? n : static_cast<int>(__umulhi(static_cast<uint32_t>(n), mul));
    }
};

struct IgemmGeom {
    // M-tile grid: per image Ho x Wo "anchor" pixels covered by tw x th boxes (tw * th == 128)
    int n_img, Ho, Wo, tw, th, tiles_x, tiles_y, n_tiles, block_n;
    // A source image extent (pixels) -- used by epilogues that need input-side bounds
    int Hi, Wi;
    // taps and K structure
    int taps, taps_w, taps_h, stride, pad;
    int tw_shift;                // log2(tw)
    FastDiv fd_ntiles, fd_tx, fd_ty;   // dividers of decode_tile (filled in by launch_igemm)
    int groups, cblocks;
    int a_c_base[kMaxGroups];    // channel coordinate of the group's first k-block in the A tensor
    int b_row_base[kMaxGroups];  // row coordinate of the group's first output column in the B tensor
    int b_k_base[kMaxGroups];    // K coordinate (within one tap) of the group's first k-block in B
    int b_k_tap_stride;          // K distance between consecutive taps in B
    uint32_t idesc;              // tcgen05 instruction descriptor (operand signedness, M, N)
    // Sequential N steps per tile (1 for conv layers).  A tile visits its steps in order on ONE CTA, so an
    // epilogue may carry state from step to step (the codebook's level-by-level argmin).  B rows of
    // (step, group) start at b_row_base[g] + step_row_base[step] + g * step_group_stride[step].
    long long* trace;   // optional [grid][kTraceTiles][16] clock64 stamps per role (qv2x_debug_trace), else nullptr
    int debug;   // bring-up knobs (qv2x_set_debug_flags): 1 skip epilogue math, 2 skip MMA issue, 4 skip A loads, 8 skip B loads
    int n_steps;
    int step_row_base[kMaxSteps];
    int step_group_stride[kMaxSteps];
    // HALO mainloop (3x3 stride-1 convs): the (th+2) x (tw+2) input window of a tile is loaded ONCE per channel block
    // and all nine taps read it through shifted shared-memory descriptors; B (weights) is either streamed one
    // (tap, channel block) stage at a time or, when the whole [BLOCK_N x K] slice fits, loaded once per CTA.
    int a_slots, b_stages, b_resident;
};

// Role timelines (qv2x_debug_trace) and role knock-outs (qv2x_set_debug_flags) are bring-up instruments: they cost
// constant loads, branches and R2UR moves inside the single-thread TMA / MMA issue loops, which is exactly what paces
// the short layers (profiles/r2_exp_mma_rate.log), so the product build compiles them out.
#ifdef QV2X_IGEMM_DEBUG
constexpr bool kIgemmDebug = true;
#else
constexpr bool kIgemmDebug = false;
#endif
constexpr int kTraceTiles = 32;
__device__ __forceinline__ void trace_stamp(const IgemmGeom& g, int tile_local, int slot) {
    if constexpr (kIgemmDebug) {
        if (g.trace != nullptr && tile_local < kTraceTiles)
            g.trace[(static_cast<long long>(blockIdx.x) * kTraceTiles + tile_local) * 16 + slot] = clock64();
    }
}
__device__ __forceinline__ int dbg_flags(const IgemmGeom& g) { return kIgemmDebug ? g.debug : 0; }

constexpr int kBarrierBytes = 1024;    // mbarriers + the TMEM base address
constexpr int kEpiSmemBytes = 12288;   // two side-input slots / epilogue scratch
constexpr int kHaloSmemBytes = 8192;   // two halo buffers, one per side warp   // scratch handed to the epilogue functor (cross-warp merges)

// TPS = taps per pipeline stage: layers with 64 input channels have only 64 bytes of K per tap, so three taps
// share one stage (one mbarrier round trip per 192 bytes of K instead of per 64).
template <int BLOCK_N, int BK, int MAX_STAGES = 8, int TPS = 1>
struct IgemmCfg {
    static constexpr int kASub = kTileM * BK;
    static constexpr int kBSub = BLOCK_N * BK;
    static constexpr int kATile = kASub * TPS;
    static constexpr int kBTile = kBSub * TPS;
    static constexpr int kStageBytes = kATile + kBTile;
    // 227 KB per CTA minus alignment slack, barriers, the epilogue / side-input scratch and the halo buffers
    static constexpr int kStagesRaw = (227 * 1024 - 1024 - kBarrierBytes - kEpiSmemBytes - kHaloSmemBytes) / kStageBytes;
    static constexpr int kStages = kStagesRaw > MAX_STAGES ? MAX_STAGES : kStagesRaw;
    static constexpr int kSlots = (512 / BLOCK_N) > 4 ? 4 : (512 / BLOCK_N);
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + kBarrierBytes + kEpiSmemBytes + kHaloSmemBytes;
};

// Shared-memory budget of the HALO mainloop: A slots hold one 18 x 10 pixel window of CB channel bytes (rounded up to
// the swizzle pattern), B stages one tap x CB bytes of K for BLOCK_N output columns.  Ring depths are chosen at launch.
constexpr int kHaloTileW = 8, kHaloTileH = 16;     // output pixel box of a tile: 8-row operand groups = 8 pixels in x
constexpr int kHaloW = kHaloTileW + 2, kHaloH = kHaloTileH + 2;
constexpr int kHaloMaxBStages = 32, kHaloMaxASlots = 4;
template <int BLOCK_N, int CB>
struct HaloCfg {
    static constexpr int kABytes = kHaloW * kHaloH * CB;                       // bytes one TMA box delivers
    static constexpr int kASlot = (kABytes + 1023) / 1024 * 1024;
    static constexpr int kBStage = BLOCK_N * CB;
    static constexpr int kSlots = (512 / BLOCK_N) > 4 ? 4 : (512 / BLOCK_N);
    static constexpr int kRingBudget = 227 * 1024 - 1024 - kBarrierBytes - kEpiSmemBytes - kHaloSmemBytes;
    static constexpr int smem_bytes(int a_slots, int b_stages) {
        return a_slots * kASlot + b_stages * kBStage + 1024 + kBarrierBytes + kEpiSmemBytes + kHaloSmemBytes;
    }
};

struct TileCoord {
    int img, ty, tx, nt;
};
// Where the accumulators of the current step live in TMEM, for epilogues that re-read them in step_end:
// addr[g] = column 0 of group g's slot for the calling warp's lane quadrant; the warp owns columns [c_begin, c_end).
struct TmemView {
    uint32_t addr[kMaxGroups];
    int c_begin, c_end;
};
__device__ __forceinline__ TileCoord decode_tile(const IgemmGeom& g, int t) {
    TileCoord c;
    int m = g.fd_ntiles.div(t);
    c.nt = t - m * g.n_tiles;
    int q = g.fd_tx.div(m);
    c.tx = m - q * g.tiles_x;
    m = q;
    q = g.fd_ty.div(m);
    c.ty = m - q * g.tiles_y;
    c.img = q;
    return c;
}

// Epilogue contract:
//   struct Epi {
//     struct Tile;                                     // per-thread, per-tile state
//     static constexpr int col_split(int block_n);     // 1, 2 or 4: epilogue warps per TMEM lane quadrant
//     static constexpr int kMaxStages;                 // cap of the smem ring depth (frees L1 for gathers)
//     static constexpr bool kSideWarp;                 // warp 3 stages the tile's side inputs in shared memory
//     static constexpr bool kSeqDrain;                 // G > 1: accumulator groups are drained one by one
//     struct Side; __device__ void side_init(Side&, const IgemmGeom&, int lane) const;   -- once per kernel
//     __device__ void side_load(const IgemmGeom&, const TileCoord&, int lane, uint8_t* slot, int32_t* halo,
//                               int& staged_nt, const Side&) const;
//         -- kSideWarp: run by the 32 lanes of warp 3, one tile AHEAD of the epilogue, into one of two
//            kEpiSmemBytes/2 slots (per-column parameters, per-row receptive-field sums, ...)
//     __device__ void begin(Tile&, const IgemmGeom&, const TileCoord&, int row, const uint8_t* slot) const;
//         -- called BEFORE the accumulators are ready
//     template <int W> __device__ void chunk(Tile&, const IgemmGeom&, const TileCoord&, int step, int col0,
//                           const int32_t (*acc)[W]) const;
//         -- W consecutive columns [col0, col0+W) of this thread's row in `step`, acc[g][j] (all groups at once)
//     kSeqDrain only:
//     template <int W> __device__ void accum(const Tile&, const IgemmGeom&, int grp, int col0, const int32_t (&acc)[W],
//                           float (&v)[W]) const;      -- fold group grp's accumulators into the running sum v
//     template <int W> __device__ void finish(Tile&, int col0, const float (&v)[W]) const;
//     static constexpr bool kHoldSlots;                -- step_end runs BEFORE the step's TMEM slots are released
//     __device__ void step_end(Tile&, const IgemmGeom&, const TileCoord&, int step, int part, int quad, int lane,
//                              uint8_t* scratch, const TmemView&) const;   -- scratch: kEpiSmemBytes of shared memory
//     __device__ void end(Tile&, const IgemmGeom&, const TileCoord&) const;
//   };
// Calls f(integral_constant<int, I * STEP>) for the I < COUNT with v == I * STEP: turns a run-time column base into
// a compile-time one (epilogues whose per-column parameters are immediate constant-bank operands).
template <int I, int COUNT, int STEP, class F>
__device__ __forceinline__ void dispatch_col(int v, F& f) {
    if constexpr (I < COUNT) {
        if (v == I * STEP) f(std::integral_constant<int, I * STEP>{});
        else dispatch_col<I + 1, COUNT, STEP>(v, f);
    }
}
template <class Epi, class = void>
struct EpiStaticCols : std::false_type {};
template <class Epi>
struct EpiStaticCols<Epi, std::enable_if_t<Epi::kStaticCols>> : std::true_type {};

// Optional epilogue trait `static constexpr int tile_split(int block_n)` (1 or 2): with 2, TWO groups of
// 4 * col_split epilogue warps alternate over the CTA's tiles (group j takes the tiles k = j mod 2 of the CTA's
// sequence, i.e. the side-input slot j), so the per-tile fixed costs of an epilogue warp (side-input barrier, row-sum
// loads, accumulator barrier + fence, slot release, tile decode: ~1.2 k cycles) overlap with the other group's chunk
// loop.  Short-K layers are bound by exactly these costs (profiles/r2_trace_short_layers.log).
template <class Epi, class = void>
struct EpiTileSplit {
    static constexpr int get(int) { return 1; }
};
template <class Epi>
struct EpiTileSplit<Epi, decltype(void(Epi::tile_split(0)))> {
    static constexpr int get(int block_n) { return Epi::tile_split(block_n); }
};

// Optional two-phase side staging (epilogue members SidePre / side_prefetch / side_store).
template <class Epi, class = void>
struct EpiHasSidePre {
    static constexpr bool value = false;
};
template <class Epi>
struct EpiHasSidePre<Epi, decltype(void(sizeof(typename Epi::SidePre)))> {
    static constexpr bool value = true;
};

template <class Epi, int BLOCK_N>
constexpr int igemm_threads() {
    return (4 + 4 * Epi::col_split(BLOCK_N) * EpiTileSplit<Epi>::get(BLOCK_N)) * 32;
}

template <int BLOCK_N, int BK, int G, class Epi, int TPS = 1, bool HALO = false, bool BRES = false, bool SRING = false>
__global__ void __launch_bounds__(igemm_threads<Epi, BLOCK_N>(), 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ IgemmGeom g, const __grid_constant__ Epi epi) {
    using Cfg = IgemmCfg<BLOCK_N, BK, Epi::kMaxStages, TPS>;
    using HCfg = HaloCfg<BLOCK_N, BK>;
    constexpr int kStages = HALO ? kHaloMaxBStages : Cfg::kStages;   // barrier slots (HALO: ring depth is g.b_stages)
    constexpr int kSlots = Cfg::kSlots;
    static_assert(G <= kSlots, "every group needs its own TMEM slot");
    static_assert(!HALO || TPS == 1, "the HALO mainloop has one tap per B stage");
    static_assert(!(BRES && SRING), "a static ring is a streamed-weights mode");
    constexpr int kColSplit = Epi::col_split(BLOCK_N);
    constexpr int kTileSplit = EpiTileSplit<Epi>::get(BLOCK_N);
    static_assert(kTileSplit == 1 || (kTileSplit == 2 && Epi::kSideWarp), "tile groups map onto the two side slots");
    constexpr int kNumEpiWarps = 4 * kColSplit;      // warps that work on ONE tile (barrier arrival counts)
    static_assert(BLOCK_N % (16 * kColSplit) == 0, "BLOCK_N must split into 16-column-aligned parts");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int ring_bytes = HALO ? g.a_slots * HCfg::kASlot + g.b_stages * HCfg::kBStage : Cfg::kStages * Cfg::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ring_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tfull_bar = bars + 2 * kStages;
    uint64_t* tempty_bar = bars + 2 * kStages + kSlots;
    uint64_t* side_full = bars + 2 * kStages + 2 * kSlots;      // [2] side-input slots (warp 3 -> epilogue)
    uint64_t* side_empty = side_full + 2;
    uint64_t* afull_bar = side_empty + 2;                        // HALO: [kHaloMaxASlots] activation windows
    uint64_t* aempty_bar = afull_bar + kHaloMaxASlots;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(aempty_bar + kHaloMaxASlots);
    static_assert((2 * kHaloMaxBStages + 2 * 4 + 4 + 2 * kHaloMaxASlots) * 8 + 4 <= kBarrierBytes, "barrier area");
    uint8_t* epi_scratch = smem + ring_bytes + kBarrierBytes;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(smem_u32(&full_bar[i]), 1);
            mbar_init(smem_u32(&empty_bar[i]), 1);
        }
        for (int i = 0; i < kSlots; ++i) {
            mbar_init(smem_u32(&tfull_bar[i]), 1);
            mbar_init(smem_u32(&tempty_bar[i]), kNumEpiWarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&side_full[i]), 1);
            mbar_init(smem_u32(&side_empty[i]), kNumEpiWarps);
        }
        if constexpr (HALO) {
            for (int i = 0; i < kHaloMaxASlots; ++i) {
                mbar_init(smem_u32(&afull_bar[i]), 1);
                mbar_init(smem_u32(&aempty_bar[i]), 1);
            }
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_ptr), 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // Programmatic dependent launch: the next layer's CTAs may move onto SMs as this grid frees them and run their
    // own set-up (and weight loads) there; every role of THIS kernel waits for the previous layer to complete before
    // it touches activations, row sums or output buffers (the producer first requests its resident weights).
    grid_dep_launch_dependents();
    if (!(HALO && BRES && warp == 0)) grid_dep_wait();

    const int total_tiles = g.n_img * g.tiles_y * g.tiles_x * g.n_tiles;
    const int kblocks_per_group = (g.taps / TPS) * g.cblocks;   // pipeline stages per accumulator group

    // The producer and MMA roles run their loops on the WHOLE warp (every lane computes the same, warp-uniform
    // values, so they live in uniform registers) and elect one lane only for the instructions that must be issued
    // once.  All addressing is incremental: a single thread's dependent integer chain (divisions, 64-bit
    // descriptor assembly) between two TMA / MMA instructions was what paced these roles before.
    if (HALO && warp == 0) {
        // ------------------------------------------------------------ TMA producer, HALO mainloop
        // Two independent streams from one warp: activation windows (ONE 4-D box per (tile, group, channel block) = the
        // tile's 18 x 10 pixel input window, zero fill outside the image = conv padding) and -- unless the weights are
        // resident -- B stages, one per tap.  Windows are prefetched as far ahead as free slots allow: while the
        // B stream waits for a free stage the warp keeps polling the window ring, so a window is requested the moment
        // its slot is released instead of after the previous group's nine B stages (one L2 round trip too late).
        const bool leader = elect_one();
        const uint32_t a_base = smem_u32(smem);
        const uint32_t b_base = a_base + g.a_slots * HCfg::kASlot;
        const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
        const uint32_t afull0 = smem_u32(&afull_bar[0]), aempty0 = smem_u32(&aempty_bar[0]);
        const uint32_t a_slots = g.a_slots, b_stages = g.b_stages;
        const int n_grp_cb = G * g.cblocks;
        if (BRES && leader) {
            // every CTA keeps one column tile (the grid is a multiple of n_tiles): its whole B slice is loaded once
            const int ncol0 = static_cast<int>(blockIdx.x % g.n_tiles) * BLOCK_N;
            int kb = 0;
            for (int grp = 0; grp < G; ++grp)
                for (int cb = 0; cb < g.cblocks; ++cb)
                    for (int tap = 0; tap < 9; ++tap, ++kb) {
                        const uint32_t fb = full0 + 8 * kb;
                        mbar_expect_tx(fb, HCfg::kBStage);
                        tma_load_2d(b_base + kb * HCfg::kBStage, &tmB, fb,
                                    g.b_k_base[grp] + tap * g.b_k_tap_stride + cb * BK, g.b_row_base[grp] + ncol0);
                    }
        }
        if constexpr (BRES) grid_dep_wait();       // the weights are on their way; activations need the previous layer
        // window stream cursor
        int a_t = blockIdx.x, a_gc = 0, a_x0 = 0, a_y0 = 0, a_img = 0;
        uint32_t sa = 0, aph = 0;
        long long a_issued = 0;                 // windows requested so far
        auto a_tile = [&]() {
            const TileCoord tc = decode_tile(g, a_t);
            a_x0 = tc.tx * kHaloTileW - g.pad;
            a_y0 = tc.ty * kHaloTileH - g.pad;
            a_img = tc.img;
        };
        if (a_t < total_tiles) a_tile();
        auto a_issue = [&]() {                  // request the next window into slot sa (which must be free)
            if (leader) {
                const uint32_t fb = afull0 + 8 * sa;
                const int grp = (G == 1) ? 0 : a_gc / g.cblocks;
                const int cb = (G == 1) ? a_gc : a_gc - grp * g.cblocks;
                mbar_expect_tx(fb, HCfg::kABytes);
                tma_load_4d(a_base + sa * HCfg::kASlot, &tmA, fb, g.a_c_base[grp] + cb * BK, a_x0, a_y0, a_img);
            }
            if (++sa == a_slots) sa = 0, aph ^= 1;
            ++a_issued;
            if (++a_gc == n_grp_cb) {
                a_gc = 0;
                a_t += gridDim.x;
                if (a_t < total_tiles) a_tile();
            }
        };
        if constexpr (BRES) {
            while (a_t < total_tiles) {
                mbar_wait(aempty0 + 8 * sa, aph ^ 1);
                a_issue();
            }
        } else {
            uint32_t sb = 0, bph = 0;
            long long b_group = 0;              // (tile, group, channel block) index of the B stream
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int ncol0 = (t - g.fd_ntiles.div(t) * g.n_tiles) * BLOCK_N;
                for (int grp = 0; grp < G; ++grp) {
                    const int brow = g.b_row_base[grp] + ncol0;
                    for (int cb = 0; cb < g.cblocks; ++cb, ++b_group) {
                        // the consumer needs this group's window before its first tap: never let B run ahead of it
                        while (a_issued <= b_group) {
                            mbar_wait(aempty0 + 8 * sa, aph ^ 1);
                            a_issue();
                        }
                        int bk = g.b_k_base[grp] + cb * BK;
                        for (int tap = 0; tap < 9; ++tap, bk += g.b_k_tap_stride) {
                            while (!mbar_try_wait(empty0 + 8 * sb, bph ^ 1)) {
                                if (a_t < total_tiles && mbar_try_wait(aempty0 + 8 * sa, aph ^ 1)) a_issue();
                            }
                            if (leader) {
                                const uint32_t fb = full0 + 8 * sb;
                                mbar_expect_tx(fb, HCfg::kBStage);
                                tma_load_2d(b_base + sb * HCfg::kBStage, &tmB, fb, bk, brow);
                            }
                            if (++sb == b_stages) sb = 0, bph ^= 1;
                        }
                        if (a_t < total_tiles && mbar_try_wait(aempty0 + 8 * sa, aph ^ 1)) a_issue();
                    }
                }
            }
        }
    } else if (HALO && warp == 1) {
        // ------------------------------------------------------------ MMA issuer, HALO mainloop
        // Tap (ky, kx) of the window is the SAME shared-memory bytes read through a descriptor whose start address is
        // moved by (ky * 10 + kx) pixels: 8-row operand groups are 8 consecutive window pixels, consecutive groups one
        // window row (10 pixels) apart (stride byte offset), and the hardware swizzles on absolute address bits
        // (profiles/r2_exp_shift_descriptor.log).  The nine taps are unrolled, so every descriptor is base + constant.
        const bool leader = elect_one();
        const uint32_t a_base = smem_u32(smem);
        const uint32_t b_base = a_base + g.a_slots * HCfg::kASlot;
        const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
        const uint32_t afull0 = smem_u32(&afull_bar[0]), aempty0 = smem_u32(&aempty_bar[0]);
        const uint32_t tfull0 = smem_u32(&tfull_bar[0]), tempty0 = smem_u32(&tempty_bar[0]);
        const uint64_t a_desc0 = umma_smem_desc_sbo(a_base, BK, kHaloW * BK);
        const uint64_t b_desc0 = umma_smem_desc(b_base, BK);
        const uint32_t a_slots = g.a_slots, b_stages = g.b_stages;
        const int cblocks = g.cblocks;
        const uint32_t idesc = g.idesc;
        uint32_t sa = 0, aph = 0, sb = 0, bph = 0, slot = 0, tph = 0;
        bool first = true;
        int tl = 0;
        // Every descriptor of a tap is a compile-time offset from two per-channel-block bases: the window slot and
        // (weights resident) the block's first B stage, or (streamed, ring of exactly nine stages) the ring itself,
        // so that tap t always lives in stage t.  The issuing thread executes ~20 instructions per tap; with a
        // run-time stage index it was ~45 and paced every BLOCK_N <= 128 layer (profiles/r2_exp_mma_rate.log).
        constexpr bool kStaticB = BRES || SRING;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
            if (leader) trace_stamp(g, tl, 2);
            uint32_t kb0 = 0;                          // resident: first B stage of the (group, channel block)
            for (int grp = 0; grp < G; ++grp) {
                mbar_wait(tempty0 + 8 * slot, tph ^ 1);
                tcgen05_fence_after();
                if (leader && grp == 0) trace_stamp(g, tl, 3);
                const uint32_t d_tmem = tmem_base + slot * BLOCK_N;
                for (int cb = 0; cb < cblocks; ++cb, kb0 += 9) {
                    mbar_wait(afull0 + 8 * sa, aph);
                    tcgen05_fence_after();
                    if (leader && grp == 0 && cb == 0) trace_stamp(g, tl, 4);
                    const uint64_t a_desc = a_desc0 + static_cast<uint64_t>(sa * (HCfg::kASlot >> 4));
                    const uint64_t b_desc_cb = b_desc0 + static_cast<uint64_t>(BRES ? kb0 * (HCfg::kBStage >> 4) : 0u);
                    const uint32_t full_cb = full0 + (BRES ? 8 * kb0 : 0u);
                    if (BRES && !first) {
                        // resident weights, steady state: nothing to wait for between taps -- all 9 x BK/32 MMAs of the
                        // channel block are issued back to back from two descriptor bases
                        if (leader) {
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                                for (int k = 0; k < BK / 32; ++k)
                                    umma_i8(d_tmem, a_desc + ((((tap / 3) * kHaloW + (tap % 3)) * BK) >> 4) + 2 * k,
                                            b_desc_cb + (tap * (HCfg::kBStage >> 4) + 2 * k), idesc,
                                            (tap | k) != 0 ? 1u : static_cast<uint32_t>(cb));
                        }
                        __syncwarp();
                    } else {
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if constexpr (kStaticB) {
                            mbar_wait(full_cb + 8 * tap, BRES ? 0u : bph);
                            tcgen05_fence_after();
                        } else {
                            mbar_wait(full0 + 8 * sb, bph);
                            tcgen05_fence_after();
                        }
                        if (leader) {
                            const uint64_t b_desc =
                                kStaticB ? b_desc_cb + static_cast<uint64_t>(tap * (HCfg::kBStage >> 4))
                                         : b_desc0 + static_cast<uint64_t>(sb * (HCfg::kBStage >> 4));
#pragma unroll
                            for (int k = 0; k < BK / 32; ++k)
                                umma_i8(d_tmem, a_desc + ((((tap / 3) * kHaloW + (tap % 3)) * BK) >> 4) + 2 * k,
                                        b_desc + 2 * k, idesc, (tap | k) != 0 ? 1u : static_cast<uint32_t>(cb));
                            if constexpr (!BRES) umma_commit(kStaticB ? empty0 + 8 * tap : empty0 + 8 * sb);
                        }
                        __syncwarp();
                        if constexpr (!kStaticB) {
                            if (++sb == b_stages) sb = 0, bph ^= 1;
                        }
                    }
                    }
                    if constexpr (SRING) bph ^= 1;
                    if (leader) umma_commit(aempty0 + 8 * sa);
                    __syncwarp();
                    if (++sa == a_slots) sa = 0, aph ^= 1;
                }
                if (leader) umma_commit(tfull0 + 8 * slot);
                __syncwarp();
                if (++slot == kSlots) slot = 0, tph ^= 1;
            }
            if (leader) trace_stamp(g, tl, 5);
            first = false;
        }
    } else if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        const bool leader = elect_one();
        const uint32_t smem_base = smem_u32(smem);
        const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
        const uint32_t tx_bytes = ((dbg_flags(g) & 4) ? 0 : Cfg::kATile) + ((dbg_flags(g) & 8) ? 0 : Cfg::kBTile);
        uint32_t s = 0, ph = 0;
        int tl = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
            const TileCoord tc = decode_tile(g, t);
            if (leader) trace_stamp(g, tl, 0);
            long long waited = 0;
            const int x0 = tc.tx * g.tw * g.stride - g.pad;
            const int y0 = tc.ty * g.th * g.stride - g.pad;
            const int ncol0 = tc.nt * BLOCK_N;
            for (int step = 0; step < g.n_steps; ++step) {
                for (int grp = 0; grp < G; ++grp) {
                    const int brow = g.b_row_base[grp] + g.step_row_base[step] + grp * g.step_group_stride[step] + ncol0;
                    const int a_c0 = g.a_c_base[grp];
                    int ky = 0, kx = 0;                 // tap coordinates of the stage's first tap
                    int b_k0 = g.b_k_base[grp];         // its K coordinate in B
                    for (int tap0 = 0; tap0 < g.taps; tap0 += TPS) {
                        for (int cb = 0; cb < g.cblocks; ++cb) {
                            if (kIgemmDebug && g.trace != nullptr) {
                                const long long w0 = clock64();
                                mbar_wait(empty0 + 8 * s, ph ^ 1);
                                waited += clock64() - w0;
                            } else {
                                mbar_wait(empty0 + 8 * s, ph ^ 1);
                            }
                            if (leader) {
                                const uint32_t fb = full0 + 8 * s;
                                mbar_expect_tx(fb, tx_bytes);
                                const uint32_t st = smem_base + s * Cfg::kStageBytes;
                                int kyi = ky, kxi = kx, bki = b_k0 + cb * BK;
#pragma unroll
                                for (int i = 0; i < TPS; ++i) {
                                    if (!(dbg_flags(g) & 4))
                                        tma_load_4d(st + i * Cfg::kASub, &tmA, fb, a_c0 + cb * BK, x0 + kxi, y0 + kyi,
                                                    tc.img);
                                    if (!(dbg_flags(g) & 8))
                                        tma_load_2d(st + Cfg::kATile + i * Cfg::kBSub, &tmB, fb, bki, brow);
                                    bki += g.b_k_tap_stride;
                                    if (++kxi == g.taps_w) kxi = 0, ++kyi;
                                }
                            }
                            if (++s == kStages) s = 0, ph ^= 1;
                        }
#pragma unroll
                        for (int i = 0; i < TPS; ++i) {
                            b_k0 += g.b_k_tap_stride;
                            if (++kx == g.taps_w) kx = 0, ++ky;
                        }
                    }
                }
            }
            if (leader) {
                trace_stamp(g, tl, 1);
                if (kIgemmDebug && g.trace != nullptr && tl < kTraceTiles)
                    g.trace[(static_cast<long long>(blockIdx.x) * kTraceTiles + tl) * 16 + 14] = waited;
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        const bool leader = elect_one();
        const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
        const uint32_t tfull0 = smem_u32(&tfull_bar[0]), tempty0 = smem_u32(&tempty_bar[0]);
        // descriptors of stage 0; stage s / tap i / K step k only add to the 14-bit (address >> 4) field
        const uint64_t a_desc0 = umma_smem_desc(smem_u32(smem), BK);
        const uint64_t b_desc0 = umma_smem_desc(smem_u32(smem) + Cfg::kATile, BK);
        uint32_t s = 0, ph = 0, slot = 0, aph = 0;
        int tl = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
            if (leader) trace_stamp(g, tl, 2);
            long long waited = 0;
            for (int sg = 0; sg < g.n_steps * G; ++sg) {
                mbar_wait(tempty0 + 8 * slot, aph ^ 1);
                tcgen05_fence_after();
                if (leader && sg == 0) trace_stamp(g, tl, 3);
                const uint32_t d_tmem = tmem_base + slot * BLOCK_N;
                for (int kb = 0; kb < kblocks_per_group; ++kb) {
                    if (kIgemmDebug && g.trace != nullptr) {
                        const long long w0 = clock64();
                        mbar_wait(full0 + 8 * s, ph);
                        waited += clock64() - w0;
                    } else {
                        mbar_wait(full0 + 8 * s, ph);
                    }
                    tcgen05_fence_after();
                    if (leader) {
                        if (sg == 0 && kb == 0) trace_stamp(g, tl, 4);
                        if (!(dbg_flags(g) & 2)) {
                            const uint64_t soff = static_cast<uint64_t>(s * (Cfg::kStageBytes >> 4));
#pragma unroll
                            for (int i = 0; i < TPS; ++i) {
#pragma unroll
                                for (int k = 0; k < BK / 32; ++k) {
                                    // +32 bytes of K inside the swizzle row = +2 in the (addr >> 4) field
                                    umma_i8(d_tmem, a_desc0 + soff + (i * (Cfg::kASub >> 4) + 2 * k),
                                            b_desc0 + soff + (i * (Cfg::kBSub >> 4) + 2 * k), g.idesc,
                                            (kb | i | k) != 0);
                                }
                            }
                        }
                        umma_commit(empty0 + 8 * s);
                    }
                    __syncwarp();
                    if (++s == kStages) s = 0, ph ^= 1;
                }
                if (leader) umma_commit(tfull0 + 8 * slot);
                __syncwarp();
                if (++slot == kSlots) slot = 0, aph ^= 1;
            }
            if (leader) {
                trace_stamp(g, tl, 5);
                if (kIgemmDebug && g.trace != nullptr && tl < kTraceTiles)
                    g.trace[(static_cast<long long>(blockIdx.x) * kTraceTiles + tl) * 16 + 15] = waited;
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ------------------------------------------------------------ side inputs, ahead of the epilogue
        // Warp 2 serves the even tiles of this CTA into slot 0, warp 3 the odd tiles into slot 1: each has two tile
        // periods for one tile's work, and each owns a private halo buffer.
        if constexpr (Epi::kSideWarp) {
            const uint32_t par = warp - 2;
            uint8_t* slot = epi_scratch + par * (kEpiSmemBytes / 2);
            int32_t* halo = reinterpret_cast<int32_t*>(epi_scratch + kEpiSmemBytes + par * (kHaloSmemBytes / 2));
            int staged_nt = -1;
            uint32_t sph = 0;
            typename Epi::Side sd;
            epi.side_init(sd, g, lane);
            for (int t = blockIdx.x + par * gridDim.x; t < total_tiles; t += 2 * gridDim.x, sph ^= 1) {
                const TileCoord tc = decode_tile(g, t);
                if constexpr (EpiHasSidePre<Epi>::value) {
                    // the tile's sums are formed (global fetch of the halo + shared-memory taps) while the epilogue
                    // may still be reading the slot; only the final stores wait for it
                    typename Epi::SidePre pre;
                    epi.side_prefetch(g, tc, lane, halo, sd, pre);
                    mbar_wait(smem_u32(&side_empty[par]), sph ^ 1);
                    epi.side_store(g, tc, lane, slot, staged_nt, sd, pre);
                } else {
                    mbar_wait(smem_u32(&side_empty[par]), sph ^ 1);
                    epi.side_load(g, tc, lane, slot, halo, staged_nt, sd);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&side_full[par]));     // release: the slot's writes are visible
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue
        const int quad = warp & 3;            // TMEM lane quadrant this warp may access
        const int part = ((warp - 4) >> 2) % kColSplit;     // which part of the step's columns (kColSplit parts)
        const int tgroup = ((warp - 4) >> 2) / kColSplit;   // which tiles of the CTA's sequence (kTileSplit groups)
        const int row = quad * 32 + lane;     // tile row == TMEM lane
        constexpr int kColsPerWarp = BLOCK_N / kColSplit;
        uint32_t ac = static_cast<uint32_t>(tgroup) * g.n_steps * G;     // accumulator index of this group's first tile
        uint32_t tile_par = tgroup;           // kTileSplit == 2: the group's side slot, fixed
        const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
        uint32_t sph = 0;
        int tl = tgroup;
        const bool tracer = (lane == 0 && quad == 0);
        for (int t = blockIdx.x + tgroup * gridDim.x; t < total_tiles; t += kTileSplit * gridDim.x, tl += kTileSplit) {
            const TileCoord tc = decode_tile(g, t);
            if (tracer && !(part == 1 && (dbg_flags(g) & 32))) trace_stamp(g, tl, part == 0 ? 6 : 11);
            typename Epi::Tile ts;
            const uint8_t* slot = epi_scratch + (Epi::kSideWarp ? tile_par * (kEpiSmemBytes / 2) : 0);
            if constexpr (Epi::kSideWarp) mbar_wait(smem_u32(&side_full[tile_par]), sph);
            epi.begin(ts, g, tc, row, slot);
            if (tracer && part == 0) trace_stamp(g, tl, 7);
            for (int step = 0; step < g.n_steps; ++step, ac += G) {
                const int c_begin = part * kColsPerWarp, c_end = (part + 1) * kColsPerWarp;
                if constexpr (Epi::kSeqDrain) {
                    // Accumulator groups are drained in order as soon as each is complete: group g's TMEM slot is
                    // released right after it has been folded into the fp32 running sums (registers), so the MMAs
                    // of the following groups / tiles never wait for a whole tile's epilogue.
                    static_assert(kColsPerWarp % 32 == 0, "sequential drain works on 32-column pairs of chunks");
                    float vsum[kColsPerWarp / 16][16];
#pragma unroll
                    for (int grp = 0; grp < G; ++grp) {
                        const uint32_t a = ac + grp;
                        mbar_wait(smem_u32(&tfull_bar[a % kSlots]), (a / kSlots) & 1);
                        tcgen05_fence_after();
                        if (tracer && step == 0 && grp == 0) trace_stamp(g, tl, part == 0 ? 8 : 12);
                        const uint32_t tbase = tmem_base + lane_base + (a % kSlots) * BLOCK_N + c_begin;
                        uint32_t acc_a[16], acc_b[16];
                        tmem_ld_x16(tbase, acc_a);
#pragma unroll
                        for (int c0 = 0; c0 < kColsPerWarp; c0 += 32) {
                            tmem_ld_wait();
                            tmem_ld_x16(tbase + c0 + 16, acc_b);
                            if (!(dbg_flags(g) & 1)) {
                                const int n0 = tc.nt * BLOCK_N + c_begin + c0;
                                epi.template accum<16>(ts, g, grp, n0, reinterpret_cast<const int32_t(&)[16]>(acc_a),
                                                       vsum[c0 / 16]);
                                if (grp == G - 1) epi.template finish<16>(ts, n0, vsum[c0 / 16]);
                            }
                            tmem_ld_wait();
                            if (c0 + 32 < kColsPerWarp) tmem_ld_x16(tbase + c0 + 32, acc_a);
                            if (!(dbg_flags(g) & 1)) {
                                const int n0 = tc.nt * BLOCK_N + c_begin + c0 + 16;
                                epi.template accum<16>(ts, g, grp, n0, reinterpret_cast<const int32_t(&)[16]>(acc_b),
                                                       vsum[c0 / 16 + 1]);
                                if (grp == G - 1) epi.template finish<16>(ts, n0, vsum[c0 / 16 + 1]);
                            }
                        }
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[a % kSlots]));
                    }
                    if (tracer && part == 0 && step == 0) trace_stamp(g, tl, 9);
                    epi.step_end(ts, g, tc, step, part, quad, lane, epi_scratch, TmemView{});
                    continue;
                }
#pragma unroll
                for (int grp = 0; grp < G; ++grp) {
                    const uint32_t a = ac + grp;
                    mbar_wait(smem_u32(&tfull_bar[a % kSlots]), (a / kSlots) & 1);
                }
                tcgen05_fence_after();
                if (tracer && step == 0 && !(part == 1 && (dbg_flags(g) & 32))) trace_stamp(g, tl, part == 0 ? 8 : 12);
                if (tracer && part == 0 && step == 1 && (dbg_flags(g) & 32)) trace_stamp(g, tl, 12);
                if constexpr (EpiStaticCols<Epi>::value) {
                    // Per-column parameters are kernel-parameter constants: make the warp's column base a compile-time
                    // value (one code copy per (column tile, part); a CTA only ever runs its own two), so that after
                    // unrolling every parameter is an immediate constant-bank operand -- no load instruction at all.
                    static_assert(kColsPerWarp % 32 == 0, "static columns work on 32-column pairs of chunks");
                    auto cols = [&](auto col0_c) {
                        constexpr int kCol0 = decltype(col0_c)::value;          // global column of the warp's first one
                        constexpr int kLoc0 = kCol0 % BLOCK_N;                  // ... and its place inside the tile
                        if constexpr (G == 1) {
                            const uint32_t tbase = tmem_base + lane_base + (ac % kSlots) * BLOCK_N + kLoc0;
                            uint32_t acc_a[1][16], acc_b[1][16];
                            tmem_ld_x16(tbase, acc_a[0]);
#pragma unroll
                            for (int c = 0; c < kColsPerWarp; c += 32) {
                                tmem_ld_wait();
                                tmem_ld_x16(tbase + c + 16, acc_b[0]);
                                epi.template chunk<16>(ts, g, tc, step, kCol0 + c,
                                                       reinterpret_cast<const int32_t(*)[16]>(acc_a));
                                tmem_ld_wait();
                                if (c + 32 < kColsPerWarp) tmem_ld_x16(tbase + c + 32, acc_a[0]);
                                epi.template chunk<16>(ts, g, tc, step, kCol0 + c + 16,
                                                       reinterpret_cast<const int32_t(*)[16]>(acc_b));
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < kColsPerWarp; c += 16) {
                                uint32_t acc[G][16];
#pragma unroll
                                for (int grp = 0; grp < G; ++grp) {
                                    const uint32_t sl = (ac + grp) % kSlots;
                                    tmem_ld_x16(tmem_base + lane_base + sl * BLOCK_N + kLoc0 + c, acc[grp]);
                                }
                                tmem_ld_wait();
                                epi.template chunk<16>(ts, g, tc, step, kCol0 + c,
                                                       reinterpret_cast<const int32_t(*)[16]>(acc));
                            }
                        }
                    };
                    dispatch_col<0, 256 / kColsPerWarp, kColsPerWarp>(tc.nt * BLOCK_N + c_begin, cols);
                } else if constexpr (G == 1 && (kColsPerWarp % 32 == 0)) {
                    // software-pipelined TMEM reads: the load of chunk i+1 is in flight while chunk i is processed
                    const uint32_t tbase = tmem_base + lane_base + (ac % kSlots) * BLOCK_N;
                    uint32_t acc_a[1][16], acc_b[1][16];
                    tmem_ld_x16(tbase + c_begin, acc_a[0]);
                    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                        tmem_ld_wait();
                        tmem_ld_x16(tbase + c0 + 16, acc_b[0]);
                        if (!(dbg_flags(g) & 1))
                            epi.template chunk<16>(ts, g, tc, step, tc.nt * BLOCK_N + c0,
                                                   reinterpret_cast<const int32_t(*)[16]>(acc_a));
                        tmem_ld_wait();
                        if (c0 + 32 < c_end) tmem_ld_x16(tbase + c0 + 32, acc_a[0]);
                        if (!(dbg_flags(g) & 1))
                            epi.template chunk<16>(ts, g, tc, step, tc.nt * BLOCK_N + c0 + 16,
                                                   reinterpret_cast<const int32_t(*)[16]>(acc_b));
                    }
                } else {
                    for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                        uint32_t acc[G][16];
#pragma unroll
                        for (int grp = 0; grp < G; ++grp) {
                            const uint32_t slot = (ac + grp) % kSlots;
                            tmem_ld_x16(tmem_base + lane_base + slot * BLOCK_N + c0, acc[grp]);
                        }
                        tmem_ld_wait();
                        if (!(dbg_flags(g) & 1))
                            epi.template chunk<16>(ts, g, tc, step, tc.nt * BLOCK_N + c0,
                                                   reinterpret_cast<const int32_t(*)[16]>(acc));
                    }
                }
                if (tracer && part == 0 && step == 0) trace_stamp(g, tl, 9);
                if (tracer && part == 0 && step == 1 && (dbg_flags(g) & 32)) trace_stamp(g, tl, 13);
                TmemView tv{};
#pragma unroll
                for (int grp = 0; grp < G; ++grp)
                    tv.addr[grp] = tmem_base + lane_base + ((ac + grp) % kSlots) * BLOCK_N;
                tv.c_begin = c_begin;
                tv.c_end = c_end;
                if constexpr (Epi::kHoldSlots) epi.step_end(ts, g, tc, step, part, quad, lane, epi_scratch, tv);
                if (tracer && part == 0 && step == 0 && (dbg_flags(g) & 32)) trace_stamp(g, tl, 11);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int grp = 0; grp < G; ++grp) mbar_arrive(smem_u32(&tempty_bar[(ac + grp) % kSlots]));
                }
                if constexpr (!Epi::kHoldSlots) epi.step_end(ts, g, tc, step, part, quad, lane, epi_scratch, tv);
            }
            epi.end(ts, g, tc);
            if constexpr (Epi::kSideWarp) {
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&side_empty[tile_par]));
            }
            if constexpr (kTileSplit == 1) {
                if ((tile_par ^= 1) == 0) sph ^= 1;
            } else {
                sph ^= 1;                                      // the group's own side slot, one phase per own tile
                ac += (kTileSplit - 1) * g.n_steps * G;        // skip the other group's tiles
            }
            if (tracer && !(part == 1 && (dbg_flags(g) & 32))) trace_stamp(g, tl, part == 0 ? 10 : 13);
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace qv2x
