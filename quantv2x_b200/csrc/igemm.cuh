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
};

constexpr int kTraceTiles = 32;
__device__ __forceinline__ void trace_stamp(const IgemmGeom& g, int tile_local, int slot) {
    if (g.trace != nullptr && tile_local < kTraceTiles)
        g.trace[(static_cast<long long>(blockIdx.x) * kTraceTiles + tile_local) * 16 + slot] = clock64();
}

constexpr int kEpiSmemBytes = 10240;   // two side-input slots / epilogue scratch
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
    static constexpr int kStagesRaw = (227 * 1024 - 1024 - 256 - kEpiSmemBytes - kHaloSmemBytes) / kStageBytes;
    static constexpr int kStages = kStagesRaw > MAX_STAGES ? MAX_STAGES : kStagesRaw;
    static constexpr int kSlots = (512 / BLOCK_N) > 4 ? 4 : (512 / BLOCK_N);
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiSmemBytes + kHaloSmemBytes;
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
template <class Epi, int BLOCK_N>
constexpr int igemm_threads() { return (4 + 4 * Epi::col_split(BLOCK_N)) * 32; }

template <int BLOCK_N, int BK, int G, class Epi, int TPS = 1>
__global__ void __launch_bounds__(igemm_threads<Epi, BLOCK_N>(), 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const IgemmGeom g,
             const Epi epi) {
    using Cfg = IgemmCfg<BLOCK_N, BK, Epi::kMaxStages, TPS>;
    constexpr int kStages = Cfg::kStages;
    constexpr int kSlots = Cfg::kSlots;
    static_assert(G <= kSlots, "every group needs its own TMEM slot");
    constexpr int kColSplit = Epi::col_split(BLOCK_N);
    constexpr int kNumEpiWarps = 4 * kColSplit;
    static_assert(BLOCK_N % (16 * kColSplit) == 0, "BLOCK_N must split into 16-column-aligned parts");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tfull_bar = bars + 2 * kStages;
    uint64_t* tempty_bar = bars + 2 * kStages + kSlots;
    uint64_t* side_full = bars + 2 * kStages + 2 * kSlots;      // [2] side-input slots (warp 3 -> epilogue)
    uint64_t* side_empty = side_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(side_empty + 2);
    uint8_t* epi_scratch = smem + kStages * Cfg::kStageBytes + 256;

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
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_ptr), 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int total_tiles = g.n_img * g.tiles_y * g.tiles_x * g.n_tiles;
    const int kblocks_per_group = (g.taps / TPS) * g.cblocks;   // pipeline stages per accumulator group

    // The producer and MMA roles run their loops on the WHOLE warp (every lane computes the same, warp-uniform
    // values, so they live in uniform registers) and elect one lane only for the instructions that must be issued
    // once.  All addressing is incremental: a single thread's dependent integer chain (divisions, 64-bit
    // descriptor assembly) between two TMA / MMA instructions was what paced these roles before.
    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        const bool leader = elect_one();
        const uint32_t smem_base = smem_u32(smem);
        const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
        const uint32_t tx_bytes = ((g.debug & 4) ? 0 : Cfg::kATile) + ((g.debug & 8) ? 0 : Cfg::kBTile);
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
                            if (g.trace != nullptr) {
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
                                    if (!(g.debug & 4))
                                        tma_load_4d(st + i * Cfg::kASub, &tmA, fb, a_c0 + cb * BK, x0 + kxi, y0 + kyi,
                                                    tc.img);
                                    if (!(g.debug & 8))
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
                if (g.trace != nullptr && tl < kTraceTiles)
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
                    if (g.trace != nullptr) {
                        const long long w0 = clock64();
                        mbar_wait(full0 + 8 * s, ph);
                        waited += clock64() - w0;
                    } else {
                        mbar_wait(full0 + 8 * s, ph);
                    }
                    tcgen05_fence_after();
                    if (leader) {
                        if (sg == 0 && kb == 0) trace_stamp(g, tl, 4);
                        if (!(g.debug & 2)) {
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
                if (g.trace != nullptr && tl < kTraceTiles)
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
                mbar_wait(smem_u32(&side_empty[par]), sph ^ 1);
                epi.side_load(g, tc, lane, slot, halo, staged_nt, sd);
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&side_full[par]));     // release: the slot's writes are visible
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue
        const int quad = warp & 3;            // TMEM lane quadrant this warp may access
        const int part = (warp - 4) >> 2;     // which part of the step's columns (kColSplit parts)
        const int row = quad * 32 + lane;     // tile row == TMEM lane
        constexpr int kColsPerWarp = BLOCK_N / kColSplit;
        uint32_t ac = 0;
        uint32_t tile_par = 0;
        const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
        uint32_t sph = 0;
        int tl = 0;
        const bool tracer = (lane == 0 && quad == 0);
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl) {
            const TileCoord tc = decode_tile(g, t);
            if (tracer) trace_stamp(g, tl, part == 0 ? 6 : 11);
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
                            if (!(g.debug & 1)) {
                                const int n0 = tc.nt * BLOCK_N + c_begin + c0;
                                epi.template accum<16>(ts, g, grp, n0, reinterpret_cast<const int32_t(&)[16]>(acc_a),
                                                       vsum[c0 / 16]);
                                if (grp == G - 1) epi.template finish<16>(ts, n0, vsum[c0 / 16]);
                            }
                            tmem_ld_wait();
                            if (c0 + 32 < kColsPerWarp) tmem_ld_x16(tbase + c0 + 32, acc_a);
                            if (!(g.debug & 1)) {
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
                if (tracer && step == 0) trace_stamp(g, tl, part == 0 ? 8 : 12);
                if constexpr (G == 1 && (kColsPerWarp % 32 == 0)) {
                    // software-pipelined TMEM reads: the load of chunk i+1 is in flight while chunk i is processed
                    const uint32_t tbase = tmem_base + lane_base + (ac % kSlots) * BLOCK_N;
                    uint32_t acc_a[1][16], acc_b[1][16];
                    tmem_ld_x16(tbase + c_begin, acc_a[0]);
                    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                        tmem_ld_wait();
                        tmem_ld_x16(tbase + c0 + 16, acc_b[0]);
                        if (!(g.debug & 1))
                            epi.template chunk<16>(ts, g, tc, step, tc.nt * BLOCK_N + c0,
                                                   reinterpret_cast<const int32_t(*)[16]>(acc_a));
                        tmem_ld_wait();
                        if (c0 + 32 < c_end) tmem_ld_x16(tbase + c0 + 32, acc_a[0]);
                        if (!(g.debug & 1))
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
                        if (!(g.debug & 1))
                            epi.template chunk<16>(ts, g, tc, step, tc.nt * BLOCK_N + c0,
                                                   reinterpret_cast<const int32_t(*)[16]>(acc));
                    }
                }
                if (tracer && part == 0 && step == 0) trace_stamp(g, tl, 9);
                TmemView tv{};
#pragma unroll
                for (int grp = 0; grp < G; ++grp)
                    tv.addr[grp] = tmem_base + lane_base + ((ac + grp) % kSlots) * BLOCK_N;
                tv.c_begin = c_begin;
                tv.c_end = c_end;
                if constexpr (Epi::kHoldSlots) epi.step_end(ts, g, tc, step, part, quad, lane, epi_scratch, tv);
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
            if ((tile_par ^= 1) == 0) sph ^= 1;
            if (tracer) trace_stamp(g, tl, part == 0 ? 10 : 13);
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
