// Fixed-point requantization epilogue of the quantized conv / deconv layers (8-bit output, zero-point 0, ReLU).
//
// Reference semantics (opencood/quant/quant_layer.py:391-410 with :132-148): fake-quant weight -> conv (+bias) ->
// folded BN = identity -> ReLU -> fake-quant activation, i.e. q = clamp(rint(relu(conv + b) / delta_out), 0, 255).
// The conv result is an exact integer combination of the int32 accumulators, so the whole output step is one affine
// map per output column c, evaluated here in 64-bit fixed point instead of fp32 (oracle/int_oracle.py restates it bit
// for bit; it is accurate to 2^-31 relative, tighter than the fp32 chain it replaces, and needs ~4 instructions per
// output where the fp32 chain needed ~12 -- the requant epilogue, not the tensor pipe, paced every layer with
// K < 2000, profiles/r2_trace_halo_v2.log):
//     t_g  = acc_g[p,c] - zpw[c] * S_g[p]                                    int32, exact
//     G groups (conv; G = 3 is the concat input with three activation scales):
//         u = sum_g t_g * M_g[c] + C[c]                                      int64, exact
//         q = clamp(u >> sh[c], 0, 255)                                      arithmetic shift = floor
//       M_g[c] = rint(r_g * 2^sh),  r_g = gs_g * cs[c] / delta_out (double),  sh in [32, 47] so that max_g M_g < 2^31
//       C[c]   = rint(bias[c] / delta_out * 2^sh) + 2^(sh-1)                 (round half up)
//     DIGITS (transposed convs: 24-bit fixed-point weights as three signed byte digits, K <= 256):
//         w = 256 * acc_mid + acc_lo   (int32, exact)        u = acc_hi * M + ((w * M) >> 16) + C'
//         q = clamp(u >> (sh - 16), 0, 255),  sh in [48, 62],  C' = rint(bias / delta_out * 2^(sh-16)) + 2^(sh-17)
// The host (qv2x_layer_create) derives M, C in double precision; layers whose parameters do not fit (r >= 0.5,
// |bias / delta_out| >= 2^15) keep the fp32 epilogue -- the oracle applies the same rule.
#pragma once
#include "epilogue_requant.cuh"

namespace qv2x {

// d = bytes {sat_u8(x0), sat_u8(x1), sat_u8(x2), sat_u8(x3)} (x0 in the low byte): two I2IP instructions
__device__ __forceinline__ uint32_t pack4_sat_u8(int x0, int x1, int x2, int x3) {
    uint32_t hi, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(x3), "r"(x2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x1), "r"(x0), "r"(hi));
    return d;
}

__device__ __forceinline__ int2 lds_i2(uint32_t saddr) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}

// ZP: the weights carry a per-column zero-point (8-bit weights); DUMP: test hook that writes the zero-point-corrected
// accumulators (a separate instantiation, so that the branch and its address arithmetic stay out of the hot kernels).
template <int G, bool DIGITS = false, bool ZP = true, bool DUMP = false>
struct FixedEpilogue {
    // Transposed convs (DIGITS) do almost no tensor work per tile (K <= 256) and three accumulator groups of epilogue:
    // four warps per lane quadrant hide the latencies of that ALU-bound code better than two.
    static constexpr int col_split(int) { return 2; }          // (four warps per quadrant measured slower for DIGITS: 38.7 vs 34.4 us)
    static constexpr int kMaxStages = 8;
    static constexpr bool kSideWarp = true;
    static constexpr bool kSeqDrain = false;       // all groups of a tile stay in TMEM until the tile is requantized
    static constexpr bool kHoldSlots = false;
    // output addressing as RequantEpilogue: pixel (oy*up + dy, ox*up + dx) of an [n_img, Hout, Wout, out_cstride] tensor
    int up, cout_sub, Hout, Wout, out_cstride, out_cbase;
    int up_shift;
    int debug;
    FastDiv fd_cout_sub;
    const int32_t* mul[kMaxGroups];       // [N_total] M_g (DIGITS: mul[0] only)
    const int2* cadd;                     // [N_total] C as {low word, high word}
    const int32_t* shr;                   // [N_total] sh - 32 (DIGITS: sh - 48): shift of the high word of u
    const int32_t* zpw;                   // [N_total] (ZP only)
    const int32_t* rowsum_in[kMaxGroups];
    uint8_t* out;
    int32_t* rowsum_out;
    int32_t* acc_dump;                    // DUMP only (DIGITS: dumped in hi, mid, lo order)
    int n_total;

    struct Tile {
        int32_t negS[G];         // -S_g of this row
        long long opix, mrow;
        int rsum;
        uint32_t sm_par;
        int n_base, ch_off;
    };
    // Side slot (kEpiSmemBytes / 2 = 6 KB): arrays of kPar columns: M_0 .. M_{G-1} (int32) | C (int2) | shift | zpw,
    // then S[G][128].  G = 1 tiles are up to 256 columns wide, G = 3 tiles up to 128.
    static constexpr int kPar = (G == 1) ? 256 : 128;
    static constexpr int kOffC = G * kPar * 4;
    static constexpr int kOffSh = kOffC + kPar * 8;
    static constexpr int kOffZ = kOffSh + kPar * 4;
    static constexpr int kSlotS = kOffZ + kPar * 4;
    static_assert(kSlotS + G * kTileM * 4 <= kEpiSmemBytes / 2, "side slot");
    static constexpr int kHaloInts = SideHalo::kHaloInts;
    using Side = SideHalo;

    __device__ __forceinline__ void side_init(Side& sd, const IgemmGeom& g, int lane) const { sd.init(g, lane); }

    __device__ __forceinline__ void side_load(const IgemmGeom& g, const TileCoord& tc, int lane, uint8_t* slot,
                                              int32_t* halo, int& staged_nt, const Side& sd) const {
        if (staged_nt != tc.nt) {
            const uint32_t sp = smem_u32(slot);
            const int n_base = tc.nt * g.block_n;
            for (int i = lane; i < g.block_n; i += 32) {
#pragma unroll
                for (int q = 0; q < (DIGITS ? 1 : G); ++q) cp_async_4(sp + 4 * (q * kPar + i), mul[q] + n_base + i, true);
                cp_async_4(sp + kOffC + 8 * i, &cadd[n_base + i].x, true);
                cp_async_4(sp + kOffC + 8 * i + 4, &cadd[n_base + i].y, true);
                cp_async_4(sp + kOffSh + 4 * i, shr + n_base + i, true);
                if constexpr (ZP) cp_async_4(sp + kOffZ + 4 * i, zpw + n_base + i, true);
            }
            staged_nt = tc.nt;
        }
        sd.template stage<G>(g, tc, lane, reinterpret_cast<int32_t*>(slot + kSlotS), halo, rowsum_in);
        cp_async_wait_all();
    }

    // Two-phase form (igemm.cuh uses it when present): the receptive-field sums of the tile are computed into
    // registers BEFORE the slot is free -- they only need the warp's private halo buffer -- and stored afterwards.
    struct SidePre {
        int32_t sums[G][4];
    };
    __device__ __forceinline__ void side_prefetch(const IgemmGeom& g, const TileCoord& tc, int lane, int32_t* halo,
                                                  const Side& sd, SidePre& pre) const {
        sd.template compute<G>(g, tc, lane, halo, rowsum_in, pre.sums);
    }
    __device__ __forceinline__ void side_store(const IgemmGeom& g, const TileCoord& tc, int lane, uint8_t* slot,
                                               int& staged_nt, const Side& sd, const SidePre& pre) const {
        if (staged_nt != tc.nt) {
            const uint32_t sp = smem_u32(slot);
            const int n_base = tc.nt * g.block_n;
            for (int i = lane; i < g.block_n; i += 32) {
#pragma unroll
                for (int q = 0; q < (DIGITS ? 1 : G); ++q) cp_async_4(sp + 4 * (q * kPar + i), mul[q] + n_base + i, true);
                cp_async_4(sp + kOffC + 8 * i, &cadd[n_base + i].x, true);
                cp_async_4(sp + kOffC + 8 * i + 4, &cadd[n_base + i].y, true);
                cp_async_4(sp + kOffSh + 4 * i, shr + n_base + i, true);
                if constexpr (ZP) cp_async_4(sp + kOffZ + 4 * i, zpw + n_base + i, true);
            }
            staged_nt = tc.nt;
        }
        sd.template store<G>(lane, reinterpret_cast<int32_t*>(slot + kSlotS), pre.sums);
        cp_async_wait_all();
    }

    __device__ __forceinline__ void begin(Tile& ts, const IgemmGeom& g, const TileCoord& tc, int row,
                                          const uint8_t* slot) const {
        ts.sm_par = smem_u32(slot);
        ts.n_base = tc.nt * g.block_n;
        ts.ch_off = (up > 1) ? fd_cout_sub.div(ts.n_base) * cout_sub : 0;
        const int lx = row & (g.tw - 1), ly = row >> g.tw_shift;
        const int ox = tc.tx * g.tw + lx, oy = tc.ty * g.th + ly;
        const bool valid = (ox < g.Wo) && (oy < g.Ho);
        ts.rsum = 0;
        ts.opix = -1;
        ts.mrow = -1;
#pragma unroll
        for (int grp = 0; grp < G; ++grp)
            ts.negS[grp] = -*reinterpret_cast<const int32_t*>(slot + kSlotS + 4 * (grp * kTileM + row));
        if (!valid) return;
        ts.mrow = (static_cast<long long>(tc.img) * g.Ho + oy) * g.Wo + ox;
        int dy = 0, dx = 0;
        if (up > 1) {
            const int sub = fd_cout_sub.div(ts.n_base);
            dy = sub >> up_shift;
            dx = sub - (dy << up_shift);
        }
        ts.opix = (static_cast<long long>(tc.img) * Hout + oy * up + dy) * Wout + ox * up + dx;
    }

    template <int W>
    __device__ __forceinline__ void load_i32(uint32_t saddr, int32_t (&v)[W]) const {
#pragma unroll
        for (int v4 = 0; v4 < W / 4; ++v4) {
            const int4 z = lds_i4(saddr + 16 * v4);
            v[4 * v4 + 0] = z.x, v[4 * v4 + 1] = z.y, v[4 * v4 + 2] = z.z, v[4 * v4 + 3] = z.w;
        }
    }
    // C of W columns as ready-made 64-bit operands (even / odd register pairs straight out of LDS.128)
    template <int W>
    __device__ __forceinline__ void load_c(uint32_t saddr, long long (&c)[W]) const {
#pragma unroll
        for (int v2 = 0; v2 < W / 2; ++v2) {
            const int4 z = lds_i4(saddr + 16 * v2);
            asm("mov.b64 %0, {%1, %2};" : "=l"(c[2 * v2]) : "r"(z.x), "r"(z.y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(c[2 * v2 + 1]) : "r"(z.z), "r"(z.w));
        }
    }
    // high word of a * b + c (one IMAD.HI: its addend is a 64-bit register pair)
    static __device__ __forceinline__ int32_t mad_hi64(int32_t a, int32_t b, long long c) {
        return static_cast<int32_t>((static_cast<long long>(a) * b + c) >> 32);
    }

    // W consecutive columns [n0, n0 + W) of this thread's row, all G accumulator groups at once.
    template <int W>
    __device__ __forceinline__ void chunk(Tile& ts, const IgemmGeom& g, const TileCoord& tc, int step, int n0,
                                          const int32_t (*acc)[W]) const {
        (void)tc;
        (void)step;
        static_assert(W == 16, "chunk width");
        const int nl = n0 - ts.n_base;
        long long c[W];
        int32_t sh[W];
        load_c<W>(ts.sm_par + kOffC + 8 * nl, c);
        load_i32<W>(ts.sm_par + kOffSh + 4 * nl, sh);
        int q[W];
        if constexpr (DIGITS) {
            // groups arrive in the order mid, lo, hi (see qv2x_layer_forward)
            if constexpr (DUMP) {
                if (acc_dump != nullptr && ts.mrow >= 0) {
                    const long long gstride = static_cast<long long>(g.n_img) * g.Ho * g.Wo;
#pragma unroll
                    for (int grp = 0; grp < 3; ++grp) {
                        const int dg = (grp == 0) ? 1 : (grp == 1 ? 2 : 0);
                        int32_t* dp = acc_dump + (dg * gstride + ts.mrow) * n_total + n0;
#pragma unroll
                        for (int j = 0; j < W; j += 4)
                            st_global_v4(dp + j, acc[grp][j], acc[grp][j + 1], acc[grp][j + 2], acc[grp][j + 3]);
                    }
                }
            }
            int32_t m[W];
            load_i32<W>(ts.sm_par + 4 * nl, m);
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const int32_t w = acc[0][j] * 256 + acc[1][j];
                const long long b = static_cast<long long>(w) * m[j];
                q[j] = mad_hi64(acc[2][j], m[j], (b >> 16) + c[j]) >> sh[j];
            }
        } else {
            int32_t t[G][W];
            if constexpr (ZP) {
                int32_t zw[W];
                load_i32<W>(ts.sm_par + kOffZ + 4 * nl, zw);
#pragma unroll
                for (int grp = 0; grp < G; ++grp)
#pragma unroll
                    for (int j = 0; j < W; ++j) t[grp][j] = zw[j] * ts.negS[grp] + acc[grp][j];
            } else {
#pragma unroll
                for (int grp = 0; grp < G; ++grp)
#pragma unroll
                    for (int j = 0; j < W; ++j) t[grp][j] = acc[grp][j];
            }
            if constexpr (DUMP) {
                if (acc_dump != nullptr && ts.mrow >= 0) {
                    const long long gstride = static_cast<long long>(g.n_img) * g.Ho * g.Wo;
#pragma unroll
                    for (int grp = 0; grp < G; ++grp) {
                        int32_t* dp = acc_dump + (grp * gstride + ts.mrow) * n_total + n0;
#pragma unroll
                        for (int j = 0; j < W; j += 4)
                            st_global_v4(dp + j, t[grp][j], t[grp][j + 1], t[grp][j + 2], t[grp][j + 3]);
                    }
                }
            }
#pragma unroll
            for (int grp = 0; grp < G - 1; ++grp) {
                int32_t m[W];
                load_i32<W>(ts.sm_par + 4 * (grp * kPar + nl), m);
#pragma unroll
                for (int j = 0; j < W; ++j) c[j] += static_cast<long long>(t[grp][j]) * m[j];
            }
            int32_t m[W];
            load_i32<W>(ts.sm_par + 4 * ((G - 1) * kPar + nl), m);
#pragma unroll
            for (int j = 0; j < W; ++j) q[j] = mad_hi64(t[G - 1][j], m[j], c[j]) >> sh[j];
        }
        uint32_t packed[W / 4];
        unsigned rsum = 0;
#pragma unroll
        for (int w = 0; w < W / 4; ++w) {
            packed[w] = pack4_sat_u8(q[4 * w], q[4 * w + 1], q[4 * w + 2], q[4 * w + 3]);
            rsum = __dp4a(packed[w], 0x01010101u, rsum);
        }
        if (ts.opix >= 0) {
            st_global_v4(out + ts.opix * out_cstride + out_cbase + (n0 - ts.ch_off), packed[0], packed[1], packed[2],
                         packed[3]);
            ts.rsum += static_cast<int>(rsum);
        }
    }

    __device__ __forceinline__ void step_end(Tile&, const IgemmGeom&, const TileCoord&, int, int, int, int, uint8_t*,
                                             const TmemView&) const {}

    __device__ __forceinline__ void end(Tile& ts, const IgemmGeom& g, const TileCoord& tc) const {
        (void)g;
        (void)tc;
        if (rowsum_out != nullptr && ts.opix >= 0) atomicAdd(rowsum_out + ts.opix, ts.rsum);
    }
};

// ------------------------------------------------------------------------------------------------------------
// The same requantizer with its per-column parameters in the KERNEL-PARAMETER constant bank instead of shared memory.
// While tcgen05.mma streams operands (128 B/clk for BLOCK_N <= 128) ordinary shared-memory loads are starved -- one
// LDS wavefront takes ~74 cycles (profiles/r2_exp_smem_contention.log) -- and the ~20 parameter loads per 16-column chunk
// were what paced every epilogue.  The column index is warp-uniform, so the constant cache serves it as a broadcast and
// the shared-memory port is left to the tensor core.  Layers of up to 256 output columns (every conv of the model).
template <int G>
struct FixedTable {
    int32_t mul[G][256];
    int2 cadd[256];
    int32_t shr[256];
    int32_t zw[256];
};

template <int G, bool ZP = true>
struct FixedEpilogueC : FixedEpilogue<G, false, ZP, false> {
    using Base = FixedEpilogue<G, false, ZP, false>;
    using Tile = typename Base::Tile;
    using Side = typename Base::Side;
    static constexpr bool kStaticCols = true;      // igemm.cuh: the column base of every chunk is a compile-time value
    // 128-column tiles of short-K layers (stage 1, K = 1152) are epilogue-bound: four warps per lane quadrant halve
    // the chunk loop of a tile (role traces: 1.9 k of a 2.9 k-cycle tile period)
    static constexpr int col_split(int) { return 2; }
    // short-K layers (stages 0 and 1) are bound by the per-tile fixed costs of the epilogue warps: two groups of
    // eight warps alternate over the tiles (igemm.cuh, EpiTileSplit)
    static constexpr int tile_split(int block_n) { return block_n <= 128 ? 2 : 1; }
    FixedTable<G> tab;

    __device__ __forceinline__ void side_load(const IgemmGeom& g, const TileCoord& tc, int lane, uint8_t* slot,
                                              int32_t* halo, int& staged_nt, const Side& sd) const {
        (void)staged_nt;
        sd.template stage<G>(g, tc, lane, reinterpret_cast<int32_t*>(slot + Base::kSlotS), halo, this->rowsum_in);
    }
    __device__ __forceinline__ void side_store(const IgemmGeom&, const TileCoord&, int lane, uint8_t* slot,
                                               int& staged_nt, const Side& sd,
                                               const typename Base::SidePre& pre) const {
        (void)staged_nt;
        sd.template store<G>(lane, reinterpret_cast<int32_t*>(slot + Base::kSlotS), pre.sums);
    }

    template <int W>
    __device__ __forceinline__ void chunk(Tile& ts, const IgemmGeom& g, const TileCoord& tc, int step, int n0,
                                          const int32_t (*acc)[W]) const {
        (void)g;
        (void)tc;
        (void)step;
        static_assert(W == 16, "chunk width");
        int q[W];
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const int n = n0 + j;
            long long u;
            asm("mov.b64 %0, {%1, %2};" : "=l"(u) : "r"(tab.cadd[n].x), "r"(tab.cadd[n].y));
#pragma unroll
            for (int grp = 0; grp < G; ++grp) {
                const int32_t t = ZP ? tab.zw[n] * ts.negS[grp] + acc[grp][j] : acc[grp][j];
                if (grp < G - 1) u += static_cast<long long>(t) * tab.mul[grp][n];
                else q[j] = Base::mad_hi64(t, tab.mul[grp][n], u) >> tab.shr[n];
            }
        }
        uint32_t packed[W / 4];
        unsigned rsum = 0;
#pragma unroll
        for (int w = 0; w < W / 4; ++w) {
            packed[w] = pack4_sat_u8(q[4 * w], q[4 * w + 1], q[4 * w + 2], q[4 * w + 3]);
            rsum = __dp4a(packed[w], 0x01010101u, rsum);
        }
        if (ts.opix >= 0) {
            st_global_v4(this->out + ts.opix * this->out_cstride + this->out_cbase + (n0 - ts.ch_off), packed[0],
                         packed[1], packed[2], packed[3]);
            ts.rsum += static_cast<int>(rsum);
        }
    }
};

}  // namespace qv2x
