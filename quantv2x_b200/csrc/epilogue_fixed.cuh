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
//       C[c]   = rint(bias[c] / delta_out * 2^sh) + 2^(sh-1)                 (round half up), low 4 bits := sh - 32
//     DIGITS (transposed convs: 24-bit fixed-point weights as three signed byte digits, K <= 256):
//         w = 256 * acc_mid + acc_lo   (int32, exact)        u = acc_hi * M + ((w * M) >> 16) + C'
//         q = clamp(u >> (sh - 16), 0, 255),  sh in [48, 62],  C' = rint(bias / delta_out * 2^(sh-16)) + 2^(sh-17),
//         low 4 bits := sh - 48
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

template <int G, bool DIGITS = false>
struct FixedEpilogue {
    static constexpr int col_split(int) { return 2; }
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
    const int32_t* c_lo;                  // [N_total] low / high words of C
    const int32_t* c_hi;
    const int32_t* zpw;                   // [N_total] or nullptr (weights already zero-centred; always for DIGITS)
    const int32_t* rowsum_in[kMaxGroups];
    uint8_t* out;
    int32_t* rowsum_out;
    int32_t* acc_dump;                    // as RequantEpilogue (DIGITS: dumped in hi, mid, lo order)
    int n_total;

    struct Tile {
        int32_t S[G];
        long long opix, mrow;
        int rsum;
        uint32_t sm_par;
        int n_base, ch_off;
    };
    // Side slot (kEpiSmemBytes / 2 = 5 KB): parameter arrays of kPar entries each: M_0 .. M_{G-1} | C_lo | C_hi | zpw,
    // then S[G][128].  G = 1 tiles are up to 256 columns wide, G = 3 tiles up to 128.
    static constexpr int kPar = (G == 1) ? 256 : 128;
    static constexpr int kNumArr = G + 3;
    static constexpr int kSlotS = kNumArr * kPar * 4;
    static_assert(kSlotS + G * kTileM * 4 <= kEpiSmemBytes / 2, "side slot");
    static constexpr int kHaloInts = SideHalo::kHaloInts;
    using Side = SideHalo;

    __device__ __forceinline__ void side_init(Side& sd, const IgemmGeom& g, int lane) const { sd.init(g, lane); }

    __device__ __forceinline__ void side_load(const IgemmGeom& g, const TileCoord& tc, int lane, uint8_t* slot,
                                              int32_t* halo, int& staged_nt, const Side& sd) const {
        if (staged_nt != tc.nt) {
            int32_t* s_p = reinterpret_cast<int32_t*>(slot);
            const int n_base = tc.nt * g.block_n;
            const bool has_zp = (zpw != nullptr);
            for (int i = lane; i < g.block_n; i += 32) {
#pragma unroll
                for (int q = 0; q < (DIGITS ? 1 : G); ++q) cp_async_4(smem_u32(s_p + q * kPar + i), mul[q] + n_base + i, true);
                cp_async_4(smem_u32(s_p + G * kPar + i), c_lo + n_base + i, true);
                cp_async_4(smem_u32(s_p + (G + 1) * kPar + i), c_hi + n_base + i, true);
                cp_async_4(smem_u32(s_p + (G + 2) * kPar + i), has_zp ? zpw + n_base + i : c_lo, has_zp);
            }
            staged_nt = tc.nt;
        }
        sd.template stage<G>(g, tc, lane, reinterpret_cast<int32_t*>(slot + kSlotS), halo, rowsum_in);
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
            ts.S[grp] = *reinterpret_cast<const int32_t*>(slot + kSlotS + 4 * (grp * kTileM + row));
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
    __device__ __forceinline__ void load_par(const Tile& ts, int arr, int nl, int32_t (&v)[W]) const {
#pragma unroll
        for (int v4 = 0; v4 < W / 4; ++v4) {
            const int4 z = lds_i4(ts.sm_par + 4 * (arr * kPar + nl) + 16 * v4);
            v[4 * v4 + 0] = z.x, v[4 * v4 + 1] = z.y, v[4 * v4 + 2] = z.z, v[4 * v4 + 3] = z.w;
        }
    }

    // W consecutive columns [n0, n0 + W) of this thread's row, all G accumulator groups at once.
    template <int W>
    __device__ __forceinline__ void chunk(Tile& ts, const IgemmGeom& g, const TileCoord& tc, int step, int n0,
                                          const int32_t (*acc)[W]) const {
        (void)tc;
        (void)step;
        static_assert(W == 16, "chunk width");
        const int nl = n0 - ts.n_base;
        int32_t clo[W], chi[W];
        load_par<W>(ts, G, nl, clo);
        load_par<W>(ts, G + 1, nl, chi);
        int q[W];
        if constexpr (DIGITS) {
            // groups arrive in the order mid, lo, hi (see qv2x_layer_forward)
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
            int32_t m[W];
            load_par<W>(ts, 0, nl, m);
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const int32_t w = acc[0][j] * 256 + acc[1][j];
                const long long b = static_cast<long long>(w) * m[j];
                const long long c = (static_cast<long long>(chi[j]) << 32) | static_cast<uint32_t>(clo[j]);
                const long long u = static_cast<long long>(acc[2][j]) * m[j] + ((b >> 16) + c);
                q[j] = static_cast<int32_t>(u >> 32) >> (clo[j] & 15);
            }
        } else {
            int32_t t[G][W];
            if (zpw != nullptr) {
                int32_t zw[W];
                load_par<W>(ts, G + 2, nl, zw);
#pragma unroll
                for (int grp = 0; grp < G; ++grp)
#pragma unroll
                    for (int j = 0; j < W; ++j) t[grp][j] = acc[grp][j] - zw[j] * ts.S[grp];
            } else {
#pragma unroll
                for (int grp = 0; grp < G; ++grp)
#pragma unroll
                    for (int j = 0; j < W; ++j) t[grp][j] = acc[grp][j];
            }
            if (acc_dump != nullptr && ts.mrow >= 0) {      // test hook, off the hot path
                const long long gstride = static_cast<long long>(g.n_img) * g.Ho * g.Wo;
#pragma unroll
                for (int grp = 0; grp < G; ++grp) {
                    int32_t* dp = acc_dump + (grp * gstride + ts.mrow) * n_total + n0;
#pragma unroll
                    for (int j = 0; j < W; j += 4) st_global_v4(dp + j, t[grp][j], t[grp][j + 1], t[grp][j + 2], t[grp][j + 3]);
                }
            }
            long long u[W];
#pragma unroll
            for (int j = 0; j < W; ++j) u[j] = (static_cast<long long>(chi[j]) << 32) | static_cast<uint32_t>(clo[j]);
#pragma unroll
            for (int grp = 0; grp < G; ++grp) {
                int32_t m[W];
                load_par<W>(ts, grp, nl, m);
#pragma unroll
                for (int j = 0; j < W; ++j) u[j] += static_cast<long long>(t[grp][j]) * m[j];
            }
#pragma unroll
            for (int j = 0; j < W; ++j) q[j] = static_cast<int32_t>(u[j] >> 32) >> (clo[j] & 15);
        }
        uint32_t packed[W / 4];
        unsigned rsum = 0;
#pragma unroll
        for (int w = 0; w < W / 4; ++w) {
            packed[w] = pack4_sat_u8(q[4 * w], q[4 * w + 1], q[4 * w + 2], q[4 * w + 3]);
            rsum = __dp4a(packed[w], 0x01010101u, rsum);
        }
        if (ts.opix >= 0 && !(debug & 16)) {
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

}  // namespace qv2x
