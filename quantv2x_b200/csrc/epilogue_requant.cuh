// Fused requantization epilogue of the quantized conv / deconv layers.
//
// Reference semantics (opencood/quant/quant_layer.py:391-410 with :132-148): fake-quant weight ->
// conv (+bias) -> folded BN = identity -> ReLU -> fake-quant activation.  Restated on integers
// (SURVEY Appendix A.1); normative fp32 operation order (the scale-and-bias step is ONE fused multiply-add):
//     t_g = float( acc_g[p,n] - zpw_g[n] * S_g[p] )                     (int32 exact, then RN convert)
//     v   = gs[0]*t_0  (+ gs[1]*t_1 + gs[2]*t_2, left to right)
//     y   = fma(v, cs[n], bias[n]) ;  y = max(y, 0) if relu                 (one rounding)
//     q   = clamp( rint(y / delta_out) + zp_out, 0, qmax )              (true division, half-to-even)
// Residual blocks (QuantBasicBlock / QuantBottleneck.forward, opencood/quant/quant_block.py:88-97, 124-134): the last
// conv of a block has no quantizer of its own; `out += residual` comes before the ReLU and the block's quantizer:
//     y   = fma(v, cs, bias) + r ,   r = fl(res_delta * code) for an identity shortcut (the block input, on its
//           quantizer's grid) or the FP32 output of the downsample conv (which is this epilogue with out_f32 set).
// S_g[p] is the sum of the group's input bytes over the receptive field of p (zero padding adds 0
// because the activation zero-point is 0 after ReLU); it is rebuilt from per-pixel channel sums
// ("rowsums") that the producing layer's epilogue emitted, so the tensor pipe does no extra work.
#pragma once
#include "igemm.cuh"

namespace qv2x {

// ------------------------------------------------------------------------------------------------------------
// Receptive-field sums S_g[row] of a tile, rebuilt by a side warp from the producing layer's per-pixel channel sums:
// the (th-1)*stride+taps_h by (tw-1)*stride+taps_w window of sums is fetched once with cp.async (zero outside the
// image), then every tile row adds its taps from shared memory -- ~10x fewer global loads than taps x rows.
// Shared by the float and the fixed-point requant epilogues.
struct SideHalo {
    static constexpr int kHaloInts = 1024;   // >= (2*127+3)*3, the widest halo (stride 2, 128 x 1 tile box)
    // Per-lane constants of the side warp, computed once per kernel: which halo elements the lane fetches and
    // where its four tile rows start inside the halo.  (A single warp runs this code, so every instruction it does
    // NOT execute per tile is latency taken off the tile period.)
    static constexpr int kFlat = 8;          // halos of up to 32 * kFlat elements take the unrolled path
    int hw, hh, n;
    int hoff[4];
    int hy[kFlat], hx[kFlat];
    // 3x3 stride-1 convs on 8 x 16 pixel tiles (every HALO layer): separable box sum -- lane (lx, q) owns the four
    // vertically adjacent rows 4q..4q+3 of column lx, reads six halo values of its column (plus six of the tile's
    // outer column on the edge lanes), forms the vertical 3-sums and gets its neighbours' by shuffle: 12 shared-memory
    // loads per lane instead of 36 (they are starved while the tensor core streams operands).
    bool box;

    __device__ __forceinline__ void init(const IgemmGeom& g, int lane) {
        hw = (g.tw - 1) * g.stride + g.taps_w;
        hh = (g.th - 1) * g.stride + g.taps_h;
        n = hw * hh;
        box = (g.stride == 1 && g.taps_h == 3 && g.taps_w == 3 && g.tw == 8 && g.th == 16);
#pragma unroll
        for (int j = 0; j < kFlat; ++j) {
            const int e = lane + 32 * j;
            hy[j] = e / hw;
            hx[j] = e - hy[j] * hw;
        }
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int row = lane + 32 * rr;
            const int lx = row & (g.tw - 1), ly = row >> g.tw_shift;      // tw is a power of two
            hoff[rr] = ly * g.stride * hw + lx * g.stride;
        }
    }

    // s_S[grp * 128 + row] for grp < G; rowsum_in[grp] == nullptr -> zeros
    template <int G>
    __device__ __forceinline__ void stage(const IgemmGeom& g, const TileCoord& tc, int lane, int32_t* s_S,
                                          int32_t* halo, const int32_t* const (&rowsum_in)[kMaxGroups]) const {
        int32_t sums[G][4];
        compute<G>(g, tc, lane, halo, rowsum_in, sums);
        store<G>(lane, s_S, sums);
    }

    template <int G>
    __device__ __forceinline__ void store(int lane, int32_t* s_S, const int32_t (&sums)[G][4]) const {
        // tile row of sums[.][rr]: lane + 32 rr, or (box) row 4q + rr of column lx = 32 q + 8 rr + lx
        const int r0 = box ? 32 * (lane >> 3) + (lane & 7) : lane;
        const int rstep = box ? 8 : 32;
#pragma unroll
        for (int grp = 0; grp < G; ++grp)
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) s_S[grp * kTileM + r0 + rstep * rr] = sums[grp][rr];
    }

    // The sums of this lane's four tile rows, in registers: touches only the warp's private halo buffer, so the side
    // warp can run it for the NEXT tile of its slot while the epilogue still reads the slot (igemm.cuh).
    template <int G>
    __device__ __forceinline__ void compute(const IgemmGeom& g, const TileCoord& tc, int lane, int32_t* halo,
                                            const int32_t* const (&rowsum_in)[kMaxGroups],
                                            int32_t (&sums)[G][4]) const {
        const int ix0 = tc.tx * g.tw * g.stride - g.pad, iy0 = tc.ty * g.th * g.stride - g.pad;
        const uint32_t halo_s = smem_u32(halo);
#pragma unroll
        for (int grp = 0; grp < G; ++grp) {
            const int32_t* rs = rowsum_in[grp];
            if (rs == nullptr) {
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) sums[grp][rr] = 0;
                continue;
            }
            rs += static_cast<long long>(tc.img) * g.Hi * g.Wi;
            __syncwarp();                                   // the previous group's readers are done with the halo
            if (n <= 32 * kFlat) {
#pragma unroll
                for (int j = 0; j < kFlat; ++j) {
                    const int e = lane + 32 * j;
                    if (e < n) {
                        const int iy = iy0 + hy[j], ix = ix0 + hx[j];
                        const bool ok = (iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi);
                        cp_async_4(halo_s + 4 * e, ok ? rs + iy * g.Wi + ix : rs, ok);
                    }
                }
            } else {
                for (int y = 0; y < hh; ++y) {
                    const int iy = iy0 + y;
                    const bool yok = (iy >= 0 && iy < g.Hi);
                    for (int x = lane; x < hw; x += 32) {
                        const int ix = ix0 + x;
                        const bool ok = yok && ix >= 0 && ix < g.Wi;
                        cp_async_4(halo_s + 4 * (y * hw + x), ok ? rs + iy * g.Wi + ix : rs, ok);
                    }
                }
            }
            cp_async_wait_all();
            __syncwarp();
            if (box) {
                const int lx = lane & 7, q = lane >> 3;
                const int32_t* col = halo + (4 * q) * 10 + lx + 1;                     // own column, halo rows 4q..4q+5
                const int32_t* edge = halo + (4 * q) * 10 + (lx == 7 ? 9 : (lx == 0 ? 0 : lx + 1));
                int32_t a[6], e[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) a[r] = col[10 * r], e[r] = edge[10 * r];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int32_t own = (a[j] + a[j + 1]) + a[j + 2];
                    const int32_t out = (e[j] + e[j + 1]) + e[j + 2];                  // only used on the edge lanes
                    const int32_t up = __shfl_up_sync(0xffffffffu, own, 1, 8);         // column lx - 1
                    const int32_t dn = __shfl_down_sync(0xffffffffu, own, 1, 8);       // column lx + 1
                    sums[grp][j] = (lx == 0 ? out : up) + own + (lx == 7 ? out : dn);
                }
                continue;
            }
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int32_t* h0 = halo + hoff[rr];
                int32_t sum = 0;
                if (g.taps == 9) {
                    const int32_t *h1 = h0 + hw, *h2 = h1 + hw;
                    sum = ((h0[0] + h0[1]) + (h0[2] + h1[0])) + ((h1[1] + h1[2]) + (h2[0] + h2[1])) + h2[2];
                } else {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
                            if (ky < g.taps_h && kx < g.taps_w) sum += h0[ky * hw + kx];
                }
                sums[grp][rr] = sum;                          // rows outside the image are never stored
            }
        }
    }
};

template <int G, bool DIGITS = false, bool FAST8 = true, bool RES = false>
struct RequantEpilogue {
    // epilogue warps per TMEM lane quadrant: more warps hide the latencies of the (ALU-pipe bound) requant math
    static constexpr int col_split(int) { return 2; }   // (4 was measured: spills and no gain -- not latency bound)
    static constexpr int kMaxStages = 8;
    static constexpr bool kSideWarp = true;        // per-column parameters and per-row input sums come via warp 3
    static constexpr bool kSeqDrain = (G > 1);     // groups are drained one by one into fp32 partial sums
    // output addressing: pixel (oy*up + dy, ox*up + dx) of an [n_img, Hout, Wout, out_cstride] u8 tensor,
    // (dy, dx) = sub-position owned by this N tile (transposed conv with kernel == stride == up)
    int up, cout_sub, Hout, Wout, out_cstride, out_cbase;
    int up_shift;                         // log2(up); up is 1, 2 or 4
    int debug;                            // qv2x_set_debug_flags: 16 skips the output stores
    FastDiv fd_cout_sub;
    int relu;
    float qmax, delta_out, zp_out;
    float rdelta;                         // fl(1 / delta_out)
    // FAST8: 8-bit output, zero-point 0, ReLU -- the saturating fast path (a template parameter, so that the
    // generic path's W inline divisions do not bloat the hot kernels)
    // DIGITS (G == 3): digit GEMM reduced in the order mid, lo, hi and combined on integers (see accum)
    float gscale[kMaxGroups];
    const float* cscale;                  // [N_total]
    const float* bias;                    // [N_total]
    const int32_t* zpw[kMaxGroups];       // [N_total] or nullptr (weights already zero-centred)
    const int32_t* rowsum_in[kMaxGroups]; // [n_img*Hi*Wi] or nullptr
    uint8_t* out;
    int32_t* rowsum_out;                  // [n_img*Hout*Wout] (atomically accumulated) or nullptr
    int32_t* acc_dump;                    // [G][n_img*Ho*Wo][N_total] zero-point-corrected accumulators, or nullptr
    int n_total;
    // shortcut added before the ReLU (generic path, or FAST8 with RES); FP32 output without a quantizer (generic only)
    const uint8_t* res_u8;                // [n_img*Hout*Wout][res_cstride] codes at channel res_cbase, scale res_delta
    const float* res_f32;                 // [n_img*Hout*Wout][res_cstride] floats at channel res_cbase
    float res_delta;
    int res_cstride, res_cbase;
    float* out_f32;                       // [n_img*Hout*Wout][out_f32_cstride]: y after the (optional) ReLU, no codes
    int out_f32_cstride;

    struct Tile {
        int32_t S[G];
        long long opix;   // output pixel index, -1 if this row is outside the image
        long long mrow;   // GEMM row index (for acc_dump)
        int rsum;
        uint32_t sm_par;         // shared-window address of this tile's per-column parameters: cs | bias | zpw
        int n_base;              // first global column of the tile
        int ch_off;              // output channel of column n = n - ch_off (a tile never straddles sub-positions)
    };
    // Shared-memory slot written by the side warp for one tile (kEpiSmemBytes / 2 = 5 KB):
    //   [0, 1K) cs[256] f32 | [1K, 2K) bias[256] f32 | [2K, 3K) zpw[256] i32 | [3K, 3K + G*512) S[G][128] i32
    // S[g][row] = sum of group g's input bytes over the receptive field of the tile's row (zero padding adds 0),
    // rebuilt from the producing layer's per-pixel channel sums.
    static constexpr int kSlotS = 3072;

    static constexpr int kHaloInts = SideHalo::kHaloInts;
    using Side = SideHalo;
    __device__ __forceinline__ void side_init(Side& sd, const IgemmGeom& g, int lane) const { sd.init(g, lane); }

    __device__ __forceinline__ void side_load(const IgemmGeom& g, const TileCoord& tc, int lane, uint8_t* slot,
                                              int32_t* halo, int& staged_nt, const Side& sd) const {
        if (staged_nt != tc.nt) {       // per-column parameters change only with the column tile
            float* s_cs = reinterpret_cast<float*>(slot);
            float* s_b = s_cs + 256;
            int32_t* s_z = reinterpret_cast<int32_t*>(s_b + 256);
            const int n_base = tc.nt * g.block_n;
            const bool has_zp = (zpw[0] != nullptr);                  // groups share one zero-point array
            for (int i = lane; i < g.block_n; i += 32) {
                cp_async_4(smem_u32(s_cs + i), cscale + n_base + i, true);
                cp_async_4(smem_u32(s_b + i), bias + n_base + i, true);
                cp_async_4(smem_u32(s_z + i), has_zp ? static_cast<const void*>(zpw[0] + n_base + i) : cscale, has_zp);
            }
            staged_nt = tc.nt;
        }
        sd.template stage<G>(g, tc, lane, reinterpret_cast<int32_t*>(slot + kSlotS), halo, rowsum_in);
        cp_async_wait_all();                       // the parameter copies, when no group waited for them
    }

    // Two-phase form (igemm.cuh): the tile's receptive-field sums are formed in registers before the slot is free.
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
            float* s_cs = reinterpret_cast<float*>(slot);
            float* s_b = s_cs + 256;
            int32_t* s_z = reinterpret_cast<int32_t*>(s_b + 256);
            const int n_base = tc.nt * g.block_n;
            const bool has_zp = (zpw[0] != nullptr);
            for (int i = lane; i < g.block_n; i += 32) {
                cp_async_4(smem_u32(s_cs + i), cscale + n_base + i, true);
                cp_async_4(smem_u32(s_b + i), bias + n_base + i, true);
                cp_async_4(smem_u32(s_z + i), has_zp ? static_cast<const void*>(zpw[0] + n_base + i) : cscale, has_zp);
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
            ts.S[grp] = *reinterpret_cast<const int32_t*>(slot + kSlotS + 4 * (grp * kTileM + row));
        if (!valid) return;
        ts.mrow = (static_cast<long long>(tc.img) * g.Ho + oy) * g.Wo + ox;
        // an N tile never straddles two sub-positions (BLOCK_N divides cout_sub)
        int dy = 0, dx = 0;
        if (up > 1) {
            const int sub = fd_cout_sub.div(ts.n_base);
            dy = sub >> up_shift;
            dx = sub - (dy << up_shift);
        }
        ts.opix = (static_cast<long long>(tc.img) * Hout + oy * up + dy) * Wout + ox * up + dx;
    }

    // Per-column parameters of W consecutive columns from shared memory: warp-uniform 16-byte loads (broadcast).
    template <int W>
    __device__ __forceinline__ void load_zw(const Tile& ts, int nl, int32_t (&zw)[W]) const {
#pragma unroll
        for (int v4 = 0; v4 < W / 4; ++v4) {
            const int4 z = lds_i4(ts.sm_par + 2048 + 4 * nl + 16 * v4);
            zw[4 * v4 + 0] = z.x, zw[4 * v4 + 1] = z.y, zw[4 * v4 + 2] = z.z, zw[4 * v4 + 3] = z.w;
        }
    }

    // One accumulator group of W columns -> running fp32 sum v (normative order: groups left to right).
    template <int W>
    __device__ __forceinline__ void accum(const Tile& ts, const IgemmGeom& g, int grp, int n0, const int32_t (&acc)[W],
                                          float (&v)[W]) const {
        if constexpr (G == 3 && DIGITS) {
            // 24-bit fixed-point weights as three signed byte digits, groups arrive in the order mid, lo, hi:
            //   v = fl32( 65536 * hi + fl32(256 * mid + lo) )     (256 * mid + lo is exact in int32, hi exact in fp32)
            // the running "sum" holds the raw mid accumulator (as bits) between the first two groups
            if (acc_dump != nullptr && ts.mrow >= 0) {      // test hook: dumped in hi, mid, lo order
                const long long gstride = static_cast<long long>(g.n_img) * g.Ho * g.Wo;
                const int dg = (grp == 0) ? 1 : (grp == 1 ? 2 : 0);
                int32_t* dp = acc_dump + (dg * gstride + ts.mrow) * n_total + n0;
#pragma unroll
                for (int j = 0; j < W; j += 4) st_global_v4(dp + j, acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            }
#pragma unroll
            for (int j = 0; j < W; ++j) {
                if (grp == 0) v[j] = __int_as_float(acc[j]);
                else if (grp == 1) v[j] = __int2float_rn(__float_as_int(v[j]) * 256 + acc[j]);
                else v[j] = fmaf(__int2float_rn(acc[j]), 65536.f, v[j]);
            }
            return;
        }
        const bool use_zp = (zpw[0] != nullptr);
        int32_t t[W];
        if (use_zp) {
            int32_t zw[W];
            load_zw<W>(ts, n0 - ts.n_base, zw);
            int32_t sg = 0;
#pragma unroll
            for (int q = 0; q < G; ++q)
                if (q == grp) sg = ts.S[q];
#pragma unroll
            for (int j = 0; j < W; ++j) t[j] = acc[j] - zw[j] * sg;
        } else {
#pragma unroll
            for (int j = 0; j < W; ++j) t[j] = acc[j];
        }
        if (acc_dump != nullptr && ts.mrow >= 0) {      // test hook, off the hot path
            const long long gstride = static_cast<long long>(g.n_img) * g.Ho * g.Wo;
            int32_t* dp = acc_dump + (grp * gstride + ts.mrow) * n_total + n0;
#pragma unroll
            for (int j = 0; j < W; j += 4) st_global_v4(dp + j, t[j], t[j + 1], t[j + 2], t[j + 3]);
        }
        float gs = 1.f;
#pragma unroll
        for (int q = 0; q < G; ++q)
            if (q == grp) gs = gscale[q];
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const float tf = __int2float_rn(t[j]);
            const float term = (G == 1) ? tf : __fmul_rn(gs, tf);   // gscale[0] == 1 when G == 1
            v[j] = (grp == 0) ? term : __fadd_rn(v[j], term);
        }
    }

    // The shortcut of a residual block for W consecutive channels of this row: fl(res_delta * code) from the block's
    // uint8 input, or the FP32 output of the downsample conv.  Returns false (and zeros) when the layer has none.
    template <int W>
    __device__ __forceinline__ bool load_res(const Tile& ts, int n0, float (&radd)[W]) const {
#pragma unroll
        for (int j = 0; j < W; ++j) radd[j] = 0.f;
        const bool has_res = (res_u8 != nullptr) || (res_f32 != nullptr);
        if (has_res && ts.opix >= 0) {
            const long long ro = ts.opix * res_cstride + res_cbase + (n0 - ts.ch_off);
            if (res_u8 != nullptr) {
#pragma unroll
                for (int w = 0; w < W / 4; ++w) {
                    const uint32_t rb = __ldg(reinterpret_cast<const uint32_t*>(res_u8 + ro) + w);
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        radd[4 * w + b] = __fmul_rn(static_cast<float>((rb >> (8 * b)) & 0xffu), res_delta);
                }
            } else {
#pragma unroll
                for (int w = 0; w < W / 4; ++w) {
                    const float4 rf = __ldg(reinterpret_cast<const float4*>(res_f32 + ro) + w);
                    radd[4 * w + 0] = rf.x, radd[4 * w + 1] = rf.y, radd[4 * w + 2] = rf.z, radd[4 * w + 3] = rf.w;
                }
            }
        }
        return has_res;
    }

    // v (the scaled accumulator sum) -> bias, ReLU, output quantization, packed store of W bytes.
    template <int W>
    __device__ __forceinline__ void finish(Tile& ts, int n0, const float (&v)[W]) const {
        static_assert(W == 8 || W == 16, "chunk width");
        uint32_t packed[W / 4] = {};
        unsigned rsum = 0;
        float cs[W], bs[W];
        const int nl = n0 - ts.n_base;
#pragma unroll
        for (int v4 = 0; v4 < W / 4; ++v4) {
            const float4 c = lds_f4(ts.sm_par + 4 * nl + 16 * v4);
            const float4 b = lds_f4(ts.sm_par + 1024 + 4 * nl + 16 * v4);
            cs[4 * v4 + 0] = c.x, cs[4 * v4 + 1] = c.y, cs[4 * v4 + 2] = c.z, cs[4 * v4 + 3] = c.w;
            bs[4 * v4 + 0] = b.x, bs[4 * v4 + 1] = b.y, bs[4 * v4 + 2] = b.z, bs[4 * v4 + 3] = b.w;
        }
        if constexpr (FAST8) {
            // q = sat_u8(rint(y / delta)), y = fma(v, cs, b), for an 8-bit output with zero-point 0 (ReLU is subsumed
            // by the clamp), entirely on the FMA/ALU pipes and mostly in immediate-operand forms (which issue at
            // twice the rate of three-register forms):
            //   t = y * fl(1/delta) lies within 5.4e-5 of the IEEE quotient for |q| < 300, so both round to the same
            //   integer unless t is within 1e-4 of a half-integer -- only then (~2e-4 of elements) the exact division
            //   runs.  The clamp is applied to t (to [-0.25, 255.25], which rounds to 0 / 255 and is never "near"),
            //   and s = t + 1.5*2^23 carries rint(t) (round-half-even) in its low mantissa byte.
            uint32_t bits[W];
            float y[W];
            float worst = 0.f;
            float radd[W];
            if constexpr (RES) load_res<W>(ts, n0, radd);
#pragma unroll
            for (int j = 0; j < W; ++j) {
                y[j] = fmaf(v[j], cs[j], bs[j]);
                if constexpr (RES) y[j] = __fadd_rn(y[j], radd[j]);
                const float t = fminf(fmaxf(__fmul_rn(y[j], rdelta), -0.25f), 255.25f);
                const float sft = __fadd_rn(t, 12582912.0f);
                const float r = __fadd_rn(sft, -12582912.0f);
                worst = fmaxf(worst, fabsf(__fadd_rn(t, -r)));
                bits[j] = __float_as_uint(sft);
            }
            // rare: exact IEEE division for the elements that sit next to a rounding boundary
            if (worst > 0.4999f) {
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    const float t = fminf(fmaxf(__fmul_rn(y[j], rdelta), -0.25f), 255.25f);
                    const float r = __fadd_rn(__fadd_rn(t, 12582912.0f), -12582912.0f);
                    if (fabsf(__fadd_rn(t, -r)) > 0.4999f) {
                        const float q = fminf(fmaxf(rintf(__fdiv_rn(y[j], delta_out)), 0.f), 255.f);
                        bits[j] = __float_as_uint(__fadd_rn(q, 12582912.0f));
                    }
                }
            }
#pragma unroll
            for (int w = 0; w < W / 4; ++w) {
                const uint32_t lo = __byte_perm(bits[4 * w + 0], bits[4 * w + 1], 0x0040);
                const uint32_t hi = __byte_perm(bits[4 * w + 2], bits[4 * w + 3], 0x0040);
                packed[w] = __byte_perm(lo, hi, 0x5410);
                rsum = __dp4a(packed[w], 0x01010101u, static_cast<unsigned>(rsum));
            }
        } else {
            const bool live = (ts.opix >= 0);
            float radd[W];
            const bool has_res = load_res<W>(ts, n0, radd);
            if (out_f32 != nullptr) {               // no output quantizer (downsample conv, occupancy head)
                float yo[W];
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    float y = fmaf(v[j], cs[j], bs[j]);
                    if (has_res) y = __fadd_rn(y, radd[j]);
                    yo[j] = relu ? fmaxf(y, 0.f) : y;
                }
                if (live && !(debug & 16)) {
                    float4* dst = reinterpret_cast<float4*>(out_f32 + ts.opix * out_f32_cstride + (n0 - ts.ch_off));
#pragma unroll
                    for (int w = 0; w < W / 4; ++w)
                        dst[w] = make_float4(yo[4 * w], yo[4 * w + 1], yo[4 * w + 2], yo[4 * w + 3]);
                }
                return;
            }
#pragma unroll
            for (int j = 0; j < W; ++j) {
                float y = fmaf(v[j], cs[j], bs[j]);
                if (has_res) y = __fadd_rn(y, radd[j]);
                if (relu) y = fmaxf(y, 0.f);
                // A zero dividend would send the whole warp through the division's slow path: divide
                // delta/delta instead and mask.
                const bool nz = (y != 0.f);
                float d = __fdiv_rn(nz ? y : delta_out, delta_out);
                d = nz ? d : 0.f;
                float q = __fadd_rn(rintf(d), zp_out);
                q = fminf(fmaxf(q, 0.f), qmax);
                const uint32_t b = static_cast<uint32_t>(q);
                rsum += static_cast<int>(b);
                packed[j >> 2] |= b << ((j & 3) * 8);
            }
        }
        if (ts.opix >= 0 && !(debug & 16)) {
            const int ch = n0 - ts.ch_off;
            if constexpr (W == 16)
                st_global_v4(out + ts.opix * out_cstride + out_cbase + ch, packed[0], packed[1], packed[2], packed[3]);
            else
                st_global_v2(out + ts.opix * out_cstride + out_cbase + ch, packed[0], packed[1]);
            ts.rsum += static_cast<int>(rsum);
        }
    }

    // G == 1 convenience: one chunk straight from the accumulator to the output.
    template <int W>
    __device__ __forceinline__ void chunk(Tile& ts, const IgemmGeom& g, const TileCoord& tc, int step, int n0,
                                          const int32_t (*acc)[W]) const {
        (void)tc;
        (void)step;
        float v[W];
        accum<W>(ts, g, 0, n0, acc[0], v);
        finish<W>(ts, n0, v);
    }

    static constexpr bool kHoldSlots = false;
    __device__ __forceinline__ void step_end(Tile&, const IgemmGeom&, const TileCoord&, int, int, int, int, uint8_t*,
                                             const TmemView&) const {}

    __device__ __forceinline__ void end(Tile& ts, const IgemmGeom& g, const TileCoord& tc) const {
        (void)g;
        (void)tc;
        if (rowsum_out != nullptr && ts.opix >= 0) atomicAdd(rowsum_out + ts.opix, ts.rsum);
    }
};

}  // namespace qv2x
