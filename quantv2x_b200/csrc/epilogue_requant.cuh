// Fused requantization epilogue of the quantized conv / deconv layers.
//
// Reference semantics (opencood/quant/quant_layer.py:391-410 with :132-148): fake-quant weight ->
// conv (+bias) -> folded BN = identity -> ReLU -> fake-quant activation.  Restated on integers
// (SURVEY Appendix A.1); normative fp32 operation order, every operation rounded separately (no FMA):
//     t_g = float( acc_g[p,n] - zpw_g[n] * S_g[p] )                     (int32 exact, then RN convert)
//     v   = gs[0]*t_0  (+ gs[1]*t_1 + gs[2]*t_2, left to right)
//     y   = v * cs[n] + bias[n] ;  y = max(y, 0) if relu
//     q   = clamp( rint(y / delta_out) + zp_out, 0, qmax )              (true division, half-to-even)
// S_g[p] is the sum of the group's input bytes over the receptive field of p (zero padding adds 0
// because the activation zero-point is 0 after ReLU); it is rebuilt from per-pixel channel sums
// ("rowsums") that the producing layer's epilogue emitted, so the tensor pipe does no extra work.
#pragma once
#include "igemm.cuh"

namespace qv2x {

template <int G>
struct RequantEpilogue {
    static constexpr int kColSplit = 2;
    static constexpr int kMaxStages = 8;
    // output addressing: pixel (oy*up + dy, ox*up + dx) of an [n_img, Hout, Wout, out_cstride] u8 tensor,
    // (dy, dx) = sub-position owned by this N tile (transposed conv with kernel == stride == up)
    int up, cout_sub, Hout, Wout, out_cstride, out_cbase;
    int relu;
    float qmax, delta_out, zp_out;
    float rdelta;                         // fl(1 / delta_out)
    int fast8;                            // 8-bit output, zero-point 0, ReLU: saturating fast path
    float gscale[kMaxGroups];
    const float* cscale;                  // [N_total]
    const float* bias;                    // [N_total]
    const int32_t* zpw[kMaxGroups];       // [N_total] or nullptr (weights already zero-centred)
    const int32_t* rowsum_in[kMaxGroups]; // [n_img*Hi*Wi] or nullptr
    uint8_t* out;
    int32_t* rowsum_out;                  // [n_img*Hout*Wout] (atomically accumulated) or nullptr
    int32_t* acc_dump;                    // [G][n_img*Ho*Wo][N_total] zero-point-corrected accumulators, or nullptr
    int n_total;

    struct Tile {
        int32_t S[G];
        long long opix;   // output pixel index, -1 if this row is outside the image
        long long mrow;   // GEMM row index (for acc_dump)
        int rsum;
    };

    __device__ __forceinline__ void begin(Tile& ts, const IgemmGeom& g, const TileCoord& tc, int row) const {
        const int lx = row % g.tw, ly = row / g.tw;
        const int ox = tc.tx * g.tw + lx, oy = tc.ty * g.th + ly;
        const bool valid = (ox < g.Wo) && (oy < g.Ho);
        ts.rsum = 0;
        ts.opix = -1;
        ts.mrow = -1;
#pragma unroll
        for (int grp = 0; grp < G; ++grp) ts.S[grp] = 0;
        if (!valid) return;
        ts.mrow = (static_cast<long long>(tc.img) * g.Ho + oy) * g.Wo + ox;
#pragma unroll
        for (int grp = 0; grp < G; ++grp) {
            if (rowsum_in[grp] == nullptr) continue;
            const int32_t* rs = rowsum_in[grp] + static_cast<long long>(tc.img) * g.Hi * g.Wi;
            int32_t s = 0;
            for (int tap = 0; tap < g.taps; ++tap) {
                const int ky = tap / g.taps_w, kx = tap - ky * g.taps_w;
                const int iy = oy * g.stride + ky - g.pad, ix = ox * g.stride + kx - g.pad;
                if (iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi) s += __ldg(rs + iy * g.Wi + ix);
            }
            ts.S[grp] = s;
        }
        // an N tile never straddles two sub-positions (BLOCK_N divides cout_sub)
        const int sub = (tc.nt * g.block_n) / cout_sub;
        const int dy = sub / up, dx = sub - dy * up;
        ts.opix = (static_cast<long long>(tc.img) * Hout + oy * up + dy) * Wout + ox * up + dx;
    }

    __device__ __forceinline__ void chunk(Tile& ts, const IgemmGeom& g, const TileCoord& tc, int step, int n0,
                                          const int32_t (*acc)[16]) const {
        (void)tc;
        (void)step;
        uint32_t packed[4] = {0, 0, 0, 0};
        int rsum = 0;
        // per-column parameters: warp-uniform 16-byte loads (L1 broadcast)
        float cs[16], bs[16];
        int32_t zw[G][16];
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
            const float4 c = __ldg(reinterpret_cast<const float4*>(cscale + n0) + v4);
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n0) + v4);
            cs[4 * v4 + 0] = c.x, cs[4 * v4 + 1] = c.y, cs[4 * v4 + 2] = c.z, cs[4 * v4 + 3] = c.w;
            bs[4 * v4 + 0] = b.x, bs[4 * v4 + 1] = b.y, bs[4 * v4 + 2] = b.z, bs[4 * v4 + 3] = b.w;
#pragma unroll
            for (int grp = 0; grp < G; ++grp) {
                if (zpw[grp] != nullptr) {
                    const int4 z = __ldg(reinterpret_cast<const int4*>(zpw[grp] + n0) + v4);
                    zw[grp][4 * v4 + 0] = z.x, zw[grp][4 * v4 + 1] = z.y, zw[grp][4 * v4 + 2] = z.z,
                                     zw[grp][4 * v4 + 3] = z.w;
                }
            }
        }
        if (acc_dump != nullptr && ts.mrow >= 0) {      // test hook, off the hot path
            const long long gstride = static_cast<long long>(g.n_img) * g.Ho * g.Wo;
#pragma unroll
            for (int grp = 0; grp < G; ++grp)
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    int32_t t = acc[grp][j];
                    if (zpw[grp] != nullptr) t -= zw[grp][j] * ts.S[grp];
                    acc_dump[(grp * gstride + ts.mrow) * n_total + n0 + j] = t;
                }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float v = 0.f;
#pragma unroll
            for (int grp = 0; grp < G; ++grp) {
                int32_t t = acc[grp][j];
                if (zpw[grp] != nullptr) t -= zw[grp][j] * ts.S[grp];
                const float tf = __int2float_rn(t);
                const float term = (G == 1) ? tf : __fmul_rn(gscale[grp], tf);   // gscale[0] == 1 when G == 1
                v = (grp == 0) ? term : __fadd_rn(v, term);
            }
            float y = __fadd_rn(__fmul_rn(v, cs[j]), bs[j]);
            uint32_t b;
            if (fast8) {
                // q = sat_u8(rint(y / delta)) for an 8-bit output with zero-point 0 (ReLU is subsumed by the
                // saturation).  The IEEE quotient is replaced by y * fl(1/delta): both lie within 5.4e-5 of
                // each other for |q| < 300, so they round to the same integer unless the product is within
                // 1e-4 of a half-integer -- only then (about 2e-4 of all elements) the exact division runs.
                const float t = __fmul_rn(y, rdelta);
                float f = rintf(t);
                if (fabsf(t - f) > 0.4999f && fabsf(t) < 300.f) f = rintf(__fdiv_rn(y, delta_out));
                asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(b) : "f"(f));
            } else {
                if (relu) y = fmaxf(y, 0.f);
                // A zero dividend would send the whole warp through the division's slow path: divide
                // delta/delta instead and mask.
                const bool nz = (y != 0.f);
                float d = __fdiv_rn(nz ? y : delta_out, delta_out);
                d = nz ? d : 0.f;
                float q = __fadd_rn(rintf(d), zp_out);
                q = fminf(fmaxf(q, 0.f), qmax);
                b = static_cast<uint32_t>(q);
            }
            rsum += static_cast<int>(b);
            packed[j >> 2] |= b << ((j & 3) * 8);
        }
        if (ts.opix >= 0) {
            const int ch = n0 % cout_sub;
            st_global_v4(out + ts.opix * out_cstride + out_cbase + ch, packed[0], packed[1], packed[2], packed[3]);
            ts.rsum += rsum;
        }
    }

    __device__ __forceinline__ void step_end(Tile&, const IgemmGeom&, const TileCoord&, int, int, int, int,
                                             uint8_t*) const {}

    __device__ __forceinline__ void end(Tile& ts, const IgemmGeom& g, const TileCoord& tc) const {
        (void)g;
        (void)tc;
        if (rowsum_out != nullptr && ts.opix >= 0) atomicAdd(rowsum_out + ts.opix, ts.rsum);
    }
};

}  // namespace qv2x
