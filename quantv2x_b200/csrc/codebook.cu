// qv2x_codebook: the multi-level residual multi-codebook quantizer (reference UMGMQuantizer,
// opencood/models/sub_modules/codebook.py:280-343) as two kernels:
//
//   encode : every head is an affine map, so the distance score of every level is affine in the input
//            row and in the codewords already chosen (oracle/codebook_oracle.py):
//                score[l][(s,k)] = G_l[(s,k),:] . x + g0_l[(s,k)] + sum_{j<l,s'} B_{l,j,s'}[code_{j,s'}][(s,k)]
//            x = delta * q with q the shrinker's uint8 output, so G_l . q is an int8 GEMM on the tensor
//            cores (G carried as 24-bit fixed point = three signed byte digits, three exact int32
//            accumulators); the rest of the score and the running argmin (ties -> lowest index) run in
//            float64 in the epilogue, level after level on the same CTA.
//   decode : the six heads collapse to  decode(codes) = const + sum_{l,s} T_{l,s}[code_{l,s}]  (fp32 tables).
#include <cmath>
#include <cstring>
#include <vector>

#include "igemm_launch.cuh"

namespace qv2x {

constexpr int kMaxLevels = 4;
constexpr int kMaxSeg = 4;

// int32 -> float64, exact, without a conversion instruction: the integer is planted in the mantissa of 2^52 + 2^31
__device__ __forceinline__ double i2d_exact(int x) {
    return __hiloint2double(0x43300000, x ^ 0x80000000) - 4503601774854144.0;
}

// FAST: fp32 filter with float64 fallback (one step per level); MSEG: segments per level as a compile-time constant
// (0 = run-time m).  Both are template parameters so that the hot loop carries no code of the other variants.
template <bool FAST, int MSEG>
struct EncodeEpilogue {
    // Epilogue warps per row quadrant (their candidates are merged per level).  The chunk loop is latency-bound
    // (TMEM read, parameter loads, table gathers, dependent min/max chains): the hot configuration (fp32 filter, one
    // segment) runs four warps per quadrant -- 16 epilogue warps, 32 columns of a 128-column level each.
    static constexpr int kSplit = (FAST && MSEG == 1) ? 4 : 2;
    static constexpr int kSegSlots = MSEG > 0 ? MSEG : kMaxSeg;      // candidate slots per row in the merge scratch
    static constexpr int col_split(int) { return kSplit; }
    static constexpr int kMaxStages = 4;   // K is only C bytes: a short ring leaves L1 room for the table gathers
    static constexpr bool kSideWarp = false;
    static constexpr bool kSeqDrain = false;
    int levels, m;
    int k[kMaxLevels];            // codewords per segment
    int n_level[kMaxLevels];      // m * k
    int colbase[kMaxLevels];      // offset of the level's columns in sc / g0
    int step_level[kMaxSteps];    // level of each step
    int step_col0[kMaxSteps];     // first column (within the level) of each step
    int step_last[kMaxSteps];     // 1 if the step closes its level
    double delta;
    const double* dsc;            // fl64(delta * sc[j]): the value of one fixed-point unit of column j
    const double* g0;             // per-column constant
    const double* btab;           // codeword cross terms, see boff
    long long boff[kMaxLevels][kMaxLevels];  // boff[l][j]: start of [m][k_j][n_level[l]] doubles
    uint8_t* codes;               // [levels][m][rows]
    long long rows;
    // fp32 filter (fast != 0; needs one step per level so that the level's accumulators are still in TMEM when a
    // row turns out to need the exact evaluation):
    //   the scores are first evaluated in fp32 with a rigorous error bound eps(row) (levels <= 3); if the best score beats the
    //   runner-up by more than 2*eps the fp32 argmin IS the float64 argmin, otherwise (~1e-4 of rows) the warp
    //   re-reads the accumulators and evaluates the level in float64 exactly as the slow path does.
    int pack16;
    const float* d32;             // float(delta * sc[j])
    const float* g32;             // float(g0[j])
    const float* btab32;          // float(btab), same layout
    float dmax[kMaxLevels];       // max_j |d32[j]| of the level
    float cabs[kMaxLevels];       // max_j |g0[j]| + sum_{j<l} max |B_{l,j}|

    __device__ __forceinline__ int nseg() const { return MSEG > 0 ? MSEG : m; }

    struct Tile {
        long long r;
        double best[kMaxSeg];
        int bidx[kMaxSeg];
        int code[kMaxLevels][kMaxSeg];
        float fbest[kMaxSeg], fsecond[kMaxSeg];
    };

    struct Side {};
    __device__ __forceinline__ void side_init(Side&, const IgemmGeom&, int) const {}
    __device__ __forceinline__ void side_load(const IgemmGeom&, const TileCoord&, int, uint8_t*, int32_t*, int&,
                                              const Side&) const {}

    __device__ __forceinline__ void begin(Tile& ts, const IgemmGeom& g, const TileCoord& tc, int row,
                                          const uint8_t*) const {
        ts.r = static_cast<long long>(tc.tx) * g.tw + row;
#pragma unroll
        for (int s = 0; s < kMaxSeg; ++s) {
            ts.best[s] = INFINITY;
            ts.bidx[s] = 0;
            ts.fbest[s] = INFINITY;
            ts.fsecond[s] = INFINITY;
        }
#pragma unroll
        for (int l = 0; l < kMaxLevels; ++l)
#pragma unroll
            for (int s = 0; s < kMaxSeg; ++s) ts.code[l][s] = 0;
    }

    template <int W>
    __device__ __forceinline__ void chunk(Tile& ts, const IgemmGeom& g, const TileCoord& tc, int step, int c0,
                                          const int32_t (*acc)[W]) const {
        static_assert(W == 16, "the encode epilogue works on 16-column chunks");
        if constexpr (FAST) chunk_fast(ts, step, c0, acc, dbg_flags(g));
        else chunk_exact<16>(ts, g, tc, step, c0, acc);
    }

    // fp32 scores of 16 columns + running (best, runner-up, index) and max |V| of the row
    __device__ __forceinline__ void chunk_fast(Tile& ts, int step, int c0, const int32_t (*acc)[16],
                                               const int dbg = 0) const {
        const int l = step_level[step];
        const int n0 = step_col0[step] + c0;
        const int nl = n_level[l];
        const int seg = n0 / k[l];
        const int kk0 = n0 - seg * k[l];
        float sc16[16];
        {
            const float4* dp = reinterpret_cast<const float4*>(d32 + colbase[l] + n0);
            const float4* gp = reinterpret_cast<const float4*>(g32 + colbase[l] + n0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 dv = __ldg(dp + q), gv = __ldg(gp + q);
                const float dd[4] = {dv.x, dv.y, dv.z, dv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = 4 * q + e;
                    const float hi = __int2float_rn(acc[0][j]);
                    // |256 * mid + lo| < 2^31 when the reduction has at most 256 terms (pack16)
                    float lo;
                    if (pack16) {
                        lo = __int2float_rn(acc[1][j] * 256 + acc[2][j]);
                    } else {
                        const float md = __int2float_rn(acc[1][j]), l2 = __int2float_rn(acc[2][j]);
                        lo = fmaf(md, 256.f, l2);
                    }
                    const float vf = fmaf(hi, 65536.f, lo);
                    sc16[j] = fmaf(vf, dd[e], gg[e]);
                }
            }
        }
#pragma unroll
        for (int jl = 0; jl < kMaxLevels - 1; ++jl) {
            if (jl < l && !(dbg & 512)) {
#pragma unroll
                for (int s2 = 0; s2 < kMaxSeg; ++s2) {
                    if (s2 < nseg()) {
                        const float* bp =
                            btab32 + boff[l][jl] + (static_cast<long long>(s2) * k[jl] + ts.code[jl][s2]) * nl + n0;
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            float b[8];
                            ldg_nc_f8(bp + 8 * q, b);
#pragma unroll
                            for (int e = 0; e < 8; ++e) sc16[8 * q + e] += b[e];
                        }
                    }
                }
            }
        }
        float best = INFINITY, second = INFINITY;
        int bidx = 0;
#pragma unroll
        for (int s = 0; s < kMaxSeg; ++s)
            if (s == seg) best = ts.fbest[s], second = ts.fsecond[s], bidx = ts.bidx[s];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float v = sc16[j];
            if (!(dbg & 128)) {
                second = fminf(second, fmaxf(best, v));
                if (v < best) bidx = kk0 + j;      // the exact index only matters when the gap test passes
            }
            best = fminf(best, v);
        }
#pragma unroll
        for (int s = 0; s < kMaxSeg; ++s)
            if (s == seg) ts.fbest[s] = best, ts.fsecond[s] = second, ts.bidx[s] = bidx;
    }

    template <int W>
    __device__ __forceinline__ void chunk_exact(Tile& ts, const IgemmGeom&, const TileCoord&, int step, int c0,
                                                const int32_t (*acc)[W]) const {
        const int l = step_level[step];
        const int n0 = step_col0[step] + c0;          // column within the level; a chunk never straddles segments
        const int nl = n_level[l];
        const int seg = n0 / k[l];
        const int kk0 = n0 - seg * k[l];
        const double* scp = dsc + colbase[l] + n0;
        const double* g0p = g0 + colbase[l] + n0;
        // The cross-term rows are per-thread gathers (random rows, 128 B per 16-column chunk): pull the NEXT chunk's
        // lines into L1 now so its loads do not expose the L2 latency.
        if (n0 + 16 < nl) {
#pragma unroll
            for (int jl = 0; jl < kMaxLevels - 1; ++jl)
                if (jl < l) {
#pragma unroll
                    for (int s2 = 0; s2 < kMaxSeg; ++s2)
                        if (s2 < nseg()) {
                            const double* nx = btab + boff[l][jl] +
                                               (static_cast<long long>(s2) * k[jl] + ts.code[jl][s2]) * nl + n0 + 16;
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(nx));
                        }
                }
        }
        double score[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            // V = acc_hi*65536 + acc_mid*256 + acc_lo, exact in float64 (|V| < 2^48); the int -> double conversions
            // go through the 2^52 mantissa trick (one LOP + one DADD) instead of the slow conversion unit
            const double lo = pack16 ? i2d_exact(acc[1][j] * 256 + acc[2][j])
                                     : fma(i2d_exact(acc[1][j]), 256.0, i2d_exact(acc[2][j]));
            const double V = fma(i2d_exact(acc[0][j]), 65536.0, lo);
            score[j] = __dadd_rn(__dmul_rn(V, __ldg(scp + j)), __ldg(g0p + j));     // scp = fl64(delta * sc)
        }
#pragma unroll
        for (int jl = 0; jl < kMaxLevels - 1; ++jl) {
            if (jl < l) {
#pragma unroll
                for (int s2 = 0; s2 < kMaxSeg; ++s2) {
                    if (s2 < nseg()) {
                        const double* bp = btab + boff[l][jl] +
                                           (static_cast<long long>(s2) * k[jl] + ts.code[jl][s2]) * nl + n0;
#pragma unroll
                        for (int j2 = 0; j2 < 8; ++j2) {
                            const double2 b = __ldg(reinterpret_cast<const double2*>(bp) + j2);
                            score[2 * j2] = __dadd_rn(score[2 * j2], b.x);
                            score[2 * j2 + 1] = __dadd_rn(score[2 * j2 + 1], b.y);
                        }
                    }
                }
            }
        }
        double best = INFINITY;
        int bidx = 0;
#pragma unroll
        for (int s = 0; s < kMaxSeg; ++s)
            if (s == seg) best = ts.best[s], bidx = ts.bidx[s];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (score[j] < best) {      // strict: ties keep the lowest index
                best = score[j];
                bidx = kk0 + j;
            }
        }
#pragma unroll
        for (int s = 0; s < kMaxSeg; ++s)
            if (s == seg) ts.best[s] = best, ts.bidx[s] = bidx;
    }

    // The float64 evaluation of one level for this warp (rare path of the fp32 filter).
    __device__ __forceinline__ void redo_level_exact(Tile& ts, const IgemmGeom& g, const TileCoord& tc, int step,
                                                  const TmemView& tv) const {
#pragma unroll
        for (int s = 0; s < kMaxSeg; ++s) ts.best[s] = INFINITY, ts.bidx[s] = 0;
        for (int c0 = tv.c_begin; c0 < tv.c_end; c0 += 16) {
            uint32_t acc[3][16];
#pragma unroll
            for (int grp = 0; grp < 3; ++grp) tmem_ld_x16(tv.addr[grp] + c0, acc[grp]);
            tmem_ld_wait();
            chunk_exact<16>(ts, g, tc, step, c0, reinterpret_cast<const int32_t(*)[16]>(acc));
        }
    }

    static constexpr bool kHoldSlots = true;     // the level's accumulators may be re-read in step_end
    static constexpr float kEpsRel = 9.5367431640625e-07f;   // 2^-20 = 16 ulp(fp32): see the bound in step_end

    __device__ __forceinline__ void step_end(Tile& ts, const IgemmGeom& g, const TileCoord& tc, int step, int part,
                                             int quad, int lane, uint8_t* scratch, const TmemView& tv) const {
        if (!step_last[step]) return;
        const int l = step_level[step];
        const int bar_id = 1 + quad;
        bool exact_merge = true;
        if constexpr (FAST) {
            // ---- merge the fp32 candidates of the two column halves and test the gap.
            // Error bound of one fp32 score s against its float64 evaluation.  V = 65536 hi + (256 mid + lo) exactly;
            // the computed vf differs from it by the conversion of the low part (an integer below 2^32: <= 2^8 units,
            // whatever the digits cancel to) plus one fma rounding.  The rounded d32, the score fma, the rounded g32
            // and each rounded cross term with its addition cost <= 2^-24 of a partial sum each, and every partial
            // sum is bounded by |vf d| + |g0| + sum |B| <= |s| + 2 C with C = max|g0| + sum max|B| (cabs): in total
            //     |s - s64| <= (6 + 2 l) 2^-24 (|s| + 2 C) + 2^8 max|d|  <=  eps(s) = 2^-20 (|s| + 2 C + 2^30 max|d|)
            // for the l <= 2 earlier levels (16 * 2^-24 against 10 * 2^-24, and 2^10 max|d| for the conversion term).
            // s - eps(s) is increasing in s, so the runner-up bounds every other column from below:
            //     second - eps(second) > best + eps(best)  =>  the fp32 argmin is the float64 argmin.
            // (Round 1 tracked max |vf| of the row for a uniform eps: one more ALU-pipe instruction per score.)
            struct FCand {
                float best, second;
                int idx;
                int pad;
            };
            // slot p - 1 of a row holds the candidates of part p; part 0 publishes the merged index in slot 0
            static_assert((kSplit - 1) * 128 * kSegSlots * 16 <= 8192, "merge scratch");
            FCand* fc = reinterpret_cast<FCand*>(scratch) + (quad * 32 + lane) * kSegSlots;
            constexpr int kPartStride = 128 * kSegSlots;
            int* qflag = reinterpret_cast<int*>(scratch + 8192);
            if (part != 0) {
#pragma unroll
                for (int s = 0; s < kMaxSeg; ++s)
                    if (s < nseg())
                        fc[(part - 1) * kPartStride + s] = FCand{ts.fbest[s], ts.fsecond[s], ts.bidx[s], 0};
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * kSplit) : "memory");
            if (part == 0) {
                bool unsafe = false;
#pragma unroll
                for (int s = 0; s < kMaxSeg; ++s)
                    if (s < nseg()) {
                        float best = ts.fbest[s], second = ts.fsecond[s];
                        int idx = ts.bidx[s];
#pragma unroll
                        for (int p = 1; p < kSplit; ++p) {       // ascending parts = ascending columns
                            const FCand o = fc[(p - 1) * kPartStride + s];
                            second = fminf(fminf(second, o.second), fmaxf(best, o.best));
                            idx = (o.best < best) ? o.idx : idx;
                            best = fminf(best, o.best);
                        }
                        const float k2 = 2.f * fmaf(1073741824.f, dmax[l], 2.f * cabs[l]);
                        const float eps2 = kEpsRel * (fabsf(second) + fabsf(best) + k2);     // eps(second) + eps(best)
                        unsafe |= !(second - best > eps2);           // also true for NaN / a single column
                        ts.bidx[s] = idx;
                        fc[s].idx = idx;
                    }
                if (dbg_flags(g) & 256) unsafe = false;
                const int any = __any_sync(0xffffffffu, unsafe) ? 1 : 0;
                if (lane == 0) qflag[quad] = any;
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * kSplit) : "memory");
            exact_merge = (qflag[quad] != 0);
            if (exact_merge) {
                // rare: some row of this quadrant sits in a near-tie -- evaluate the level exactly for all 32 rows
                // (the accumulators are still in TMEM because the slots are released only after step_end)
                redo_level_exact(ts, g, tc, step, tv);
                // part 0 may still be reading fc[] above when the other parts start overwriting the slots below
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * kSplit) : "memory");
            } else {
#pragma unroll
                for (int s = 0; s < kMaxSeg; ++s)
                    if (s < nseg() && part != 0) ts.bidx[s] = fc[s].idx;
            }
        }
        struct Cand {
            double best;
            int idx;
            int pad;
        };
        static_assert(sizeof(Cand) == 16, "Cand shares the FCand slots");
        Cand* cand = reinterpret_cast<Cand*>(scratch) + (quad * 32 + lane) * kSegSlots;
        constexpr int kCandStride = 128 * kSegSlots;
        if (exact_merge) {
            // merge the column parts of this row quadrant: parts 1.. publish their running minima, part 0 keeps
            // the lowest index on ties and publishes the final codes back (slot 0).
            if (part != 0) {
#pragma unroll
                for (int s = 0; s < kMaxSeg; ++s)
                    if (s < nseg())
                        cand[(part - 1) * kCandStride + s].best = ts.best[s],
                        cand[(part - 1) * kCandStride + s].idx = ts.bidx[s];
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * kSplit) : "memory");
            if (part == 0) {
#pragma unroll
                for (int s = 0; s < kMaxSeg; ++s)
                    if (s < nseg()) {
#pragma unroll
                        for (int p = 1; p < kSplit; ++p) {
                            const double ob = cand[(p - 1) * kCandStride + s].best;
                            const int oi = cand[(p - 1) * kCandStride + s].idx;
                            if (ob < ts.best[s] || (ob == ts.best[s] && oi < ts.bidx[s]))      // ties -> lowest
                                ts.best[s] = ob, ts.bidx[s] = oi;
                        }
                        cand[s].idx = ts.bidx[s];
                    }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * kSplit) : "memory");
            if (part != 0) {
#pragma unroll
                for (int s = 0; s < kMaxSeg; ++s)
                    if (s < nseg()) ts.bidx[s] = cand[s].idx;
            }
        }
#pragma unroll
        for (int s = 0; s < kMaxSeg; ++s) {
            if (s < nseg()) {
                const int code = ts.bidx[s];
#pragma unroll
                for (int ll = 0; ll < kMaxLevels; ++ll)
                    if (ll == l) ts.code[ll][s] = code;
                if (part == 0 && ts.r < rows)
                    codes[(static_cast<long long>(l) * m + s) * rows + ts.r] = static_cast<uint8_t>(code);
                ts.best[s] = INFINITY;
                ts.bidx[s] = 0;
                ts.fbest[s] = INFINITY;
                ts.fsecond[s] = INFINITY;
            }
        }
        if constexpr (FAST) {
            // The next level gathers one row of every cross-term table per earlier level, chosen by the codes that
            // are final as of now: pull this thread's slice of those rows into L1 while the tensor core is busy with
            // the next level's GEMM, so that the chunk loop does not expose the L2 latency of the gathers.
            if (l + 1 < levels && !(dbg_flags(g) & 1024)) {
                const int nl = n_level[l + 1];
#pragma unroll
                for (int jl = 0; jl < kMaxLevels - 1; ++jl)
                    if (jl <= l) {
#pragma unroll
                        for (int s2 = 0; s2 < kMaxSeg; ++s2)
                            if (s2 < nseg()) {
                                const float* bp = btab32 + boff[l + 1][jl] +
                                                  (static_cast<long long>(s2) * k[jl] + ts.code[jl][s2]) * nl;
                                for (int c = tv.c_begin; c < tv.c_end && c < nl; c += 32)
                                    asm volatile("prefetch.global.L1 [%0];" ::"l"(bp + c));
                            }
                    }
            }
        }
        // the other parts must have read the final codes before anyone overwrites the slots at the next level end
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * kSplit) : "memory");
    }

    __device__ __forceinline__ void end(Tile&, const IgemmGeom&, const TileCoord&) const {}
};

// Decode: out[r][:] = ((const + T_0[code_0]) + T_1[code_1]) + ... in fp32, fixed summation order.
//
// The tables (levels*m*k*C floats, 384 KB at k=128 C=256) would be re-read from L2 three times per output row if
// gathered from global memory (3x the bytes the kernel writes).  Instead the channel axis is cut into S slices so
// that one slice of ALL tables fits in shared memory (<= 200 KB); a persistent CTA owns one slice, stages it once,
// and then streams rows: 3 code bytes in, Cs*4 contiguous bytes out.  The kernel is then bound by the HBM write.
// Rows are either [0, rows) or the union of up to 8 rectangles of per-agent row grids (multi-GPU ego tiles).
constexpr int kMaxTables = kMaxLevels * kMaxSeg;
struct DecodeRegions {
    int n;                   // 0: plain rows [0, total)
    long long base[8];
    int y0[8], x0[8], w[8];
    long long start[9];      // prefix sums of the region sizes; start[n] (or start[0] when n == 0) = total rows
    int pitch;
};
struct DecodeTables {
    int nt, C, S, Cs;
    int k[kMaxTables];
    long long toff[kMaxTables];   // float offset of table i in the global tables array ([k][C])
    int soff[kMaxTables];         // float offset of table i's slice in shared memory ([k][Cs])
};

__device__ __forceinline__ long long decode_row_of(const DecodeRegions& rg, long long t) {
    if (rg.n == 0) return t;
    int a = 0;
    while (a + 1 < rg.n && t >= rg.start[a + 1]) ++a;
    const long long local = t - rg.start[a];
    const int yy = rg.y0[a] + static_cast<int>(local / rg.w[a]), xx = rg.x0[a] + static_cast<int>(local % rg.w[a]);
    return rg.base[a] + static_cast<long long>(yy) * rg.pitch + xx;
}

template <int NT>
__global__ void __launch_bounds__(1024, 1)
codebook_decode_kernel(const uint8_t* __restrict__ codes, long long plane_stride, const float* __restrict__ dconst,
                       const float* __restrict__ tables, float* __restrict__ out, const DecodeTables tb,
                       const DecodeRegions rg) {
    extern __shared__ float4 dsm4[];
    float* s_tab = reinterpret_cast<float*>(dsm4);
    const int split = blockIdx.x % tb.S;
    const int group = blockIdx.x / tb.S, ngroups = gridDim.x / tb.S;
    const int c0 = split * tb.Cs;
    const int vec = tb.Cs / 4;
    // stage this channel slice of every table (+ the constant as the last row)
    const int s_const = tb.soff[NT - 1] + tb.k[NT - 1] * tb.Cs;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        const float4* src = reinterpret_cast<const float4*>(tables + tb.toff[i] + c0);
        float4* dst = reinterpret_cast<float4*>(s_tab + tb.soff[i]);
        for (int idx = threadIdx.x; idx < tb.k[i] * vec; idx += blockDim.x) {
            const int row = idx / vec, v = idx - row * vec;
            dst[idx] = __ldg(src + static_cast<long long>(row) * (tb.C / 4) + v);
        }
    }
    for (int v = threadIdx.x; v < vec; v += blockDim.x)
        reinterpret_cast<float4*>(s_tab + s_const)[v] = __ldg(reinterpret_cast<const float4*>(dconst + c0) + v);
    __syncthreads();

    constexpr int R = NT <= 3 ? 4 : (NT <= 6 ? 2 : 1);   // rows per warp iteration (loads / stores in flight)
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const long long total = rg.start[rg.n];
    const long long nwarps = static_cast<long long>(ngroups) * wpb;
    const long long wid = static_cast<long long>(group) * wpb + wib;
    long long rowi[R];
    int code[R][NT];
    auto fetch = [&](long long t0) {
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const long long t = t0 + rr;
            rowi[rr] = (t < total) ? decode_row_of(rg, t) : -1;
#pragma unroll
            for (int i = 0; i < NT; ++i)
                code[rr][i] = (rowi[rr] >= 0) ? __ldg(codes + static_cast<long long>(i) * plane_stride + rowi[rr]) : 0;
        }
    };
    long long t0 = wid * R;
    if (t0 < total) fetch(t0);
    while (t0 < total) {
        long long crow[R];
        int soffs[R][NT];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            crow[rr] = rowi[rr];
#pragma unroll
            for (int i = 0; i < NT; ++i)      // an out-of-range code (corrupt input) must not index past the table
                soffs[rr][i] = tb.soff[i] + min(code[rr][i], tb.k[i] - 1) * tb.Cs;
        }
        const long long tn = t0 + nwarps * R;
        if (tn < total) fetch(tn);      // next iteration's codes are in flight while this one is gathered and stored
        for (int v = lane; v < vec; v += 32) {
            const float4 cst = reinterpret_cast<const float4*>(s_tab + s_const)[v];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                float4 a = cst;
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    const float4 t = reinterpret_cast<const float4*>(s_tab + soffs[rr][i])[v];
                    a.x = __fadd_rn(a.x, t.x);
                    a.y = __fadd_rn(a.y, t.y);
                    a.z = __fadd_rn(a.z, t.z);
                    a.w = __fadd_rn(a.w, t.w);
                }
                if (crow[rr] >= 0) reinterpret_cast<float4*>(out + crow[rr] * tb.C + c0)[v] = a;
            }
        }
        t0 = tn;
    }
}

// ----------------------------------------------------------------------------------------------- host math
struct Mat {
    int r = 0, c = 0;
    std::vector<double> v;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), v(static_cast<size_t>(r_) * c_, 0.0) {}
    double& at(int i, int j) { return v[static_cast<size_t>(i) * c + j]; }
    double at(int i, int j) const { return v[static_cast<size_t>(i) * c + j]; }
};
static Mat matmul(const Mat& a, const Mat& b) {
    Mat o(a.r, b.c);
    for (int i = 0; i < a.r; ++i)
        for (int k = 0; k < a.c; ++k) {
            const double x = a.at(i, k);
            if (x == 0.0) continue;
            const double* bp = &b.v[static_cast<size_t>(k) * b.c];
            double* op = &o.v[static_cast<size_t>(i) * o.c];
            for (int j = 0; j < b.c; ++j) op[j] += x * bp[j];
        }
    return o;
}
static std::vector<double> matvec(const Mat& a, const std::vector<double>& x) {
    std::vector<double> o(a.r, 0.0);
    for (int i = 0; i < a.r; ++i) {
        double s = 0.0;
        for (int j = 0; j < a.c; ++j) s += a.at(i, j) * x[j];
        o[i] = s;
    }
    return o;
}
static Mat eye(int n) {
    Mat o(n, n);
    for (int i = 0; i < n; ++i) o.at(i, i) = 1.0;
    return o;
}
static Mat from_f32(const float* p, int r, int c) {
    Mat o(r, c);
    for (size_t i = 0; i < o.v.size(); ++i) o.v[i] = p[i];
    return o;
}
static std::vector<double> vec_f32(const float* p, int n) {
    std::vector<double> o(n, 0.0);
    if (p) for (int i = 0; i < n; ++i) o[i] = p[i];
    return o;
}

}  // namespace qv2x

using namespace qv2x;

struct qv2x_codebook {
    qv2x_codebook_desc d;
    int C, m, levels, dseg;
    int k[kMaxLevels], n_level[kMaxLevels], colbase[kMaxLevels], rowbase[kMaxLevels];
    int block_n, bk, n_steps;
    long long boff[kMaxLevels][kMaxLevels];
    // host copies (for qv2x_codebook_folded_copy) and device buffers
    std::vector<int8_t> h_digits;
    std::vector<double> h_sc, h_g0, h_btab;
    std::vector<float> h_dconst, h_tables;
    std::vector<long long> h_toff;
    int8_t* d_digits = nullptr;
    double *d_sc = nullptr, *d_g0 = nullptr, *d_btab = nullptr;
    float *d_dconst = nullptr, *d_tables = nullptr;
    long long* d_toff = nullptr;
    // fp32 filter of the encoder: copies of g0 / btab, and delta * sc for the delta seen last (re-derived, with a
    // synchronous upload, only when a call passes a different delta)
    float *d_g32 = nullptr, *d_btab32 = nullptr, *d_d32 = nullptr;
    double* d_dsc = nullptr;      // fl64(delta * sc[j])
    float cached_delta = -1.f;
    float dmax[kMaxLevels] = {}, cabs[kMaxLevels] = {};
    bool fast = false;
};

extern "C" {

void qv2x_codebook_destroy(qv2x_codebook* cb) {
    if (!cb) return;
    cudaFree(cb->d_digits);
    cudaFree(cb->d_sc);
    cudaFree(cb->d_g0);
    cudaFree(cb->d_btab);
    cudaFree(cb->d_dconst);
    cudaFree(cb->d_tables);
    cudaFree(cb->d_toff);
    cudaFree(cb->d_g32);
    cudaFree(cb->d_btab32);
    cudaFree(cb->d_d32);
    cudaFree(cb->d_dsc);
    delete cb;
}

int qv2x_codebook_create(const qv2x_codebook_desc* desc, const float* const* codebooks, const float* const* weights,
                         const float* const* biases, qv2x_codebook** out) {
    QV2X_REQUIRE(desc && codebooks && weights && biases && out, "qv2x_codebook_create: null argument");
    QV2X_CHECK_SIZE(desc, qv2x_codebook_desc);
    const int C = desc->channel, m = desc->m, L = desc->levels;
    QV2X_REQUIRE(L >= 1 && L <= kMaxLevels, "levels must be 1..%d", kMaxLevels);
    QV2X_REQUIRE(m >= 1 && m <= kMaxSeg && C % m == 0, "m must be 1..%d and divide channel", kMaxSeg);
    QV2X_REQUIRE(C % 64 == 0, "channel must be a multiple of 64");
    for (int l = 0; l < L; ++l)
        QV2X_REQUIRE(desc->k[l] >= 16 && desc->k[l] <= 256 && desc->k[l] % 16 == 0,
                     "dict size must be a multiple of 16 in [16, 256] (codes are bytes)");
    const int d = C / m;
    auto W = [&](int l, int h) { return weights[l * 6 + h]; };
    auto Bv = [&](int l, int h) { return biases[l * 6 + h]; };
    enum { H_ENC = 0, H_QH = 1, H_LAT = 2, H_DEQ = 3, H_SIDE = 4, H_RES = 5 };
    for (int l = 0; l < L; ++l) {
        QV2X_REQUIRE(W(l, H_ENC) && W(l, H_QH) && W(l, H_DEQ) && W(l, H_RES) && codebooks[l],
                     "level %d: latentStageEncoder / quantizationHead / dequantizationHead / restoreHead required", l);
        if (l < L - 1) QV2X_REQUIRE(W(l, H_LAT) && W(l, H_SIDE), "level %d: latentHead and sideHead required", l);
    }
    auto cb = new qv2x_codebook();
    cb->d = *desc;
    cb->C = C;
    cb->m = m;
    cb->levels = L;
    cb->dseg = d;
    int ncols = 0, nrows = 0;
    for (int l = 0; l < L; ++l) {
        cb->k[l] = desc->k[l];
        cb->n_level[l] = m * desc->k[l];
        cb->colbase[l] = ncols;
        cb->rowbase[l] = nrows;
        ncols += cb->n_level[l];
        nrows += 3 * cb->n_level[l];
    }
    cb->bk = (C % 128 == 0) ? 128 : 64;
    // one BLOCK_N for all levels: 128 when every level's width allows it, else 64, else 32... keep 64/128
    cb->block_n = 128;
    for (int l = 0; l < L; ++l)
        if (cb->n_level[l] % 128 != 0) cb->block_n = 64;
    for (int l = 0; l < L; ++l)
        if (cb->n_level[l] % cb->block_n != 0) {
            delete cb;
            return set_error(QV2X_ERR_INVALID, "m * dict_size must be a multiple of 64 (level %d has %d)", l,
                             m * desc->k[l]);
        }
    cb->n_steps = 0;
    for (int l = 0; l < L; ++l) cb->n_steps += cb->n_level[l] / cb->block_n;
    if (cb->n_steps > kMaxSteps) {
        delete cb;
        return set_error(QV2X_ERR_INVALID, "too many N steps (%d > %d)", cb->n_steps, kMaxSteps);
    }

    // ---------------------------------------------------------------- fold the encoder (float64)
    cb->h_digits.assign(static_cast<size_t>(nrows) * C, 0);
    cb->h_sc.assign(ncols, 1.0);
    cb->h_g0.assign(ncols, 0.0);
    Mat A = eye(C);
    std::vector<double> a(C, 0.0);
    std::vector<Mat> P;
    long long bcount = 0;
    for (int l = 0; l < L; ++l)
        for (int j = 0; j < l; ++j) {
            cb->boff[l][j] = bcount;
            bcount += static_cast<long long>(m) * cb->k[j] * cb->n_level[l];
        }
    cb->h_btab.assign(static_cast<size_t>(bcount > 0 ? bcount : 1), 0.0);
    for (int l = 0; l < L; ++l) {
        const int k = cb->k[l], nl = cb->n_level[l];
        const Mat We = from_f32(W(l, H_ENC), C, C), Wq = from_f32(W(l, H_QH), C, C);
        const std::vector<double> be = vec_f32(Bv(l, H_ENC), C), bq = vec_f32(Bv(l, H_QH), C);
        const Mat M = matmul(Wq, We);
        std::vector<double> hb = matvec(Wq, be);
        for (int i = 0; i < C; ++i) hb[i] += bq[i];
        Mat Cmat(nl, C);
        std::vector<double> c2(nl, 0.0);
        for (int s = 0; s < m; ++s)
            for (int kk = 0; kk < k; ++kk)
                for (int t = 0; t < d; ++t) {
                    const double cv = codebooks[l][(static_cast<size_t>(s) * k + kk) * d + t];
                    Cmat.at(s * k + kk, s * d + t) = cv;
                    c2[s * k + kk] += cv * cv;
                }
        Mat T = matmul(Cmat, M);
        for (auto& x : T.v) x *= -2.0;
        const Mat G = matmul(T, A);
        const std::vector<double> Ta = matvec(T, a), Chb = matvec(Cmat, hb);
        for (int n = 0; n < nl; ++n) {
            cb->h_g0[cb->colbase[l] + n] = c2[n] - 2.0 * Chb[n] + Ta[n];
            double mx = 0.0;
            for (int i = 0; i < C; ++i) mx = std::max(mx, std::fabs(G.at(n, i)));
            const double sc = mx > 0.0 ? mx / kDigitMax : 1.0;
            cb->h_sc[cb->colbase[l] + n] = sc;
            for (int i = 0; i < C; ++i) {
                int8_t dg[3];
                split_digits(static_cast<long long>(std::nearbyint(G.at(n, i) / sc)), dg);
                for (int g3 = 0; g3 < 3; ++g3)
                    cb->h_digits[(static_cast<size_t>(cb->rowbase[l]) + static_cast<size_t>(g3) * nl + n) * C + i] =
                        dg[g3];
            }
        }
        for (int j = 0; j < l; ++j) {
            const Mat TP = matmul(T, P[j]);   // [nl, C]
            const int kj = cb->k[j];
            for (int s2 = 0; s2 < m; ++s2)
                for (int kk = 0; kk < kj; ++kk)
                    for (int n = 0; n < nl; ++n) {
                        double sum = 0.0;
                        for (int t = 0; t < d; ++t)
                            sum += codebooks[j][(static_cast<size_t>(s2) * kj + kk) * d + t] * TP.at(n, s2 * d + t);
                        cb->h_btab[cb->boff[l][j] + (static_cast<long long>(s2) * kj + kk) * nl + n] = -sum;
                    }
        }
        if (l < L - 1) {
            const Mat Wl = from_f32(W(l, H_LAT), C, C);
            const std::vector<double> bl = vec_f32(Bv(l, H_LAT), C);
            const Mat N = matmul(Wl, We);
            for (auto& Pj : P) Pj = matmul(N, Pj);
            P.push_back(eye(C));
            std::vector<double> na = matvec(N, a), wbe = matvec(Wl, be);
            for (int i = 0; i < C; ++i) a[i] = na[i] + wbe[i] + bl[i];
            A = matmul(N, A);
        }
    }
    // ---------------------------------------------------------------- fold the decoder (float64 -> fp32 tables)
    {
        std::vector<std::vector<Mat>> tabs(L);   // tabs[l][s]: [k, C]
        std::vector<double> cst;
        for (int l = L - 1; l >= 0; --l) {
            const Mat Wd = from_f32(W(l, H_DEQ), C, C), Wr = from_f32(W(l, H_RES), C, C);
            const std::vector<double> bd = vec_f32(Bv(l, H_DEQ), C), br = vec_f32(Bv(l, H_RES), C);
            const Mat RD = matmul(Wr, Wd);   // [C, C]
            std::vector<double> inner = bd;
            if (l < L - 1) {
                const Mat Ws = from_f32(W(l, H_SIDE), C, C);
                const std::vector<double> bs = vec_f32(Bv(l, H_SIDE), C);
                const Mat RS = matmul(Wr, Ws);
                for (int ll = l + 1; ll < L; ++ll)
                    for (auto& t : tabs[ll]) {
                        Mat nt(t.r, C);
                        for (int r = 0; r < t.r; ++r)
                            for (int i = 0; i < C; ++i) {
                                double sum = 0.0;
                                for (int j = 0; j < C; ++j) sum += t.at(r, j) * RS.at(i, j);
                                nt.at(r, i) = sum;
                            }
                        t = nt;
                    }
                const std::vector<double> wsc = matvec(Ws, cst);
                for (int i = 0; i < C; ++i) inner[i] += wsc[i] + bs[i];
            }
            cst = matvec(Wr, inner);
            for (int i = 0; i < C; ++i) cst[i] += br[i];
            tabs[l].resize(m);
            for (int s = 0; s < m; ++s) {
                Mat t(cb->k[l], C);
                for (int kk = 0; kk < cb->k[l]; ++kk)
                    for (int i = 0; i < C; ++i) {
                        double sum = 0.0;
                        for (int t2 = 0; t2 < d; ++t2)
                            sum += codebooks[l][(static_cast<size_t>(s) * cb->k[l] + kk) * d + t2] * RD.at(i, s * d + t2);
                        t.at(kk, i) = sum;
                    }
                tabs[l][s] = t;
            }
        }
        cb->h_dconst.resize(C);
        for (int i = 0; i < C; ++i) cb->h_dconst[i] = static_cast<float>(cst[i]);
        long long off = 0;
        for (int l = 0; l < L; ++l)
            for (int s = 0; s < m; ++s) {
                cb->h_toff.push_back(off);
                for (double x : tabs[l][s].v) cb->h_tables.push_back(static_cast<float>(x));
                off += static_cast<long long>(cb->k[l]) * C;
            }
    }
    int rc = upload(&cb->d_digits, cb->h_digits.data(), cb->h_digits.size());
    if (!rc) rc = upload(&cb->d_sc, cb->h_sc.data(), cb->h_sc.size());
    if (!rc) rc = upload(&cb->d_g0, cb->h_g0.data(), cb->h_g0.size());
    if (!rc) rc = upload(&cb->d_btab, cb->h_btab.data(), cb->h_btab.size());
    if (!rc) rc = upload(&cb->d_dconst, cb->h_dconst.data(), cb->h_dconst.size());
    if (!rc) rc = upload(&cb->d_tables, cb->h_tables.data(), cb->h_tables.size());
    if (!rc) rc = upload(&cb->d_toff, cb->h_toff.data(), cb->h_toff.size());
    {   // fp32 filter tables and the per-level magnitudes of its error bound (rounded up)
        std::vector<float> g32(cb->h_g0.size()), b32(cb->h_btab.size());
        for (size_t i = 0; i < g32.size(); ++i) g32[i] = static_cast<float>(cb->h_g0[i]);
        for (size_t i = 0; i < b32.size(); ++i) b32[i] = static_cast<float>(cb->h_btab[i]);
        if (!rc) rc = upload(&cb->d_g32, g32.data(), g32.size());
        if (!rc) rc = upload(&cb->d_btab32, b32.data(), b32.size());
        if (!rc && (cudaMalloc(reinterpret_cast<void**>(&cb->d_d32), cb->h_sc.size() * sizeof(float)) != cudaSuccess ||
                    cudaMalloc(reinterpret_cast<void**>(&cb->d_dsc), cb->h_sc.size() * sizeof(double)) != cudaSuccess))
            rc = set_error(QV2X_ERR_CUDA, "cudaMalloc failed");
        for (int l = 0; l < L; ++l) {
            double c = 0.0;
            for (int n = 0; n < cb->n_level[l]; ++n) c = std::max(c, std::fabs(cb->h_g0[cb->colbase[l] + n]));
            for (int j = 0; j < l; ++j) {
                double bm = 0.0;
                const long long cnt = static_cast<long long>(m) * cb->k[j] * cb->n_level[l];
                for (long long i = 0; i < cnt; ++i) bm = std::max(bm, std::fabs(cb->h_btab[cb->boff[l][j] + i]));
                c += bm * m;     // one row per segment of level j is added
            }
            cb->cabs[l] = std::nextafter(static_cast<float>(c * (1.0 + 1e-6)), INFINITY);
        }
        // the level's accumulators must still be in TMEM at the end of the level: one N step per level
        cb->fast = (cb->n_steps == L);
    }
    if (rc) {
        qv2x_codebook_destroy(cb);
        return rc;
    }
    *out = cb;
    return 0;
}

int qv2x_codebook_encode(const qv2x_codebook* cb, long long rows, const uint8_t* d_feat, int feat_cstride,
                         float delta, uint8_t* d_codes, void* stream_) {
    QV2X_REQUIRE(cb && d_feat && d_codes, "qv2x_codebook_encode: null argument");
    if (rows <= 0) return 0;
    QV2X_REQUIRE(rows < (1ll << 31), "too many rows");
    QV2X_REQUIRE(feat_cstride >= cb->C && feat_cstride % 16 == 0, "bad feature stride");
    QV2X_REQUIRE(delta > 0.f, "delta must be positive");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    IgemmGeom g{};
    g.n_img = 1;
    g.Ho = 1;
    g.Wo = static_cast<int>(rows);
    g.Hi = 1;
    g.Wi = static_cast<int>(rows);
    g.tw = 128;
    g.th = 1;
    g.tiles_x = static_cast<int>((rows + 127) / 128);
    g.tiles_y = 1;
    g.n_tiles = 1;
    g.block_n = cb->block_n;
    g.taps = 1;
    g.taps_w = 1;
    g.stride = 1;
    g.pad = 0;
    g.groups = 3;
    g.cblocks = cb->C / cb->bk;
    g.idesc = make_idesc_i8(cb->block_n, true);
    g.n_steps = cb->n_steps;
    const bool fast = cb->fast && !(g_debug_flags & 64);       // debug 64: force the float64 path
    auto run = [&](auto e) -> int {
    e.levels = cb->levels;
        e.m = cb->m;
        int step = 0;
        for (int l = 0; l < cb->levels; ++l) {
            e.k[l] = cb->k[l];
            e.n_level[l] = cb->n_level[l];
            e.colbase[l] = cb->colbase[l];
            for (int j = 0; j < kMaxLevels; ++j) e.boff[l][j] = cb->boff[l][j];
            const int spl = cb->n_level[l] / cb->block_n;
            for (int c = 0; c < spl; ++c, ++step) {
                g.step_row_base[step] = cb->rowbase[l] + c * cb->block_n;
                g.step_group_stride[step] = cb->n_level[l];
                e.step_level[step] = l;
                e.step_col0[step] = c * cb->block_n;
                e.step_last[step] = (c == spl - 1);
            }
        }
        if (delta != cb->cached_delta) {
            auto* mcb = const_cast<qv2x_codebook*>(cb);
            std::vector<float> d32(cb->h_sc.size());
            std::vector<double> dsc(cb->h_sc.size());
            for (size_t i = 0; i < d32.size(); ++i) {
                dsc[i] = static_cast<double>(delta) * cb->h_sc[i];       // one float64 rounding, as the oracle does
                d32[i] = static_cast<float>(dsc[i]);
            }
            QV2X_CUDA_OK(cudaMemcpy(cb->d_dsc, dsc.data(), dsc.size() * sizeof(double), cudaMemcpyHostToDevice));
            for (int l = 0; l < cb->levels; ++l) {
                float mx = 0.f;
                for (int n = 0; n < cb->n_level[l]; ++n) mx = std::max(mx, std::fabs(d32[cb->colbase[l] + n]));
                mcb->dmax[l] = mx;
            }
            QV2X_CUDA_OK(cudaMemcpy(cb->d_d32, d32.data(), d32.size() * sizeof(float), cudaMemcpyHostToDevice));
            mcb->cached_delta = delta;
        }
        e.pack16 = (cb->C <= 256) ? 1 : 0;
        e.d32 = cb->d_d32;
        e.g32 = cb->d_g32;
        e.btab32 = cb->d_btab32;
        for (int l = 0; l < cb->levels; ++l) {
            e.dmax[l] = cb->dmax[l];
            e.cabs[l] = cb->cabs[l];
        }
        e.delta = static_cast<double>(delta);
        e.dsc = cb->d_dsc;
        e.g0 = cb->d_g0;
        e.btab = cb->d_btab;
        e.codes = d_codes;
        e.rows = rows;
        CUtensorMap tmA, tmB;
        int rc = make_act_tmap(&tmA, d_feat, 1, 1, static_cast<int>(rows), feat_cstride, 128, 1, 1, cb->bk);
        if (rc) return rc;
        int total_rows = 0;
        for (int l = 0; l < cb->levels; ++l) total_rows += 3 * cb->n_level[l];
        rc = make_weight_tmap(&tmB, cb->d_digits, total_rows, cb->C, cb->block_n, cb->bk);
        if (rc) return rc;
        return dispatch_igemm<3>(cb->block_n, cb->bk, tmA, tmB, g, e, stream);
    };
    if (cb->m == 1) return fast ? run(EncodeEpilogue<true, 1>{}) : run(EncodeEpilogue<false, 1>{});
    if (cb->m == 2) return fast ? run(EncodeEpilogue<true, 2>{}) : run(EncodeEpilogue<false, 2>{});
    return fast ? run(EncodeEpilogue<true, 0>{}) : run(EncodeEpilogue<false, 0>{});
}

// Shared launcher of the two decode entry points.
static int launch_decode(const qv2x_codebook* cb, long long plane_stride, const DecodeRegions& rg,
                         const uint8_t* d_codes, float* d_out, cudaStream_t stream) {
    const int nt = cb->levels * cb->m;
    QV2X_REQUIRE(nt >= 1 && nt <= kMaxTables, "too many code planes");
    QV2X_REQUIRE(cb->C % 4 == 0, "C must be a multiple of 4");
    DecodeTables tb{};
    tb.nt = nt;
    tb.C = cb->C;
    long long krows = 1;       // + the constant row
    for (int l = 0; l < cb->levels; ++l)
        for (int sgm = 0; sgm < cb->m; ++sgm) {
            tb.k[l * cb->m + sgm] = cb->k[l];
            tb.toff[l * cb->m + sgm] = cb->h_toff[l * cb->m + sgm];
            krows += cb->k[l];
        }
    // smallest number of channel slices whose table slice fits in 200 KB (slice width a multiple of 4 floats)
    int S = 0;
    for (int cand = 1; cand <= cb->C / 4; ++cand) {
        if (cb->C % (4 * cand) != 0) continue;
        if (krows * (cb->C / cand) * 4 <= 200 * 1024) {
            S = cand;
            break;
        }
    }
    QV2X_REQUIRE(S > 0 && S <= num_sms(), "decode tables do not fit in shared memory at any channel split");
    tb.S = S;
    tb.Cs = cb->C / S;
    int off = 0;
    for (int i = 0; i < nt; ++i) {
        tb.soff[i] = off;
        off += tb.k[i] * tb.Cs;
    }
    const int smem = static_cast<int>(krows * tb.Cs * 4);
    const long long total = rg.start[rg.n];
    const int threads = 1024;
    // persistent: one CTA per SM, a whole number of slice groups; small inputs use fewer groups
    const long long want_groups = (total + (threads / 32) * 4 - 1) / ((threads / 32) * 4);
    const int groups = static_cast<int>(std::max<long long>(1, std::min<long long>(num_sms() / S, want_groups)));
    const int grid = groups * S;
#define QV2X_DECODE_CASE(N)                                                                                        \
    case N: {                                                                                                      \
        static bool attr = false;                                                                                  \
        if (!attr) {                                                                                               \
            QV2X_CUDA_OK(cudaFuncSetAttribute(codebook_decode_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                              208 * 1024));                                                        \
            attr = true;                                                                                           \
        }                                                                                                          \
        codebook_decode_kernel<N><<<grid, threads, smem, stream>>>(d_codes, plane_stride, cb->d_dconst, cb->d_tables, \
                                                                   d_out, tb, rg);                                 \
        break;                                                                                                     \
    }
    switch (nt) {
        QV2X_DECODE_CASE(1) QV2X_DECODE_CASE(2) QV2X_DECODE_CASE(3) QV2X_DECODE_CASE(4) QV2X_DECODE_CASE(5)
        QV2X_DECODE_CASE(6) QV2X_DECODE_CASE(7) QV2X_DECODE_CASE(8) QV2X_DECODE_CASE(9) QV2X_DECODE_CASE(10)
        QV2X_DECODE_CASE(11) QV2X_DECODE_CASE(12) QV2X_DECODE_CASE(13) QV2X_DECODE_CASE(14) QV2X_DECODE_CASE(15)
        QV2X_DECODE_CASE(16)
    }
#undef QV2X_DECODE_CASE
    g_launch_count.fetch_add(1);
    QV2X_CUDA_OK(cudaGetLastError());
    return 0;
}

int qv2x_codebook_decode(const qv2x_codebook* cb, long long rows, const uint8_t* d_codes, float* d_out, void* stream_) {
    QV2X_REQUIRE(cb && d_codes && d_out, "qv2x_codebook_decode: null argument");
    if (rows <= 0) return 0;
    DecodeRegions rg{};
    rg.n = 0;
    rg.start[0] = rows;
    return launch_decode(cb, rows, rg, d_codes, d_out, static_cast<cudaStream_t>(stream_));
}

int qv2x_codebook_decode_regions(const qv2x_codebook* cb, long long plane_stride, int pitch, int n_regions,
                                 const long long* base_row, const int* rect, const uint8_t* d_codes, float* d_out,
                                 void* stream_) {
    QV2X_REQUIRE(cb && base_row && rect && d_codes && d_out, "qv2x_codebook_decode_regions: null argument");
    QV2X_REQUIRE(n_regions >= 1 && n_regions <= 8, "1..8 regions");
    DecodeRegions rg{};
    rg.pitch = pitch;
    rg.start[0] = 0;
    int n = 0;
    for (int a = 0; a < n_regions; ++a) {
        const int y0 = rect[4 * a + 0], y1 = rect[4 * a + 1], x0 = rect[4 * a + 2], x1 = rect[4 * a + 3];
        QV2X_REQUIRE(y1 >= y0 && x1 >= x0 && y0 >= 0 && x0 >= 0 && x1 <= pitch, "bad region %d", a);
        if (y1 == y0 || x1 == x0) continue;     // empty regions are dropped (the row mapping divides by the width)
        rg.base[n] = base_row[a];
        rg.y0[n] = y0;
        rg.x0[n] = x0;
        rg.w[n] = x1 - x0;
        rg.start[n + 1] = rg.start[n] + static_cast<long long>(y1 - y0) * (x1 - x0);
        ++n;
    }
    if (n == 0) return 0;
    rg.n = n;
    return launch_decode(cb, plane_stride, rg, d_codes, d_out, static_cast<cudaStream_t>(stream_));
}

int qv2x_codebook_desc_get(const qv2x_codebook* cb, qv2x_codebook_desc* desc) {
    QV2X_REQUIRE(cb && desc, "qv2x_codebook_desc_get: null argument");
    QV2X_CHECK_SIZE(desc, qv2x_codebook_desc);
    *desc = cb->d;
    return 0;
}

long long qv2x_codebook_folded_size(const qv2x_codebook* cb, int which) {
    if (!cb) return -1;
    switch (which) {
        case 0: return static_cast<long long>(cb->h_digits.size());
        case 1: return static_cast<long long>(cb->h_sc.size());
        case 2: return static_cast<long long>(cb->h_g0.size());
        case 3: return static_cast<long long>(cb->h_btab.size());
        case 4: return static_cast<long long>(cb->h_dconst.size());
        case 5: return static_cast<long long>(cb->h_tables.size());
        default: return -1;
    }
}

int qv2x_codebook_folded_copy(const qv2x_codebook* cb, int which, void* host_buf) {
    QV2X_REQUIRE(cb && host_buf, "qv2x_codebook_folded_copy: null argument");
    switch (which) {
        case 0: memcpy(host_buf, cb->h_digits.data(), cb->h_digits.size()); break;
        case 1: memcpy(host_buf, cb->h_sc.data(), cb->h_sc.size() * sizeof(double)); break;
        case 2: memcpy(host_buf, cb->h_g0.data(), cb->h_g0.size() * sizeof(double)); break;
        case 3: memcpy(host_buf, cb->h_btab.data(), cb->h_btab.size() * sizeof(double)); break;
        case 4: memcpy(host_buf, cb->h_dconst.data(), cb->h_dconst.size() * sizeof(float)); break;
        case 5: memcpy(host_buf, cb->h_tables.data(), cb->h_tables.size() * sizeof(float)); break;
        default: return set_error(QV2X_ERR_INVALID, "unknown table id %d", which);
    }
    return 0;
}

}  // extern "C"
