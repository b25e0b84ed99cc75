// Ego-side kernels: warp neighbours into the ego frame + fuse (max / attention), detection heads,
// and the layout / quantization converters used at the drop-in module boundaries.  All HBM-bound.
//
//   warp   : reference warp_affine_simple (opencood/models/sub_modules/torch_transformation_utils.py:323-332)
//            = F.affine_grid + F.grid_sample(bilinear, zeros, align_corners=False)
//   max    : reference MaxFusion.forward  (opencood/models/fuse_modules/fusion_in_one.py:87-124)
//   att    : reference AttFusion.forward  (fusion_in_one.py:126-151) -- only the ego query row is kept by the
//            reference (`[0, ...]`), so only that row is computed: out = sum_j softmax_j(x0.xj/sqrt(C)) xj
//   weighted : reference weighted_fuse (opencood/models/fuse_modules/pyramid_fuse.py:17-62) as called per pyramid level
//            by QuantPyramidFusion.forward_collab (opencood/quant/quant_block.py:516-539): the per-agent occupancy
//            score sigmoid(occ) + 1e-4 is warped like the features, a warped score of exactly 0 (no tap inside the
//            agent's map) excludes the agent, out = sum_j softmax_j(score_j) x_j
//   heads  : cls/reg/dir 1x1 convs with fake-quant weights on FP32 features (quant_model.py:129-136)
#include <algorithm>
#include <cmath>
#include <vector>

#include "host_common.h"
#include "ptx.cuh"

namespace qv2x {

constexpr int kMaxAgents = 8;

struct FuseParams {
    int n, H, W, C;
    int y0, x0, th, tw;         // output tile (rows [y0, y0+th), columns [x0, x0+tw)); the result is stored compactly
    const float* aff;           // DEVICE [n][6]: row-major 2x3, normalized coordinates (ego <- agent j)
    float inv_sqrt_c;
    const float* score;         // MODE 2: DEVICE [n][H][W] occupancy logits (score_is_logit) or ready-made scores
    int score_is_logit;
    float u8_delta;             // U8 input: features are uint8 codes of scale u8_delta (value = fl(delta * code))
};

// One warp per output pixel; lane owns float4 chunks v = lane + 32*t of the channel vector.
// NA (agents) and VPL are compile-time so that every load of a pixel -- NA agents x 4 bilinear taps x VPL chunks --
// is independent of every branch: out-of-image taps read a clamped address with weight 0 (adds +-0, exact), and
// the compiler can keep all of an agent's loads (and the next agent's) in flight together.  The kernel is
// HBM/L2-latency bound, so memory-level parallelism per warp is what sets its speed.
// U8: `feat` points at uint8 codes [n][H][W][C] instead of floats; a lane's chunk of four channels is one 32-bit word,
// de-quantized on load exactly as qv2x_dequant_u8 does (fl(delta * code)), so the result equals dequantize + fuse bit
// for bit while reading a quarter of the bytes and skipping the FP32 copy of every agent's map.
template <int MODE, int NA, int VPL /* float4 chunks per lane */, bool U8 = false>
__global__ void __launch_bounds__(256, (VPL > 2 ? 1 : 2)) fuse_kernel(const float* __restrict__ feat, float* __restrict__ out,
                                                   const FuseParams p) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const long long npix = static_cast<long long>(p.H) * p.W;
    const long long tpix = static_cast<long long>(p.th) * p.tw;
    const int vec = p.C / 4;
    __shared__ float am[kMaxAgents][6];
    if (threadIdx.x < NA * 6) am[threadIdx.x / 6][threadIdx.x % 6] = __ldg(p.aff + threadIdx.x);
    __syncthreads();
    for (long long pix = warp; pix < tpix; pix += nwarps) {
        const int ti = static_cast<int>(pix / p.tw);
        const int i = p.y0 + ti, j = p.x0 + static_cast<int>(pix - static_cast<long long>(ti) * p.tw);
        const float xn = (2.f * j + 1.f) / p.W - 1.f;
        const float yn = (2.f * i + 1.f) / p.H - 1.f;
        // Sampling geometry: lane a (< NA) works out agent a's four tap offsets and weights ONCE; the other lanes
        // receive them by shuffle (every lane computing all NA agents' geometry was a third of the kernel's
        // instructions).  Out-of-image taps get a clamped offset and weight 0.
        float tw_[4] = {0.f, 0.f, 0.f, 0.f};
        int to_[4] = {0, 0, 0, 0};
        float wscore = 0.f;                                          // MODE 2: agent `lane`'s warped score
        if (lane < NA) {
            const int a = lane;
            const float xs = am[a][0] * xn + am[a][1] * yn + am[a][2];
            const float ys = am[a][3] * xn + am[a][4] * yn + am[a][5];
            const float ix = ((xs + 1.f) * p.W - 1.f) * 0.5f;
            const float iy = ((ys + 1.f) * p.H - 1.f) * 0.5f;
            // clamp far-away coordinates before the int conversion (every tap is out of the image there anyway)
            const float fx = fminf(fmaxf(floorf(ix), -2.f), static_cast<float>(p.W) + 1.f);
            const float fy = fminf(fmaxf(floorf(iy), -2.f), static_cast<float>(p.H) + 1.f);
            const float tx = ix - fx, ty = iy - fy;
            const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
#pragma unroll
            for (int tap = 0; tap < 4; ++tap) {
                const int xx = x0 + (tap & 1), yy = y0 + (tap >> 1);
                const bool inb = (xx >= 0 && xx < p.W && yy >= 0 && yy < p.H);
                const float wt = ((tap & 1) ? tx : 1.f - tx) * ((tap >> 1) ? ty : 1.f - ty);
                tw_[tap] = inb ? wt : 0.f;
                const int xc = min(max(xx, 0), p.W - 1), yc = min(max(yy, 0), p.H - 1);
                to_[tap] = (yc * p.W + xc) * vec;                    // float4 offset inside the agent's map
                if (MODE == 2) {
                    float sv = __ldg(p.score + static_cast<long long>(a) * npix + yc * p.W + xc);
                    if (p.score_is_logit) sv = 1.f / (1.f + expf(-sv)) + 1e-4f;
                    wscore = fmaf(tw_[tap], sv, wscore);
                }
            }
        }
        float4 xa[NA][VPL];
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            float4 s[4][VPL];
            float w[4];
            if constexpr (U8) {
                const uint32_t* base = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(feat) +
                                                                         static_cast<long long>(a) * npix * p.C);
#pragma unroll
                for (int tap = 0; tap < 4; ++tap) {
                    w[tap] = __shfl_sync(0xffffffffu, tw_[tap], a);
                    const uint32_t* src = base + __shfl_sync(0xffffffffu, to_[tap], a);
#pragma unroll
                    for (int t = 0; t < VPL; ++t) {
                        const int v = lane + 32 * t;
                        const uint32_t q = (v < vec) ? __ldg(src + v) : 0u;
                        s[tap][t] = make_float4(__fmul_rn(static_cast<float>(q & 0xffu), p.u8_delta),
                                                __fmul_rn(static_cast<float>((q >> 8) & 0xffu), p.u8_delta),
                                                __fmul_rn(static_cast<float>((q >> 16) & 0xffu), p.u8_delta),
                                                __fmul_rn(static_cast<float>(q >> 24), p.u8_delta));
                    }
                }
            } else {
            const float4* base = reinterpret_cast<const float4*>(feat + static_cast<long long>(a) * npix * p.C);
#pragma unroll
            for (int tap = 0; tap < 4; ++tap) {
                w[tap] = __shfl_sync(0xffffffffu, tw_[tap], a);
                const float4* src = base + __shfl_sync(0xffffffffu, to_[tap], a);
#pragma unroll
                for (int t = 0; t < VPL; ++t) {
                    const int v = lane + 32 * t;
                    s[tap][t] = (v < vec) ? __ldg(src + v) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            }
#pragma unroll
            for (int t = 0; t < VPL; ++t) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int tap = 0; tap < 4; ++tap) {
                    acc.x = fmaf(w[tap], s[tap][t].x, acc.x);
                    acc.y = fmaf(w[tap], s[tap][t].y, acc.y);
                    acc.z = fmaf(w[tap], s[tap][t].z, acc.z);
                    acc.w = fmaf(w[tap], s[tap][t].w, acc.w);
                }
                xa[a][t] = acc;
            }
        }
        float4 o[VPL];
        if (MODE == 0) {
#pragma unroll
            for (int t = 0; t < VPL; ++t) o[t] = xa[0][t];
#pragma unroll
            for (int a = 1; a < NA; ++a) {
#pragma unroll
                for (int t = 0; t < VPL; ++t) {
                    o[t].x = fmaxf(o[t].x, xa[a][t].x);
                    o[t].y = fmaxf(o[t].y, xa[a][t].y);
                    o[t].z = fmaxf(o[t].z, xa[a][t].z);
                    o[t].w = fmaxf(o[t].w, xa[a][t].w);
                }
            }
        } else if (MODE == 2) {
            float sc[NA];
            float mx = -INFINITY;
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const float sv = __shfl_sync(0xffffffffu, wscore, a);
                sc[a] = (sv == 0.f) ? -INFINITY : sv;              // masked_fill_(scores == 0, -inf)
                mx = fmaxf(mx, sc[a]);
            }
            float den = 0.f;
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                sc[a] = (mx == -INFINITY) ? 0.f : expf(sc[a] - mx);  // all agents excluded: softmax NaN -> 0
                den += sc[a];
            }
#pragma unroll
            for (int t = 0; t < VPL; ++t) o[t] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const float w = den > 0.f ? sc[a] / den : 0.f;
#pragma unroll
                for (int t = 0; t < VPL; ++t) {
                    o[t].x += w * xa[a][t].x;
                    o[t].y += w * xa[a][t].y;
                    o[t].z += w * xa[a][t].z;
                    o[t].w += w * xa[a][t].w;
                }
            }
        } else {
            float sc[NA];
            float mx = -INFINITY;
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                float d = 0.f;
#pragma unroll
                for (int t = 0; t < VPL; ++t)
                    d += xa[0][t].x * xa[a][t].x + xa[0][t].y * xa[a][t].y + xa[0][t].z * xa[a][t].z +
                         xa[0][t].w * xa[a][t].w;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
                sc[a] = d * p.inv_sqrt_c;
                mx = fmaxf(mx, sc[a]);
            }
            float den = 0.f;
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                sc[a] = expf(sc[a] - mx);
                den += sc[a];
            }
#pragma unroll
            for (int t = 0; t < VPL; ++t) o[t] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const float w = sc[a] / den;
#pragma unroll
                for (int t = 0; t < VPL; ++t) {
                    o[t].x += w * xa[a][t].x;
                    o[t].y += w * xa[a][t].y;
                    o[t].z += w * xa[a][t].z;
                    o[t].w += w * xa[a][t].w;
                }
            }
        }
        float4* dst = reinterpret_cast<float4*>(out + pix * p.C);
#pragma unroll
        for (int t = 0; t < VPL; ++t) {
            const int v = lane + 32 * t;
            if (v < vec) dst[v] = o[t];
        }
    }
}

template <int MODE, int VPL, bool U8 = false>
static void launch_fuse(int n, int grid, int threads, cudaStream_t stream, const float* feat, float* out,
                        const FuseParams& p) {
    switch (n) {
        case 1: fuse_kernel<MODE, 1, VPL, U8><<<grid, threads, 0, stream>>>(feat, out, p); break;
        case 2: fuse_kernel<MODE, 2, VPL, U8><<<grid, threads, 0, stream>>>(feat, out, p); break;
        case 3: fuse_kernel<MODE, 3, VPL, U8><<<grid, threads, 0, stream>>>(feat, out, p); break;
        case 4: fuse_kernel<MODE, 4, VPL, U8><<<grid, threads, 0, stream>>>(feat, out, p); break;
        case 5: fuse_kernel<MODE, 5, VPL, U8><<<grid, threads, 0, stream>>>(feat, out, p); break;
        case 6: fuse_kernel<MODE, 6, VPL, U8><<<grid, threads, 0, stream>>>(feat, out, p); break;
        case 7: fuse_kernel<MODE, 7, VPL, U8><<<grid, threads, 0, stream>>>(feat, out, p); break;
        default: fuse_kernel<MODE, 8, VPL, U8><<<grid, threads, 0, stream>>>(feat, out, p); break;
    }
}

// ------------------------------------------------------------------------------------------ heads
// out[o][p] = bias[o] + sum_k x[p][k] * w[o][k] (k ascending, one fma per term: the same order for every pixel, so
// results do not depend on how the map is tiled).  Persistent CTAs: the k-major weight matrix [C][72] is staged in
// shared memory once, pixel tiles of 64 x C are double-buffered with cp.async so that the next tile's HBM read
// overlaps the current tile's math.  Warp w owns outputs 8w..8w+7, lane l pixels l and l+32.  The inner product
// runs on packed fp32 pairs (FFMA2: two IEEE fp32 fmas per instruction, the pixel value as the broadcast operand).
constexpr int kHeadsPix = 64, kHeadsOut = 72, kHeadsPitchPad = 4, kHeadsThreads = 288;

// Quantizing output of the FP32 GEMM (QOUT instantiations): the transposed convs of the pyramid path (kernel = stride
// = s) on FP32 fused maps.  Column o of the GEMM is (dy, dx, c) = (o / (s * cs), (o / cs) % s, o % cs) with cs the
// conv's output channels; input pixel (y, x) of a w-wide map lands at output pixel (y s + dy, x s + dx) of the uint8
// NHWC concat buffer (channel stride cstride, first channel cbase).  A thread owns 8 consecutive columns = 8
// consecutive channels of one sub-position: one 8-byte store per pixel.  The codes equal the quantizing converter's
// (IEEE division): the product with fl(1 / delta) is redone exactly within 1e-3 of a rounding boundary.
struct HeadsQOut {
    uint8_t* out;
    float delta, inv_delta;
    int s, cs, w, cstride, cbase;
};
__device__ __forceinline__ unsigned heads_quant_u8(float v, float delta, float inv_delta) {
    const float t = v * inv_delta;
    float r = rintf(t);
    if (fabsf(t - r) > 0.499f) r = rintf(__fdiv_rn(v, delta));
    return static_cast<unsigned>(fminf(fmaxf(r, 0.f), 255.f));
}
// 8 columns starting at global column gcol0 of input pixel pp -> 8 bytes
__device__ __forceinline__ void heads_store_q8(const HeadsQOut& q, long long pp, int gcol0, const float (&v)[8]) {
    const int sub = gcol0 / q.cs, c0 = gcol0 - sub * q.cs;
    const int dy = sub / q.s, dx = sub - dy * q.s;
    const long long y = pp / q.w, x = pp - y * q.w;
    const long long opix = (y * q.s + dy) * (static_cast<long long>(q.w) * q.s) + x * q.s + dx;
    uint2 pk;
    pk.x = heads_quant_u8(v[0], q.delta, q.inv_delta) | (heads_quant_u8(v[1], q.delta, q.inv_delta) << 8) |
           (heads_quant_u8(v[2], q.delta, q.inv_delta) << 16) | (heads_quant_u8(v[3], q.delta, q.inv_delta) << 24);
    pk.y = heads_quant_u8(v[4], q.delta, q.inv_delta) | (heads_quant_u8(v[5], q.delta, q.inv_delta) << 8) |
           (heads_quant_u8(v[6], q.delta, q.inv_delta) << 16) | (heads_quant_u8(v[7], q.delta, q.inv_delta) << 24);
    *reinterpret_cast<uint2*>(q.out + opix * q.cstride + q.cbase + c0) = pk;
}

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}

template <bool QOUT>
__global__ void __launch_bounds__(kHeadsThreads) heads_kernel(const float* __restrict__ x, const float* __restrict__ wt,
                                                              const float* __restrict__ bias, float* __restrict__ out,
                                                              long long npix, int C, int cout, int tile_w,
                                                              long long out_w, long long out_npix,
                                                              const HeadsQOut qo) {
    extern __shared__ float4 hsm4[];
    float* hsm = reinterpret_cast<float*>(hsm4);
    // blockIdx.y = chunk of 72 outputs (wider GEMMs: the FP32 deblocks of the pyramid path): own weight slab, bias
    // slice and output rows
    wt += static_cast<long long>(blockIdx.y) * C * kHeadsOut;
    if (bias) bias += blockIdx.y * kHeadsOut;
    out += static_cast<long long>(blockIdx.y) * kHeadsOut * out_npix;
    cout = min(cout - static_cast<int>(blockIdx.y) * kHeadsOut, kHeadsOut);
    const int pitch = C + kHeadsPitchPad;
    float* ws = hsm;                                   // [C][72]   k-major, outputs contiguous (pairs for FFMA2)
    float* xs0 = hsm + C * kHeadsOut;                  // 2 x [64][pitch] pixel-major
    const int tid = threadIdx.x;
    const int vec = C / 4;
    const long long ntiles = (npix + kHeadsPix - 1) / kHeadsPix;

    auto load_tile = [&](long long t, int buf) {
        const long long p0 = t * kHeadsPix;
        const uint32_t dst = smem_u32(xs0 + buf * kHeadsPix * pitch);
        for (int idx = tid; idx < kHeadsPix * vec; idx += kHeadsThreads) {
            const int pp = idx / vec, v = idx - pp * vec;
            const bool ok = (p0 + pp < npix);
            cp_async_16(dst + (pp * pitch + 4 * v) * 4, ok ? x + (p0 + pp) * C + 4 * v : x, ok);
        }
    };
    for (int idx = tid; idx < C * kHeadsOut / 4; idx += kHeadsThreads)
        cp_async_16(smem_u32(ws) + idx * 16, wt + idx * 4, true);
    if (blockIdx.x < ntiles) load_tile(blockIdx.x, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");

    const int pg = tid & 31, og = tid >> 5;    // pixels pg, pg + 32; outputs og*8 + 2*j, og*8 + 2*j + 1
    int buf = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, buf ^= 1) {
        if (t + gridDim.x < ntiles) load_tile(t + gridDim.x, buf ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");       // everything but the tile just requested
        __syncthreads();
        const float* xs = xs0 + buf * kHeadsPix * pitch;
        f32x2 acc[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = pack2(0.f, 0.f);
        // unrolled so that the shared-memory loads of the next k groups are in flight under the FMAs of the current one
        // (nine warps per SM cannot hide the LDS latency by themselves)
#pragma unroll 4
        for (int k = 0; k < C; k += 4) {
            float4 xv[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + (pg + 32 * i) * pitch + k);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 w01 = *reinterpret_cast<const float4*>(ws + (k + kk) * kHeadsOut + og * 8);
                const float4 w23 = *reinterpret_cast<const float4*>(ws + (k + kk) * kHeadsOut + og * 8 + 4);
                const f32x2 wp[4] = {pack2(w01.x, w01.y), pack2(w01.z, w01.w), pack2(w23.x, w23.y),
                                     pack2(w23.z, w23.w)};
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float xk = (kk == 0) ? xv[i].x : (kk == 1) ? xv[i].y : (kk == 2) ? xv[i].z : xv[i].w;
                    const f32x2 xx = pack2(xk, xk);
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fma2(xx, wp[j], acc[i][j]);
                }
            }
        }
        const long long p0 = t * kHeadsPix;
        // pixel pp of the (compact, tile_w wide) input lands at row pp / tile_w, column pp % tile_w of an out_w wide
        // map (out_w == tile_w: the plain [cout][npix] layout)
        long long off[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const long long pp = p0 + pg + 32 * i;
            const int ty = static_cast<int>(pp / tile_w);
            off[i] = (pp < npix) ? ty * out_w + (pp - static_cast<long long>(ty) * tile_w) : -1;
        }
        if constexpr (QOUT) {
            if (og * 8 < cout) {
                float bb[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) bb[k] = bias ? __ldg(bias + og * 8 + k) : 0.f;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const long long pp = p0 + pg + 32 * i;
                    if (pp >= npix) continue;
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        unpack2(acc[i][j], v[2 * j], v[2 * j + 1]);
                        v[2 * j] += bb[2 * j];
                        v[2 * j + 1] += bb[2 * j + 1];
                    }
                    heads_store_q8(qo, pp, static_cast<int>(blockIdx.y) * kHeadsOut + og * 8, v);
                }
            }
        } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int o = og * 8 + 2 * j + h;
                if (o < cout) {
                    const float b = bias ? __ldg(bias + o) : 0.f;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        float lo, hi;
                        unpack2(acc[i][j], lo, hi);
                        if (off[i] >= 0) out[static_cast<long long>(o) * out_npix + off[i]] = (h == 0 ? lo : hi) + b;
                    }
                }
            }
        }
        }
        __syncthreads();          // everyone is done with this buffer before the next iteration refills it
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// The same GEMM with a 4 pixel x 8 output register tile (C a multiple of 128): half the weight loads per FMA of the
// 2 x 8 kernel above.  Pixel tiles of 128 are staged in K halves of 128 channels (two 66 KB buffers next to the 72 KB
// weight matrix), the accumulators live across the K chunks of a tile.  Per output the fma order is k ascending as
// above, so both kernels give identical bits.
constexpr int kHeadsPix4 = 128, kHeadsKC = 128;

template <bool QOUT>
__global__ void __launch_bounds__(kHeadsThreads) heads_kernel_p4(const float* __restrict__ x,
                                                                 const float* __restrict__ wt,
                                                                 const float* __restrict__ bias,
                                                                 float* __restrict__ out, long long npix, int C,
                                                                 int cout, int tile_w, long long out_w,
                                                                 long long out_npix, const HeadsQOut qo) {
    extern __shared__ float4 hsm4[];
    float* hsm = reinterpret_cast<float*>(hsm4);
    wt += static_cast<long long>(blockIdx.y) * C * kHeadsOut;           // chunk of 72 outputs, as in heads_kernel
    if (bias) bias += blockIdx.y * kHeadsOut;
    out += static_cast<long long>(blockIdx.y) * kHeadsOut * out_npix;
    cout = min(cout - static_cast<int>(blockIdx.y) * kHeadsOut, kHeadsOut);
    constexpr int pitch = kHeadsKC + kHeadsPitchPad;
    float* ws = hsm;                                   // [C][72]
    float* xs0 = hsm + C * kHeadsOut;                  // 2 x [128][pitch]
    const int tid = threadIdx.x;
    const int nkc = C / kHeadsKC;
    const long long ntiles = (npix + kHeadsPix4 - 1) / kHeadsPix4;
    const long long n_my = (static_cast<long long>(blockIdx.x) < ntiles)
                               ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long n_chunks = n_my * nkc;

    auto load_chunk = [&](long long c, int buf) {
        const long long t = blockIdx.x + (c / nkc) * gridDim.x;
        const int kc = static_cast<int>(c % nkc);
        const long long p0 = t * kHeadsPix4;
        const uint32_t dst = smem_u32(xs0 + buf * kHeadsPix4 * pitch);
        for (int idx = tid; idx < kHeadsPix4 * (kHeadsKC / 4); idx += kHeadsThreads) {
            const int pp = idx / (kHeadsKC / 4), v = idx % (kHeadsKC / 4);
            const bool ok = (p0 + pp < npix);
            cp_async_16(dst + (pp * pitch + 4 * v) * 4, ok ? x + (p0 + pp) * C + kc * kHeadsKC + 4 * v : x, ok);
        }
    };
    for (int idx = tid; idx < C * kHeadsOut / 4; idx += kHeadsThreads)
        cp_async_16(smem_u32(ws) + idx * 16, wt + idx * 4, true);
    if (n_chunks > 0) load_chunk(0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");

    const int pg = tid & 31, og = tid >> 5;    // pixels pg + 32 i; outputs og*8 + 2*j, og*8 + 2*j + 1
    f32x2 acc[4][4];
    int buf = 0;
    for (long long c = 0; c < n_chunks; ++c, buf ^= 1) {
        const int kc = static_cast<int>(c % nkc);
        if (c + 1 < n_chunks) load_chunk(c + 1, buf ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const float* xs = xs0 + buf * kHeadsPix4 * pitch;
        const float* wk = ws + kc * kHeadsKC * kHeadsOut + og * 8;
        if (kc == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = pack2(0.f, 0.f);
        }
#pragma unroll 2
        for (int k = 0; k < kHeadsKC; k += 4) {
            float4 xv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + (pg + 32 * i) * pitch + k);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 w01 = *reinterpret_cast<const float4*>(wk + (k + kk) * kHeadsOut);
                const float4 w23 = *reinterpret_cast<const float4*>(wk + (k + kk) * kHeadsOut + 4);
                const f32x2 wp[4] = {pack2(w01.x, w01.y), pack2(w01.z, w01.w), pack2(w23.x, w23.y),
                                     pack2(w23.z, w23.w)};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float xk = (kk == 0) ? xv[i].x : (kk == 1) ? xv[i].y : (kk == 2) ? xv[i].z : xv[i].w;
                    const f32x2 xx = pack2(xk, xk);
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fma2(xx, wp[j], acc[i][j]);
                }
            }
        }
        if (kc == nkc - 1) {
            const long long p0 = (blockIdx.x + (c / nkc) * gridDim.x) * kHeadsPix4;
            long long off[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long pp = p0 + pg + 32 * i;
                const int ty = static_cast<int>(pp / tile_w);
                off[i] = (pp < npix) ? ty * out_w + (pp - static_cast<long long>(ty) * tile_w) : -1;
            }
            if constexpr (QOUT) {
                if (og * 8 < cout) {
                    float bb[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) bb[k] = bias ? __ldg(bias + og * 8 + k) : 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const long long pp = p0 + pg + 32 * i;
                        if (pp >= npix) continue;
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            unpack2(acc[i][j], v[2 * j], v[2 * j + 1]);
                            v[2 * j] += bb[2 * j];
                            v[2 * j + 1] += bb[2 * j + 1];
                        }
                        heads_store_q8(qo, pp, static_cast<int>(blockIdx.y) * kHeadsOut + og * 8, v);
                    }
                }
            } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int o = og * 8 + 2 * j + h;
                    if (o < cout) {
                        const float b = bias ? __ldg(bias + o) : 0.f;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float lo, hi;
                            unpack2(acc[i][j], lo, hi);
                            if (off[i] >= 0)
                                out[static_cast<long long>(o) * out_npix + off[i]] = (h == 0 ? lo : hi) + b;
                        }
                    }
                }
            }
            }
        }
        __syncthreads();          // everyone is done with this buffer before the next iteration refills it
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// ------------------------------------------------------------------------------------------ layout converters
// Batched transpose between channel-major [n][C][P] and pixel-major [n][P][C] with a per-element conversion.
struct QuantizeOp {   // float -> uint8 activation code (reference quant_layer.py:132-133)
    float delta, zp, qmax;
    __device__ uint8_t operator()(float v) const {
        const float q = fminf(fmaxf(__fadd_rn(rintf(__fdiv_rn(v, delta)), zp), 0.f), qmax);
        return static_cast<uint8_t>(q);
    }
};
struct DequantOp {    // uint8 code -> float (reference quant_layer.py:148)
    float delta, zp;
    __device__ float operator()(uint8_t q) const { return __fmul_rn(__fsub_rn(static_cast<float>(q), zp), delta); }
};
struct CopyOp {
    __device__ float operator()(float v) const { return v; }
};

// src [n][R][S] -> dst [n][S][R'] (dst row pitch dpitch >= R, written at column offset doff)
template <class TI, class TO, class Op>
__global__ void transpose_kernel(const TI* __restrict__ src, TO* __restrict__ dst, int R, long long S, int dpitch,
                                 int doff, Op op) {
    __shared__ TO tile[32][33];
    const long long s0 = static_cast<long long>(blockIdx.x) * 32;
    const int r0 = blockIdx.y * 32;
    const int b = blockIdx.z;
    const TI* sp = src + static_cast<long long>(b) * R * S;
    TO* dp = dst + static_cast<long long>(b) * S * dpitch;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int r = r0 + y;
        const long long s = s0 + threadIdx.x;
        if (r < R && s < S) tile[y][threadIdx.x] = op(sp[static_cast<long long>(r) * S + s]);
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const long long s = s0 + y;
        const int r = r0 + threadIdx.x;
        if (r < R && s < S) dp[s * dpitch + doff + r] = tile[threadIdx.x][y];
    }
}

template <class TI, class TO, class Op>
static int launch_transpose(const TI* src, TO* dst, int n, int R, long long S, int dpitch, int doff, Op op,
                            cudaStream_t stream) {
    dim3 grid(static_cast<unsigned>((S + 31) / 32), static_cast<unsigned>((R + 31) / 32), static_cast<unsigned>(n));
    transpose_kernel<TI, TO, Op><<<grid, dim3(32, 8), 0, stream>>>(src, dst, R, S, dpitch, doff, op);
    g_launch_count.fetch_add(1);
    QV2X_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace qv2x

using namespace qv2x;

struct qv2x_heads {
    int cin, cout;
    float* d_w = nullptr;
    float* d_b = nullptr;
};

static int fuse_impl(int mode, int n_agents, int H, int W, int C, const float* d_feat, const float* d_score,
                     int score_is_logit, const float* d_affine, float* d_out, int y0, int y1, int x0, int x1,
                     void* stream_, float u8_delta = 0.f);

extern "C" {

int qv2x_fuse(int mode, int n_agents, int H, int W, int C, const float* d_feat, const float* d_affine, float* d_out,
              void* stream_) {
    return qv2x_fuse_tile(mode, n_agents, H, W, C, d_feat, d_affine, d_out, 0, H, 0, W, stream_);
}

int qv2x_fuse_tile(int mode, int n_agents, int H, int W, int C, const float* d_feat, const float* d_affine,
                   float* d_out, int y0, int y1, int x0, int x1, void* stream_) {
    QV2X_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (max) or 1 (attention)");
    return fuse_impl(mode, n_agents, H, W, C, d_feat, nullptr, 0, d_affine, d_out, y0, y1, x0, x1, stream_);
}

int qv2x_fuse_weighted(int n_agents, int H, int W, int C, const float* d_feat, const float* d_score,
                       int score_is_logit, const float* d_affine, float* d_out, void* stream_) {
    QV2X_REQUIRE(d_score, "qv2x_fuse_weighted: null score map");
    return fuse_impl(2, n_agents, H, W, C, d_feat, d_score, score_is_logit, d_affine, d_out, 0, H, 0, W, stream_);
}

int qv2x_fuse_weighted_u8(int n_agents, int H, int W, int C, const uint8_t* d_feat_u8, float delta,
                          const float* d_score, int score_is_logit, const float* d_affine, float* d_out,
                          void* stream_) {
    QV2X_REQUIRE(d_score, "qv2x_fuse_weighted_u8: null score map");
    QV2X_REQUIRE(delta > 0.f, "qv2x_fuse_weighted_u8: delta must be positive");
    return fuse_impl(2, n_agents, H, W, C, reinterpret_cast<const float*>(d_feat_u8), d_score, score_is_logit,
                     d_affine, d_out, 0, H, 0, W, stream_, delta);
}

}  // extern "C"

static int fuse_impl(int mode, int n_agents, int H, int W, int C, const float* d_feat, const float* d_score,
                     int score_is_logit, const float* d_affine, float* d_out, int y0, int y1, int x0, int x1,
                     void* stream_, float u8_delta) {
    QV2X_REQUIRE(d_feat && d_affine && d_out, "qv2x_fuse: null argument");
    QV2X_REQUIRE(0 <= y0 && y0 < y1 && y1 <= H && 0 <= x0 && x0 < x1 && x1 <= W, "bad output tile");
    QV2X_REQUIRE(n_agents >= 1 && n_agents <= kMaxAgents, "n_agents must be 1..%d", kMaxAgents);
    QV2X_REQUIRE(C % 4 == 0 && C <= 512, "C must be a multiple of 4 and <= 512");
    QV2X_REQUIRE(H > 0 && W > 0, "empty feature map");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    FuseParams p{};
    p.n = n_agents;
    p.H = H;
    p.W = W;
    p.C = C;
    p.aff = d_affine;
    p.y0 = y0;
    p.x0 = x0;
    p.th = y1 - y0;
    p.tw = x1 - x0;
    p.inv_sqrt_c = 1.0f / sqrtf(static_cast<float>(C));
    p.score = d_score;
    p.score_is_logit = score_is_logit;
    p.u8_delta = u8_delta;
    const long long npix = static_cast<long long>(p.th) * p.tw;
    const int threads = 256;
    const int grid = static_cast<int>(std::min<long long>((npix * 32 + threads - 1) / threads,
                                                          static_cast<long long>(num_sms()) * 8));
    const int vpl = (C / 4 + 31) / 32;
    if (mode == 0) {
        if (vpl == 1) launch_fuse<0, 1>(n_agents, grid, threads, stream, d_feat, d_out, p);
        else if (vpl == 2) launch_fuse<0, 2>(n_agents, grid, threads, stream, d_feat, d_out, p);
        else launch_fuse<0, 4>(n_agents, grid, threads, stream, d_feat, d_out, p);
    } else if (mode == 1) {
        if (vpl == 1) launch_fuse<1, 1>(n_agents, grid, threads, stream, d_feat, d_out, p);
        else if (vpl == 2) launch_fuse<1, 2>(n_agents, grid, threads, stream, d_feat, d_out, p);
        else launch_fuse<1, 4>(n_agents, grid, threads, stream, d_feat, d_out, p);
    } else if (u8_delta > 0.f) {
        if (vpl == 1) launch_fuse<2, 1, true>(n_agents, grid, threads, stream, d_feat, d_out, p);
        else if (vpl == 2) launch_fuse<2, 2, true>(n_agents, grid, threads, stream, d_feat, d_out, p);
        else launch_fuse<2, 4, true>(n_agents, grid, threads, stream, d_feat, d_out, p);
    } else {
        if (vpl == 1) launch_fuse<2, 1>(n_agents, grid, threads, stream, d_feat, d_out, p);
        else if (vpl == 2) launch_fuse<2, 2>(n_agents, grid, threads, stream, d_feat, d_out, p);
        else launch_fuse<2, 4>(n_agents, grid, threads, stream, d_feat, d_out, p);
    }
    g_launch_count.fetch_add(1);
    QV2X_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" {

int qv2x_heads_create(int cin, int cout, const float* w, const float* bias, qv2x_heads** out) {
    QV2X_REQUIRE(w && out, "qv2x_heads_create: null argument");
    QV2X_REQUIRE(cin % 4 == 0 && cin <= 256, "cin must be a multiple of 4 and <= 256");
    QV2X_REQUIRE(cout >= 1 && cout <= 65535 * kHeadsOut, "cout must be 1..%d", 65535 * kHeadsOut);
    auto h = new qv2x_heads();
    h->cin = cin;
    h->cout = cout;
    // chunks of 72 outputs, each k-major and zero-padded: wt[chunk][k][o] = w[chunk * 72 + o][k] (what a CTA stages in
    // shared memory); the bias is padded the same way
    const int chunks = (cout + kHeadsOut - 1) / kHeadsOut;
    std::vector<float> wt(static_cast<size_t>(chunks) * cin * kHeadsOut, 0.f), bp(static_cast<size_t>(chunks) * kHeadsOut, 0.f);
    for (int o = 0; o < cout; ++o) {
        const size_t base = static_cast<size_t>(o / kHeadsOut) * cin * kHeadsOut + (o % kHeadsOut);
        for (int k = 0; k < cin; ++k) wt[base + static_cast<size_t>(k) * kHeadsOut] = w[static_cast<size_t>(o) * cin + k];
        if (bias) bp[o] = bias[o];          // chunk stride == chunk width: only the tail is padding
    }
    int rc = upload(&h->d_w, wt.data(), wt.size());
    if (!rc && bias) rc = upload(&h->d_b, bp.data(), bp.size());
    if (rc) {
        cudaFree(h->d_w);
        delete h;
        return rc;
    }
    *out = h;
    return 0;
}

void qv2x_heads_destroy(qv2x_heads* h) {
    if (!h) return;
    cudaFree(h->d_w);
    cudaFree(h->d_b);
    delete h;
}

int qv2x_heads_forward(const qv2x_heads* h, long long pixels, const float* d_x, float* d_out, void* stream_) {
    return qv2x_heads_forward_tile(h, pixels, d_x, d_out, static_cast<int>(std::min<long long>(pixels, 1 << 30)),
                                   pixels, pixels, stream_);
}

static int heads_launch(const qv2x_heads* h, long long pixels, const float* d_x, float* d_out, int tile_w,
                        long long out_w, long long out_pixels, const HeadsQOut* qo, void* stream_) {
    if (pixels <= 0) return 0;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    static bool attr = false;
    if (!attr) {
        QV2X_CUDA_OK(cudaFuncSetAttribute(heads_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        QV2X_CUDA_OK(cudaFuncSetAttribute(heads_kernel_p4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        QV2X_CUDA_OK(cudaFuncSetAttribute(heads_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        QV2X_CUDA_OK(cudaFuncSetAttribute(heads_kernel_p4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr = true;
    }
    static const bool p4_enabled = [] {
        const char* e = getenv("QV2X_HEADS_P4");
        return !(e && e[0] == '0');
    }();
    // one CTA per SM overall: the output chunks of 72 columns are blockIdx.y, the pixel tiles of a chunk are shared
    // by ceil(#SMs / chunks) persistent CTAs
    const int chunks = (h->cout + kHeadsOut - 1) / kHeadsOut;
    const int per_chunk = std::max(1, (num_sms() + chunks - 1) / chunks);
    const HeadsQOut q0{};
    if (p4_enabled && h->cin % kHeadsKC == 0 && pixels * chunks >= 4LL * kHeadsPix4 * num_sms() / 8) {
        const int smem = (2 * kHeadsPix4 * (kHeadsKC + kHeadsPitchPad) + h->cin * kHeadsOut) *
                         static_cast<int>(sizeof(float));
        const long long ntiles = (pixels + kHeadsPix4 - 1) / kHeadsPix4;
        const dim3 grid(static_cast<unsigned>(std::min<long long>(ntiles, per_chunk)), chunks);
        if (qo)
            heads_kernel_p4<true><<<grid, kHeadsThreads, smem, stream>>>(d_x, h->d_w, h->d_b, d_out, pixels, h->cin,
                                                                         h->cout, tile_w, out_w, out_pixels, *qo);
        else
            heads_kernel_p4<false><<<grid, kHeadsThreads, smem, stream>>>(d_x, h->d_w, h->d_b, d_out, pixels, h->cin,
                                                                          h->cout, tile_w, out_w, out_pixels, q0);
    } else {
        const int smem = (2 * kHeadsPix * (h->cin + kHeadsPitchPad) + h->cin * kHeadsOut) *
                         static_cast<int>(sizeof(float));
        const long long ntiles = (pixels + kHeadsPix - 1) / kHeadsPix;
        const dim3 grid(static_cast<unsigned>(std::min<long long>(ntiles, per_chunk)), chunks);
        if (qo)
            heads_kernel<true><<<grid, kHeadsThreads, smem, stream>>>(d_x, h->d_w, h->d_b, d_out, pixels, h->cin,
                                                                      h->cout, tile_w, out_w, out_pixels, *qo);
        else
            heads_kernel<false><<<grid, kHeadsThreads, smem, stream>>>(d_x, h->d_w, h->d_b, d_out, pixels, h->cin,
                                                                       h->cout, tile_w, out_w, out_pixels, q0);
    }
    g_launch_count.fetch_add(1);
    QV2X_CUDA_OK(cudaGetLastError());
    return 0;
}

int qv2x_heads_forward_tile(const qv2x_heads* h, long long pixels, const float* d_x, float* d_out, int tile_w,
                            long long out_w, long long out_pixels, void* stream_) {
    QV2X_REQUIRE(h && d_x && d_out, "qv2x_heads_forward: null argument");
    QV2X_REQUIRE(tile_w >= 1 && out_w >= tile_w && out_pixels >= 1, "bad output tile geometry");
    return heads_launch(h, pixels, d_x, d_out, tile_w, out_w, out_pixels, nullptr, stream_);
}

int qv2x_heads_forward_deconv_u8(const qv2x_heads* h, int in_h, int in_w, const float* d_x, int stride, float out_delta,
                                 uint8_t* d_out_u8, int out_cstride, int out_cbase, void* stream_) {
    QV2X_REQUIRE(h && d_x && d_out_u8, "qv2x_heads_forward_deconv_u8: null argument");
    QV2X_REQUIRE(in_h > 0 && in_w > 0 && stride >= 1 && out_delta > 0.f, "bad map size / stride / scale");
    QV2X_REQUIRE(h->cout % (stride * stride) == 0, "the GEMM has %d columns, not a multiple of stride^2", h->cout);
    const int cs = h->cout / (stride * stride);
    QV2X_REQUIRE(cs % 8 == 0 && out_cbase % 8 == 0 && out_cstride % 8 == 0 && out_cbase + cs <= out_cstride,
                 "output channels, channel base and channel stride must be multiples of 8");
    HeadsQOut qo{};
    qo.out = d_out_u8;
    qo.delta = out_delta;
    qo.inv_delta = 1.0f / out_delta;
    qo.s = stride;
    qo.cs = cs;
    qo.w = in_w;
    qo.cstride = out_cstride;
    qo.cbase = out_cbase;
    const long long pixels = static_cast<long long>(in_h) * in_w;
    return heads_launch(h, pixels, d_x, reinterpret_cast<float*>(d_out_u8), in_w, pixels, pixels, &qo, stream_);
}

int qv2x_quantize_nchw_to_nhwc_u8(const float* d_x, int n, int c, long long pixels, float delta, float zero_point,
                                  int bits, uint8_t* d_y, int out_cstride, int out_cbase, void* stream) {
    QV2X_REQUIRE(d_x && d_y && delta > 0.f && bits >= 2 && bits <= 8, "qv2x_quantize_nchw_to_nhwc_u8: bad argument");
    QuantizeOp op{delta, zero_point, static_cast<float>((1 << bits) - 1)};
    return launch_transpose<float, uint8_t>(d_x, d_y, n, c, pixels, out_cstride, out_cbase, op,
                                            static_cast<cudaStream_t>(stream));
}

int qv2x_dequant_nhwc_u8_to_nchw_f32(const uint8_t* d_x, int n, int c, long long pixels, float delta,
                                     float zero_point, float* d_y, void* stream) {
    QV2X_REQUIRE(d_x && d_y, "qv2x_dequant_nhwc_u8_to_nchw_f32: null argument");
    DequantOp op{delta, zero_point};
    // src is [n][pixels][c] -> dst [n][c][pixels]
    return launch_transpose<uint8_t, float>(d_x, d_y, n, static_cast<int>(pixels), c, static_cast<int>(pixels), 0, op,
                                            static_cast<cudaStream_t>(stream));
}

int qv2x_nchw_to_nhwc_f32(const float* d_x, int n, int c, long long pixels, float* d_y, void* stream) {
    QV2X_REQUIRE(d_x && d_y, "qv2x_nchw_to_nhwc_f32: null argument");
    return launch_transpose<float, float>(d_x, d_y, n, c, pixels, c, 0, CopyOp{}, static_cast<cudaStream_t>(stream));
}

int qv2x_nhwc_to_nchw_f32(const float* d_x, int n, int c, long long pixels, float* d_y, void* stream) {
    QV2X_REQUIRE(d_x && d_y, "qv2x_nhwc_to_nchw_f32: null argument");
    return launch_transpose<float, float>(d_x, d_y, n, static_cast<int>(pixels), c, static_cast<int>(pixels), 0,
                                          CopyOp{}, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
