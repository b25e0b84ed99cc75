// qv2x_pillar: the quantized PointPillars front end as ONE kernel:
//   pillars [M, 32, 4] -> point decoration (offsets to the pillar mean and to the cell centre, padded points
//   zeroed) -> Linear(10 -> 64, fake-quant weights, BN folded) -> pre-ReLU activation quantizer -> ReLU -> block
//   activation quantizer -> max over the pillar's 32 points -> scatter into the uint8 NHWC BEV map.
// Reference: QuantPillarVFE.forward / QuantPFNLayer.forward (opencood/quant/quant_block.py:611-630, 666-716),
// PointPillarScatter.forward (opencood/models/sub_modules/point_pillar_scatter.py:19-75).  The reference
// materialises [M, 32, 64] FP32 twice and a dense FP32 BEV tensor; here a warp owns one pillar, the decorated
// points live in shared memory, lane l evaluates outputs l and l + 32 for all 32 points, and only the 64 result
// bytes per pillar are written.  Both quantizers and the ReLU are monotone non-decreasing, so the maximum over the
// points is taken on the FP32 linear output and quantized once (identical to quantizing all 32 and taking the max).
//
// Normative FP32 order (oracle/pillar_oracle.py):
//   sum_xyz  : xor-butterfly over the 32 lanes (16, 8, 4, 2, 1); mean = sum / float(num_points)
//   f[0..3]  = x, y, z, i;  f[4..6] = xyz - mean;  f[7..9] = xyz - fl(fl(cell * voxel) + offset);  padded points -> 0
//   y[o]     = bias[o], then y = fma(f[c], w[o][c], y) for c = 0..9
//   q1 = clamp(rint(y / d1) + z1, 0, 2^b1 - 1);  y1 = (q1 - z1) * d1;  r = max(y1, 0)
//   q  = clamp(rint(r / d2) + z2, 0, 2^b2 - 1)
#include <algorithm>
#include <cstring>
#include <vector>

#include "host_common.h"
#include "ptx.cuh"

namespace qv2x {

constexpr int kPillarPoints = 32, kPillarFeat = 10, kPillarFeatPad = 12, kPillarOut = 64;

struct PillarParams {
    float vx, vy, vz, ox, oy, oz;
    int nx, ny;
    int has_q1;
    float d1, z1, qmax1;
    float d2, z2, qmax2;
};

__global__ void __launch_bounds__(256) pillar_bev_kernel(const float4* __restrict__ pts, const int4* __restrict__ coords,
                                                         const int* __restrict__ num, int n_pillars, int batch,
                                                         const float* __restrict__ w /*[64][12]*/,
                                                         const float* __restrict__ bias, uint8_t* __restrict__ bev,
                                                         const PillarParams p, int32_t* __restrict__ rowsum) {
    __shared__ __align__(16) float s_f[8][kPillarPoints][kPillarFeatPad];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // this lane's two output channels: weights and bias stay in registers for the whole kernel
    // (as packed fp32 pairs: one FFMA2 advances both outputs, the point feature is the broadcast operand)
    f32x2 wp[kPillarFeat];
#pragma unroll
    for (int c = 0; c < kPillarFeat; ++c)
        wp[c] = pack2(__ldg(w + lane * kPillarFeatPad + c), __ldg(w + (lane + 32) * kPillarFeatPad + c));
    const f32x2 bp = pack2(__ldg(bias + lane), __ldg(bias + lane + 32));
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < n_pillars; m += warps) {
        const float4 pt = __ldg(pts + static_cast<long long>(m) * kPillarPoints + lane);
        const int4 cd = __ldg(coords + m);            // (batch, z, y, x)
        const int n = __ldg(num + m);
        float sx = pt.x, sy = pt.y, sz = pt.z;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sx = __fadd_rn(sx, __shfl_xor_sync(0xffffffffu, sx, off));
            sy = __fadd_rn(sy, __shfl_xor_sync(0xffffffffu, sy, off));
            sz = __fadd_rn(sz, __shfl_xor_sync(0xffffffffu, sz, off));
        }
        const float fn = static_cast<float>(n);
        const float mx = __fdiv_rn(sx, fn), my = __fdiv_rn(sy, fn), mz = __fdiv_rn(sz, fn);
        const float cx = __fadd_rn(__fmul_rn(static_cast<float>(cd.w), p.vx), p.ox);
        const float cy = __fadd_rn(__fmul_rn(static_cast<float>(cd.z), p.vy), p.oy);
        const float cz = __fadd_rn(__fmul_rn(static_cast<float>(cd.y), p.vz), p.oz);
        const bool live = lane < n;
        float f[kPillarFeatPad];
        f[0] = pt.x, f[1] = pt.y, f[2] = pt.z, f[3] = pt.w;
        f[4] = __fsub_rn(pt.x, mx), f[5] = __fsub_rn(pt.y, my), f[6] = __fsub_rn(pt.z, mz);
        f[7] = __fsub_rn(pt.x, cx), f[8] = __fsub_rn(pt.y, cy), f[9] = __fsub_rn(pt.z, cz);
        f[10] = 0.f, f[11] = 0.f;
#pragma unroll
        for (int c = 0; c < kPillarFeat; ++c) f[c] = live ? f[c] : 0.f;
        __syncwarp();                                  // the previous pillar's readers are done
        float4* dst = reinterpret_cast<float4*>(&s_f[wib][lane][0]);
        dst[0] = make_float4(f[0], f[1], f[2], f[3]);
        dst[1] = make_float4(f[4], f[5], f[6], f[7]);
        dst[2] = make_float4(f[8], f[9], 0.f, 0.f);
        __syncwarp();
        // Padded rows have all-zero features, so their linear output is the bias exactly (fma(0, w, u) = u) and they
        // are all alike: the maximum over the 32 rows = max(bias if any row is padded, rows of the nv valid points).
        // The loop runs over the valid points only (half the work at a uniform 1..32 points per pillar).
        const int nv = min(max(n, 0), kPillarPoints);
        float y0 = -INFINITY, y1 = -INFINITY;
        if (nv < kPillarPoints) unpack2(bp, y0, y1);
#pragma unroll 4
        for (int q = 0; q < nv; ++q) {
            const float4* src = reinterpret_cast<const float4*>(&s_f[wib][q][0]);     // broadcast reads
            const float4 a = src[0], b = src[1], c = src[2];
            const float g[kPillarFeat] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y};
            f32x2 u = bp;
#pragma unroll
            for (int k = 0; k < kPillarFeat; ++k) u = fma2(pack2(g[k], g[k]), wp[k], u);
            float u0, u1;
            unpack2(u, u0, u1);
            y0 = fmaxf(y0, u0);
            y1 = fmaxf(y1, u1);
        }
        auto quant = [&](float y) -> uint8_t {
            if (p.has_q1) {
                float t = __fadd_rn(rintf(__fdiv_rn(y, p.d1)), p.z1);
                t = fminf(fmaxf(t, 0.f), p.qmax1);
                y = __fmul_rn(__fsub_rn(t, p.z1), p.d1);
            }
            y = fmaxf(y, 0.f);
            float q = __fadd_rn(rintf(__fdiv_rn(y, p.d2)), p.z2);
            q = fminf(fmaxf(q, 0.f), p.qmax2);
            return static_cast<uint8_t>(q);
        };
        const uint8_t q0 = quant(y0), q1 = quant(y1);
        // voxel_coords rows outside the grid / batch (never produced by the reference voxelizer) are dropped
        if (cd.x >= 0 && cd.x < batch && cd.z >= 0 && cd.z < p.ny && cd.w >= 0 && cd.w < p.nx) {
            uint8_t* cell = bev + ((static_cast<long long>(cd.x) * p.ny + cd.z) * p.nx + cd.w) * kPillarOut;
            cell[lane] = q0;
            cell[lane + 32] = q1;
            if (rowsum != nullptr) {
                // the cell's channel sum (the uint8 x uint8 zero-point term of the first conv needs it): saves the
                // separate pass over the 95 % empty map
                int sum = static_cast<int>(q0) + static_cast<int>(q1);
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                if (lane == 0) rowsum[(static_cast<long long>(cd.x) * p.ny + cd.z) * p.nx + cd.w] = sum;
            }
        }
    }
}

// Zero the cells (64 code bytes + the row sum) of a set of pillars: 16 threads per pillar, 4 bytes each.  With the
// scatter-only entry point this replaces the memset of the whole (95 % empty) map: the map is all-zero between frames.
__global__ void pillar_clear_kernel(const int4* __restrict__ coords, int n_pillars, int batch, int ny, int nx,
                                    uint8_t* __restrict__ bev, int32_t* __restrict__ rowsum) {
    const long long total = static_cast<long long>(n_pillars) * 16;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int m = static_cast<int>(i >> 4), part = static_cast<int>(i & 15);
        const int4 cd = __ldg(coords + m);
        if (cd.x >= 0 && cd.x < batch && cd.z >= 0 && cd.z < ny && cd.w >= 0 && cd.w < nx) {
            const long long cell = (static_cast<long long>(cd.x) * ny + cd.z) * nx + cd.w;
            reinterpret_cast<uint32_t*>(bev + cell * kPillarOut)[part] = 0u;
            if (rowsum != nullptr && part == 0) rowsum[cell] = 0;
        }
    }
}

}  // namespace qv2x

using namespace qv2x;

struct qv2x_pillar {
    qv2x_pillar_desc d;
    float* d_w = nullptr;      // [64][12]
    float* d_b = nullptr;      // [64]
};

extern "C" {

int qv2x_pillar_create(const qv2x_pillar_desc* desc, const float* w_hat, const float* bias, qv2x_pillar** out) {
    QV2X_REQUIRE(desc && w_hat && out, "qv2x_pillar_create: null argument");
    QV2X_CHECK_SIZE(desc, qv2x_pillar_desc);
    QV2X_REQUIRE(desc->n_feat == kPillarFeat && desc->cout == kPillarOut && desc->max_points == kPillarPoints,
                 "the pillar kernel is built for 10 decorated features -> 64 channels over 32 points (got %d -> %d, %d)",
                 desc->n_feat, desc->cout, desc->max_points);
    QV2X_REQUIRE(desc->nx > 0 && desc->ny > 0 && desc->out_delta > 0.f, "bad grid / output scale");
    QV2X_REQUIRE(!desc->has_pre_quant || desc->pre_delta > 0.f, "bad pre-ReLU quantizer scale");
    QV2X_REQUIRE(desc->out_bits >= 2 && desc->out_bits <= 8 && desc->pre_bits >= 2 && desc->pre_bits <= 16,
                 "bit widths out of range");
    auto h = new qv2x_pillar();
    h->d = *desc;
    std::vector<float> wp(static_cast<size_t>(kPillarOut) * kPillarFeatPad, 0.f), bp(kPillarOut, 0.f);
    for (int o = 0; o < kPillarOut; ++o) {
        for (int c = 0; c < kPillarFeat; ++c) wp[o * kPillarFeatPad + c] = w_hat[o * kPillarFeat + c];
        if (bias) bp[o] = bias[o];
    }
    int rc = upload(&h->d_w, wp.data(), wp.size());
    if (!rc) rc = upload(&h->d_b, bp.data(), bp.size());
    if (rc) {
        cudaFree(h->d_w);
        delete h;
        return rc;
    }
    *out = h;
    return 0;
}

void qv2x_pillar_destroy(qv2x_pillar* h) {
    if (!h) return;
    cudaFree(h->d_w);
    cudaFree(h->d_b);
    delete h;
}

int qv2x_pillar_forward(const qv2x_pillar* h, int n_pillars, const float* d_points, const int* d_coords,
                        const int* d_num_points, int batch, uint8_t* d_bev, void* stream_) {
    return qv2x_pillar_forward_rs(h, n_pillars, d_points, d_coords, d_num_points, batch, d_bev, nullptr, stream_);
}

int qv2x_pillar_forward_rs(const qv2x_pillar* h, int n_pillars, const float* d_points, const int* d_coords,
                           const int* d_num_points, int batch, uint8_t* d_bev, int32_t* d_rowsum, void* stream_) {
    QV2X_REQUIRE(h && d_bev, "qv2x_pillar_forward: null argument");
    QV2X_REQUIRE(batch >= 1 && n_pillars >= 0, "bad batch / pillar count");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t bytes = static_cast<size_t>(batch) * h->d.ny * h->d.nx * kPillarOut;
    QV2X_CUDA_OK(cudaMemsetAsync(d_bev, 0, bytes, stream));       // empty cells are code 0 (zero-point 0)
    if (d_rowsum) QV2X_CUDA_OK(cudaMemsetAsync(d_rowsum, 0, bytes / kPillarOut * sizeof(int32_t), stream));
    return qv2x_pillar_scatter(h, n_pillars, d_points, d_coords, d_num_points, batch, d_bev, d_rowsum, stream_);
}

int qv2x_pillar_clear(const qv2x_pillar* h, int n_pillars, const int* d_coords, int batch, uint8_t* d_bev,
                      int32_t* d_rowsum, void* stream_) {
    QV2X_REQUIRE(h && d_bev, "qv2x_pillar_clear: null argument");
    QV2X_REQUIRE(batch >= 1 && n_pillars >= 0, "bad batch / pillar count");
    if (n_pillars == 0) return 0;
    QV2X_REQUIRE(d_coords, "qv2x_pillar_clear: null argument");
    const int threads = 256;
    const int grid = std::min((n_pillars * 16 + threads - 1) / threads, num_sms() * 8);
    pillar_clear_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<const int4*>(d_coords), n_pillars, batch, h->d.ny, h->d.nx, d_bev, d_rowsum);
    g_launch_count.fetch_add(1);
    QV2X_CUDA_OK(cudaGetLastError());
    return 0;
}

int qv2x_pillar_scatter(const qv2x_pillar* h, int n_pillars, const float* d_points, const int* d_coords,
                        const int* d_num_points, int batch, uint8_t* d_bev, int32_t* d_rowsum, void* stream_) {
    QV2X_REQUIRE(h && d_bev, "qv2x_pillar_scatter: null argument");
    QV2X_REQUIRE(batch >= 1 && n_pillars >= 0, "bad batch / pillar count");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n_pillars == 0) return 0;
    QV2X_REQUIRE(d_points && d_coords && d_num_points, "qv2x_pillar_forward: null argument");
    PillarParams p{};
    p.vx = h->d.voxel_size[0], p.vy = h->d.voxel_size[1], p.vz = h->d.voxel_size[2];
    p.ox = h->d.offset[0], p.oy = h->d.offset[1], p.oz = h->d.offset[2];
    p.nx = h->d.nx, p.ny = h->d.ny;
    p.has_q1 = h->d.has_pre_quant;
    p.d1 = h->d.pre_delta, p.z1 = h->d.pre_zero_point, p.qmax1 = static_cast<float>((1 << h->d.pre_bits) - 1);
    p.d2 = h->d.out_delta, p.z2 = h->d.out_zero_point, p.qmax2 = static_cast<float>((1 << h->d.out_bits) - 1);
    const int threads = 256;
    const int grid = std::min((n_pillars * 32 + threads - 1) / threads, num_sms() * 8);
    pillar_bev_kernel<<<grid, threads, 0, stream>>>(reinterpret_cast<const float4*>(d_points),
                                                    reinterpret_cast<const int4*>(d_coords), d_num_points, n_pillars,
                                                    batch, h->d_w, h->d_b, d_bev, p, d_rowsum);
    g_launch_count.fetch_add(1);
    QV2X_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
