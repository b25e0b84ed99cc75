// qv2x_decode_linear: codebook decode followed by a 1x1 convolution and an activation quantizer, folded over the
// codeword tables -- the entry of the pyramid model's ego stage.
//
// Reference chain: UMGMQuantizer.decode (opencood/models/sub_modules/codebook.py:192-201, 263-269) -> conv1 of the
// first ResNeXt bottleneck of PyramidFusion (opencood/models/sub_modules/resblock.py:67-122, wrapped by QuantBottleneck,
// opencood/quant/quant_block.py:100-134: QuantModule conv1 -> BN folded -> ReLU -> act quantizer).
//   decode(codes) = const + sum_i T_i[code_i]   and conv1 is linear, so
//   q1[row][c]    = clamp(rint((b[c] + W const + sum_i (W T_i)[code_i][c]) / delta), 0, 255)
// with the tables FT_i = T_i W^T and the constant folded at create time in float64 (stored fp32).  Three 512-byte row
// reads from shared memory per pixel replace a 64 -> 128 FP32 GEMM on decoded features plus a transposing quantizer
// pass; the per-pixel sums of the codes (the next layer's zero-point term) come out of the same kernel.
#include <algorithm>
#include <cmath>
#include <vector>

#include "host_common.h"
#include "ptx.cuh"

namespace qv2x {

constexpr int kDLMaxTables = 8;

struct DecodeLinearParams {
    int nt, R, cout;
    int rowbase[kDLMaxTables], kk[kDLMaxTables];
    long long rows, plane_stride;
    const uint8_t* codes;
    const float* ft;          // [R][cout]
    const float* bias;        // [cout]  (b + W const)
    float delta, inv_delta;
    uint8_t* out;             // [rows][cout]
    int32_t* rowsum;          // [rows] or nullptr
};

// q = clamp(rint(v / delta), 0, 255), bit-identical to the IEEE division of the converter kernels
// (fusion.cu QuantizeOp): the product with fl(1/delta) is within a few ulps of the quotient, so only values that land
// within 1e-3 of a rounding boundary are redone with the exact division.
__device__ __forceinline__ float quant_u8(float v, float delta, float inv_delta) {
    const float t = v * inv_delta;
    float r = rintf(t);
    if (fabsf(t - r) > 0.499f) r = rintf(__fdiv_rn(v, delta));
    return fminf(fmaxf(r, 0.f), 255.f);
}

template <int NT>
__global__ void __launch_bounds__(1024, 1) decode_linear_kernel(const DecodeLinearParams p) {
    extern __shared__ float4 dlsm4[];
    float* s_ft = reinterpret_cast<float*>(dlsm4);          // [R][cout] then bias [cout]
    const int cout = p.cout, vec = cout >> 2;
    for (int idx = threadIdx.x; idx < p.R * vec; idx += blockDim.x)
        reinterpret_cast<float4*>(s_ft)[idx] = __ldg(reinterpret_cast<const float4*>(p.ft) + idx);
    float* s_bias = s_ft + p.R * cout;
    for (int idx = threadIdx.x; idx < vec; idx += blockDim.x)
        reinterpret_cast<float4*>(s_bias)[idx] = __ldg(reinterpret_cast<const float4*>(p.bias) + idx);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    constexpr int RW = 4;                                   // rows per warp iteration (code loads in flight)
    int code[RW][NT];
    auto fetch = [&](long long r0) {
#pragma unroll
        for (int rr = 0; rr < RW; ++rr)
#pragma unroll
            for (int i = 0; i < NT; ++i)
                code[rr][i] = (r0 + rr < p.rows) ? __ldg(p.codes + i * p.plane_stride + r0 + rr) : 0;
    };
    long long r0 = wid * RW;
    if (r0 < p.rows) fetch(r0);
    while (r0 < p.rows) {
        int off[RW][NT];
#pragma unroll
        for (int rr = 0; rr < RW; ++rr)
#pragma unroll
            for (int i = 0; i < NT; ++i) off[rr][i] = (p.rowbase[i] + min(code[rr][i], p.kk[i] - 1)) * cout;
        const long long rn = r0 + nwarps * RW;
        if (rn < p.rows) fetch(rn);
        int rs[RW];
#pragma unroll
        for (int rr = 0; rr < RW; ++rr) rs[rr] = 0;
        for (int v = lane; v < vec; v += 32) {
            const float4 b = reinterpret_cast<const float4*>(s_bias)[v];
#pragma unroll
            for (int rr = 0; rr < RW; ++rr) {
                float4 a = b;
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    const float4 t = reinterpret_cast<const float4*>(s_ft + off[rr][i])[v];
                    a.x = __fadd_rn(a.x, t.x);
                    a.y = __fadd_rn(a.y, t.y);
                    a.z = __fadd_rn(a.z, t.z);
                    a.w = __fadd_rn(a.w, t.w);
                }
                const unsigned q0 = static_cast<unsigned>(quant_u8(a.x, p.delta, p.inv_delta));
                const unsigned q1 = static_cast<unsigned>(quant_u8(a.y, p.delta, p.inv_delta));
                const unsigned q2 = static_cast<unsigned>(quant_u8(a.z, p.delta, p.inv_delta));
                const unsigned q3 = static_cast<unsigned>(quant_u8(a.w, p.delta, p.inv_delta));
                rs[rr] += static_cast<int>(q0 + q1 + q2 + q3);
                if (r0 + rr < p.rows)
                    reinterpret_cast<uint32_t*>(p.out + (r0 + rr) * cout)[v] = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
            }
        }
        if (p.rowsum != nullptr) {
#pragma unroll
            for (int rr = 0; rr < RW; ++rr) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) rs[rr] += __shfl_xor_sync(0xffffffffu, rs[rr], o);
            }
            if (lane < RW && r0 + lane < p.rows) {
                int vsum = rs[0];
#pragma unroll
                for (int rr = 1; rr < RW; ++rr) vsum = (lane == rr) ? rs[rr] : vsum;
                p.rowsum[r0 + lane] = vsum;
            }
        }
        r0 = rn;
    }
}

}  // namespace qv2x

using namespace qv2x;

struct qv2x_decode_linear {
    int nt, R, cout, C;
    int rowbase[kDLMaxTables], kk[kDLMaxTables];
    float delta;
    float *d_ft = nullptr, *d_bias = nullptr;
    int smem;
};

static bool decode_linear_fits(int R, int cout, int* smem) {
    const long long need = 4ll * (static_cast<long long>(R) + 1) * cout;
    if (smem) *smem = static_cast<int>(need);
    return need <= 220 * 1024;
}

extern "C" {

int qv2x_decode_linear_supported(const qv2x_codebook* cb, int cout) {
    qv2x_codebook_desc d{};
    d.struct_size = sizeof(d);
    if (!cb || qv2x_codebook_desc_get(cb, &d) != 0) return 0;
    const int nt = d.levels * d.m;
    int R = 0;
    for (int l = 0; l < d.levels; ++l) R += d.m * d.k[l];
    return (nt >= 1 && nt <= 4 && cout >= 4 && cout % 4 == 0 && decode_linear_fits(R, cout, nullptr)) ? 1 : 0;
}

int qv2x_decode_linear_create(const qv2x_codebook* cb, int cout, const float* w, const float* bias, float out_delta,
                              qv2x_decode_linear** out) {
    QV2X_REQUIRE(cb && w && out, "qv2x_decode_linear_create: null argument");
    QV2X_REQUIRE(out_delta > 0.f, "out_delta must be positive");
    QV2X_REQUIRE(qv2x_decode_linear_supported(cb, cout),
                 "qv2x_decode_linear: unsupported configuration (levels*m <= 4, cout a multiple of 4 and a folded "
                 "table of sum(k)*m x cout floats that fits in shared memory)");
    qv2x_codebook_desc d{};
    d.struct_size = sizeof(d);
    int rc = qv2x_codebook_desc_get(cb, &d);
    if (rc) return rc;
    const int C = d.channel;
    auto h = new qv2x_decode_linear();
    h->C = C;
    h->nt = d.levels * d.m;
    h->cout = cout;
    h->delta = out_delta;
    int R = 0;
    for (int l = 0; l < d.levels; ++l)
        for (int s = 0; s < d.m; ++s) {
            h->rowbase[l * d.m + s] = R;
            h->kk[l * d.m + s] = d.k[l];
            R += d.k[l];
        }
    h->R = R;
    decode_linear_fits(R, cout, &h->smem);
    std::vector<float> tab(static_cast<size_t>(qv2x_codebook_folded_size(cb, 5))), cst(static_cast<size_t>(C));
    if (static_cast<long long>(tab.size()) != static_cast<long long>(R) * C || qv2x_codebook_folded_size(cb, 4) != C) {
        delete h;
        return set_error(QV2X_ERR_INVALID, "qv2x_decode_linear_create: unexpected decode table size");
    }
    qv2x_codebook_folded_copy(cb, 5, tab.data());
    qv2x_codebook_folded_copy(cb, 4, cst.data());
    std::vector<float> ft(static_cast<size_t>(R) * cout), bp(static_cast<size_t>(cout));
    for (int r = 0; r < R; ++r)
        for (int o = 0; o < cout; ++o) {
            double s = 0.0;
            for (int c = 0; c < C; ++c)
                s += static_cast<double>(w[static_cast<size_t>(o) * C + c]) * static_cast<double>(tab[static_cast<size_t>(r) * C + c]);
            ft[static_cast<size_t>(r) * cout + o] = static_cast<float>(s);
        }
    for (int o = 0; o < cout; ++o) {
        double s = bias ? static_cast<double>(bias[o]) : 0.0;
        for (int c = 0; c < C; ++c) s += static_cast<double>(w[static_cast<size_t>(o) * C + c]) * static_cast<double>(cst[c]);
        bp[o] = static_cast<float>(s);
    }
    rc = upload(&h->d_ft, ft.data(), ft.size());
    if (!rc) rc = upload(&h->d_bias, bp.data(), bp.size());
    if (rc) {
        qv2x_decode_linear_destroy(h);
        return rc;
    }
    *out = h;
    return 0;
}

void qv2x_decode_linear_destroy(qv2x_decode_linear* h) {
    if (!h) return;
    cudaFree(h->d_ft);
    cudaFree(h->d_bias);
    delete h;
}

int qv2x_decode_linear_forward(const qv2x_decode_linear* h, long long rows, const uint8_t* d_codes,
                               long long plane_stride, uint8_t* d_out, int32_t* d_rowsum, void* stream_) {
    QV2X_REQUIRE(h && d_codes && d_out, "qv2x_decode_linear_forward: null argument");
    QV2X_REQUIRE(plane_stride >= rows, "code planes are shorter than the row count");
    if (rows <= 0) return 0;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DecodeLinearParams p{};
    p.nt = h->nt;
    p.R = h->R;
    p.cout = h->cout;
    for (int i = 0; i < h->nt; ++i) {
        p.rowbase[i] = h->rowbase[i];
        p.kk[i] = h->kk[i];
    }
    p.rows = rows;
    p.plane_stride = plane_stride;
    p.codes = d_codes;
    p.ft = h->d_ft;
    p.bias = h->d_bias;
    p.delta = h->delta;
    p.inv_delta = 1.0f / h->delta;
    p.out = d_out;
    p.rowsum = d_rowsum;
    const int threads = 1024;
    const long long want = (rows + (threads / 32) * 4 - 1) / ((threads / 32) * 4);
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(num_sms(), want)));
#define QV2X_DL_CASE(N)                                                                                            \
    case N: {                                                                                                      \
        static bool attr = false;                                                                                  \
        if (!attr) {                                                                                               \
            QV2X_CUDA_OK(cudaFuncSetAttribute(decode_linear_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                              224 * 1024));                                                        \
            attr = true;                                                                                           \
        }                                                                                                          \
        decode_linear_kernel<N><<<grid, threads, h->smem, stream>>>(p);                                            \
        break;                                                                                                     \
    }
    switch (h->nt) {
        QV2X_DL_CASE(1) QV2X_DL_CASE(2) QV2X_DL_CASE(3) QV2X_DL_CASE(4)
        default: return set_error(QV2X_ERR_INVALID, "qv2x_decode_linear_forward: unsupported table count %d", h->nt);
    }
#undef QV2X_DL_CASE
    g_launch_count.fetch_add(1);
    QV2X_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
