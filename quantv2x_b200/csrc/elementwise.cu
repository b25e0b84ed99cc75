// Small HBM-bound helper kernels around the igemm path.
#include <algorithm>

#include "host_common.h"

namespace qv2x {

// Per-pixel channel sums of a uint8 NHWC tensor (the S term of the uint8 x uint8 zero-point algebra).
// A group of L lanes owns one pixel; every lane reads 16-byte chunks, sums bytes with dp4a, then the
// group reduces with shuffles.  Coalesced: consecutive lanes read consecutive 16-byte chunks.
template <int L>
__global__ void rowsum_u8_kernel(const uint8_t* __restrict__ x, long long n_pixels, int cstride, int cbase, int c,
                                 int32_t* __restrict__ out) {
    const int lane_in_group = threadIdx.x % L;
    const long long group = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / L;
    const long long n_groups = static_cast<long long>(gridDim.x) * blockDim.x / L;
    const int chunks = c / 16;
    for (long long p = group; p < n_pixels; p += n_groups) {
        const uint4* src = reinterpret_cast<const uint4*>(x + p * cstride + cbase);
        unsigned int s = 0;
        for (int i = lane_in_group; i < chunks; i += L) {
            const uint4 v = __ldg(src + i);
            s = __dp4a(v.x, 0x01010101u, s);
            s = __dp4a(v.y, 0x01010101u, s);
            s = __dp4a(v.z, 0x01010101u, s);
            s = __dp4a(v.w, 0x01010101u, s);
        }
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, L);
        if (lane_in_group == 0) out[p] = static_cast<int32_t>(s);
    }
}

// Multi-GPU exchange of the code planes without a collective: every rank stores its own planes straight into the
// code buffer of every peer (peer-mapped pointers over NVLink / NVSwitch), 16 bytes per thread.
//   local : [planes][rows_local] bytes;  peer p destination: base[p] + plane * dst_plane_stride + dst_row0 + row
struct PushPeers {
    uint8_t* base[8];
    int n;
};
// SCATTER = false: every peer receives all `rows` of each plane (all-gather);
// SCATTER = true : peer p receives rows [p * rows, (p + 1) * rows) of each plane (all-to-all; src_plane_stride =
//                  n_peers * rows).
template <bool SCATTER>
__global__ void push_planes_kernel(const uint8_t* __restrict__ local, int planes, long long rows,
                                   long long src_plane_stride, long long dst_plane_stride, long long dst_row0,
                                   const PushPeers peers) {
    const long long vec_per_plane = rows / 16;
    const long long total = static_cast<long long>(peers.n) * planes * vec_per_plane;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long v = i % vec_per_plane;
        const long long pp = i / vec_per_plane;
        const int plane = static_cast<int>(pp % planes), peer = static_cast<int>(pp / planes);
        const uint8_t* src = local + plane * src_plane_stride + (SCATTER ? peer * rows : 0);
        const uint4 val = __ldg(reinterpret_cast<const uint4*>(src) + v);
        *reinterpret_cast<uint4*>(peers.base[peer] + plane * dst_plane_stride + dst_row0 + 16 * v) = val;
    }
}

// uint8 NHWC codes -> float32 NHWC values fl(delta * code): 16 codes in, four float4 out per thread.
__global__ void dequant_u8_kernel(const uint4* __restrict__ x, long long n_vec, float delta, float4* __restrict__ out) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint4 v = __ldg(x + i);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            out[4 * i + k] = make_float4(__fmul_rn(static_cast<float>(w[k] & 0xffu), delta),
                                         __fmul_rn(static_cast<float>((w[k] >> 8) & 0xffu), delta),
                                         __fmul_rn(static_cast<float>((w[k] >> 16) & 0xffu), delta),
                                         __fmul_rn(static_cast<float>(w[k] >> 24), delta));
    }
}

}  // namespace qv2x

using namespace qv2x;

extern "C" int qv2x_dequant_u8(const uint8_t* d_x, long long n, float delta, float* d_out, void* stream_) {
    QV2X_REQUIRE(d_x && d_out && n >= 0 && n % 16 == 0, "qv2x_dequant_u8: null argument or n not a multiple of 16");
    QV2X_REQUIRE((reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0,
                 "qv2x_dequant_u8: pointers must be 16-byte aligned");
    if (n == 0) return 0;
    const long long n_vec = n / 16;
    const int threads = 256;
    const int grid = static_cast<int>(std::min<long long>((n_vec + threads - 1) / threads, num_sms() * 16LL));
    dequant_u8_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream_)>>>(
        reinterpret_cast<const uint4*>(d_x), n_vec, delta, reinterpret_cast<float4*>(d_out));
    g_launch_count.fetch_add(1);
    QV2X_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int qv2x_push_planes(const uint8_t* d_local, int planes, long long rows_local, long long dst_plane_stride,
                                long long dst_row0, void* const* peer_bases, int n_peers, void* stream_) {
    QV2X_REQUIRE(d_local && peer_bases, "qv2x_push_planes: null argument");
    QV2X_REQUIRE(n_peers >= 1 && n_peers <= 8, "1..8 peers");
    QV2X_REQUIRE(rows_local % 16 == 0 && dst_plane_stride % 16 == 0 && dst_row0 % 16 == 0,
                 "plane sizes and offsets must be multiples of 16 bytes");
    if (planes <= 0 || rows_local <= 0) return 0;
    PushPeers pp{};
    pp.n = n_peers;
    for (int i = 0; i < n_peers; ++i) {
        QV2X_REQUIRE(peer_bases[i] != nullptr, "null peer buffer %d", i);
        pp.base[i] = static_cast<uint8_t*>(peer_bases[i]);
    }
    const long long total = static_cast<long long>(n_peers) * planes * (rows_local / 16);
    const int threads = 256;
    const int grid = static_cast<int>(std::min<long long>((total + threads - 1) / threads, num_sms() * 4LL));
    push_planes_kernel<false><<<grid, threads, 0, static_cast<cudaStream_t>(stream_)>>>(
        d_local, planes, rows_local, rows_local, dst_plane_stride, dst_row0, pp);
    g_launch_count.fetch_add(1);
    QV2X_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int qv2x_scatter_planes(const uint8_t* d_local, int planes, long long rows_per_peer,
                                   long long dst_plane_stride, long long dst_row0, void* const* peer_bases, int n_peers,
                                   void* stream_) {
    QV2X_REQUIRE(d_local && peer_bases, "qv2x_scatter_planes: null argument");
    QV2X_REQUIRE(n_peers >= 1 && n_peers <= 8, "1..8 peers");
    QV2X_REQUIRE(rows_per_peer % 16 == 0 && dst_plane_stride % 16 == 0 && dst_row0 % 16 == 0,
                 "plane sizes and offsets must be multiples of 16 bytes");
    if (planes <= 0 || rows_per_peer <= 0) return 0;
    PushPeers pp{};
    pp.n = n_peers;
    for (int i = 0; i < n_peers; ++i) {
        QV2X_REQUIRE(peer_bases[i] != nullptr, "null peer buffer %d", i);
        pp.base[i] = static_cast<uint8_t*>(peer_bases[i]);
    }
    const long long total = static_cast<long long>(n_peers) * planes * (rows_per_peer / 16);
    const int threads = 256;
    const int grid = static_cast<int>(std::min<long long>((total + threads - 1) / threads, num_sms() * 4LL));
    push_planes_kernel<true><<<grid, threads, 0, static_cast<cudaStream_t>(stream_)>>>(
        d_local, planes, rows_per_peer, rows_per_peer * n_peers, dst_plane_stride, dst_row0, pp);
    g_launch_count.fetch_add(1);
    QV2X_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int qv2x_rowsum_u8(const uint8_t* d_x, long long n_pixels, int cstride, int cbase, int c, int32_t* d_out,
                              void* stream_) {
    QV2X_REQUIRE(d_x && d_out, "qv2x_rowsum_u8: null argument");
    QV2X_REQUIRE(c > 0 && c % 16 == 0 && cstride % 16 == 0 && cbase % 16 == 0, "channels must be multiples of 16");
    if (n_pixels <= 0) return 0;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int threads = 256;
    auto launch = [&](auto kern, int L) {
        long long want = (n_pixels * L + threads - 1) / threads;
        const int grid = static_cast<int>(std::min<long long>(want, static_cast<long long>(num_sms()) * 16));
        kern<<<grid, threads, 0, stream>>>(d_x, n_pixels, cstride, cbase, c, d_out);
        g_launch_count.fetch_add(1);
    };
    if (c <= 64) launch(rowsum_u8_kernel<4>, 4);
    else if (c <= 128) launch(rowsum_u8_kernel<8>, 8);
    else if (c <= 256) launch(rowsum_u8_kernel<16>, 16);
    else launch(rowsum_u8_kernel<32>, 32);
    QV2X_CUDA_OK(cudaGetLastError());
    return 0;
}
