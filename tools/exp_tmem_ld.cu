// Micro-experiment (bring-up tool): tcgen05.ld (TMEM -> registers) throughput and latency per SM, with the tensor pipe
// idle and with a full-rate int8 MMA stream running next to it.  4 or 8 reader warps (one or two per lane quadrant)
// sweep 256 accumulator columns repeatedly; variants: x16 loads with 1 / 2 / 4 loads in flight per wait, x32, x64.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../quantv2x_b200/csrc/host_common.h"
#include "../quantv2x_b200/csrc/ptx.cuh"

using namespace qv2x;

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

// mode 0: x16, wait after each; 1: x16, two in flight; 2: x16, four in flight; 3: x32, wait after each; 4: x32 two in flight
template <int MODE>
__global__ void __launch_bounds__(384, 1) ld_kernel(int iters, int mma_on, int n_readers, uint32_t idesc,
                                                    unsigned long long* stats, uint32_t* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
    volatile int* stop = reinterpret_cast<volatile int*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        fence_mbar_init();
        *stop = 0;
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_ptr), 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp == 1) {
        // MMA stream into columns [256, 512) until the readers are done
        if (mma_on && elect_one()) {
            const uint64_t ad = umma_smem_desc(smem_u32(smem), 128);
            const uint64_t bd = umma_smem_desc(smem_u32(smem) + 32 * 1024, 128);
            const unsigned long long m0 = clock64();
            unsigned long long n_mma = 0;
            while (*stop < n_readers) {
#pragma unroll
                for (int k = 0; k < 16; ++k) umma_i8(tmem_base + 256, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, 1);
                n_mma += 16;
            }
            umma_commit(smem_u32(&bars[0]));
            mbar_wait(smem_u32(&bars[0]), 0);
            stats[148 * 8 + blockIdx.x * 2] = clock64() - m0;
            stats[148 * 8 + blockIdx.x * 2 + 1] = n_mma;
        }
    } else if (warp >= 4 && warp < 4 + n_readers) {
        const int quad = warp & 3;
        const uint32_t base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + ((warp - 4) >> 2) * 128;
        uint32_t acc = 0;
        __syncwarp();
        const unsigned long long c0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if constexpr (MODE == 5) {
                // shared-memory pressure instead of TMEM loads: 64 x LDS.128 (different rows per lane: 512 B per instruction)
#pragma unroll 8
                for (int c = 0; c < 64; ++c) {
                    const int4 v = lds_i4(smem_u32(smem) + 64 * 1024 + ((c * 512 + lane * 16) & 16383));
                    acc ^= v.x ^ v.y ^ v.z ^ v.w;
                }
            } else if constexpr (MODE <= 2) {
                constexpr int FL = MODE == 0 ? 1 : (MODE == 1 ? 2 : 4);
                uint32_t r[FL][16];
#pragma unroll
                for (int c = 0; c < 128; c += 16 * FL) {
#pragma unroll
                    for (int f = 0; f < FL; ++f) tmem_ld_x16(base + c + 16 * f, r[f]);
                    tmem_ld_wait();
#pragma unroll
                    for (int f = 0; f < FL; ++f)
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc ^= r[f][j];
                }
            } else {
                constexpr int FL = MODE == 3 ? 1 : 2;
                uint32_t r[FL][32];
#pragma unroll
                for (int c = 0; c < 128; c += 32 * FL) {
#pragma unroll
                    for (int f = 0; f < FL; ++f) tmem_ld_x32(base + c + 32 * f, r[f]);
                    tmem_ld_wait();
#pragma unroll
                    for (int f = 0; f < FL; ++f)
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc ^= r[f][j];
                }
            }
        }
        const unsigned long long c1 = clock64();
        if (lane == 0) {
            stats[blockIdx.x * 8 + (warp - 4)] = c1 - c0;
            atomicAdd(const_cast<int*>(stop), 1);
        }
        sink[blockIdx.x * 384 + threadIdx.x] = acc;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int MODE>
static void run(const char* name, int mma_n, int n_readers) {
    const int iters = 2000;
    const int smem_bytes = 100 * 1024;
    const int n = mma_n ? mma_n : 128;
    uint32_t idesc = (2u << 4) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    cudaFuncSetAttribute(ld_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    unsigned long long* d_stats;
    uint32_t* d_sink;
    cudaMalloc(&d_stats, 148 * 10 * 8);
    cudaMemset(d_stats, 0, 148 * 10 * 8);
    cudaMalloc(&d_sink, 148 * 384 * 4);
    for (int rep = 0; rep < 2; ++rep) {
        ld_kernel<MODE><<<148, 384, smem_bytes>>>(iters, mma_n != 0, n_readers, idesc, d_stats, d_sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    std::vector<unsigned long long> st(148 * 10);
    cudaMemcpy(st.data(), d_stats, st.size() * 8, cudaMemcpyDeviceToHost);
    double cyc = 0;
    for (int b = 0; b < 148; ++b)
        for (int w = 0; w < n_readers; ++w) cyc += st[b * 8 + w];
    cyc /= 148.0 * n_readers;
    const double bytes_sm = static_cast<double>(iters) * 128 * 32 * 4 * n_readers;     // per SM
    double mc = 0, mn = 0;
    for (int b = 0; b < 148; ++b) mc += st[148 * 8 + 2 * b], mn += st[148 * 8 + 2 * b + 1];
    printf("%-28s mma N=%3d readers=%d: %7.1f cyc per 128-col sweep per warp, %6.1f B/clk/SM   MMA: %6.1f cyc each\n", name,
           mma_n, n_readers, cyc / iters, bytes_sm / cyc, mn > 0 ? mc / mn : 0.0);
    cudaFree(d_stats);
    cudaFree(d_sink);
}

int main() {
    for (int mma_n : {64, 128, 256})
        for (int nr : {1, 8}) {
            run<0>("x16, 1 in flight", mma_n, nr);
            run<5>("LDS.128 x64 per sweep", mma_n, nr);
        }
    return 0;
}
