// qv2x_int8_mma_peak: the raw tcgen05.mma kind::i8 rate of this GPU, measured by the library's own issue loop
// (M = 128, N = 256, K = 32 per instruction, operands resident in shared memory, one CTA per SM, descriptors loop
// invariant -- nothing but the tensor pipe paces it).  This is the denominator of the conv kernels' roofline:
// MEASURED_PEAKS.json carries no int8 figure, and a library GEMM (cuBLASLt int8, ~3.0-3.1 POP/s on this pool) is
// itself a kernel with operand traffic and an epilogue, not the pipe's ceiling (profiles/r2_exp_mma_rate.log).
#include "host_common.h"
#include "ptx.cuh"

namespace qv2x {

__device__ __forceinline__ unsigned long long peak_gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(128, 1) int8_peak_kernel(int iters, uint32_t idesc, unsigned long long* stats) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 16 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
    fence_proxy_async_smem();
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_ptr), 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (warp == 1) {
        const unsigned long long c0 = clock64(), t0 = peak_gtimer();
        if (elect_one()) {
            const uint64_t ad = umma_smem_desc(smem_u32(smem), 128);
            const uint64_t bd = umma_smem_desc(smem_u32(smem) + 16 * 1024, 128);
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k = 0; k < 8; ++k) umma_i8(tmem_base + (k & 1) * 256, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, 1);
            }
            umma_commit(smem_u32(&bars[0]));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bars[0]), 0);
        const unsigned long long c1 = clock64(), t1 = peak_gtimer();
        if (threadIdx.x == 32) {
            stats[2 * blockIdx.x] = c1 - c0;
            stats[2 * blockIdx.x + 1] = t1 - t0;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace qv2x

extern "C" int qv2x_int8_mma_peak(double* tops, double* sm_mhz, void* stream_) {
    using namespace qv2x;
    QV2X_REQUIRE(tops != nullptr, "qv2x_int8_mma_peak: null argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int iters = 4000, smem_bytes = 66 * 1024 + 1024, sms = num_sms();
    const uint32_t idesc = (2u << 4) | (1u << 10) | (static_cast<uint32_t>(256 >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    QV2X_CUDA_OK(cudaFuncSetAttribute(int8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    unsigned long long* d_stats = nullptr;
    QV2X_CUDA_OK(cudaMalloc(&d_stats, sizeof(unsigned long long) * 2 * sms));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, stream);
        int8_peak_kernel<<<sms, 128, smem_bytes, stream>>>(iters, idesc, d_stats);
        cudaEventRecord(e1, stream);
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) {
            cudaFree(d_stats);
            return set_error(QV2X_ERR_CUDA, "int8 peak kernel failed: %s", cudaGetErrorString(e));
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    g_launch_count.fetch_add(5);
    unsigned long long h[2] = {0, 0};
    cudaMemcpy(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d_stats);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double ops = 2.0 * iters * 8.0 * 128.0 * 256.0 * 32.0 * sms;
    *tops = ops / (best * 1e-3) / 1e12;
    if (sm_mhz) *sm_mhz = h[1] ? 1e3 * static_cast<double>(h[0]) / static_cast<double>(h[1]) : 0.0;
    return 0;
}
