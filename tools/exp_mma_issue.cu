// Micro-experiment (bring-up tool): what does ONE tcgen05.mma cost the issuing thread when its shared-memory descriptors
// are produced right before it (base + constant, as in a real mainloop) instead of being loop invariant?
//   mode 0: descriptors loop-invariant (hoisted)                 -- the tensor-pipe floor
//   mode 1: 64-bit descriptor = (run-time base) + constant       -- UIADD3.64 per operand
//   mode 2: only the LOW 32-bit word is recomputed, high word kept in a fixed register (mov.b64 pack)
//   mode 3: mode 1 with the two operands sharing ONE run-time offset (one add feeds both)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../quantv2x_b200/csrc/host_common.h"
#include "../quantv2x_b200/csrc/ptx.cuh"

using namespace qv2x;

template <int N, int MODE, int CB = 128>
__global__ void __launch_bounds__(128, 1) issue_kernel(int iters, uint32_t idesc, unsigned long long* stats,
                                                       const unsigned long long* bases) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
    const int warp = threadIdx.x >> 5;
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
        const bool leader = elect_one();
        const unsigned long long c0 = clock64();
        const uint64_t ad0 = umma_smem_desc_sbo(smem_u32(smem), CB, MODE == 4 ? 8 * CB : 10 * CB);
        const uint64_t bd0 = umma_smem_desc(smem_u32(smem) + 32 * 1024, CB);
        for (int it = 0; it < iters; ++it) {
            // a run-time, warp-uniform slot offset (0 in practice) that the compiler cannot fold
            const uint64_t slot = bases[it & 7];
            if (leader) {
                const uint64_t ad = ad0 + slot, bd = bd0 + slot;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t aoff = (((tap / 3) * 10 + (tap % 3)) * CB >> 4) + 2 * (k % (CB / 32));
                        const uint32_t boff = tap * (N * CB >> 4) % 1024 + 2 * (k % (CB / 32));
                        if constexpr (MODE == 0) {
                            umma_i8(tmem_base, ad0 + 2 * k, bd0 + 2 * k, idesc, 1);
                        } else if constexpr (MODE == 1) {
                            umma_i8(tmem_base, ad + aoff, bd + boff, idesc, 1);
                        } else if constexpr (MODE == 2) {
                            uint64_t a2, b2;
                            const uint32_t alo = static_cast<uint32_t>(ad) + aoff, blo = static_cast<uint32_t>(bd) + boff;
                            asm("mov.b64 %0, {%1, %2};" : "=l"(a2) : "r"(alo), "r"(static_cast<uint32_t>(ad0 >> 32)));
                            asm("mov.b64 %0, {%1, %2};" : "=l"(b2) : "r"(blo), "r"(static_cast<uint32_t>(bd0 >> 32)));
                            umma_i8(tmem_base, a2, b2, idesc, 1);
                        } else if constexpr (MODE == 4) {      // unshifted A (tile-aligned groups), run-time base
                            umma_i8(tmem_base, ad + 2 * (k % (CB / 32)), bd + boff, idesc, 1);
                        } else {
                            umma_i8(tmem_base, ad + aoff, bd + aoff, idesc, 1);
                        }
                    }
            }
            __syncwarp();
        }
        if (leader) umma_commit(smem_u32(&bars[0]));
        __syncwarp();
        mbar_wait(smem_u32(&bars[0]), 0);
        const unsigned long long c1 = clock64();
        if (threadIdx.x == 32) stats[blockIdx.x] = c1 - c0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int N, int MODE, int CB = 128>
static void run(const unsigned long long* d_bases) {
    const int iters = 1000;
    const int smem_bytes = 100 * 1024;
    uint32_t idesc = (2u << 4) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    cudaFuncSetAttribute(issue_kernel<N, MODE, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    unsigned long long* d_stats;
    cudaMalloc(&d_stats, 148 * 8);
    for (int rep = 0; rep < 2; ++rep) {
        issue_kernel<N, MODE, CB><<<148, 128, smem_bytes>>>(iters, idesc, d_stats, d_bases);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    std::vector<unsigned long long> st(148);
    cudaMemcpy(st.data(), d_stats, 148 * 8, cudaMemcpyDeviceToHost);
    double cyc = 0;
    for (int i = 0; i < 148; ++i) cyc += st[i];
    cyc /= 148;
    printf("N=%3d CB=%3d mode=%d: %6.1f cycles per MMA\n", N, CB, MODE, cyc / (iters * 36.0));
    cudaFree(d_stats);
}

int main() {
    unsigned long long* d_bases;
    cudaMalloc(&d_bases, 64);
    cudaMemset(d_bases, 0, 64);
    run<64, 0>(d_bases); run<64, 1>(d_bases); run<64, 2>(d_bases); run<64, 3>(d_bases);
    run<128, 0>(d_bases); run<128, 1>(d_bases); run<128, 2>(d_bases); run<128, 3>(d_bases);
    run<32, 1>(d_bases);
    run<64, 1, 64>(d_bases); run<64, 4, 64>(d_bases); run<128, 1, 64>(d_bases); run<128, 4, 64>(d_bases); run<64, 4, 128>(d_bases); run<128, 4, 128>(d_bases);
    return 0;
}
