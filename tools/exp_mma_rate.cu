// Micro-experiment (bring-up tool, not product code): what paces a tcgen05 kind::i8 mainloop on this B200?
//   * MMA rate vs N (64 / 128 / 256) with operands resident in shared memory (no loads at all)
//   * the same with the A operand read through a shifted "halo" descriptor (8-row groups 10 pixels apart)
//   * the same while TMA streams `bytes` per k-block (128 bytes of K) into a ring: B only (N x 128) or A + B
// Every CTA (one per SM) reports its cycles (clock64) and nanoseconds (globaltimer), so the effective SM clock under
// the load is measured too.  Usage: exp_mma_rate
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../quantv2x_b200/csrc/host_common.h"
#include "../quantv2x_b200/csrc/ptx.cuh"

using namespace qv2x;

struct Params {
    int n;            // MMA N
    int halo;         // 1: A through a shifted halo descriptor (stride 10 pixels), tap cycles 0..8
    int a_rows;       // rows of A streamed per k-block (0 or 128)
    int b_rows;       // rows of B streamed per k-block (0 or N)
    int iters;        // k-blocks (128 bytes of K each = 4 MMAs) per CTA
    int stages;
    int stage_bytes;
    int n_acc;        // independent TMEM accumulators the MMAs rotate over (1 = one dependent chain)
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(128, 1)
rate_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Params p, uint32_t idesc,
            unsigned long long* stats) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sHalo = smem;                       // 24 KB resident A halo
    uint8_t* ring = smem + 24 * 1024;            // stages x stage_bytes: [A 16 KB (optional)] [B n x 128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 24 * 1024 + p.stages * p.stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + 8;
    uint64_t* done = bars + 16;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 17);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) {
            mbar_init(smem_u32(&full[i]), 1);
            mbar_init(smem_u32(&empty[i]), 1);
        }
        mbar_init(smem_u32(done), 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_ptr), 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const bool streaming = (p.a_rows + p.b_rows) > 0;
    const int a_bytes = p.a_rows * 128;
    unsigned long long c0 = 0, t0 = 0;
    if (warp == 1) {
        c0 = clock64();
        t0 = gtimer();
    }
    if (warp == 0 && streaming) {
        const bool leader = elect_one();
        uint32_t s = 0, ph = 0;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(smem_u32(&empty[s]), ph ^ 1);
            if (leader) {
                const uint32_t fb = smem_u32(&full[s]);
                mbar_expect_tx(fb, a_bytes + p.b_rows * 128);
                const uint32_t st = smem_u32(ring) + s * p.stage_bytes;
                if (p.a_rows) tma_load_2d(st, &tmA, fb, 0, ((it * 131 + blockIdx.x * 17) & 1023) * 128);
                if (p.b_rows) tma_load_2d(st + a_bytes, &tmB, fb, (it % 18) * 128, 0);
            }
            if (++s == static_cast<uint32_t>(p.stages)) s = 0, ph ^= 1;
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        uint32_t s = 0, ph = 0;
        int tap = 0;
        for (int it = 0; it < p.iters; ++it) {
            if (streaming) {
                mbar_wait(smem_u32(&full[s]), ph);
                tcgen05_fence_after();
            }
            if (leader) {
                const uint32_t st = smem_u32(ring) + s * p.stage_bytes;
                uint64_t ad, bd;
                if (p.halo) {
                    const int ky = tap / 3, kx = tap - 3 * ky;
                    const uint32_t a_addr = smem_u32(sHalo) + (ky * 10 + kx) * 128;
                    ad = static_cast<uint64_t>((a_addr & 0x3ffffu) >> 4) | (1ull << 16) |
                         (static_cast<uint64_t>((10u * 128u) >> 4) << 32) | (1ull << 46) | (2ull << 61);
                } else {
                    ad = umma_smem_desc(p.a_rows ? st : smem_u32(sHalo), 128);
                }
                bd = umma_smem_desc(st + a_bytes, 128);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_i8(tmem_base + ((it * 4 + k) % p.n_acc) * p.n, ad + 2 * k, bd + 2 * k, idesc, 1);
                if (streaming) umma_commit(smem_u32(&empty[s]));
            }
            __syncwarp();
            if (++tap == 9) tap = 0;
            if (++s == static_cast<uint32_t>(p.stages)) s = 0, ph ^= 1;
        }
        if (leader) umma_commit(smem_u32(done));
        __syncwarp();
        mbar_wait(smem_u32(done), 0);
        const unsigned long long c1 = clock64(), t1 = gtimer();
        if (leader) {
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


// Lean issue loop: descriptors hoisted, nothing but MMAs in the loop body (is the ~97-cycle floor per MMA hardware or
// the issuing thread's instruction stream?)
template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) lean_kernel(int iters, uint32_t idesc, unsigned long long* stats) {
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
        const unsigned long long c0 = clock64(), t0 = gtimer();
        if (elect_one()) {
            const uint64_t ad = umma_smem_desc(smem_u32(smem), 128);
            const uint64_t bd = umma_smem_desc(smem_u32(smem) + 32 * 1024, 128);
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_i8(tmem_base + (k % NACC) * N, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, 1);
            }
            umma_commit(smem_u32(&bars[0]));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bars[0]), 0);
        const unsigned long long c1 = clock64(), t1 = gtimer();
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

template <int N, int NACC>
static void run_lean(int iters) {
    const int smem_bytes = 100 * 1024;
    uint32_t idesc = (2u << 4) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    cudaFuncSetAttribute(lean_kernel<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    unsigned long long* d_stats;
    cudaMalloc(&d_stats, 148 * 2 * 8);
    std::vector<unsigned long long> st(296);
    for (int rep = 0; rep < 3; ++rep) {
        lean_kernel<N, NACC><<<148, 128, smem_bytes>>>(iters, idesc, d_stats);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(st.data(), d_stats, 296 * 8, cudaMemcpyDeviceToHost);
    double cyc = 0, ns = 0;
    for (int i = 0; i < 148; ++i) cyc += st[2 * i], ns += st[2 * i + 1];
    cyc /= 148, ns /= 148;
    const double macs = static_cast<double>(iters) * 8 * 128.0 * N * 32;
    printf("lean N=%3d acc=%d: %6.1f cyc/MMA  %6.0f MAC/clk/SM  %5.2f GHz\n", N, NACC, cyc / (iters * 8.0), macs / cyc, cyc / ns);
    cudaFree(d_stats);
}

static uint8_t* g_src = nullptr;

static void run(int n, int halo, int a_rows, int b_rows, int iters, int n_acc = 1) {
    Params p{};
    p.n_acc = n_acc;
    p.n = n;
    p.halo = halo;
    p.a_rows = a_rows;
    p.b_rows = b_rows;
    p.iters = iters;
    p.stage_bytes = a_rows * 128 + n * 128;        // the B slot exists even when B is not streamed
    p.stages = std::min(8, (180 * 1024) / p.stage_bytes);
    const int smem_bytes = 24 * 1024 + p.stages * p.stage_bytes + 1024 + 256;
    CUtensorMap tmA, tmB;
    {
        const uint64_t dims[2] = {128, 1024 * 128 + 128};
        const uint64_t strides[1] = {128};
        const uint32_t box[2] = {128, static_cast<uint32_t>(a_rows ? a_rows : 128)};
        const uint32_t es[2] = {1, 1};
        if (encode_tmap_u8(&tmA, g_src, 2, dims, strides, box, es, 128)) { printf("tmap: %s\n", qv2x_last_error()); return; }
    }
    {
        const uint64_t dims[2] = {2304, 256};
        const uint64_t strides[1] = {2304};
        const uint32_t box[2] = {128, static_cast<uint32_t>(b_rows ? b_rows : 64)};
        const uint32_t es[2] = {1, 1};
        if (encode_tmap_u8(&tmB, g_src + (64 << 20), 2, dims, strides, box, es, 128)) { printf("tmap: %s\n", qv2x_last_error()); return; }
    }
    uint32_t idesc = (2u << 4) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    unsigned long long* d_stats;
    cudaMalloc(&d_stats, 148 * 2 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    std::vector<unsigned long long> st(296);
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        rate_kernel<<<148, 128, smem_bytes>>>(tmA, tmB, p, idesc, d_stats);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) {
            best = ms;
            cudaMemcpy(st.data(), d_stats, 296 * 8, cudaMemcpyDeviceToHost);
        }
    }
    double cyc = 0, ns = 0;
    for (int i = 0; i < 148; ++i) cyc += st[2 * i], ns += st[2 * i + 1];
    cyc /= 148, ns /= 148;
    const double macs = static_cast<double>(iters) * 4 * 128.0 * n * 32;
    printf("N=%3d acc=%d halo=%d stream A=%3d B=%3d rows (%5d B/kblock, %d stages): %8.1f us  %7.1f cyc/kblock  %6.0f MAC/clk/SM  "
           "%5.2f GHz  %7.1f TOP/s  fill %5.1f B/clk/SM\n",
           n, n_acc, halo, a_rows, b_rows, (a_rows + b_rows) * 128, p.stages, best * 1e3, cyc / iters, macs / cyc, cyc / ns,
           2 * macs * 148 / (best * 1e-3) / 1e12, (a_rows + b_rows) * 128.0 * iters / cyc);
    cudaFree(d_stats);
}

int main() {
    cudaMalloc(&g_src, 80 << 20);
    cudaMemset(g_src, 1, 80 << 20);
    run_lean<256, 1>(10000);
    run_lean<256, 2>(10000);
    run_lean<128, 1>(10000);
    run_lean<128, 2>(10000);
    run_lean<128, 4>(10000);
    run_lean<64, 1>(10000);
    run_lean<64, 2>(10000);
    run_lean<64, 4>(10000);
    run_lean<32, 1>(10000);
    return 0;
    const int iters = 20000;
    for (int n : {256, 128, 64}) {
        run(n, 0, 0, 0, iters);          // resident operands
        run(n, 1, 0, 0, iters);          // resident, A through the halo descriptor
        run(n, 1, 0, n, iters);          // B streamed, A halo resident
        run(n, 0, 128, n, iters);        // A and B streamed (the round-1 mainloop)
    }
    for (int na : {2, 4}) {
        run(256, 1, 0, 0, iters, std::min(na, 2));
        run(128, 1, 0, 0, iters, na);
        run(128, 1, 0, 128, iters, na);
        run(128, 1, 0, 64, iters, na);
        run(64, 1, 0, 0, iters, na);
        run(64, 1, 0, 0, iters, 2 * na);
        run(64, 1, 0, 64, iters, 2 * na);
    }
    run(256, 1, 0, 128, iters);          // half of B streamed (what a CTA pair / two M tiles per B stage would need)
    run(128, 1, 0, 64, iters);
    return 0;
}
