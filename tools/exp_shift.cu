// Micro-experiment (bring-up tool, not product code): can ONE shared-memory copy of an activation halo serve all
// nine taps of a 3x3 conv through SHIFTED tcgen05 shared-memory descriptors?
//
// A halo of hh x hw pixels x CB channel bytes (CB = 128: SWIZZLE_128B, CB = 64: SWIZZLE_64B) is written once by a
// 4-D TMA box.  The GEMM tile is an 8-wide x 16-high pixel box, so an 8-row group of the UMMA operand is 8 consecutive
// halo pixels and consecutive groups are hw pixels apart (stride byte offset = hw * CB).  Tap (ky, kx) only moves the
// descriptor's start address by (ky * hw + kx) * CB bytes -- NOT a multiple of the swizzle pattern.  Whether the
// hardware swizzles on absolute address bits (then any hw works) or on a row counter plus the descriptor's base-offset
// field (then hw must keep the stride a multiple of the pattern and the field must be set) is not documented;
// this program measures it.  Usage: exp_shift  (prints mismatch counts per variant).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../quantv2x_b200/csrc/host_common.h"
#include "../quantv2x_b200/csrc/ptx.cuh"

using namespace qv2x;

constexpr int kN = 32;          // GEMM columns per tap
constexpr int kHH = 18;         // halo rows (16 + 2)

struct Params {
    int hw, cb, bo_mode;        // halo width (pixels), channel bytes per pixel, base-offset mode
};

__global__ void __launch_bounds__(128, 1)
shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Params p, uint32_t idesc,
             int32_t* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                         // halo: hh * hw * cb bytes (<= 18 * 16 * 128 = 36 KB)
    uint8_t* sB = smem + 60 * 1024;             // 9 taps x kN rows x cb bytes (<= 36 KB)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 100 * 1024);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_ptr), 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (threadIdx.x == 0) {
        const uint32_t fb = smem_u32(&bars[0]);
        mbar_expect_tx(fb, kHH * p.hw * p.cb + 9 * kN * p.cb);
        tma_load_4d(smem_u32(sA), &tmA, fb, 0, 0, 0, 0);
        for (int t = 0; t < 9; ++t) tma_load_2d(smem_u32(sB) + t * kN * p.cb, &tmB, fb, t * p.cb, 0);
        mbar_wait(fb, 0);
        tcgen05_fence_after();
        const uint64_t layout = (p.cb == 128) ? 2ull : 4ull;
        for (int t = 0; t < 9; ++t) {
            const int ky = t / 3, kx = t % 3;
            const uint32_t a_addr = smem_u32(sA) + (ky * p.hw + kx) * p.cb;
            const uint32_t b_addr = smem_u32(sB) + t * kN * p.cb;
            uint64_t ad = 0;
            ad |= static_cast<uint64_t>((a_addr & 0x3ffffu) >> 4);
            ad |= static_cast<uint64_t>(1) << 16;
            ad |= static_cast<uint64_t>((static_cast<uint32_t>(p.hw * p.cb)) >> 4) << 32;   // 8-row groups are hw pixels apart
            ad |= static_cast<uint64_t>(1) << 46;
            if (p.bo_mode == 1) ad |= static_cast<uint64_t>((a_addr >> 7) & 7u) << 49;
            ad |= layout << 61;
            const uint64_t bd = umma_smem_desc(b_addr, p.cb);
            for (int k = 0; k < p.cb / 32; ++k)
                umma_i8(tmem_base + t * kN, ad + 2 * k, bd + 2 * k, idesc, k != 0);
        }
        umma_commit(smem_u32(&bars[1]));
    }
    __syncthreads();
    mbar_wait(smem_u32(&bars[1]), 0);
    tcgen05_fence_after();
    const int row = warp * 32 + lane;
    for (int t = 0; t < 9; ++t) {
        for (int c0 = 0; c0 < kN; c0 += 16) {
            uint32_t acc[16];
            tmem_ld_x16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + t * kN + c0, acc);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) out[(t * 128 + row) * kN + c0 + j] = static_cast<int32_t>(acc[j]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static int run(int hw, int cb, int bo_mode) {
    const int H = kHH, W = hw;
    std::vector<uint8_t> hA(static_cast<size_t>(H) * W * cb), hB(static_cast<size_t>(kN) * 9 * cb);
    srand(1234 + hw * 7 + cb);
    for (auto& v : hA) v = static_cast<uint8_t>(rand() & 255);
    for (auto& v : hB) v = static_cast<uint8_t>(rand() & 255);      // int8 bit patterns
    uint8_t *dA, *dB;
    int32_t* dO;
    cudaMalloc(&dA, hA.size());
    cudaMalloc(&dB, hB.size());
    cudaMalloc(&dO, 9 * 128 * kN * 4);
    cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
    cudaMemset(dO, 0xff, 9 * 128 * kN * 4);
    CUtensorMap tmA, tmB;
    {
        const uint64_t dims[4] = {(uint64_t)cb, (uint64_t)W, (uint64_t)H, 1};
        const uint64_t strides[3] = {(uint64_t)cb, (uint64_t)cb * W, (uint64_t)cb * W * H};
        const uint32_t box[4] = {(uint32_t)cb, (uint32_t)hw, (uint32_t)kHH, 1};
        const uint32_t es[4] = {1, 1, 1, 1};
        if (encode_tmap_u8(&tmA, dA, 4, dims, strides, box, es, cb)) { printf("tmap A: %s\n", qv2x_last_error()); return -1; }
    }
    {
        const uint64_t dims[2] = {(uint64_t)9 * cb, (uint64_t)kN};
        const uint64_t strides[1] = {(uint64_t)9 * cb};
        const uint32_t box[2] = {(uint32_t)cb, (uint32_t)kN};
        const uint32_t es[2] = {1, 1};
        if (encode_tmap_u8(&tmB, dB, 2, dims, strides, box, es, cb)) { printf("tmap B: %s\n", qv2x_last_error()); return -1; }
    }
    uint32_t idesc = 0;
    idesc |= 2u << 4;
    idesc |= 1u << 10;                                  // B signed
    idesc |= static_cast<uint32_t>(kN >> 3) << 17;
    idesc |= static_cast<uint32_t>(128 >> 4) << 24;
    cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    Params p{hw, cb, bo_mode};
    shift_kernel<<<1, 128, 110 * 1024>>>(tmA, tmB, p, idesc, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("hw=%d cb=%d bo=%d: CUDA error %s\n", hw, cb, bo_mode, cudaGetErrorString(e)); return -1; }
    std::vector<int32_t> hO(9 * 128 * kN);
    cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
    printf("hw=%2d cb=%3d bo_mode=%d : mismatches per tap:", hw, cb, bo_mode);
    long total = 0;
    for (int t = 0; t < 9; ++t) {
        const int ky = t / 3, kx = t % 3;
        long bad = 0;
        for (int r = 0; r < 128; ++r) {
            const int ly = r / 8, lx = r % 8;
            const uint8_t* a = &hA[(static_cast<size_t>(ly + ky) * W + lx + kx) * cb];
            for (int n = 0; n < kN; ++n) {
                const int8_t* b = reinterpret_cast<const int8_t*>(&hB[static_cast<size_t>(n) * 9 * cb + t * cb]);
                int32_t ref = 0;
                for (int k = 0; k < cb; ++k) ref += static_cast<int32_t>(a[k]) * b[k];
                if (ref != hO[(t * 128 + r) * kN + n]) ++bad;
            }
        }
        printf(" %ld", bad);
        total += bad;
    }
    printf("  => %s\n", total == 0 ? "OK" : "WRONG");
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dO);
    return total == 0 ? 0 : 1;
}

int main() {
    for (int cb : {128, 64})
        for (int hw : {10, 16, 24})
            for (int bo : {0, 1}) run(hw, cb, bo);
    return 0;
}
