// Throughput of int->float conversions vs plain FP32 ops on one SM with 4/8/16 warps (the encode epilogue runs 8).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/exp_i2f tools/exp_i2f.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(int* out, int iters, long long* cyc) {
    int a[8];
    float f[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 7 + i, f[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {            // I2F
                f[i] += __int2float_rn(a[i]);
                a[i] += 3;
            } else if (MODE == 1) {     // same loop without the conversion (baseline LOP + FADD + IADD)
                f[i] += __int_as_float(a[i] & 0x3fffffff);
                a[i] += 3;
            } else if (MODE == 2) {     // magic-number conversion: valid for |x| < 2^22
                f[i] += __int_as_float(0x4B400000 + a[i]) - 12582912.f;
                a[i] += 3;
            } else if (MODE == 3) {     // float64 mantissa trick + F2F
                double d = __hiloint2double(0x43300000, a[i] ^ 0x80000000) - 4503601774854144.0;
                f[i] += (float)d;
                a[i] += 3;
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (int)s + a[0];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    int* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    const char* nm[] = {"I2F.rn + FADD + IADD", "LOP + FADD + IADD (no conversion)", "magic IADD+FADD + FADD + IADD",
                        "f64 mantissa trick + F2F.F32.F64"};
    for (int warps : {4, 8, 16}) {
        for (int mode = 0; mode < 4; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, cyc);
                if (mode == 1) k<1><<<148, warps * 32>>>(out, iters, cyc);
                if (mode == 2) k<2><<<148, warps * 32>>>(out, iters, cyc);
                if (mode == 3) k<3><<<148, warps * 32>>>(out, iters, cyc);
                cudaDeviceSynchronize();
            }
            long long h;
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("warps=%2d %-36s: %7.2f cycles per 8 elements per warp, %6.1f elements/clk/SM\n", warps, nm[mode],
                   (double)h / iters, warps * 32.0 * 8 * iters / h);
        }
    }
    return 0;
}
