// ffma_peak.cu -- microbenchmark: FP32 FFMA vs packed FFMA2 issue rate on sm_100a, and LDS.128 rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o benchmarks/_build/ffma_peak benchmarks/ffma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k_fma(float *out, int iters, float s) {
    float a[32], b[8];
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = s + i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], b[j], b[(j + 1) & 7]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float2 x = make_float2(a[i], a[i + 1]);
                    float2 y = make_float2(b[j], b[j]);
                    float2 z = make_float2(b[(j + 1) & 7], b[(j + 1) & 7]);
                    x = __ffma2_rn(x, y, z);
                    a[i] = x.x; a[i + 1] = x.y;
                }
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// LDS.128 rate: every lane reads conflict-free 16-byte chunks in a loop
__global__ void __launch_bounds__(256) k_lds(float *out, int iters) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float4 v = sm[(idx + j * 32) & 1023];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx += 7;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000, grid = 148 * 8;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k_fma<0><<<grid, 256>>>(out, iters, 1.0001f);
            else k_fma<1><<<grid, 256>>>(out, iters, 1.0001f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 256.0 * grid * iters * 256.0;
            printf("%s rep %d: %.3f ms  %.1f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", rep, ms, flops / ms / 1e9);
        }
    }
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_lds<<<grid, 256>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double bytes = 16.0 * 16 * 256.0 * grid * iters;
        printf("LDS.128 rep %d: %.3f ms  %.1f TB/s  (%.1f B/clk/SM at 1.9 GHz)\n", rep, ms, bytes / ms / 1e9,
               bytes / ms / 1e6 / 148 / 1.9e3);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
