// mma_peak.cu -- legacy mma.sync issue rate on sm_100a (TF32 m16n8k8, BF16 m16n8k16), to decide whether a
// banded-GEMM formulation of the correlation can beat the FFMA kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o benchmarks/_build/mma_peak benchmarks/mma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>  // 0: tf32 m16n8k8, 1: bf16 m16n8k16
__global__ void __launch_bounds__(256) k_mma(float *out, int iters) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 11u}, b[2] = {threadIdx.x * 5u, 13u};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000, grid = 148 * 4;
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k_mma<0><<<grid, 256>>>(out, iters); else k_mma<1><<<grid, 256>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double macs = (mode == 0 ? 16.0 * 8 * 8 : 16.0 * 8 * 16) * 8.0 * iters * (256 / 32) * grid;
            printf("%s rep %d: %.3f ms  %.1f TFLOP/s dense  (%.0f MAC/clk/SM at 1.9 GHz)\n", mode ? "mma.sync bf16 m16n8k16" : "mma.sync tf32 m16n8k8 ",
                   rep, ms, 2 * macs / ms / 1e9, macs / ms / 1e6 / 148 / 1.9e3);
        }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
