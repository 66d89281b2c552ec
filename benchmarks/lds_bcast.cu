// lds_bcast.cu -- does LDS.128 merge identical addresses ACROSS quarter-warps?  (sm_100a)
// mode 0: 32 distinct 16-byte chunks per warp instruction (4 wavefronts expected)
// mode 1: 8 lanes share each chunk, lanes (dy = lane>>2, q = lane&3) -> 4 distinct chunks per instruction
// mode 2: all 32 lanes read the same chunk
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters) {
    __shared__ float4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = MODE == 0 ? lane : MODE == 1 ? (lane & 3) : 0;
    base += warp * 64;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float4 v = sm[(base + j * 36) & 2047];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        base += 5;
    }
    out[blockIdx.x * 256 + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 4000, grid = 148 * 8;
    for (int mode = 0; mode < 3; ++mode) for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a);
        if (mode == 0) k<0><<<grid, 256>>>(out, iters); else if (mode == 1) k<1><<<grid, 256>>>(out, iters); else k<2><<<grid, 256>>>(out, iters);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double instr = 16.0 * iters * 8 * grid;  // warp-level LDS.128 instructions
        printf("mode %d: %.3f ms  %.2f cycles per LDS.128 warp-instruction per SM (1.965 GHz)\n", mode, ms,
               ms * 1e-3 * 1.965e9 / (instr / 148));
    }
    return 0;
}
