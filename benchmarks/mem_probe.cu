// mem_probe.cu -- what the memory system of a B200 gives to small streaming / gather kernels.
// Sizes are those of the flow_warp levels (27 MB feature map, 20 MB RGB image): they take only a few
// microseconds at the HBM roof, so ramp-up, request width and requests in flight decide the achieved rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mem_probe mem_probe.cu && ./mem_probe
// Prints GB/s (read + written bytes) for every variant, cold (L2 flushed by a 256 MB memset) and warm.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// 1 float per thread per iteration, grid-stride, U independent loads in flight
template <int U>
__global__ void copy32(const float *__restrict__ a, float *__restrict__ b, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        float v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) v[k] = __ldg(a + i + k * stride);
#pragma unroll
        for (int k = 0; k < U; ++k) b[i + k * stride] = v[k];
    }
    for (; i < n; i += stride) b[i] = __ldg(a + i);
}
template <int U>
__global__ void copy128(const float4 *__restrict__ a, float4 *__restrict__ b, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) v[k] = __ldg(a + i + k * stride);
#pragma unroll
        for (int k = 0; k < U; ++k) b[i + k * stride] = v[k];
    }
    for (; i < n4; i += stride) b[i] = __ldg(a + i);
}
// read-only: sum into a dummy (tests the load side alone)
template <int U>
__global__ void read32(const float *__restrict__ a, float *__restrict__ sink, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.f;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        float v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) v[k] = __ldg(a + i + k * stride);
#pragma unroll
        for (int k = 0; k < U; ++k) s += v[k];
    }
    if (s == 123.456f) sink[0] = s;
}
template <int U>
__global__ void read128(const float4 *__restrict__ a, float *__restrict__ sink, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.f;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) v[k] = __ldg(a + i + k * stride);
#pragma unroll
        for (int k = 0; k < U; ++k) s += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    if (s == 123.456f) sink[0] = s;
}
// 4-tap "bilinear" gather with a constant sub-pixel shift: out[c][p] = sum of 4 neighbours (W-strided rows);
// CH channels per thread in flight (one pixel per thread, planes of HW floats), like the scalar warp kernel
template <int CH>
__global__ void gather4(const float *__restrict__ a, float *__restrict__ b, int HW, int W, int planes) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW - W - 2) return;
    for (int c0 = blockIdx.y * CH; c0 < planes; c0 += gridDim.y * CH) {
        float r[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const float *q = a + (size_t)(c0 + k) * HW + p + 1;
            r[k] = 0.25f * __ldg(q) + 0.25f * __ldg(q + 1) + 0.25f * __ldg(q + W) + 0.25f * __ldg(q + W + 1);
        }
#pragma unroll
        for (int k = 0; k < CH; ++k) b[(size_t)(c0 + k) * HW + p] = r[k];
    }
}

static float *flushbuf;
static const size_t kFlush = 256u << 20;
template <typename F>
static void timeit(const char *name, double bytes, F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int cold = 0; cold < 2; ++cold) {
        std::vector<float> ts;
        for (int it = 0; it < 25; ++it) {
            if (cold) CK(cudaMemsetAsync(flushbuf, it, kFlush));
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it >= 5) ts.push_back(ms * 1e3f);
        }
        std::sort(ts.begin(), ts.end());
        const float us = ts[ts.size() / 2];
        printf("%-34s %s  %7.2f us  %7.1f GB/s\n", name, cold ? "cold" : "warm", us, bytes / us / 1e3);
    }
    CK(cudaGetLastError());
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaMalloc(&flushbuf, kFlush));
    const int planes = 8 * 32, H = 64, W = 208, HW = H * W;
    const size_t n = (size_t)planes * HW;  // 27 MB
    float *a, *b, *sink;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4));
    // spin the clocks up
    for (int i = 0; i < 200; ++i) copy128<4><<<sms * 8, 256>>>((const float4 *)a, (float4 *)b, n / 4);
    CK(cudaDeviceSynchronize());
    printf("array: %zu floats (%.1f MB), %d SMs\n", n, n * 4 / 1e6, sms);
    for (int occ : {2, 4, 8}) {
        char nm[64];
        const int g = sms * occ;
        snprintf(nm, 64, "read32<1>  grid %d x256", g);  timeit(nm, n * 4.0, [&] { read32<1><<<g, 256>>>(a, sink, n); });
        snprintf(nm, 64, "read32<4>  grid %d x256", g);  timeit(nm, n * 4.0, [&] { read32<4><<<g, 256>>>(a, sink, n); });
        snprintf(nm, 64, "read32<16> grid %d x256", g);  timeit(nm, n * 4.0, [&] { read32<16><<<g, 256>>>(a, sink, n); });
        snprintf(nm, 64, "read128<1> grid %d x256", g);  timeit(nm, n * 4.0, [&] { read128<1><<<g, 256>>>((const float4 *)a, sink, n / 4); });
        snprintf(nm, 64, "read128<4> grid %d x256", g);  timeit(nm, n * 4.0, [&] { read128<4><<<g, 256>>>((const float4 *)a, sink, n / 4); });
        snprintf(nm, 64, "copy32<4>  grid %d x256", g);  timeit(nm, n * 8.0, [&] { copy32<4><<<g, 256>>>(a, b, n); });
        snprintf(nm, 64, "copy32<16> grid %d x256", g);  timeit(nm, n * 8.0, [&] { copy32<16><<<g, 256>>>(a, b, n); });
        snprintf(nm, 64, "copy128<4> grid %d x256", g);  timeit(nm, n * 8.0, [&] { copy128<4><<<g, 256>>>((const float4 *)a, (float4 *)b, n / 4); });
    }
    // full-grid (non persistent) scalar copy, one element per thread
    timeit("copy32 1 elt/thread, full grid", n * 8.0, [&] { copy32<1><<<(unsigned)((n + 255) / 256), 256>>>(a, b, n); });
    timeit("copy128 1 vec/thread, full grid", n * 8.0, [&] { copy128<1><<<(unsigned)((n / 4 + 255) / 256), 256>>>((const float4 *)a, (float4 *)b, n / 4); });
    // gathers: 8 images x 32 channels handled as 256 planes; gridDim.y channel slices
    for (int ys : {32, 64, 128}) {
        char nm[64];
        snprintf(nm, 64, "gather4<4> slices %d", ys);
        timeit(nm, n * 8.0, [&] { gather4<4><<<dim3((HW + 255) / 256, ys), 256>>>(a, b, HW, W, planes); });
        snprintf(nm, 64, "gather4<8> slices %d", ys);
        timeit(nm, n * 8.0, [&] { gather4<8><<<dim3((HW + 255) / 256, ys), 256>>>(a, b, HW, W, planes); });
    }
    // cudaMemcpy D2D as the driver's own answer
    timeit("cudaMemcpyAsync D2D", n * 8.0, [&] { CK(cudaMemcpyAsync(b, a, n * 4, cudaMemcpyDeviceToDevice)); });
    return 0;
}
