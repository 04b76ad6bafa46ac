// Which pipe carries the packed 16-bit min/max forms on sm_100a?  Times long dependent-free streams of
//   0: max/min.u16x2 (VIMNMX)   1: max/min.f16x2 (HMNMX2)   2: both, half and half   3: max.bf16x2   4: u16x2 + bf16x2
// with 8 warps per SM sub-partition so the issue rate is the pipe's, not the latency's.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/minmax_pipe_probe scripts/minmax_pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t umax2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("max.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t umin2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("min.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hmax2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hmin2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t bmax2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t bmin2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("min.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }

template <int MODE>
__global__ void __launch_bounds__(1024) probe(uint32_t* out, const uint32_t* in, int iters, long long* clk) {
    uint32_t x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = in[threadIdx.x + i * 1024] & 0x3FFF3FFFu; y[i] = in[threadIdx.x + (8 + i) * 1024] & 0x3FFF3FFFu; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0)      { x[i] = umax2(x[i], y[i]); y[i] = umin2(y[i], x[(i + 1) & 7]); }
            else if (MODE == 1) { x[i] = hmax2(x[i], y[i]); y[i] = hmin2(y[i], x[(i + 1) & 7]); }
            else if (MODE == 2) { x[i] = umax2(x[i], y[i]); y[i] = hmin2(y[i], x[(i + 1) & 7]); }
            else if (MODE == 3) { x[i] = bmax2(x[i], y[i]); y[i] = bmin2(y[i], x[(i + 1) & 7]); }
            else                { x[i] = umax2(x[i], y[i]); y[i] = bmin2(y[i], x[(i + 1) & 7]); }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i] ^ y[i];
    out[blockIdx.x * 1024 + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

int main() {
    uint32_t *in, *out; long long* clk;
    cudaMalloc(&in, 16 * 1024 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
    uint32_t* h = new uint32_t[16 * 1024];
    for (int i = 0; i < 16 * 1024; ++i) h[i] = 0x12345u * (i + 7) ^ (i << 13);
    cudaMemcpy(in, h, 16 * 1024 * 4, cudaMemcpyHostToDevice);
    const int iters = 4096;
    const char* names[5] = {"max/min.u16x2", "max/min.f16x2", "u16x2 + f16x2", "max/min.bf16x2", "u16x2 + bf16x2"};
    for (int m = 0; m < 5; ++m) {
        for (int rep = 0; rep < 2; ++rep) {
            if (m == 0) probe<0><<<148, 1024>>>(out, in, iters, clk);
            if (m == 1) probe<1><<<148, 1024>>>(out, in, iters, clk);
            if (m == 2) probe<2><<<148, 1024>>>(out, in, iters, clk);
            if (m == 3) probe<3><<<148, 1024>>>(out, in, iters, clk);
            if (m == 4) probe<4><<<148, 1024>>>(out, in, iters, clk);
        }
        long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
        // per sub-partition: 8 warps x 16 instructions per iteration
        printf("%-16s %lld clocks, %.3f clocks per warp instruction per sub-partition (%s)\n", names[m], c, double(c) / (double(iters) * 8 * 16),
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
