// tmem_probe.cu — TMEM read bandwidth of tcgen05.ld.32x32b.x32 per SM (8 / 16 warps), with 1, 2 or 4 loads in flight per warp
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "../uzliti_slam_b200/csrc/uz_knn2_mma.cuh"
using namespace uz;
#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e__), __LINE__); exit(2); } } while (0)
template <int DEPTH>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, long long* cycles, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t a[32], b[32], c[32], d[32], acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t col = (uint32_t)((it * 128 + (warp >> 2) * 32) & 511) & ~31u;
        tc_ld32(base + (col & 384), a);
        if (DEPTH >= 2) tc_ld32(base + ((col + 32) & 480), b);
        if (DEPTH >= 4) { tc_ld32(base + ((col + 64) & 480), c); tc_ld32(base + ((col + 96) & 480), d); }
        tc_wait_ld();
        tc_pin(a); acc ^= a[0] ^ a[31];
        if (DEPTH >= 2) { tc_pin(b); acc ^= b[5]; }
        if (DEPTH >= 4) { tc_pin(c); tc_pin(d); acc ^= c[7] ^ d[9]; }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345u) out[0] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory"); }
}
template <int DEPTH> void run(int threads) {
    uint32_t* o; long long* c; CK(cudaMalloc(&o, 64)); CK(cudaMalloc(&c, 148 * 8));
    const int iters = 20000;
    k<DEPTH><<<148, threads>>>(o, c, iters); CK(cudaDeviceSynchronize());
    k<DEPTH><<<148, threads>>>(o, c, iters); CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost));
    const double bytes = (double)(threads / 32) * iters * DEPTH * 4096.0;
    printf("%2d warps, %d loads in flight: %.1f B/clk per SM (%.0f clk per round)\n", threads / 32, DEPTH, bytes / (double)h[0], (double)h[0] / iters);
}
int main() {
    run<1>(256); run<2>(256); run<4>(256); run<1>(512); run<2>(512); run<4>(512); run<4>(128); run<2>(128);
    return 0;
}
