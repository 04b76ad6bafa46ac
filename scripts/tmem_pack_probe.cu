// tmem_pack_probe.cu - what does tcgen05.ld.32x32b.x32.pack::16b return?  Column c of lane l holds (l << 16 | c) + 0x70000000
// (written with tcgen05.st.32x32b.x32); the packed load of 32 registers is printed for lane 0 and lane 37.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "../uzliti_slam_b200/csrc/uz_knn2_mma.cuh"
using namespace uz;
#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e__), __LINE__); exit(2); } } while (0)
__global__ void __launch_bounds__(128, 1) k(uint32_t* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
    uint32_t v[32];
    for (int chunk = 0; chunk < 4; ++chunk) {
        for (int j = 0; j < 32; ++j) v[j] = 0x70000000u + ((uint32_t)(warp * 32 + lane) << 16) + (uint32_t)(chunk * 32 + j);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                     "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                     "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                     ::"r"(base + chunk * 32), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                     "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
                     "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]),
                     "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(base)
        : "memory");
    tc_wait_ld(); tc_pin(r);
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = r[j];
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(128u) : "memory"); }
}
int main() {
    uint32_t* o; CK(cudaMalloc(&o, 128 * 32 * 4));
    k<<<1, 128>>>(o); CK(cudaDeviceSynchronize());
    static uint32_t h[128 * 32]; CK(cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost));
    for (int l : {0, 37}) { printf("lane %d:", l); for (int j = 0; j < 32; ++j) printf(" %08x", h[l * 32 + j]); printf("\n"); }
    return 0;
}
