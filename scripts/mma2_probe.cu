// mma2_probe.cu - issue rate of tcgen05.mma kind::i8 with operands resident in shared memory (no loads, no epilogue):
// one CTA per SM issuing 128 x 256 x 32 (cta_group::1) against a CTA pair issuing 256 x 256 x 32 (cta_group::2, each CTA holds its
// 128 A rows and HALF of the B rows).  Question: does pairing lift the ~72 B/clk shared-memory operand fetch that holds the
// one-CTA instruction at 171 clocks (profiles/mma_experiments_r02.txt)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mma2_probe scripts/mma2_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {   // K-major, no swizzle: LBO 128 B, SBO 2048 B, version 1
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(2048 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}

template <int G>
__global__ void __launch_bounds__(128, 1) probe(long long* cycles, int iters, int n_cols) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;                 // 128 rows x 256 B
    uint8_t* sB = smem + 32768;         // G = 1: 256 rows, G = 2: 128 rows (this CTA's half of N)
    __shared__ uint64_t bar[2];         // commits alternate, so a barrier never runs more than one phase ahead of its waiter
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = G == 2 ? cta_rank() : 0;
    for (int i = tid; i < (32768 + 65536) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01ff01ffu, 0xff01ff01u, 0x0101ffffu, 0xffff0101u);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        if (G == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (G == 2) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (rank == 0 && tid == 32) {
        // c_format S32 | a, b INT8 | K-major | N >> 3 at [17,23) | M >> 4 at [24,29)
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)((G * 128) >> 4) << 24);
        const uint64_t ad = make_desc(smem_u32(sA)), bd = make_desc(smem_u32(sB));
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t acc = tm + (uint32_t)(it & 1) * 256u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t en = k > 0;
                if (G == 1)
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                                 ::"r"(acc), "l"(ad + (uint64_t)(k * 16)), "l"(bd + (uint64_t)(k * 16)), "r"(idesc), "r"(en) : "memory");
                else
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                                 ::"r"(acc), "l"(ad + (uint64_t)(k * 16)), "l"(bd + (uint64_t)(k * 16)), "r"(idesc), "r"(en) : "memory");
            }
            if (G == 1)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[it & 1])) : "memory");
            else
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                             ::"r"(smem_u32(&bar[it & 1])), "h"((uint16_t)1) : "memory");
            // keep at most two commits in flight (two accumulators), like the real pipeline
            if (it >= 1) {
                const uint32_t parity = (uint32_t)(((it - 1) >> 1) & 1);
                uint32_t spins = 0;
                while (!mbar_try(&bar[(it - 1) & 1], parity)) if (++spins > (1u << 24)) __trap();
            }
        }
        {
            const uint32_t parity = (uint32_t)(((iters - 1) >> 1) & 1);
            uint32_t spins = 0;
            while (!mbar_try(&bar[(iters - 1) & 1], parity)) if (++spins > (1u << 24)) __trap();
        }
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (G == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 0) {
        if (G == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
    }
}

// ---- the same instruction stream fed by a live load pipeline (no epilogue): producer warp streams B tiles (256 train rows x 256 B)
// from an L2-resident buffer into NS shared-memory stages with 1-D bulk copies; per tile the issuer runs 2 x 8 instructions
// (two accumulators), tcgen05.commit frees the stage.  G = 2: each CTA loads HALF of every B tile into its own shared memory,
// CTA 1 relays "my half has landed" to the leader with a remote mbarrier arrive, the leader's commit frees the stage in both.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try(bar, parity)) if (++spins > (1u << 24)) __trap();
}
__device__ __forceinline__ bool mbar_try_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}

template <int G, int NS>
__global__ void __launch_bounds__(96, 1) pipe(const uint8_t* __restrict__ src, size_t src_bytes, long long* cycles, int tiles) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr uint32_t kStage = 65536 / G;          // this CTA's bytes of one B tile
    uint8_t* sA = smem;                             // 2 x 32 KB
    uint8_t* sB = smem + 65536;                     // NS stages
    __shared__ uint64_t b_full[NS], b_empty[NS], b_peer[NS], done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = G == 2 ? cta_rank() : 0;
    for (int i = tid; i < (int)((65536 + NS * kStage) / 16); i += 96) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01ff01ffu, 0xff01ff01u, 0x0101ffffu, 0xffff0101u);
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&b_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&b_empty[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&b_peer[s])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        if (G == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (G == 2) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    const size_t n_src_tiles = src_bytes / 65536;
    if (tid == 0) {
        // producer: tile t of this CTA / cluster comes from a pseudo-random 64 KB slot of the buffer
        uint32_t x = 0x9E3779B9u * (uint32_t)(blockIdx.x / G + 1);
        for (int t = 0; t < tiles; ++t) {
            const int s = t % NS;
            mbar_wait_spin(&b_empty[s], (uint32_t)(((t / NS) & 1) ^ 1));
            x = x * 1664525u + 1013904223u;
            const uint8_t* from = src + (size_t)(x % n_src_tiles) * 65536 + (size_t)rank * kStage;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&b_full[s])), "r"(kStage) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sB + (size_t)s * kStage)), "l"(from), "r"(kStage), "r"(smem_u32(&b_full[s])) : "memory");
        }
    } else if (G == 2 && rank == 1 && tid == 64) {
        // relay: this CTA's half of stage s has landed -> tell the leader
        for (int t = 0; t < tiles; ++t) {
            const int s = t % NS;
            mbar_wait_spin(&b_full[s], (uint32_t)((t / NS) & 1));
            uint32_t ra;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(&b_peer[s])), "r"(0));
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
        }
    } else if (rank == 0 && tid == 32) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)((G * 128) >> 4) << 24);
        const long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            const int s = t % NS;
            mbar_wait_spin(&b_full[s], (uint32_t)((t / NS) & 1));
            if (G == 2) { uint32_t spins = 0; while (!mbar_try_cluster(&b_peer[s], (uint32_t)((t / NS) & 1))) if (++spins > (1u << 24)) __trap(); }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t bd = make_desc(smem_u32(sB + (size_t)s * kStage));
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint64_t ad = make_desc(smem_u32(sA + i * 32768));
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t en = k > 0;
                    if (G == 1)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                                     ::"r"(tm + (uint32_t)i * 256u), "l"(ad + (uint64_t)(k * 16)), "l"(bd + (uint64_t)(k * 16)), "r"(idesc), "r"(en) : "memory");
                    else
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                                     ::"r"(tm + (uint32_t)i * 256u), "l"(ad + (uint64_t)(k * 16)), "l"(bd + (uint64_t)(k * 16)), "r"(idesc), "r"(en) : "memory");
                }
            }
            if (G == 1)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&b_empty[s])) : "memory");
            else
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                             ::"r"(smem_u32(&b_empty[s])), "h"((uint16_t)3) : "memory");
        }
        if (G == 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
        else
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&done)), "h"((uint16_t)1) : "memory");
        mbar_wait_spin(&done, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (G == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 0) {
        if (G == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
    }
}

template <int G, int NS>
static void run_pipe(int sms, int tiles, const uint8_t* src, size_t src_bytes) {
    long long* d;
    const int grid = G == 2 ? (sms / 2) * 2 : sms;
    CK(cudaMalloc(&d, grid * sizeof(long long)));
    CK(cudaMemset(d, 0, grid * sizeof(long long)));
    const size_t smem = 65536 + (size_t)NS * (65536 / G);
    CK(cudaFuncSetAttribute(pipe<G, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(96); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        CK(cudaLaunchKernelEx(&cfg, pipe<G, NS>, src, src_bytes, d, tiles));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> h(grid);
        CK(cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0; int cnt = 0;
        for (int i = 0; i < grid; ++i) if (h[i]) { mx = h[i] > mx ? h[i] : mx; ++cnt; }
        printf("pipeline cta_group::%d, %d stages of %d KB, source %zu MB: %d issuers, %.1f clocks per instruction (128 x 256 x 32 per SM), kernel %.3f ms\n",
               G, NS, 64 / G, src_bytes >> 20, cnt, (double)mx / (16.0 * tiles), ms);
    }
    cudaFree(d);
}

template <int G>
static void run(int sms, int iters, int n_cols) {
    long long* d;
    const int grid = G == 2 ? (sms / 2) * 2 : sms;
    CK(cudaMalloc(&d, grid * sizeof(long long)));
    CK(cudaMemset(d, 0, grid * sizeof(long long)));
    const size_t smem = 32768 + 65536;
    CK(cudaFuncSetAttribute(probe<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        CK(cudaLaunchKernelEx(&cfg, probe<G>, d, iters, n_cols));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> h(grid);
        CK(cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0, mn = 1ll << 60; int cnt = 0;
        for (int i = 0; i < grid; ++i) if (h[i]) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; ++cnt; }
        const double per = (double)mx / (8.0 * iters);
        // work per instruction: G * 128 x n_cols x 32 MACs on G SMs -> clocks per (128 x 256 x 32)-equivalent per SM
        printf("cta_group::%d  M=%d N=%d: %d issuers, %.1f clocks per instruction (min %.1f), = %.1f clocks per 128x256x32 per SM, kernel %.3f ms\n",
               G, G * 128, n_cols, cnt, per, (double)mn / (8.0 * iters), per * 256.0 / n_cols, ms);
    }
    cudaFree(d);
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int iters = argc > 1 ? atoi(argv[1]) : 4000;
    printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
    run<1>(p.multiProcessorCount, iters, 256);
    run<1>(p.multiProcessorCount, iters, 128);
    run<2>(p.multiProcessorCount, iters, 256);
    run<2>(p.multiProcessorCount, iters, 128);
    const int tiles = iters / 2;
    for (size_t mb : {(size_t)32, (size_t)2048}) {
        uint8_t* src;
        CK(cudaMalloc(&src, mb << 20));
        CK(cudaMemset(src, 1, mb << 20));
        run_pipe<1, 2>(p.multiProcessorCount, tiles, src, mb << 20);
        run_pipe<2, 2>(p.multiProcessorCount, tiles, src, mb << 20);
        run_pipe<2, 4>(p.multiProcessorCount, tiles, src, mb << 20);
        cudaFree(src);
    }
    return 0;
}
