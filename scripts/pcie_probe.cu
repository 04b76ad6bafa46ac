// pcie_probe.cu — how fast do scattered keyframe arrays cross PCIe?  (a) one big cudaMemcpyAsync, (b) the gather kernel pulling
// from pinned mapped memory (grid sizes), (c) cudaMemcpyBatchAsync with many small copies, (d) many cudaMemcpyAsync calls.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../uzliti_slam_b200/csrc/uz_knn2.cuh"
using namespace uz;
#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e__), __LINE__); exit(2); } } while (0)
int main() {
    const size_t total = (size_t)512 << 20;
    uint8_t* h; CK(cudaHostAlloc(&h, total, cudaHostAllocMapped | cudaHostAllocPortable));
    uint8_t* d; CK(cudaMalloc(&d, total));
    for (size_t i = 0; i < total; i += 4096) h[i] = (uint8_t)i;
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, size_t bytes, auto fn) {
        fn(); CK(cudaStreamSynchronize(s));
        auto t0 = std::chrono::steady_clock::now();
        cudaEventRecord(e0, s); fn(); cudaEventRecord(e1, s);
        auto t1 = std::chrono::steady_clock::now();
        CK(cudaStreamSynchronize(s));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-52s %8.3f ms device  %6.1f GB/s   host enqueue %.3f ms\n", name, ms, bytes / ms * 1e-6, std::chrono::duration<double, std::milli>(t1 - t0).count());
    };
    timeit("one cudaMemcpyAsync 256 MB", (size_t)256 << 20, [&] { CK(cudaMemcpyAsync(d, h, (size_t)256 << 20, cudaMemcpyHostToDevice, s)); });
    // scattered: 4600 keyframes x (32000 + 24000 + 1000) B, sources at random keyframe slots
    const int nk = 4600;
    std::vector<CopyChunk> cc; std::vector<void*> dsts, srcs; std::vector<size_t> sizes;
    size_t at = 0, bytes = 0;
    srand(1);
    for (int k = 0; k < nk; ++k) {
        const size_t slot = (size_t)(rand() % 8000);
        const size_t off[3] = {slot * 32000, (size_t)8000 * 32000 + slot * 24000, (size_t)8000 * 56000 + slot * 1000};
        const size_t len[3] = {32000, 24000, 1000};
        for (int f = 0; f < 3; ++f) {
            dsts.push_back(d + at); srcs.push_back(h + off[f]); sizes.push_back(len[f]);
            for (size_t o = 0; o < len[f]; o += 16384) { CopyChunk c; c.src = h + off[f] + o; c.dst = d + at + o; c.bytes = (uint32_t)std::min<size_t>(16384, len[f] - o); c.pad = 0; cc.push_back(c); }
            at += (len[f] + 255) & ~(size_t)255; bytes += len[f];
        }
    }
    CopyChunk* dcc; CK(cudaMalloc(&dcc, cc.size() * sizeof(CopyChunk)));
    CK(cudaMemcpy(dcc, cc.data(), cc.size() * sizeof(CopyChunk), cudaMemcpyHostToDevice));
    for (int ctas : {16, 32, 64, 148, 296, 1184}) {
        char nm[96]; snprintf(nm, sizeof nm, "gather kernel %d CTAs x 128 thr, %zu chunks", ctas, cc.size());
        timeit(nm, bytes, [&] { gather_copy_kernel<<<ctas, 128, 0, s>>>(dcc, (int)cc.size()); });
    }
    timeit("gather kernel 1184 CTAs x 256 thr", bytes, [&] { gather_copy_kernel<<<1184, 256, 0, s>>>(dcc, (int)cc.size()); });
    {
        cudaMemcpyAttributes attr; memset(&attr, 0, sizeof(attr));
        attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
        size_t idx = 0, fail = 0;
        timeit("cudaMemcpyBatchAsync 13800 copies", bytes, [&] {
            cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), dsts.size(), &attr, &idx, 1, &fail, s);
            if (e != cudaSuccess) { printf("batch failed: %s (idx %zu)\n", cudaGetErrorString(e), fail); cudaGetLastError(); }
        });
        attr.srcAccessOrder = cudaMemcpySrcAccessOrderAny;
        timeit("cudaMemcpyBatchAsync 13800 copies (order any)", bytes, [&] {
            cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), dsts.size(), &attr, &idx, 1, &fail, s);
            if (e != cudaSuccess) { printf("batch failed: %s (idx %zu)\n", cudaGetErrorString(e), fail); cudaGetLastError(); }
        });
    }
    timeit("13800 x cudaMemcpyAsync", bytes, [&] { for (size_t i = 0; i < dsts.size(); ++i) cudaMemcpyAsync(dsts[i], srcs[i], sizes[i], cudaMemcpyHostToDevice, s); });
    return 0;
}
