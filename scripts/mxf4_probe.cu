// mxf4_probe.cu - can the Hamming contraction run on the 4-bit tensor-core path?  tcgen05.mma kind::mxf4 (e2m1 operands, UE8M0
// block scales, fp32 accumulate), 128 x 256 x 64 per instruction, operands +-4 with constant block scales 2^1 (a product is
// +-64, as in the int8 kernels), the accumulator PRE-LOADED with 2^23 + 16384 so that its low 16 bits are the integer.
// Questions: (1) is every accumulator bit exact, (2) how many clocks per instruction.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mxf4_probe scripts/mxf4_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes) {   // K-major, no swizzle: LBO 128 B, version 1
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                 "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
                 "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]),
                 "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// sign of element (row, k): a fixed pseudo-random function, the same on the host
__host__ __device__ inline int sgn(uint32_t seed, int row, int k) {
    uint32_t x = seed ^ (uint32_t)(row * 0x9E3779B1u) ^ (uint32_t)(k * 0x85EBCA77u);
    x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12; x *= 0x297A2D39u; x ^= x >> 15;
    return (x & 1u) ? 1 : -1;
}

constexpr int kM = 128, kN = 256, kK = 256;             // one accumulator of compares: 128 query rows x 256 train rows x 256 bits
constexpr int kRowBytes = kK / 2;                      // 128 B of nibbles per row

// mode 0: correctness (one pass, accumulator written out); mode 1: issue rate (iters passes, two accumulators in turn)
__global__ void __launch_bounds__(128, 1) probe(uint32_t* out, long long* cycles, int iters, int mode, uint32_t init_bits, uint32_t sf_byte,
                                                uint32_t nib_plus, uint32_t nib_minus) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;                          // 128 rows x 128 B
    uint8_t* sB = smem + kM * kRowBytes;         // 256 rows x 128 B
    uint8_t* sTA = sB + kN * kRowBytes;          // 128 rows x 32 B  e5m2: [2048, 2048, 128, 64, 8, 1, 0 ...]
    uint8_t* sTB = sTA + kM * 32;                // 256 rows x 32 B  e5m2: [2048, 2048, 128, d2, d1, d0, 0 ...], 127 - (row & 127) = 64 d2 + 8 d1 + d0
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // operands: element k of a row lives in byte k / 2, low nibble first; canonical K-major layout with 16-byte core-matrix rows
    for (int i = tid; i < (kM + kN) * kRowBytes; i += 128) {
        const bool isB = i >= kM * kRowBytes;
        const int j = isB ? i - kM * kRowBytes : i;
        const int row = j / kRowBytes, byte = j % kRowBytes;
        const int s0 = sgn(isB ? 77u : 11u, row, 2 * byte), s1 = sgn(isB ? 77u : 11u, row, 2 * byte + 1);
        const uint8_t v = (uint8_t)((s0 > 0 ? nib_plus : nib_minus) | ((s1 > 0 ? nib_plus : nib_minus) << 4));
        (isB ? sB : sA)[(row >> 3) * (8 * kRowBytes) + (byte >> 4) * 128 + (row & 7) * 16 + (byte & 15)] = v;
    }
    for (int r = tid; r < kM + kN; r += 128) {
        const bool isB = r >= kM;
        const int row = isB ? r - kM : r;
        const int v = 127 - (row & 127);
        const uint8_t dig[8] = {0x00, 0x3C, 0x40, 0x42, 0x44, 0x45, 0x46, 0x47};      // e5m2 of 0..7
        uint8_t e[32] = {0};
        e[0] = 0x68; e[1] = 0x68; e[2] = 0x58;                                          // 2048, 2048, 128
        if (isB) { e[3] = dig[v >> 6]; e[4] = dig[(v >> 3) & 7]; e[5] = dig[v & 7]; }
        else { e[3] = 0x54; e[4] = 0x48; e[5] = 0x3C; }                                 // 64, 8, 1
        uint8_t* dst = (isB ? sTB : sTA) + (row >> 3) * 256 + (row & 7) * 16;
        for (int k = 0; k < 32; ++k) dst[(k >> 4) * 128 + (k & 15)] = e[k];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
    {   // scale factors: columns 480..511 of every lane hold the same byte four times; accumulator 0 starts from `init_bits`
        uint32_t v[32];
        for (int j = 0; j < 32; ++j) v[j] = sf_byte * 0x01010101u;
        st32(lane_base + 480, v);
        for (int j = 0; j < 32; ++j) v[j] = init_bits;
        for (int c = 0; c < 256; c += 32) st32(lane_base + c, v);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 32) {
        // a, b format E2M1 (MXF4Format 1) | K-major | N >> 3 at [17,23) | scale format UE8M0 at 23 | M >> 4 at [24,29)
        const uint32_t idesc = (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | (1u << 23) | ((uint32_t)(kM >> 4) << 24);
        const uint64_t ad = make_desc(smem_u32(sA), 8 * kRowBytes), bd = make_desc(smem_u32(sB), 8 * kRowBytes);
        const uint32_t sfa = tm + 480, sfb = tm + 484;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t acc = tm + (mode & 1 ? (uint32_t)(it & 1) * 224u : 0u);       // (rate modes: the second accumulator overlaps nothing we read)
            if (mode >= 2) {
                // kind::f8f6f4, e5m2 x e5m2, K = 32, NOT accumulating: the accumulator becomes 2^23 + 16384 + 127 - (column & 127)
                const uint32_t idt = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
                const uint64_t ta = make_desc(smem_u32(sTA), 256), tb = make_desc(smem_u32(sTB), 256);
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 0, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%4, %4, %4, %4}, p;\n}\n"
                             ::"r"(acc), "l"(ta), "l"(tb), "r"(idt), "r"(0u) : "memory");
            }
#pragma unroll
            for (int k = 0; k < kK / 64; ++k) {
                const uint32_t en = (mode == 0 || mode >= 2 || k > 0) ? 1u : 0u;           // mode 0: onto the pre-loaded values; 2, 3: onto the tail
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n}\n"
                             ::"r"(acc), "l"(ad + (uint64_t)(k * 16)), "l"(bd + (uint64_t)(k * 16)), "r"(idesc), "r"(en), "r"(sfa), "r"(sfb) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[it & 1])) : "memory");
            if (it >= 1) { uint32_t spins = 0; while (!mbar_try(&bar[(it - 1) & 1], (uint32_t)(((it - 1) >> 1) & 1))) if (++spins > (1u << 24)) __trap(); }
        }
        { uint32_t spins = 0; while (!mbar_try(&bar[(iters - 1) & 1], (uint32_t)(((iters - 1) >> 1) & 1))) if (++spins > (1u << 24)) __trap(); }
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if ((mode == 0 || mode == 2) && blockIdx.x == 0) {
        uint32_t r[32];
        for (int c = 0; c < 256; c += 32) {
            ld32(lane_base + c, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + lane) * 256 + c + j] = r[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

static int check(const char* what, float init, uint32_t sf_byte, uint32_t nib_plus, uint32_t nib_minus, float unit, int mode = 0) {
    uint32_t* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, kM * kN * 4)); CK(cudaMalloc(&d_cyc, 148 * 8));
    uint32_t init_bits; memcpy(&init_bits, &init, 4);
    const size_t smem = (kM + kN) * kRowBytes + (kM + kN) * 32;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe<<<1, 128, smem>>>(d_out, d_cyc, 1, mode, init_bits, sf_byte, nib_plus, nib_minus);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[%s] kernel failed: %s\n", what, cudaGetErrorString(e)); exit(2); }
    std::vector<uint32_t> h(kM * kN);
    CK(cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < kM; ++i)
        for (int j = 0; j < kN; ++j) {
            int dot = 0;
            for (int k = 0; k < kK; ++k) dot += sgn(11u, i, k) * sgn(77u, j, k);
            const float want = (mode == 2 ? 8388608.0f + 16384.0f + (float)(127 - (j & 127)) : init) + unit * (float)dot;
            float got; memcpy(&got, &h[(size_t)i * kN + j], 4);
            if (got != want) { if (bad < 4) printf("  [%s] (%d,%d): got %.1f (0x%08x) want %.1f\n", what, i, j, got, h[(size_t)i * kN + j], want); ++bad; }
        }
    printf("[%s] %s (%d of %d differ)\n", what, bad ? "MISMATCH" : "exact", bad, kM * kN);
    cudaFree(d_out); cudaFree(d_cyc);
    return bad;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
    int bad = 0;
    bad += check("+-1, scales 2^0, from 0", 0.0f, 127, 0x2, 0xA, 1.0f);
    bad += check("+-4, scales 2^1, from 0 (unit 64)", 0.0f, 128, 0x6, 0xE, 64.0f);
    bad += check("+-4, scales 2^1, from 2^23 + 16384", 8388608.0f + 16384.0f, 128, 0x6, 0xE, 64.0f);
    bad += check("+-4, scales 2^1, onto a kind::f8f6f4 e5m2 instruction that sets 2^23 + 16384 + 127 - column", 12345.0f, 128, 0x6, 0xE, 64.0f, 2);
    // issue rate
    uint32_t* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, kM * kN * 4)); CK(cudaMalloc(&d_cyc, 148 * 8));
    const size_t smem = (kM + kN) * kRowBytes + (kM + kN) * 32;
    const int iters = 4000;
    for (int rep = 0; rep < 2; ++rep) {
        probe<<<p.multiProcessorCount, 128, smem>>>(d_out, d_cyc, iters, 1, 0u, 127, 0x2, 0xA);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(p.multiProcessorCount);
        CK(cudaMemcpy(h.data(), d_cyc, h.size() * 8, cudaMemcpyDeviceToHost));
        long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
        printf("kind::mxf4 128 x 256 x 64: %.1f clocks per instruction (= %.1f per 128 x 256 x 256 bits; the int8 path needs 8 x 128 = 1024)\n",
               (double)mx / (4.0 * iters), (double)mx / iters);
    }
    for (int rep = 0; rep < 2; ++rep) {
        probe<<<p.multiProcessorCount, 128, smem>>>(d_out, d_cyc, iters, 3, 0u, 128, 0x6, 0xE);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(p.multiProcessorCount);
        CK(cudaMemcpy(h.data(), d_cyc, h.size() * 8, cudaMemcpyDeviceToHost));
        long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
        printf("1 x kind::f8f6f4 (K = 32) + 4 x kind::mxf4 (K = 64) per accumulator: %.1f clocks (knn2_mmak_kernel: 9 x 128 = 1152)\n", (double)mx / iters);
    }
    return bad ? 1 : 0;
}
