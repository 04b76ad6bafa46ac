// uz_knn2_mma.cuh — K1 on the 5th-generation tensor cores: Hamming kNN-2 as an exact int8 contraction.
// This file: the operand layout, the tcgen05 helpers every tensor-core match kernel shares, and the FIRST kernel of the family
// (knn2_mma_kernel, IMAD epilogue; UZ_MATCH_MMA=7).  The default kernel is knn2_mmak_kernel (uz_knn2_mmak.cuh), 512-bit rows run
// in knn2_mmaw_kernel (uz_knn2_mmaw.cuh), the CTA-pair experiment is knn2_mma2_kernel (uz_knn2_mma2.cuh).
//
// Same contract as knn2_kernel (uz_knn2.cuh): cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, k = 2)
// (/root/reference/transformation_estimation/src/feature_transformation_estimator.cpp:38,58), two packed keys per
// query row, key = (distance << 16) | trainIdx, m1 < m2 == OpenCV's order by (distance, trainIdx).
//
// Idea ("is the POPC ceiling the right ceiling?").  With bits mapped to +-1,
//   hamming(q, t) = (256 - <q, t>) / 2
// is a dense integer contraction over K = 256, exact in int32: tcgen05.mma kind::i8 (sm_100a), 128 x 256 x 32 per
// instruction at 8192 MAC/clk/SM = one 128 x 256 tile of compares per 1024 clocks, against 128 x 256 x 4 POPC / 16 per
// clock = 8192 clocks on the XU pipe.  The north star describes the POPC design ("tensor cores are not used"); this
// kernel deliberately steps outside that sentence and is kept only because it is bit-exact and measured faster
// (DESIGN.md section 4).
//
// Data.  Descriptors are expanded ONCE at ingestion into the "E8 layout": one int8 per bit (+8 set, -8 clear: the accumulator
// holds 64 <q, t>, which is what lets knn2_mmak_kernel form the packed key inside the MMA), stored
// directly in the no-swizzle K-major canonical layout of the UMMA shared-memory descriptor, so that any run of 8-row
// groups is one contiguous byte range and a tile arrives with ONE 1-D TMA bulk copy (no tensor map):
//   byte(row i, k) = (i >> 3) * 2048 + (k >> 4) * 128 + (i & 7) * 16 + (k & 15),   k = bit index 0..255
// (core matrix = 8 rows x 16 bytes contiguous; LBO = 128 B between K-adjacent core matrices, SBO = 2048 B between
// 8-row groups).
//
// Mapping.  Persistent grid, one CTA per SM, 320 threads, warp-specialised:
//   warp 0      producer: bulk copies of the item's two 128-row query tiles (A, 2 x 32 KB) and of 256-row train tiles
//               (B, 2 stages x 64 KB), mbarrier complete_tx
//   warp 1      MMA issuer (one elected lane): per train tile and query tile 8 x tcgen05.mma 128 x 256 x 32 into one of
//               two 256-column TMEM accumulators; tcgen05.commit frees the smem stage / publishes the accumulator
//   warps 2..9  epilogue: warp w owns TMEM lanes 32 (w % 4) .. +31 (its hardware lane quarter) and one 128-column half
//               of every accumulator; a thread is ONE query row.  tcgen05.ld 32 columns at a time, then per compare
//               1 IMAD (FMA pipe) + 1.25 VIMNMX.U16x2 (ALU pipe):
//                 key16 = (hamming << 7) | (column & 127) = (16384 + column) - accumulator     (accumulator = 64 dot)
//               two columns share a register (low half: even columns, high half: odd columns), running top-2 per half
//               with the packed min/max of knn2_kernel, widened into the 32-bit keys every 128 columns.
// An item is 256 query rows of one matching against all of its train rows; each train tile is used by both query tiles
// (B traffic halves), and the two accumulators ping-pong so the epilogue of one overlaps the MMAs of the other.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "uz_knn2.cuh"

namespace uz {

constexpr int kMmaM = 128;                 // query rows per MMA / per TMEM accumulator (lanes)
constexpr int kMmaN = 256;                 // train rows per stage / accumulator columns
constexpr int kMmaItemRows = 2 * kMmaM;    // query rows per item (two accumulators)
constexpr int kE8RowBytes = 256;           // one int8 per descriptor bit
constexpr uint32_t kE8Set = 0x08u, kE8Clear = 0xF8u;   // +8 / -8: a product is +-64, so <q, t> arrives as 64 x (agreements - disagreements)
constexpr int kE8Dot = 64;                 // ... and the key unit 64 x 2 per bit of distance needs no multiply (uz_knn2_mmak.cuh)
constexpr int kE8GroupBytes = 8 * kE8RowBytes;
constexpr int kMmaThreads = 320;
constexpr int kMmaABytes = kMmaM * kE8RowBytes;     // 32 KB
constexpr int kMmaBBytes = kMmaN * kE8RowBytes;     // 64 KB
constexpr int kMmaSmemBytes = 2 * kMmaABytes + 2 * kMmaBBytes + 2 * kMmaItemRows * 8 + 256;

__host__ __device__ constexpr size_t e8_bytes(int n) { return (size_t)((n + 7) / 8) * kE8GroupBytes; }

// The kernel reads the batch's MatchTask table (uz_knn2.cuh): for a matching that runs here, q_desc / t_desc point at the
// E8 layouts of the "to" (OpenCV query) and "from" (train) cameras.
using MmaTask = MatchTask;
__device__ __forceinline__ const uint8_t* mma_q(const MmaTask* tk) { return reinterpret_cast<const uint8_t*>(tk->q_desc); }
__device__ __forceinline__ const uint8_t* mma_t(const MmaTask* tk) { return reinterpret_cast<const uint8_t*>(tk->t_desc); }

// Descriptor fields the host passes in (so that the probe can A/B them): see uz_knn2_mma_desc()
struct MmaDesc {
    uint32_t lbo16, sbo16;    // leading / stride byte offset, 16-byte units
    uint32_t idesc_base;      // instruction descriptor without N
};
__host__ __device__ inline MmaDesc uz_knn2_mma_desc() {
    MmaDesc d;
    d.lbo16 = 128 >> 4;                       // K-adjacent core matrices
    d.sbo16 = kE8GroupBytes >> 4;             // M/N-adjacent 8-row groups
    // c_format S32 (2) [4,6) | a_format INT8 (1) [7,10) | b_format INT8 (1) [10,13) | K-major A, B | M >> 4 at [24,29)
    d.idesc_base = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kMmaM >> 4) << 24);
    return d;
}

// ---- bits -> E8 layout (ingestion) ------------------------------------------------------------------------
// One thread per (row, 16-bit chunk): 16 int8 = one uint4 store.  raw: n rows x 8 words as given.
__global__ void __launch_bounds__(256) expand_e8_kernel(const uint32_t* __restrict__ raw, int n, uint8_t* __restrict__ e8) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 16) return;
    const int row = i >> 4, c = i & 15;
    const uint32_t w = raw[(size_t)row * 8 + (c >> 1)];
    const uint32_t bits = (c & 1) ? (w >> 16) : (w & 0xFFFFu);
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) v |= (((bits >> (4 * k + b)) & 1u) ? kE8Set : kE8Clear) << (8 * b);
        o[k] = v;
    }
    *reinterpret_cast<uint4*>(e8 + (size_t)(row >> 3) * kE8GroupBytes + c * 128 + (row & 7) * 16) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ---- tcgen05 helpers ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 columns of 32-bit accumulators: thread l of the warp receives lane (quarter * 32 + l)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// The registers of a tcgen05.ld are written asynchronously until wait::ld; the compiler only sees plain values, so every
// consumer is pinned behind the wait by passing the registers through an empty volatile asm.
__device__ __forceinline__ void tc_pin(uint32_t (&r)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) asm volatile("" : "+r"(r[j])::"memory");
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, const MmaDesc& d) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(d.lbo16 & 0x3FFFu) << 16) |
           ((uint64_t)(d.sbo16 & 0x3FFFu) << 32) | (1ull << 46);      // version 1 (sm_100), no swizzle, base offset 0
}

// mbarrier wait with a watchdog: a protocol bug must abort the launch, not hang the device
#ifndef UZ_MMA_WATCHDOG
#define UZ_MMA_WATCHDOG 1
#endif
// one lane of a converged warp, in the form ptxas recognises as "a single thread": tcgen05 instructions issued under it keep
// their descriptors in uniform registers and come out back to back.  Behind a `lane == 0` test every one of them is wrapped
// in a per-thread loop with R2UR moves (~70 clocks apiece next to busy epilogue warps).
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
#if UZ_MMA_WATCHDOG
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && spins > (1u << 26)) __trap();
    }
#else
    mbar_wait(bar, parity);
#endif
}

// Epilogue multipliers from the constant bank: with immediate powers of two ptxas turns the multiply-add into
// ALU-pipe shifts/LEAs, and the ALU pipe carries the packed min/max
__constant__ int32_t kMmaKeyMul[2] = {-1, -65536};        // accumulators hold 64 x dot (operands are +-8)

// two columns (ja even lane, jb odd lane) of one query row -> packed key16 pair
__device__ __forceinline__ uint32_t mma_pack2(uint32_t dot_a, uint32_t dot_b, const int ja /* a constant after unrolling */) {
    const uint32_t C = (16384u + (uint32_t)ja) | ((16384u + (uint32_t)ja + 1u) << 16);
    const uint32_t lo = mad_u32(dot_a, (uint32_t)kMmaKeyMul[0], C);
    return mad_u32(dot_b, (uint32_t)kMmaKeyMul[1], lo);
}

// 32 accumulator columns, all valid: columns 32 L .. 32 L + 31 of the thread's 128-column block
template <int L>
__device__ __forceinline__ void mma_chunk_full(const uint32_t (&d)[32], uint32_t& p1, uint32_t& p2) {
#pragma unroll
    for (int m = 0; m < 32; m += 4) {
        const uint32_t ka = mma_pack2(d[m], d[m + 1], 32 * L + m);
        const uint32_t kb = mma_pack2(d[m + 2], d[m + 3], 32 * L + m + 2);
        top2_update2_u16x2(p1, p2, ka, kb);
    }
}
// ragged end of the train rows: the first `valid` columns only, straight into the 32-bit keys
__device__ __forceinline__ void mma_chunk_masked(const uint32_t (&d)[32], int valid, uint32_t t_first, uint32_t& m1, uint32_t& m2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j < valid) {
            const uint32_t ham = (uint32_t)(256 * kE8Dot - (int32_t)d[j]) >> 7;       // (256 - dot) / 2, accumulator = 64 dot
            top2_update(m1, m2, (ham << 16) | (t_first + (uint32_t)j));
        }
    }
}

// UZ_MMA_PROF (probe builds only): per CTA, clocks the MMA issuer spent waiting for train tiles, query tiles and free
// accumulators, and its whole loop
#ifdef UZ_MMA_PROF
__device__ long long g_mma_prof[256][4];
#define UZ_PROF_T(x) const long long x = clock64()
#define UZ_PROF_ADD(k, a, b) prof[k] += (b) - (a)
#else
#define UZ_PROF_T(x)
#define UZ_PROF_ADD(k, a, b)
#endif

// items[k] = (task, first query row); CTA b takes items b, b + gridDim.x, ...
__global__ void __launch_bounds__(kMmaThreads, 1) knn2_mma_kernel(const MmaTask* __restrict__ tasks, const int2* __restrict__ items,
                                                                  int n_items, uint2* __restrict__ keys, MmaDesc dsc,
                                                                  int* __restrict__ pair_pending, unsigned int* __restrict__ progress) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;                                    // [2][32 KB]
    uint8_t* sB = smem + 2 * kMmaABytes;                   // [2][64 KB]
    uint2* xchg = reinterpret_cast<uint2*>(sB + 2 * kMmaBBytes);            // [2 parities][256 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + 2 * kMmaItemRows);
    uint64_t* a_full = bars;          // [2]
    uint64_t* a_empty = bars + 2;     // [2]
    uint64_t* b_full = bars + 4;      // [2]
    uint64_t* b_empty = bars + 6;     // [2]
    uint64_t* acc_full = bars + 8;    // [2]
    uint64_t* acc_empty = bars + 10;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1);
            mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {          // the whole TMEM: two 256-column accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== producer =====================
        if (lane == 0) {
            uint32_t uB = 0, uA[2] = {0, 0};
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const int2 item = items[it];
                const MmaTask* tk = tasks + item.x;
                const int nq = tk->nq, nt = tk->nt, q0 = item.y;
                const int nqt = (nq - q0 > kMmaM) ? 2 : 1;
                const int T = (nt + kMmaN - 1) / kMmaN;
                for (int t = 0; t < T; ++t) {
                    if (t == 0) {
                        mbar_wait_wd(&a_empty[0], (uA[0] & 1u) ^ 1u);
                        const uint32_t bytes = (uint32_t)e8_bytes(min(kMmaM, nq - q0));
                        mbar_expect_tx(&a_full[0], bytes);
                        bulk_g2s(sA, mma_q(tk) + (size_t)(q0 >> 3) * kE8GroupBytes, bytes, &a_full[0]);
                        uA[0]++;
                    }
                    {
                        const uint32_t slot = uB & 1u;
                        mbar_wait_wd(&b_empty[slot], ((uB >> 1) & 1u) ^ 1u);
                        const uint32_t bytes = (uint32_t)e8_bytes(min(kMmaN, nt - t * kMmaN));
                        mbar_expect_tx(&b_full[slot], bytes);
                        bulk_g2s(sB + slot * kMmaBBytes, mma_t(tk) + (size_t)t * kMmaBBytes, bytes, &b_full[slot]);
                        uB++;
                    }
                    if (t == 0 && nqt == 2) {
                        mbar_wait_wd(&a_empty[1], (uA[1] & 1u) ^ 1u);
                        const uint32_t bytes = (uint32_t)e8_bytes(min(kMmaM, nq - q0 - kMmaM));
                        mbar_expect_tx(&a_full[1], bytes);
                        bulk_g2s(sA + kMmaABytes, mma_q(tk) + (size_t)((q0 + kMmaM) >> 3) * kE8GroupBytes, bytes, &a_full[1]);
                        uA[1]++;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        uint32_t uB = 0, uA[2] = {0, 0}, uAcc[2] = {0, 0};
#ifdef UZ_MMA_PROF
        long long prof[4] = {0, 0, 0, 0};
        const long long prof_begin = clock64();
#endif
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const int2 item = items[it];
            const MmaTask* tk = tasks + item.x;
            const int nq = tk->nq, nt = tk->nt, q0 = item.y;
            const int nqt = (nq - q0 > kMmaM) ? 2 : 1;
            const int T = (nt + kMmaN - 1) / kMmaN;
            for (int t = 0; t < T; ++t) {
                const uint32_t slot = uB & 1u;
                UZ_PROF_T(p0);
                mbar_wait_wd(&b_full[slot], (uB >> 1) & 1u);
                UZ_PROF_T(p1);
                UZ_PROF_ADD(0, p0, p1);
                const int rows = min(kMmaN, nt - t * kMmaN);
                const uint32_t n_mma = (uint32_t)((rows + 15) & ~15);            // N: multiple of 16 at M = 128
                const uint32_t idesc = dsc.idesc_base | ((n_mma >> 3) << 17);
                for (int i = 0; i < nqt; ++i) {
                    UZ_PROF_T(p2);
                    if (t == 0) mbar_wait_wd(&a_full[i], uA[i] & 1u);
                    UZ_PROF_T(p3);
                    mbar_wait_wd(&acc_empty[i], (uAcc[i] & 1u) ^ 1u);
                    UZ_PROF_T(p4);
                    UZ_PROF_ADD(1, p2, p3);
                    UZ_PROF_ADD(2, p3, p4);
                    tc_fence_after();
                    if (tc_elect_one()) {
                        const uint32_t a_addr = smem_u32(sA + i * kMmaABytes), b_addr = smem_u32(sB + slot * kMmaBBytes);
#pragma unroll
                        for (int k = 0; k < kE8RowBytes / 32; ++k)
                            tc_mma_i8(tmem_base + (uint32_t)i * kMmaN, make_smem_desc(a_addr + k * 256, dsc),
                                      make_smem_desc(b_addr + k * 256, dsc), idesc, k > 0 ? 1u : 0u);
                        tc_commit(&acc_full[i]);
                        if (t == T - 1) tc_commit(&a_empty[i]);
                    }
                    __syncwarp();
                    uAcc[i]++;
                }
                if (tc_elect_one()) tc_commit(&b_empty[slot]);
                __syncwarp();
                uB++;
            }
            if (T > 0) for (int i = 0; i < nqt; ++i) uA[i]++;
        }
#ifdef UZ_MMA_PROF
        if (lane == 0 && blockIdx.x < 256) {
            prof[3] = clock64() - prof_begin;
            for (int k = 0; k < 4; ++k) g_mma_prof[blockIdx.x][k] = prof[k];
        }
#endif
    } else {
        // ===================== epilogue =====================
        const int ew = warp - 2;                   // 0..7
        const int quarter = warp & 3;              // the TMEM lanes this warp may touch: 32 * (warp % 4) ..
        const int half = ew >> 2;                  // which 128 columns of every accumulator
        const int row_in_tile = quarter * 32 + lane;
        uint32_t uAcc[2] = {0, 0};
        uint32_t item_parity = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, item_parity ^= 1u) {
            const int2 item = items[it];
            const MmaTask* tk = tasks + item.x;
            const int nq = tk->nq, nt = tk->nt, q0 = item.y;
            const int nqt = (nq - q0 > kMmaM) ? 2 : 1;
            const int T = (nt + kMmaN - 1) / kMmaN;
            uint32_t m1[2] = {kNoKey, kNoKey}, m2[2] = {kNoKey, kNoKey};
            for (int t = 0; t < T; ++t) {
                const int cvalid = min(kMmaN, nt - t * kMmaN) - half * 128;     // valid columns of this warp's half
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (i < nqt) {
                        mbar_wait_wd(&acc_full[i], uAcc[i] & 1u);
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(i * kMmaN + half * 128);
                        const uint32_t tbase = (uint32_t)(t * kMmaN + half * 128);       // global train row of column 0
                        uint32_t p1 = 0xFFFFFFFFu, p2 = 0xFFFFFFFFu;
                        uint32_t dA[32], dB[32];
                        if (cvalid >= 128) {
                            tc_ld32(taddr, dA);
                            tc_ld32(taddr + 32, dB);
                            tc_wait_ld(); tc_pin(dA); tc_pin(dB);
                            mma_chunk_full<0>(dA, p1, p2);
                            tc_ld32(taddr + 64, dA);
                            mma_chunk_full<1>(dB, p1, p2);
                            tc_wait_ld(); tc_pin(dA);
                            tc_ld32(taddr + 96, dB);
                            mma_chunk_full<2>(dA, p1, p2);
                            tc_wait_ld(); tc_pin(dB);
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&acc_empty[i]);
                            mma_chunk_full<3>(dB, p1, p2);
                            merge_block16(m1[i], m2[i], p1 & 0xFFFFu, p2 & 0xFFFFu, tbase);
                            merge_block16(m1[i], m2[i], p1 >> 16, p2 >> 16, tbase);
                        } else {
                            // ragged last tile: chunk by chunk, full chunks packed, the partial one masked
                            int nfull = 0;
                            if (cvalid > 0) {
                                tc_ld32(taddr, dA); tc_wait_ld(); tc_pin(dA);
                                if (cvalid >= 32) { mma_chunk_full<0>(dA, p1, p2); nfull++; } else mma_chunk_masked(dA, cvalid, tbase, m1[i], m2[i]);
                            }
                            if (cvalid > 32) {
                                tc_ld32(taddr + 32, dA); tc_wait_ld(); tc_pin(dA);
                                if (cvalid >= 64) { mma_chunk_full<1>(dA, p1, p2); nfull++; } else mma_chunk_masked(dA, cvalid - 32, tbase + 32, m1[i], m2[i]);
                            }
                            if (cvalid > 64) {
                                tc_ld32(taddr + 64, dA); tc_wait_ld(); tc_pin(dA);
                                if (cvalid >= 96) { mma_chunk_full<2>(dA, p1, p2); nfull++; } else mma_chunk_masked(dA, cvalid - 64, tbase + 64, m1[i], m2[i]);
                            }
                            if (cvalid > 96) {
                                tc_ld32(taddr + 96, dA); tc_wait_ld(); tc_pin(dA);
                                mma_chunk_masked(dA, cvalid - 96, tbase + 96, m1[i], m2[i]);
                            }
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&acc_empty[i]);
                            if (nfull > 0) {
                                merge_block16(m1[i], m2[i], p1 & 0xFFFFu, p2 & 0xFFFFu, tbase);
                                merge_block16(m1[i], m2[i], p1 >> 16, p2 >> 16, tbase);
                            }
                        }
                        uAcc[i]++;
                    }
                }
            }
            // fold the two column halves of every row (half 1 -> shared memory -> half 0) and publish the keys
            uint2* xc = xchg + item_parity * kMmaItemRows;
            if (half == 1) {
#pragma unroll
                for (int i = 0; i < 2; ++i) xc[i * kMmaM + row_in_tile] = make_uint2(m1[i], m2[i]);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (half == 0) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int q = q0 + i * kMmaM + row_in_tile;
                    if (q < nq) {
                        const uint2 o = xc[i * kMmaM + row_in_tile];
                        const uint32_t hi = max(m1[i], o.x);
                        const uint32_t a = min(m1[i], o.x);
                        const uint32_t b = min(hi, min(m2[i], o.y));
                        keys[(size_t)tk->key_off + q] = make_uint2(a, b);
                    }
                }
            }
            if (pair_pending != nullptr) {            // streaming hand-over, as in knn2_kernel (one count per item)
                if (half == 0) {
                    asm volatile("bar.sync 2, 128;" ::: "memory");   // the four half-0 warps: their key stores are done
                    if (row_in_tile == 0) {
                        __threadfence();
                        atomicSub(pair_pending + tk->pair, 1);
                        atomicAdd(progress, 1u);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace uz
