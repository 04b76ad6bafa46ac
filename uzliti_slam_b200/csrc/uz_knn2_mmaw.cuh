// uz_knn2_mmaw.cuh — the tensor-core match kernel for 512-bit rows (BRISK / FREAK: cv::BRISK is FeatureExtractionCore's
// default extractor, /root/reference/feature_extraction/src/feature_extraction_core.cpp:46-49,69-77).
//
// Same contract as knn2_wide_kernel (uz_knn2.cuh) and the same idea as knn2_mmak_kernel: bits -> +-8, the contraction over
// K = 512 delivers 64 <q, t>, and a constant K-slice (query side [32 x16, 1, 0 ...], train side [64 x16, 127 - column, 0 ...]:
// 16 * 32 * 64 = 32768) turns the accumulator into
//     K' = 64 <q, t> + 32768 + 127 - column = 65663 - ((hamming << 7) | column),      column = train row & 127
// The key needs 17 bits here (hamming 0..512), so the epilogue keeps the two LARGEST 32-bit values per thread and tile
// (2.5 VIMNMX per compare) - still no multiply, and a 512-bit compare buys twice the tensor-core time of a 256-bit one.
//
// Operand layout ("E8W"): a camera's rows as TWO planes, plane p = bits [256 p, 256 p + 256) in exactly the E8 layout of the
// 256-bit kernels (uz_knn2_mma.cuh), plane 1 behind plane 0 at e8_bytes(n).  A stage is then one contiguous bulk copy of
// 128 rows of ONE plane (32 KB), and the K loop runs over the two planes.
//
// Mapping (one CTA per SM, 320 threads, roles as in knn2_mmak_kernel).  An item is 256 query rows = two 128-row tiles, both
// planes resident (128 KB); train tiles are 128 rows (N = 128 per instruction), two 32 KB stages; per train tile and query
// tile 2 x 8 + 1 instructions into one of FOUR 128-column accumulators (query tile x parity of the train tile), so the
// epilogue of a train tile overlaps the instructions of the next.  A stage feeds both query tiles.
#pragma once
#include "uz_knn2_mmak.cuh"

namespace uz {

constexpr int kMmawN = 128;                               // train rows per tile / accumulator columns
constexpr int kMmawStageBytes = kMmawN * kE8RowBytes;     // 128 rows of one plane: 32 KB
constexpr int kMmawSmemBytes = 2 * 2 * kMmaABytes + 2 * kMmawStageBytes + 2 * kMmakTailABytes + 2 * kMmaItemRows * 8 + 256;
static_assert(kMmawSmemBytes <= 232448, "CTA exceeds the 227 KB of shared memory");
constexpr uint32_t kMmawTop = 65663u;                     // (hamming << 7) | column = kMmawTop - accumulator

__host__ __device__ constexpr size_t e8w_bytes(int n) { return 2 * e8_bytes(n); }

// running two LARGEST 32-bit values, two new ones per step (5 min/max)
__device__ __forceinline__ void top2max_update2(uint32_t& p1, uint32_t& p2, uint32_t ka, uint32_t kb) {
    const uint32_t hi = max(ka, kb), lo = min(ka, kb);
    const uint32_t t = min(p1, hi);
    p1 = max(p1, hi);
    p2 = max(max(p2, t), lo);
}
__device__ __forceinline__ uint32_t mmaw_key(uint32_t acc, uint32_t t_first) {     // accumulator -> (distance << 16) | trainIdx
    const uint32_t k17 = kMmawTop - acc;
    return ((k17 >> 7) << 16) | (t_first + (k17 & 127u));
}

// items[k] = (task, first query row); CTA b takes items b, b + gridDim.x, ...
__global__ void __launch_bounds__(kMmaThreads, 1) knn2_mmaw_kernel(const MmaTask* __restrict__ tasks, const int2* __restrict__ items,
                                                                   int n_items, uint2* __restrict__ keys, MmaDesc dsc) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;                                                     // [2 tiles][2 planes][32 KB]
    uint8_t* sB = smem + 4 * kMmaABytes;                                    // [2 stages][32 KB]
    uint8_t* sTailA = sB + 2 * kMmawStageBytes;                             // [128 rows x 32 B]
    uint8_t* sTailB = sTailA + kMmakTailABytes;                             // [128 rows x 32 B]
    uint2* xchg = reinterpret_cast<uint2*>(sTailB + kMmakTailABytes);       // [2 parities][256 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + 2 * kMmaItemRows);
    uint64_t* a_full = bars;          // [2]
    uint64_t* a_empty = bars + 2;     // [2]
    uint64_t* b_full = bars + 4;      // [2]
    uint64_t* b_empty = bars + 6;     // [2]
    uint64_t* acc_full = bars + 8;    // [4]  query tile * 2 + parity of the train tile
    uint64_t* acc_empty = bars + 12;  // [4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1);
            mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 4; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the constant K-slice: 16 x (32 * 64) = 32768, + 1 * (127 - train row of the tile); compact canonical layout (8-row groups of 256 B)
    for (int i = tid; i < 2 * kMmaM * 2; i += kMmaThreads) {
        const bool isB = i >= kMmaM * 2;
        const int r = (isB ? i - kMmaM * 2 : i) >> 1, kc = i & 1;
        uint4 v;
        if (kc == 0) v = isB ? make_uint4(0x40404040u, 0x40404040u, 0x40404040u, 0x40404040u) : make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
        else v = make_uint4(isB ? (uint32_t)(127 - r) : 1u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>((isB ? sTailB : sTailA) + (r >> 3) * 256 + kc * 128 + (r & 7) * 16) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {          // the whole TMEM: four 128-column accumulators
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
                const int T = (nt + kMmawN - 1) / kMmawN;
                const size_t q_plane = e8_bytes(nq), t_plane = e8_bytes(nt);
                auto load_a = [&](int i) {
                    mbar_wait_wd(&a_empty[i], (uA[i] & 1u) ^ 1u);
                    const int first = q0 + i * kMmaM;
                    const uint32_t bytes = (uint32_t)e8_bytes(min(kMmaM, nq - first));
                    mbar_expect_tx(&a_full[i], 2 * bytes);
                    for (int p = 0; p < 2; ++p)
                        bulk_g2s(sA + (i * 2 + p) * kMmaABytes, mma_q(tk) + p * q_plane + (size_t)(first >> 3) * kE8GroupBytes, bytes, &a_full[i]);
                    uA[i]++;
                };
                for (int t = 0; t < T; ++t) {
                    if (t == 0) load_a(0);
                    for (int p = 0; p < 2; ++p) {
                        const uint32_t slot = uB & 1u;
                        mbar_wait_wd(&b_empty[slot], ((uB >> 1) & 1u) ^ 1u);
                        const uint32_t bytes = (uint32_t)e8_bytes(min(kMmawN, nt - t * kMmawN));
                        mbar_expect_tx(&b_full[slot], bytes);
                        bulk_g2s(sB + slot * kMmawStageBytes, mma_t(tk) + p * t_plane + (size_t)t * kMmawStageBytes, bytes, &b_full[slot]);
                        uB++;
                        if (t == 0 && p == 0 && nqt == 2) load_a(1);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        uint32_t uB = 0, uA[2] = {0, 0}, uAcc[4] = {0, 0, 0, 0};
        MmaDesc tail_dsc = dsc;
        tail_dsc.lbo16 = 128 >> 4; tail_dsc.sbo16 = 256 >> 4;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const int2 item = items[it];
            const MmaTask* tk = tasks + item.x;
            const int nq = tk->nq, nt = tk->nt, q0 = item.y;
            const int nqt = (nq - q0 > kMmaM) ? 2 : 1;
            const int T = (nt + kMmawN - 1) / kMmawN;
            for (int t = 0; t < T; ++t) {
                const int rows = min(kMmawN, nt - t * kMmawN);
                const uint32_t n_mma = (uint32_t)((rows + 15) & ~15);            // N: multiple of 16 at M = 128
                const uint32_t idesc = dsc.idesc_base | ((n_mma >> 3) << 17);
                for (int p = 0; p < 2; ++p) {
                    const uint32_t slot = uB & 1u;
                    mbar_wait_wd(&b_full[slot], (uB >> 1) & 1u);
                    for (int i = 0; i < nqt; ++i) {
                        const int ai = i * 2 + (t & 1);
                        if (p == 0) {
                            if (t == 0) mbar_wait_wd(&a_full[i], uA[i] & 1u);
                            mbar_wait_wd(&acc_empty[ai], (uAcc[ai] & 1u) ^ 1u);
                        }
                        tc_fence_after();
                        if (tc_elect_one()) {
                            const uint32_t a_addr = smem_u32(sA + (i * 2 + p) * kMmaABytes), b_addr = smem_u32(sB + slot * kMmawStageBytes);
                            const uint32_t d = tmem_base + (uint32_t)ai * kMmawN;
#pragma unroll
                            for (int k = 0; k < kE8RowBytes / 32; ++k)
                                tc_mma_i8(d, make_smem_desc(a_addr + k * 256, dsc), make_smem_desc(b_addr + k * 256, dsc), idesc,
                                          (p > 0 || k > 0) ? 1u : 0u);
                            if (p == 1) {
                                tc_mma_i8(d, make_smem_desc(smem_u32(sTailA), tail_dsc), make_smem_desc(smem_u32(sTailB), tail_dsc), idesc, 1u);
                                tc_commit(&acc_full[ai]);
                                if (t == T - 1) tc_commit(&a_empty[i]);
                            }
                        }
                        __syncwarp();
                        if (p == 1) uAcc[ai]++;
                    }
                    if (tc_elect_one()) tc_commit(&b_empty[slot]);
                    __syncwarp();
                    uB++;
                }
            }
            if (T > 0) for (int i = 0; i < nqt; ++i) uA[i]++;
        }
    } else {
        // ===================== epilogue =====================
        const int ew = warp - 2;                   // 0..7
        const int quarter = warp & 3;              // the TMEM lanes this warp may touch: 32 * (warp % 4) ..
        const int half = ew >> 2;                  // which 64 columns of every accumulator
        const int row_in_tile = quarter * 32 + lane;
        uint32_t uAcc[4] = {0, 0, 0, 0};
        uint32_t item_parity = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, item_parity ^= 1u) {
            const int2 item = items[it];
            const MmaTask* tk = tasks + item.x;
            const int nq = tk->nq, nt = tk->nt, q0 = item.y;
            const int nqt = (nq - q0 > kMmaM) ? 2 : 1;
            const int T = (nt + kMmawN - 1) / kMmawN;
            uint32_t m1[2] = {kNoKey, kNoKey}, m2[2] = {kNoKey, kNoKey};
            for (int t = 0; t < T; ++t) {
                const int cvalid = min(kMmawN, nt - t * kMmawN) - half * 64;     // valid columns of this warp's half
                const uint32_t tbase = (uint32_t)(t * kMmawN);                     // train row of column 0 of the tile
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (i < nqt) {
                        const int ai = i * 2 + (t & 1);
                        // (uAcc is indexed with a value the unrolled code knows only up to the parity of t)
                        uint32_t& ua = (t & 1) ? uAcc[i * 2 + 1] : uAcc[i * 2];
                        mbar_wait_wd(&acc_full[ai], ua & 1u);
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ai * kMmawN + half * 64);
                        uint32_t dA[32], dB[32];
                        if (cvalid > 0) tc_ld32(taddr, dA);
                        if (cvalid > 32) tc_ld32(taddr + 32, dB);
                        tc_wait_ld(); tc_pin(dA); tc_pin(dB);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[ai]);
                        if (cvalid >= 64) {
                            uint32_t p1 = 0u, p2 = 0u;
#pragma unroll
                            for (int m = 0; m < 32; m += 2) top2max_update2(p1, p2, dA[m], dA[m + 1]);
#pragma unroll
                            for (int m = 0; m < 32; m += 2) top2max_update2(p1, p2, dB[m], dB[m + 1]);
                            top2_update(m1[i], m2[i], mmaw_key(p1, tbase));
                            top2_update(m1[i], m2[i], mmaw_key(p2, tbase));
                        } else {
                            // ragged last tile: column by column
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < cvalid) top2_update(m1[i], m2[i], mmaw_key(dA[j], tbase));
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j + 32 < cvalid) top2_update(m1[i], m2[i], mmaw_key(dB[j], tbase));
                        }
                        ua++;
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
