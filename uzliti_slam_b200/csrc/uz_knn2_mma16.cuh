// uz_knn2_mma16.cuh — knn2_mma_kernel with SIXTEEN epilogue warps.
//
// Why.  The one-CTA kernel is bound by its epilogue, not by the tensor pipe: ~830 warp-instructions per scheduler and
// accumulator at an issue rate of 0.54 with two epilogue warps per scheduler = 1530 clocks per 128 x 256 accumulator, against
// 1024 clocks for the eight instructions that fill it (profiles/mma_experiments_r02.txt).  Four warps per scheduler, each
// owning a 64-column quarter of the accumulator: both of its TMEM loads are in flight at once, the accumulator goes back to
// the issuer as soon as they have landed, and the dependent packed min/max chains of four warps interleave.
// 18 warps x 32 lanes at <= kMma16Regs registers leave room for the copy / layout CTAs of the next chunk beside the match CTA
// (uz_estimate_edges_host).  Same operands, same arithmetic, bit-identical keys.
#pragma once
#include "uz_knn2_mma.cuh"

namespace uz {

constexpr int kMma16Threads = 18 * 32;
constexpr int kMma16Regs = 96;        // five warps per scheduler x 96 x 32 = 15 360 of its 16 384 registers
constexpr int kMma16SmemBytes = 2 * kMmaABytes + 2 * kMmaBBytes + 2 * 3 * kMmaItemRows * 8 + 256;
static_assert(kMma16SmemBytes <= 232448, "CTA exceeds the 227 KB of shared memory");

// items[k] = (task, first query row); CTA b takes items b, b + gridDim.x, ...
__global__ void __maxnreg__(kMma16Regs) knn2_mma16_kernel(const MmaTask* __restrict__ tasks, const int2* __restrict__ items,
                                                                  int n_items, uint2* __restrict__ keys, MmaDesc dsc,
                                                                  int* __restrict__ pair_pending, unsigned int* __restrict__ progress) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;                                    // [2][32 KB]
    uint8_t* sB = smem + 2 * kMmaABytes;                   // [2][64 KB]
    uint2* xchg = reinterpret_cast<uint2*>(sB + 2 * kMmaBBytes);            // [2 parities][3 column quarters][256 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + 2 * 3 * kMmaItemRows);
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
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 16);
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
                    if (lane == 0) {
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
                if (lane == 0) tc_commit(&b_empty[slot]);
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
        // ===================== epilogue: 16 warps, four per scheduler =====================
        // warp w owns TMEM lanes 32 (w % 4) .. +31 and one 64-column quarter of every accumulator: both 32-column loads are
        // issued at once, the accumulator is released as soon as they have landed (the issuer never waits for arithmetic),
        // and four warps per scheduler hide each other's dependent min/max chains.
        const int ew = warp - 2;                   // 0..15
        const int quarter = warp & 3;              // hardware lane quarter of this warp
        const int colq = ew >> 2;                  // 64-column quarter of the accumulator
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
                const int cvalid = min(kMmaN, nt - t * kMmaN) - colq * 64;      // valid columns of this warp's quarter
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (i < nqt) {
                        mbar_wait_wd(&acc_full[i], uAcc[i] & 1u);
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(i * kMmaN + colq * 64);
                        const uint32_t kbase = (uint32_t)(t * kMmaN + (colq >> 1) * 128);   // train row of column 0 of the 128-column key block
                        const uint32_t tbase = (uint32_t)(t * kMmaN + colq * 64);           // train row of this warp's first column
                        uint32_t p1 = 0xFFFFFFFFu, p2 = 0xFFFFFFFFu;
                        uint32_t dA[32], dB[32];
                        if (cvalid > 0) { tc_ld32(taddr, dA); if (cvalid > 32) tc_ld32(taddr + 32, dB); tc_wait_ld(); tc_pin(dA); tc_pin(dB); }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[i]);
                        if (cvalid >= 64) {
                            if (colq & 1) { mma_chunk_full<2>(dA, p1, p2); mma_chunk_full<3>(dB, p1, p2); }
                            else { mma_chunk_full<0>(dA, p1, p2); mma_chunk_full<1>(dB, p1, p2); }
                            merge_block16(m1[i], m2[i], p1 & 0xFFFFu, p2 & 0xFFFFu, kbase);
                            merge_block16(m1[i], m2[i], p1 >> 16, p2 >> 16, kbase);
                        } else if (cvalid > 0) {
                            // ragged last tile: the full chunk packed, the partial one masked
                            if (cvalid >= 32) {
                                if (colq & 1) mma_chunk_full<2>(dA, p1, p2); else mma_chunk_full<0>(dA, p1, p2);
                                merge_block16(m1[i], m2[i], p1 & 0xFFFFu, p2 & 0xFFFFu, kbase);
                                merge_block16(m1[i], m2[i], p1 >> 16, p2 >> 16, kbase);
                                if (cvalid > 32) mma_chunk_masked(dB, cvalid - 32, tbase + 32, m1[i], m2[i]);
                            } else {
                                mma_chunk_masked(dA, cvalid, tbase, m1[i], m2[i]);
                            }
                        }
                        uAcc[i]++;
                    }
                }
            }
            // fold the four column quarters of every row (quarters 1..3 -> shared memory -> quarter 0) and publish the keys
            uint2* xc = xchg + item_parity * 3 * kMmaItemRows;
            if (colq > 0) {
#pragma unroll
                for (int i = 0; i < 2; ++i) xc[(colq - 1) * kMmaItemRows + i * kMmaM + row_in_tile] = make_uint2(m1[i], m2[i]);
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (colq == 0) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int q = q0 + i * kMmaM + row_in_tile;
                    if (q < nq) {
                        uint32_t a = m1[i], b = m2[i];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const uint2 o = xc[c * kMmaItemRows + i * kMmaM + row_in_tile];
                            const uint32_t hi = max(a, o.x);
                            a = min(a, o.x);
                            b = min(hi, min(b, o.y));
                        }
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
