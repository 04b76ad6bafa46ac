// uz_knn2_mmak.cuh — the tensor-core match kernel with the KEYS formed by the tensor core.
//
// knn2_mma_kernel's epilogue spends, per compare, one IMAD (half-rate pipe) to turn a dot product into a packed 16-bit key and
// 1.25 packed min/max to keep the two best: ~900 clocks of pipe time per 128 x 256 accumulator against 1024 clocks for the
// eight instructions that fill it, and the kernel ran at ~1530 clocks per accumulator (profiles/mma_experiments_r02.txt).
// Here the operands are +-8 (a product is +-64) and every accumulator gets a NINTH K-slice from two constant shared-memory
// blocks: query side [32 x16, 1, 0 ...], train side [32 x16, 127 - (row & 127), 0 ...], so that
//     accumulator = 64 <q, t> + 16384 + 127 - (column & 127) = 32895 - key16,   key16 = (hamming << 7) | (column & 127)
// - the packed key of uz_knn2.cuh, complemented, in 16 bits.  tcgen05.ld ... .pack::16b then returns TWO columns per register
// (scripts/tmem_pack_probe.cu) and the epilogue is nothing but packed max: 1.25 VIMNMX.U16x2 per compare, no multiply, half
// the TMEM load instructions, and the accumulator goes back to the issuer as soon as its two loads have landed.  Costs one
// more tcgen05.mma per eight.  Same E8 layout in HBM, same items, bit-identical keys.
#pragma once
#include "uz_knn2_mma.cuh"

namespace uz {

constexpr int kMmakTailABytes = kMmaM * 32;        // 4 KB
constexpr int kMmakTailBBytes = kMmaN * 32;        // 8 KB
constexpr int kMmakSmemBytes = kMmaSmemBytes + kMmakTailABytes + kMmakTailBBytes;
static_assert(kMmakSmemBytes <= 232448, "CTA exceeds the 227 KB of shared memory");
constexpr uint32_t kMmakTop = 32895u;              // key16 = kMmakTop - accumulator

// 32 lanes x 64 columns, low halves of adjacent columns packed: register j = column 2j | column 2j + 1 << 16
__device__ __forceinline__ void tc_ld64p(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// running two LARGEST per 16-bit lane (p1 >= p2), two registers = four columns per step
__device__ __forceinline__ void top2max_update2_u16x2(uint32_t& p1, uint32_t& p2, uint32_t ka, uint32_t kb) {
    const uint32_t hi = max_u16x2(ka, kb), lo = min_u16x2(ka, kb);
    const uint32_t t = min_u16x2(p1, hi);
    p1 = max_u16x2(p1, hi);
    p2 = max_u16x2(max_u16x2(p2, t), lo);
}
__device__ __forceinline__ void mmak_group(const uint32_t (&d)[32], uint32_t& p1, uint32_t& p2) {
#pragma unroll
    for (int m = 0; m < 32; m += 2) top2max_update2_u16x2(p1, p2, d[m], d[m + 1]);
}
// the four winners of a 128-column block (even and odd columns, best and second) -> full keys
__device__ __forceinline__ void mmak_merge(uint32_t& m1, uint32_t& m2, uint32_t& p1, uint32_t& p2, uint32_t tbase) {
    merge_block16(m1, m2, kMmakTop - (p1 & 0xFFFFu), kMmakTop - (p2 & 0xFFFFu), tbase);
    merge_block16(m1, m2, kMmakTop - (p1 >> 16), kMmakTop - (p2 >> 16), tbase);
    p1 = 0u; p2 = 0u;
}
// unpacked columns (ragged end of the train rows): the first `valid` of 32, first column = train row t_first
__device__ __forceinline__ void mmak_masked(const uint32_t (&d)[32], int valid, uint32_t t_first, uint32_t& m1, uint32_t& m2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j < valid) {
            const uint32_t ham = (kMmakTop - d[j]) >> 7;
            top2_update(m1, m2, (ham << 16) | (t_first + (uint32_t)j));
        }
    }
}

// items[k] = (task, first query row); CTA b takes items b, b + gridDim.x, ...
__global__ void __launch_bounds__(kMmaThreads, 1) knn2_mmak_kernel(const MmaTask* __restrict__ tasks, const int2* __restrict__ items,
                                                                  int n_items, uint2* __restrict__ keys, MmaDesc dsc,
                                                                  int* __restrict__ pair_pending, unsigned int* __restrict__ progress) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;                                    // [2][32 KB]
    uint8_t* sB = smem + 2 * kMmaABytes;                   // [2][64 KB]
    uint8_t* sTailA = sB + 2 * kMmaBBytes;                                  // [128 rows x 32 B]  constant 9th K-slice of the query operand
    uint8_t* sTailB = sTailA + kMmakTailABytes;                             // [256 rows x 32 B]  ... of the train operand: carries 127 - (row & 127)
    uint2* xchg = reinterpret_cast<uint2*>(sTailB + kMmakTailBBytes);       // [2 parities][256 rows]
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
    // the constant K-slice: 16 x (32 * 32) = 16384, + 1 * (127 - (train row & 127)); compact canonical layout (8-row groups of 256 B)
    for (int i = tid; i < (kMmaM + kMmaN) * 2; i += kMmaThreads) {
        const bool isB = i >= kMmaM * 2;
        const int r = (isB ? i - kMmaM * 2 : i) >> 1, kc = i & 1;          // row, 16-byte K chunk
        uint4 v;
        if (kc == 0) v = make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
        else v = make_uint4(isB ? (uint32_t)(127 - (r & 127)) : 1u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>((isB ? sTailB : sTailA) + (r >> 3) * 256 + kc * 128 + (r & 7) * 16) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> visible to the tensor core
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
        MmaDesc tail_dsc = dsc;
        tail_dsc.lbo16 = 128 >> 4; tail_dsc.sbo16 = 256 >> 4;
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
                        tc_mma_i8(tmem_base + (uint32_t)i * kMmaN, make_smem_desc(smem_u32(sTailA), tail_dsc),
                                  make_smem_desc(smem_u32(sTailB), tail_dsc), idesc, 1u);
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
                        // The accumulator IS the key: 64 dot + 16384 + 127 - (column & 127) = 32895 - ((hamming << 7) | (column & 127)),
                        // 16 bits; a packed TMEM load returns two columns per register and the LARGEST values win.
                        uint32_t p1 = 0u, p2 = 0u;
                        uint32_t dA[32], dB[32];
                        if (cvalid >= 128) {
                            tc_ld64p(taddr, dA);
                            tc_ld64p(taddr + 64, dB);
                            tc_wait_ld(); tc_pin(dA); tc_pin(dB);
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&acc_empty[i]);
                            mmak_group(dA, p1, p2);
                            mmak_group(dB, p1, p2);
                            mmak_merge(m1[i], m2[i], p1, p2, tbase);
                        } else {
                            // ragged last tile: whole 64-column groups packed, the rest column by column
                            const int n64 = cvalid > 0 ? cvalid >> 6 : 0;
                            const int rem = cvalid > 0 ? cvalid & 63 : 0;
                            if (n64 > 0) { tc_ld64p(taddr, dA); tc_wait_ld(); tc_pin(dA); mmak_group(dA, p1, p2); mmak_merge(m1[i], m2[i], p1, p2, tbase); }
                            if (rem > 0) {
                                tc_ld32(taddr + n64 * 64, dA); tc_wait_ld(); tc_pin(dA);
                                mmak_masked(dA, min(rem, 32), tbase + n64 * 64, m1[i], m2[i]);
                            }
                            if (rem > 32) {
                                tc_ld32(taddr + n64 * 64 + 32, dA); tc_wait_ld(); tc_pin(dA);
                                mmak_masked(dA, rem - 32, tbase + n64 * 64 + 32, m1[i], m2[i]);
                            }
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&acc_empty[i]);
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
