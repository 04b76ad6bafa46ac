// uz_knn2_mma2.cuh — the tensor-core match kernel (uz_knn2_mma.cuh) on CTA PAIRS: tcgen05.mma.cta_group::2, 256 x 256 x 32.
//
// Why.  In the one-CTA kernel the MMA issuer waits 8 % of its time for train tiles (two 64 KB stages are all that fits beside
// the 64 KB of query tiles) and its instructions take ~148 clocks instead of 128 while the stages are being refilled
// (scripts/mma_probe.cu with UZ_MMA_PROF, scripts/mma2_probe.cu).  A CTA pair splits every 256-row train tile between the two
// SMs: each CTA stages only ITS 128 train rows (32 KB), so four stages fit, the shared-memory operand traffic per
// instruction drops from 12 KB to 8 KB per SM, and the L2 -> SM traffic per compare halves (a train tile now serves 512
// query rows).  scripts/mma2_probe.cu: four 32 KB stages keep a pair at 128.4 clocks per instruction even when every tile
// comes from DRAM; two 64 KB stages hold one CTA at 160.
//
// Same contract, same E8 operand layout, same epilogue arithmetic as knn2_mma_kernel; results are bit-identical.
//
// Mapping.  Cluster of two CTAs (one TPC), 320 threads each.  An item is 512 query rows of one matching = two 256-row
// blocks; block i is the M = 256 operand of one instruction stream: CTA r holds rows [256 i + 128 r, +128) of it in its own
// shared memory and receives exactly those rows of the accumulator in its own TMEM (lanes 0..127, 256 columns), two
// accumulators per CTA as before.
//   warp 0      producer (both CTAs): its two 128-row query tiles per item, its 128-row half of every train tile, 4 stages
//   warp 1      CTA 0: MMA issuer - waits for both halves of a stage (its own mbarrier + a "peer" mbarrier), issues
//                      8 x tcgen05.mma.cta_group::2 per block and train tile, commits with multicast to both CTAs
//               CTA 1: relay - forwards "my half has landed" to the leader's peer barriers (remote mbarrier arrive); a 1-D
//                      bulk copy can only signal a barrier next to its destination
//   warps 2..9  epilogue, unchanged: a thread is one query row and one 128-column half of the accumulator; CTA 1's warps
//               release an accumulator with a remote arrive on the leader's barrier
// The last train tile is always issued with N = 256 (at cta_group::2 a shorter N would re-split the columns between the
// CTAs); the epilogue masks the columns past the last train row, as it does in the one-CTA kernel.
#pragma once
#include "uz_knn2_mma.cuh"

namespace uz {

constexpr int kMma2ItemRows = 4 * kMmaM;            // 512 query rows per item and CTA pair
constexpr int kMma2BBytes = kMmaM * kE8RowBytes;    // this CTA's half of a train tile: 128 rows = 32 KB
// ASETS query-tile sets (2: the next item's query tiles load while this item computes), NS train stages
__host__ __device__ constexpr int mma2_xchg_parities(int asets) { return asets == 1 ? 2 : 1; }
__host__ __device__ constexpr int mma2_smem_bytes(int asets, int ns) {
    return asets * 2 * kMmaABytes + ns * kMma2BBytes + mma2_xchg_parities(asets) * 2 * kMmaM * 8 + 256;
}
static_assert(mma2_smem_bytes(2, 3) <= 232448 && mma2_smem_bytes(1, 4) <= 232448, "CTA exceeds the 227 KB of shared memory");

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(rank));
    return ra;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
#ifdef UZ_MMA2_STRONG_SCOPE
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#else
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");      // as CUTLASS' ClusterBarrier::arrive(cta_id)
#endif
}
// wait on a barrier that CTAs of the cluster arrive on
__device__ __forceinline__ void mbar_wait_cluster_wd(uint64_t* bar, uint32_t parity) {
#ifndef UZ_MMA2_STRONG_SCOPE
    mbar_wait_wd(bar, parity);
    return;
#endif
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tc2_commit(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tc2_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

#ifdef UZ_MMA_PROF
__device__ long long g_mma_prof2[256][16];      // [cluster][0..3 train tile t & 3 | 4..5 query block | 8..15 accumulator (t & 3) * 2 + block]
#define UZ_PROF2_ADD(k, a, b) prof2[k] += (b) - (a)
#else
#define UZ_PROF2_ADD(k, a, b)
#endif

// rows of [first, first + 128) that exist in a camera of n rows
__device__ __forceinline__ int rows_in(int n, int first) { return max(0, min(kMmaM, n - first)); }

// items[k] = (task, first query row, a multiple of 512); cluster c takes items c, c + clusters, ...
template <int ASETS, int NS>
__global__ void __launch_bounds__(kMmaThreads, 1) knn2_mma2_kernel(const MmaTask* __restrict__ tasks, const int2* __restrict__ items,
                                                                   int n_items, uint2* __restrict__ keys, MmaDesc dsc) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int NA = 2 * ASETS;                                           // query-tile buffers: [set][block]
    constexpr int XP = mma2_xchg_parities(ASETS);
    uint8_t* sA = smem;                                                     // [NA][32 KB]  this CTA's 128 rows of a block
    uint8_t* sB = smem + NA * kMmaABytes;                                   // [NS][32 KB]  this CTA's half of a train tile
    uint2* xchg = reinterpret_cast<uint2*>(sB + NS * kMma2BBytes);          // [XP][256 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + XP * 2 * kMmaM);
    uint64_t* a_full = bars;                          // [NA]  own query tile landed
    uint64_t* a_peer = bars + NA;                     // [NA]  (leader) the peer's query tile landed
    uint64_t* a_empty = bars + 2 * NA;                // [NA]  multicast commit
    uint64_t* b_full = bars + 3 * NA;                 // [NS]
    uint64_t* b_peer = b_full + NS;                   // [NS]
    uint64_t* b_empty = b_full + 2 * NS;              // [NS]  multicast commit
    uint64_t* acc_full = b_full + 3 * NS;             // [2]  multicast commit
    uint64_t* acc_empty = acc_full + 2;               // [2]  (leader) 8 local + 8 remote epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    static_assert((3 * NA + 3 * NS + 4) * 8 + 4 <= 256, "barrier block");

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (tid == 0) {
        for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_peer[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 16); }
        for (int s = 0; s < NS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_peer[s], 1); mbar_init(&b_empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {          // the whole TMEM of both SMs: two 256-column accumulators per CTA
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();       // barriers of both CTAs are initialised before anyone arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== producer (both CTAs) =====================
        if (lane == 0) {
            uint32_t uB = 0, uA[NA], seq = 0;
            for (int i = 0; i < NA; ++i) uA[i] = 0;
            for (int it = cluster; it < n_items; it += n_clusters, ++seq) {
                const int2 item = items[it];
                const MmaTask* tk = tasks + item.x;
                const int nq = tk->nq, nt = tk->nt, q0 = item.y;
                const int nblk = (nq - q0 > 2 * kMmaM) ? 2 : 1;
                const int T = (nt + kMmaN - 1) / kMmaN;
                if (T > 0) {
                    // both query tiles up front (with two sets they were free long ago: this item's tiles load while the
                    // previous item computes)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        if (i >= nblk) break;
                        const uint32_t ab = (seq % ASETS) * 2 + i;
                        mbar_wait_wd(&a_empty[ab], (uA[ab] & 1u) ^ 1u);
                        const int first = q0 + i * 2 * kMmaM + (int)rank * kMmaM;
                        const int rows = rows_in(nq, first);
                        if (rows > 0) {
                            const uint32_t bytes = (uint32_t)e8_bytes(rows);
                            mbar_expect_tx(&a_full[ab], bytes);
                            bulk_g2s(sA + ab * kMmaABytes, mma_q(tk) + (size_t)(first >> 3) * kE8GroupBytes, bytes, &a_full[ab]);
                        } else {
                            mbar_arrive(&a_full[ab]);       // nothing of this block in this CTA: the buffer keeps stale bytes, rows unused
                        }
                        uA[ab]++;
                    }
                }
                for (int t = 0; t < T; ++t) {
                    const uint32_t slot = uB % NS;
                    mbar_wait_wd(&b_empty[slot], ((uB / NS) & 1u) ^ 1u);
                    const int first = t * kMmaN + (int)rank * kMmaM;
                    const int rows = rows_in(nt, first);
                    if (rows > 0) {
                        const uint32_t bytes = (uint32_t)e8_bytes(rows);
                        mbar_expect_tx(&b_full[slot], bytes);
                        bulk_g2s(sB + slot * kMma2BBytes, mma_t(tk) + (size_t)(first >> 3) * kE8GroupBytes, bytes, &b_full[slot]);
                    } else {
                        mbar_arrive(&b_full[slot]);
                    }
                    uB++;
                }
            }
        }
    } else if (warp == 1 && rank == 1) {
        // ===================== relay (CTA 1) =====================
        if (lane == 0) {
            uint32_t uB = 0, uA[NA], seq = 0;
            uint32_t ra_peer[NA], rb_peer[NS];
            for (int i = 0; i < NA; ++i) { uA[i] = 0; ra_peer[i] = mapa_u32(&a_peer[i], 0); }
            for (int s = 0; s < NS; ++s) rb_peer[s] = mapa_u32(&b_peer[s], 0);
            for (int it = cluster; it < n_items; it += n_clusters, ++seq) {
                const int2 item = items[it];
                const MmaTask* tk = tasks + item.x;
                const int nq = tk->nq, nt = tk->nt, q0 = item.y;
                const int nblk = (nq - q0 > 2 * kMmaM) ? 2 : 1;
                const int T = (nt + kMmaN - 1) / kMmaN;
                if (T > 0) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        if (i >= nblk) break;
                        const uint32_t ab = (seq % ASETS) * 2 + i;
                        mbar_wait_wd(&a_full[ab], uA[ab] & 1u);
                        mbar_arrive_remote(ra_peer[ab]);
                        uA[ab]++;
                    }
                }
                for (int t = 0; t < T; ++t) {
                    const uint32_t slot = uB % NS;
                    mbar_wait_wd(&b_full[slot], (uB / NS) & 1u);
                    mbar_arrive_remote(rb_peer[slot]);
                    uB++;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (CTA 0) =====================
        uint32_t uB = 0, uA[NA], uAcc[2] = {0, 0}, seq = 0;
        for (int i = 0; i < NA; ++i) uA[i] = 0;
        // M = 256 (both CTAs), N = 256
        const uint32_t idesc = (dsc.idesc_base & ~(0x1Fu << 24)) | ((uint32_t)((2 * kMmaM) >> 4) << 24) | ((uint32_t)(kMmaN >> 3) << 17);
#ifdef UZ_MMA_PROF
        long long prof[4] = {0, 0, 0, 0};
        long long prof2[16];
        for (int k = 0; k < 16; ++k) prof2[k] = 0;
        const long long prof_begin = clock64();
#endif
        for (int it = cluster; it < n_items; it += n_clusters, ++seq) {
            const int2 item = items[it];
            const MmaTask* tk = tasks + item.x;
            const int nq = tk->nq, nt = tk->nt, q0 = item.y;
            const int nblk = (nq - q0 > 2 * kMmaM) ? 2 : 1;
            const int T = (nt + kMmaN - 1) / kMmaN;
            for (int t = 0; t < T; ++t) {
                const uint32_t slot = uB % NS;
                const uint32_t bpar = (uB / NS) & 1u;
                UZ_PROF_T(p0);
                mbar_wait_wd(&b_full[slot], bpar);
                mbar_wait_cluster_wd(&b_peer[slot], bpar);
                UZ_PROF_T(p1);
                UZ_PROF_ADD(0, p0, p1);
                UZ_PROF2_ADD(t & 3, p0, p1);
                for (int i = 0; i < nblk; ++i) {
                    const uint32_t ab = (seq % ASETS) * 2 + i;
                    UZ_PROF_T(p2);
                    if (t == 0) {
                        mbar_wait_wd(&a_full[ab], uA[ab] & 1u);
                        mbar_wait_cluster_wd(&a_peer[ab], uA[ab] & 1u);
                    }
                    UZ_PROF_T(p3);
                    mbar_wait_cluster_wd(&acc_empty[i], (uAcc[i] & 1u) ^ 1u);
                    UZ_PROF_T(p4);
                    UZ_PROF_ADD(1, p2, p3);
                    UZ_PROF_ADD(2, p3, p4);
                    UZ_PROF2_ADD(4 + i, p2, p3);
                    UZ_PROF2_ADD(8 + (t & 3) * 2 + i, p3, p4);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_addr = smem_u32(sA + ab * kMmaABytes), b_addr = smem_u32(sB + slot * kMma2BBytes);
#pragma unroll
                        for (int k = 0; k < kE8RowBytes / 32; ++k)
                            tc2_mma_i8(tmem_base + (uint32_t)i * kMmaN, make_smem_desc(a_addr + k * 256, dsc),
                                       make_smem_desc(b_addr + k * 256, dsc), idesc, k > 0 ? 1u : 0u);
                        tc2_commit(&acc_full[i], 3);
                        if (t == T - 1) tc2_commit(&a_empty[ab], 3);
                    }
                    __syncwarp();
                    uAcc[i]++;
                }
                if (lane == 0) tc2_commit(&b_empty[slot], 3);
                __syncwarp();
                uB++;
            }
            if (T > 0) for (int i = 0; i < nblk; ++i) uA[(seq % ASETS) * 2 + i]++;
        }
#ifdef UZ_MMA_PROF
        if (lane == 0 && cluster < 256) {
            prof[3] = clock64() - prof_begin;
            for (int k = 0; k < 4; ++k) g_mma_prof[cluster][k] = prof[k];
            for (int k = 0; k < 16; ++k) g_mma_prof2[cluster][k] = prof2[k];
        }
#endif
    } else {
        // ===================== epilogue (both CTAs) =====================
        const int ew = warp - 2;                   // 0..7
        const int quarter = warp & 3;              // the TMEM lanes this warp may touch: 32 * (warp % 4) ..
        const int half = ew >> 2;                  // which 128 columns of every accumulator
        const int row_in_tile = quarter * 32 + lane;
        uint32_t uAcc[2] = {0, 0};
        uint32_t item_parity = 0;
        uint32_t r_acc_empty[2];
        for (int i = 0; i < 2; ++i) r_acc_empty[i] = mapa_u32(&acc_empty[i], 0);
        for (int it = cluster; it < n_items; it += n_clusters, item_parity ^= 1u) {
            const int2 item = items[it];
            const MmaTask* tk = tasks + item.x;
            const int nq = tk->nq, nt = tk->nt, q0 = item.y;
            const int nblk = (nq - q0 > 2 * kMmaM) ? 2 : 1;
            const int T = (nt + kMmaN - 1) / kMmaN;
            uint32_t m1[2] = {kNoKey, kNoKey}, m2[2] = {kNoKey, kNoKey};
            for (int t = 0; t < T; ++t) {
                const int cvalid = min(kMmaN, nt - t * kMmaN) - half * 128;     // valid columns of this warp's half
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (i < nblk) {
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
                            if (lane == 0) mbar_arrive_remote(r_acc_empty[i]);
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
                            if (lane == 0) mbar_arrive_remote(r_acc_empty[i]);
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
            uint2* xc = xchg + (XP == 2 ? item_parity : 0u) * 2 * kMmaM;
            if (half == 1) {
#pragma unroll
                for (int i = 0; i < 2; ++i) xc[i * kMmaM + row_in_tile] = make_uint2(m1[i], m2[i]);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (half == 0) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int q = q0 + i * 2 * kMmaM + (int)rank * kMmaM + row_in_tile;
                    if (i < nblk && q < nq) {
                        const uint2 o = xc[i * kMmaM + row_in_tile];
                        const uint32_t hi = max(m1[i], o.x);
                        const uint32_t a = min(m1[i], o.x);
                        const uint32_t b = min(hi, min(m2[i], o.y));
                        keys[(size_t)tk->key_off + q] = make_uint2(a, b);
                    }
                }
            }
            if (XP == 1) asm volatile("bar.sync 1, 256;" ::: "memory");      // one exchange buffer: read before the next item's write
        }
    }

    // nobody leaves while the partner may still read its shared memory or arrive on its barriers
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace uz
