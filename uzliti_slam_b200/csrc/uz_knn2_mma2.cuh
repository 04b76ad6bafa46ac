// uz_knn2_mma2.cuh — the tensor-core match kernel (uz_knn2_mmak.cuh: keys formed by the MMA) on CTA PAIRS:
// tcgen05.mma.cta_group::2, 256 x 256 x 32.
//
// Why.  With the epilogue reduced to packed max, knn2_mmak_kernel is bound by what feeds the tensor pipe: two 64 KB train
// stages are all that fits beside 64 KB of query tiles, and at every item boundary the pipe idles while the next query
// tile loads into the buffer that just became free (~1350 of 9200 clocks per item).  A CTA pair splits every 256-row train
// tile between its two SMs - each CTA stages only ITS 128 train rows (32 KB) - so four stages fit, and a train tile serves
// 512 query rows (the L2 -> SM traffic per compare halves).  Four stages hold TWO train tiles in use plus two in flight, which
// allows the order (t, block 0), (t + 1, block 0), (t, block 1), (t + 1, block 1): a query buffer is free two instruction
// groups (2300 clocks) before the next item needs it, enough to reload it.
// scripts/mma2_probe.cu: four 32 KB stages keep a pair at 128 clocks per instruction even when every tile comes from DRAM.
//
// Same contract, same E8 operand layout and key arithmetic as knn2_mmak_kernel; results are bit-identical.
//
// Mapping.  Cluster of two CTAs (one TPC), 320 threads each.  An item is 512 query rows of one matching = two 256-row
// blocks; block i is the M = 256 operand: CTA r holds rows [256 i + 128 r, +128) of it in its own shared memory and receives
// exactly those rows of the accumulator in its own TMEM (lanes 0..127, 256 columns).  Two accumulators per CTA, used in turn
// by the steps of the order above.
//   warp 0      producer (both CTAs): its 128-row query tile of each block, its 128-row half of every train tile, 4 stages
//   warp 1      CTA 0: MMA issuer - waits for both halves of a stage (its own mbarrier + a "peer" mbarrier), issues
//                      9 x tcgen05.mma.cta_group::2 per step, commits with multicast to both CTAs
//               CTA 1: relay - forwards "my half has landed" to the leader's peer barriers (remote mbarrier arrive); a 1-D
//                      bulk copy can only signal a barrier next to its destination
//   warps 2..9  epilogue as in knn2_mmak_kernel; CTA 1's warps release an accumulator with a remote arrive
// Barrier operations between the CTAs use the default (CTA-scope) semantics, as CUTLASS' ClusterBarrier does: with
// .release.cluster / .acquire.cluster on every arrive and poll the kernel ran 30 % slower.
// The last train tile is always issued with N = 256 (at cta_group::2 a shorter N would re-split the columns between the
// CTAs); the epilogue masks the columns past the last train row.
#pragma once
#include "uz_knn2_mmak.cuh"

namespace uz {

constexpr int kMma2ItemRows = 4 * kMmaM;            // 512 query rows per item and CTA pair
constexpr int kMma2Stages = 4;
constexpr int kMma2BBytes = kMmaM * kE8RowBytes;    // this CTA's half of a train tile: 128 rows = 32 KB
constexpr int kMma2SmemBytes = 2 * kMmaABytes + kMma2Stages * kMma2BBytes + 2 * kMmakTailABytes + 2 * 2 * kMmaM * 8 + 256;
static_assert(kMma2SmemBytes <= 232448, "CTA exceeds the 227 KB of shared memory");

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
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
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

// rows of [first, first + 128) that exist in a camera of n rows
__device__ __forceinline__ int rows_in(int n, int first) { return max(0, min(kMmaM, n - first)); }

// one item as every role sees it
struct Mma2Item {
    const MmaTask* tk;
    int nq, nt, q0, nblk, T;
};
__device__ __forceinline__ Mma2Item mma2_item(const MmaTask* tasks, const int2* items, int it) {
    Mma2Item m;
    const int2 item = items[it];
    m.tk = tasks + item.x;
    m.nq = m.tk->nq; m.nt = m.tk->nt; m.q0 = item.y;
    m.nblk = (m.nq - m.q0 > 2 * kMmaM) ? 2 : 1;
    m.T = (m.nt + kMmaN - 1) / kMmaN;
    return m;
}

// items[k] = (task, first query row, a multiple of 512); cluster c takes items c, c + clusters, ...
__global__ void __launch_bounds__(kMmaThreads, 1) knn2_mma2_kernel(const MmaTask* __restrict__ tasks, const int2* __restrict__ items,
                                                                   int n_items, uint2* __restrict__ keys, MmaDesc dsc) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int NS = kMma2Stages;
    uint8_t* sA = smem;                                                     // [2][32 KB]  block i, this CTA's 128 rows
    uint8_t* sB = smem + 2 * kMmaABytes;                                    // [NS][32 KB] this CTA's half of a train tile
    uint8_t* sTailA = sB + NS * kMma2BBytes;                                // [128 rows x 32 B] constant K-slice, query side
    uint8_t* sTailB = sTailA + kMmakTailABytes;                             // [128 rows x 32 B] ... train side (this CTA's half)
    uint2* xchg = reinterpret_cast<uint2*>(sTailB + kMmakTailABytes);       // [2 parities][256 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + 2 * 2 * kMmaM);
    uint64_t* a_full = bars;                          // [2]  own query tile landed
    uint64_t* a_peer = bars + 2;                      // [2]  (leader) the peer's query tile landed
    uint64_t* a_empty = bars + 4;                     // [2]  multicast commit
    uint64_t* b_full = bars + 6;                      // [NS]
    uint64_t* b_peer = b_full + NS;                   // [NS]
    uint64_t* b_empty = b_full + 2 * NS;              // [NS]  multicast commit
    uint64_t* acc_full = b_full + 3 * NS;             // [2]   multicast commit
    uint64_t* acc_empty = acc_full + 2;               // [2]   (leader) 8 local + 8 remote epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    static_assert((6 + 3 * NS + 4) * 8 + 4 <= 256, "barrier block");

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1); mbar_init(&a_peer[i], 1); mbar_init(&a_empty[i], 1);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 16);
        }
        for (int s = 0; s < NS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_peer[s], 1); mbar_init(&b_empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the constant K-slice (uz_knn2_mmak.cuh): 16 x (32 * 32) = 16384, + 1 * (127 - (train row & 127)); a CTA's 128 train rows are
    // rows 128 r .. of the tile, so (row & 127) is the local row in both CTAs
    for (int i = tid; i < 2 * kMmaM * 2; i += kMmaThreads) {
        const bool isB = i >= kMmaM * 2;
        const int r = (isB ? i - kMmaM * 2 : i) >> 1, kc = i & 1;
        uint4 v;
        if (kc == 0) v = make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
        else v = make_uint4(isB ? (uint32_t)(127 - r) : 1u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>((isB ? sTailB : sTailA) + (r >> 3) * 256 + kc * 128 + (r & 7) * 16) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 1) {          // the whole TMEM of both SMs: two 256-column accumulators per CTA
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();       // barriers and constant slices of both CTAs are ready before anyone touches them remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== producer (both CTAs) =====================
        // Issue order = the order in which the issuer frees and needs things: the first two train tiles of an item go out
        // while the previous item still computes (their stages are free two steps before its query buffers are).
        if (lane == 0) {
            uint32_t uB = 0, uA[2] = {0, 0};
            auto load_b = [&](const Mma2Item& m, int t) {
                const uint32_t slot = uB % NS;
                mbar_wait_wd(&b_empty[slot], ((uB / NS) & 1u) ^ 1u);
                const int first = t * kMmaN + (int)rank * kMmaM;
                const int rows = rows_in(m.nt, first);
                if (rows > 0) {
                    const uint32_t bytes = (uint32_t)e8_bytes(rows);
                    mbar_expect_tx(&b_full[slot], bytes);
                    bulk_g2s(sB + slot * kMma2BBytes, mma_t(m.tk) + (size_t)(first >> 3) * kE8GroupBytes, bytes, &b_full[slot]);
                } else {
                    mbar_arrive(&b_full[slot]);          // nothing of this tile in this CTA: the stage keeps stale bytes, columns masked
                }
                uB++;
            };
            int it = cluster;
            if (it < n_items) {
                const Mma2Item m = mma2_item(tasks, items, it);
                for (int t = 0; t < min(2, m.T); ++t) load_b(m, t);
            }
            for (; it < n_items; it += n_clusters) {
                const Mma2Item m = mma2_item(tasks, items, it);
                if (m.T > 0) {
                    for (int i = 0; i < m.nblk; ++i) {
                        mbar_wait_wd(&a_empty[i], (uA[i] & 1u) ^ 1u);
                        const int first = m.q0 + i * 2 * kMmaM + (int)rank * kMmaM;
                        const int rows = rows_in(m.nq, first);
                        if (rows > 0) {
                            const uint32_t bytes = (uint32_t)e8_bytes(rows);
                            mbar_expect_tx(&a_full[i], bytes);
                            bulk_g2s(sA + i * kMmaABytes, mma_q(m.tk) + (size_t)(first >> 3) * kE8GroupBytes, bytes, &a_full[i]);
                        } else {
                            mbar_arrive(&a_full[i]);     // nothing of this block in this CTA: stale bytes, rows unused
                        }
                        uA[i]++;
                    }
                }
                for (int t = 2; t < m.T; ++t) load_b(m, t);
                if (it + n_clusters < n_items) {
                    const Mma2Item nx = mma2_item(tasks, items, it + n_clusters);
                    for (int t = 0; t < min(2, nx.T); ++t) load_b(nx, t);
                }
            }
        }
    } else if (warp == 1 && rank == 1) {
        // ===================== relay (CTA 1), in the producer's order =====================
        if (lane == 0) {
            uint32_t uB = 0, uA[2] = {0, 0};
            uint32_t ra_peer[2], rb_peer[NS];
            for (int i = 0; i < 2; ++i) ra_peer[i] = mapa_u32(&a_peer[i], 0);
            for (int s = 0; s < NS; ++s) rb_peer[s] = mapa_u32(&b_peer[s], 0);
            auto fwd_b = [&]() {
                const uint32_t slot = uB % NS;
                mbar_wait_wd(&b_full[slot], (uB / NS) & 1u);
                mbar_arrive_remote(rb_peer[slot]);
                uB++;
            };
            int it = cluster;
            if (it < n_items) {
                const Mma2Item m = mma2_item(tasks, items, it);
                for (int t = 0; t < min(2, m.T); ++t) fwd_b();
            }
            for (; it < n_items; it += n_clusters) {
                const Mma2Item m = mma2_item(tasks, items, it);
                if (m.T > 0) {
                    for (int i = 0; i < m.nblk; ++i) {
                        mbar_wait_wd(&a_full[i], uA[i] & 1u);
                        mbar_arrive_remote(ra_peer[i]);
                        uA[i]++;
                    }
                }
                for (int t = 2; t < m.T; ++t) fwd_b();
                if (it + n_clusters < n_items) {
                    const Mma2Item nx = mma2_item(tasks, items, it + n_clusters);
                    for (int t = 0; t < min(2, nx.T); ++t) fwd_b();
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (CTA 0) =====================
        uint32_t tile0 = 0, uA[2] = {0, 0}, step = 0;
        // M = 256 (both CTAs), N = 256
        const uint32_t idesc = (dsc.idesc_base & ~(0x1Fu << 24)) | ((uint32_t)((2 * kMmaM) >> 4) << 24) | ((uint32_t)(kMmaN >> 3) << 17);
        MmaDesc tail_dsc = dsc;
        tail_dsc.lbo16 = 128 >> 4; tail_dsc.sbo16 = 256 >> 4;
        const uint64_t tail_a = make_smem_desc(smem_u32(sTailA), tail_dsc), tail_b = make_smem_desc(smem_u32(sTailB), tail_dsc);
#ifdef UZ_MMA_PROF
        long long prof[4] = {0, 0, 0, 0};
        const long long prof_begin = clock64();
#endif
        for (int it = cluster; it < n_items; it += n_clusters) {
            const Mma2Item m = mma2_item(tasks, items, it);
            for (int tp = 0; tp < m.T; tp += 2) {
                const int gn = min(2, m.T - tp);
                for (int i = 0; i < m.nblk; ++i) {
                    for (int g = 0; g < gn; ++g) {
                        const int t = tp + g;
                        const uint32_t ub = tile0 + (uint32_t)t, slot = ub % NS, bpar = (ub / NS) & 1u, acc = step & 1u;
                        UZ_PROF_T(p0);
                        if (i == 0) { mbar_wait_wd(&b_full[slot], bpar); mbar_wait_wd(&b_peer[slot], bpar); }
                        UZ_PROF_T(p1);
                        if (tp == 0 && g == 0) { mbar_wait_wd(&a_full[i], uA[i] & 1u); mbar_wait_wd(&a_peer[i], uA[i] & 1u); }
                        UZ_PROF_T(p2);
                        mbar_wait_wd(&acc_empty[acc], ((step >> 1) & 1u) ^ 1u);
                        UZ_PROF_T(p3);
                        UZ_PROF_ADD(0, p0, p1); UZ_PROF_ADD(1, p1, p2); UZ_PROF_ADD(2, p2, p3);
                        tc_fence_after();
                        if (tc_elect_one()) {
                            const uint32_t a_addr = smem_u32(sA + i * kMmaABytes), b_addr = smem_u32(sB + slot * kMma2BBytes);
#pragma unroll
                            for (int k = 0; k < kE8RowBytes / 32; ++k)
                                tc2_mma_i8(tmem_base + acc * kMmaN, make_smem_desc(a_addr + k * 256, dsc),
                                           make_smem_desc(b_addr + k * 256, dsc), idesc, k > 0 ? 1u : 0u);
                            tc2_mma_i8(tmem_base + acc * kMmaN, tail_a, tail_b, idesc, 1u);
                            tc2_commit(&acc_full[acc], 3);
                            if (t == m.T - 1) tc2_commit(&a_empty[i], 3);
                            if (i == m.nblk - 1) tc2_commit(&b_empty[slot], 3);
                        }
                        __syncwarp();
                        ++step;
                    }
                }
            }
            if (m.T > 0) for (int i = 0; i < m.nblk; ++i) uA[i]++;
            tile0 += (uint32_t)m.T;
        }
#ifdef UZ_MMA_PROF
        if (lane == 0 && cluster < 256) {
            prof[3] = clock64() - prof_begin;
            for (int k = 0; k < 4; ++k) g_mma_prof[cluster][k] = prof[k];
        }
#endif
    } else {
        // ===================== epilogue (both CTAs) =====================
        const int ew = warp - 2;                   // 0..7
        const int quarter = warp & 3;              // the TMEM lanes this warp may touch: 32 * (warp % 4) ..
        const int half = ew >> 2;                  // which 128 columns of every accumulator
        const int row_in_tile = quarter * 32 + lane;
        uint32_t step = 0, item_parity = 0;
        uint32_t r_acc_empty[2];
        for (int i = 0; i < 2; ++i) r_acc_empty[i] = mapa_u32(&acc_empty[i], 0);
        for (int it = cluster; it < n_items; it += n_clusters, item_parity ^= 1u) {
            const Mma2Item m = mma2_item(tasks, items, it);
            uint32_t m1[2] = {kNoKey, kNoKey}, m2[2] = {kNoKey, kNoKey};
            for (int tp = 0; tp < m.T; tp += 2) {
                const int gn = min(2, m.T - tp);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (i >= m.nblk) break;
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        if (g >= gn) break;
                        const int t = tp + g;
                        const uint32_t acc = step & 1u;
                        const int cvalid = min(kMmaN, m.nt - t * kMmaN) - half * 128;     // valid columns of this warp's half
                        mbar_wait_wd(&acc_full[acc], (step >> 1) & 1u);
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kMmaN + (uint32_t)(half * 128);
                        const uint32_t tbase = (uint32_t)(t * kMmaN + half * 128);       // global train row of column 0
                        uint32_t p1 = 0u, p2 = 0u;
                        uint32_t dA[32], dB[32];
                        if (cvalid >= 128) {
                            tc_ld64p(taddr, dA);
                            tc_ld64p(taddr + 64, dB);
                            tc_wait_ld(); tc_pin(dA); tc_pin(dB);
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_remote(r_acc_empty[acc]);
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
                            if (lane == 0) mbar_arrive_remote(r_acc_empty[acc]);
                        }
                        ++step;
                    }
                }
            }
            // fold the two column halves of every row (half 1 -> shared memory -> half 0) and publish the keys
            uint2* xc = xchg + item_parity * 2 * kMmaM;
            if (half == 1) {
#pragma unroll
                for (int i = 0; i < 2; ++i) xc[i * kMmaM + row_in_tile] = make_uint2(m1[i], m2[i]);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (half == 0) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int q = m.q0 + i * 2 * kMmaM + (int)rank * kMmaM + row_in_tile;
                    if (i < m.nblk && q < m.nq) {
                        const uint2 o = xc[i * kMmaM + row_in_tile];
                        const uint32_t hi = max(m1[i], o.x);
                        const uint32_t a = min(m1[i], o.x);
                        const uint32_t b = min(hi, min(m2[i], o.y));
                        keys[(size_t)m.tk->key_off + q] = make_uint2(a, b);
                    }
                }
            }
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
