// uz_knn2.cuh — K1: brute-force Hamming kNN-2 over 256-bit descriptors on the sm_100a integer pipes.
//
// Replaces cv::BFMatcher(NORM_HAMMING).knnMatch(query = to->features_, train = from->features_, k = 2)
// (/root/reference/transformation_estimation/src/feature_transformation_estimator.cpp:38,58).
// Output per query row: two packed keys  key = (distance << 16) | trainIdx,  m1 < m2, which is exactly
// OpenCV's order by (distance, trainIdx); 0xFFFFFFFF marks a missing neighbour (nt < 2).
//
// Mapping: one CTA per (matching, query tile).  Every thread keeps QPT query descriptors in registers
// (8 x u32 each; QPT = 2 in the shapes the host picks, 40 registers, 48 resident warps per SM); train
// descriptors stream through shared memory in 2-stage tiles filled by 1-D TMA bulk copies
// (cp.async.bulk + mbarrier) and are read with warp-broadcast 128-bit loads, so one LDS.128 pair feeds
// 32 x QPT compares.
//
// Arithmetic.  A 256-bit compare is 8 XOR + 8 POPC in the textbook form; POPC issues at a quarter of
// the LOP3 rate, so the kernel trades POPCs for LOP3s with a carry-save adder tree.  Descriptors are
// stored in a "CSA layout" (an invertible XOR transform done once at ingestion, see csa_pack):
//   W0=w0  W1=w1  W2=w0^w1^w2  W3=w3  W4=w4  W5=w3^w4^w5  W6=w0^..^w6  W7=w7
// With x_i = q_i ^ t_i, the XOR of two CSA-layout rows yields x0,x1,S1=x0^x1^x2,x3,x4,S2=x3^x4^x5,
// S3=x0^..^x6,x7 directly, and the carries follow with one LOP3 each:
//   C1=maj(x0,x1,x2)  C2=maj(x3,x4,x5)  C3=maj(S1,S2,x6)       (x2,x5,x6 are implied: lut 0xD4)
//   S5=C1^C2^C3       C5=maj(C1,C2,C3)
//   distance = popc(S3) + popc(x7) + 2 popc(S5) + 4 popc(C5)
// = 13 LOP3 + 4 POPC + 4 IMAD (weights and the key shift folded into the key build); the top-2 update costs
// 2.5 VIMNMX per compare on 32-bit keys and 1.25 on the packed 16-bit keys below (PACK16, the default).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace uz {

constexpr uint32_t kNoKey = 0xFFFFFFFFu;
constexpr int kStages = 2;
// Shape of one CTA: THREADS x QPT query rows, train rows staged in kStages tiles of knn_train_rows() rows.  The
// two-queries-per-thread shapes scale their staging with the CTA (128 B of shared memory per thread) and are compiled
// for 40 registers, so that every one of them keeps 48 warps resident per SM (32 for the one-warp CTA: 32 CTAs per SM is
// the hardware limit); small CTAs exist so that ragged query counts waste little (a 400-row camera costs 448 rows in
// one-warp CTAs, 512 in anything larger).
__host__ __device__ constexpr int knn_train_rows(int threads, int qpt) { return qpt == 2 ? (threads >= 64 ? 2 * threads : 128) : 512; }
__host__ __device__ constexpr int knn_smem_bytes(int threads, int qpt) { return kStages * knn_train_rows(threads, qpt) * 32 + 64; }
__host__ __device__ constexpr int knn_min_ctas(int threads, int qpt) {
    return qpt == 2 ? (threads == 256 ? 6 : threads == 128 ? 12 : threads == 64 ? 24 : 32) : 1;
}


struct MatchTask {
    const uint32_t* q_desc;   // "to" camera descriptors (OpenCV query), layout per UZ variant, 32 B rows
    const uint32_t* t_desc;   // "from" camera descriptors (OpenCV train)
    int32_t nq, nt;
    uint32_t key_off;         // first row of this matching in the keys scratch
    int32_t pair;             // owning pair (index within the launch)
    // solve-side views of the two cameras
    const double* q_pos; const uint8_t* q_valid;   // to
    const double* t_pos; const uint8_t* t_valid;   // from
    int32_t cam_from, cam_to;
    uint32_t rev_key_off;     // cross-check: first keys row of the REVERSED matching (query = from, train = to), kNoRev = off
    uint32_t pad_;
};
constexpr uint32_t kNoRev = 0xFFFFFFFFu;

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Popcount weights (<<16 folded in).  They live in constant memory on purpose: with immediate powers of two
// ptxas strength-reduces one of the four multiply-adds into an ALU-pipe LEA, and the ALU pipe is this
// kernel's binding unit (13 LOP3 + 3 VIMNMX per compare); a constant-bank operand keeps all four on the
// otherwise idle FMA pipe as IMAD.
__constant__ uint32_t kPopcWeight[3] = {1u << 16, 2u << 16, 4u << 16};

// raw 8 words -> CSA layout (also its own inverse is csa_unpack; both are XOR-linear)
__host__ __device__ __forceinline__ void csa_pack(const uint32_t* w, uint32_t* o) {
    const uint32_t s1 = w[0] ^ w[1] ^ w[2];
    const uint32_t s2 = w[3] ^ w[4] ^ w[5];
    o[0] = w[0]; o[1] = w[1]; o[2] = s1; o[3] = w[3]; o[4] = w[4]; o[5] = s2; o[6] = s1 ^ s2 ^ w[6]; o[7] = w[7];
}

// key = (hamming << 16) + jkey for two CSA-layout rows
__device__ __forceinline__ uint32_t csa_key(const uint32_t (&U)[8], const uint4& a, const uint4& b, uint32_t jkey) {
    const uint32_t x0 = U[0] ^ a.x, x1 = U[1] ^ a.y, S1 = U[2] ^ a.z;
    const uint32_t C1 = lop3<0xD4>(x0, x1, S1);
    const uint32_t x3 = U[3] ^ a.w, x4 = U[4] ^ b.x, S2 = U[5] ^ b.y;
    const uint32_t C2 = lop3<0xD4>(x3, x4, S2);
    const uint32_t S3 = U[6] ^ b.z;
    const uint32_t C3 = lop3<0xD4>(S1, S2, S3);
    const uint32_t x7 = U[7] ^ b.w;
    const uint32_t S5 = lop3<0x96>(C1, C2, C3);
    const uint32_t C5 = lop3<0xE8>(C1, C2, C3);
    uint32_t k = mad_u32(__popc(C5), kPopcWeight[2], jkey);
    k = mad_u32(__popc(S5), kPopcWeight[1], k);
    k = mad_u32(__popc(x7), kPopcWeight[0], k);
    k = mad_u32(__popc(S3), kPopcWeight[0], k);
    return k;
}

// textbook form on raw rows (8 XOR + 8 POPC) — kept as the measured alternative (UZ_KNN_VARIANT=1)
__device__ __forceinline__ uint32_t plain_key(const uint32_t (&U)[8], const uint4& a, const uint4& b, uint32_t jkey) {
    const uint32_t d0 = __popc(U[0] ^ a.x) + __popc(U[1] ^ a.y) + __popc(U[2] ^ a.z) + __popc(U[3] ^ a.w);
    const uint32_t d1 = __popc(U[4] ^ b.x) + __popc(U[5] ^ b.y) + __popc(U[6] ^ b.z) + __popc(U[7] ^ b.w);
    return mad_u32(d0 + d1, 1u << 16, jkey);
}

__device__ __forceinline__ void top2_update(uint32_t& m1, uint32_t& m2, uint32_t k) {
    const uint32_t hi = max(m1, k);
    m1 = min(m1, k);
    m2 = min(m2, hi);
}
// two new keys at once: 5 min/max (one of them a 3-input VIMNMX3) instead of 6
__device__ __forceinline__ void top2_update2(uint32_t& m1, uint32_t& m2, uint32_t ka, uint32_t kb) {
    const uint32_t lo = min(ka, kb), hi = max(ka, kb);
    const uint32_t t = max(m1, lo);
    m1 = min(m1, lo);
    m2 = min(min(m2, t), hi);
}

// ---- packed 16-bit keys: two queries per register ---------------------------------------------------
// Inside a block of 128 train rows a key needs only 9 bits of distance (0..256) and 7 bits of row:
//   key16 = (distance << 7) | (row & 127)   (<= 32895, 0xFFFF = none)
// so the keys of TWO queries against the same train row share one register and the running top-2 of both
// is kept with VIMNMX.U16x2 / VIMNMX3.U16x2: 2.5 min/max per two compares instead of 2.5 per compare (the
// min/max share the ALU pipe with the 13 LOP3 of a compare).  The weighted popcounts of both queries are
// accumulated by one IMAD chain (weights << 7 for the low half, << 23 for the high half; no carry crosses
// the halves).  Every 128 rows the four 16-bit winners are widened to full keys and merged into m1/m2.
__constant__ uint32_t kPopcWeightLo[3] = {1u << 7, 2u << 7, 4u << 7};
__constant__ uint32_t kPopcWeightHi[3] = {1u << 23, 2u << 23, 4u << 23};

template <bool HI>
__device__ __forceinline__ uint32_t csa_acc16(const uint32_t (&U)[8], const uint4& a, const uint4& b, uint32_t acc) {
    const uint32_t x0 = U[0] ^ a.x, x1 = U[1] ^ a.y, S1 = U[2] ^ a.z;
    const uint32_t C1 = lop3<0xD4>(x0, x1, S1);
    const uint32_t x3 = U[3] ^ a.w, x4 = U[4] ^ b.x, S2 = U[5] ^ b.y;
    const uint32_t C2 = lop3<0xD4>(x3, x4, S2);
    const uint32_t S3 = U[6] ^ b.z;
    const uint32_t C3 = lop3<0xD4>(S1, S2, S3);
    const uint32_t x7 = U[7] ^ b.w;
    const uint32_t S5 = lop3<0x96>(C1, C2, C3);
    const uint32_t C5 = lop3<0xE8>(C1, C2, C3);
    const uint32_t w1 = HI ? kPopcWeightHi[0] : kPopcWeightLo[0];
    const uint32_t w2 = HI ? kPopcWeightHi[1] : kPopcWeightLo[1];
    const uint32_t w4 = HI ? kPopcWeightHi[2] : kPopcWeightLo[2];
    uint32_t k = mad_u32(__popc(C5), w4, acc);
    k = mad_u32(__popc(S5), w2, k);
    k = mad_u32(__popc(x7), w1, k);
    k = mad_u32(__popc(S3), w1, k);
    return k;
}
__device__ __forceinline__ uint32_t min_u16x2(uint32_t a, uint32_t b) {
    uint32_t d; asm("min.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ uint32_t max_u16x2(uint32_t a, uint32_t b) {
    uint32_t d; asm("max.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d;
}
__device__ __forceinline__ void top2_update2_u16x2(uint32_t& m1, uint32_t& m2, uint32_t ka, uint32_t kb) {
    const uint32_t lo = min_u16x2(ka, kb), hi = max_u16x2(ka, kb);
    const uint32_t t = max_u16x2(m1, lo);
    m1 = min_u16x2(m1, lo);
    m2 = min_u16x2(min_u16x2(m2, t), hi);
}
// widen one 16-bit winner pair (v1 < v2, both real) of the block starting at train row `base` and merge it
__device__ __forceinline__ void merge_block16(uint32_t& m1, uint32_t& m2, uint32_t v1, uint32_t v2, uint32_t base) {
    const uint32_t k1 = ((v1 & 0xFF80u) << 9) | (base + (v1 & 127u));
    const uint32_t k2 = ((v2 & 0xFF80u) << 9) | (base + (v2 & 127u));
    const uint32_t t = max(m1, k1);
    m1 = min(m1, k1);
    m2 = min(min(m2, t), k2);
}

// ---- fused cross-check (uz_params.cross_check): column minima beside the row top-2 -----------------------
// cv::BFMatcher(crossCheck = true) keeps (q, t) only if q is also the nearest query row of t (lowest index on ties).
// The distances that decide this are the ones the forward matching computes anyway, so the match kernel tracks, per
// train row, the minimum over its query tile instead of running the whole matching a second time, reversed:
//   per thread  v = min over its two queries of  (key16 << 16) | local query index      (1 IMAD + 1 LOP3 + 1 VIMNMX)
//   per warp    REDUX.MIN over the 32 lanes, one shared-memory ATOMS.MIN per warp and train row
//   per tile    one global atomicMin per train row on keys[rev_key_off + t].x = (distance << 16) | queryIdx
// The key16 of both queries carries the same train row, so the order of v is (distance, query index).  Threads
// beyond the last query of a ragged tile recompute the last valid query (same candidates, no effect on a minimum).
__constant__ uint32_t kShift16 = 65536u;
// v is warp-uniform (the REDUX result): lane 0 alone issues the shared-memory reduction.  The address arrives as a
// per-thread shared-window offset (col_base: base + 4 * lane, used by lane 0 only): with an address it can prove warp-uniform, ptxas
// wraps every such reduction in its own warp aggregation - a leader election and a second REDUX per call.
__device__ __forceinline__ uint32_t col_base(const uint32_t* s_col) {
    uint32_t a;             // base + 4 * lane: the right address in lane 0, the only lane that ever uses it
    asm volatile("mad.lo.u32 %0, %1, 4, %2;" : "=r"(a) : "r"(threadIdx.x & 31u), "r"((uint32_t)__cvta_generic_to_shared(s_col)));
    return a;
}
__device__ __forceinline__ void col_publish(uint32_t v, uint32_t s_col_row) {
    if ((threadIdx.x & 31) == 0) asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(s_col_row), "r"(v) : "memory");
}
template <int DSHIFT /* distance position inside key16: 7 (256-bit rows) or 6 (512-bit rows) */>
__device__ __forceinline__ void col_update16(uint32_t k, uint32_t c0, uint32_t c1, uint32_t s_col_row) {
    const uint32_t x = mad_u32(k, kShift16, c0);                 // query 0: (key16 << 16) + index       (FMA pipe)
    const uint32_t y = lop3<0xEA>(k, 0xFFFF0000u, c1);           // query 1: (k & 0xFFFF0000) | index
    col_publish(__reduce_min_sync(0xffffffffu, min(x, y)), s_col_row);
}
// the same from two full 32-bit keys (tail rows): only the distance and the index matter for the flush
template <int DSHIFT>
__device__ __forceinline__ void col_update32(uint32_t key0, uint32_t key1, uint32_t c0, uint32_t c1, uint32_t s_col_row) {
    const uint32_t x = ((key0 >> 16) << (16 + DSHIFT)) | c0;
    const uint32_t y = ((key1 >> 16) << (16 + DSHIFT)) | c1;
    col_publish(__reduce_min_sync(0xffffffffu, min(x, y)), s_col_row);
}
// end of a train tile: publish the tile's column minima and re-arm the shared array
template <int THREADS, int DSHIFT>
__device__ __forceinline__ void col_flush(uint32_t* s_col, int nrows, uint2* col_keys, int r0, int q_base) {
    for (int r = threadIdx.x; r < nrows; r += THREADS) {
        const uint32_t v = s_col[r];
        s_col[r] = 0xFFFFFFFFu;
        if (v != 0xFFFFFFFFu) atomicMin(&col_keys[r0 + r].x, ((v >> (16 + DSHIFT)) << 16) | (uint32_t)(q_base + (int)(v & 0xFFFFu)));
    }
}

// ---- mbarrier / bulk-copy helpers (PTX) --------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// tiles[blockIdx.x] = (task index, first query row of the tile)
// SEG (small launches): a tile also names a SEGMENT of the train rows, tiles[] holds int4 (task, first query row, first train
// row, (segment index << 16) | train rows of the segment), and the CTA writes its partial neighbours to
// keys[key_off + segment * nq + q]; merge_segments_kernel folds the segments into segment 0.  A batch of a few pairs then
// spreads over the whole chip instead of a handful of CTAs that each walk every train row.
template <int THREADS, int QPT, bool CSA, bool PACK16 = false, bool XCHK = false, bool SEG = false>
__global__ void __launch_bounds__(THREADS, knn_min_ctas(THREADS, QPT)) knn2_kernel(const MatchTask* __restrict__ tasks,
                                                       const int2* __restrict__ tiles,
                                                       uint2* __restrict__ keys,
                                                       int* __restrict__ pair_pending,
                                                       unsigned int* __restrict__ progress) {
    constexpr int kTrainTileRows = knn_train_rows(THREADS, QPT);
    static_assert(!XCHK || (PACK16 && QPT == 2), "the fused cross-check lives in the packed-key, two-queries-per-thread shapes");
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kTrainTileRows * 32);
    uint32_t* s_col = reinterpret_cast<uint32_t*>(smem + kStages * kTrainTileRows * 32 + 64);     // XCHK: [kTrainTileRows]
    const uint32_t s_col_a = XCHK ? col_base(s_col) : 0u;

    static_assert(!SEG || (PACK16 && QPT == 2), "segmented launches use the packed-key, two-queries-per-thread shapes");
    int4 tile4 = make_int4(0, 0, 0, 0);
    if (SEG) tile4 = reinterpret_cast<const int4*>(tiles)[blockIdx.x];
    const int2 tile = SEG ? make_int2(tile4.x, tile4.y) : tiles[blockIdx.x];
    const MatchTask* tk = tasks + tile.x;
    const int t0 = SEG ? tile4.z : 0;                                   // first train row of this CTA's segment
    const uint32_t* __restrict__ qd = tk->q_desc;
    const uint32_t* __restrict__ td = tk->t_desc + (size_t)t0 * 8;
    const int nq = tk->nq, nt = SEG ? (tile4.w & 0xFFFF) : tk->nt;      // train rows this CTA walks
    const int tid = threadIdx.x;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (XCHK) for (int r = tid; r < kTrainTileRows; r += THREADS) s_col[r] = 0xFFFFFFFFu;
    __syncthreads();

    const int ntiles = (nt + kTrainTileRows - 1) / kTrainTileRows;
    if (tid == 0 && ntiles > 0) {
        const uint32_t bytes = (uint32_t)min(nt, kTrainTileRows) * 32u;
        mbar_expect_tx(&bars[0], bytes);
        bulk_g2s(smem, td, bytes, &bars[0]);
    }

    // query rows of this thread: q0 + k*THREADS + tid (coalesced 32 B per thread)
    uint32_t U[QPT][8];
    uint32_t m1[QPT], m2[QPT];
    uint32_t cq[2] = {0u, 0u};          // XCHK: local index (inside the tile) of the query each slot really holds
#pragma unroll
    for (int k = 0; k < QPT; ++k) {
        int q = tile.y + k * THREADS + tid;
        if (XCHK) { q = min(q, nq - 1); cq[k & 1] = (uint32_t)(q - tile.y); }      // ragged tile: recompute the last query
        uint4 a = make_uint4(0, 0, 0, 0), b = a;
        if (q < nq) {
            const uint4* p = reinterpret_cast<const uint4*>(qd + (size_t)q * 8);
            a = __ldg(p);
            b = __ldg(p + 1);
        }
        U[k][0] = a.x; U[k][1] = a.y; U[k][2] = a.z; U[k][3] = a.w;
        U[k][4] = b.x; U[k][5] = b.y; U[k][6] = b.z; U[k][7] = b.w;
        m1[k] = kNoKey; m2[k] = kNoKey;
    }

    for (int ti = 0; ti < ntiles; ++ti) {
        const int stage = ti & 1;
        if (tid == 0 && ti + 1 < ntiles) {       // prefetch next tile into the other stage (freed by the
            const int r0 = (ti + 1) * kTrainTileRows;   // __syncthreads that ended iteration ti-1)
            const uint32_t bytes = (uint32_t)min(nt - r0, kTrainTileRows) * 32u;
            mbar_expect_tx(&bars[stage ^ 1], bytes);
            bulk_g2s(smem + (stage ^ 1) * kTrainTileRows * 32, td + (size_t)r0 * 8, bytes, &bars[stage ^ 1]);
        }
        mbar_wait(&bars[stage], (ti >> 1) & 1);
        const uint4* __restrict__ rows = reinterpret_cast<const uint4*>(smem + stage * kTrainTileRows * 32);
        const int r0 = ti * kTrainTileRows;
        const int nrows = min(nt - r0, kTrainTileRows);
        if (PACK16) {
            static_assert(!PACK16 || (CSA && QPT % 2 == 0), "packed keys need the CSA layout and an even QPT");
            for (int b0 = 0; b0 < nrows; b0 += 128) {
                const int nb = min(nrows - b0, 128);
                const uint4* __restrict__ brows = rows + 2 * b0;
                uint32_t p1[QPT / 2], p2[QPT / 2];
#pragma unroll
                for (int kk = 0; kk < QPT / 2; ++kk) { p1[kk] = 0xFFFFFFFFu; p2[kk] = 0xFFFFFFFFu; }
                int j = 0;
#pragma unroll 1
                for (; j + 4 <= nb; j += 4) {
#pragma unroll
                    for (int u = 0; u < 4; u += 2) {
                        const uint4 a0 = brows[2 * (j + u)], b0v = brows[2 * (j + u) + 1];
                        const uint4 a1 = brows[2 * (j + u) + 2], b1v = brows[2 * (j + u) + 3];
                        const uint32_t jj = (uint32_t)(j + u) * 0x00010001u;
#pragma unroll
                        for (int kk = 0; kk < QPT / 2; ++kk) {
                            uint32_t k0 = csa_acc16<false>(U[2 * kk], a0, b0v, jj);
                            uint32_t k1 = csa_acc16<false>(U[2 * kk], a1, b1v, jj + 0x00010001u);
                            k0 = csa_acc16<true>(U[2 * kk + 1], a0, b0v, k0);
                            k1 = csa_acc16<true>(U[2 * kk + 1], a1, b1v, k1);
                            top2_update2_u16x2(p1[kk], p2[kk], k0, k1);
                            if (XCHK) {
                                col_update16<7>(k0, cq[0], cq[1], s_col_a + 4u * (uint32_t)(b0 + j + u));
                                col_update16<7>(k1, cq[0], cq[1], s_col_a + 4u * (uint32_t)(b0 + j + u + 1));
                            }
                        }
                    }
                }
                if (j > 0) {
                    const uint32_t base = (uint32_t)(t0 + r0 + b0);
#pragma unroll
                    for (int kk = 0; kk < QPT / 2; ++kk) {
                        merge_block16(m1[2 * kk], m2[2 * kk], p1[kk] & 0xFFFFu, p2[kk] & 0xFFFFu, base);
                        merge_block16(m1[2 * kk + 1], m2[2 * kk + 1], p1[kk] >> 16, p2[kk] >> 16, base);
                    }
                }
                for (; j < nb; ++j) {
                    const uint4 a = brows[2 * j], b = brows[2 * j + 1];
                    const uint32_t jkey = (uint32_t)(t0 + r0 + b0 + j);
                    uint32_t kq[QPT];
#pragma unroll
                    for (int k = 0; k < QPT; ++k) { kq[k] = csa_key(U[k], a, b, jkey); top2_update(m1[k], m2[k], kq[k]); }
                    if (XCHK) col_update32<7>(kq[0], kq[QPT - 1], cq[0], cq[1], s_col_a + 4u * (uint32_t)(b0 + j));
                }
            }
        } else {
            int j = 0;
#pragma unroll 1
            for (; j + 4 <= nrows; j += 4) {
#pragma unroll
                for (int u = 0; u < 4; u += 2) {
                    const uint4 a0 = rows[2 * (j + u)], b0 = rows[2 * (j + u) + 1];
                    const uint4 a1 = rows[2 * (j + u) + 2], b1 = rows[2 * (j + u) + 3];
                    const uint32_t jkey = (uint32_t)(r0 + j + u);
#pragma unroll
                    for (int k = 0; k < QPT; ++k) {
                        const uint32_t k0 = CSA ? csa_key(U[k], a0, b0, jkey) : plain_key(U[k], a0, b0, jkey);
                        const uint32_t k1 = CSA ? csa_key(U[k], a1, b1, jkey + 1) : plain_key(U[k], a1, b1, jkey + 1);
                        top2_update2(m1[k], m2[k], k0, k1);
                    }
                }
            }
            for (; j < nrows; ++j) {
                const uint4 a = rows[2 * j], b = rows[2 * j + 1];
                const uint32_t jkey = (uint32_t)(r0 + j);
#pragma unroll
                for (int k = 0; k < QPT; ++k) {
                    const uint32_t key = CSA ? csa_key(U[k], a, b, jkey) : plain_key(U[k], a, b, jkey);
                    top2_update(m1[k], m2[k], key);
                }
            }
        }
        __syncthreads();
        if (XCHK) {
            col_flush<THREADS, 7>(s_col, nrows, keys + tk->rev_key_off, t0 + r0, tile.y);
            __syncthreads();
        }
    }

    const uint32_t out_off = tk->key_off + (SEG ? (uint32_t)(tile4.w >> 16) * (uint32_t)nq : 0u);
#pragma unroll
    for (int k = 0; k < QPT; ++k) {
        const int q = tile.y + k * THREADS + tid;
        if (q < nq) keys[out_off + q] = make_uint2(m1[k], m2[k]);
    }
    // Streaming hand-over to the solve kernel that runs beside this one (uz_solve.cuh, solve_stream_kernel):
    // pair_pending[pair] counts the tiles of the pair that have not published their keys yet.  bar.sync orders
    // every thread's key stores before thread 0's fence + atomic (release); the consumer polls with ld.acquire.
    if (pair_pending != nullptr) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicSub(pair_pending + tk->pair, 1);
            atomicAdd(progress, 1u);          // liveness signal for the consumer's starvation test
        }
    }
}

// ---- 512-bit rows (BRISK / FREAK: 64-byte descriptors, feature_transformation_estimator.cpp:56-57) ----------
// A wide row is stored as two independent CSA-layout halves (csa_pack applied to words 0..7 and 8..15), so a compare
// is two 256-bit CSA trees feeding ONE weighted-popcount chain: 26 LOP3 + 8 POPC + 8 IMAD per 512-bit compare, the
// same POPC-per-bit cost as the 256-bit kernel.  Distances reach 512, so the packed 16-bit keys carry 10 bits of
// distance and 6 bits of row:  key16 = (distance << 6) | (row & 63)  (<= 32831), merged into the 32-bit keys every
// 64 train rows.  Shape: THREADS x 2 queries per CTA (32 query words per thread, 64 registers, 32 resident warps per
// SM), THREADS train rows per stage.
#ifndef UZ_WIDE_MERGE
#define UZ_WIDE_MERGE 0        // measured alternative, see csa_acc16w_merged: 15.56 ms against 15.10 ms for the plain halves
#endif
__constant__ uint32_t kPopcWeightLo6[3] = {1u << 6, 2u << 6, 4u << 6};
__constant__ uint32_t kPopcWeightHi6[3] = {1u << 22, 2u << 22, 4u << 22};

__host__ __device__ constexpr int knn_wide_train_rows(int threads) { return threads; }
__host__ __device__ constexpr int knn_wide_smem_bytes(int threads) { return kStages * knn_wide_train_rows(threads) * 64 + 64; }
__host__ __device__ constexpr int knn_wide_min_ctas(int threads) { return threads == 256 ? 4 : threads == 128 ? 8 : 16; }

// one 256-bit half: U points at the 8 query words of that half (registers after unrolling)
template <bool HI>
__device__ __forceinline__ uint32_t csa_acc16w(const uint32_t* U, const uint4& a, const uint4& b, uint32_t acc) {
    const uint32_t x0 = U[0] ^ a.x, x1 = U[1] ^ a.y, S1 = U[2] ^ a.z;
    const uint32_t C1 = lop3<0xD4>(x0, x1, S1);
    const uint32_t x3 = U[3] ^ a.w, x4 = U[4] ^ b.x, S2 = U[5] ^ b.y;
    const uint32_t C2 = lop3<0xD4>(x3, x4, S2);
    const uint32_t S3 = U[6] ^ b.z;
    const uint32_t C3 = lop3<0xD4>(S1, S2, S3);
    const uint32_t x7 = U[7] ^ b.w;
    const uint32_t S5 = lop3<0x96>(C1, C2, C3);
    const uint32_t C5 = lop3<0xE8>(C1, C2, C3);
    const uint32_t w1 = HI ? kPopcWeightHi6[0] : kPopcWeightLo6[0];
    const uint32_t w2 = HI ? kPopcWeightHi6[1] : kPopcWeightLo6[1];
    const uint32_t w4 = HI ? kPopcWeightHi6[2] : kPopcWeightLo6[2];
    uint32_t k = mad_u32(__popc(C5), w4, acc);
    k = mad_u32(__popc(S5), w2, k);
    k = mad_u32(__popc(x7), w1, k);
    k = mad_u32(__popc(S3), w1, k);
    return k;
}
// One whole 512-bit compare with the two halves' weight-1 planes merged by one more full adder (UZ_WIDE_MERGE):
// (S3a, x7a, S3b) -> sum (weight 1) + carry (weight 2), so 7 POPC + 28 LOP3 instead of 8 + 26.  The XU pipe (POPC) is
// this kernel's binding unit at 94 % busy and the ALU pipe has slack on paper, but the trade LOSES on B200: 514 against
// 530 G cmp512/s (same instruction count, the two pipes share dispatch bandwidth).  Kept as the measured alternative.
template <bool HI>
__device__ __forceinline__ uint32_t csa_acc16w_merged(const uint32_t* U, const uint4& a, const uint4& b, const uint4& c,
                                                      const uint4& d, uint32_t acc) {
    const uint32_t w1 = HI ? kPopcWeightHi6[0] : kPopcWeightLo6[0];
    const uint32_t w2 = HI ? kPopcWeightHi6[1] : kPopcWeightLo6[1];
    const uint32_t w4 = HI ? kPopcWeightHi6[2] : kPopcWeightLo6[2];
    uint32_t S3a, x7a;
    {
        const uint32_t x0 = U[0] ^ a.x, x1 = U[1] ^ a.y, S1 = U[2] ^ a.z;
        const uint32_t C1 = lop3<0xD4>(x0, x1, S1);
        const uint32_t x3 = U[3] ^ a.w, x4 = U[4] ^ b.x, S2 = U[5] ^ b.y;
        const uint32_t C2 = lop3<0xD4>(x3, x4, S2);
        S3a = U[6] ^ b.z;
        const uint32_t C3 = lop3<0xD4>(S1, S2, S3a);
        x7a = U[7] ^ b.w;
        const uint32_t S5 = lop3<0x96>(C1, C2, C3);
        const uint32_t C5 = lop3<0xE8>(C1, C2, C3);
        acc = mad_u32(__popc(C5), w4, acc);
        acc = mad_u32(__popc(S5), w2, acc);
    }
    {
        const uint32_t x0 = U[8] ^ c.x, x1 = U[9] ^ c.y, S1 = U[10] ^ c.z;
        const uint32_t C1 = lop3<0xD4>(x0, x1, S1);
        const uint32_t x3 = U[11] ^ c.w, x4 = U[12] ^ d.x, S2 = U[13] ^ d.y;
        const uint32_t C2 = lop3<0xD4>(x3, x4, S2);
        const uint32_t S3b = U[14] ^ d.z;
        const uint32_t C3 = lop3<0xD4>(S1, S2, S3b);
        const uint32_t x7b = U[15] ^ d.w;
        const uint32_t S5 = lop3<0x96>(C1, C2, C3);
        const uint32_t C5 = lop3<0xE8>(C1, C2, C3);
        const uint32_t s1 = lop3<0x96>(S3a, x7a, S3b);       // weight 1
        const uint32_t c1 = lop3<0xE8>(S3a, x7a, S3b);       // weight 2
        acc = mad_u32(__popc(C5), w4, acc);
        acc = mad_u32(__popc(S5), w2, acc);
        acc = mad_u32(__popc(c1), w2, acc);
        acc = mad_u32(__popc(s1), w1, acc);
        acc = mad_u32(__popc(x7b), w1, acc);
    }
    return acc;
}
// full 32-bit key of one wide compare (tail rows)
__device__ __forceinline__ uint32_t csa_key_wide(const uint32_t* U, const uint4& a, const uint4& b, const uint4& c,
                                                 const uint4& d, uint32_t jkey) {
    uint32_t Ulo[8], Uhi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { Ulo[i] = U[i]; Uhi[i] = U[8 + i]; }
    return csa_key(Uhi, c, d, csa_key(Ulo, a, b, jkey));
}
__device__ __forceinline__ void merge_block16w(uint32_t& m1, uint32_t& m2, uint32_t v1, uint32_t v2, uint32_t base) {
    const uint32_t k1 = ((v1 & 0xFFC0u) << 10) | (base + (v1 & 63u));
    const uint32_t k2 = ((v2 & 0xFFC0u) << 10) | (base + (v2 & 63u));
    const uint32_t t = max(m1, k1);
    m1 = min(m1, k1);
    m2 = min(min(m2, t), k2);
}

template <int THREADS, bool XCHK = false, bool SEG = false>
__global__ void __launch_bounds__(THREADS, knn_wide_min_ctas(THREADS)) knn2_wide_kernel(const MatchTask* __restrict__ tasks,
                                                       const int2* __restrict__ tiles,
                                                       uint2* __restrict__ keys,
                                                       int* __restrict__ pair_pending,
                                                       unsigned int* __restrict__ progress) {
    constexpr int kRows = knn_wide_train_rows(THREADS);
    static_assert(kRows % 64 == 0, "train tiles are cut into 64-row key blocks");
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kRows * 64);
    uint32_t* s_col = reinterpret_cast<uint32_t*>(smem + kStages * kRows * 64 + 64);       // XCHK: [kRows]
    const uint32_t s_col_a = XCHK ? col_base(s_col) : 0u;

    int4 tile4 = make_int4(0, 0, 0, 0);
    if (SEG) tile4 = reinterpret_cast<const int4*>(tiles)[blockIdx.x];             // see knn2_kernel
    const int2 tile = SEG ? make_int2(tile4.x, tile4.y) : tiles[blockIdx.x];
    const MatchTask* tk = tasks + tile.x;
    const int t0 = SEG ? tile4.z : 0;
    const uint32_t* __restrict__ qd = tk->q_desc;
    const uint32_t* __restrict__ td = tk->t_desc + (size_t)t0 * 16;
    const int nq = tk->nq, nt = SEG ? (tile4.w & 0xFFFF) : tk->nt;
    const int tid = threadIdx.x;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (XCHK) for (int r = tid; r < kRows; r += THREADS) s_col[r] = 0xFFFFFFFFu;
    __syncthreads();

    const int ntiles = (nt + kRows - 1) / kRows;
    if (tid == 0 && ntiles > 0) {
        const uint32_t bytes = (uint32_t)min(nt, kRows) * 64u;
        mbar_expect_tx(&bars[0], bytes);
        bulk_g2s(smem, td, bytes, &bars[0]);
    }

    uint32_t U[2][16];
    uint32_t m1[2], m2[2];
    uint32_t cq[2] = {0u, 0u};          // XCHK: local index of the query each slot really holds (see col_update16)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        int q = tile.y + k * THREADS + tid;
        if (XCHK) { q = min(q, nq - 1); cq[k] = (uint32_t)(q - tile.y); }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            uint4 a = make_uint4(0, 0, 0, 0);
            if (q < nq) a = __ldg(reinterpret_cast<const uint4*>(qd + (size_t)q * 16) + h);
            U[k][4 * h] = a.x; U[k][4 * h + 1] = a.y; U[k][4 * h + 2] = a.z; U[k][4 * h + 3] = a.w;
        }
        m1[k] = kNoKey; m2[k] = kNoKey;
    }

    for (int ti = 0; ti < ntiles; ++ti) {
        const int stage = ti & 1;
        if (tid == 0 && ti + 1 < ntiles) {
            const int r0n = (ti + 1) * kRows;
            const uint32_t bytes = (uint32_t)min(nt - r0n, kRows) * 64u;
            mbar_expect_tx(&bars[stage ^ 1], bytes);
            bulk_g2s(smem + (stage ^ 1) * kRows * 64, td + (size_t)r0n * 16, bytes, &bars[stage ^ 1]);
        }
        mbar_wait(&bars[stage], (ti >> 1) & 1);
        const uint4* __restrict__ rows = reinterpret_cast<const uint4*>(smem + stage * kRows * 64);
        const int r0 = ti * kRows;
        const int nrows = min(nt - r0, kRows);
        for (int b0 = 0; b0 < nrows; b0 += 64) {
            const int nb = min(nrows - b0, 64);
            const uint4* __restrict__ brows = rows + 4 * b0;
            uint32_t p1 = 0xFFFFFFFFu, p2 = 0xFFFFFFFFu;
            int j = 0;
#pragma unroll 1
            for (; j + 2 <= nb; j += 2) {
                const uint32_t jj = (uint32_t)j * 0x00010001u;
                uint32_t k0, k1;
#if UZ_WIDE_MERGE
                {
                    const uint4 a0 = brows[4 * j], b0v = brows[4 * j + 1], c0 = brows[4 * j + 2], d0 = brows[4 * j + 3];
                    k0 = csa_acc16w_merged<false>(U[0], a0, b0v, c0, d0, jj);
                    k0 = csa_acc16w_merged<true>(U[1], a0, b0v, c0, d0, k0);
                }
                {
                    const uint4 a1 = brows[4 * j + 4], b1v = brows[4 * j + 5], c1 = brows[4 * j + 6], d1 = brows[4 * j + 7];
                    k1 = csa_acc16w_merged<false>(U[0], a1, b1v, c1, d1, jj + 0x00010001u);
                    k1 = csa_acc16w_merged<true>(U[1], a1, b1v, c1, d1, k1);
                }
#else
                {
                    const uint4 a0 = brows[4 * j], b0v = brows[4 * j + 1];
                    const uint4 a1 = brows[4 * j + 4], b1v = brows[4 * j + 5];
                    k0 = csa_acc16w<false>(U[0], a0, b0v, jj);
                    k1 = csa_acc16w<false>(U[0], a1, b1v, jj + 0x00010001u);
                    k0 = csa_acc16w<true>(U[1], a0, b0v, k0);
                    k1 = csa_acc16w<true>(U[1], a1, b1v, k1);
                }
                {
                    const uint4 c0 = brows[4 * j + 2], d0 = brows[4 * j + 3];
                    const uint4 c1 = brows[4 * j + 6], d1 = brows[4 * j + 7];
                    k0 = csa_acc16w<false>(U[0] + 8, c0, d0, k0);
                    k1 = csa_acc16w<false>(U[0] + 8, c1, d1, k1);
                    k0 = csa_acc16w<true>(U[1] + 8, c0, d0, k0);
                    k1 = csa_acc16w<true>(U[1] + 8, c1, d1, k1);
                }
#endif
                top2_update2_u16x2(p1, p2, k0, k1);
                if (XCHK) {
                    col_update16<6>(k0, cq[0], cq[1], s_col_a + 4u * (uint32_t)(b0 + j));
                    col_update16<6>(k1, cq[0], cq[1], s_col_a + 4u * (uint32_t)(b0 + j + 1));
                }
            }
            if (j > 0) {
                const uint32_t base = (uint32_t)(t0 + r0 + b0);
                merge_block16w(m1[0], m2[0], p1 & 0xFFFFu, p2 & 0xFFFFu, base);
                merge_block16w(m1[1], m2[1], p1 >> 16, p2 >> 16, base);
            }
            for (; j < nb; ++j) {
                const uint4 a = brows[4 * j], b = brows[4 * j + 1], c = brows[4 * j + 2], d = brows[4 * j + 3];
                const uint32_t jkey = (uint32_t)(t0 + r0 + b0 + j);
                uint32_t kq[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) { kq[k] = csa_key_wide(U[k], a, b, c, d, jkey); top2_update(m1[k], m2[k], kq[k]); }
                if (XCHK) col_update32<6>(kq[0], kq[1], cq[0], cq[1], s_col_a + 4u * (uint32_t)(b0 + j));
            }
        }
        __syncthreads();
        if (XCHK) {
            col_flush<THREADS, 6>(s_col, nrows, keys + tk->rev_key_off, t0 + r0, tile.y);
            __syncthreads();
        }
    }

    const uint32_t out_off = tk->key_off + (SEG ? (uint32_t)(tile4.w >> 16) * (uint32_t)nq : 0u);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int q = tile.y + k * THREADS + tid;
        if (q < nq) keys[out_off + q] = make_uint2(m1[k], m2[k]);
    }
    if (pair_pending != nullptr) {          // streaming hand-over, as in knn2_kernel
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicSub(pair_pending + tk->pair, 1);
            atomicAdd(progress, 1u);
        }
    }
}

// Segmented launches: fold the partial neighbours of segments 1..S-1 into segment 0 (top-2 of sorted pairs; a missing
// neighbour is the largest key).  table[blockIdx.y] = (key_off, nq, S, 0).
__global__ void __launch_bounds__(256) merge_segments_kernel(const int4* __restrict__ table, uint2* __restrict__ keys) {
    const int4 e = table[blockIdx.y];
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < e.y; q += gridDim.x * blockDim.x) {
        uint2 m = keys[(size_t)e.x + q];
        for (int s = 1; s < e.z; ++s) {
            const uint2 b = keys[(size_t)e.x + (size_t)s * e.y + q];
            const uint32_t hi = max(m.x, b.x);
            m.x = min(m.x, b.x);
            m.y = min(hi, min(m.y, b.y));
        }
        keys[(size_t)e.x + q] = m;
    }
}

// raw descriptor rows (any byte stride) -> packed rows, raw and CSA layout.  A row is `halves` 256-bit halves
// (1: ORB/BRIEF, 2: BRISK/FREAK); one thread per half, n counts halves, every half gets its own CSA transform.
// src and raw may be the SAME buffer (rows already packed: every thread reads its half before it writes it back), so
// neither carries __restrict__.
__global__ void pack_descriptors_kernel(const uint8_t* src, int n, int stride,
                                        uint32_t* raw, uint32_t* __restrict__ csa, int halves) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = halves == 1 ? src + (size_t)i * stride : src + (size_t)(i >> 1) * stride + (i & 1) * 32;
    uint32_t w[8], o[8];
    if ((((uintptr_t)p) & 15) == 0) {
        const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 16);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            w[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) |
                   ((uint32_t)p[4 * k + 3] << 24);
    }
    csa_pack(w, o);
    uint4* r = reinterpret_cast<uint4*>(raw + (size_t)i * 8);
    uint4* c = reinterpret_cast<uint4*>(csa + (size_t)i * 8);
    r[0] = make_uint4(w[0], w[1], w[2], w[3]); r[1] = make_uint4(w[4], w[5], w[6], w[7]);
    c[0] = make_uint4(o[0], o[1], o[2], o[3]); c[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// One launch moves many host buffers: every CTA copies one chunk (<= 16 KB) of a buffer that lives in pinned,
// device-mapped host memory straight over PCIe/NVLink-C2C into HBM.  Replaces thousands of small
// cudaMemcpyAsync calls (2 us of driver time each) when a batch references scattered keyframes.
struct CopyChunk { const uint8_t* src; uint8_t* dst; uint32_t bytes; uint32_t pad; };

__global__ void __launch_bounds__(256) gather_copy_kernel(const CopyChunk* __restrict__ chunks, int n_chunks) {
    // grid-stride over the chunk table: a full-size grid when nothing else runs, a small persistent grid when the
    // copy runs beside the match kernel (a PCIe pull needs ~200 KB in flight, not the whole chip)
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
        const CopyChunk c = chunks[ci];
        const uintptr_t a = (uintptr_t)c.src | (uintptr_t)c.dst;
        if ((a & 15) == 0) {
            const uint4* __restrict__ s = reinterpret_cast<const uint4*>(c.src);
            uint4* __restrict__ d = reinterpret_cast<uint4*>(c.dst);
            const int n = (int)(c.bytes >> 4);
            for (int i = tid; i < n; i += 4 * nth) {          // four independent 16 B loads in flight per thread
                uint4 v0, v1, v2, v3;
                v0 = s[i];
                if (i + nth < n) v1 = s[i + nth];
                if (i + 2 * nth < n) v2 = s[i + 2 * nth];
                if (i + 3 * nth < n) v3 = s[i + 3 * nth];
                d[i] = v0;
                if (i + nth < n) d[i + nth] = v1;
                if (i + 2 * nth < n) d[i + 2 * nth] = v2;
                if (i + 3 * nth < n) d[i + 3 * nth] = v3;
            }
            for (int i = (n << 4) + tid; i < (int)c.bytes; i += nth) c.dst[i] = c.src[i];
        } else if ((a & 7) == 0) {
            const uint2* __restrict__ s = reinterpret_cast<const uint2*>(c.src);
            uint2* __restrict__ d = reinterpret_cast<uint2*>(c.dst);
            const int n = (int)(c.bytes >> 3);
            for (int i = tid; i < n; i += nth) d[i] = s[i];
            for (int i = (n << 3) + tid; i < (int)c.bytes; i += nth) c.dst[i] = c.src[i];
        } else {
            for (int i = tid; i < (int)c.bytes; i += nth) c.dst[i] = c.src[i];
        }
    }
}

// keys -> (idx, dist) int32 pairs for uz_match_knn2
__global__ void unpack_keys_kernel(const uint2* __restrict__ keys, int nq, int32_t* __restrict__ idx,
                                   int32_t* __restrict__ dist) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint2 k = keys[q];
    idx[2 * q] = k.x == kNoKey ? -1 : (int32_t)(k.x & 0xFFFFu);
    dist[2 * q] = k.x == kNoKey ? -1 : (int32_t)(k.x >> 16);
    idx[2 * q + 1] = k.y == kNoKey ? -1 : (int32_t)(k.y & 0xFFFFu);
    dist[2 * q + 1] = k.y == kNoKey ? -1 : (int32_t)(k.y >> 16);
}

// ---- integer-pipe microbenchmarks (roofline denominators) ---------------------------------------
// Each thread runs ITER iterations of 8 independent chains of one op; ops = grid*block*ITER*8.
template <int OP>
__global__ void __launch_bounds__(256) intpipe_bench_kernel(uint32_t* out, uint32_t seed, int iters) {
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u + blockIdx.x;
    const uint32_t c1 = seed | 1u, c2 = seed ^ 0x5bd1e995u;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) {            // POPC (result fed back through the popc input to keep a chain)
                    asm volatile("popc.b32 %0, %0;" : "+r"(v[i]));
                } else if (OP == 1) {     // LOP3
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(c1), "r"(c2));
                } else if (OP == 2) {     // IMAD
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(c1), "r"(c2));
                } else if (OP == 3) {     // VIMNMX
                    asm volatile("min.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(c2 + i + r));
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    if (s == 0x12345678u) out[0] = s;     // keep the chains live
}

}  // namespace uz
