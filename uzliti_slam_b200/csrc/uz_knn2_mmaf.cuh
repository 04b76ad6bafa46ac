// uz_knn2_mmaf.cuh — the tensor-core match kernel on the 4-BIT path: tcgen05.mma kind::mxf4 (e2m1 operands, UE8M0 block scales,
// fp32 accumulate), 128 x 240 x 64 per instruction.  knn2_mmaf_kernel<false>: 32-byte descriptor rows (K = 256, four
// instructions per accumulator); knn2_mmaf_kernel<true>: 64-byte rows (K = 512, eight instructions).
//
// scripts/mxf4_probe.cu: +-4 operands with constant block scales accumulate EXACTLY onto an fp32 accumulator pre-loaded with
// 2^23 + 16384 - every bit - and the instruction takes 141-155 clocks for twice the K of an int8 one.
//
// Operand layout ("E4"): one NIBBLE per descriptor bit (0x6 = +4 set, 0xE = -4 clear), K-major canonical layout with 16-byte
// core-matrix rows: byte(row i, bit k) = (i >> 3) * G + (k >> 5) * 128 + (i & 7) * 16 + ((k >> 1) & 15), low nibble = even k,
// G = 1024 for 32-byte rows and 2048 for 64-byte rows.  Half the bytes of the int8 layouts.
// Keys.  Every accumulator is STARTED by one kind::f8f6f4 instruction (e5m2 x e5m2, K = 32, not accumulating) over two constant
// shared-memory blocks - query side [2048, 2048, 128, 64, 8, 1, 0 ...], train side [2048, 2048, 128, d2, d1, d0, 0 ...] with
// B - 1 - (row & (B - 1)) = 64 d2 + 8 d1 + d0, B = columns per key block - which sets it to 2^23 + 16384 + B - 1 - (column &
// (B - 1)); the mxf4 instructions add U dot on top, exactly.  The low 16 bits of that fp32 value ARE the integer
// TOP - key16, key16 = hamming * B + (column & (B - 1)), so a packed 16-bit TMEM load delivers two keys per register:
//   32-byte rows: B = 128, block scales 2 x 2, U = 64, TOP = 32895 (257 distances x 128 columns fit 16 bits)
//   64-byte rows: B = 64,  block scales 2 x 1, U = 32, TOP = 32831 (513 distances x 64 columns)
// TMEM.  The block scales live in TMEM next to the accumulators, so a train tile is 240 rows, not 256: accumulators at
// columns 0 and 240, scales from 480.
#pragma once
#include "uz_knn2_mmak.cuh"

namespace uz {

constexpr int kF4N = 240;                              // train rows per tile / accumulator columns
constexpr int kF4SfCol = 480;                          // TMEM column of the block scales
constexpr int kF4TailABytes = kMmaM * 32, kF4TailBBytes = kF4N * 32;

template <bool WIDE>
struct F4 {
    static constexpr int kRowBytes = WIDE ? 256 : 128;     // one nibble per descriptor bit
    static constexpr int kGroupBytes = 8 * kRowBytes;
    static constexpr int kABytes = kMmaM * kRowBytes;      // 16 / 32 KB
    static constexpr int kBBytes = kF4N * kRowBytes;       // 30 / 60 KB
    static constexpr int kStages = WIDE ? 2 : 4;           // train tiles in flight
    static constexpr int kABufs = WIDE ? 1 : 2;            // items whose query tiles are in shared memory
    static constexpr int kKSteps = kRowBytes / 32;         // instructions per accumulator (K = 64 nibbles = 32 bytes each)
    static constexpr int kKeyCols = WIDE ? 64 : 128;       // columns per key block
    static constexpr uint32_t kTop = WIDE ? 32831u : 32895u;
    static constexpr uint32_t kScaleA = 0x80808080u;       // UE8M0 2^1
    static constexpr uint32_t kScaleB = WIDE ? 0x7F7F7F7Fu : 0x80808080u;      // 2^0 / 2^1: a product is +-32 / +-64
    static constexpr int kSfbCol = WIDE ? kF4SfCol + 16 : kF4SfCol + 4;
    static constexpr int kSmemBytes = 2 * kABufs * kABytes + kStages * kBBytes + kF4TailABytes + 3 * kF4TailBBytes + 2 * kMmaItemRows * 8 + 256;
    static_assert(kSmemBytes <= 232448, "CTA exceeds the 227 KB of shared memory");
    __host__ __device__ static constexpr size_t bytes(int n) { return (size_t)((n + 7) / 8) * kGroupBytes; }
};
constexpr int kF4RowBytes = F4<false>::kRowBytes, kF4GroupBytes = F4<false>::kGroupBytes, kF4BBytes = F4<false>::kBBytes;
constexpr int kF4SmemBytes = F4<false>::kSmemBytes;

constexpr uint32_t kE4Set = 0x6u, kE4Clear = 0xEu;      // e2m1: +4 for a set bit, -4 for a clear one
__host__ __device__ constexpr size_t e4_bytes(int n) { return F4<false>::bytes(n); }
__host__ __device__ constexpr size_t e4w_bytes(int n) { return F4<true>::bytes(n); }
// The rows [n, round_up(n, 8)) of the last 8-row group hold zero nibbles (value +0), and a ragged train tile is completed to
// 128 or 240 rows from a page of zeros: a missing row contributes nothing to any sum.
constexpr size_t kF4ZeroPageBytes = F4<true>::kBBytes;
// train rows the instructions of a tile with `rows` real rows cover
__host__ __device__ constexpr int f4_tile_cols(int rows) { return rows <= 128 ? 128 : kF4N; }

// bits -> E4 layout (ingestion): one thread per (row, 32-bit word): 32 nibbles = one uint4 store
__global__ void __launch_bounds__(256) expand_e4_kernel(const uint32_t* __restrict__ raw, int n, uint8_t* __restrict__ e4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((n + 7) & ~7) * 8) return;
    const int row = i >> 3, w = i & 7;
    if (row >= n) { *reinterpret_cast<uint4*>(e4 + (size_t)(row >> 3) * kF4GroupBytes + w * 128 + (row & 7) * 16) = make_uint4(0u, 0u, 0u, 0u); return; }
    const uint32_t bits = raw[(size_t)row * 8 + w];
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) v |= (((bits >> (8 * k + b)) & 1u) ? kE4Set : kE4Clear) << (4 * b);
        o[k] = v;
    }
    *reinterpret_cast<uint4*>(e4 + (size_t)(row >> 3) * kF4GroupBytes + w * 128 + (row & 7) * 16) = make_uint4(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ void tc_mma_mxf4(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t sfa, uint32_t sfb) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, 1, 0;\n"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%4], [%5], p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(sfa), "r"(sfb) : "memory");
}
// 32 lanes x 32 columns, every cell the same value (the block scales)
__device__ __forceinline__ void f4_fill32(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
                 ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void f4_fill16(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
                 ::"r"(taddr), "r"(v) : "memory");
}
// the instruction that starts an accumulator: kind::f8f6f4, K = 32, D = A B (no accumulate)
__device__ __forceinline__ void tc_mma_f8_start(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, 0, 0;\n"                  // false: D = A B, the accumulator is not read
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%4, %4, %4, %4}, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0u) : "memory");
}
// A whole 128-column block (dA: 64 columns, dB: the first 2 * REGSB columns of the next 64) in two sweeps instead of the
// five-operation insertion network, because the epilogue is bound by the ALU pipe (every packed min / max, 2 clocks each):
//   sweep 1  P = the packed maximum of the block, three-input max                                  0.5 ALU operations / register
//   sweep 2  y = x - P + 65536 as ONE 32-bit multiply-add (fma pipe), then the packed maximum of y   0.5 ALU + 1 FMA / register
// In 16-bit lanes y is (x - P) mod 2^16: the winner becomes the smallest value, everybody else keeps its order just below
// 2^16, so max(y) is the runner-up.  The 32-bit form is exact for the low lane; the low lane's borrow reaches the high lane
// in every register except the one holding the low lane's winner, which the +65536 pre-pays: there the high lane is one too
// large.  High-lane keys are those of odd columns and 32895 is odd, so they are all even and two of them differ by at
// least 2: the stray +1 can neither reorder them nor wrap, and is masked off when the runner-up is rebuilt.
__device__ __forceinline__ uint32_t f4_mad(uint32_t a, uint32_t one, uint32_t c) {
    uint32_t d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c)); return d;
}
template <int REGSB>
__device__ __forceinline__ void f4_block(const uint32_t (&dA)[32], const uint32_t (&dB)[32], uint32_t one,
                                         uint32_t& m1, uint32_t& m2, uint32_t tbase) {
    uint32_t g0 = 0u, g1 = 0u;
#pragma unroll
    for (int m = 0; m < 32; m += 4) {
        g0 = max_u16x2(max_u16x2(g0, dA[m]), dA[m + 1]);
        g1 = max_u16x2(max_u16x2(g1, dA[m + 2]), dA[m + 3]);
    }
#pragma unroll
    for (int m = 0; m < REGSB; m += 4) {
        g0 = max_u16x2(max_u16x2(g0, dB[m]), dB[m + 1]);
        g1 = max_u16x2(max_u16x2(g1, dB[m + 2]), dB[m + 3]);
    }
    uint32_t P = max_u16x2(g0, g1);
    const uint32_t c = 65536u - P;
    uint32_t y0 = 0u, y1 = 0u;
#pragma unroll
    for (int m = 0; m < 32; m += 4) {
        y0 = max_u16x2(max_u16x2(y0, f4_mad(dA[m], one, c)), f4_mad(dA[m + 1], one, c));
        y1 = max_u16x2(max_u16x2(y1, f4_mad(dA[m + 2], one, c)), f4_mad(dA[m + 3], one, c));
    }
#pragma unroll
    for (int m = 0; m < REGSB; m += 4) {
        y0 = max_u16x2(max_u16x2(y0, f4_mad(dB[m], one, c)), f4_mad(dB[m + 1], one, c));
        y1 = max_u16x2(max_u16x2(y1, f4_mad(dB[m + 2], one, c)), f4_mad(dB[m + 3], one, c));
    }
    uint32_t S = (max_u16x2(y0, y1) - c) & 0xFFFEFFFFu;
    mmak_merge(m1, m2, P, S, tbase);
}
// 64-byte rows: a 64-column block per array.  key16 = hamming * 64 + (column & 63) <= 32831 = TOP; odd columns again give even
// high-lane values (TOP is odd), so the second sweep works unchanged.
template <int REGS>
__device__ __forceinline__ void f4_block64(const uint32_t (&d)[32], uint32_t one, uint32_t& m1, uint32_t& m2, uint32_t tbase) {
    uint32_t g0 = 0u, g1 = 0u;
#pragma unroll
    for (int m = 0; m < REGS; m += 4) {
        g0 = max_u16x2(max_u16x2(g0, d[m]), d[m + 1]);
        g1 = max_u16x2(max_u16x2(g1, d[m + 2]), d[m + 3]);
    }
    const uint32_t P = max_u16x2(g0, g1);
    const uint32_t c = 65536u - P;
    uint32_t y0 = 0u, y1 = 0u;
#pragma unroll
    for (int m = 0; m < REGS; m += 4) {
        y0 = max_u16x2(max_u16x2(y0, f4_mad(d[m], one, c)), f4_mad(d[m + 1], one, c));
        y1 = max_u16x2(max_u16x2(y1, f4_mad(d[m + 2], one, c)), f4_mad(d[m + 3], one, c));
    }
    const uint32_t S = (max_u16x2(y0, y1) - c) & 0xFFFEFFFFu;
    // the four winners (even and odd columns, best and second) -> (distance << 16) | train row
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t v1 = F4<true>::kTop - (h ? P >> 16 : P & 0xFFFFu), v2 = F4<true>::kTop - (h ? S >> 16 : S & 0xFFFFu);
        const uint32_t k1 = ((v1 & 0xFFC0u) << 10) | (tbase + (v1 & 63u));
        const uint32_t k2 = ((v2 & 0xFFC0u) << 10) | (tbase + (v2 & 63u));
        const uint32_t t = max(m1, k1);
        m1 = min(m1, k1);
        m2 = min(min(m2, t), k2);
    }
}
// 20 warps: the producer, two MMA issuers, the writer of the ragged tiles' start operands, and two groups of eight epilogue
// warps, one group per accumulator, so that the load latency of one accumulator hides behind the sweeps of the other.  Five
// warps share a sub-partition's 16384 registers: 96 each.
//
// Ragged train tiles (the last of a matching: `rows` < 240 real rows) are made whole instead of being masked in the epilogue:
// the producer completes the operand tile to 128 or 240 rows from a page of zeros, and warp 3 writes a copy of the start
// operand whose rows >= `rows` are [2048, 2048, 0, ...]: the accumulator of such a column is exactly 2^23, its key 0.  A real
// key of a block with missing columns is never 0 (that is distance 256 in column 127), and a lane with fewer than two real
// columns can only surface a key-0 "row" whose index is >= the number of train rows: it loses every tie against real rows and
// is dropped when the keys are published.  The epilogue therefore knows one path.
#ifdef UZ_F4_TRACE
__device__ long long g_f4_trace[2][128][8];      // probe builds: CTA 0's first 128 accumulators, issuer and epilogue clocks
#endif
#ifndef F4_EPI_GROUPS
#define F4_EPI_GROUPS 2
#endif
constexpr int kF4EpiGroups = F4_EPI_GROUPS;   // 2: eight epilogue warps per accumulator; 1: eight warps serve both in turn
constexpr int kF4Threads = (4 + 8 * kF4EpiGroups) * 32;
#ifndef F4_ISSUERS
#define F4_ISSUERS 2
#endif
constexpr int kF4Issuers = F4_ISSUERS;        // 1: warp 1 issues for both accumulators; 2: warp 1 for the first, warp 2 for the second
constexpr int kF4MaxRegs = kF4EpiGroups == 2 ? 96 : 128;

// What a role needs to know about an item, fetched one item ahead (and the item's (task, row) pair two ahead): the two
// dependent global loads at the top of an item cost every role ~1000 clocks otherwise, with the tensor pipe idle behind them.
struct F4Item { const uint8_t* q; const uint8_t* t; int nq, nt, q0; uint32_t key_off; int pair; };
__device__ __forceinline__ int2 f4_item_id(const int2* __restrict__ items, int it, int n_items) {
    return it < n_items ? items[it] : make_int2(-1, 0);
}
__device__ __forceinline__ F4Item f4_item(const MmaTask* __restrict__ tasks, int2 id) {
    F4Item m = {nullptr, nullptr, 0, 0, 0, 0u, 0};
    if (id.x >= 0) {
        const MmaTask* tk = tasks + id.x;
        m.q = mma_q(tk); m.t = mma_t(tk); m.nq = tk->nq; m.nt = tk->nt; m.q0 = id.y; m.key_off = tk->key_off; m.pair = tk->pair;
    }
    return m;
}
#define F4_ITEM_LOOP_BEGIN                                                                                   \
    int2 id1 = f4_item_id(items, blockIdx.x + gridDim.x, n_items);                                           \
    F4Item cur = f4_item(tasks, f4_item_id(items, blockIdx.x, n_items));                                     \
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {                                               \
        const F4Item nxt = f4_item(tasks, id1);                                                              \
        const int2 id2 = f4_item_id(items, it + 2 * gridDim.x, n_items);                                     \
        const int nq = cur.nq, nt = cur.nt, q0 = cur.q0;
#define F4_ITEM_LOOP_END                                                                                     \
        cur = nxt; id1 = id2;                                                                                \
    }

// items[k] = (task, first query row); CTA b takes items b, b + gridDim.x, ...
template <bool WIDE>
__global__ void __maxnreg__(kF4MaxRegs) knn2_mmaf_kernel(const MmaTask* __restrict__ tasks, const int2* __restrict__ items,
                                                                  int n_items, uint2* __restrict__ keys, MmaDesc dsc,
                                                                  int* __restrict__ pair_pending, unsigned int* __restrict__ progress,
                                                                  const uint8_t* __restrict__ zero_page) {
    using C = F4<WIDE>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;                                    // [kABufs items][2][16 / 32 KB]  query tiles, 128 rows of nibbles; with two
                                                           // buffers the next item's are fetched while this one's are in use
    uint8_t* sB = smem + 2 * C::kABufs * C::kABytes;       // [kStages][30 / 60 KB]  train tiles, 240 rows
    uint8_t* sTailA = sB + C::kStages * C::kBBytes;        // [128 rows x 32 B] e5m2, the instruction that starts an accumulator
    uint8_t* sTailB = sTailA + kF4TailABytes;               // [240 rows x 32 B]
    uint8_t* sTailR = sTailB + kF4TailBBytes;               // [2][240 rows x 32 B]  start operands of ragged tiles
    uint2* xchg = reinterpret_cast<uint2*>(sTailR + 2 * kF4TailBBytes);     // [2 parities][256 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + 2 * kMmaItemRows);
    uint64_t* a_full = bars;          // [2 items][2]  (eight slots reserved)
    uint64_t* b_full = bars + 8;      // [C::kStages]
    uint64_t* b_empty = b_full + C::kStages;
    uint64_t* acc_full = b_empty + C::kStages;    // [2]
    uint64_t* acc_empty = acc_full + 2;          // [2]
    uint64_t* rag_full = acc_empty + 2;          // [2]
    uint64_t* rag_empty = rag_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rag_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&a_full[i], 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); mbar_init(&rag_full[i], 1); mbar_init(&rag_empty[i], kF4Issuers); }
        for (int s = 0; s < C::kStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], kF4Issuers); }       // every issuer returns a train tile
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the constant operands of the starting instruction (e5m2 bytes; compact canonical layout, 8-row groups of 256 B)
    for (int r = tid; r < kMmaM + kF4N; r += kF4Threads) {
        const bool isB = r >= kMmaM;
        const int row = isB ? r - kMmaM : r;
        const int v = (C::kKeyCols - 1) - (row & (C::kKeyCols - 1));
        const uint32_t dig[8] = {0x00u, 0x3Cu, 0x40u, 0x42u, 0x44u, 0x45u, 0x46u, 0x47u};      // e5m2 of 0..7
        const uint32_t w0 = 0x68u | (0x68u << 8) | (0x58u << 16) | ((isB ? dig[v >> 6] : 0x54u) << 24);     // 2048, 2048, 128, d2 | 64
        const uint32_t w1 = isB ? (dig[(v >> 3) & 7] | (dig[v & 7] << 8)) : (0x48u | (0x3Cu << 8));        // d1, d0 | 8, 1
        uint8_t* dst = (isB ? sTailB : sTailA) + (row >> 3) * 256 + (row & 7) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(w0, w1, 0u, 0u);
        *reinterpret_cast<uint4*>(dst + 128) = make_uint4(0u, 0u, 0u, 0u);
        if (isB) {           // the upper K half of the ragged copies never changes
            *reinterpret_cast<uint4*>(sTailR + (row >> 3) * 256 + (row & 7) * 16 + 128) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(sTailR + kF4TailBBytes + (row >> 3) * 256 + (row & 7) * 16 + 128) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {          // the whole TMEM: two 240-column accumulators, the block scales behind them
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= 4 && warp < 8) {
        // 32-byte rows: every block scale is 2^1 (UE8M0 128): operands +-4 count as +-8, a product is +-64
        if (WIDE) {          // 64-byte rows: query-side scales 2^1, train-side scales 2^0 (a product is +-32), sixteen columns each
            f4_fill16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kF4SfCol, C::kScaleA);
            f4_fill16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C::kSfbCol, C::kScaleB);
        } else {
            f4_fill32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kF4SfCol, C::kScaleA);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        // ===================== producer =====================
        if (lane == 0) {
            // Query tiles have no "empty" barrier of their own: an item's query tiles are free when the instructions on its LAST
            // train tile are done, which is what that tile's stage barrier says (both issuers arrive on it) - one
            // tcgen05.commit per item and issuer less.  a_last[b] = that tile's running number for query-tile buffer b.
            uint32_t uB = 0, k = 0, a_last0 = 0, a_last1 = 0, a_used = 0;
            F4_ITEM_LOOP_BEGIN
                const int nqt = (nq - q0 > kMmaM) ? 2 : 1;
                const int T = (nt + kF4N - 1) / kF4N;
                const uint32_t ab = C::kABufs == 2 ? (k & 1u) * 2u : 0u;
                for (int t = 0; t < T; ++t) {
                    if (t == 0) {
                        if (a_used & (1u << (ab >> 1))) {
                            const uint32_t u = ab ? a_last1 : a_last0;
                            // (once the stage has been refilled - tile u + kStages is behind us -, the wait before that refill
                            // has covered this already, and asking again could alias with a later phase of the barrier)
                            if (uB <= u + C::kStages) mbar_wait_wd(&b_empty[u % C::kStages], (u / C::kStages) & 1u);
                        }
                        const uint32_t bytes = (uint32_t)C::bytes(min(kMmaM, nq - q0));
                        mbar_expect_tx(&a_full[ab], bytes);
                        bulk_g2s(sA + ab * C::kABytes, cur.q + (size_t)(q0 >> 3) * C::kGroupBytes, bytes, &a_full[ab]);
                    }
                    {
                        const uint32_t slot = uB % C::kStages;
                        mbar_wait_wd(&b_empty[slot], ((uB / C::kStages) & 1u) ^ 1u);
                        const int rows = min(kF4N, nt - t * kF4N);
                        const uint32_t bytes = (uint32_t)C::bytes(rows), fill = (uint32_t)C::bytes(f4_tile_cols(rows)) - bytes;
                        mbar_expect_tx(&b_full[slot], bytes + fill);
                        bulk_g2s(sB + slot * C::kBBytes, cur.t + (size_t)t * C::kBBytes, bytes, &b_full[slot]);
                        if (fill) bulk_g2s(sB + slot * C::kBBytes + bytes, zero_page, fill, &b_full[slot]);
                        uB++;
                    }
                    if (t == 0 && nqt == 2) {
                        const uint32_t bytes = (uint32_t)C::bytes(min(kMmaM, nq - q0 - kMmaM));
                        mbar_expect_tx(&a_full[ab + 1], bytes);
                        bulk_g2s(sA + (ab + 1) * C::kABytes, cur.q + (size_t)((q0 + kMmaM) >> 3) * C::kGroupBytes, bytes, &a_full[ab + 1]);
                    }
                }
                if (T > 0) {
                    if (ab) a_last1 = uB - 1; else a_last0 = uB - 1;
                    a_used |= 1u << (ab >> 1);
                }
                ++k;
            F4_ITEM_LOOP_END
        }
    } else if (warp >= 1 && warp <= kF4Issuers) {
        // ===================== MMA issuer(s) =====================
        // The five instructions of an accumulator and their commits are issued by ONE ELECTED lane of the converged warp
        // (elect.sync): ptxas then keeps every descriptor in uniform registers and emits the tcgen05 instructions back to
        // back.  Behind a `lane == 0` test it wraps each of them in a per-thread loop with R2UR moves, ~70 clocks apiece
        // next to busy epilogue warps - as long as the instruction runs - and the issue, not the tensor pipe, bounds the kernel.
        const int i_first = kF4Issuers == 2 ? warp - 1 : 0, i_end = kF4Issuers == 2 ? warp : 2;
        uint32_t uB = 0, phA = 0, k = 0, nrag = 0, uAcc[2] = {0, 0};
#ifdef UZ_F4_TRACE
        int tr = 0;
#endif
        MmaDesc fdsc = dsc;
        fdsc.lbo16 = 128 >> 4; fdsc.sbo16 = C::kGroupBytes >> 4;
        const uint32_t sfa = tmem_base + kF4SfCol, sfb = tmem_base + C::kSfbCol;
        MmaDesc tdsc = dsc;
        tdsc.lbo16 = 128 >> 4; tdsc.sbo16 = 256 >> 4;
        const uint64_t tail_a = make_smem_desc(smem_u32(sTailA), tdsc), tail_b = make_smem_desc(smem_u32(sTailB), tdsc);
        F4_ITEM_LOOP_BEGIN
            const int nqt = (nq - q0 > kMmaM) ? 2 : 1;
            const int T = (nt + kF4N - 1) / kF4N;
            const uint32_t ab = C::kABufs == 2 ? (k & 1u) * 2u : 0u;
            for (int t = 0; t < T; ++t) {
                const uint32_t slot = uB % C::kStages;
#ifdef UZ_F4_TRACE
                const long long trA = clock64();
#endif
                mbar_wait_wd(&b_full[slot], (uB / C::kStages) & 1u);
#ifdef UZ_F4_TRACE
                const long long trB = clock64();
#endif
                const int rows = min(kF4N, nt - t * kF4N);
                const uint32_t n_mma = (uint32_t)f4_tile_cols(rows);
                // a ragged tile starts from its own copy of the start operand (written by warp 3)
                const bool ragged = rows < kF4N;
                const uint32_t rb = nrag & 1u;
                if (ragged) mbar_wait_wd(&rag_full[rb], (nrag >> 1) & 1u);
                const uint64_t tail_bt = ragged ? make_smem_desc(smem_u32(sTailR + rb * kF4TailBBytes), tdsc) : tail_b;
                // a, b format E2M1 (MXF4Format 1) | K-major | N >> 3 at [17,23) | scale format UE8M0 at 23 | M >> 4 at [24,29)
                const uint32_t idesc = (1u << 7) | (1u << 10) | ((n_mma >> 3) << 17) | (1u << 23) | ((uint32_t)(kMmaM >> 4) << 24);
                // the starting instruction: c format F32 | a, b format E5M2 | K-major | N | M
                const uint32_t idesc_start = (1u << 4) | (1u << 7) | (1u << 10) | ((n_mma >> 3) << 17) | ((uint32_t)(kMmaM >> 4) << 24);
                bool issued = false;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (i >= i_first && i < i_end && i < nqt) {
                        if (t == 0) mbar_wait_wd(&a_full[ab + i], (phA >> (ab + i)) & 1u);
                        mbar_wait_wd(&acc_empty[i], (uAcc[i] & 1u) ^ 1u);
                        tc_fence_after();
#ifdef UZ_F4_TRACE
                        const long long tr0 = clock64();
#endif
                        if (tc_elect_one()) {
                            const uint32_t a_addr = smem_u32(sA + (ab + i) * C::kABytes), b_addr = smem_u32(sB + slot * C::kBBytes);
                            const uint32_t d_acc = tmem_base + (uint32_t)i * kF4N;
                            tc_mma_f8_start(d_acc, tail_a, tail_bt, idesc_start);
#pragma unroll
                            for (int k = 0; k < C::kKSteps; ++k)       // K = 64 nibbles = 32 bytes of a row per instruction, on top of the start value
                                tc_mma_mxf4(d_acc, make_smem_desc(a_addr + k * 256, fdsc), make_smem_desc(b_addr + k * 256, fdsc), idesc, sfa, sfb);
                            tc_commit(&acc_full[i]);
                        }
#ifdef UZ_F4_TRACE
                        if (lane == 0 && blockIdx.x == 0 && 2 * tr + i < 128) {
                            long long* row = g_f4_trace[0][2 * tr + i];
                            row[0] = tr0; row[1] = clock64(); row[2] = trA; row[3] = trB;
                        }
#endif
                        __syncwarp();
                        uAcc[i]++;
                        issued = true;
                    }
                }
                // the train tile (and a ragged tile's start operand) goes back when this issuer's instructions on it are done
                // (an issuer with nothing to do on an item of one query tile returns its share at once)
                if (tc_elect_one()) {
                    if (issued) tc_commit(&b_empty[slot]); else mbar_arrive(&b_empty[slot]);
                    if (ragged) { if (issued) tc_commit(&rag_empty[rb]); else mbar_arrive(&rag_empty[rb]); }
                }
                __syncwarp();
                if (ragged) ++nrag;
#ifdef UZ_F4_TRACE
                ++tr;
#endif
                uB++;
            }
            if (T > 0) for (int i = 0; i < nqt; ++i) phA ^= 1u << (ab + i);
            ++k;
        F4_ITEM_LOOP_END
    } else if (warp == 3) {
        // ===================== start operands of the ragged tiles =====================
        uint32_t nrag = 0;
        F4_ITEM_LOOP_BEGIN
            (void)nq; (void)q0;
            const int rows = nt % kF4N;          // real rows of the last tile, if it is ragged
            if (rows != 0) {
            const uint32_t rb = nrag & 1u;
            mbar_wait_wd(&rag_empty[rb], ((nrag >> 1) & 1u) ^ 1u);
            const uint32_t dig[8] = {0x00u, 0x3Cu, 0x40u, 0x42u, 0x44u, 0x45u, 0x46u, 0x47u};      // e5m2 of 0..7
            for (int row = lane; row < kF4N; row += 32) {
                const int v = (C::kKeyCols - 1) - (row & (C::kKeyCols - 1));
                const bool real = row < rows;
                const uint32_t w0 = 0x68u | (0x68u << 8) | (real ? (0x58u << 16) | (dig[v >> 6] << 24) : 0u);       // 2048, 2048, 128 | 0, d2 | 0
                const uint32_t w1 = real ? (dig[(v >> 3) & 7] | (dig[v & 7] << 8)) : 0u;
                *reinterpret_cast<uint4*>(sTailR + rb * kF4TailBBytes + (row >> 3) * 256 + (row & 7) * 16) = make_uint4(w0, w1, 0u, 0u);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&rag_full[rb]);
            ++nrag;
            }
        F4_ITEM_LOOP_END
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;                   // 0 .. 8 * kF4EpiGroups - 1
        const int quarter = warp & 3;              // the TMEM lanes this warp may touch: 32 * (warp % 4) ..
        const int half = (ew >> 2) & 1;            // which 128 columns of the accumulator
        // which accumulators (query tiles of the item) this group of eight warps serves
        const int i_first = kF4EpiGroups == 2 ? ew >> 3 : 0, i_end = kF4EpiGroups == 2 ? i_first + 1 : 2;
        const int row_in_tile = quarter * 32 + lane;
        uint32_t uAcc[2] = {0, 0};
        uint32_t item_parity = 0;
#ifdef UZ_F4_TRACE
        int tr = 0;
#endif
        const uint32_t one = 1u + (uint32_t)(n_items < 0);        // 1, but not a constant ptxas could fold the multiply-add with
        F4_ITEM_LOOP_BEGIN
            const int nqt = (nq - q0 > kMmaM) ? 2 : 1;
            const int T = (nt + kF4N - 1) / kF4N;
            if (i_first < nqt) {                   // (a group has nothing to do on an item of one query tile)
            uint32_t m1[2] = {kNoKey, kNoKey}, m2[2] = {kNoKey, kNoKey};
            for (int t = 0; t < T; ++t) {
                const bool active = half == 0 || nt - t * kF4N > 128;        // a tile of <= 128 rows has no second half
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (i >= i_first && i < i_end && i < nqt) {
                        mbar_wait_wd(&acc_full[i], uAcc[i] & 1u);
                        tc_fence_after();
#ifdef UZ_F4_TRACE
                        const long long tr0 = clock64();
                        long long tr1 = 0;
#endif
                        if (active) {
                            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(i * kF4N + half * 128);
                            const uint32_t tbase = (uint32_t)(t * kF4N + half * 128);        // global train row of column 0
                            // low 16 bits of the fp32 accumulator = 32895 - ((hamming << 7) | (column & 127)): the key of
                            // knn2_mmak_kernel; 0 in the columns a ragged tile does not have
                            uint32_t dA[32], dB[32];
                            tc_ld64p(taddr, dA);
                            tc_ld64p(taddr + 64, dB);            // (half 1: its last 16 columns belong to nobody and are not looked at)
                            tc_wait_ld(); tc_pin(dA); tc_pin(dB);
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&acc_empty[i]);
#ifdef UZ_F4_TRACE
                            tr1 = clock64();
#endif
                            if (WIDE) {
                                f4_block64<32>(dA, one, m1[i], m2[i], tbase);
                                if (half == 0) f4_block64<32>(dB, one, m1[i], m2[i], tbase + 64); else f4_block64<24>(dB, one, m1[i], m2[i], tbase + 64);
                            } else {
                                if (half == 0) f4_block<32>(dA, dB, one, m1[i], m2[i], tbase); else f4_block<24>(dA, dB, one, m1[i], m2[i], tbase);
                            }
                        } else {
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&acc_empty[i]);
                        }
#ifdef UZ_F4_TRACE
                        if (lane == 0 && blockIdx.x == 0 && (ew & 7) == 0 && 2 * tr + i < 128) {
                            long long* row = g_f4_trace[1][2 * tr + i];
                            row[0] = tr0; row[1] = tr1; row[2] = clock64();
                        }
#endif
                        uAcc[i]++;
                    }
                }
#ifdef UZ_F4_TRACE
                ++tr;
#endif
            }
            // fold the two column halves of every row (half 1 -> shared memory -> half 0) and publish the keys
            uint2* xc = xchg + item_parity * kMmaItemRows;
            if (half == 1) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    if (i >= i_first && i < i_end && i < nqt) xc[i * kMmaM + row_in_tile] = make_uint2(m1[i], m2[i]);
            }
            if (i_first == 0) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 2, 256;" ::: "memory");
            if (half == 0) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int q = q0 + i * kMmaM + row_in_tile;
                    if (i >= i_first && i < i_end && i < nqt && q < nq) {
                        const uint2 o = xc[i * kMmaM + row_in_tile];
                        const uint32_t hi = max(m1[i], o.x);
                        const uint32_t a = min(m1[i], o.x);
                        uint32_t b = min(hi, min(m2[i], o.y));
                        if ((b & 0xFFFFu) >= (uint32_t)nt) b = kNoKey;          // a key-0 column of a ragged tile (one train row in all)
                        keys[(size_t)cur.key_off + q] = make_uint2(a, b);
                    }
                }
            }
            if (pair_pending != nullptr && half == 0) {            // streaming hand-over, as in knn2_kernel (one count per item)
                // the half-0 warps of every group at work on this item: their key stores are done
                if (kF4EpiGroups == 2 && nqt == 2) asm volatile("bar.sync 3, 256;" ::: "memory"); else asm volatile("bar.sync 4, 128;" ::: "memory");
                if (i_first == 0 && row_in_tile == 0) {
                    __threadfence();
                    atomicSub(pair_pending + cur.pair, 1);
                    atomicAdd(progress, 1u);
                }
            }
            }
            item_parity ^= 1u;
        F4_ITEM_LOOP_END
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


}  // namespace uz
