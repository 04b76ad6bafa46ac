// uz_internal.cuh — host-side state shared by the translation unit's parts (uz_capi.cu includes uz_upload.inl,
// uz_batch.inl, uz_capi_places.inl, uz_capi_ingest.inl, uz_group.inl): memory helpers, the keyframe store's records and
// the context.  Nothing here is part of the C-ABI (include/uzliti_edge.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/uzliti_edge.h"
#include "uz_ingest.cuh"
#include "uz_knn2.cuh"
#include "uz_knn2_mma.cuh"
#include "uz_knn2_mma2.cuh"
#include "uz_knn2_mmak.cuh"
#include "uz_knn2_mmaf.cuh"
#include "uz_knn2_mmaw.cuh"
#include "uz_places.cuh"
#include "uz_samples.h"
#include "uz_solve.cuh"

using namespace uz;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// Bump pools for per-call scratch that several in-flight copies/kernels read (copy-chunk tables): blocks stay
// alive until reset(), which the entry points call only when the stream is known to be idle.
template <bool PINNED_HOST>
struct BumpPool {
    struct Block { uint8_t* base; size_t size, used; };
    std::vector<Block> blocks;
    void* alloc(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        for (auto& b : blocks)
            if (b.size - b.used >= bytes) { void* r = b.base + b.used; b.used += bytes; return r; }
        const size_t sz = std::max(bytes, (size_t)1 << 20);
        void* p = nullptr;
        const cudaError_t e = PINNED_HOST ? cudaMallocHost(&p, sz) : cudaMalloc(&p, sz);
        if (e != cudaSuccess) { cudaGetLastError(); return nullptr; }
        blocks.push_back(Block{(uint8_t*)p, sz, bytes});
        return p;
    }
    void reset() { for (auto& b : blocks) b.used = 0; }
    void release() {
        for (auto& b : blocks) { if (PINNED_HOST) cudaFreeHost(b.base); else cudaFree(b.base); }
        blocks.clear();
    }
};

// Device allocator of the keyframe store and of per-call transients: large cudaMalloc'ed chunks (HBM3e: 180 GB, chunks
// grow geometrically), first-fit over per-chunk free maps with coalescing.  A keyframe owns ONE range (all layouts of all
// its cameras), so uz_store_remove gives back exactly what uz_store_add took (the reference removes and merges nodes for
// as long as it runs: graph_slam_node.cpp:641,665-777,1050).  reset() frees everything at once (transients, uz_store_clear).
struct Arena {
    struct Chunk { uint8_t* base; size_t size; std::map<size_t, size_t> free; /* offset -> length */ };
    std::vector<Chunk> chunks;
    size_t chunk_bytes = (size_t)64 << 20;
    size_t total = 0, used = 0;
    static size_t round(size_t b) { return b == 0 ? 256 : (b + 255) & ~(size_t)255; }
    void* alloc(size_t bytes) {
        bytes = round(bytes);
        for (auto& c : chunks)
            for (auto it = c.free.begin(); it != c.free.end(); ++it)
                if (it->second >= bytes) {
                    const size_t off = it->first, len = it->second;
                    c.free.erase(it);
                    if (len > bytes) c.free.emplace(off + bytes, len - bytes);
                    used += bytes;
                    return c.base + off;
                }
        size_t sz = std::max(std::max(chunk_bytes, bytes), std::min(total / 2, (size_t)4 << 30));
        sz = (sz + 255) & ~(size_t)255;
        void* p = nullptr;
        if (cudaMalloc(&p, sz) != cudaSuccess) {
            cudaGetLastError();
            sz = round(bytes);
            if (cudaMalloc(&p, sz) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        }
        chunks.push_back(Chunk{(uint8_t*)p, sz, {}});
        if (sz > bytes) chunks.back().free.emplace(bytes, sz - bytes);
        total += sz; used += bytes;
        return p;
    }
    void free(void* ptr, size_t bytes) {
        if (!ptr) return;
        bytes = round(bytes);
        for (auto& c : chunks) {
            if ((uint8_t*)ptr < c.base || (uint8_t*)ptr >= c.base + c.size) continue;
            size_t off = (size_t)((uint8_t*)ptr - c.base), len = bytes;
            auto nx = c.free.lower_bound(off);
            if (nx != c.free.begin()) {
                auto pv = std::prev(nx);
                if (pv->first + pv->second == off) { off = pv->first; len += pv->second; c.free.erase(pv); }
            }
            if (nx != c.free.end() && nx->first == off + len) { len += nx->second; c.free.erase(nx); }
            c.free.emplace(off, len);
            used -= bytes;
            return;
        }
    }
    void reset() { for (auto& c : chunks) { c.free.clear(); c.free.emplace(0, c.size); } used = 0; }
    void release() { for (auto& c : chunks) cudaFree(c.base); chunks.clear(); total = 0; used = 0; }
};

// A few persistent host threads for packing pageable buffers into the pinned staging ring (a single memcpy stream tops out
// near 10 GB/s; spawning threads per group costs more than a small group's copy).
struct HostPool {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv_go, cv_done;
    std::function<void(unsigned)> job;
    unsigned generation = 0, pending = 0;
    bool quit = false;
    void start(unsigned n) {
        for (unsigned i = 0; i < n; ++i)
            th.emplace_back([this, i]() {
                unsigned seen = 0;
                std::unique_lock<std::mutex> lk(m);
                for (;;) {
                    cv_go.wait(lk, [&] { return quit || generation != seen; });
                    if (quit) return;
                    seen = generation;
                    auto fn = job;
                    lk.unlock();
                    fn(i + 1);
                    lk.lock();
                    if (--pending == 0) cv_done.notify_all();
                }
            });
    }
    // runs fn(0) here and fn(1..n) on the workers; returns when all are through
    void run(const std::function<void(unsigned)>& fn) {
        {
            std::lock_guard<std::mutex> lk(m);
            job = fn; pending = (unsigned)th.size(); ++generation;
        }
        cv_go.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
    void stop() {
        { std::lock_guard<std::mutex> lk(m); quit = true; }
        cv_go.notify_all();
        for (auto& t : th) if (t.joinable()) t.join();
        th.clear();
    }
};

struct Cam {
    uint32_t* raw = nullptr;   // n x dbytes/4 words, bytes as given
    uint32_t* csa = nullptr;   // same rows, every 256-bit half in CSA layout (uz_knn2.cuh)
    uint8_t* e8 = nullptr;     // the tensor-core operand layout: one 4-bit value per bit (uz_knn2_mmaf.cuh); when the context runs an int8
                               // kernel, one int8 per bit (32-byte rows: uz_knn2_mma.cuh; 64-byte rows: two planes, uz_knn2_mmaw.cuh)
    double* pos = nullptr;     // 3 x n column-major
    uint8_t* valid = nullptr;  // n
    int32_t n = 0, feature_type = 0, sensor_frame = 0;
    int32_t dbytes = UZ_DESC_BYTES;   // descriptor width: 32 or 64
};

// where the layouts of one camera live inside a block (offsets are multiples of 256)
struct CamLayout { size_t raw, csa, e8, pos, valid, end; };
// fmt: bit 0 = 32-byte rows keep the 4-bit layout, bit 1 = 64-byte rows do (else the int8 layouts)
inline size_t operand_bytes(int n, int dbytes, int fmt) {
    return dbytes == UZ_DESC_BYTES ? ((fmt & 1) ? e4_bytes(n) : e8_bytes(n)) : ((fmt & 2) ? e4w_bytes(n) : e8w_bytes(n));
}
inline CamLayout cam_layout(size_t at, int n, int dbytes, int fmt) {
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    CamLayout L;
    L.raw = at;
    L.pos = up(L.raw + (size_t)n * dbytes);
    L.valid = up(L.pos + (size_t)n * 24);
    L.csa = up(L.valid + (size_t)n);
    L.e8 = up(L.csa + (size_t)n * dbytes);
    L.end = up(L.e8 + operand_bytes(n, dbytes, fmt));
    return L;
}

struct Keyframe {
    std::vector<Cam> cams;
    bool live = false;
    void* block = nullptr;     // the keyframe's range of the store arena
    size_t block_bytes = 0;
};

struct BlockRef { void* p; size_t bytes; };

// one keyframe pair as two camera spans (store keyframes or transient uploads)
struct PairRef { const Cam* from; int n_from; const Cam* to; int n_to; };

// Host mirror + device buffers of the place recogniser (uz_places.cuh)
struct PlaceInfo { int32_t handle; long long stamp_ns; bool live; uint32_t ins_begin, ins_count; };
struct PlacesState {
    uz_place_params params;
    std::vector<PlaceInfo> places;                      // index = place index (place_count_ == places.size())
    std::unordered_map<int32_t, int32_t> by_handle;     // live places only (place_id_map_.right)
    std::unordered_set<uint64_t> checked;               // checked_: (from handle << 32) | to handle
    std::vector<PlaceCam> inserted;                     // every camera ever inserted (relink on growth); n = 0: its keyframe left the store
    std::unordered_map<int32_t, std::vector<BlockRef>> retired;   // place -> store ranges it still reads although the keyframe was
                                                        // replaced since (uz_store_replace): the reference's recogniser keeps its own
                                                        // copy of the descriptors it was given (lsh_set_recognizer.cpp:253-259)
    PlaceSlot* d_slots = nullptr; uint32_t n_slots = 0;
    PlaceNode* d_nodes = nullptr; size_t node_cap = 0, n_nodes = 0;
    size_t live_entries = 0;                            // upper bound of distinct keys (for the load factor)
    long long* d_stamps = nullptr; uint8_t* d_live = nullptr; size_t place_cap = 0;
    DevBuf d_cams, d_votes, d_out, d_out_votes;
    int64_t last_votes_bytes = 0;
};

}  // namespace

struct uz_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaStream_t side = nullptr;     // high-priority stream the uploads of a chunked batch run on
    cudaStream_t alt = nullptr;      // second compute stream: odd chunks of a chunked host batch run here, so that the
                                     // next chunk's match CTAs fill the SMs while the previous chunk's last wave drains
    int alt_chunks = 1;              // UZ_ALT_CHUNKS=0: every chunk on the context stream (kernel boundaries serialise)
    cudaStream_t solve_stream = nullptr;   // high-priority stream of the streaming solve grid (runs beside the POPC match kernel)
    int stream_probe = 0;            // UZ_STREAM_PROBE: measurement / test hooks of the streaming solve (scripts/gpu_stream_probe.py)
    int stream_min_pairs = 0;        // UZ_STREAM_SOLVE_MIN_PAIRS: smallest batch that takes the streaming form (0 = two pairs per CTA)
    int stream_solve_ctas = 1;       // UZ_STREAM_SOLVE: persistent solve CTAs per SM (0 = off: one solve CTA per pair behind the match kernel)
    int force_cfg = -1;              // UZ_KNN_CFG: force a knn2 tile shape (tuning knob)
    int xcheck_fused = 1;            // UZ_XCHECK_FUSED=0: cross-check by a second, reversed matching (the measured alternative)
    int force_wide_cfg = -1;         // UZ_KNN_WIDE_CFG: force a knn2_wide tile shape (0 = 256 x 2, 1 = 64 x 2)
    int match_mma_wide = 1;          // UZ_MATCH_MMA_WIDE: 1 = 512-bit rows on 4-bit operands (knn2_mmaf_kernel<true>, default); 2 = on two int8
                                     // planes (knn2_mmaw_kernel, the measured alternative); 0 = on the integer pipes (knn2_wide_kernel)
    int match_mma = 1;               // UZ_MATCH_MMA: 0 = 256-bit rows on the integer pipes (knn2_kernel); 1 = tensor cores, 4-bit operands
                                     // (kind::mxf4, knn2_mmaf_kernel, default); measured alternatives on int8 operands: 4 = keys formed
                                     // by the MMA (knn2_mmak_kernel), 2 / 3 = CTA pairs (knn2_mma2_kernel) for launches that fill the
                                     // chip / always, 7 = IMAD epilogue (knn2_mma_kernel)
    uint8_t* f4_zeros = nullptr;     // kF4ZeroPageBytes of zeros: what knn2_mmaf_kernel completes a ragged train tile from
    bool narrow_e4 = true;           // 32-byte rows keep the 4-bit operand layout (match_mma == 1), else the int8 one
    bool wide_e4 = true;             // 64-byte rows: the same (match_mma_wide == 1)
    int operand_fmt() const { return (narrow_e4 ? 1 : 0) | (wide_e4 ? 2 : 0); }
    std::vector<uint8_t> task_wide;  // per task of the batch being prepared: 64-byte rows
    std::vector<int4> merge_table;   // per batch: tasks whose train rows were cut into segments
    int solve_wide = 1;              // UZ_SOLVE_WIDE=0: never use the 512-thread solve CTA for small launches
    int segment_small = 1;           // UZ_SEGMENT=0: never cut small launches along the train rows
    uz_params params;
    std::string err;
    int variant_csa = 1;
    int variant_pack16 = 1;          // UZ_KNN_VARIANT=2: CSA layout with 32-bit keys (the previous kernel), =1: textbook 8-POPC
    int sm_count = 148;

    Arena store_arena, transient;
    std::vector<Keyframe> kfs;
    std::vector<int32_t> free_handles;
    int32_t live = 0;
    int32_t store_max_n = 0;

    // pinned, device-mapped host ranges seen during the CURRENT entry-point call (host begin, host end, device address
    // of begin); cleared at every call, because the host may free or re-map a range between calls
    struct MappedRange { uintptr_t hb, he, db; };
    std::vector<MappedRange> mapped;
    int map_hits = 0, map_misses = 0;
    void* pfn_ptr_attr = nullptr;    // cuPointerGetAttribute via cudaGetDriverEntryPoint (no link-time libcuda)
    int gather_upload = 1;           // UZ_GATHER_UPLOAD=0 forces the cudaMemcpyAsync path
    int copy_beside_compute = 0;     // set while uploads are enqueued that overlap the match kernel
    int copy_ctas = 24;              // UZ_COPY_CTAS: gather CTAs beside the compute kernels.  16 CTAs already pull 49 GB/s over PCIe; every one
                                     // takes a solve CTA's place on its SM (C4 end to end: 8: 2.21 M, 16: 2.75 M, 24: 2.81 M, 32: 2.75 M, 64: 2.64 M edges/s)
    cudaEvent_t trace_mid = nullptr; // UZ_TRACE=2: recorded between the gather and the layout pass of an upload
    // pageable sources are staged through this pinned ring (two halves, an event each) and pulled by the same gather kernel
    PinBuf ring;
    size_t ring_half = (size_t)32 << 20;          // UZ_RING_MB
    size_t ring_used[2] = {0, 0};
    cudaEvent_t ring_free[2] = {nullptr, nullptr};
    bool ring_busy[2] = {false, false};
    int ring_cur = 0;
    uintptr_t ring_dev = 0;          // device address of the ring
    HostPool* pool = nullptr;        // packs pageable buffers into the ring (UZ_STAGE_THREADS, default min(8, cores / 2))
    int stage_threads = 0;
    PlacesState places;
    double places_ms[3] = {0, 0, 0};
    int host_chunks = 0;             // UZ_HOST_CHUNKS: upload/compute pipeline depth of uz_estimate_edges_host (0 = auto)
    BumpPool<false> d_chunks;
    BumpPool<true> h_chunks;
    PinBuf h_results;                // pinned landing zone of the edge records of uz_estimate_edges_host

    // sample tables, keyed by (iterations, do_prosac): the path alternates estimateEdge (100, prosac) with
    // calcValidEdges (200, no prosac) on one context
    struct SampleTable { int iters = -1, prosac = -1, cap = -1; DevBuf d; uint64_t last_use = 0; };
    SampleTable samples[4];
    uint64_t sample_clock = 0;

    // per-batch staging, double buffered: the host prepares batch i+1 (task/tile tables in pinned memory) while
    // the GPU still works on batch i; a slot is reused once the event recorded behind its last kernel fired
    struct Slot {
        DevBuf d_tasks, d_tiles, d_pair_tasks, d_keys, d_pending, d_tables;
        PinBuf h_tasks, h_tiles, h_pair_tasks, h_pending, h_tables;
        cudaEvent_t done = nullptr;
        bool used = false;
    };
    // Ring of launch slots: a batch waits for the batch that used its slot `slot_depth` launches ago.  Two deep for
    // store-resident calls (their tables are large); the chunked host path runs kSlots deep so that the host can enqueue
    // uploads several chunks ahead of the compute streams.
    static constexpr int kSlots = 6;
    Slot slots[kSlots];
    int cur_slot = 0, slot_depth = 2, host_slots = kSlots;
    int solve_smem_pad = 0;          // UZ_SOLVE_SMEM_PAD (measurement only): extra dynamic shared memory per solve CTA = fewer CTAs per SM
    int pipeline_calls = 1;          // UZ_PIPELINE_CALLS=0: synchronous store-resident calls run as one launch pair (run_pairs_pipelined)
    int host_first_waves = 1;        // UZ_HOST_FIRST_WAVES: size of the first chunk of uz_estimate_edges_host in solve waves (0: like the others)
    DevBuf d_results, d_dbg_matches, d_dbg_mask, d_dbg_counts, d_dbg_phase, d_misc;

    // where the solve kernel writes the records of the batch being launched (group mode: a peer-mapped buffer)
    // parity taps
    int debug = 0;
    int dbg_cap = 0, dbg_pairs = 0, dbg_iters = 0;

    // introspection
    int64_t launches = 0;
    int timers = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double match_ms = 0, solve_ms = 0;
    int64_t match_launches = 0, solve_launches = 0, compares = 0;
    int64_t mma_launches = 0;        // match launches that ran on the tensor cores
    // lazily resolved event triples (start, after K1, after solve): recording costs ~1 us and no sync,
    // so the timers can stay on inside a timed region; uz_get_timers() synchronises and folds them in
    struct Timed { cudaEvent_t e[4]; bool has_solve; };   // knn2 begin/end, solve begin/end
    std::vector<Timed> pending;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t get_event() {
        if (!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
};

namespace {

void places_release(uz_context* ctx);
uz_status places_reset(uz_context* ctx);
void places_forget_handle(uz_context* ctx, int32_t handle);

// UZ_TRACE=1: host-side stage times of the batched entry points on stderr
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    const char* name;
    Trace(const char* n) : on(getenv("UZ_TRACE") != nullptr), name(n) { if (on) t0 = std::chrono::steady_clock::now(); }
    void lap(const char* what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[uz trace] %s: %s %.3f ms\n", name, what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// UZ_TRACE=1: accumulated host time per stage of the chunked host path (printed by uz_estimate_edges_host)
struct StageClock {
    bool on = getenv("UZ_TRACE") != nullptr;
    double acc[16] = {0};
    std::chrono::steady_clock::time_point t;
    void start() { if (on) t = std::chrono::steady_clock::now(); }
    void stop(int k) { if (on) { auto n = std::chrono::steady_clock::now(); acc[k] += std::chrono::duration<double, std::milli>(n - t).count(); t = n; } }
    void report(const char* const* names, int n) {
        if (!on) return;
        for (int i = 0; i < n; ++i) fprintf(stderr, "[uz trace]   %-28s %.3f ms\n", names[i], acc[i]);
        for (double& a : acc) a = 0;
    }
};
StageClock g_stage;

std::string g_create_err = "";   // why the last uz_create failed (uz_last_error(NULL))

uz_status fail(uz_context* ctx, uz_status st, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_create_err = msg;
    return st;
}

#define UZ_CUDA(ctx, call)                                                                              \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            cudaGetLastError();                                                                         \
            return fail((ctx), UZ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));       \
        }                                                                                               \
    } while (0)

int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// features_.cols as the ABI carries it: 0 = 32; anything but 32 / 64 is unsupported (0)
int desc_width(int desc_bytes) {
    if (desc_bytes == 0) return UZ_DESC_BYTES;
    return (desc_bytes == UZ_DESC_BYTES || desc_bytes == UZ_MAX_DESC_BYTES) ? desc_bytes : 0;
}

bool is_binary_type(int t) { return t >= UZ_FEATURE_BRIEF && t <= UZ_FEATURE_FREAK; }   // :54-57

double thr_sq_star(double thr) {
    // smallest double s with sqrt(s) >= thr (sqrt correctly rounded), so sqrt(s) < thr <=> s < s*
    if (!(thr > 0.0)) return 0.0;
    if (std::isinf(thr)) return thr;
    double c = thr * thr;
    while (c > 0.0 && std::sqrt(c) >= thr) c = std::nextafter(c, -INFINITY);
    while (std::sqrt(std::nextafter(c, INFINITY)) < thr) c = std::nextafter(c, INFINITY);
    return std::nextafter(c, INFINITY);
}

uz_status check_ctx(uz_context* ctx) {
    if (!ctx) return UZ_ERR_INVALID;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(ctx, UZ_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    ctx->mapped.clear(); ctx->map_hits = 0; ctx->map_misses = 0;
    return UZ_OK;
}

}  // namespace
