// uz_capi.cu — context, device-resident keyframe store and the C-ABI of include/uzliti_edge.h.
//
// Host side of the batched feature-edge path.  What the reference does per pair on one worker thread
// (/root/reference/transformation_estimation/src/transformation_estimator.cpp:45-62: LIFO pop, impl,
// callback, 1 ms sleep) becomes: enumerate the camera-pair matchings of all pairs on the host
// (feature_transformation_estimator.cpp:40-49), one knn2 launch over every (matching, query tile), one
// solve launch with a CTA per pair, one result copy.  No CPU compute fallback exists: without a
// usable device every compute entry point returns UZ_ERR_CUDA.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/uzliti_edge.h"
#include "uz_ingest.cuh"
#include "uz_knn2.cuh"
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

// Bump allocator over large device chunks (HBM3e: 180 GB — chunks are cheap, fragmentation is not an issue
// for append-mostly keyframe maps).  Memory of removed keyframes is reclaimed by uz_store_clear().
struct Arena {
    struct Chunk { uint8_t* base; size_t size, used; };
    std::vector<Chunk> chunks;
    size_t chunk_bytes = (size_t)64 << 20;
    size_t total = 0;
    void* alloc(size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        if (bytes == 0) bytes = 256;
        for (auto& c : chunks)
            if (c.size - c.used >= bytes) { void* r = c.base + c.used; c.used += bytes; return r; }
        size_t sz = std::max(chunk_bytes, bytes);
        void* p = nullptr;
        if (cudaMalloc(&p, sz) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        chunks.push_back(Chunk{(uint8_t*)p, sz, bytes});
        total += sz;
        return p;
    }
    void reset() { for (auto& c : chunks) c.used = 0; }
    void release() { for (auto& c : chunks) cudaFree(c.base); chunks.clear(); total = 0; }
};

struct Cam {
    uint32_t* raw = nullptr;   // n x dbytes/4 words, bytes as given
    uint32_t* csa = nullptr;   // same rows, every 256-bit half in CSA layout (uz_knn2.cuh)
    double* pos = nullptr;     // 3 x n column-major
    uint8_t* valid = nullptr;  // n
    int32_t n = 0, feature_type = 0, sensor_frame = 0;
    int32_t dbytes = UZ_DESC_BYTES;   // descriptor width: 32 or 64
};

struct Keyframe {
    std::vector<Cam> cams;
    bool live = false;
};

// one keyframe pair as two camera spans (store keyframes or transient uploads)
struct PairRef { const Cam* from; int n_from; const Cam* to; int n_to; };

// Host mirror + device buffers of the place recogniser (uz_places.cuh)
struct PlaceInfo { int32_t handle; long long stamp_ns; bool live; };
struct PlacesState {
    uz_place_params params;
    std::vector<PlaceInfo> places;                      // index = place index (place_count_ == places.size())
    std::unordered_map<int32_t, int32_t> by_handle;     // live places only (place_id_map_.right)
    std::unordered_set<uint64_t> checked;               // checked_: (from handle << 32) | to handle
    std::vector<PlaceCam> inserted;                     // every camera ever inserted (relink on growth)
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
    cudaStream_t side = nullptr;     // high-priority stream the solve kernels of a chunked batch run on
    cudaStream_t alt = nullptr;      // second compute stream: odd chunks of a chunked host batch run here, so that the
                                     // next chunk's match CTAs fill the SMs while the previous chunk's last wave drains
    int alt_chunks = 1;              // UZ_ALT_CHUNKS=0: every chunk on the context stream (kernel boundaries serialise)
    cudaStream_t solve_stream = nullptr;   // high-priority stream of the streaming solve grid (runs beside the match kernel)
    int stream_probe = 0;            // UZ_STREAM_PROBE: measurement / test hooks of the streaming solve (scripts/gpu_stream_probe.py)
    int stream_min_pairs = 0;        // UZ_STREAM_SOLVE_MIN_PAIRS: smallest batch that takes the streaming form (0 = two pairs per CTA)
    int stream_solve_ctas = 1;       // UZ_STREAM_SOLVE: persistent solve CTAs per SM (0 = off: one solve CTA per pair behind the match kernel)
    int force_cfg = -1;              // UZ_KNN_CFG: force a knn2 tile shape (tuning knob)
    int xcheck_fused = 1;            // UZ_XCHECK_FUSED=0: cross-check by a second, reversed matching (the measured alternative)
    int force_wide_cfg = -1;         // UZ_KNN_WIDE_CFG: force a knn2_wide tile shape (0 = 256 x 2, 1 = 64 x 2)
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

    // pinned, device-mapped host ranges seen so far (host begin, host end, device address of begin)
    struct MappedRange { uintptr_t hb, he, db; };
    std::vector<MappedRange> mapped;
    void* pfn_ptr_attr = nullptr;    // cuPointerGetAttribute via cudaGetDriverEntryPoint (no link-time libcuda)
    int gather_upload = 1;           // UZ_GATHER_UPLOAD=0 forces the cudaMemcpyAsync path
    int copy_beside_compute = 0;     // set while uploads are enqueued that overlap the match kernel
    int copy_ctas = 64;              // UZ_COPY_CTAS
    PlacesState places;
    double places_ms[3] = {0, 0, 0};
    int host_chunks = 0;             // UZ_HOST_CHUNKS: upload/compute pipeline depth of uz_estimate_edges_host (0 = auto)
    BumpPool<false> d_chunks;
    BumpPool<true> h_chunks;
    PinBuf h_results;                // pinned landing zone of the edge records of uz_estimate_edges_host

    // sample table
    DevBuf d_samples;
    int samp_cap = -1, samp_iters = -1, samp_prosac = -1;

    // per-batch staging, double buffered: the host prepares batch i+1 (task/tile tables in pinned memory) while
    // the GPU still works on batch i; a slot is reused once the event recorded behind its last kernel fired
    struct Slot {
        DevBuf d_tasks, d_tiles, d_pair_tasks, d_keys, d_pending, d_tables;
        PinBuf h_tasks, h_tiles, h_pair_tasks, h_pending, h_tables;
        cudaEvent_t done = nullptr;
        bool used = false;
    };
    Slot slots[2];
    int cur_slot = 0;
    DevBuf d_results, d_dbg_matches, d_dbg_mask, d_dbg_counts, d_dbg_phase, d_misc;

    // parity taps
    int debug = 0;
    int dbg_cap = 0, dbg_pairs = 0, dbg_iters = 0;

    // introspection
    int64_t launches = 0;
    int timers = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double match_ms = 0, solve_ms = 0;
    int64_t match_launches = 0, solve_launches = 0, compares = 0;
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

// ---- uploads -------------------------------------------------------------------------------------
struct Span { const uint8_t* host; size_t bytes; uint8_t** dev_slot; };

// A host-contiguous run of spans and where it goes on the device.
struct Run { const uint8_t* host; size_t bytes; uint8_t* dev; };

// Lays a set of host spans out in ONE device block, merging host-contiguous spans into runs (a map laid out
// as one big array on the host becomes one run per field) and sharing one device copy between identical
// host pointers.  No data moves here; the runs are appended to `runs` for flush_runs().
uz_status plan_spans(uz_context* ctx, Arena& arena, std::vector<Span>& spans, std::vector<Run>& runs,
                     uint8_t** block_out = nullptr, size_t* block_bytes_out = nullptr) {
    std::sort(spans.begin(), spans.end(), [](const Span& a, const Span& b) {
        return a.host != b.host ? a.host < b.host : a.bytes > b.bytes;
    });
    struct Tmp { size_t first, last; const uint8_t* b; const uint8_t* e; size_t off; };
    std::vector<Tmp> tmp;
    size_t total = 0;
    size_t i = 0;
    while (i < spans.size()) {
        size_t j = i;
        const uint8_t* run_begin = spans[i].host;
        const uint8_t* run_end = spans[i].host + spans[i].bytes;
        while (j + 1 < spans.size() && spans[j + 1].host <= run_end &&
               (spans[j + 1].host == run_end || spans[j + 1].host + spans[j + 1].bytes <= run_end ||
                spans[j + 1].host == spans[j].host)) {
            // contiguous continuation, a span nested in the run, or a duplicate
            run_end = std::max(run_end, spans[j + 1].host + spans[j + 1].bytes);
            ++j;
        }
        tmp.push_back(Tmp{i, j, run_begin, run_end, total});
        total += ((size_t)(run_end - run_begin) + 31) & ~(size_t)31;      // runs stay 32 B aligned inside the block
        i = j + 1;
    }
    uint8_t* block = (uint8_t*)arena.alloc(std::max<size_t>(total, 1));
    if (!block) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
    for (const Tmp& t : tmp) {
        runs.push_back(Run{t.b, (size_t)(t.e - t.b), block + t.off});
        for (size_t k = t.first; k <= t.last; ++k) *spans[k].dev_slot = block + t.off + (spans[k].host - t.b);
    }
    if (block_out) *block_out = block;
    if (block_bytes_out) *block_bytes_out = total;
    return UZ_OK;
}

// Device-side address of a pinned, device-mapped host range, or 0 if the range is not (entirely) mapped.
uintptr_t mapped_device_address(uz_context* ctx, const uint8_t* host, size_t bytes) {
    const uintptr_t h = (uintptr_t)host;
    for (const auto& r : ctx->mapped)
        if (h >= r.hb && h + bytes <= r.he) return r.db + (h - r.hb);
    typedef CUresult (*attr_fn)(void*, CUpointer_attribute, CUdeviceptr);
    if (!ctx->pfn_ptr_attr) {
        cudaDriverEntryPointQueryResult q;
        void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
            cudaGetLastError();
            return 0;
        }
        ctx->pfn_ptr_attr = fn;
    }
    attr_fn get = (attr_fn)ctx->pfn_ptr_attr;
    unsigned int mtype = 0;
    CUdeviceptr start = 0, dptr = 0;
    size_t size = 0;
    if (get(&mtype, CU_POINTER_ATTRIBUTE_MEMORY_TYPE, (CUdeviceptr)h) != CUDA_SUCCESS || mtype != CU_MEMORYTYPE_HOST) return 0;
    if (get(&start, CU_POINTER_ATTRIBUTE_RANGE_START_ADDR, (CUdeviceptr)h) != CUDA_SUCCESS) return 0;
    if (get(&size, CU_POINTER_ATTRIBUTE_RANGE_SIZE, (CUdeviceptr)h) != CUDA_SUCCESS) return 0;
    if (get(&dptr, CU_POINTER_ATTRIBUTE_DEVICE_POINTER, (CUdeviceptr)h) != CUDA_SUCCESS || !dptr) return 0;
    uz_context::MappedRange r;
    r.hb = (uintptr_t)start; r.he = r.hb + size; r.db = (uintptr_t)dptr - (h - r.hb);
    if (ctx->mapped.size() < 4096) ctx->mapped.push_back(r);
    if (h >= r.hb && h + bytes <= r.he) return r.db + (h - r.hb);
    return 0;
}

// Moves the planned runs.  Pinned, device-mapped sources are pulled by ONE gather kernel; anything else
// (pageable memory, or few large runs where the DMA engines are the better tool) goes through cudaMemcpyAsync.
uz_status flush_runs(uz_context* ctx, const std::vector<Run>& runs) {
    bool gather = ctx->gather_upload && runs.size() > 16;
    std::vector<uintptr_t> dev_src;
    if (gather) {
        dev_src.resize(runs.size());
        for (size_t i = 0; i < runs.size() && gather; ++i) {
            dev_src[i] = runs[i].bytes ? mapped_device_address(ctx, runs[i].host, runs[i].bytes) : 1;
            if (!dev_src[i]) gather = false;
        }
    }
    if (!gather) {
        for (const Run& r : runs)
            if (r.bytes) UZ_CUDA(ctx, cudaMemcpyAsync(r.dev, r.host, r.bytes, cudaMemcpyHostToDevice, ctx->stream));
        return UZ_OK;
    }
    const size_t kChunk = 16384;
    size_t n_chunks = 0;
    for (const Run& r : runs) n_chunks += (r.bytes + kChunk - 1) / kChunk;
    if (n_chunks == 0) return UZ_OK;
    CopyChunk* cc = (CopyChunk*)ctx->h_chunks.alloc(n_chunks * sizeof(CopyChunk));
    CopyChunk* d_cc = (CopyChunk*)ctx->d_chunks.alloc(n_chunks * sizeof(CopyChunk));
    if (!cc || !d_cc) return fail(ctx, UZ_ERR_NOMEM, "copy-chunk table allocation failed");
    size_t k = 0;
    for (size_t i = 0; i < runs.size(); ++i)
        for (size_t off = 0; off < runs[i].bytes; off += kChunk) {
            cc[k].src = (const uint8_t*)(dev_src[i] + off);
            cc[k].dst = runs[i].dev + off;
            cc[k].bytes = (uint32_t)std::min(kChunk, runs[i].bytes - off);
            cc[k].pad = 0;
            ++k;
        }
    UZ_CUDA(ctx, cudaMemcpyAsync(d_cc, cc, n_chunks * sizeof(CopyChunk), cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->copy_beside_compute)      // few small CTAs: they fit in the registers the match kernel leaves free
        gather_copy_kernel<<<(unsigned)std::min<size_t>(n_chunks, (size_t)ctx->copy_ctas), 128, 0, ctx->stream>>>(d_cc, (int)n_chunks);
    else
        gather_copy_kernel<<<(unsigned)std::min<size_t>(n_chunks, (size_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(d_cc, (int)n_chunks);
    ctx->launches++;
    UZ_CUDA(ctx, cudaGetLastError());
    return UZ_OK;
}

uz_status validate_features(uz_context* ctx, const uz_features* f) {
    if (f->n < 0 || f->n > UZ_MAX_FEATURES) return fail(ctx, UZ_ERR_INVALID, "feature count out of range (0..UZ_MAX_FEATURES)");
    if (f->n > 0 && (!f->descriptors || !f->positions || !f->valid_3d)) return fail(ctx, UZ_ERR_INVALID, "null feature buffer");
    const int db = desc_width(f->desc_bytes);
    if (db == 0) return fail(ctx, UZ_ERR_UNSUPPORTED, "descriptor width must be 32 or 64 bytes (256- or 512-bit binary descriptors)");
    if (f->n > 0 && f->desc_stride < db) return fail(ctx, UZ_ERR_INVALID, "descriptor stride < descriptor width");
    return UZ_OK;
}

// Uploads cameras (descriptors, positions, valid) and builds both descriptor layouts on the device.
uz_status upload_cams(uz_context* ctx, Arena& arena, const std::vector<const uz_features*>& feats,
                      std::vector<Cam>& out) {
    out.assign(feats.size(), Cam());
    std::vector<Span> dspans, pspans, vspans;
    std::vector<uint8_t*> d_raw(feats.size(), nullptr), d_pos(feats.size(), nullptr), d_val(feats.size(), nullptr);
    std::vector<size_t> strided;
    for (size_t i = 0; i < feats.size(); ++i) {
        const uz_features* f = feats[i];
        uz_status st = validate_features(ctx, f);
        if (st != UZ_OK) return st;
        out[i].n = f->n; out[i].feature_type = f->feature_type; out[i].sensor_frame = f->sensor_frame;
        out[i].dbytes = desc_width(f->desc_bytes);
        if (f->n == 0) continue;
        if (f->desc_stride == out[i].dbytes) dspans.push_back(Span{f->descriptors, (size_t)f->n * out[i].dbytes, &d_raw[i]});
        else strided.push_back(i);
        pspans.push_back(Span{(const uint8_t*)f->positions, (size_t)f->n * 24, &d_pos[i]});
        vspans.push_back(Span{f->valid_3d, (size_t)f->n, &d_val[i]});
    }
    uz_status st;
    std::vector<Run> runs;
    uint8_t* raw_block = nullptr;
    size_t raw_bytes = 0;
    if ((st = plan_spans(ctx, arena, dspans, runs, &raw_block, &raw_bytes)) != UZ_OK) return st;
    if ((st = plan_spans(ctx, arena, pspans, runs)) != UZ_OK) return st;
    if ((st = plan_spans(ctx, arena, vspans, runs)) != UZ_OK) return st;
    if ((st = flush_runs(ctx, runs)) != UZ_OK) return st;
    // CSA layout of the packed block: one launch, the CSA block mirrors the raw block 256-bit half by half (a 64-byte
    // row is two halves, each with its own transform; runs are 32 B aligned inside the block)
    if (raw_bytes) {
        uint8_t* csa = (uint8_t*)arena.alloc(raw_bytes);
        if (!csa) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
        const size_t rows = raw_bytes / 32;
        pack_descriptors_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(
            raw_block, (int)rows, 32, (uint32_t*)raw_block, (uint32_t*)csa, 1);
        ctx->launches++;
        for (size_t i = 0; i < feats.size(); ++i)
            if (d_raw[i]) { out[i].raw = (uint32_t*)d_raw[i]; out[i].csa = (uint32_t*)(csa + (d_raw[i] - raw_block)); }
    }
    for (size_t i : strided) {   // padded cv::Mat rows: pitch copy into packed rows, then their own CSA pass
        const uz_features* f = feats[i];
        const int db = out[i].dbytes, halves = f->n * (db / 32);
        d_raw[i] = (uint8_t*)arena.alloc((size_t)f->n * db);
        uint8_t* csa = (uint8_t*)arena.alloc((size_t)f->n * db);
        if (!d_raw[i] || !csa) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
        UZ_CUDA(ctx, cudaMemcpy2DAsync(d_raw[i], db, f->descriptors, f->desc_stride, db, f->n,
                                       cudaMemcpyHostToDevice, ctx->stream));
        pack_descriptors_kernel<<<(halves + 255) / 256, 256, 0, ctx->stream>>>(d_raw[i], halves, 32, (uint32_t*)d_raw[i], (uint32_t*)csa, 1);
        ctx->launches++;
        out[i].raw = (uint32_t*)d_raw[i]; out[i].csa = (uint32_t*)csa;
    }
    UZ_CUDA(ctx, cudaGetLastError());
    for (size_t k = 0; k < feats.size(); ++k) { out[k].pos = (double*)d_pos[k]; out[k].valid = d_val[k]; }
    return UZ_OK;
}

// ---- sample table --------------------------------------------------------------------------------
uz_status ensure_samples(uz_context* ctx, int iterations, int do_prosac, int max_m) {
    if (ctx->samp_iters == iterations && ctx->samp_prosac == do_prosac && ctx->samp_cap >= max_m) return UZ_OK;
    int cap = std::max(256, (max_m + 255) & ~255);
    if (ctx->samp_iters == iterations && ctx->samp_prosac == do_prosac) cap = std::max(cap, ctx->samp_cap);
    std::vector<uint16_t> table;
    build_sample_table(iterations, do_prosac != 0, cap, table);
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // previous launches may still read the old table
    UZ_CUDA(ctx, ctx->d_samples.ensure(table.size() * sizeof(uint16_t)));
    // stream-ordered: a plain cudaMemcpy from pageable memory runs on the legacy stream, which a
    // non-blocking stream does not wait for
    UZ_CUDA(ctx, cudaMemcpyAsync(ctx->d_samples.p, table.data(), table.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->samp_cap = cap; ctx->samp_iters = iterations; ctx->samp_prosac = do_prosac;
    return UZ_OK;
}

// ---- launches --------------------------------------------------------------------------------------
template <int THREADS, int QPT>
void launch_knn2(uz_context* ctx, const MatchTask* d_tasks, const int2* d_tiles, int n_tiles, uint2* d_keys, int* d_pending,
                 unsigned int* d_progress, bool xchk, bool seg) {
    if (xchk || seg) {
        if constexpr (QPT == 2) {    // fused cross-check: column minima in the same pass (+ 4 B of shared memory per staged train row);
                                     // seg: tiles are int4 and name a segment of the train rows (small launches)
            const size_t sm = knn_smem_bytes(THREADS, QPT) + (xchk ? knn_train_rows(THREADS, QPT) * 4 : 0);
            if (xchk && seg) knn2_kernel<THREADS, QPT, true, true, true, true><<<n_tiles, THREADS, sm, ctx->stream>>>(d_tasks, d_tiles, d_keys, d_pending, d_progress);
            else if (xchk) knn2_kernel<THREADS, QPT, true, true, true, false><<<n_tiles, THREADS, sm, ctx->stream>>>(d_tasks, d_tiles, d_keys, d_pending, d_progress);
            else knn2_kernel<THREADS, QPT, true, true, false, true><<<n_tiles, THREADS, sm, ctx->stream>>>(d_tasks, d_tiles, d_keys, d_pending, d_progress);
        }
    } else if (ctx->variant_csa && ctx->variant_pack16)
        knn2_kernel<THREADS, QPT, true, true><<<n_tiles, THREADS, knn_smem_bytes(THREADS, QPT), ctx->stream>>>(d_tasks, d_tiles, d_keys, d_pending, d_progress);
    else if (ctx->variant_csa)
        knn2_kernel<THREADS, QPT, true><<<n_tiles, THREADS, knn_smem_bytes(THREADS, QPT), ctx->stream>>>(d_tasks, d_tiles, d_keys, d_pending, d_progress);
    else
        knn2_kernel<THREADS, QPT, false><<<n_tiles, THREADS, knn_smem_bytes(THREADS, QPT), ctx->stream>>>(d_tasks, d_tiles, d_keys, d_pending, d_progress);
}

template <int THREADS>
void launch_knn2_wide(uz_context* ctx, const MatchTask* d_tasks, const int2* d_tiles, int n_tiles, uint2* d_keys, int* d_pending,
                      unsigned int* d_progress, bool xchk, bool seg) {
    const size_t sm = knn_wide_smem_bytes(THREADS) + (xchk ? knn_wide_train_rows(THREADS) * 4 : 0);
    if (xchk && seg) knn2_wide_kernel<THREADS, true, true><<<n_tiles, THREADS, sm, ctx->stream>>>(d_tasks, d_tiles, d_keys, d_pending, d_progress);
    else if (xchk) knn2_wide_kernel<THREADS, true, false><<<n_tiles, THREADS, sm, ctx->stream>>>(d_tasks, d_tiles, d_keys, d_pending, d_progress);
    else if (seg) knn2_wide_kernel<THREADS, false, true><<<n_tiles, THREADS, sm, ctx->stream>>>(d_tasks, d_tiles, d_keys, d_pending, d_progress);
    else knn2_wide_kernel<THREADS><<<n_tiles, THREADS, sm, ctx->stream>>>(d_tasks, d_tiles, d_keys, d_pending, d_progress);
}

// One solve CTA per pair (or per direct problem).  Launches of at most one pair per SM take the wide CTA: the chip is
// otherwise idle, so only the latency of that one CTA counts (UZ_SOLVE_WIDE=0: always the 128-thread CTA).
void launch_solve(uz_context* ctx, int n_ctas, int cap, cudaStream_t st, const MatchTask* d_tasks, const int2* d_pair_tasks,
                  const uint2* d_keys, const SolveParams& sp, uz_edge_result* d_results) {
    if (ctx->solve_wide && n_ctas <= ctx->sm_count)
        solve_kernel<kSolveThreadsWide><<<n_ctas, kSolveThreadsWide, solve_smem_bytes(cap, kSolveThreadsWide), st>>>(d_tasks, d_pair_tasks, d_keys, sp, d_results);
    else
        solve_kernel<kSolveThreads><<<n_ctas, kSolveThreads, solve_smem_bytes(cap), st>>>(d_tasks, d_pair_tasks, d_keys, sp, d_results);
}

// Kernels of one library that are meant to run beside each other must agree on the shared-memory carve-out of the SM:
// an SM is only reconfigured when it is empty, so a kernel that asks for another split waits until the resident
// kernel's CTAs have drained - which serialises the two (and starves a consumer that polls its producer).
template <typename K>
cudaError_t max_shared_carveout(K kernel) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}
template <int THREADS, int QPT>
cudaError_t knn2_carveout() {
    cudaError_t e = max_shared_carveout(knn2_kernel<THREADS, QPT, true, true>);
    if (e == cudaSuccess) e = max_shared_carveout(knn2_kernel<THREADS, QPT, true, false>);
    if (e == cudaSuccess) e = max_shared_carveout(knn2_kernel<THREADS, QPT, false, false>);
    if constexpr (QPT == 2) { if (e == cudaSuccess) e = max_shared_carveout(knn2_kernel<THREADS, QPT, true, true, true>); }
    return e;
}
cudaError_t set_carveouts() {
    cudaError_t e = knn2_carveout<256, 4>();
    if (e == cudaSuccess) e = knn2_carveout<128, 4>();
    if (e == cudaSuccess) e = knn2_carveout<64, 2>();
    if (e == cudaSuccess) e = knn2_carveout<256, 2>();
    if (e == cudaSuccess) e = knn2_carveout<128, 2>();
    if (e == cudaSuccess) e = knn2_carveout<32, 2>();
    if (e == cudaSuccess) e = max_shared_carveout(knn2_wide_kernel<256>);
    if (e == cudaSuccess) e = max_shared_carveout(knn2_wide_kernel<64>);
    if (e == cudaSuccess) e = max_shared_carveout(knn2_wide_kernel<256, true>);
    if (e == cudaSuccess) e = max_shared_carveout(knn2_wide_kernel<64, true>);
    if (e == cudaSuccess) e = max_shared_carveout(solve_kernel<kSolveThreads>);
    if (e == cudaSuccess) e = max_shared_carveout(solve_stream_kernel<kSolveThreads>);
    if (e == cudaSuccess) e = max_shared_carveout(gather_copy_kernel);
    if (e == cudaSuccess) e = max_shared_carveout(pack_descriptors_kernel);
    return e;
}

// relative throughput of a shape at full occupancy (measured on C4-sized work; the one-warp CTA is capped at 32 warps per SM)
struct KnnConfig { int threads, qpt; double speed; };
const KnnConfig kKnnConfigs[6] = {{256, 4, 0.92}, {128, 4, 0.92}, {64, 2, 0.985}, {256, 2, 1.0}, {128, 2, 1.0}, {32, 2, 0.95}};

// Runs K1 (+ optionally K2..K5) for a list of pairs whose cameras are already on the device.
// join = false (chunked callers): the caller's stream is NOT made to wait for a streaming solve, so that the next chunk's
// match kernel starts while this chunk's last pairs are still being solved; *result_stream is then the stream behind
// which the records are complete.
uz_status run_pairs(uz_context* ctx, const std::vector<PairRef>& pairs, uz_edge_result* d_results, bool join = true,
                    cudaStream_t* result_stream = nullptr) {
    const uz_params prm = ctx->params;     // snapshot (setConfig may race with a batch in the reference)
    const int n_pairs = (int)pairs.size();
    if (result_stream) *result_stream = ctx->stream;
    if (n_pairs == 0) return UZ_OK;
    cudaStream_t results_on = ctx->stream;
    ctx->cur_slot ^= 1;
    uz_context::Slot& sl = ctx->slots[ctx->cur_slot];
    if (!sl.done) UZ_CUDA(ctx, cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    if (sl.used) UZ_CUDA(ctx, cudaEventSynchronize(sl.done));      // normally long complete

    // 1. enumerate matchings (:40-49) and pick the tile shape
    size_t max_tasks = 0;
    for (const PairRef& p : pairs) max_tasks += (size_t)p.n_from * (size_t)p.n_to;
    const bool cross = prm.cross_check != 0 && d_results != nullptr;
    // with cross-check every matching is also run reversed; the reversed tasks live behind the forward ones
    UZ_CUDA(ctx, sl.h_tasks.ensure(std::max<size_t>(max_tasks, 1) * (cross ? 2 : 1) * sizeof(MatchTask)));
    UZ_CUDA(ctx, sl.h_pair_tasks.ensure((size_t)n_pairs * sizeof(int2)));
    MatchTask* tasks = (MatchTask*)sl.h_tasks.p;
    int2* pair_tasks = (int2*)sl.h_pair_tasks.p;
    size_t n_tasks = 0, key_rows = 0;
    int max_nq = 0;
    int64_t compares = 0;
    std::vector<uint8_t>& task_wide = ctx->task_wide;      // host-side: 1 = 64-byte rows (knn2_wide_kernel)
    task_wide.assign(std::max<size_t>(max_tasks, 1) * (cross ? 2 : 1), 0);
    size_t n_wide = 0;
    for (int i = 0; i < n_pairs; ++i) {
        const int first = (int)n_tasks;
        const Cam* fc = pairs[i].from;
        const Cam* tc = pairs[i].to;
        for (int a = 0; a < pairs[i].n_from; ++a)
            for (int b = 0; b < pairs[i].n_to; ++b) {
                const Cam& F = fc[a];
                const Cam& T = tc[b];
                if (F.n >= prm.min_keypoints && T.n >= prm.min_keypoints && F.feature_type == T.feature_type &&
                    F.sensor_frame == T.sensor_frame && F.dbytes == T.dbytes) {
                    if (F.dbytes != UZ_DESC_BYTES) { task_wide[n_tasks] = 1; ++n_wide; }
                    MatchTask& tk = tasks[n_tasks++];
                    const bool bin = is_binary_type(F.feature_type);   // unknown type: empty matches (:60-62)
                    const bool use_csa = ctx->variant_csa || F.dbytes != UZ_DESC_BYTES;     // the wide kernel has the CSA form only
                    tk.q_desc = use_csa ? T.csa : T.raw;
                    tk.t_desc = use_csa ? F.csa : F.raw;
                    tk.nq = bin ? T.n : 0; tk.nt = F.n;
                    tk.key_off = (uint32_t)key_rows; tk.pair = i;
                    tk.q_pos = T.pos; tk.q_valid = T.valid; tk.t_pos = F.pos; tk.t_valid = F.valid;
                    tk.cam_from = (int)a; tk.cam_to = (int)b;
                    tk.rev_key_off = kNoRev; tk.pad_ = 0;
                    key_rows += (size_t)tk.nq;
                    max_nq = std::max(max_nq, tk.nq);
                    compares += (int64_t)tk.nq * tk.nt;
                }
            }
        pair_tasks[i] = make_int2(first, (int)n_tasks - first);
    }
    const size_t n_fwd = n_tasks;
    // Cross-check.  Fused form (default): the forward match kernel also keeps, per train row, the minimum over the
    // queries (uz_knn2.cuh, col_update16); the matching's column keys live behind the row keys in the same scratch.
    // Reversed form (UZ_XCHECK_FUSED=0, and the kernel variants without the packed-key shapes): every matching runs
    // a second time with query and train swapped.
    const bool fused = cross && ctx->xcheck_fused && ctx->variant_csa && ctx->variant_pack16 &&
                       !(ctx->force_cfg >= 0 && ctx->force_cfg < 2);
    if (!fused && cross) {
        for (size_t t = 0; t < n_fwd; ++t) {
            MatchTask& f = tasks[t];
            if (f.nq == 0) continue;                 // non-binary type: no matches to check
            MatchTask& r = tasks[n_tasks];
            r = f;
            r.q_desc = f.t_desc; r.t_desc = f.q_desc; r.nq = f.nt; r.nt = f.nq;
            r.rev_key_off = kNoRev;
            task_wide[n_tasks] = task_wide[t];
            f.rev_key_off = 0;                       // "has a reversed task"; the offset is assigned below

            compares += (int64_t)r.nq * r.nt;
            f.pad_ = (uint32_t)n_tasks;              // index of the reversed task (host-side only)
            ++n_tasks;
        }
    }
    // tile shape per descriptor width.  256-bit rows: two queries per thread (40-56 registers: 6 resident CTAs per SM,
    // measured 8 % faster than the four-query shapes, which stay reachable through UZ_KNN_CFG), largest tile first.
    // 512-bit rows: the 256 x 2 and 64 x 2 shapes of knn2_wide_kernel.
    auto pick = [&](const int* cand, int n_cand, const int* threads, const int* qpt, const double* speed, int warps_per_sm,
                    bool wide) {
        int best = cand[0];
        double best_cost = 1e300;
        for (int ci = 0; ci < n_cand; ++ci) {
            const int c = cand[ci];
            const int tile = threads[c] * qpt[c];
            double padded = 0; size_t tiles = 0;
            for (size_t t = 0; t < n_tasks; ++t) {
                if ((task_wide[t] != 0) != wide) continue;
                const size_t nt = ((size_t)tasks[t].nq + tile - 1) / tile;
                tiles += nt; padded += (double)nt * tile * tasks[t].nt;
            }
            // a launch that cannot fill the chip pays for its idle warp slots
            const double fill = std::min(1.0, (double)tiles * threads[c] / ((double)ctx->sm_count * 32 * warps_per_sm));
            const double cost = padded / (speed[c] * std::max(fill, 1e-3));
            if (cost < best_cost * 0.999) { best_cost = cost; best = c; }
        }
        return best;
    };
    int best_cfg = 3, wide_cfg = 0;
    {
        static const int kCandidates[4] = {3, 4, 2, 5};
        int th[6], qp[6]; double sp[6];
        for (int c = 0; c < 6; ++c) { th[c] = kKnnConfigs[c].threads; qp[c] = kKnnConfigs[c].qpt; sp[c] = kKnnConfigs[c].speed; }
        if (n_wide < n_tasks) best_cfg = pick(kCandidates, 4, th, qp, sp, 48, false);
        if (ctx->force_cfg >= 0 && ctx->force_cfg < 6) best_cfg = ctx->force_cfg;
        static const int kWideCand[2] = {0, 1};
        static const int wth[2] = {256, 64}, wqp[2] = {2, 2};
        static const double wsp[2] = {1.0, 0.97};
        if (n_wide) wide_cfg = pick(kWideCand, 2, wth, wqp, wsp, 32, true);
        if (ctx->force_wide_cfg >= 0 && ctx->force_wide_cfg < 2) wide_cfg = ctx->force_wide_cfg;
    }
    const int tile_rows = kKnnConfigs[best_cfg].threads * kKnnConfigs[best_cfg].qpt;
    const int wide_threads = wide_cfg == 0 ? 256 : 64;
    const int wide_tile_rows = 2 * wide_threads;
    auto rows_of = [&](size_t t) { return task_wide[t] ? wide_tile_rows : tile_rows; };
    auto qtiles_of = [&](size_t t) { return ((size_t)tasks[t].nq + rows_of(t) - 1) / rows_of(t); };

    // Small launches are cut along the train rows as well (SEG kernels): a handful of pairs - the online case, one new
    // keyframe against its candidates - would otherwise run on a handful of CTAs that each walk every train row.  Every
    // (query tile, train segment) CTA writes partial neighbours, merge_segments_kernel folds them.  Not for launches
    // that fill the chip anyway, and not beside the streaming solve (it consumes keys tile by tile).
    const bool with_solve = d_results != nullptr;
    const int cap = with_solve ? std::max(128, pow2ceil(std::max(max_nq, 1))) : 0;
    const int stream_ctas = ctx->stream_solve_ctas * ctx->sm_count;
    const bool may_stream = with_solve && ctx->solve_stream != nullptr && ctx->stream_solve_ctas > 0 && cap <= 1024 &&
                            n_pairs >= (ctx->stream_min_pairs > 0 ? ctx->stream_min_pairs : 2 * stream_ctas);
    int seg_target = 1;
    if (ctx->segment_small && !may_stream && ctx->variant_csa && ctx->variant_pack16 &&
        kKnnConfigs[best_cfg].qpt == 2) {
        double warps = 0;
        for (size_t t = 0; t < n_tasks; ++t)
            warps += (double)qtiles_of(t) * (task_wide[t] ? wide_threads : kKnnConfigs[best_cfg].threads) / 32.0;
        const double capacity = (double)ctx->sm_count * 32.0;
        if (warps > 0 && warps * 2 <= capacity) seg_target = (int)std::min(16.0, std::floor(capacity / warps));
    }
    const bool seg = seg_target > 1;
    // segment of a task: a multiple of 64 train rows (the key blocks of both kernels), at least 64
    auto seg_rows_of = [&](size_t t) {
        if (!seg) return std::max(tasks[t].nt, 1);
        const int want = (tasks[t].nt + seg_target - 1) / seg_target;
        return std::max(64, (want + 63) & ~63);
    };
    auto nseg_of = [&](size_t t) { return seg ? std::max(1, (tasks[t].nt + seg_rows_of(t) - 1) / seg_rows_of(t)) : 1; };

    // key scratch: per task nq rows per segment (segment 0 holds the final neighbours), then the column keys of the
    // fused cross-check
    key_rows = 0;
    for (size_t t = 0; t < n_tasks; ++t) {
        tasks[t].key_off = (uint32_t)std::min<size_t>(key_rows, 0xFFFFFFFFu);
        key_rows += (size_t)tasks[t].nq * (size_t)nseg_of(t);
    }
    const size_t col_begin = key_rows;
    for (size_t t = 0; t < n_fwd; ++t) {
        if (tasks[t].nq == 0) continue;
        if (fused) {
            tasks[t].rev_key_off = (uint32_t)std::min<size_t>(key_rows, 0xFFFFFFFFu);
            key_rows += (size_t)tasks[t].nt;
        } else if (cross) {
            tasks[t].rev_key_off = tasks[tasks[t].pad_].key_off;
        }
    }
    if (key_rows >= ((size_t)1 << 32)) return fail(ctx, UZ_ERR_INVALID, "batch too large: split it (key scratch > 2^32 rows)");

    size_t n_tiles = 0, n_tiles_wide = 0;
    std::vector<int4>& merges = ctx->merge_table;         // (key_off, nq, segments, 0) of every task cut into segments
    merges.clear();
    for (size_t t = 0; t < n_tasks; ++t) {
        const size_t k = qtiles_of(t) * (size_t)nseg_of(t);
        n_tiles += k;
        if (task_wide[t]) n_tiles_wide += k;
        if (nseg_of(t) > 1 && tasks[t].nq > 0) merges.push_back(make_int4((int)tasks[t].key_off, tasks[t].nq, nseg_of(t), 0));
    }
    const size_t n_tiles_narrow = n_tiles - n_tiles_wide;
    const size_t tile_bytes = seg ? sizeof(int4) : sizeof(int2);
    UZ_CUDA(ctx, sl.h_tiles.ensure(std::max<size_t>(n_tiles, 1) * tile_bytes + merges.size() * sizeof(int4)));
    int2* tiles = (int2*)sl.h_tiles.p;
    int4* tiles4 = (int4*)sl.h_tiles.p;
    {
        // 256-bit tiles first, 512-bit tiles behind them (one launch each); inside a list: forward tiles, then the
        // reversed tiles of the same matching
        size_t kn = 0, kw = n_tiles_narrow;
        auto emit_tiles = [&](size_t t, size_t& k) {
            const int tr = rows_of(t);
            if (!seg) {
                for (int q0 = 0; q0 < tasks[t].nq; q0 += tr) tiles[k++] = make_int2((int)t, q0);
                return;
            }
            const int sr = seg_rows_of(t), ns = nseg_of(t);
            for (int q0 = 0; q0 < tasks[t].nq; q0 += tr)
                for (int sg = 0; sg < ns; ++sg) {
                    const int tb = sg * sr, rows = std::max(0, std::min(tasks[t].nt - tb, sr));
                    tiles4[k++] = make_int4((int)t, q0, tb, (sg << 16) | rows);
                }
        };
        for (size_t t = 0; t < n_fwd; ++t) {
            size_t& k = task_wide[t] ? kw : kn;
            emit_tiles(t, k);
            if (!fused && cross && tasks[t].nq > 0) emit_tiles((size_t)tasks[t].pad_, k);
        }
    }
    int4* h_merges = (int4*)((uint8_t*)sl.h_tiles.p + std::max<size_t>(n_tiles, 1) * tile_bytes);
    if (!merges.empty()) memcpy(h_merges, merges.data(), merges.size() * sizeof(int4));

    // 2. device buffers
    UZ_CUDA(ctx, sl.d_tasks.ensure(std::max<size_t>(n_tasks, 1) * sizeof(MatchTask)));
    UZ_CUDA(ctx, sl.d_tiles.ensure(std::max<size_t>(n_tiles, 1) * tile_bytes + merges.size() * sizeof(int4)));
    UZ_CUDA(ctx, sl.d_pair_tasks.ensure((size_t)n_pairs * sizeof(int2)));
    UZ_CUDA(ctx, sl.d_keys.ensure(std::max<size_t>(key_rows, 1) * sizeof(uint2)));
    if (fused && key_rows > col_begin)       // column keys start at "none"; the match kernel lowers them with atomicMin
        UZ_CUDA(ctx, cudaMemsetAsync((uint2*)sl.d_keys.p + col_begin, 0xFF, (key_rows - col_begin) * sizeof(uint2), ctx->stream));
    // small launches: the three tables travel as ONE copy (a copy command costs more than its few KB)
    const size_t b_tasks = (n_tasks * sizeof(MatchTask) + 255) & ~(size_t)255;
    const size_t b_tiles = (n_tiles * tile_bytes + merges.size() * sizeof(int4) + 255) & ~(size_t)255;
    const size_t b_pairs = ((size_t)n_pairs * sizeof(int2) + 255) & ~(size_t)255;
    const bool one_copy = b_tasks + b_tiles + b_pairs <= ((size_t)64 << 10);
    const MatchTask* t_tasks = (const MatchTask*)sl.d_tasks.p;
    const uint8_t* t_tiles = (const uint8_t*)sl.d_tiles.p;
    const int2* t_pair_tasks = (const int2*)sl.d_pair_tasks.p;
    if (one_copy) {
        UZ_CUDA(ctx, sl.h_tables.ensure(b_tasks + b_tiles + b_pairs));
        UZ_CUDA(ctx, sl.d_tables.ensure(b_tasks + b_tiles + b_pairs));
        uint8_t* hb = (uint8_t*)sl.h_tables.p;
        memcpy(hb, tasks, n_tasks * sizeof(MatchTask));
        memcpy(hb + b_tasks, tiles, n_tiles * tile_bytes + merges.size() * sizeof(int4));
        memcpy(hb + b_tasks + b_tiles, pair_tasks, (size_t)n_pairs * sizeof(int2));
        UZ_CUDA(ctx, cudaMemcpyAsync(sl.d_tables.p, hb, b_tasks + b_tiles + b_pairs, cudaMemcpyHostToDevice, ctx->stream));
        t_tasks = (const MatchTask*)sl.d_tables.p;
        t_tiles = (const uint8_t*)sl.d_tables.p + b_tasks;
        t_pair_tasks = (const int2*)((const uint8_t*)sl.d_tables.p + b_tasks + b_tiles);
    } else {
        if (n_tasks) UZ_CUDA(ctx, cudaMemcpyAsync(sl.d_tasks.p, tasks, n_tasks * sizeof(MatchTask), cudaMemcpyHostToDevice, ctx->stream));
        if (n_tiles) UZ_CUDA(ctx, cudaMemcpyAsync(sl.d_tiles.p, tiles, n_tiles * tile_bytes + merges.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
        UZ_CUDA(ctx, cudaMemcpyAsync(sl.d_pair_tasks.p, pair_tasks, (size_t)n_pairs * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    }

    // 3. launches.  Large batches: ONE match launch plus the persistent streaming solve beside it (uz_solve.cuh);
    // small batches, the parity taps and UZ_STREAM_SOLVE=0: match launch, then one solve CTA per pair behind it.
    SolveParams sp;
    memset(&sp, 0, sizeof(sp));
    if (with_solve) {
        uz_status st = ensure_samples(ctx, prm.ransac_iterations, prm.do_prosac, max_nq);
        if (st != UZ_OK) return st;
        sp.thr = prm.ransac_threshold; sp.thr_sq_star = thr_sq_star(prm.ransac_threshold);
        sp.break_pct = prm.break_percentage; sp.iterations = prm.ransac_iterations;
        sp.ratio_num = prm.ratio_num; sp.ratio_den = prm.ratio_den; sp.cap = cap;
        sp.samples = (const uint16_t*)ctx->d_samples.p; sp.samples_by_m = 1;
        ctx->dbg_pairs = 0;
        if (ctx->debug) {
            UZ_CUDA(ctx, ctx->d_dbg_matches.ensure((size_t)n_pairs * cap * 3 * sizeof(int32_t)));
            UZ_CUDA(ctx, ctx->d_dbg_mask.ensure((size_t)n_pairs * cap));
            UZ_CUDA(ctx, ctx->d_dbg_counts.ensure((size_t)n_pairs * prm.ransac_iterations * sizeof(int32_t)));
            UZ_CUDA(ctx, cudaMemsetAsync(ctx->d_dbg_counts.p, 0xFF, (size_t)n_pairs * prm.ransac_iterations * sizeof(int32_t), ctx->stream));
            sp.dbg_matches = (int32_t*)ctx->d_dbg_matches.p; sp.dbg_mask = (uint8_t*)ctx->d_dbg_mask.p;
            sp.dbg_counts = (int32_t*)ctx->d_dbg_counts.p;
            UZ_CUDA(ctx, ctx->d_dbg_phase.ensure((size_t)n_pairs * 8 * sizeof(long long)));
            UZ_CUDA(ctx, cudaMemsetAsync(ctx->d_dbg_phase.p, 0, (size_t)n_pairs * 8 * sizeof(long long), ctx->stream));
            sp.dbg_phase = (long long*)ctx->d_dbg_phase.p;
            ctx->dbg_cap = cap; ctx->dbg_pairs = n_pairs; ctx->dbg_iters = prm.ransac_iterations;
        }
    }
    // the streaming solve needs its 96-register CTAs to fit beside the match CTAs: one pair's on-chip state must
    // leave room for them (cap <= 1024: 44 KB; at cap 2048 the gain measured on C5 rigs was 1 %), and the batch must be large enough to be worth a persistent grid
    const bool streaming = may_stream && n_tiles > 0;
    if (ctx->timers) ctx->compares += compares;

    int* d_pending = nullptr;
    int* d_deferred = nullptr;
    StreamCtl* d_ctl = nullptr;
    if (streaming) {
        UZ_CUDA(ctx, sl.h_pending.ensure((size_t)n_pairs * sizeof(int)));
        UZ_CUDA(ctx, sl.d_pending.ensure((size_t)n_pairs * 2 * sizeof(int) + 256));
        int* pend = (int*)sl.h_pending.p;
        memset(pend, 0, (size_t)n_pairs * sizeof(int));
        for (size_t k = 0; k < n_tiles; ++k) pend[tasks[tiles[k].x].pair]++;
        d_ctl = (StreamCtl*)sl.d_pending.p;                       // control block first, counters 256 B behind it
        d_pending = (int*)((uint8_t*)sl.d_pending.p + 256);
        d_deferred = d_pending + n_pairs;
        UZ_CUDA(ctx, cudaMemsetAsync(d_ctl, 0, sizeof(StreamCtl), ctx->stream));
        UZ_CUDA(ctx, cudaMemcpyAsync(d_pending, pend, (size_t)n_pairs * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }

    uz_context::Timed tm;
    tm.e[0] = tm.e[1] = tm.e[2] = tm.e[3] = nullptr; tm.has_solve = false;
    if (ctx->timers) { tm.e[0] = ctx->get_event(); tm.e[1] = ctx->get_event(); cudaEventRecord(tm.e[0], ctx->stream); }
    cudaEvent_t ev_tables = nullptr;
    if (streaming) {            // the side grid may start once the tables, counters and everything before them are in place
        ev_tables = ctx->get_event();
        UZ_CUDA(ctx, cudaEventRecord(ev_tables, ctx->stream));
        if (ctx->stream_probe == 9) delay_kernel<<<1, 1, 0, ctx->stream>>>(3 * kStallNs);     // test hook: starve the streaming grid
    }
    if (n_tiles > 0) {
        const MatchTask* d_tk = t_tasks;
        const int2* d_t = (const int2*)t_tiles;
        const int2* d_tw = (const int2*)(t_tiles + n_tiles_narrow * tile_bytes);
        uint2* d_k = (uint2*)sl.d_keys.p;
        const int nt = (int)n_tiles_narrow;
        unsigned int* d_prog = d_ctl ? &d_ctl->progress : nullptr;
        if (n_tiles_wide > 0) {
            if (wide_cfg == 0) launch_knn2_wide<256>(ctx, d_tk, d_tw, (int)n_tiles_wide, d_k, d_pending, d_prog, fused, seg);
            else launch_knn2_wide<64>(ctx, d_tk, d_tw, (int)n_tiles_wide, d_k, d_pending, d_prog, fused, seg);
            ctx->launches++;
            UZ_CUDA(ctx, cudaGetLastError());
            if (ctx->timers) ctx->match_launches++;
        }
        if (nt > 0) switch (best_cfg) {
            case 0: launch_knn2<256, 4>(ctx, d_tk, d_t, nt, d_k, d_pending, d_prog, fused, seg); break;
            case 1: launch_knn2<128, 4>(ctx, d_tk, d_t, nt, d_k, d_pending, d_prog, fused, seg); break;
            case 2: launch_knn2<64, 2>(ctx, d_tk, d_t, nt, d_k, d_pending, d_prog, fused, seg); break;
            case 3: launch_knn2<256, 2>(ctx, d_tk, d_t, nt, d_k, d_pending, d_prog, fused, seg); break;
            case 4: launch_knn2<128, 2>(ctx, d_tk, d_t, nt, d_k, d_pending, d_prog, fused, seg); break;
            default: launch_knn2<32, 2>(ctx, d_tk, d_t, nt, d_k, d_pending, d_prog, fused, seg); break;
        }
        if (nt > 0) {
            ctx->launches++;
            UZ_CUDA(ctx, cudaGetLastError());
            if (ctx->timers) ctx->match_launches++;
        }
        if (!merges.empty()) {
            const int4* d_m = (const int4*)(t_tiles + std::max<size_t>(n_tiles, 1) * tile_bytes);
            merge_segments_kernel<<<dim3((unsigned)((max_nq + 255) / 256), (unsigned)merges.size(), 1), 256, 0, ctx->stream>>>(d_m, d_k);
            ctx->launches++;
            UZ_CUDA(ctx, cudaGetLastError());
        }
    }
    if (ctx->timers) cudaEventRecord(tm.e[1], ctx->stream);
    if (with_solve) {
        cudaStream_t sB = streaming ? ctx->solve_stream : ctx->stream;
        if (streaming) {
            UZ_CUDA(ctx, cudaStreamWaitEvent(sB, ev_tables, 0));
            ctx->event_pool.push_back(ev_tables);         // safe to recycle: the wait has captured it
        }
        if (ctx->timers) { tm.e[2] = ctx->get_event(); tm.e[3] = ctx->get_event(); tm.has_solve = true; cudaEventRecord(tm.e[2], sB); }
        sp.pair_base = 0;
        sp.dbg_skip = streaming ? ctx->stream_probe : 0;
        if (streaming)
            solve_stream_kernel<kSolveThreads><<<stream_ctas, kSolveThreads, solve_smem_bytes(cap), sB>>>(
                t_tasks, t_pair_tasks, (const uint2*)sl.d_keys.p, sp, d_results,
                n_pairs, d_pending, d_ctl, d_deferred, 0);
        else
            launch_solve(ctx, n_pairs, cap, sB, t_tasks, t_pair_tasks, (const uint2*)sl.d_keys.p, sp, d_results);
        ctx->launches++;
        UZ_CUDA(ctx, cudaGetLastError());
        if (ctx->timers) { cudaEventRecord(tm.e[3], sB); ctx->solve_launches++; }
        if (streaming) {
            // cleanup form, ordered behind the match kernel by an event: pairs the streaming grid deferred or never
            // drew (only if it starved - normally every CTA of this launch exits on its first look)
            cudaEvent_t ev_match = ctx->get_event();
            UZ_CUDA(ctx, cudaEventRecord(ev_match, ctx->stream));
            UZ_CUDA(ctx, cudaStreamWaitEvent(sB, ev_match, 0));
            ctx->event_pool.push_back(ev_match);
            solve_stream_kernel<kSolveThreads><<<4 * ctx->sm_count, kSolveThreads, solve_smem_bytes(cap), sB>>>(
                t_tasks, t_pair_tasks, (const uint2*)sl.d_keys.p, sp, d_results,
                n_pairs, d_pending, d_ctl, d_deferred, 1);
            ctx->launches++;
            UZ_CUDA(ctx, cudaGetLastError());
            if (join) {         // rejoin: everything the caller enqueues next on its stream sees the results
                cudaEvent_t ev = ctx->get_event();
                UZ_CUDA(ctx, cudaEventRecord(ev, sB));
                UZ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev, 0));
                ctx->event_pool.push_back(ev);
            } else {
                results_on = sB;
            }
        }
    }
    if (ctx->timers) ctx->pending.push_back(tm);
    UZ_CUDA(ctx, cudaEventRecord(sl.done, results_on));     // the slot's tables and keys are free once the solve is through
    sl.used = true;
    if (result_stream) *result_stream = results_on;
    return UZ_OK;
}

uz_status resolve_timers(uz_context* ctx) {
    if (ctx->pending.empty()) return UZ_OK;
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->side) UZ_CUDA(ctx, cudaStreamSynchronize(ctx->side));
    if (ctx->alt) UZ_CUDA(ctx, cudaStreamSynchronize(ctx->alt));
    if (ctx->solve_stream) UZ_CUDA(ctx, cudaStreamSynchronize(ctx->solve_stream));
    for (auto& t : ctx->pending) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, t.e[0], t.e[1]);
        ctx->match_ms += a;
        if (t.has_solve) { cudaEventElapsedTime(&b, t.e[2], t.e[3]); ctx->solve_ms += b; }
        for (int i = 0; i < 4; ++i) if (t.e[i]) ctx->event_pool.push_back(t.e[i]);
    }
    ctx->pending.clear();
    return UZ_OK;
}

uz_status check_ctx(uz_context* ctx) {
    if (!ctx) return UZ_ERR_INVALID;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(ctx, UZ_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return UZ_OK;
}

uz_status validate_params(uz_context* ctx, const uz_params* p) {
    if (p->ransac_iterations < 1 || p->ransac_iterations > UZ_MAX_ITERATIONS) return fail(ctx, UZ_ERR_INVALID, "ransac_iterations out of range");
    if (p->ratio_den <= 0 || p->ratio_num < 0 || p->ratio_den > 4096 || p->ratio_num > 4096) return fail(ctx, UZ_ERR_INVALID, "ratio out of range");
    if (p->cross_check != 0 && p->cross_check != 1) return fail(ctx, UZ_ERR_INVALID, "cross_check must be 0 or 1");
    if (p->min_keypoints < 0) return fail(ctx, UZ_ERR_INVALID, "min_keypoints < 0");
    return UZ_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

const char* uz_version(void) { return "uzliti_edge_b200 0.1 (sm_100a)"; }

void uz_default_params(uz_params* p) {
    if (!p) return;
    p->ransac_threshold = 0.1;     // iti_slam_launch/yaml/slam.yaml:35
    p->break_percentage = 0.6;     // cfg/FeatureLinkEstimation.cfg:12
    p->ransac_iterations = 100;    // slam.yaml:36
    p->do_prosac = 1;
    p->ratio_num = 99; p->ratio_den = 100;   // feature_transformation_estimator.cpp:67
    p->min_keypoints = 7;          // :47
    p->cross_check = 0;
}

uz_status uz_create(int32_t device, uz_context** out) {
    if (!out) return UZ_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) { cudaGetLastError(); return fail(nullptr, UZ_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e)); }
    if (device < 0 || device >= count) return fail(nullptr, UZ_ERR_INVALID, "device index out of range");
    if ((e = cudaSetDevice(device)) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, UZ_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { cudaGetLastError(); return fail(nullptr, UZ_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e)); }
    if (prop.major != 10) return fail(nullptr, UZ_ERR_CUDA, "device is not sm_100 (this library carries sm_100a code only)");
    uz_context* ctx = new uz_context();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    uz_default_params(&ctx->params);
    uz_default_place_params(&ctx->places.params);
    const char* v = getenv("UZ_KNN_VARIANT");
    if (v && v[0] == '1') ctx->variant_csa = 0;
    if (v && v[0] == '2') ctx->variant_pack16 = 0;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) { delete ctx; return fail(nullptr, UZ_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
    for (int i = 0; i < 4; ++i) cudaEventCreate(&ctx->ev[i]);
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if ((e = cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, hi)) != cudaSuccess) { uz_destroy(ctx); return fail(nullptr, UZ_ERR_CUDA, std::string("cudaStreamCreateWithPriority: ") + cudaGetErrorString(e)); }
        if ((e = cudaStreamCreateWithFlags(&ctx->alt, cudaStreamNonBlocking)) != cudaSuccess) { uz_destroy(ctx); return fail(nullptr, UZ_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
        const char* ac = getenv("UZ_ALT_CHUNKS");
        if (ac) ctx->alt_chunks = atoi(ac) != 0;
        if ((e = cudaStreamCreateWithPriority(&ctx->solve_stream, cudaStreamNonBlocking, hi)) != cudaSuccess) { uz_destroy(ctx); return fail(nullptr, UZ_ERR_CUDA, std::string("cudaStreamCreateWithPriority: ") + cudaGetErrorString(e)); }
        const char* mp = getenv("UZ_STREAM_SOLVE_MIN_PAIRS");
        if (mp && atoi(mp) > 0) ctx->stream_min_pairs = atoi(mp);
        const char* sq = getenv("UZ_STREAM_PROBE");
        if (sq) ctx->stream_probe = atoi(sq);
        const char* ss = getenv("UZ_STREAM_SOLVE");
        if (ss && atoi(ss) >= 0 && atoi(ss) <= 4) ctx->stream_solve_ctas = atoi(ss);
        const char* fc = getenv("UZ_KNN_CFG");
        if (fc) ctx->force_cfg = atoi(fc);
        const char* xf = getenv("UZ_XCHECK_FUSED");
        if (xf) ctx->xcheck_fused = atoi(xf) != 0;
        const char* sw = getenv("UZ_SOLVE_WIDE");
        if (sw) ctx->solve_wide = atoi(sw) != 0;
        const char* sg = getenv("UZ_SEGMENT");
        if (sg) ctx->segment_small = atoi(sg) != 0;
        const char* fw = getenv("UZ_KNN_WIDE_CFG");
        if (fw) ctx->force_wide_cfg = atoi(fw);
        const char* gu = getenv("UZ_GATHER_UPLOAD");
        if (gu) ctx->gather_upload = atoi(gu);
        const char* cc = getenv("UZ_COPY_CTAS");
        if (cc && atoi(cc) > 0) ctx->copy_ctas = atoi(cc);
        const char* hc = getenv("UZ_HOST_CHUNKS");
        if (hc) ctx->host_chunks = atoi(hc);
    }
    e = cudaFuncSetAttribute(solve_kernel<kSolveThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)solve_smem_bytes(UZ_MAX_FEATURES));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(solve_kernel<kSolveThreadsWide>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)solve_smem_bytes(UZ_MAX_FEATURES, kSolveThreadsWide));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(solve_stream_kernel<kSolveThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)solve_smem_bytes(1024));
    if (e == cudaSuccess) e = set_carveouts();
    if (e != cudaSuccess) { cudaGetLastError(); uz_destroy(ctx); return fail(nullptr, UZ_ERR_CUDA, std::string("cudaFuncSetAttribute(solve_kernel smem): ") + cudaGetErrorString(e)); }
    *out = ctx;
    return UZ_OK;
}

void uz_destroy(uz_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    places_release(ctx);
    ctx->store_arena.release(); ctx->transient.release();
    for (auto& sl : ctx->slots) {
        sl.d_tasks.release(); sl.d_tiles.release(); sl.d_pair_tasks.release(); sl.d_keys.release(); sl.d_pending.release(); sl.d_tables.release();
        sl.h_tasks.release(); sl.h_tiles.release(); sl.h_pair_tasks.release(); sl.h_pending.release(); sl.h_tables.release();
        if (sl.done) cudaEventDestroy(sl.done);
    }
    ctx->d_samples.release(); ctx->d_results.release(); ctx->d_dbg_matches.release(); ctx->d_dbg_mask.release(); ctx->d_dbg_counts.release(); ctx->d_dbg_phase.release(); ctx->d_chunks.release(); ctx->h_chunks.release(); ctx->h_results.release();
    ctx->d_misc.release();
    for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (auto& t : ctx->pending) for (int i = 0; i < 4; ++i) if (t.e[i]) cudaEventDestroy(t.e[i]);
    if (ctx->side) { cudaStreamSynchronize(ctx->side); cudaStreamDestroy(ctx->side); }
    if (ctx->alt) { cudaStreamSynchronize(ctx->alt); cudaStreamDestroy(ctx->alt); }
    if (ctx->solve_stream) { cudaStreamSynchronize(ctx->solve_stream); cudaStreamDestroy(ctx->solve_stream); }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* uz_last_error(const uz_context* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

uz_status uz_set_params(uz_context* ctx, const uz_params* p) {
    if (!ctx || !p) return UZ_ERR_INVALID;
    uz_status st = validate_params(ctx, p);
    if (st != UZ_OK) return st;
    ctx->params = *p;
    return UZ_OK;
}

uz_status uz_get_params(const uz_context* ctx, uz_params* p) {
    if (!ctx || !p) return UZ_ERR_INVALID;
    *p = ctx->params;
    return UZ_OK;
}

uz_status uz_set_stream(uz_context* ctx, void* cuda_stream) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (cuda_stream) { ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false; }
    else { UZ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    return UZ_OK;
}

// ---- store -------------------------------------------------------------------------------------------
uz_status uz_store_add_bulk(uz_context* ctx, const uz_features* cams, const int32_t* cams_per_keyframe,
                            int32_t n_keyframes, int32_t* handles_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (n_keyframes < 0 || (n_keyframes > 0 && (!cams_per_keyframe || !handles_out))) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    size_t total = 0;
    for (int i = 0; i < n_keyframes; ++i) {
        if (cams_per_keyframe[i] < 0) return fail(ctx, UZ_ERR_INVALID, "negative camera count");
        total += (size_t)cams_per_keyframe[i];
    }
    if (total > 0 && !cams) return fail(ctx, UZ_ERR_INVALID, "null camera array");
    std::vector<const uz_features*> feats(total);
    for (size_t i = 0; i < total; ++i) feats[i] = cams + i;
    std::vector<Cam> up;
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->h_chunks.reset(); ctx->d_chunks.reset();
    st = upload_cams(ctx, ctx->store_arena, feats, up);
    if (st != UZ_OK) return st;
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // host buffers are borrowed only for the call
    size_t k = 0;
    for (int i = 0; i < n_keyframes; ++i) {
        int32_t h;
        if (!ctx->free_handles.empty()) { h = ctx->free_handles.back(); ctx->free_handles.pop_back(); }
        else { h = (int32_t)ctx->kfs.size(); ctx->kfs.emplace_back(); }
        Keyframe& kf = ctx->kfs[h];
        kf.cams.assign(up.begin() + k, up.begin() + k + cams_per_keyframe[i]);
        for (const Cam& c : kf.cams) ctx->store_max_n = std::max(ctx->store_max_n, c.n);
        kf.live = true;
        k += (size_t)cams_per_keyframe[i];
        ctx->live++;
        handles_out[i] = h;
    }
    return UZ_OK;
}

uz_status uz_store_add(uz_context* ctx, const uz_features* cams, int32_t n_cams, int32_t* handle_out) {
    if (!handle_out) return UZ_ERR_INVALID;
    return uz_store_add_bulk(ctx, cams, &n_cams, 1, handle_out);
}

uz_status uz_store_remove(uz_context* ctx, int32_t handle) {
    if (!ctx) return UZ_ERR_INVALID;
    if (handle < 0 || handle >= (int32_t)ctx->kfs.size() || !ctx->kfs[handle].live) return fail(ctx, UZ_ERR_INVALID, "unknown keyframe handle");
    ctx->kfs[handle].live = false;
    ctx->kfs[handle].cams.clear();
    ctx->free_handles.push_back(handle);
    ctx->live--;
    return UZ_OK;
}

uz_status uz_store_clear(uz_context* ctx) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->kfs.clear(); ctx->free_handles.clear(); ctx->live = 0; ctx->store_max_n = 0;
    ctx->store_arena.reset();
    if ((st = places_reset(ctx)) != UZ_OK) return st;      // the recogniser's nodes point into the store
    return UZ_OK;
}

int32_t uz_store_size(const uz_context* ctx) { return ctx ? ctx->live : 0; }

int64_t uz_store_bytes(const uz_context* ctx) {
    if (!ctx) return 0;
    int64_t b = 0;
    for (const auto& c : ctx->store_arena.chunks) b += (int64_t)c.used;
    return b;
}

// ---- stage entry points ----------------------------------------------------------------------------
uz_status uz_match_knn2(uz_context* ctx, int32_t desc_bytes, const uint8_t* query, int32_t nq, int32_t q_stride,
                        const uint8_t* train, int32_t nt, int32_t t_stride, int32_t* idx_out, int32_t* dist_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (nq < 0 || nt < 0 || nq > 65535 || nt > 65535) return fail(ctx, UZ_ERR_INVALID, "nq/nt out of range (0..65535)");
    if (nq == 0) return UZ_OK;
    if (!query || !idx_out || !dist_out || (nt > 0 && !train)) return fail(ctx, UZ_ERR_INVALID, "null buffer");
    const int db = desc_width(desc_bytes);
    if (db == 0) return fail(ctx, UZ_ERR_UNSUPPORTED, "descriptor width must be 32 or 64 bytes");
    if (q_stride < db || (nt > 0 && t_stride < db)) return fail(ctx, UZ_ERR_INVALID, "descriptor stride < descriptor width");
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->transient.reset();
    // two position-less cameras
    Keyframe kq, kt;
    kq.cams.resize(1); kt.cams.resize(1);
    auto up = [&](const uint8_t* h, int n, int stride, Cam& c) -> uz_status {
        c.n = n; c.feature_type = UZ_FEATURE_ORB; c.dbytes = db;
        if (n == 0) return UZ_OK;
        const int halves = n * (db / 32);
        c.raw = (uint32_t*)ctx->transient.alloc((size_t)n * db);
        c.csa = (uint32_t*)ctx->transient.alloc((size_t)n * db);
        uint8_t* stage = (uint8_t*)ctx->transient.alloc((size_t)n * stride);
        if (!c.raw || !c.csa || !stage) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
        UZ_CUDA(ctx, cudaMemcpyAsync(stage, h, (size_t)(n - 1) * stride + db, cudaMemcpyHostToDevice, ctx->stream));
        pack_descriptors_kernel<<<(halves + 255) / 256, 256, 0, ctx->stream>>>(stage, halves, stride, c.raw, c.csa, db / 32);
        ctx->launches++;
        return UZ_OK;
    };
    if ((st = up(query, nq, q_stride, kq.cams[0])) != UZ_OK) return st;
    if ((st = up(train, nt, t_stride, kt.cams[0])) != UZ_OK) return st;
    uz_params saved = ctx->params;
    ctx->params.min_keypoints = 0;
    std::vector<PairRef> pairs(1, PairRef{kt.cams.data(), 1, kq.cams.data(), 1});
    st = run_pairs(ctx, pairs, nullptr);
    ctx->params = saved;
    if (st != UZ_OK) return st;
    UZ_CUDA(ctx, ctx->d_misc.ensure((size_t)nq * 16));
    int32_t* d_idx = (int32_t*)ctx->d_misc.p;
    int32_t* d_dist = d_idx + 2 * (size_t)nq;
    unpack_keys_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>((const uint2*)ctx->slots[ctx->cur_slot].d_keys.p, nq, d_idx, d_dist);
    ctx->launches++;
    UZ_CUDA(ctx, cudaGetLastError());
    UZ_CUDA(ctx, cudaMemcpyAsync(idx_out, d_idx, (size_t)nq * 8, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaMemcpyAsync(dist_out, d_dist, (size_t)nq * 8, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UZ_OK;
}

uz_status uz_sample_list(uz_context* ctx, int32_t M, int32_t iterations, int32_t do_prosac, int32_t* out) {
    if (!ctx || !out) return UZ_ERR_INVALID;
    if (M < 0 || M > 65535 || iterations < 1 || iterations > UZ_MAX_ITERATIONS) return fail(ctx, UZ_ERR_INVALID, "M/iterations out of range");
    std::vector<uint32_t> rnd;
    glibc_rand_stream(1, (size_t)iterations * (size_t)std::max(M, 1), rnd);
    std::vector<uint16_t> rows((size_t)(M + 1) * iterations * 3, 0);
    build_sample_rows(iterations, do_prosac != 0, M, M + 1, rnd, rows.data());
    const uint16_t* r = rows.data() + (size_t)M * iterations * 3;
    for (int i = 0; i < iterations * 3; ++i) out[i] = r[i];
    return UZ_OK;
}

uz_status uz_estimate_svd(uz_context* ctx, const double* P, const double* Q, int32_t M, double max_error,
                          int32_t iterations, double break_percentage, int32_t do_prosac, const int32_t* samples,
                          double* T16_out, int32_t* consensus_out, double* mse_out, uint8_t* inlier_mask_out,
                          int32_t* best_iteration_out, int32_t* iterations_run_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (M < 0 || M > UZ_MAX_FEATURES) return fail(ctx, UZ_ERR_INVALID, "M out of range");
    if (iterations < 1 || iterations > UZ_MAX_ITERATIONS) return fail(ctx, UZ_ERR_INVALID, "iterations out of range");
    if ((M > 0 && (!P || !Q)) || !T16_out || !consensus_out || !mse_out) return fail(ctx, UZ_ERR_INVALID, "null buffer");
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->transient.reset();
    const int cap = std::max(128, pow2ceil(std::max(M, 1)));
    double* dP = (double*)ctx->transient.alloc((size_t)std::max(M, 1) * 24);
    double* dQ = (double*)ctx->transient.alloc((size_t)std::max(M, 1) * 24);
    uint8_t* dmask = (uint8_t*)ctx->transient.alloc((size_t)cap);
    uz_edge_result* dres = (uz_edge_result*)ctx->transient.alloc(sizeof(uz_edge_result));
    uint16_t* dsamp = nullptr;
    if (!dP || !dQ || !dmask || !dres) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
    if (M > 0) {
        UZ_CUDA(ctx, cudaMemcpyAsync(dP, P, (size_t)M * 24, cudaMemcpyHostToDevice, ctx->stream));
        UZ_CUDA(ctx, cudaMemcpyAsync(dQ, Q, (size_t)M * 24, cudaMemcpyHostToDevice, ctx->stream));
    }
    UZ_CUDA(ctx, cudaMemsetAsync(dmask, 0, cap, ctx->stream));
    SolveParams sp;
    memset(&sp, 0, sizeof(sp));
    std::vector<uint16_t> s16;
    if (samples) {
        s16.resize((size_t)iterations * 3);
        for (size_t i = 0; i < s16.size(); ++i) {
            if (samples[i] < 0 || samples[i] >= std::max(M, 1)) return fail(ctx, UZ_ERR_INVALID, "sample index out of range");
            s16[i] = (uint16_t)samples[i];
        }
        dsamp = (uint16_t*)ctx->transient.alloc(s16.size() * 2);
        if (!dsamp) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
        UZ_CUDA(ctx, cudaMemcpyAsync(dsamp, s16.data(), s16.size() * 2, cudaMemcpyHostToDevice, ctx->stream));
        sp.samples = dsamp; sp.samples_by_m = 0;
    } else {
        st = ensure_samples(ctx, iterations, do_prosac, M);
        if (st != UZ_OK) return st;
        sp.samples = (const uint16_t*)ctx->d_samples.p; sp.samples_by_m = 1;
    }
    sp.thr = max_error; sp.thr_sq_star = thr_sq_star(max_error); sp.break_pct = break_percentage;
    sp.iterations = iterations; sp.ratio_num = 99; sp.ratio_den = 100; sp.cap = cap;
    sp.direct_P = dP; sp.direct_Q = dQ; sp.direct_M = M;
    sp.dbg_mask = dmask;
    launch_solve(ctx, 1, cap, ctx->stream, nullptr, nullptr, nullptr, sp, dres);
    ctx->launches++;
    UZ_CUDA(ctx, cudaGetLastError());
    uz_edge_result r;
    UZ_CUDA(ctx, cudaMemcpyAsync(&r, dres, sizeof(r), cudaMemcpyDeviceToHost, ctx->stream));
    if (inlier_mask_out && M > 0) UZ_CUDA(ctx, cudaMemcpyAsync(inlier_mask_out, dmask, (size_t)M, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(T16_out, r.T, sizeof(r.T));
    *consensus_out = r.consensus; *mse_out = r.mse;
    if (best_iteration_out) *best_iteration_out = r.best_iteration;
    if (iterations_run_out) *iterations_run_out = r.iterations_run;
    return UZ_OK;
}

}  // extern "C"

namespace {
__global__ void consensus_kernel(const double* __restrict__ P, const double* __restrict__ Q, int M, const double* __restrict__ T16,
                                 double thr_star, uint8_t* __restrict__ set, int32_t* __restrict__ count) {
    double T[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) T[i] = T16[i];
    int c = 0;
    for (int base = 0; base < M; base += blockDim.x) {
        const int i = base + threadIdx.x;
        bool in = false;
        if (i < M) {
            in = residual_sq(T, P[3 * i], P[3 * i + 1], P[3 * i + 2], Q[3 * i], Q[3 * i + 1], Q[3 * i + 2]) < thr_star;
            set[i] = in;
        }
        c += __syncthreads_count(in);
    }
    if (threadIdx.x == 0) *count = c;
}
}  // namespace

extern "C" {

uz_status uz_consensus3d(uz_context* ctx, const double* P, const double* Q, int32_t M, const double* T16,
                         double thresh, uint8_t* set_out, int32_t* count_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (M < 0 || M > (1 << 24) || !T16 || !count_out || (M > 0 && (!P || !Q || !set_out))) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->transient.reset();
    const size_t m1 = (size_t)std::max(M, 1);
    double* dP = (double*)ctx->transient.alloc(m1 * 24);
    double* dQ = (double*)ctx->transient.alloc(m1 * 24);
    double* dT = (double*)ctx->transient.alloc(128);
    uint8_t* dset = (uint8_t*)ctx->transient.alloc(m1);
    int32_t* dcount = (int32_t*)ctx->transient.alloc(4);
    if (!dP || !dQ || !dT || !dset || !dcount) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
    if (M > 0) {
        UZ_CUDA(ctx, cudaMemcpyAsync(dP, P, (size_t)M * 24, cudaMemcpyHostToDevice, ctx->stream));
        UZ_CUDA(ctx, cudaMemcpyAsync(dQ, Q, (size_t)M * 24, cudaMemcpyHostToDevice, ctx->stream));
    }
    UZ_CUDA(ctx, cudaMemcpyAsync(dT, T16, 128, cudaMemcpyHostToDevice, ctx->stream));
    consensus_kernel<<<1, 256, 0, ctx->stream>>>(dP, dQ, M, dT, thr_sq_star(thresh), dset, dcount);
    ctx->launches++;
    UZ_CUDA(ctx, cudaGetLastError());
    if (M > 0) UZ_CUDA(ctx, cudaMemcpyAsync(set_out, dset, (size_t)M, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaMemcpyAsync(count_out, dcount, 4, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UZ_OK;
}

// ---- after the path: cluster RANSAC of TransformationFilter::calcValidEdges, edge acceptance gate ----
// One estimateSVD per cluster in ONE launch (transformation_filter.cpp:266-275: estimateSVD(P, Q, T, consensus, mse,
// 0.3, 200, 1.0, false) then consensus3D(P, Q, T, 0.3) — the mask returned here IS that second call's set, both use the
// refitted T and the same threshold).
uz_status uz_estimate_svd_batch(uz_context* ctx, const double* P, const double* Q, const int32_t* offsets, int32_t n_problems,
                                double max_error, int32_t iterations, double break_percentage, int32_t do_prosac,
                                double* T16_out, int32_t* consensus_out, double* mse_out, uint8_t* inlier_mask_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (n_problems < 0 || (n_problems > 0 && (!offsets || !T16_out || !consensus_out || !mse_out))) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    if (iterations < 1 || iterations > UZ_MAX_ITERATIONS) return fail(ctx, UZ_ERR_INVALID, "iterations out of range");
    if (n_problems == 0) return UZ_OK;
    int max_m = 0;
    if (offsets[0] != 0) return fail(ctx, UZ_ERR_INVALID, "offsets[0] must be 0");
    for (int i = 0; i < n_problems; ++i) {
        const int m = offsets[i + 1] - offsets[i];
        if (m < 0 || m > UZ_MAX_FEATURES) return fail(ctx, UZ_ERR_INVALID, "problem size out of range");
        max_m = std::max(max_m, m);
    }
    const int total = offsets[n_problems];
    if (total > 0 && (!P || !Q)) return fail(ctx, UZ_ERR_INVALID, "null points");
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->transient.reset();
    const int cap = std::max(128, pow2ceil(std::max(max_m, 1)));
    double* dP = (double*)ctx->transient.alloc((size_t)std::max(total, 1) * 24);
    double* dQ = (double*)ctx->transient.alloc((size_t)std::max(total, 1) * 24);
    int32_t* doff = (int32_t*)ctx->transient.alloc((size_t)(n_problems + 1) * 4);
    uint8_t* dmask = (uint8_t*)ctx->transient.alloc((size_t)cap * n_problems);
    uz_edge_result* dres = (uz_edge_result*)ctx->transient.alloc(sizeof(uz_edge_result) * (size_t)n_problems);
    if (!dP || !dQ || !doff || !dmask || !dres) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
    if (total > 0) {
        UZ_CUDA(ctx, cudaMemcpyAsync(dP, P, (size_t)total * 24, cudaMemcpyHostToDevice, ctx->stream));
        UZ_CUDA(ctx, cudaMemcpyAsync(dQ, Q, (size_t)total * 24, cudaMemcpyHostToDevice, ctx->stream));
    }
    UZ_CUDA(ctx, cudaMemcpyAsync(doff, offsets, (size_t)(n_problems + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    UZ_CUDA(ctx, cudaMemsetAsync(dmask, 0, (size_t)cap * n_problems, ctx->stream));
    st = ensure_samples(ctx, iterations, do_prosac, max_m);
    if (st != UZ_OK) return st;
    SolveParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.samples = (const uint16_t*)ctx->d_samples.p; sp.samples_by_m = 1;
    sp.thr = max_error; sp.thr_sq_star = thr_sq_star(max_error); sp.break_pct = break_percentage;
    sp.iterations = iterations; sp.ratio_num = 99; sp.ratio_den = 100; sp.cap = cap;
    sp.direct_P = dP; sp.direct_Q = dQ; sp.direct_M = 0; sp.direct_offsets = doff;
    sp.dbg_mask = dmask;
    launch_solve(ctx, n_problems, cap, ctx->stream, nullptr, nullptr, nullptr, sp, dres);
    ctx->launches++;
    UZ_CUDA(ctx, cudaGetLastError());
    std::vector<uz_edge_result> r((size_t)n_problems);
    std::vector<uint8_t> hm(inlier_mask_out ? (size_t)cap * n_problems : 0);
    UZ_CUDA(ctx, cudaMemcpyAsync(r.data(), dres, sizeof(uz_edge_result) * (size_t)n_problems, cudaMemcpyDeviceToHost, ctx->stream));
    if (inlier_mask_out) UZ_CUDA(ctx, cudaMemcpyAsync(hm.data(), dmask, hm.size(), cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n_problems; ++i) {
        memcpy(T16_out + 16 * (size_t)i, r[i].T, sizeof(r[i].T));
        consensus_out[i] = r[i].consensus; mse_out[i] = r[i].mse;
        if (inlier_mask_out) memcpy(inlier_mask_out + offsets[i], hm.data() + (size_t)i * cap, (size_t)(offsets[i + 1] - offsets[i]));
    }
    return UZ_OK;
}

void uz_default_gate_params(uz_gate_params* g) {
    if (!g) return;
    g->min_matching_score = 20.0;      // iti_slam_launch/yaml/slam.yaml:27
    g->max_edge_distance_T = 1.5;      // slam.yaml:25
    g->max_edge_distance_R = 30.0;     // slam.yaml:26
}

uz_status uz_gate_edges_device(uz_context* ctx, const void* results_device, int32_t n, const uz_gate_params* gate,
                               void* accept_device, void* translation_norm_device, void* rotation_deg_device) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (n < 0 || !gate || (n > 0 && (!results_device || !accept_device))) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    if (n == 0) return UZ_OK;
    GateParams g{gate->min_matching_score, gate->max_edge_distance_T, gate->max_edge_distance_R};
    gate_edges_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>((const uz_edge_result*)results_device, n, g, (uint8_t*)accept_device,
                                                                 (double*)translation_norm_device, (double*)rotation_deg_device);
    ctx->launches++;
    UZ_CUDA(ctx, cudaGetLastError());
    return UZ_OK;
}

uz_status uz_gate_edges(uz_context* ctx, const uz_edge_result* results, int32_t n, const uz_gate_params* gate,
                        uint8_t* accept_out, double* translation_norm_out, double* rotation_deg_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (n < 0 || !gate || (n > 0 && (!results || !accept_out))) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    if (n == 0) return UZ_OK;
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->transient.reset();
    uz_edge_result* dres = (uz_edge_result*)ctx->transient.alloc(sizeof(uz_edge_result) * (size_t)n);
    uint8_t* dacc = (uint8_t*)ctx->transient.alloc((size_t)n);
    double* dt = (double*)ctx->transient.alloc((size_t)n * 8);
    double* dr = (double*)ctx->transient.alloc((size_t)n * 8);
    if (!dres || !dacc || !dt || !dr) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
    UZ_CUDA(ctx, cudaMemcpyAsync(dres, results, sizeof(uz_edge_result) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    st = uz_gate_edges_device(ctx, dres, n, gate, dacc, dt, dr);
    if (st != UZ_OK) return st;
    UZ_CUDA(ctx, cudaMemcpyAsync(accept_out, dacc, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (translation_norm_out) UZ_CUDA(ctx, cudaMemcpyAsync(translation_norm_out, dt, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (rotation_deg_out) UZ_CUDA(ctx, cudaMemcpyAsync(rotation_deg_out, dr, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UZ_OK;
}

// ---- the batched path ----------------------------------------------------------------------------------
static uz_status pairs_from_handles(uz_context* ctx, const int32_t* from_handles, const int32_t* to_handles,
                                    int32_t n_pairs, std::vector<PairRef>& pairs) {
    if (n_pairs < 0 || (n_pairs > 0 && (!from_handles || !to_handles))) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    pairs.resize((size_t)n_pairs);
    const int32_t nk = (int32_t)ctx->kfs.size();
    for (int i = 0; i < n_pairs; ++i) {
        const int32_t a = from_handles[i], b = to_handles[i];
        if (a < 0 || a >= nk || b < 0 || b >= nk || !ctx->kfs[a].live || !ctx->kfs[b].live)
            return fail(ctx, UZ_ERR_INVALID, "unknown keyframe handle in pair list");
        pairs[i] = PairRef{ctx->kfs[a].cams.data(), (int)ctx->kfs[a].cams.size(), ctx->kfs[b].cams.data(), (int)ctx->kfs[b].cams.size()};
    }
    return UZ_OK;
}

uz_status uz_estimate_edges_device(uz_context* ctx, const int32_t* from_handles, const int32_t* to_handles,
                                   int32_t n_pairs, void* results_device) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (n_pairs > 0 && !results_device) return fail(ctx, UZ_ERR_INVALID, "null results");
    std::vector<PairRef> pairs;
    st = pairs_from_handles(ctx, from_handles, to_handles, n_pairs, pairs);
    if (st != UZ_OK) return st;
    return run_pairs(ctx, pairs, (uz_edge_result*)results_device);   // asynchronous: staging is double buffered
}

uz_status uz_estimate_edges(uz_context* ctx, const int32_t* from_handles, const int32_t* to_handles,
                            int32_t n_pairs, uz_edge_result* results) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (n_pairs > 0 && !results) return fail(ctx, UZ_ERR_INVALID, "null results");
    if (n_pairs <= 0) return n_pairs == 0 ? UZ_OK : fail(ctx, UZ_ERR_INVALID, "n_pairs < 0");
    UZ_CUDA(ctx, ctx->d_results.ensure((size_t)n_pairs * sizeof(uz_edge_result)));
    st = uz_estimate_edges_device(ctx, from_handles, to_handles, n_pairs, ctx->d_results.p);
    if (st != UZ_OK) return st;
    UZ_CUDA(ctx, cudaMemcpyAsync(results, ctx->d_results.p, (size_t)n_pairs * sizeof(uz_edge_result), cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UZ_OK;
}

uz_status uz_estimate_edges_host(uz_context* ctx, const uz_features* from_cams, const int32_t* n_from,
                                 const uz_features* to_cams, const int32_t* n_to, int32_t n_pairs,
                                 uz_edge_result* results) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (n_pairs < 0 || (n_pairs > 0 && (!n_from || !n_to || !results))) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    if (n_pairs == 0) return UZ_OK;
    Trace tr("estimate_edges_host");
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->transient.reset();
    ctx->h_chunks.reset(); ctx->d_chunks.reset();
    size_t tf = 0, tt = 0;
    for (int i = 0; i < n_pairs; ++i) {
        if (n_from[i] < 0 || n_to[i] < 0) return fail(ctx, UZ_ERR_INVALID, "negative camera count");
        tf += (size_t)n_from[i]; tt += (size_t)n_to[i];
    }
    if ((tf && !from_cams) || (tt && !to_cams)) return fail(ctx, UZ_ERR_INVALID, "null camera array");

    // The batch is cut into chunks of pairs and every chunk goes through the whole host pipeline on its own -
    // intern its cameras, enqueue their upload on the high-priority side stream, enqueue match + solve on the main
    // stream behind the upload's event, queue its records' way home - so the GPU starts after the host has looked at
    // the FIRST chunk only, uploads of later chunks run beside the matching of earlier ones, and records are copied
    // out to the caller while later chunks still compute.  What stays exposed is the first chunk's upload.
    // Chunk ends as fractions of the batch: equal chunks of about 768 pairs, between 4 and 40 of them.  Consecutive chunks
    // compute on two alternating streams (below), so a chunk boundary costs next to nothing and small chunks win: the
    // exposed first upload shrinks and records go home earlier, until the chunks get too small for the persistent solve
    // grid.  Measured on C4 (25 000 pairs, store-resident 937 k edges/s): 8 chunks 894 k, 16: 916 k, 32: 926 k, 48: 930 k,
    // 64: 903 k, 80: 839 k; on one stream the same 9-chunk split that gave 869 k gives 909 k on two.
    // UZ_HOST_CHUNKS = k > 0 forces k equal parts.
    std::vector<double> fracs;
    int want_chunks = 1;
    if (ctx->debug) want_chunks = 1;                            // the parity taps describe ONE launch pair
    else if (ctx->host_chunks > 0) want_chunks = std::min(ctx->host_chunks, n_pairs);
    else if (n_pairs < 2048) want_chunks = 1;
    else want_chunks = std::min(40, std::max(4, n_pairs / 768));
    for (int c = 1; c <= want_chunks; ++c) fracs.push_back((double)c / want_chunks);
    const int n_chunks = (int)fracs.size();
    std::vector<size_t> chunk_pair_end((size_t)n_chunks);
    for (int c = 0; c < n_chunks; ++c)
        chunk_pair_end[c] = c + 1 == n_chunks ? (size_t)n_pairs : std::min<size_t>((size_t)n_pairs, (size_t)(fracs[c] * n_pairs + 0.5));

    // unique cameras (a keyframe that appears in many pairs - one query vs many candidates - is uploaded once),
    // numbered in order of first use so that every chunk uploads exactly the cameras nobody before it needed
    struct CamKey {
        const void* d; const void* p; const void* v; int32_t n, stride, bytes, type, frame;
        bool operator==(const CamKey& o) const {
            return d == o.d && p == o.p && v == o.v && n == o.n && stride == o.stride && bytes == o.bytes && type == o.type &&
                   frame == o.frame;
        }
    };
    struct CamKeyHash {
        size_t operator()(const CamKey& k) const {
            uint64_t h = (uint64_t)(uintptr_t)k.d * 0x9E3779B97F4A7C15ull;
            h ^= ((uint64_t)(uintptr_t)k.p + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
            h ^= (uint64_t)(uint32_t)k.n * 0x165667B19E3779F9ull + (uint64_t)(uint32_t)k.frame;
            return (size_t)(h ^ (h >> 29));
        }
    };
    std::unordered_map<CamKey, uint32_t, CamKeyHash> seen;
    seen.reserve(std::min<size_t>(tf + tt, (size_t)1 << 20) * 2);
    std::vector<const uz_features*> uniq;
    std::vector<Cam> up;                       // device views of the unique cameras, in order of first use
    auto intern = [&](const uz_features* f) -> uint32_t {
        const CamKey k{f->descriptors, f->positions, f->valid_3d, f->n, f->desc_stride, f->desc_bytes, f->feature_type, f->sensor_frame};
        auto it = seen.find(k);
        if (it != seen.end()) return it->second;
        const uint32_t id = (uint32_t)uniq.size();
        seen.emplace(k, id);
        uniq.push_back(f);
        return id;
    };

    cudaStream_t main_stream = ctx->stream;
    const bool piped = n_chunks > 1 && ctx->side != nullptr;
    if (piped) {            // the side stream must not run ahead of work the caller queued on the main stream
        cudaEvent_t ev = ctx->get_event();
        UZ_CUDA(ctx, cudaEventRecord(ev, main_stream));
        UZ_CUDA(ctx, cudaStreamWaitEvent(ctx->side, ev, 0));
        if (ctx->alt) UZ_CUDA(ctx, cudaStreamWaitEvent(ctx->alt, ev, 0));
        ctx->event_pool.push_back(ev);
    }
    const bool alternate = piped && ctx->alt != nullptr && ctx->alt_chunks;
    UZ_CUDA(ctx, ctx->d_results.ensure((size_t)n_pairs * sizeof(uz_edge_result)));
    UZ_CUDA(ctx, ctx->h_results.ensure((size_t)n_pairs * sizeof(uz_edge_result)));
    uz_edge_result* h_res = (uz_edge_result*)ctx->h_results.p;
    std::vector<cudaEvent_t> home((size_t)n_chunks, nullptr);
    std::vector<char> copied((size_t)n_chunks, 0);
    std::vector<uint32_t> from_u, to_u;
    std::vector<Cam> fcams, tcams;
    std::vector<PairRef> part;
    auto chunk_begin = [&](int c) { return c ? chunk_pair_end[c - 1] : (size_t)0; };
    // records of finished chunks go out to the caller while the GPU works on later ones
    auto drain = [&](int upto, bool wait) {
        for (int d = 0; d < upto && st == UZ_OK; ++d) {
            if (copied[d] || !home[d]) continue;
            if (!wait && cudaEventQuery(home[d]) != cudaSuccess) { cudaGetLastError(); break; }
            if (wait && cudaEventSynchronize(home[d]) != cudaSuccess) { st = fail(ctx, UZ_ERR_CUDA, "cudaEventSynchronize failed"); break; }
            memcpy(results + chunk_begin(d), h_res + chunk_begin(d), (chunk_pair_end[d] - chunk_begin(d)) * sizeof(uz_edge_result));
            copied[d] = 1;
        }
    };
    size_t cf = 0, ct = 0;
    cudaEvent_t last_ready = nullptr;
    for (int c = 0; c < n_chunks && st == UZ_OK; ++c) {
        const size_t p0 = chunk_begin(c), p1 = chunk_pair_end[c];
        if (p1 <= p0) continue;
        // (1) intern the chunk's cameras
        const size_t u0 = uniq.size(), cf0 = cf, ct0 = ct;
        from_u.clear(); to_u.clear();
        for (size_t i = p0; i < p1; ++i) {
            for (int k = 0; k < n_from[i]; ++k, ++cf) from_u.push_back(intern(from_cams + cf));
            for (int k = 0; k < n_to[i]; ++k, ++ct) to_u.push_back(intern(to_cams + ct));
        }
        // (2) upload (+ CSA pass) of the cameras nobody before this chunk needed, on the side stream
        cudaEvent_t ready = nullptr;
        if (uniq.size() > u0) {
            std::vector<const uz_features*> fresh(uniq.begin() + u0, uniq.end());
            std::vector<Cam> got;
            if (piped) ctx->stream = ctx->side;
            ctx->copy_beside_compute = piped && c > 0;       // chunk 0 has the chip to itself
            st = upload_cams(ctx, ctx->transient, fresh, got);
            ctx->copy_beside_compute = 0;
            if (st == UZ_OK && piped) {
                ready = ctx->get_event();
                if (cudaEventRecord(ready, ctx->stream) != cudaSuccess) st = fail(ctx, UZ_ERR_CUDA, "cudaEventRecord failed");
            }
            ctx->stream = main_stream;
            if (st != UZ_OK) { if (ready) ctx->event_pool.push_back(ready); break; }
            up.insert(up.end(), got.begin(), got.end());
        }
        // (3) match + solve behind the upload, records home behind the solve
        fcams.resize(cf - cf0); tcams.resize(ct - ct0);
        for (size_t i = 0; i < fcams.size(); ++i) fcams[i] = up[from_u[i]];
        for (size_t i = 0; i < tcams.size(); ++i) tcams[i] = up[to_u[i]];
        part.resize(p1 - p0);
        {
            size_t a = 0, b2 = 0;
            for (size_t i = p0; i < p1; ++i) {
                part[i - p0] = PairRef{fcams.data() + a, n_from[i], tcams.data() + b2, n_to[i]};
                a += (size_t)n_from[i]; b2 += (size_t)n_to[i];
            }
        }
        // odd chunks compute on the second stream: two match kernels of one stream run strictly one after the other, and
        // every such boundary leaves the SMs partly idle while the last CTAs of the earlier kernel finish
        if (alternate && (c & 1)) ctx->stream = ctx->alt;
        // the upload stream is in order, so the newest upload event covers every camera uploaded so far; a chunk on the
        // other compute stream needs it even when it brought no camera of its own
        if (ready) {
            if (last_ready) ctx->event_pool.push_back(last_ready);
            last_ready = ready;
        }
        if (last_ready && (ready || alternate)) cudaStreamWaitEvent(ctx->stream, last_ready, 0);
        cudaStream_t rs = ctx->stream;
        st = run_pairs(ctx, part, (uz_edge_result*)ctx->d_results.p + p0, /*join=*/false, &rs);
        ctx->stream = main_stream;
        // into pinned memory: a pageable destination would make the copy synchronous and stall the next chunk's enqueue
        if (st == UZ_OK && cudaMemcpyAsync(h_res + p0, (uz_edge_result*)ctx->d_results.p + p0, (p1 - p0) * sizeof(uz_edge_result),
                                           cudaMemcpyDeviceToHost, rs) != cudaSuccess)
            st = fail(ctx, UZ_ERR_CUDA, "cudaMemcpyAsync(results) failed");
        if (st == UZ_OK) { home[c] = ctx->get_event(); cudaEventRecord(home[c], rs); }
        if (c == 0) tr.lap("first chunk enqueued");
        drain(c, false);
    }
    tr.lap("all chunks enqueued");
    drain(n_chunks, true);
    for (auto e : home) if (e) ctx->event_pool.push_back(e);
    if (last_ready) ctx->event_pool.push_back(last_ready);
    if (st != UZ_OK) {
        cudaStreamSynchronize(ctx->stream);
        if (ctx->side) cudaStreamSynchronize(ctx->side);
        if (ctx->alt) cudaStreamSynchronize(ctx->alt);
        if (ctx->solve_stream) cudaStreamSynchronize(ctx->solve_stream);
        cudaGetLastError();
        return st;
    }
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->alt) UZ_CUDA(ctx, cudaStreamSynchronize(ctx->alt));
    if (ctx->solve_stream) UZ_CUDA(ctx, cudaStreamSynchronize(ctx->solve_stream));
    tr.lap("wait GPU + records out");
    return UZ_OK;
}

uz_status uz_set_debug(uz_context* ctx, int32_t enable) {
    if (!ctx) return UZ_ERR_INVALID;
    ctx->debug = enable ? 1 : 0;
    if (!enable) ctx->dbg_pairs = 0;
    return UZ_OK;
}

uz_status uz_debug_pair(uz_context* ctx, int32_t pair_index, int32_t* matches_out, uint8_t* inlier_mask_out,
                        int32_t capacity, int32_t* n_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (pair_index < 0 || pair_index >= ctx->dbg_pairs) return fail(ctx, UZ_ERR_INVALID, "no debug data for that pair (uz_set_debug before the call)");
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int cap = ctx->dbg_cap;
    const int n = std::min(capacity, cap);
    if (matches_out && n > 0)
        UZ_CUDA(ctx, cudaMemcpyAsync(matches_out, (int32_t*)ctx->d_dbg_matches.p + (size_t)pair_index * cap * 3, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if (inlier_mask_out && n > 0)
        UZ_CUDA(ctx, cudaMemcpyAsync(inlier_mask_out, (uint8_t*)ctx->d_dbg_mask.p + (size_t)pair_index * cap, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_out) *n_out = n;
    return UZ_OK;
}

uz_status uz_debug_counts(uz_context* ctx, int32_t pair_index, int32_t* counts_out, int32_t capacity, int32_t* n_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (pair_index < 0 || pair_index >= ctx->dbg_pairs || !counts_out) return fail(ctx, UZ_ERR_INVALID, "no debug data for that pair (uz_set_debug before the call)");
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int n = std::min(capacity, ctx->dbg_iters);
    if (n > 0)
        UZ_CUDA(ctx, cudaMemcpyAsync(counts_out, (int32_t*)ctx->d_dbg_counts.p + (size_t)pair_index * ctx->dbg_iters, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_out) *n_out = n;
    return UZ_OK;
}

uz_status uz_debug_phases(uz_context* ctx, int32_t pair_index, int64_t* clocks8_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (pair_index < 0 || pair_index >= ctx->dbg_pairs || !clocks8_out) return fail(ctx, UZ_ERR_INVALID, "no debug data for that pair (uz_set_debug before the call)");
    UZ_CUDA(ctx, cudaMemcpyAsync(clocks8_out, (long long*)ctx->d_dbg_phase.p + (size_t)pair_index * 8, 64, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UZ_OK;
}

// ---- introspection -------------------------------------------------------------------------------------
uz_status uz_set_stream_solve(uz_context* ctx, int32_t ctas_per_sm) {
    if (!ctx) return UZ_ERR_INVALID;
    if (ctas_per_sm < 0 || ctas_per_sm > 4) return fail(ctx, UZ_ERR_INVALID, "ctas_per_sm must be 0..4");
    ctx->stream_solve_ctas = ctas_per_sm;
    return UZ_OK;
}

int64_t uz_launch_count(const uz_context* ctx) { return ctx ? ctx->launches : 0; }

uz_status uz_enable_timers(uz_context* ctx, int32_t enable) {
    if (!ctx) return UZ_ERR_INVALID;
    ctx->timers = enable ? 1 : 0;
    return UZ_OK;
}

uz_status uz_reset_timers(uz_context* ctx) {
    if (!ctx) return UZ_ERR_INVALID;
    uz_status st = resolve_timers(ctx);
    if (st != UZ_OK) return st;
    ctx->match_ms = ctx->solve_ms = 0; ctx->match_launches = ctx->solve_launches = ctx->compares = 0;
    return UZ_OK;
}

uz_status uz_get_timers(uz_context* ctx, double* match_ms, double* solve_ms, int64_t* match_launches,
                        int64_t* solve_launches, int64_t* descriptor_compares) {
    if (!ctx) return UZ_ERR_INVALID;
    uz_status st = resolve_timers(ctx);
    if (st != UZ_OK) return st;
    if (match_ms) *match_ms = ctx->match_ms;
    if (solve_ms) *solve_ms = ctx->solve_ms;
    if (match_launches) *match_launches = ctx->match_launches;
    if (solve_launches) *solve_launches = ctx->solve_launches;
    if (descriptor_compares) *descriptor_compares = ctx->compares;
    return UZ_OK;
}

uz_status uz_microbench(uz_context* ctx, int32_t op, double* gops_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (!gops_out || op < 0 || op > 3) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    UZ_CUDA(ctx, ctx->d_misc.ensure(256));
    const int blocks = ctx->sm_count * 8, iters = 4096;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(ctx->ev[0], ctx->stream);
        switch (op) {
            case 0: intpipe_bench_kernel<0><<<blocks, 256, 0, ctx->stream>>>((uint32_t*)ctx->d_misc.p, 12345u + rep, iters); break;
            case 1: intpipe_bench_kernel<1><<<blocks, 256, 0, ctx->stream>>>((uint32_t*)ctx->d_misc.p, 12345u + rep, iters); break;
            case 2: intpipe_bench_kernel<2><<<blocks, 256, 0, ctx->stream>>>((uint32_t*)ctx->d_misc.p, 12345u + rep, iters); break;
            default: intpipe_bench_kernel<3><<<blocks, 256, 0, ctx->stream>>>((uint32_t*)ctx->d_misc.p, 12345u + rep, iters); break;
        }
        ctx->launches++;
        cudaEventRecord(ctx->ev[1], ctx->stream);
        UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        if (rep > 0) best = std::min(best, ms);
    }
    const double ops = (double)blocks * 256.0 * iters * 32.0;
    *gops_out = ops / (best * 1e-3) * 1e-9;
    return UZ_OK;
}

}  // extern "C"

#include "uz_capi_places.inl"
#include "uz_capi_ingest.inl"
