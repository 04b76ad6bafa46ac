// uz_group.inl — the batched path on several GPUs of one box behind the same C-ABI (SURVEY.md 8e, 8b(1) "device list").
//
// The reference owns ONE worker thread per estimator (transformation_estimator.cpp:26); pairs are independent, so here one
// process owns one context per device and one host worker thread per context:
//   * store: replicated.  A keyframe goes up ONCE over PCIe (to the first device); every other device pulls the three
//     primary arrays (descriptor rows, positions, valid flags) from that device's HBM over NVLink with the gather kernel
//     reading peer memory, and derives the CSA / E8 layouts locally.  Handles are the same on every device.
//   * batch: the from-sorted pair list is cut into contiguous shards (shard_bounds, the rule of
//     uzliti_slam_b200/sharding.py); every device runs its shard.
//   * gather: no collective.  Results wanted in device memory (uz_group_estimate_edges_device): FUSED INTO THE SOLVE - each
//     device's solve kernel writes its 176-byte records (eleven 16-byte stores) straight into the first device's buffer at the
//     pair's batch-wide index through a peer-mapped pointer over NVLink.  Results wanted on the host
//     (uz_group_estimate_edges): every device copies its shard of records into one pinned array over its OWN PCIe link
//     (gather mode 1, default); the solve kernels writing through a host-mapped pointer (mode 0) is the measured
//     alternative - small posted writes over PCIe cost more than one DMA per device (2 GPUs, 50 000 pairs: 9.7 against
//     11-20 ms).
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

struct uz_group {
    struct Worker {
        std::thread th;
        std::mutex m;
        std::condition_variable cv;
        std::function<uz_status()> job;
        bool has_job = false, done = true, quit = false;
        uz_status status = UZ_OK;
    };
    std::vector<uz_context*> ctx;
    std::vector<Worker*> workers;
    std::string err;
    int gather_mode = 1;              // host results: 1 = local records + one copy per device (default, measured faster), 0 = host-mapped sink
    PinBuf h_results;                 // portable + mapped: every device may write into it
    uz_edge_result* h_results_dev = nullptr;
    std::vector<double> last_ms;      // device time of the last batch per rank
    std::vector<cudaEvent_t> ev0, ev1;
    bool in_flight = false;           // between uz_group_estimate_edges_begin and _end
};

namespace {

void group_worker_loop(uz_group::Worker* w) {
    std::unique_lock<std::mutex> lk(w->m);
    for (;;) {
        w->cv.wait(lk, [w] { return w->has_job || w->quit; });
        if (w->quit) return;
        std::function<uz_status()> job;
        job.swap(w->job);
        w->has_job = false;
        lk.unlock();
        const uz_status st = job();
        lk.lock();
        w->status = st;
        w->done = true;
        w->cv.notify_all();
    }
}

// hands fn(rank) to every device's worker thread
void group_dispatch(uz_group* g, const std::function<uz_status(int)>& fn, int first_rank = 0) {
    const int n = (int)g->ctx.size();
    for (int r = first_rank; r < n; ++r) {
        uz_group::Worker* w = g->workers[r];
        std::lock_guard<std::mutex> lk(w->m);
        w->job = [fn, r]() { return fn(r); };
        w->has_job = true; w->done = false;
        w->cv.notify_all();
    }
}
// waits for the workers; returns the first failure
uz_status group_wait(uz_group* g, int first_rank = 0) {
    const int n = (int)g->ctx.size();
    uz_status out = UZ_OK;
    for (int r = first_rank; r < n; ++r) {
        uz_group::Worker* w = g->workers[r];
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [w] { return w->done; });
        if (w->status != UZ_OK && out == UZ_OK) {
            out = w->status;
            g->err = "device " + std::to_string(g->ctx[r]->device) + ": " + g->ctx[r]->err;
        }
    }
    return out;
}
// runs fn(rank) on every device's worker thread and waits; returns the first failure
uz_status group_run(uz_group* g, const std::function<uz_status(int)>& fn, int first_rank = 0) {
    group_dispatch(g, fn, first_rank);
    return group_wait(g, first_rank);
}

void group_shard(int n_pairs, int world, int rank, int* lo, int* hi) {
    const int base = n_pairs / world, rem = n_pairs % world;
    *lo = rank * base + std::min(rank, rem);
    *hi = *lo + base + (rank < rem ? 1 : 0);
}

// replica side of a store add: same layout as on the first device, primary arrays pulled from its HBM
uz_status group_replicate(uz_context* ctx, const uz_context* src, const int32_t* handles, int32_t n_keyframes, bool replace = false) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->h_chunks.reset(); ctx->d_chunks.reset();
    std::vector<Cam> up;
    std::vector<BlockRef> blocks;
    std::vector<int32_t> counts((size_t)n_keyframes);
    std::vector<CopyChunk> cc;
    for (int i = 0; i < n_keyframes; ++i) {
        const Keyframe& sk = src->kfs[handles[i]];
        counts[i] = (int32_t)sk.cams.size();
        size_t at = 0;
        std::vector<CamLayout> lay(sk.cams.size());
        for (size_t c = 0; c < sk.cams.size(); ++c) { lay[c] = cam_layout(at, sk.cams[c].n, sk.cams[c].dbytes, ctx->operand_fmt()); at = lay[c].end; }
        uint8_t* base = (uint8_t*)ctx->store_arena.alloc(std::max<size_t>(at, 1));
        if (!base) {
            for (auto& b : blocks) ctx->store_arena.free(b.p, b.bytes);
            return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
        }
        blocks.push_back(BlockRef{base, std::max<size_t>(at, 1)});
        for (size_t c = 0; c < sk.cams.size(); ++c) {
            const Cam& s = sk.cams[c];
            Cam o = s;
            if (s.n > 0) {
                o.raw = (uint32_t*)(base + lay[c].raw); o.pos = (double*)(base + lay[c].pos); o.valid = base + lay[c].valid;
                o.csa = (uint32_t*)(base + lay[c].csa);
                o.e8 = base + lay[c].e8;
                push_chunks(cc, (uintptr_t)s.raw, (uint8_t*)o.raw, (size_t)s.n * s.dbytes);      // peer reads over NVLink
                push_chunks(cc, (uintptr_t)s.pos, (uint8_t*)o.pos, (size_t)s.n * 24);
                push_chunks(cc, (uintptr_t)s.valid, o.valid, (size_t)s.n);
            }
            up.push_back(o);
        }
    }
    if ((st = launch_gather(ctx, cc)) == UZ_OK) st = derive_layouts(ctx, up.data(), up.size());
    if (st == UZ_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, UZ_ERR_CUDA, "replication over peer memory failed");
    if (st != UZ_OK) {
        cudaStreamSynchronize(ctx->stream); cudaGetLastError();
        for (auto& b : blocks) ctx->store_arena.free(b.p, b.bytes);
        return st;
    }
    if (replace) {                    // uz_group_store_replace: one keyframe, the handle stays
        if (handles[0] < 0 || handles[0] >= (int32_t)ctx->kfs.size() || !ctx->kfs[handles[0]].live) {
            ctx->store_arena.free(blocks[0].p, blocks[0].bytes);
            return fail(ctx, UZ_ERR_INVALID, "replicated store out of step: unknown keyframe handle");
        }
        swap_keyframe(ctx, handles[0], up, blocks[0]);
        return UZ_OK;
    }
    std::vector<int32_t> got((size_t)n_keyframes);
    register_keyframes(ctx, up, blocks, counts.data(), n_keyframes, got.data());
    for (int i = 0; i < n_keyframes; ++i)
        if (got[i] != handles[i]) return fail(ctx, UZ_ERR_INVALID, "replicated store out of step: handle differs from the first device's");
    return UZ_OK;
}

}  // namespace

extern "C" {

uz_status uz_group_create(const int32_t* devices, int32_t n_devices, uz_group** out) {
    if (!out) return UZ_ERR_INVALID;
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > 64) return fail(nullptr, UZ_ERR_INVALID, "device list empty or too long");
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return fail(nullptr, UZ_ERR_INVALID, "device listed twice");
    uz_group* g = new uz_group();
    for (int i = 0; i < n_devices; ++i) {
        uz_context* c = nullptr;
        const uz_status st = uz_create(devices[i], &c);
        if (st != UZ_OK) { uz_group_destroy(g); return st; }
        g->ctx.push_back(c);
    }
    // peer access in both directions between the first device (store source, record sink) and every other one
    for (int i = 1; i < n_devices; ++i) {
        int can = 0, can2 = 0;
        cudaDeviceCanAccessPeer(&can, devices[i], devices[0]);
        cudaDeviceCanAccessPeer(&can2, devices[0], devices[i]);
        if (!can || !can2) { uz_group_destroy(g); return fail(nullptr, UZ_ERR_CUDA, "devices of a group must have peer access to the first one (NVLink / NVSwitch)"); }
        cudaSetDevice(devices[i]);
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); uz_group_destroy(g); return fail(nullptr, UZ_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
        cudaGetLastError();
        cudaSetDevice(devices[0]);
        e = cudaDeviceEnablePeerAccess(devices[i], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); uz_group_destroy(g); return fail(nullptr, UZ_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
        cudaGetLastError();
    }
    g->last_ms.assign((size_t)n_devices, 0.0);
    g->ev0.assign((size_t)n_devices, nullptr); g->ev1.assign((size_t)n_devices, nullptr);
    for (int i = 0; i < n_devices; ++i) {
        cudaSetDevice(devices[i]);
        cudaEventCreate(&g->ev0[i]); cudaEventCreate(&g->ev1[i]);
        uz_group::Worker* w = new uz_group::Worker();
        w->th = std::thread(group_worker_loop, w);
        g->workers.push_back(w);
    }
    *out = g;
    return UZ_OK;
}

void uz_group_destroy(uz_group* g) {
    if (!g) return;
    for (auto* w : g->workers) {
        { std::lock_guard<std::mutex> lk(w->m); w->quit = true; w->cv.notify_all(); }
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    for (size_t i = 0; i < g->ev0.size(); ++i) {
        if (i < g->ctx.size()) cudaSetDevice(g->ctx[i]->device);
        if (g->ev0[i]) cudaEventDestroy(g->ev0[i]);
        if (g->ev1[i]) cudaEventDestroy(g->ev1[i]);
    }
    if (g->h_results.p) { cudaFreeHost(g->h_results.p); g->h_results.p = nullptr; }
    for (auto* c : g->ctx) uz_destroy(c);
    delete g;
}

const char* uz_group_last_error(const uz_group* g) { return g ? g->err.c_str() : g_create_err.c_str(); }
int32_t uz_group_size(const uz_group* g) { return g ? (int32_t)g->ctx.size() : 0; }
uz_context* uz_group_context(uz_group* g, int32_t rank) { return (g && rank >= 0 && rank < (int32_t)g->ctx.size()) ? g->ctx[rank] : nullptr; }

uz_status uz_group_set_params(uz_group* g, const uz_params* p) {
    if (!g || !p) return UZ_ERR_INVALID;
    if (g->in_flight) { g->err = "a batch is in flight: call uz_group_estimate_edges_end first"; return UZ_ERR_INVALID; }
    for (auto* c : g->ctx) {
        const uz_status st = uz_set_params(c, p);
        if (st != UZ_OK) { g->err = c->err; return st; }
    }
    return UZ_OK;
}

uz_status uz_group_set_gather(uz_group* g, int32_t mode) {
    if (!g || mode < 0 || mode > 1) return UZ_ERR_INVALID;
    g->gather_mode = mode;
    return UZ_OK;
}

uz_status uz_group_store_add_bulk(uz_group* g, const uz_features* cams, const int32_t* cams_per_keyframe, int32_t n_keyframes,
                                  int32_t* handles_out) {
    if (!g) return UZ_ERR_INVALID;
    if (g->in_flight) { g->err = "a batch is in flight: call uz_group_estimate_edges_end first"; return UZ_ERR_INVALID; }
    uz_status st = uz_store_add_bulk(g->ctx[0], cams, cams_per_keyframe, n_keyframes, handles_out);      // the one trip over PCIe
    if (st != UZ_OK) { g->err = g->ctx[0]->err; return st; }
    if (g->ctx.size() == 1 || n_keyframes == 0) return UZ_OK;
    const uz_context* src = g->ctx[0];
    return group_run(g, [g, src, handles_out, n_keyframes](int r) { return group_replicate(g->ctx[r], src, handles_out, n_keyframes); }, 1);
}

uz_status uz_group_store_add(uz_group* g, const uz_features* cams, int32_t n_cams, int32_t* handle_out) {
    if (!handle_out) return UZ_ERR_INVALID;
    return uz_group_store_add_bulk(g, cams, &n_cams, 1, handle_out);
}

uz_status uz_group_store_replace(uz_group* g, int32_t handle, const uz_features* cams, int32_t n_cams) {
    if (!g) return UZ_ERR_INVALID;
    if (g->in_flight) { g->err = "a batch is in flight: call uz_group_estimate_edges_end first"; return UZ_ERR_INVALID; }
    uz_status st = uz_store_replace(g->ctx[0], handle, cams, n_cams);
    if (st != UZ_OK) { g->err = g->ctx[0]->err; return st; }
    if (g->ctx.size() == 1) return UZ_OK;
    const uz_context* src = g->ctx[0];
    return group_run(g, [g, src, handle](int r) {
        uz_status s2 = sync_compute_streams(g->ctx[r]);
        return s2 != UZ_OK ? s2 : group_replicate(g->ctx[r], src, &handle, 1, true);
    }, 1);
}

uz_status uz_group_store_remove(uz_group* g, int32_t handle) {
    if (!g) return UZ_ERR_INVALID;
    if (g->in_flight) { g->err = "a batch is in flight: call uz_group_estimate_edges_end first"; return UZ_ERR_INVALID; }
    return group_run(g, [g, handle](int r) { return uz_store_remove(g->ctx[r], handle); });
}

uz_status uz_group_store_clear(uz_group* g) {
    if (!g) return UZ_ERR_INVALID;
    if (g->in_flight) { g->err = "a batch is in flight: call uz_group_estimate_edges_end first"; return UZ_ERR_INVALID; }
    return group_run(g, [g](int r) { return uz_store_clear(g->ctx[r]); });
}

int32_t uz_group_store_size(const uz_group* g) { return g ? uz_store_size(g->ctx[0]) : 0; }

// shard r of the batch on device r; records go to `sink + pair index` (any address every device can write)
static void group_estimate_dispatch(uz_group* g, const int32_t* from_handles, const int32_t* to_handles, int32_t n_pairs,
                                    uz_edge_result* sink, uz_edge_result* host_out) {
    const int world = (int)g->ctx.size();
    group_dispatch(g, [=](int r) -> uz_status {
        uz_context* ctx = g->ctx[r];
        uz_status st = check_ctx(ctx);
        if (st != UZ_OK) return st;
        int lo, hi;
        group_shard(n_pairs, world, r, &lo, &hi);
        cudaEventRecord(g->ev0[r], ctx->stream);
        bool copied_home = false;
        if (hi > lo) {
            std::vector<PairRef> pairs;
            st = pairs_from_handles(ctx, from_handles + lo, to_handles + lo, hi - lo, pairs);
            if (st != UZ_OK) return st;
            if (g->gather_mode == 0 || !host_out) {
                st = run_pairs_pipelined(ctx, pairs, sink + lo);         // the solve writes through the peer- / host-mapped pointer
                if (st != UZ_OK) return st;
            } else {                                                     // local records, copied home chunk by chunk over the device's own PCIe link
                UZ_CUDA(ctx, ctx->d_results.ensure((size_t)(hi - lo) * sizeof(uz_edge_result)));
                st = run_pairs_to_host(ctx, pairs, (uz_edge_result*)ctx->d_results.p, (uz_edge_result*)g->h_results.p + lo, host_out + lo);
                if (st != UZ_OK) return st;
                copied_home = true;
            }
        }
        cudaEventRecord(g->ev1[r], ctx->stream);
        UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, g->ev0[r], g->ev1[r]);
        g->last_ms[r] = ms;
        if (host_out && hi > lo && !copied_home)       // host-mapped sink: every worker carries its own shard from the landing zone to the caller's array
            memcpy(host_out + lo, (const uz_edge_result*)g->h_results.p + lo, (size_t)(hi - lo) * sizeof(uz_edge_result));
        return UZ_OK;
    });
}
static uz_status group_estimate(uz_group* g, const int32_t* from_handles, const int32_t* to_handles, int32_t n_pairs,
                                uz_edge_result* sink, uz_edge_result* host_out) {
    group_estimate_dispatch(g, from_handles, to_handles, n_pairs, sink, host_out);
    return group_wait(g);
}

static uz_status group_result_zone(uz_group* g, int32_t n_pairs) {
    const size_t need = (size_t)n_pairs * sizeof(uz_edge_result);
    if (need > g->h_results.cap) {
        if (g->h_results.p) cudaFreeHost(g->h_results.p);
        g->h_results.p = nullptr; g->h_results.cap = 0;
        cudaSetDevice(g->ctx[0]->device);
        void* p = nullptr;
        const size_t want = need + need / 4 + 4096;
        if (cudaHostAlloc(&p, want, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            g->err = "pinned result buffer allocation failed";
            return UZ_ERR_NOMEM;
        }
        g->h_results.p = p; g->h_results.cap = want;
        void* d = nullptr;
        cudaHostGetDevicePointer(&d, p, 0);       // unified addressing: the same address on every device
        g->h_results_dev = (uz_edge_result*)d;
    }
    return UZ_OK;
}

uz_status uz_group_estimate_edges(uz_group* g, const int32_t* from_handles, const int32_t* to_handles, int32_t n_pairs,
                                  uz_edge_result* results) {
    if (!g || n_pairs < 0 || (n_pairs > 0 && (!from_handles || !to_handles || !results))) return UZ_ERR_INVALID;
    if (g->in_flight) { g->err = "a batch is in flight: call uz_group_estimate_edges_end first"; return UZ_ERR_INVALID; }
    if (n_pairs == 0) return UZ_OK;
    const uz_status st = group_result_zone(g, n_pairs);
    if (st != UZ_OK) return st;
    return group_estimate(g, from_handles, to_handles, n_pairs, g->h_results_dev, results);
}

uz_status uz_group_estimate_edges_begin(uz_group* g, const int32_t* from_handles, const int32_t* to_handles, int32_t n_pairs,
                                        uz_edge_result* results) {
    if (!g || n_pairs < 0 || (n_pairs > 0 && (!from_handles || !to_handles || !results))) return UZ_ERR_INVALID;
    if (g->in_flight) { g->err = "a batch is in flight: call uz_group_estimate_edges_end first"; return UZ_ERR_INVALID; }
    if (n_pairs == 0) return UZ_OK;
    const uz_status st = group_result_zone(g, n_pairs);
    if (st != UZ_OK) return st;
    group_estimate_dispatch(g, from_handles, to_handles, n_pairs, g->h_results_dev, results);
    g->in_flight = true;
    return UZ_OK;
}

uz_status uz_group_estimate_edges_end(uz_group* g) {
    if (!g) return UZ_ERR_INVALID;
    if (!g->in_flight) return UZ_OK;
    g->in_flight = false;
    return group_wait(g);
}

uz_status uz_group_estimate_edges_device(uz_group* g, const int32_t* from_handles, const int32_t* to_handles, int32_t n_pairs,
                                         void* results_on_first_device) {
    if (!g || n_pairs < 0 || (n_pairs > 0 && (!from_handles || !to_handles || !results_on_first_device))) return UZ_ERR_INVALID;
    if (g->in_flight) { g->err = "a batch is in flight: call uz_group_estimate_edges_end first"; return UZ_ERR_INVALID; }
    if (n_pairs == 0) return UZ_OK;
    return group_estimate(g, from_handles, to_handles, n_pairs, (uz_edge_result*)results_on_first_device, nullptr);
}

uz_status uz_group_last_timing(const uz_group* g, double* ms_per_device, int32_t capacity) {
    if (!g || !ms_per_device) return UZ_ERR_INVALID;
    for (int i = 0; i < capacity && i < (int)g->last_ms.size(); ++i) ms_per_device[i] = g->last_ms[i];
    return UZ_OK;
}

}  // extern "C"
