// uz_capi.cu — the C-ABI of include/uzliti_edge.h: context, device-resident keyframe store, stage entry points and the
// batched path.  The translation unit is assembled from parts:
//   uz_internal.cuh     state (allocator, store records, context)
//   uz_upload.inl       host buffers -> device layouts (gather kernel, pinned staging ring), sample tables
//   uz_batch.inl        one batch from task enumeration to the last launch (run_pairs)
//   uz_capi_places.inl  candidate generation        uz_capi_ingest.inl  ingestion        uz_group.inl  several GPUs
// No CPU compute fallback exists: without a usable device every compute entry point returns UZ_ERR_CUDA.
#include "uz_internal.cuh"
#include "uz_upload.inl"
#include "uz_batch.inl"

namespace {

uz_status validate_params(uz_context* ctx, const uz_params* p) {
    if (p->ransac_iterations < 1 || p->ransac_iterations > UZ_MAX_ITERATIONS) return fail(ctx, UZ_ERR_INVALID, "ransac_iterations out of range");
    if (p->ratio_den <= 0 || p->ratio_num < 0 || p->ratio_den > 4096 || p->ratio_num > 4096) return fail(ctx, UZ_ERR_INVALID, "ratio out of range");
    if (p->cross_check != 0 && p->cross_check != 1) return fail(ctx, UZ_ERR_INVALID, "cross_check must be 0 or 1");
    if (p->min_keypoints < 0) return fail(ctx, UZ_ERR_INVALID, "min_keypoints < 0");
    return UZ_OK;
}

// registers keyframes whose cameras are already placed on the device; one block per keyframe
void register_keyframes(uz_context* ctx, const std::vector<Cam>& up, const std::vector<BlockRef>& blocks, const int32_t* cams_per_keyframe,
                        int32_t n_keyframes, int32_t* handles_out) {
    size_t k = 0;
    for (int i = 0; i < n_keyframes; ++i) {
        int32_t h;
        if (!ctx->free_handles.empty()) { h = ctx->free_handles.back(); ctx->free_handles.pop_back(); }
        else { h = (int32_t)ctx->kfs.size(); ctx->kfs.emplace_back(); }
        Keyframe& kf = ctx->kfs[h];
        kf.cams.assign(up.begin() + k, up.begin() + k + cams_per_keyframe[i]);
        for (const Cam& c : kf.cams) ctx->store_max_n = std::max(ctx->store_max_n, c.n);
        kf.live = true;
        kf.block = blocks[i].p; kf.block_bytes = blocks[i].bytes;
        k += (size_t)cams_per_keyframe[i];
        ctx->live++;
        handles_out[i] = h;
    }
}

// uz_store_replace: the keyframe's new cameras are placed; its old range goes back to the arena unless a place still reads it
void swap_keyframe(uz_context* ctx, int32_t handle, const std::vector<Cam>& up, const BlockRef& block) {
    Keyframe& kf = ctx->kfs[handle];
    const BlockRef old{kf.block, kf.block_bytes};
    auto it = ctx->places.by_handle.find(handle);
    if (it != ctx->places.by_handle.end() && ctx->places.places[it->second].ins_count > 0) ctx->places.retired[it->second].push_back(old);
    else ctx->store_arena.free(old.p, old.bytes);
    kf.cams = up;
    kf.block = block.p; kf.block_bytes = block.bytes;
    for (const Cam& c : kf.cams) ctx->store_max_n = std::max(ctx->store_max_n, c.n);
}

uz_status sync_compute_streams(uz_context* ctx) {
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->alt) UZ_CUDA(ctx, cudaStreamSynchronize(ctx->alt));
    if (ctx->solve_stream) UZ_CUDA(ctx, cudaStreamSynchronize(ctx->solve_stream));
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
        const char* hs = getenv("UZ_HOST_SLOTS");
        if (hs && atoi(hs) >= 2) ctx->host_slots = std::min(atoi(hs), (int)uz_context::kSlots);
        const char* sp = getenv("UZ_SOLVE_SMEM_PAD");
        if (sp) ctx->solve_smem_pad = std::max(0, atoi(sp));
        const char* pc = getenv("UZ_PIPELINE_CALLS");
        if (pc) ctx->pipeline_calls = atoi(pc);
        const char* hf = getenv("UZ_HOST_FIRST_WAVES");
        if (hf) ctx->host_first_waves = atoi(hf);
        const char* hc = getenv("UZ_HOST_CHUNKS");
        if (hc) ctx->host_chunks = atoi(hc);
        const char* mm = getenv("UZ_MATCH_MMA");
        if (mm) ctx->match_mma = atoi(mm);
        ctx->narrow_e4 = ctx->match_mma == 1;
        const char* mw = getenv("UZ_MATCH_MMA_WIDE");
        if (mw) ctx->match_mma_wide = atoi(mw);
        ctx->wide_e4 = ctx->match_mma_wide == 1;

        const char* stt = getenv("UZ_STAGE_THREADS");
        if (stt && atoi(stt) >= 1 && atoi(stt) <= 64) ctx->stage_threads = atoi(stt);
        const char* rm = getenv("UZ_RING_MB");
        if (rm && atoi(rm) >= 1 && atoi(rm) <= 4096) ctx->ring_half = (size_t)atoi(rm) << 20;
    }
    e = cudaFuncSetAttribute(solve_kernel<kSolveThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)solve_smem_bytes(UZ_MAX_FEATURES));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(solve_kernel<kSolveThreadsWide>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)solve_smem_bytes(UZ_MAX_FEATURES, kSolveThreadsWide));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(solve_stream_kernel<kSolveThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)solve_smem_bytes(1024));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(knn2_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmaSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(knn2_mmak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmakSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(knn2_mmaf_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F4<false>::kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(knn2_mmaf_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F4<true>::kSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(knn2_mmaw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmawSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(knn2_mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMma2SmemBytes);
    if (e == cudaSuccess) e = set_carveouts();
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->f4_zeros, kF4ZeroPageBytes);
    if (e == cudaSuccess) e = cudaMemset(ctx->f4_zeros, 0, kF4ZeroPageBytes);
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
    for (auto& t : ctx->samples) t.d.release();
    if (ctx->pool) { ctx->pool->stop(); delete ctx->pool; ctx->pool = nullptr; }
    ctx->ring.release();
    for (int i = 0; i < 2; ++i) if (ctx->ring_free[i]) cudaEventDestroy(ctx->ring_free[i]);
    ctx->d_results.release(); ctx->d_dbg_matches.release(); ctx->d_dbg_mask.release(); ctx->d_dbg_counts.release(); ctx->d_dbg_phase.release(); ctx->d_chunks.release(); ctx->h_chunks.release(); ctx->h_results.release();
    ctx->d_misc.release();
    if (ctx->f4_zeros) cudaFree(ctx->f4_zeros);
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
    std::vector<BlockRef> blocks;
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->h_chunks.reset(); ctx->d_chunks.reset();
    // one range of the arena per keyframe: uz_store_remove gives it back
    if ((st = place_cams(ctx, ctx->store_arena, feats, cams_per_keyframe, (size_t)n_keyframes, up, blocks)) != UZ_OK) return st;
    st = fill_cams(ctx, feats, up);
    if (st == UZ_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)      // host buffers are borrowed only for the call
        st = fail(ctx, UZ_ERR_CUDA, "keyframe upload failed");
    if (st != UZ_OK) {
        cudaStreamSynchronize(ctx->stream); cudaGetLastError();
        for (auto& b : blocks) ctx->store_arena.free(b.p, b.bytes);
        return st;
    }
    register_keyframes(ctx, up, blocks, cams_per_keyframe, n_keyframes, handles_out);
    return UZ_OK;
}

uz_status uz_store_add(uz_context* ctx, const uz_features* cams, int32_t n_cams, int32_t* handle_out) {
    if (!handle_out) return UZ_ERR_INVALID;
    return uz_store_add_bulk(ctx, cams, &n_cams, 1, handle_out);
}

uz_status uz_store_replace(uz_context* ctx, int32_t handle, const uz_features* cams, int32_t n_cams) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (handle < 0 || handle >= (int32_t)ctx->kfs.size() || !ctx->kfs[handle].live) return fail(ctx, UZ_ERR_INVALID, "unknown keyframe handle");
    if (n_cams < 0 || (n_cams > 0 && !cams)) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    std::vector<const uz_features*> feats((size_t)n_cams);
    for (int i = 0; i < n_cams; ++i) feats[i] = cams + i;
    if ((st = sync_compute_streams(ctx)) != UZ_OK) return st;
    ctx->h_chunks.reset(); ctx->d_chunks.reset();
    std::vector<Cam> up;
    std::vector<BlockRef> blocks;
    if ((st = place_cams(ctx, ctx->store_arena, feats, &n_cams, 1, up, blocks)) != UZ_OK) return st;
    st = fill_cams(ctx, feats, up);
    if (st == UZ_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, UZ_ERR_CUDA, "keyframe upload failed");
    if (st != UZ_OK) {
        cudaStreamSynchronize(ctx->stream); cudaGetLastError();
        for (auto& b : blocks) ctx->store_arena.free(b.p, b.bytes);
        return st;
    }
    swap_keyframe(ctx, handle, up, blocks[0]);
    return UZ_OK;
}

uz_status uz_store_remove(uz_context* ctx, int32_t handle) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (handle < 0 || handle >= (int32_t)ctx->kfs.size() || !ctx->kfs[handle].live) return fail(ctx, UZ_ERR_INVALID, "unknown keyframe handle");
    // launches that still read the keyframe must be through before its range can be handed out again
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->alt) UZ_CUDA(ctx, cudaStreamSynchronize(ctx->alt));
    if (ctx->solve_stream) UZ_CUDA(ctx, cudaStreamSynchronize(ctx->solve_stream));
    places_forget_handle(ctx, handle);           // the handle is recycled: no place, no checked_ pair may keep naming it
    Keyframe& kf = ctx->kfs[handle];
    ctx->store_arena.free(kf.block, kf.block_bytes);
    kf.block = nullptr; kf.block_bytes = 0;
    kf.live = false;
    kf.cams.clear();
    ctx->free_handles.push_back(handle);
    ctx->live--;
    return UZ_OK;
}

uz_status uz_store_clear(uz_context* ctx) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->alt) UZ_CUDA(ctx, cudaStreamSynchronize(ctx->alt));
    if (ctx->solve_stream) UZ_CUDA(ctx, cudaStreamSynchronize(ctx->solve_stream));
    ctx->kfs.clear(); ctx->free_handles.clear(); ctx->live = 0; ctx->store_max_n = 0;
    ctx->store_arena.reset();
    if ((st = places_reset(ctx)) != UZ_OK) return st;      // the recogniser's nodes point into the store
    return UZ_OK;
}

int32_t uz_store_size(const uz_context* ctx) { return ctx ? ctx->live : 0; }

int64_t uz_store_bytes(const uz_context* ctx) {
    if (!ctx) return 0;
    return (int64_t)ctx->store_arena.used;
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
    ctx->h_chunks.reset(); ctx->d_chunks.reset();
    // two position-less cameras
    Keyframe kq, kt;
    kq.cams.resize(1); kt.cams.resize(1);
    auto up = [&](const uint8_t* h, int n, int stride, Cam& c) -> uz_status {
        c.n = n; c.feature_type = UZ_FEATURE_ORB; c.dbytes = db;
        if (n == 0) return UZ_OK;
        const int halves = n * (db / 32);
        c.raw = (uint32_t*)ctx->transient.alloc((size_t)n * db);
        c.csa = (uint32_t*)ctx->transient.alloc((size_t)n * db);
        c.e8 = (uint8_t*)ctx->transient.alloc(operand_bytes(n, db, ctx->operand_fmt()));
        uint8_t* stage = (uint8_t*)ctx->transient.alloc((size_t)n * stride);
        if (!c.raw || !c.csa || !stage || !c.e8) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
        UZ_CUDA(ctx, cudaMemcpyAsync(stage, h, (size_t)(n - 1) * stride + db, cudaMemcpyHostToDevice, ctx->stream));
        pack_descriptors_kernel<<<(halves + 255) / 256, 256, 0, ctx->stream>>>(stage, halves, stride, c.raw, c.csa, db / 32);
        ctx->launches++;
        return derive_layouts(ctx, &c, 1);        // (recomputes the CSA form too; one launch)
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
        const uint16_t* table = nullptr;
        st = ensure_samples(ctx, iterations, do_prosac, M, &table);
        if (st != UZ_OK) return st;
        sp.samples = table; sp.samples_by_m = 1;
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
    const uint16_t* table = nullptr;
    st = ensure_samples(ctx, iterations, do_prosac, max_m, &table);
    if (st != UZ_OK) return st;
    SolveParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.samples = table; sp.samples_by_m = 1;
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
    if (ctx->debug || !ctx->match_mma || !ctx->pipeline_calls || (size_t)n_pairs < (size_t)8 * 5 * (size_t)ctx->sm_count) {
        // small batches (the online case): one launch pair, records straight into the caller's array
        st = uz_estimate_edges_device(ctx, from_handles, to_handles, n_pairs, ctx->d_results.p);
        if (st != UZ_OK) return st;
        UZ_CUDA(ctx, cudaMemcpyAsync(results, ctx->d_results.p, (size_t)n_pairs * sizeof(uz_edge_result), cudaMemcpyDeviceToHost, ctx->stream));
        UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return UZ_OK;
    }
    UZ_CUDA(ctx, ctx->h_results.ensure((size_t)n_pairs * sizeof(uz_edge_result)));
    std::vector<PairRef> pairs;
    st = pairs_from_handles(ctx, from_handles, to_handles, n_pairs, pairs);
    if (st != UZ_OK) return st;
    return run_pairs_to_host(ctx, pairs, (uz_edge_result*)ctx->d_results.p, (uz_edge_result*)ctx->h_results.p, results);
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
    // Chunks of about 1600 pairs, between 4 and 40 of them.  Consecutive chunks
    // compute on two alternating streams (below), so a chunk boundary costs little; small chunks shrink the exposed first
    // upload and send records home earlier, large ones amortise the launches of a chunk.  Measured on C4 (25 000 pairs) with
    // the tensor-core match kernel (store-resident 3.28 M edges/s): 8 chunks 2.49 M, 16: 2.52 M, 32: 2.37 M, 64: 2.12 M;
    // chunks that grow by 15-60 % each (short first upload, fewer launches later) measured the same as equal ones
    // (with the integer-pipe kernels of round 1, 768-pair chunks were best: 926-930 k against 937 k store-resident).
    // UZ_HOST_CHUNKS = k > 0 forces k equal parts.
    // A chunk is a whole number of solve waves (5 CTAs per SM; that is also a whole number of tensor-core items per match
    // CTA): the next chunk's match CTAs need SMs that are completely free, so a last wave that fills a quarter of the
    // chip costs a tenth of a millisecond per chunk.
    int want_chunks = 1;
    const int wave = 5 * ctx->sm_count;
    int per_chunk = std::max(wave, ((1600 + wave / 2) / wave) * wave);          // 1480 pairs on a B200
    if (ctx->debug) want_chunks = 1;                            // the parity taps describe ONE launch pair
    else if (ctx->host_chunks > 0) { want_chunks = std::min(ctx->host_chunks, n_pairs); per_chunk = (n_pairs + want_chunks - 1) / want_chunks; }
    else if (n_pairs < 2048) want_chunks = 1;
    else {
        want_chunks = (n_pairs + per_chunk - 1) / per_chunk;
        if (want_chunks > 40) { per_chunk = ((n_pairs / 40 + wave - 1) / wave) * wave; want_chunks = (n_pairs + per_chunk - 1) / per_chunk; }
        if (want_chunks < 4) { want_chunks = 4; per_chunk = (n_pairs + 3) / 4; }
    }
    if (want_chunks <= 1) per_chunk = n_pairs;
    // the first chunk is ONE solve wave (UZ_HOST_FIRST_WAVES): its upload is the only one nothing overlaps
    const int first = (ctx->host_first_waves > 0 && !ctx->debug && ctx->host_chunks <= 0 && want_chunks >= 4 &&
                       ctx->host_first_waves * wave < per_chunk) ? ctx->host_first_waves * wave : 0;
    std::vector<size_t> chunk_pair_end;
    if (first > 0) chunk_pair_end.push_back((size_t)first);
    for (size_t at = (size_t)first; at < (size_t)n_pairs;) {
        at = std::min<size_t>((size_t)n_pairs, at + (size_t)per_chunk);
        if ((size_t)n_pairs - at < (size_t)wave / 2) at = (size_t)n_pairs;          // no crumb at the end
        chunk_pair_end.push_back(at);
    }
    const int n_chunks = (int)chunk_pair_end.size();

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
    // UZ_TRACE=2: device timeline of the chunks (upload begin/end on the side stream, compute begin/end)
    const bool tl_on = getenv("UZ_TRACE") != nullptr && atoi(getenv("UZ_TRACE")) >= 2;
    std::vector<cudaEvent_t> tl(tl_on ? (size_t)n_chunks * 5 : 0, nullptr);
    // the host runs several chunks ahead of the compute streams (a two-deep ring made every upload wait for the compute
    // of the chunk three before it: 3 chunks per 1.6 ms on C4, measured with UZ_TRACE=2)
    struct DepthGuard { uz_context* c; ~DepthGuard() { c->slot_depth = 2; } } depth_guard{ctx};
    if (n_chunks > 2) ctx->slot_depth = ctx->host_slots;
    for (int c = 0; c < n_chunks && st == UZ_OK; ++c) {
        const size_t p0 = chunk_begin(c), p1 = chunk_pair_end[c];
        if (p1 <= p0) continue;
        // (1) intern the chunk's cameras
        g_stage.start();
        const size_t u0 = uniq.size(), cf0 = cf, ct0 = ct;
        from_u.clear(); to_u.clear();
        for (size_t i = p0; i < p1; ++i) {
            for (int k = 0; k < n_from[i]; ++k, ++cf) from_u.push_back(intern(from_cams + cf));
            for (int k = 0; k < n_to[i]; ++k, ++ct) to_u.push_back(intern(to_cams + ct));
        }
        g_stage.stop(0);
        // (2) upload (+ layout pass) of the cameras nobody before this chunk needed, on the side stream
        cudaEvent_t ready = nullptr;
        if (uniq.size() > u0) {
            std::vector<const uz_features*> fresh(uniq.begin() + u0, uniq.end());
            std::vector<Cam> got;
            if (piped) ctx->stream = ctx->side;
            ctx->copy_beside_compute = piped && c > 0;       // chunk 0 has the chip to itself
            if (tl_on) { tl[5 * c] = ctx->get_event(); cudaEventRecord(tl[5 * c], ctx->stream); ctx->trace_mid = tl[5 * c + 4] = ctx->get_event(); }
            st = upload_cams(ctx, ctx->transient, fresh, got);
            if (tl_on) { tl[5 * c + 1] = ctx->get_event(); cudaEventRecord(tl[5 * c + 1], ctx->stream); ctx->trace_mid = nullptr; }
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
        g_stage.start();
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
        g_stage.stop(3);
        if (alternate && (c & 1)) ctx->stream = ctx->alt;
        // the upload stream is in order, so the newest upload event covers every camera uploaded so far; a chunk on the
        // other compute stream needs it even when it brought no camera of its own
        if (ready) {
            if (last_ready) ctx->event_pool.push_back(last_ready);
            last_ready = ready;
        }
        if (last_ready && (ready || alternate)) cudaStreamWaitEvent(ctx->stream, last_ready, 0);
        cudaStream_t rs = ctx->stream;
        if (tl_on) { tl[5 * c + 2] = ctx->get_event(); cudaEventRecord(tl[5 * c + 2], ctx->stream); }
        st = run_pairs(ctx, part, (uz_edge_result*)ctx->d_results.p + p0, /*join=*/false, &rs);
        if (tl_on) { tl[5 * c + 3] = ctx->get_event(); cudaEventRecord(tl[5 * c + 3], rs); }
        ctx->stream = main_stream;
        // into pinned memory: a pageable destination would make the copy synchronous and stall the next chunk's enqueue
        if (st == UZ_OK && cudaMemcpyAsync(h_res + p0, (uz_edge_result*)ctx->d_results.p + p0, (p1 - p0) * sizeof(uz_edge_result),
                                           cudaMemcpyDeviceToHost, rs) != cudaSuccess)
            st = fail(ctx, UZ_ERR_CUDA, "cudaMemcpyAsync(results) failed");
        g_stage.start();
        if (st == UZ_OK) { home[c] = ctx->get_event(); cudaEventRecord(home[c], rs); }
        if (c == 0) tr.lap("first chunk enqueued");
        drain(c, false);
        g_stage.stop(9);
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
    if (tl_on) {
        cudaEvent_t base = nullptr;
        for (auto e : tl) if (e) { base = e; break; }
        for (int c = 0; c < n_chunks && base; ++c) {
            float t[5] = {-1, -1, -1, -1, -1};
            for (int k = 0; k < 5; ++k)
                if (tl[5 * c + k] && cudaEventElapsedTime(&t[k], base, tl[5 * c + k]) != cudaSuccess) { cudaGetLastError(); t[k] = -1; }
            fprintf(stderr, "[uz timeline] chunk %2d: upload %7.3f .. (gathered %7.3f) .. %7.3f ms   compute %7.3f .. %7.3f ms\n", c, t[0], t[4], t[1], t[2], t[3]);
        }
        for (auto e : tl) if (e) ctx->event_pool.push_back(e);
    }
    {
        static const char* const names[10] = {"intern cameras", "place cameras", "copy lists + gather + derive", "pair views", "enumerate tasks",
                                              "choose shapes", "keys + tiles", "table copies", "launches", "records home + drain"};
        g_stage.report(names, 10);
    }
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
#include "uz_group.inl"
