// uz_batch.inl — one batch of keyframe pairs from task enumeration to the last launch (part of the uz_capi.cu
// translation unit).  What the reference does per pair on one worker thread
// (/root/reference/transformation_estimation/src/transformation_estimator.cpp:45-62: LIFO pop, impl, callback, 1 ms
// sleep) becomes:  enumerate the comparable camera pairs of all pairs (feature_transformation_estimator.cpp:40-49)
// -> pick kernel forms and tile shapes -> lay out the key scratch and the tile lists -> tables to the device in one
// copy -> ONE match launch per descriptor width -> the solve (one CTA per pair, or the persistent streaming grid beside
// the integer-pipe match kernel).
namespace {

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
        solve_kernel<kSolveThreads><<<n_ctas, kSolveThreads, solve_smem_bytes(cap) + (size_t)ctx->solve_smem_pad, st>>>(d_tasks, d_pair_tasks, d_keys, sp, d_results);
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
    if (e == cudaSuccess) e = max_shared_carveout(knn2_mma_kernel);
    if (e == cudaSuccess) e = max_shared_carveout(knn2_mmak_kernel);
    if (e == cudaSuccess) e = max_shared_carveout(knn2_mmaf_kernel<false>);
    if (e == cudaSuccess) e = max_shared_carveout(knn2_mmaf_kernel<true>);
    if (e == cudaSuccess) e = max_shared_carveout(knn2_mmaw_kernel);
    if (e == cudaSuccess) e = max_shared_carveout(knn2_mma2_kernel);
    if (e == cudaSuccess) e = max_shared_carveout(derive_layouts_kernel);
    return e;
}

// relative throughput of a shape at full occupancy (measured on C4-sized work; the one-warp CTA is capped at 32 warps per SM)
struct KnnConfig { int threads, qpt; double speed; };
const KnnConfig kKnnConfigs[6] = {{256, 4, 0.92}, {128, 4, 0.92}, {64, 2, 0.985}, {256, 2, 1.0}, {128, 2, 1.0}, {32, 2, 0.95}};

// Everything the host decides about one batch before anything is launched.
struct BatchPlan {
    uz_params prm;
    int n_pairs = 0;
    bool with_solve = false, cross = false;
    MatchTask* tasks = nullptr;          // pinned staging of the slot
    int2* pair_tasks = nullptr;
    size_t n_tasks = 0, n_fwd = 0, n_wide = 0;
    int max_nq = 0, cap = 0;
    int64_t compares = 0;
    bool mma = false;                    // 256-bit matchings on the tensor cores (knn2_mmaf_kernel; int8 alternatives)
    bool mma_wide = false;               // 512-bit matchings on the tensor cores (knn2_mmaf_kernel<true>; knn2_mmaw_kernel)
    bool mma2 = false;                   // ... on CTA pairs (knn2_mma2_kernel): launches that keep every pair of SMs busy
    int best_cfg = 3, wide_cfg = 0;      // integer-pipe tile shapes
    int tile_rows = 512, wide_threads = 256, wide_tile_rows = 512;
    bool fused = false;                  // cross-check inside the match kernel (integer-pipe kernels only)
    int seg_target = 1;
    bool seg_narrow = false, seg_wide = false;
    size_t key_rows = 0, col_begin = 0;
    size_t n_tiles_narrow = 0, n_tiles_wide = 0;
    size_t narrow_tile_bytes = 8, wide_tile_bytes = 8;
    bool may_stream = false, streaming = false;
    // device views of the tables
    const MatchTask* d_tasks = nullptr;
    const uint8_t* d_tiles_narrow = nullptr;
    const uint8_t* d_tiles_wide = nullptr;
    const int4* d_merges = nullptr;
    const int2* d_pair_tasks = nullptr;
    size_t tiles_bytes() const { return n_tiles_narrow * narrow_tile_bytes + n_tiles_wide * wide_tile_bytes; }
    int rows_of(size_t t, const std::vector<uint8_t>& wide) const { return wide[t] ? wide_tile_rows : tile_rows; }
};

// (1) comparable camera pairs of every pair (:40-49): one MatchTask each; with the reversed-matching form of the
// cross-check every matching is listed a second time with the roles swapped, behind the forward ones
uz_status enumerate_tasks(uz_context* ctx, const std::vector<PairRef>& pairs, uz_context::Slot& sl, BatchPlan& bp) {
    const uz_params& prm = bp.prm;
    size_t max_tasks = 0;
    for (const PairRef& p : pairs) max_tasks += (size_t)p.n_from * (size_t)p.n_to;
    UZ_CUDA(ctx, sl.h_tasks.ensure(std::max<size_t>(max_tasks, 1) * (bp.cross ? 2 : 1) * sizeof(MatchTask)));
    UZ_CUDA(ctx, sl.h_pair_tasks.ensure((size_t)bp.n_pairs * sizeof(int2)));
    bp.tasks = (MatchTask*)sl.h_tasks.p;
    bp.pair_tasks = (int2*)sl.h_pair_tasks.p;
    std::vector<uint8_t>& task_wide = ctx->task_wide;      // host-side: 1 = 64-byte rows (knn2_wide_kernel)
    task_wide.assign(std::max<size_t>(max_tasks, 1) * (bp.cross ? 2 : 1), 0);
    size_t n_tasks = 0, key_rows = 0;
    // the tensor-core form needs the E8 layout of both cameras; the measured alternatives (UZ_KNN_VARIANT) stay on the integer pipes
    bp.mma = ctx->match_mma && ctx->variant_csa && ctx->variant_pack16 && ctx->force_cfg < 0;
    bp.mma_wide = bp.mma && ctx->match_mma_wide != 0;
    auto lacks_e8 = [](const Cam& c, bool wide) { return (c.dbytes != UZ_DESC_BYTES) == wide && c.n > 0 && !c.e8; };
    for (int i = 0; i < bp.n_pairs && (bp.mma || bp.mma_wide); ++i) {
        for (int a = 0; a < pairs[i].n_from; ++a) { if (lacks_e8(pairs[i].from[a], false)) bp.mma = false; if (lacks_e8(pairs[i].from[a], true)) bp.mma_wide = false; }
        for (int b = 0; b < pairs[i].n_to; ++b) { if (lacks_e8(pairs[i].to[b], false)) bp.mma = false; if (lacks_e8(pairs[i].to[b], true)) bp.mma_wide = false; }
    }
    bp.mma_wide = bp.mma_wide && bp.mma;
    for (int i = 0; i < bp.n_pairs; ++i) {
        const int first = (int)n_tasks;
        const Cam* fc = pairs[i].from;
        const Cam* tc = pairs[i].to;
        for (int a = 0; a < pairs[i].n_from; ++a)
            for (int b = 0; b < pairs[i].n_to; ++b) {
                const Cam& F = fc[a];
                const Cam& T = tc[b];
                if (F.n >= prm.min_keypoints && T.n >= prm.min_keypoints && F.feature_type == T.feature_type &&
                    F.sensor_frame == T.sensor_frame && F.dbytes == T.dbytes) {
                    const bool wide = F.dbytes != UZ_DESC_BYTES;
                    if (wide) { task_wide[n_tasks] = 1; ++bp.n_wide; }
                    MatchTask& tk = bp.tasks[n_tasks++];
                    const bool bin = is_binary_type(F.feature_type);   // unknown type: empty matches (:60-62)
                    if (wide ? bp.mma_wide : bp.mma) {
                        tk.q_desc = (const uint32_t*)T.e8; tk.t_desc = (const uint32_t*)F.e8;
                    } else {
                        const bool use_csa = ctx->variant_csa || wide;     // the wide kernel has the CSA form only
                        tk.q_desc = use_csa ? T.csa : T.raw;
                        tk.t_desc = use_csa ? F.csa : F.raw;
                    }
                    tk.nq = bin ? T.n : 0; tk.nt = F.n;
                    tk.key_off = (uint32_t)key_rows; tk.pair = i;
                    tk.q_pos = T.pos; tk.q_valid = T.valid; tk.t_pos = F.pos; tk.t_valid = F.valid;
                    tk.cam_from = (int)a; tk.cam_to = (int)b;
                    tk.rev_key_off = kNoRev; tk.pad_ = 0;
                    key_rows += (size_t)tk.nq;
                    bp.max_nq = std::max(bp.max_nq, tk.nq);
                    bp.compares += (int64_t)tk.nq * tk.nt;
                }
            }
        bp.pair_tasks[i] = make_int2(first, (int)n_tasks - first);
    }
    bp.n_fwd = n_tasks;
    // Cross-check.  Fused form (integer-pipe kernels, default there): the forward match kernel also keeps, per train row,
    // the minimum over the queries (uz_knn2.cuh, col_update16); the matching's column keys live behind the row keys in
    // the same scratch.  Reversed form (tensor-core kernel, UZ_XCHECK_FUSED=0, and the kernel variants without the
    // packed-key shapes): every matching runs a second time with query and train swapped.
    const bool narrow_on_mma = (bp.mma && bp.n_wide < bp.n_fwd) || (bp.mma_wide && bp.n_wide > 0);       // anything on the tensor cores
    bp.fused = bp.cross && ctx->xcheck_fused && ctx->variant_csa && ctx->variant_pack16 && !narrow_on_mma &&
               !(ctx->force_cfg >= 0 && ctx->force_cfg < 2);
    if (!bp.fused && bp.cross) {
        for (size_t t = 0; t < bp.n_fwd; ++t) {
            MatchTask& f = bp.tasks[t];
            if (f.nq == 0) continue;                 // non-binary type: no matches to check
            MatchTask& r = bp.tasks[n_tasks];
            r = f;
            r.q_desc = f.t_desc; r.t_desc = f.q_desc; r.nq = f.nt; r.nt = f.nq;
            r.rev_key_off = kNoRev;
            task_wide[n_tasks] = task_wide[t];
            if (task_wide[t]) ++bp.n_wide;
            f.rev_key_off = 0;                       // "has a reversed task"; the offset is assigned below
            bp.compares += (int64_t)r.nq * r.nt;
            bp.max_nq = std::max(bp.max_nq, r.nq);
            f.pad_ = (uint32_t)n_tasks;              // index of the reversed task (host-side only)
            ++n_tasks;
        }
    }
    bp.n_tasks = n_tasks;
    return UZ_OK;
}

// (2) kernel forms and tile shapes
void choose_shapes(uz_context* ctx, BatchPlan& bp) {
    const std::vector<uint8_t>& task_wide = ctx->task_wide;
    const MatchTask* tasks = bp.tasks;
    // tile shape per descriptor width.  256-bit rows on the integer pipes: two queries per thread (40-56 registers: 6
    // resident CTAs per SM, measured 8 % faster than the four-query shapes, which stay reachable through UZ_KNN_CFG),
    // largest tile first.  512-bit rows: the 256 x 2 and 64 x 2 shapes of knn2_wide_kernel.
    auto pick = [&](const int* cand, int n_cand, const int* threads, const int* qpt, const double* speed, int warps_per_sm,
                    bool wide) {
        int best = cand[0];
        double best_cost = 1e300;
        for (int ci = 0; ci < n_cand; ++ci) {
            const int c = cand[ci];
            const int tile = threads[c] * qpt[c];
            double padded = 0; size_t tiles = 0;
            for (size_t t = 0; t < bp.n_tasks; ++t) {
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
    {
        static const int kCandidates[4] = {3, 4, 2, 5};
        int th[6], qp[6]; double sp[6];
        for (int c = 0; c < 6; ++c) { th[c] = kKnnConfigs[c].threads; qp[c] = kKnnConfigs[c].qpt; sp[c] = kKnnConfigs[c].speed; }
        if (bp.n_wide < bp.n_tasks && !bp.mma) bp.best_cfg = pick(kCandidates, 4, th, qp, sp, 48, false);
        if (ctx->force_cfg >= 0 && ctx->force_cfg < 6) bp.best_cfg = ctx->force_cfg;
        static const int kWideCand[2] = {0, 1};
        static const int wth[2] = {256, 64}, wqp[2] = {2, 2};
        static const double wsp[2] = {1.0, 0.97};
        if (bp.n_wide) bp.wide_cfg = pick(kWideCand, 2, wth, wqp, wsp, 32, true);
        if (ctx->force_wide_cfg >= 0 && ctx->force_wide_cfg < 2) bp.wide_cfg = ctx->force_wide_cfg;
    }
    if (bp.mma && (ctx->match_mma == 2 || ctx->match_mma == 3)) {
        // CTA pairs take 512-row items; worth it only when there are at least as many items as pairs of SMs
        size_t items2 = 0;
        for (size_t t = 0; t < bp.n_tasks; ++t)
            if (!task_wide[t]) items2 += ((size_t)tasks[t].nq + kMma2ItemRows - 1) / kMma2ItemRows;
        bp.mma2 = items2 >= (size_t)(ctx->sm_count / 2) || (ctx->match_mma == 3 && items2 > 0);     // 3: always (tests)
    }
    bp.tile_rows = bp.mma2 ? kMma2ItemRows : bp.mma ? kMmaItemRows : kKnnConfigs[bp.best_cfg].threads * kKnnConfigs[bp.best_cfg].qpt;
    bp.wide_threads = bp.wide_cfg == 0 ? 256 : 64;
    bp.wide_tile_rows = bp.mma_wide ? kMmaItemRows : 2 * bp.wide_threads;

    // the persistent streaming solve runs beside the INTEGER-PIPE match kernels only: a tensor-core match CTA owns its SM's
    // shared memory (200 KB of operand stages), no solve CTA fits next to it
    bp.cap = bp.with_solve ? std::max(128, pow2ceil(std::max(bp.max_nq, 1))) : 0;
    const int stream_ctas = ctx->stream_solve_ctas * ctx->sm_count;
    const bool narrow_on_mma = (bp.mma && bp.n_wide < bp.n_tasks) || (bp.mma_wide && bp.n_wide > 0);     // anything on the tensor cores
    bp.may_stream = bp.with_solve && !narrow_on_mma && ctx->solve_stream != nullptr && ctx->stream_solve_ctas > 0 && bp.cap <= 1024 &&
                    bp.n_pairs >= (ctx->stream_min_pairs > 0 ? ctx->stream_min_pairs : 2 * stream_ctas);

    // Small launches of the integer-pipe kernels are cut along the train rows as well (SEG kernels): a handful of pairs -
    // the online case, one new keyframe against its candidates - would otherwise run on a handful of CTAs that each walk
    // every train row.  Every (query tile, train segment) CTA writes partial neighbours, merge_segments_kernel folds
    // them.  Not for launches that fill the chip anyway, and not beside the streaming solve (it consumes keys tile by
    // tile).  The tensor-core kernel needs none of this: one 256-row item against 1000 train rows takes 4 us.
    bp.seg_target = 1;
    if (ctx->segment_small && !bp.may_stream && ctx->variant_csa && ctx->variant_pack16 &&
        (bp.mma || kKnnConfigs[bp.best_cfg].qpt == 2)) {
        double warps = 0;
        for (size_t t = 0; t < bp.n_tasks; ++t) {
            if (task_wide[t] ? bp.mma_wide : bp.mma) continue;
            const int rows = bp.rows_of(t, task_wide);
            warps += (double)(((size_t)tasks[t].nq + rows - 1) / rows) * (task_wide[t] ? bp.wide_threads : kKnnConfigs[bp.best_cfg].threads) / 32.0;
        }
        const double capacity = (double)ctx->sm_count * 32.0;
        if (warps > 0 && warps * 2 <= capacity) bp.seg_target = (int)std::min(16.0, std::floor(capacity / warps));
    }
    bp.seg_wide = bp.seg_target > 1 && !bp.mma_wide;
    bp.seg_narrow = bp.seg_target > 1 && !bp.mma;
    bp.narrow_tile_bytes = bp.seg_narrow ? sizeof(int4) : sizeof(int2);
    bp.wide_tile_bytes = bp.seg_wide ? sizeof(int4) : sizeof(int2);
}

// Runs K1 (+ optionally K2..K5) for a list of pairs whose cameras are already on the device.
// join = false (chunked callers): the caller's stream is NOT made to wait for a streaming solve, so that the next chunk's
// match kernel starts while this chunk's last pairs are still being solved; *result_stream is then the stream behind
// which the records are complete.
uz_status run_pairs(uz_context* ctx, const std::vector<PairRef>& pairs, uz_edge_result* d_results, bool join = true,
                    cudaStream_t* result_stream = nullptr) {
    BatchPlan bp;
    bp.prm = ctx->params;     // snapshot (setConfig may race with a batch in the reference)
    const uz_params& prm = bp.prm;
    bp.n_pairs = (int)pairs.size();
    const int n_pairs = bp.n_pairs;
    if (result_stream) *result_stream = ctx->stream;
    if (n_pairs == 0) return UZ_OK;
    cudaStream_t results_on = ctx->stream;
    ctx->cur_slot = (ctx->cur_slot + 1) % std::max(2, std::min(ctx->slot_depth, (int)uz_context::kSlots));
    uz_context::Slot& sl = ctx->slots[ctx->cur_slot];
    if (!sl.done) UZ_CUDA(ctx, cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    if (sl.used) UZ_CUDA(ctx, cudaEventSynchronize(sl.done));      // normally long complete

    bp.with_solve = d_results != nullptr;
    bp.cross = prm.cross_check != 0 && bp.with_solve;
    g_stage.start();
    uz_status st = enumerate_tasks(ctx, pairs, sl, bp);
    if (st != UZ_OK) return st;
    g_stage.stop(4);
    choose_shapes(ctx, bp);
    g_stage.stop(5);
    MatchTask* tasks = bp.tasks;
    const std::vector<uint8_t>& task_wide = ctx->task_wide;
    const size_t n_tasks = bp.n_tasks, n_fwd = bp.n_fwd;
    const bool fused = bp.fused, cross = bp.cross;
    auto seg_of = [&](size_t t) { return task_wide[t] ? bp.seg_wide : bp.seg_narrow; };
    auto qtiles_of = [&](size_t t) { const int r = bp.rows_of(t, task_wide); return ((size_t)tasks[t].nq + r - 1) / r; };
    // segment of a task: a multiple of 64 train rows (the key blocks of both kernels), at least 64
    auto seg_rows_of = [&](size_t t) {
        if (!seg_of(t)) return std::max(tasks[t].nt, 1);
        const int want = (tasks[t].nt + bp.seg_target - 1) / bp.seg_target;
        return std::max(64, (want + 63) & ~63);
    };
    auto nseg_of = [&](size_t t) { return seg_of(t) ? std::max(1, (tasks[t].nt + seg_rows_of(t) - 1) / seg_rows_of(t)) : 1; };

    // (3) key scratch: per task nq rows per segment (segment 0 holds the final neighbours), then the column keys of the
    // fused cross-check; tile lists: 256-bit tiles first, 512-bit tiles behind them (one launch each); inside a list:
    // forward tiles, then the reversed tiles of the same matching
    size_t key_rows = 0;
    for (size_t t = 0; t < n_tasks; ++t) {
        tasks[t].key_off = (uint32_t)std::min<size_t>(key_rows, 0xFFFFFFFFu);
        key_rows += (size_t)tasks[t].nq * (size_t)nseg_of(t);
    }
    bp.col_begin = key_rows;
    for (size_t t = 0; t < n_fwd; ++t) {
        if (tasks[t].nq == 0) continue;
        if (fused) {
            tasks[t].rev_key_off = (uint32_t)std::min<size_t>(key_rows, 0xFFFFFFFFu);
            key_rows += (size_t)tasks[t].nt;
        } else if (cross) {
            tasks[t].rev_key_off = tasks[tasks[t].pad_].key_off;
        }
    }
    bp.key_rows = key_rows;
    if (key_rows >= ((size_t)1 << 32)) return fail(ctx, UZ_ERR_INVALID, "batch too large: split it (key scratch > 2^32 rows)");

    std::vector<int4>& merges = ctx->merge_table;         // (key_off, nq, segments, 0) of every task cut into segments
    merges.clear();
    for (size_t t = 0; t < n_tasks; ++t) {
        const size_t k = qtiles_of(t) * (size_t)nseg_of(t);
        if (task_wide[t]) bp.n_tiles_wide += k; else bp.n_tiles_narrow += k;
        if (nseg_of(t) > 1 && tasks[t].nq > 0) merges.push_back(make_int4((int)tasks[t].key_off, tasks[t].nq, nseg_of(t), 0));
    }
    const size_t n_tiles = bp.n_tiles_narrow + bp.n_tiles_wide;
    const size_t narrow_bytes = (bp.n_tiles_narrow * bp.narrow_tile_bytes + 15) & ~(size_t)15;
    const size_t tiles_bytes = narrow_bytes + ((bp.n_tiles_wide * bp.wide_tile_bytes + 15) & ~(size_t)15);
    UZ_CUDA(ctx, sl.h_tiles.ensure(std::max<size_t>(tiles_bytes, 16) + merges.size() * sizeof(int4)));
    uint8_t* h_tiles = (uint8_t*)sl.h_tiles.p;
    bp.streaming = bp.may_stream && n_tiles > 0;
    int* pend = nullptr;
    if (bp.streaming) {
        UZ_CUDA(ctx, sl.h_pending.ensure((size_t)n_pairs * sizeof(int)));
        pend = (int*)sl.h_pending.p;
        memset(pend, 0, (size_t)n_pairs * sizeof(int));
    }
    {
        size_t kn = 0, kw = 0;
        auto emit_tiles = [&](size_t t) {
            const bool wide = task_wide[t] != 0;
            const int tr = bp.rows_of(t, task_wide);
            uint8_t* base = wide ? h_tiles + narrow_bytes : h_tiles;
            size_t& k = wide ? kw : kn;
            if (!seg_of(t)) {
                for (int q0 = 0; q0 < tasks[t].nq; q0 += tr) { ((int2*)base)[k++] = make_int2((int)t, q0); if (pend) pend[tasks[t].pair]++; }
                return;
            }
            const int sr = seg_rows_of(t), ns = nseg_of(t);
            for (int q0 = 0; q0 < tasks[t].nq; q0 += tr)
                for (int sg = 0; sg < ns; ++sg) {
                    const int tb = sg * sr, rows = std::max(0, std::min(tasks[t].nt - tb, sr));
                    ((int4*)base)[k++] = make_int4((int)t, q0, tb, (sg << 16) | rows);
                }
        };
        for (size_t t = 0; t < n_fwd; ++t) {
            emit_tiles(t);
            if (!fused && cross && tasks[t].nq > 0) emit_tiles((size_t)tasks[t].pad_);
        }
    }
    int4* h_merges = (int4*)(h_tiles + std::max<size_t>(tiles_bytes, 16));
    if (!merges.empty()) memcpy(h_merges, merges.data(), merges.size() * sizeof(int4));
    g_stage.stop(6);

    // (4) device buffers; small launches: the three tables travel as ONE copy (a copy command costs more than its few KB)
    const size_t all_tiles_bytes = std::max<size_t>(tiles_bytes, 16) + merges.size() * sizeof(int4);
    UZ_CUDA(ctx, sl.d_tasks.ensure(std::max<size_t>(n_tasks, 1) * sizeof(MatchTask)));
    UZ_CUDA(ctx, sl.d_tiles.ensure(all_tiles_bytes));
    UZ_CUDA(ctx, sl.d_pair_tasks.ensure((size_t)n_pairs * sizeof(int2)));
    UZ_CUDA(ctx, sl.d_keys.ensure(std::max<size_t>(key_rows, 1) * sizeof(uint2)));
    if (fused && key_rows > bp.col_begin)       // column keys start at "none"; the match kernel lowers them with atomicMin
        UZ_CUDA(ctx, cudaMemsetAsync((uint2*)sl.d_keys.p + bp.col_begin, 0xFF, (key_rows - bp.col_begin) * sizeof(uint2), ctx->stream));
    const size_t b_tasks = (n_tasks * sizeof(MatchTask) + 255) & ~(size_t)255;
    const size_t b_tiles = (all_tiles_bytes + 255) & ~(size_t)255;
    const size_t b_pairs = ((size_t)n_pairs * sizeof(int2) + 255) & ~(size_t)255;
    const bool one_copy = b_tasks + b_tiles + b_pairs <= ((size_t)64 << 10);
    const uint8_t* t_tiles = (const uint8_t*)sl.d_tiles.p;
    bp.d_tasks = (const MatchTask*)sl.d_tasks.p;
    bp.d_pair_tasks = (const int2*)sl.d_pair_tasks.p;
    if (one_copy) {
        UZ_CUDA(ctx, sl.h_tables.ensure(b_tasks + b_tiles + b_pairs));
        UZ_CUDA(ctx, sl.d_tables.ensure(b_tasks + b_tiles + b_pairs));
        uint8_t* hb = (uint8_t*)sl.h_tables.p;
        memcpy(hb, tasks, n_tasks * sizeof(MatchTask));
        memcpy(hb + b_tasks, h_tiles, all_tiles_bytes);
        memcpy(hb + b_tasks + b_tiles, bp.pair_tasks, (size_t)n_pairs * sizeof(int2));
        UZ_CUDA(ctx, cudaMemcpyAsync(sl.d_tables.p, hb, b_tasks + b_tiles + b_pairs, cudaMemcpyHostToDevice, ctx->stream));
        bp.d_tasks = (const MatchTask*)sl.d_tables.p;
        t_tiles = (const uint8_t*)sl.d_tables.p + b_tasks;
        bp.d_pair_tasks = (const int2*)((const uint8_t*)sl.d_tables.p + b_tasks + b_tiles);
    } else {
        if (n_tasks) UZ_CUDA(ctx, cudaMemcpyAsync(sl.d_tasks.p, tasks, n_tasks * sizeof(MatchTask), cudaMemcpyHostToDevice, ctx->stream));
        UZ_CUDA(ctx, cudaMemcpyAsync(sl.d_tiles.p, h_tiles, all_tiles_bytes, cudaMemcpyHostToDevice, ctx->stream));
        UZ_CUDA(ctx, cudaMemcpyAsync(sl.d_pair_tasks.p, bp.pair_tasks, (size_t)n_pairs * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    }
    g_stage.stop(7);
    bp.d_tiles_narrow = t_tiles;
    bp.d_tiles_wide = t_tiles + narrow_bytes;
    bp.d_merges = (const int4*)(t_tiles + std::max<size_t>(tiles_bytes, 16));

    // (5) launches.  Tensor-core match kernel: one persistent launch, then one solve CTA per pair behind it.  Integer-pipe
    // kernels, large batches: ONE match launch plus the persistent streaming solve beside it (uz_solve.cuh); small
    // batches, the parity taps and UZ_STREAM_SOLVE=0: match launch, then one solve CTA per pair behind it.
    const int cap = bp.cap;
    SolveParams sp;
    memset(&sp, 0, sizeof(sp));
    if (bp.with_solve) {
        const uint16_t* table = nullptr;
        st = ensure_samples(ctx, prm.ransac_iterations, prm.do_prosac, bp.max_nq, &table);
        if (st != UZ_OK) return st;
        sp.thr = prm.ransac_threshold; sp.thr_sq_star = thr_sq_star(prm.ransac_threshold);
        sp.break_pct = prm.break_percentage; sp.iterations = prm.ransac_iterations;
        sp.ratio_num = prm.ratio_num; sp.ratio_den = prm.ratio_den; sp.cap = cap;
        sp.samples = table; sp.samples_by_m = 1;
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
    const bool streaming = bp.streaming;
    const int stream_ctas = ctx->stream_solve_ctas * ctx->sm_count;
    if (ctx->timers) ctx->compares += bp.compares;

    int* d_pending = nullptr;
    int* d_deferred = nullptr;
    StreamCtl* d_ctl = nullptr;
    if (streaming) {
        UZ_CUDA(ctx, sl.d_pending.ensure((size_t)n_pairs * 2 * sizeof(int) + 256));
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
        const MatchTask* d_tk = bp.d_tasks;
        const int2* d_t = (const int2*)bp.d_tiles_narrow;
        const int2* d_tw = (const int2*)bp.d_tiles_wide;
        uint2* d_k = (uint2*)sl.d_keys.p;
        const int nt = (int)bp.n_tiles_narrow;
        unsigned int* d_prog = d_ctl ? &d_ctl->progress : nullptr;
        if (bp.n_tiles_wide > 0 && bp.mma_wide) {
            const int ntw = (int)bp.n_tiles_wide;
            if (ctx->wide_e4)            // default: 4-bit operands, K = 512 in eight instructions
                knn2_mmaf_kernel<true><<<std::min(ntw, ctx->sm_count), kF4Threads, F4<true>::kSmemBytes, ctx->stream>>>(
                    reinterpret_cast<const MmaTask*>(d_tk), d_tw, ntw, d_k, uz_knn2_mma_desc(), nullptr, nullptr, ctx->f4_zeros);
            else                         // two int8 planes (UZ_MATCH_MMA_WIDE=2)
                knn2_mmaw_kernel<<<std::min(ntw, ctx->sm_count), kMmaThreads, kMmawSmemBytes, ctx->stream>>>(
                    reinterpret_cast<const MmaTask*>(d_tk), d_tw, ntw, d_k, uz_knn2_mma_desc());
            ctx->launches++; ctx->mma_launches++;
            UZ_CUDA(ctx, cudaGetLastError());
            if (ctx->timers) ctx->match_launches++;
        } else if (bp.n_tiles_wide > 0) {
            if (bp.wide_cfg == 0) launch_knn2_wide<256>(ctx, d_tk, d_tw, (int)bp.n_tiles_wide, d_k, d_pending, d_prog, fused, bp.seg_wide);
            else launch_knn2_wide<64>(ctx, d_tk, d_tw, (int)bp.n_tiles_wide, d_k, d_pending, d_prog, fused, bp.seg_wide);
            ctx->launches++;
            UZ_CUDA(ctx, cudaGetLastError());
            if (ctx->timers) ctx->match_launches++;
        }
        if (nt > 0 && bp.mma2) {
            // persistent grid of CTA pairs (clusters of two = one TPC), items dealt round-robin
            cudaLaunchConfig_t cfg = {};
            const int clusters = std::min(nt, ctx->sm_count / 2);
            cfg.gridDim = dim3((unsigned)(2 * clusters)); cfg.blockDim = dim3(kMmaThreads); cfg.stream = ctx->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            const MmaTask* mt = reinterpret_cast<const MmaTask*>(d_tk);
            cfg.dynamicSmemBytes = kMma2SmemBytes;
            UZ_CUDA(ctx, cudaLaunchKernelEx(&cfg, knn2_mma2_kernel, mt, d_t, nt, d_k, uz_knn2_mma_desc()));
            ctx->mma_launches++;
        } else if (nt > 0 && bp.mma) {
            // persistent grid, one CTA per SM; items are dealt round-robin, so neighbouring SMs work on the same pair and
            // share its train rows in L2
            const MmaTask* mt = reinterpret_cast<const MmaTask*>(d_tk);
            const int grid = std::min(nt, ctx->sm_count);
            if (ctx->match_mma == 7)         // measured alternatives, kept for A/B (profiles/mma_experiments_r02.txt)
                knn2_mma_kernel<<<grid, kMmaThreads, kMmaSmemBytes, ctx->stream>>>(mt, d_t, nt, d_k, uz_knn2_mma_desc(), nullptr, nullptr);
            else if (!ctx->narrow_e4)        // int8 operands, keys out of the tensor core
                knn2_mmak_kernel<<<grid, kMmaThreads, kMmakSmemBytes, ctx->stream>>>(mt, d_t, nt, d_k, uz_knn2_mma_desc(), nullptr, nullptr);
            else                             // default: 4-bit operands, keys out of the tensor core
                knn2_mmaf_kernel<false><<<grid, kF4Threads, F4<false>::kSmemBytes, ctx->stream>>>(mt, d_t, nt, d_k, uz_knn2_mma_desc(), nullptr, nullptr, ctx->f4_zeros);
            ctx->mma_launches++;
        } else if (nt > 0) switch (bp.best_cfg) {
            case 0: launch_knn2<256, 4>(ctx, d_tk, d_t, nt, d_k, d_pending, d_prog, fused, bp.seg_narrow); break;
            case 1: launch_knn2<128, 4>(ctx, d_tk, d_t, nt, d_k, d_pending, d_prog, fused, bp.seg_narrow); break;
            case 2: launch_knn2<64, 2>(ctx, d_tk, d_t, nt, d_k, d_pending, d_prog, fused, bp.seg_narrow); break;
            case 3: launch_knn2<256, 2>(ctx, d_tk, d_t, nt, d_k, d_pending, d_prog, fused, bp.seg_narrow); break;
            case 4: launch_knn2<128, 2>(ctx, d_tk, d_t, nt, d_k, d_pending, d_prog, fused, bp.seg_narrow); break;
            default: launch_knn2<32, 2>(ctx, d_tk, d_t, nt, d_k, d_pending, d_prog, fused, bp.seg_narrow); break;
        }
        if (nt > 0) {
            ctx->launches++;
            UZ_CUDA(ctx, cudaGetLastError());
            if (ctx->timers) ctx->match_launches++;
        }
        if (!merges.empty()) {
            merge_segments_kernel<<<dim3((unsigned)((bp.max_nq + 255) / 256), (unsigned)merges.size(), 1), 256, 0, ctx->stream>>>(bp.d_merges, d_k);
            ctx->launches++;
            UZ_CUDA(ctx, cudaGetLastError());
        }
    }
    if (ctx->timers) cudaEventRecord(tm.e[1], ctx->stream);
    if (bp.with_solve) {
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
                bp.d_tasks, bp.d_pair_tasks, (const uint2*)sl.d_keys.p, sp, d_results,
                n_pairs, d_pending, d_ctl, d_deferred, 0);
        else
            launch_solve(ctx, n_pairs, cap, sB, bp.d_tasks, bp.d_pair_tasks, (const uint2*)sl.d_keys.p, sp, d_results);
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
                bp.d_tasks, bp.d_pair_tasks, (const uint2*)sl.d_keys.p, sp, d_results,
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
    g_stage.stop(8);
    UZ_CUDA(ctx, cudaEventRecord(sl.done, results_on));     // the slot's tables and keys are free once the solve is through
    sl.used = true;
    if (result_stream) *result_stream = results_on;
    return UZ_OK;
}

// A synchronous store-resident call (uz_estimate_edges, a group's shard) as a short pipeline: the batch is cut into chunks
// that double in size, so the GPU starts after the host has prepared the FIRST chunk (1480 pairs) and the tables of chunk
// c + 1 are built while chunk c runs.  One launch pair for 25 000 pairs leaves the GPU idle for the ~1.3 ms the host needs to
// enumerate tasks and tiles.  Tensor-core match path only (the integer-pipe kernels overlap their own solve per launch), not
// with the parity taps (they describe one launch pair).
// after_chunk(first pair, pairs): called when a chunk has been enqueued (e.g. to queue its records' way home).
uz_status run_pairs_pipelined(uz_context* ctx, const std::vector<PairRef>& pairs, uz_edge_result* d_results,
                              const std::function<uz_status(size_t, size_t)>& after_chunk = nullptr) {
    const size_t n = pairs.size();
    const size_t wave = (size_t)5 * (size_t)ctx->sm_count;
    if (ctx->debug || !ctx->match_mma || !ctx->pipeline_calls || n < 8 * wave) {
        const uz_status st = run_pairs(ctx, pairs, d_results);
        return st != UZ_OK || !after_chunk ? st : after_chunk(0, n);
    }
    std::vector<PairRef> part;
    size_t at = 0, take = 2 * wave;
    while (at < n) {
        size_t k = std::min(take, n - at);
        if (n - at - k < wave) k = n - at;              // no crumb at the end
        part.assign(pairs.begin() + at, pairs.begin() + at + k);
        uz_status st = run_pairs(ctx, part, d_results + at);
        if (st == UZ_OK && after_chunk) st = after_chunk(at, k);
        if (st != UZ_OK) return st;
        at += k;
        take *= 2;
    }
    return UZ_OK;
}

// run_pairs_pipelined with the records delivered to a host array: every chunk's records are copied into pinned memory behind
// its solve and from there into the caller's array while later chunks compute.  Returns with the stream idle.
uz_status run_pairs_to_host(uz_context* ctx, const std::vector<PairRef>& pairs, uz_edge_result* d_res, uz_edge_result* h_pinned,
                            uz_edge_result* results) {
    struct Home { size_t at, k; cudaEvent_t ev; };
    std::vector<Home> home;
    size_t drained = 0;
    auto drain = [&](bool wait) {
        for (; drained < home.size(); ++drained) {
            const Home& hm = home[drained];
            if (wait) { if (cudaEventSynchronize(hm.ev) != cudaSuccess) return; }
            else if (cudaEventQuery(hm.ev) != cudaSuccess) { cudaGetLastError(); return; }
            memcpy(results + hm.at, h_pinned + hm.at, hm.k * sizeof(uz_edge_result));
        }
    };
    uz_status st = run_pairs_pipelined(ctx, pairs, d_res, [&](size_t at, size_t k) -> uz_status {
        if (cudaMemcpyAsync(h_pinned + at, d_res + at, k * sizeof(uz_edge_result), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
            return fail(ctx, UZ_ERR_CUDA, "cudaMemcpyAsync(results) failed");
        cudaEvent_t ev = ctx->get_event();
        cudaEventRecord(ev, ctx->stream);
        home.push_back(Home{at, k, ev});
        drain(false);
        return UZ_OK;
    });
    if (st == UZ_OK) drain(true);
    for (auto& hm : home) ctx->event_pool.push_back(hm.ev);
    if (st != UZ_OK) { cudaStreamSynchronize(ctx->stream); return st; }
    if (drained != home.size()) return fail(ctx, UZ_ERR_CUDA, "cudaEventSynchronize failed");
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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

}  // namespace
