// uz_upload.inl — host buffers -> device layouts (part of the uz_capi.cu translation unit).
//
// A camera's arrays (descriptors, positions, valid flags; feature_transformation_estimator.cpp:40-49 reads them from
// FeatureData) land in ONE range per keyframe, laid out by cam_layout(); whatever the host layout is, the bytes travel in
// one gather launch per group:
//   * pinned, device-mapped sources are pulled straight over PCIe by gather_copy_kernel;
//   * pageable sources (what a cv::Mat / Eigen matrix normally is) are first packed into a pinned ring by a few host
//     threads and pulled from there by the same kernel - no per-array cudaMemcpyAsync, no synchronous staging in the driver;
//   * a handful of large arrays goes through the DMA engines (cudaMemcpyAsync) as before.
// Then one launch derives the CSA layout (integer-pipe kernels) and the E8 layout (tensor-core kernel) of every camera.
namespace {

struct CopyItem {
    const uint8_t* host; uint8_t* dev;
    size_t bytes;             // packed bytes on the device (rows * row_bytes when strided)
    int32_t rows, row_bytes, stride;   // strided source (cv::Mat with padded rows): rows > 0
};

struct DeriveJob { const uint32_t* raw; uint32_t* csa; uint8_t* e8; int32_t n; int16_t halves_per_row; int16_t e4; };

// blockIdx.y = camera.  CSA transform of every 256-bit half (uz_knn2.cuh) and the tensor-core operand layout: 4-bit values
// (uz_knn2_mmaf.cuh) or int8 (uz_knn2_mma.cuh) in the UMMA canonical layout.
__global__ void __launch_bounds__(256) derive_layouts_kernel(const DeriveJob* __restrict__ jobs) {
    const DeriveJob j = jobs[blockIdx.y];
    const int halves = j.n * j.halves_per_row;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < halves; i += gridDim.x * blockDim.x) {
        const uint4* p = reinterpret_cast<const uint4*>(j.raw + (size_t)i * 8);
        const uint4 a = p[0], b = p[1];
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t o[8];
        csa_pack(w, o);
        uint4* c = reinterpret_cast<uint4*>(j.csa + (size_t)i * 8);
        c[0] = make_uint4(o[0], o[1], o[2], o[3]); c[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
    if (j.e8 == nullptr) return;
    if (j.e4) {
        // 4-bit operands (uz_knn2_mmaf.cuh): one thread per (row, 32-bit word): 32 nibbles = one uint4 store; a 64-byte row
        // is sixteen words and an 8-row group 2 KB
        // (the rows that complete the last 8-row group hold zero nibbles: they add nothing to a sum)
        const int words = 8 * j.halves_per_row;
        const size_t group = (size_t)kF4GroupBytes * j.halves_per_row;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ((j.n + 7) & ~7) * words; i += gridDim.x * blockDim.x) {
            const int row = i / words, w = i % words;
            if (row >= j.n) {
                *reinterpret_cast<uint4*>(j.e8 + (size_t)(row >> 3) * group + w * 128 + (row & 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
                continue;
            }
            const uint32_t bits = j.raw[(size_t)row * words + w];
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t v = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b) v |= (((bits >> (8 * k + b)) & 1u) ? kE4Set : kE4Clear) << (4 * b);
                o[k] = v;
            }
            *reinterpret_cast<uint4*>(j.e8 + (size_t)(row >> 3) * group + w * 128 + (row & 7) * 16) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        return;
    }
    // one thread per (row, 16-bit chunk): 16 int8 = one uint4 store.  64-byte rows: chunk 16..31 goes to the second plane.
    const int chunks = 16 * j.halves_per_row;
    const size_t plane = e8_bytes(j.n);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.n * chunks; i += gridDim.x * blockDim.x) {
        const int row = i / chunks, c = i % chunks;
        const uint32_t w = j.raw[(size_t)row * (8 * j.halves_per_row) + (c >> 1)];
        const uint32_t bits = (c & 1) ? (w >> 16) : (w & 0xFFFFu);
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) v |= (((bits >> (4 * k + b)) & 1u) ? kE8Set : kE8Clear) << (8 * b);
            o[k] = v;
        }
        *reinterpret_cast<uint4*>(j.e8 + (size_t)(c >> 4) * plane + (size_t)(row >> 3) * kE8GroupBytes + (c & 15) * 128 + (row & 7) * 16) =
            make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// Device-side address of a pinned, device-mapped host range, or 0 if the range is not (entirely) mapped.  The cache
// lives for one entry-point call (check_ctx clears it).
uintptr_t mapped_device_address(uz_context* ctx, const uint8_t* host, size_t bytes) {
    const uintptr_t h = (uintptr_t)host;
    for (const auto& r : ctx->mapped)
        if (h >= r.hb && h + bytes <= r.he) { ctx->map_hits++; return r.db + (h - r.hb); }
    // a call whose buffers all turned out pageable so far stops asking the driver (staging a pinned buffer through the
    // ring is still correct, only slower)
    if (ctx->map_hits == 0 && ctx->map_misses >= 16) return 0;
    ctx->map_misses++;
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
    if (h >= r.hb && h + bytes <= r.he) { ctx->map_hits++; ctx->map_misses--; return r.db + (h - r.hb); }
    return 0;
}

uz_status ring_ensure(uz_context* ctx) {
    if (ctx->ring.p) return UZ_OK;
    UZ_CUDA(ctx, ctx->ring.ensure(2 * ctx->ring_half));
    void* d = nullptr;
    UZ_CUDA(ctx, cudaHostGetDevicePointer(&d, ctx->ring.p, 0));
    ctx->ring_dev = (uintptr_t)d;
    for (int i = 0; i < 2; ++i) UZ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ring_free[i], cudaEventDisableTiming));
    return UZ_OK;
}

// one gather launch over a table of chunks (<= 16 KB each)
uz_status launch_gather(uz_context* ctx, const std::vector<CopyChunk>& cc) {
    if (cc.empty()) return UZ_OK;
    CopyChunk* h = (CopyChunk*)ctx->h_chunks.alloc(cc.size() * sizeof(CopyChunk));
    CopyChunk* d = (CopyChunk*)ctx->d_chunks.alloc(cc.size() * sizeof(CopyChunk));
    if (!h || !d) return fail(ctx, UZ_ERR_NOMEM, "copy-chunk table allocation failed");
    memcpy(h, cc.data(), cc.size() * sizeof(CopyChunk));
    UZ_CUDA(ctx, cudaMemcpyAsync(d, h, cc.size() * sizeof(CopyChunk), cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->copy_beside_compute)      // few small CTAs: they fit beside the compute kernels of the previous chunk
        gather_copy_kernel<<<(unsigned)std::min<size_t>(cc.size(), (size_t)ctx->copy_ctas), 128, 0, ctx->stream>>>(d, (int)cc.size());
    else
        gather_copy_kernel<<<(unsigned)std::min<size_t>(cc.size(), (size_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(d, (int)cc.size());
    ctx->launches++;
    UZ_CUDA(ctx, cudaGetLastError());
    return UZ_OK;
}

void push_chunks(std::vector<CopyChunk>& cc, uintptr_t src, uint8_t* dst, size_t bytes) {
    const size_t kChunk = 16384;
    for (size_t off = 0; off < bytes; off += kChunk) {
        CopyChunk c;
        c.src = (const uint8_t*)(src + off); c.dst = dst + off; c.bytes = (uint32_t)std::min(kChunk, bytes - off); c.pad = 0;
        cc.push_back(c);
    }
}

void stage_copy(const CopyItem& it, size_t from, size_t count, uint8_t* dst) {
    // bytes [from, from + count) of the item's PACKED image -> dst
    if (it.rows <= 0 || it.stride == it.row_bytes) { memcpy(dst, it.host + from, count); return; }
    size_t done = 0;
    while (done < count) {
        const size_t at = from + done, r = at / (size_t)it.row_bytes, o = at % (size_t)it.row_bytes;
        const size_t k = std::min(count - done, (size_t)it.row_bytes - o);
        memcpy(dst + done, it.host + r * (size_t)it.stride + o, k);
        done += k;
    }
}

// Moves a set of host buffers.  Order of the device writes is the stream's; the host buffers are free again once the
// stream has passed the last gather launch (pinned sources) or when this function returns (pageable sources, which are
// copied into the ring here).
uz_status flush_copies(uz_context* ctx, const std::vector<CopyItem>& items) {
    if (items.empty()) return UZ_OK;
    size_t total = 0;
    bool strided = false;
    for (const auto& it : items) { total += it.bytes; strided |= it.rows > 0 && it.stride != it.row_bytes; }
    if (!ctx->gather_upload) {                                            // UZ_GATHER_UPLOAD=0: plain copies, one per array
        for (const auto& it : items) {
            if (!it.bytes) continue;
            if (it.rows > 0 && it.stride != it.row_bytes)
                UZ_CUDA(ctx, cudaMemcpy2DAsync(it.dev, it.row_bytes, it.host, it.stride, it.row_bytes, it.rows, cudaMemcpyHostToDevice, ctx->stream));
            else
                UZ_CUDA(ctx, cudaMemcpyAsync(it.dev, it.host, it.bytes, cudaMemcpyHostToDevice, ctx->stream));
        }
        return UZ_OK;
    }
    if (items.size() <= 16 && !strided) {                                 // few (large) arrays: the DMA engines
        for (const auto& it : items)
            if (it.bytes) UZ_CUDA(ctx, cudaMemcpyAsync(it.dev, it.host, it.bytes, cudaMemcpyHostToDevice, ctx->stream));
        return UZ_OK;
    }
    std::vector<CopyChunk> cc;
    cc.reserve(total / 16384 + items.size());
    struct Staged { size_t item, from, count, ring_off; };
    std::vector<Staged> group;            // what the current ring half holds
    uz_status st = UZ_OK;
    auto flush_group = [&]() -> uz_status {
        const int half = ctx->ring_cur;
        if (!group.empty()) {
            // pack into the pinned ring with a few persistent host threads
            size_t bytes = 0;
            for (const auto& g : group) bytes += g.count;
            uint8_t* base = (uint8_t*)ctx->ring.p + (size_t)half * ctx->ring_half;
            if (!ctx->pool && bytes >= ((size_t)1 << 20)) {
                unsigned want = ctx->stage_threads > 0 ? (unsigned)ctx->stage_threads
                                                       : std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
                if (want > 1) { ctx->pool = new HostPool(); ctx->pool->start(want - 1); }
            }
            if (ctx->pool && bytes >= ((size_t)1 << 20)) {
                const unsigned nthr = (unsigned)ctx->pool->th.size() + 1;
                ctx->pool->run([&](unsigned t) {
                    for (size_t k = t; k < group.size(); k += nthr)
                        stage_copy(items[group[k].item], group[k].from, group[k].count, base + group[k].ring_off);
                });
            } else {
                for (const auto& g : group) stage_copy(items[g.item], g.from, g.count, base + g.ring_off);
            }
            for (const auto& g : group)
                push_chunks(cc, ctx->ring_dev + (size_t)half * ctx->ring_half + g.ring_off, items[g.item].dev + g.from, g.count);
        }
        uz_status s = launch_gather(ctx, cc);
        cc.clear();
        if (s != UZ_OK) return s;
        if (!group.empty()) {
            UZ_CUDA(ctx, cudaEventRecord(ctx->ring_free[half], ctx->stream));
            ctx->ring_busy[half] = true;
            ctx->ring_cur ^= 1;
            ctx->ring_used[ctx->ring_cur] = 0;
            if (ctx->ring_busy[ctx->ring_cur]) {
                UZ_CUDA(ctx, cudaEventSynchronize(ctx->ring_free[ctx->ring_cur]));
                ctx->ring_busy[ctx->ring_cur] = false;
            }
            group.clear();
        }
        return UZ_OK;
    };
    for (size_t i = 0; i < items.size(); ++i) {
        const CopyItem& it = items[i];
        if (!it.bytes) continue;
        const bool is_strided = it.rows > 0 && it.stride != it.row_bytes;
        const uintptr_t dsrc = is_strided ? 0 : mapped_device_address(ctx, it.host, it.bytes);
        if (dsrc) { push_chunks(cc, dsrc, it.dev, it.bytes); continue; }
        if ((st = ring_ensure(ctx)) != UZ_OK) return st;
        if (ctx->ring_busy[ctx->ring_cur] && ctx->ring_used[ctx->ring_cur] == 0) {      // left busy by an earlier call
            UZ_CUDA(ctx, cudaEventSynchronize(ctx->ring_free[ctx->ring_cur]));
            ctx->ring_busy[ctx->ring_cur] = false;
        }
        size_t from = 0;
        while (from < it.bytes) {
            size_t& used = ctx->ring_used[ctx->ring_cur];
            if (used + 16 > ctx->ring_half) { if ((st = flush_group()) != UZ_OK) return st; continue; }
            const size_t k = std::min(it.bytes - from, (ctx->ring_half - used) & ~(size_t)15);
            group.push_back(Staged{i, from, k, used});
            used = (used + k + 15) & ~(size_t)15;
            from += k;
        }
    }
    return flush_group();
}

uz_status validate_features(uz_context* ctx, const uz_features* f) {
    if (f->n < 0 || f->n > UZ_MAX_FEATURES) return fail(ctx, UZ_ERR_INVALID, "feature count out of range (0..UZ_MAX_FEATURES)");
    if (f->n > 0 && (!f->descriptors || !f->positions || !f->valid_3d)) return fail(ctx, UZ_ERR_INVALID, "null feature buffer");
    const int db = desc_width(f->desc_bytes);
    if (db == 0) return fail(ctx, UZ_ERR_UNSUPPORTED, "descriptor width must be 32 or 64 bytes (256- or 512-bit binary descriptors)");
    if (f->n > 0 && f->desc_stride < db) return fail(ctx, UZ_ERR_INVALID, "descriptor stride < descriptor width");
    return UZ_OK;
}

// CSA + E8 layouts of cameras whose raw rows are already on the device (one launch per 32768 cameras)
uz_status derive_layouts(uz_context* ctx, const Cam* cams, size_t n_cams) {
    std::vector<DeriveJob> jobs;
    int max_units = 1;
    for (size_t i = 0; i < n_cams; ++i) {
        const Cam& c = cams[i];
        if (c.n == 0) continue;
        jobs.push_back(DeriveJob{c.raw, c.csa, c.e8, c.n, (int16_t)(c.dbytes / 32), (int16_t)(c.dbytes == UZ_DESC_BYTES ? ctx->narrow_e4 : ctx->wide_e4)});
        max_units = std::max(max_units, c.n * 16 * (c.dbytes / 32));
    }
    for (size_t j0 = 0; j0 < jobs.size(); j0 += 32768) {
        const size_t cnt = std::min<size_t>(32768, jobs.size() - j0);
        DeriveJob* h = (DeriveJob*)ctx->h_chunks.alloc(cnt * sizeof(DeriveJob));
        DeriveJob* d = (DeriveJob*)ctx->d_chunks.alloc(cnt * sizeof(DeriveJob));
        if (!h || !d) return fail(ctx, UZ_ERR_NOMEM, "layout job table allocation failed");
        memcpy(h, jobs.data() + j0, cnt * sizeof(DeriveJob));
        UZ_CUDA(ctx, cudaMemcpyAsync(d, h, cnt * sizeof(DeriveJob), cudaMemcpyHostToDevice, ctx->stream));
        derive_layouts_kernel<<<dim3((unsigned)std::min((max_units + 255) / 256, 64), (unsigned)cnt, 1), 256, 0, ctx->stream>>>(d);
        ctx->launches++;
        UZ_CUDA(ctx, cudaGetLastError());
    }
    return UZ_OK;
}

// Lays cameras out on the device.  group_sizes == nullptr: ONE range for all cameras (transients; *blocks gets one
// entry); otherwise one range per keyframe (group_sizes[k] cameras each), so that a keyframe can be freed on its own.
uz_status place_cams(uz_context* ctx, Arena& arena, const std::vector<const uz_features*>& feats, const int32_t* group_sizes,
                     size_t n_groups, std::vector<Cam>& out, std::vector<BlockRef>& blocks) {
    out.assign(feats.size(), Cam());
    blocks.clear();
    for (size_t i = 0; i < feats.size(); ++i) {
        uz_status st = validate_features(ctx, feats[i]);
        if (st != UZ_OK) return st;
    }
    size_t k = 0;
    const size_t groups = group_sizes ? n_groups : 1;
    for (size_t g = 0; g < groups; ++g) {
        const size_t cnt = group_sizes ? (size_t)group_sizes[g] : feats.size();
        size_t at = 0;
        std::vector<CamLayout> lay(cnt);
        for (size_t c = 0; c < cnt; ++c) {
            const uz_features* f = feats[k + c];
            lay[c] = cam_layout(at, f->n, desc_width(f->desc_bytes), ctx->operand_fmt());
            at = lay[c].end;
        }
        uint8_t* base = (uint8_t*)arena.alloc(std::max<size_t>(at, 1));
        if (!base) {
            for (auto& b : blocks) arena.free(b.p, b.bytes);
            blocks.clear();
            return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
        }
        blocks.push_back(BlockRef{base, std::max<size_t>(at, 1)});
        for (size_t c = 0; c < cnt; ++c) {
            const uz_features* f = feats[k + c];
            Cam& o = out[k + c];
            o.n = f->n; o.feature_type = f->feature_type; o.sensor_frame = f->sensor_frame; o.dbytes = desc_width(f->desc_bytes);
            if (f->n == 0) continue;
            o.raw = (uint32_t*)(base + lay[c].raw); o.pos = (double*)(base + lay[c].pos); o.valid = base + lay[c].valid;
            o.csa = (uint32_t*)(base + lay[c].csa);
            o.e8 = base + lay[c].e8;
        }
        k += cnt;
    }
    return UZ_OK;
}

// Uploads cameras (descriptors, positions, valid) into already placed device views and derives the other layouts.
uz_status fill_cams(uz_context* ctx, const std::vector<const uz_features*>& feats, const std::vector<Cam>& cams) {
    std::vector<CopyItem> items;
    items.reserve(feats.size() * 3);
    for (size_t i = 0; i < feats.size(); ++i) {
        const uz_features* f = feats[i];
        const Cam& c = cams[i];
        if (f->n == 0) continue;
        items.push_back(CopyItem{f->descriptors, (uint8_t*)c.raw, (size_t)f->n * c.dbytes, f->n, c.dbytes, f->desc_stride});
        items.push_back(CopyItem{(const uint8_t*)f->positions, (uint8_t*)c.pos, (size_t)f->n * 24, 0, 0, 0});
        items.push_back(CopyItem{f->valid_3d, c.valid, (size_t)f->n, 0, 0, 0});
    }
    uz_status st = flush_copies(ctx, items);
    if (st != UZ_OK) return st;
    if (ctx->trace_mid) cudaEventRecord(ctx->trace_mid, ctx->stream);
    return derive_layouts(ctx, cams.data(), cams.size());
}

// transient upload: one range for everything, freed by the arena's reset()
uz_status upload_cams(uz_context* ctx, Arena& arena, const std::vector<const uz_features*>& feats, std::vector<Cam>& out) {
    std::vector<BlockRef> blocks;
    g_stage.start();
    uz_status st = place_cams(ctx, arena, feats, nullptr, 0, out, blocks);
    g_stage.stop(1);
    if (st != UZ_OK) return st;
    st = fill_cams(ctx, feats, out);
    g_stage.stop(2);
    return st;
}

// ---- sample tables ---------------------------------------------------------------------------------
// One table per (iterations, do_prosac), a few kept (LRU): the reference alternates estimateEdge (100 iterations, PROSAC)
// with calcValidEdges (200, none) on one estimator (transformation_filter.cpp:272).
uz_status ensure_samples(uz_context* ctx, int iterations, int do_prosac, int max_m, const uint16_t** table_out) {
    uz_context::SampleTable* hit = nullptr;
    uz_context::SampleTable* lru = &ctx->samples[0];
    for (auto& t : ctx->samples) {
        if (t.iters == iterations && t.prosac == do_prosac) hit = &t;
        if (t.last_use < lru->last_use) lru = &t;
    }
    if (hit && hit->cap >= max_m) { hit->last_use = ++ctx->sample_clock; *table_out = (const uint16_t*)hit->d.p; return UZ_OK; }
    uz_context::SampleTable* t = hit ? hit : lru;
    int cap = std::max(256, (max_m + 255) & ~255);
    if (hit) cap = std::max(cap, hit->cap);
    std::vector<uint16_t> table;
    build_sample_table(iterations, do_prosac != 0, cap, table);
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // previous launches may still read the old table
    if (ctx->solve_stream) UZ_CUDA(ctx, cudaStreamSynchronize(ctx->solve_stream));
    if (ctx->alt) UZ_CUDA(ctx, cudaStreamSynchronize(ctx->alt));
    UZ_CUDA(ctx, t->d.ensure(table.size() * sizeof(uint16_t)));
    // stream-ordered: a plain cudaMemcpy from pageable memory runs on the legacy stream, which a
    // non-blocking stream does not wait for
    UZ_CUDA(ctx, cudaMemcpyAsync(t->d.p, table.data(), table.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    t->cap = cap; t->iters = iterations; t->prosac = do_prosac; t->last_use = ++ctx->sample_clock;
    *table_out = (const uint16_t*)t->d.p;
    return UZ_OK;
}

}  // namespace
