// uz_capi_ingest.inl — host side of device ingestion (included at the end of uz_capi.cu; kernels in uz_ingest.cuh).

namespace {

CameraModel to_model(const uz_camera* c) {
    CameraModel m;
    m.fx = c->fx; m.fy = c->fy; m.cx = c->cx; m.cy = c->cy; m.max_depth = c->max_depth; m.width = c->width; m.height = c->height;
    return m;
}

uz_status check_camera(uz_context* ctx, const uz_camera* c, const float* depth, int32_t stride_bytes) {
    if (!c || !depth) return fail(ctx, UZ_ERR_INVALID, "null camera / depth image");
    if (c->width <= 0 || c->height <= 0 || c->width > 16384 || c->height > 16384) return fail(ctx, UZ_ERR_INVALID, "depth image size out of range");
    if (stride_bytes < c->width * 4 || (stride_bytes & 3)) return fail(ctx, UZ_ERR_INVALID, "depth stride must be a multiple of 4 and >= 4*width");
    return UZ_OK;
}

// registers one single-camera keyframe whose buffers are already on the device (one range of the store arena)
int32_t register_keyframe(uz_context* ctx, const Cam& c, const BlockRef& block) {
    int32_t h = -1;
    std::vector<Cam> up(1, c);
    std::vector<BlockRef> blocks(1, block);
    const int32_t one = 1;
    register_keyframes(ctx, up, blocks, &one, 1, &h);
    return h;
}

// one camera's layouts in one range of the arena (the same layout uz_store_add uses)
uz_status alloc_cam(uz_context* ctx, Arena& arena, int n, int dbytes, int feature_type, int sensor_frame, Cam& c, BlockRef& block) {
    c = Cam();
    c.n = n; c.feature_type = feature_type; c.sensor_frame = sensor_frame; c.dbytes = dbytes;
    const CamLayout L = cam_layout(0, n, dbytes, ctx->operand_fmt());
    uint8_t* base = (uint8_t*)arena.alloc(std::max<size_t>(L.end, 1));
    if (!base) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
    block = BlockRef{base, std::max<size_t>(L.end, 1)};
    if (n == 0) return UZ_OK;
    c.raw = (uint32_t*)(base + L.raw); c.pos = (double*)(base + L.pos); c.valid = base + L.valid;
    c.csa = (uint32_t*)(base + L.csa);
    c.e8 = base + L.e8;
    return UZ_OK;
}

}  // namespace

extern "C" {

uz_status uz_backproject(uz_context* ctx, const int32_t* u, const int32_t* v, int32_t n, const float* depth,
                         int32_t depth_stride_bytes, const uz_camera* cam, int32_t reverse, double* positions_out,
                         uint8_t* valid_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (n < 0 || (n > 0 && (!u || !v || !positions_out || !valid_out))) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    if ((st = check_camera(ctx, cam, depth, depth_stride_bytes)) != UZ_OK) return st;
    if (n == 0) return UZ_OK;
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->transient.reset();
    const size_t img = (size_t)depth_stride_bytes * cam->height;
    int32_t* du = (int32_t*)ctx->transient.alloc((size_t)n * 4);
    int32_t* dv = (int32_t*)ctx->transient.alloc((size_t)n * 4);
    float* dd = (float*)ctx->transient.alloc(img);
    double* dp = (double*)ctx->transient.alloc((size_t)n * 24);
    uint8_t* dval = (uint8_t*)ctx->transient.alloc((size_t)n);
    if (!du || !dv || !dd || !dp || !dval) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
    UZ_CUDA(ctx, cudaMemcpyAsync(du, u, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    UZ_CUDA(ctx, cudaMemcpyAsync(dv, v, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    UZ_CUDA(ctx, cudaMemcpyAsync(dd, depth, img, cudaMemcpyHostToDevice, ctx->stream));
    backproject_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(du, dv, n, dd, depth_stride_bytes / 4, to_model(cam), reverse ? 1 : 0, dp, dval);
    ctx->launches++;
    UZ_CUDA(ctx, cudaGetLastError());
    UZ_CUDA(ctx, cudaMemcpyAsync(positions_out, dp, (size_t)n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaMemcpyAsync(valid_out, dval, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UZ_OK;
}

uz_status uz_store_add_rgbd(uz_context* ctx, const uint8_t* descriptors, int32_t desc_stride, int32_t desc_bytes,
                            const int32_t* u, const int32_t* v, int32_t n, const float* depth, int32_t depth_stride_bytes, const uz_camera* cam, int32_t feature_type,
                            int32_t sensor_frame, int32_t reverse, int32_t* handle_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (!handle_out || n < 0 || n > UZ_MAX_FEATURES || (n > 0 && (!descriptors || !u || !v))) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    const int db = desc_width(desc_bytes);
    if (db == 0) return fail(ctx, UZ_ERR_UNSUPPORTED, "descriptor width must be 32 or 64 bytes");
    if (n > 0 && desc_stride < db) return fail(ctx, UZ_ERR_INVALID, "descriptor stride < descriptor width");
    if ((st = check_camera(ctx, cam, depth, depth_stride_bytes)) != UZ_OK) return st;
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->transient.reset();
    Cam c;
    BlockRef block;
    ctx->h_chunks.reset(); ctx->d_chunks.reset();
    if ((st = alloc_cam(ctx, ctx->store_arena, n, db, feature_type, sensor_frame, c, block)) != UZ_OK) return st;
    if (n > 0) {
        const int halves = n * (db / 32);
        const size_t img = (size_t)depth_stride_bytes * cam->height;
        const size_t dbytes = (size_t)(n - 1) * desc_stride + db;
        int32_t* du = (int32_t*)ctx->transient.alloc((size_t)n * 4);
        int32_t* dv = (int32_t*)ctx->transient.alloc((size_t)n * 4);
        float* dd = (float*)ctx->transient.alloc(img);
        uint8_t* ddesc = (uint8_t*)ctx->transient.alloc(dbytes);
        if (!du || !dv || !dd || !ddesc) { ctx->store_arena.free(block.p, block.bytes); return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed"); }
        UZ_CUDA(ctx, cudaMemcpyAsync(du, u, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
        UZ_CUDA(ctx, cudaMemcpyAsync(dv, v, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
        UZ_CUDA(ctx, cudaMemcpyAsync(dd, depth, img, cudaMemcpyHostToDevice, ctx->stream));
        UZ_CUDA(ctx, cudaMemcpyAsync(ddesc, descriptors, dbytes, cudaMemcpyHostToDevice, ctx->stream));
        backproject_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(du, dv, n, dd, depth_stride_bytes / 4, to_model(cam), reverse ? 1 : 0, c.pos, c.valid);
        if (reverse) {
            reverse_rows_kernel<<<(n * (db / 4) + 255) / 256, 256, 0, ctx->stream>>>(ddesc, n, desc_stride, db / 4, c.raw);
            pack_descriptors_kernel<<<(halves + 255) / 256, 256, 0, ctx->stream>>>((const uint8_t*)c.raw, halves, 32, c.raw, c.csa, 1);
            ctx->launches++;
        } else {
            pack_descriptors_kernel<<<(halves + 255) / 256, 256, 0, ctx->stream>>>(ddesc, halves, desc_stride, c.raw, c.csa, db / 32);
        }
        ctx->launches += 2;
        UZ_CUDA(ctx, cudaGetLastError());
        if ((st = derive_layouts(ctx, &c, 1)) != UZ_OK) return st;      // (the CSA pass above is repeated there; E8 is what it adds)
        UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));        // host buffers are borrowed only for the call
    }
    *handle_out = register_keyframe(ctx, c, block);
    return UZ_OK;
}

// blob: the serialised graph_slam_msgs/Feature[] field (uint32 count, then the elements)
static uz_status wire_prepare(uz_context* ctx, const uint8_t* blob, size_t blob_bytes, int32_t* n_out, int32_t* cols_out) {
    *cols_out = UZ_DESC_BYTES;
    if (!blob || blob_bytes < 4) return fail(ctx, UZ_ERR_INVALID, "feature blob too short");
    uint32_t n;
    memcpy(&n, blob, 4);
    if (n > UZ_MAX_FEATURES) return fail(ctx, UZ_ERR_INVALID, "feature count out of range (0..UZ_MAX_FEATURES)");
    if (n > 0) {
        if (blob_bytes < 4 + 17) return fail(ctx, UZ_ERR_INVALID, "feature blob truncated");
        uint32_t len;
        memcpy(&len, blob + 4 + 13, 4);
        if (len != UZ_DESC_BYTES && len != UZ_MAX_DESC_BYTES) return fail(ctx, UZ_ERR_UNSUPPORTED, "descriptor length must be 32 or 64 (256- or 512-bit binary descriptors)");
        if (blob_bytes < 4 + (size_t)n * wire_elem_bytes((int)len)) return fail(ctx, UZ_ERR_INVALID, "feature blob truncated");
        *cols_out = (int32_t)len;
    }
    *n_out = (int32_t)n;
    return UZ_OK;
}

uz_status uz_wire_decode(uz_context* ctx, const uint8_t* blob, size_t blob_bytes, int32_t capacity, int32_t* n_out,
                         int32_t* desc_bytes_out, uint8_t* descriptors_out, double* positions_out, uint8_t* valid_out,
                         int32_t* uv_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    int32_t n = 0, cols = UZ_DESC_BYTES;
    if ((st = wire_prepare(ctx, blob, blob_bytes, &n, &cols)) != UZ_OK) return st;
    if (n_out) *n_out = n;
    if (desc_bytes_out) *desc_bytes_out = cols;
    if (n == 0) return UZ_OK;
    if (n > capacity || !descriptors_out || !positions_out || !valid_out) return fail(ctx, UZ_ERR_INVALID, "output capacity too small / null output");
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->transient.reset();
    const size_t body = (size_t)n * wire_elem_bytes(cols);
    uint8_t* db = (uint8_t*)ctx->transient.alloc(body);
    uint8_t* dd = (uint8_t*)ctx->transient.alloc((size_t)n * cols);
    double* dp = (double*)ctx->transient.alloc((size_t)n * 24);
    uint8_t* dv = (uint8_t*)ctx->transient.alloc((size_t)n);
    int32_t* duv = (int32_t*)ctx->transient.alloc((size_t)n * 8);
    int* dstat = (int*)ctx->transient.alloc(4);
    if (!db || !dd || !dp || !dv || !duv || !dstat) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
    UZ_CUDA(ctx, cudaMemcpyAsync(db, blob + 4, body, cudaMemcpyHostToDevice, ctx->stream));
    UZ_CUDA(ctx, cudaMemsetAsync(dstat, 0, 4, ctx->stream));
    wire_decode_kernel<<<(n * 32 + 255) / 256, 256, 0, ctx->stream>>>(db, n, cols, dd, dp, dv, duv, dstat);
    ctx->launches++;
    UZ_CUDA(ctx, cudaGetLastError());
    int stat = 0;
    UZ_CUDA(ctx, cudaMemcpyAsync(&stat, dstat, 4, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaMemcpyAsync(descriptors_out, dd, (size_t)n * cols, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaMemcpyAsync(positions_out, dp, (size_t)n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaMemcpyAsync(valid_out, dv, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (uv_out) UZ_CUDA(ctx, cudaMemcpyAsync(uv_out, duv, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (stat) return fail(ctx, UZ_ERR_UNSUPPORTED, "an element's descriptor length differs from the first element's");
    return UZ_OK;
}

uz_status uz_store_add_wire(uz_context* ctx, const uint8_t* blob, size_t blob_bytes, int32_t feature_type, int32_t sensor_frame,
                            int32_t* handle_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (!handle_out) return UZ_ERR_INVALID;
    int32_t n = 0, cols = UZ_DESC_BYTES;
    if ((st = wire_prepare(ctx, blob, blob_bytes, &n, &cols)) != UZ_OK) return st;
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->transient.reset();
    Cam c;
    BlockRef block;
    ctx->h_chunks.reset(); ctx->d_chunks.reset();
    if ((st = alloc_cam(ctx, ctx->store_arena, n, cols, feature_type, sensor_frame, c, block)) != UZ_OK) return st;
    if (n > 0) {
        const size_t body = (size_t)n * wire_elem_bytes(cols);
        uint8_t* db = (uint8_t*)ctx->transient.alloc(body);
        int* dstat = (int*)ctx->transient.alloc(4);
        if (!db || !dstat) { ctx->store_arena.free(block.p, block.bytes); return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed"); }
        UZ_CUDA(ctx, cudaMemcpyAsync(db, blob + 4, body, cudaMemcpyHostToDevice, ctx->stream));
        UZ_CUDA(ctx, cudaMemsetAsync(dstat, 0, 4, ctx->stream));
        wire_decode_kernel<<<(n * 32 + 255) / 256, 256, 0, ctx->stream>>>(db, n, cols, (uint8_t*)c.raw, c.pos, c.valid, nullptr, dstat);
        ctx->launches++;
        UZ_CUDA(ctx, cudaGetLastError());
        if ((st = derive_layouts(ctx, &c, 1)) != UZ_OK) return st;
        int stat = 0;
        UZ_CUDA(ctx, cudaMemcpyAsync(&stat, dstat, 4, cudaMemcpyDeviceToHost, ctx->stream));
        UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (stat) { ctx->store_arena.free(block.p, block.bytes); return fail(ctx, UZ_ERR_UNSUPPORTED, "an element's descriptor length differs from the first element's"); }
    }
    *handle_out = register_keyframe(ctx, c, block);
    return UZ_OK;
}

// ---- resume: many serialised keyframes at once (second half) ------------------------------------------
uz_status uz_store_add_wire_bulk(uz_context* ctx, const uint8_t* const* blobs, const size_t* blob_bytes, const int32_t* feature_types,
                                 const int32_t* sensor_frames, const int32_t* cams_per_keyframe, int32_t n_keyframes,
                                 int32_t* handles_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (n_keyframes < 0 || (n_keyframes > 0 && (!cams_per_keyframe || !handles_out))) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    size_t total = 0;
    for (int i = 0; i < n_keyframes; ++i) {
        if (cams_per_keyframe[i] < 0) return fail(ctx, UZ_ERR_INVALID, "negative camera count");
        total += (size_t)cams_per_keyframe[i];
    }
    if (total > 0 && (!blobs || !blob_bytes || !feature_types || !sensor_frames)) return fail(ctx, UZ_ERR_INVALID, "null blob arrays");
    if (n_keyframes == 0) return UZ_OK;
    // pseudo feature views: only sizes and tags matter for the placement
    std::vector<uz_features> views(total);
    std::vector<const uz_features*> feats(total);
    std::vector<int32_t> ns(total), cols(total);
    size_t blob_total = 0;
    static const uint8_t kDummy = 0;
    for (size_t i = 0; i < total; ++i) {
        if ((st = wire_prepare(ctx, blobs[i], blob_bytes[i], &ns[i], &cols[i])) != UZ_OK) return st;
        memset(&views[i], 0, sizeof(uz_features));
        views[i].n = ns[i]; views[i].desc_bytes = cols[i]; views[i].desc_stride = cols[i];
        views[i].feature_type = feature_types[i]; views[i].sensor_frame = sensor_frames[i];
        views[i].descriptors = &kDummy; views[i].positions = (const double*)&kDummy; views[i].valid_3d = &kDummy;
        feats[i] = &views[i];
        blob_total += ((size_t)ns[i] * wire_elem_bytes(cols[i]) + 255) & ~(size_t)255;
    }
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->transient.reset();
    ctx->h_chunks.reset(); ctx->d_chunks.reset();
    std::vector<Cam> up;
    std::vector<BlockRef> blocks;
    if ((st = place_cams(ctx, ctx->store_arena, feats, cams_per_keyframe, (size_t)n_keyframes, up, blocks)) != UZ_OK) return st;
    auto undo = [&]() { cudaStreamSynchronize(ctx->stream); cudaGetLastError(); for (auto& b : blocks) ctx->store_arena.free(b.p, b.bytes); };
    // the serialised bodies travel like any other host buffers (pinned sources pulled directly, pageable ones through the ring)
    uint8_t* d_blobs = (uint8_t*)ctx->transient.alloc(std::max<size_t>(blob_total, 1));
    int* dstat = (int*)ctx->transient.alloc(4);
    if (!d_blobs || !dstat) { undo(); return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed"); }
    std::vector<CopyItem> items;
    std::vector<WireJob> jobs;
    size_t at = 0;
    int max_n = 1;
    for (size_t i = 0; i < total; ++i) {
        if (ns[i] == 0) continue;
        const size_t body = (size_t)ns[i] * wire_elem_bytes(cols[i]);
        items.push_back(CopyItem{blobs[i] + 4, d_blobs + at, body, 0, 0, 0});
        jobs.push_back(WireJob{d_blobs + at, (uint8_t*)up[i].raw, up[i].pos, up[i].valid, ns[i], cols[i]});
        max_n = std::max(max_n, ns[i]);
        at += (body + 255) & ~(size_t)255;
    }
    st = flush_copies(ctx, items);
    if (st == UZ_OK && cudaMemsetAsync(dstat, 0, 4, ctx->stream) != cudaSuccess) st = fail(ctx, UZ_ERR_CUDA, "cudaMemsetAsync failed");
    for (size_t j0 = 0; j0 < jobs.size() && st == UZ_OK; j0 += 32768) {
        const size_t cnt = std::min<size_t>(32768, jobs.size() - j0);
        WireJob* h = (WireJob*)ctx->h_chunks.alloc(cnt * sizeof(WireJob));
        WireJob* d = (WireJob*)ctx->d_chunks.alloc(cnt * sizeof(WireJob));
        if (!h || !d) { st = fail(ctx, UZ_ERR_NOMEM, "wire job table allocation failed"); break; }
        memcpy(h, jobs.data() + j0, cnt * sizeof(WireJob));
        if (cudaMemcpyAsync(d, h, cnt * sizeof(WireJob), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { st = fail(ctx, UZ_ERR_CUDA, "cudaMemcpyAsync failed"); break; }
        wire_decode_bulk_kernel<<<dim3((unsigned)std::min((max_n * 32 + 255) / 256, 32), (unsigned)cnt, 1), 256, 0, ctx->stream>>>(d, dstat);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) st = fail(ctx, UZ_ERR_CUDA, "wire_decode_bulk_kernel launch failed");
    }
    if (st == UZ_OK) st = derive_layouts(ctx, up.data(), up.size());
    int stat = 0;
    if (st == UZ_OK && cudaMemcpyAsync(&stat, dstat, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) st = fail(ctx, UZ_ERR_CUDA, "cudaMemcpyAsync failed");
    if (st == UZ_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, UZ_ERR_CUDA, "bulk wire ingest failed");
    if (st == UZ_OK && stat) st = fail(ctx, UZ_ERR_UNSUPPORTED, "an element's descriptor length differs from the first element's");
    if (st != UZ_OK) { undo(); return st; }
    register_keyframes(ctx, up, blocks, cams_per_keyframe, n_keyframes, handles_out);
    return UZ_OK;
}

// ---- envelope walk (host, no device): where the Feature[] fields sit inside serialised messages -------------------------
namespace {
struct WireCursor {
    const uint8_t* p; size_t n, at; bool ok;
    bool need(size_t k) { if (!ok || k > n - at) { ok = false; return false; } return true; }
    uint32_t u32() { if (!need(4)) return 0; uint32_t v; memcpy(&v, p + at, 4); at += 4; return v; }
    void skip(size_t k) { if (need(k)) at += k; }
    void str(size_t* off, int32_t* len) { const uint32_t l = u32(); if (off) *off = at; if (len) *len = (int32_t)l; skip(l); }
    void header() { skip(4 + 8); str(nullptr, nullptr); }                       // std_msgs/Header: seq, stamp, frame_id
    void arr(size_t elem) { const uint32_t c = u32(); if (ok && (size_t)c > (n - at) / (elem ? elem : 1)) { ok = false; return; } skip((size_t)c * elem); }
    void image() { header(); skip(4 + 4); str(nullptr, nullptr); skip(1 + 4); arr(1); }   // sensor_msgs/Image
};

// one graph_slam_msgs/SensorData (SensorData.msg) starting at the cursor
bool walk_sensor_data(WireCursor& c, uz_wire_sensor* out) {
    uz_wire_sensor s;
    memset(&s, 0, sizeof(s));
    c.header();
    s.sensor_type = (int32_t)c.u32();
    s.displacement_offset = c.at; c.skip(56);                                   // geometry_msgs/Pose: 7 float64
    c.str(&s.sensor_frame_offset, &s.sensor_frame_len);
    // graph_slam_msgs/Features: header, descriptor_type, Feature[] features, sensor_msgs/CameraInfo camera_model
    c.header();
    s.descriptor_type = (int32_t)c.u32();
    s.features_offset = c.at;
    const uint32_t nf = c.u32();
    s.n_features = (int32_t)nf;
    for (uint32_t i = 0; i < nf && c.ok; ++i) { c.skip(4 + 4 + 1 + 4); c.arr(4); c.skip(24); }     // Feature.msg
    s.features_bytes = c.at - s.features_offset;
    c.header(); c.skip(4 + 4); c.str(nullptr, nullptr); c.arr(8); c.skip(8 * (9 + 9 + 12)); c.skip(4 + 4); c.skip(4 * 4 + 1);   // CameraInfo
    c.image(); c.image();                                                        // graph_slam_msgs/DepthImage: depth, color
    c.arr(4);                                                                    // float32[] gist_descriptor
    c.header(); c.skip(7 * 4); c.arr(4); c.arr(4);                               // sensor_msgs/LaserScan
    c.skip(24);                                                                  // geometry_msgs/Point scan_center
    if (c.ok && out) *out = s;
    return c.ok;
}
}  // namespace

uz_status uz_wire_walk_sensor_data(const uint8_t* msg, size_t bytes, uz_wire_sensor* sensor_out, size_t* consumed_out) {
    if (!msg || !sensor_out) return UZ_ERR_INVALID;
    WireCursor c{msg, bytes, 0, true};
    if (!walk_sensor_data(c, sensor_out)) return UZ_ERR_INVALID;
    if (consumed_out) *consumed_out = c.at;
    return UZ_OK;
}

uz_status uz_wire_walk_node(const uint8_t* msg, size_t bytes, uz_wire_sensor* sensors_out, int32_t capacity, int32_t* n_sensors_out,
                            size_t* id_offset_out, int32_t* id_len_out) {
    if (!msg || !n_sensors_out || capacity < 0 || (capacity > 0 && !sensors_out)) return UZ_ERR_INVALID;
    WireCursor c{msg, bytes, 0, true};
    c.arr(8);                                                                    // time[] stamps
    c.str(id_offset_out, id_len_out);                                            // string id
    c.skip(56 + 56);                                                             // pose, odom_pose
    c.header();                                                                  // SensorDataArray.header
    const uint32_t ns = c.u32();
    int32_t n = 0;
    for (uint32_t i = 0; i < ns && c.ok; ++i) {
        uz_wire_sensor s;
        const size_t base = c.at;
        (void)base;
        if (!walk_sensor_data(c, &s)) break;
        if (n < capacity) sensors_out[n] = s;
        ++n;
    }
    c.arr(0);                                                                    // string[] edge_ids: count, then the strings
    if (c.ok) {
        uint32_t cnt; memcpy(&cnt, msg + c.at - 4, 4);
        for (uint32_t i = 0; i < cnt && c.ok; ++i) c.str(nullptr, nullptr);
    }
    c.skip(1 + 8);                                                               // bool fixed, float64 uncertainty
    if (!c.ok) return UZ_ERR_INVALID;
    *n_sensors_out = n;
    return UZ_OK;
}

uz_status uz_wire_encode(uz_context* ctx, int32_t handle, int32_t cam, const int32_t* uv, uint8_t* blob_out, size_t capacity,
                         size_t* bytes_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (handle < 0 || handle >= (int32_t)ctx->kfs.size() || !ctx->kfs[handle].live) return fail(ctx, UZ_ERR_INVALID, "unknown keyframe handle");
    if (cam < 0 || cam >= (int32_t)ctx->kfs[handle].cams.size()) return fail(ctx, UZ_ERR_INVALID, "camera index out of range");
    if (!blob_out || !bytes_out) return fail(ctx, UZ_ERR_INVALID, "null output");
    const Cam& c = ctx->kfs[handle].cams[cam];
    const size_t need = 4 + (size_t)c.n * wire_elem_bytes(c.dbytes);
    if (need > capacity) return fail(ctx, UZ_ERR_INVALID, "output capacity too small");
    const uint32_t cnt = (uint32_t)c.n;
    memcpy(blob_out, &cnt, 4);
    *bytes_out = need;
    if (c.n == 0) return UZ_OK;
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->transient.reset();
    uint8_t* db = (uint8_t*)ctx->transient.alloc(need - 4);
    int32_t* duv = uv ? (int32_t*)ctx->transient.alloc((size_t)c.n * 8) : nullptr;
    if (!db || (uv && !duv)) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
    if (uv) UZ_CUDA(ctx, cudaMemcpyAsync(duv, uv, (size_t)c.n * 8, cudaMemcpyHostToDevice, ctx->stream));
    wire_encode_kernel<<<(c.n * 32 + 255) / 256, 256, 0, ctx->stream>>>((const uint8_t*)c.raw, c.n, c.dbytes, c.pos, c.valid, duv, db);
    ctx->launches++;
    UZ_CUDA(ctx, cudaGetLastError());
    UZ_CUDA(ctx, cudaMemcpyAsync(blob_out + 4, db, need - 4, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UZ_OK;
}

// read a stored camera back (parity tap for the ingestion paths; also what a toMsg adapter would serialise)
uz_status uz_store_read(uz_context* ctx, int32_t handle, int32_t cam, int32_t capacity, int32_t* n_out,
                        int32_t* desc_bytes_out, uint8_t* descriptors_out, double* positions_out, uint8_t* valid_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (handle < 0 || handle >= (int32_t)ctx->kfs.size() || !ctx->kfs[handle].live) return fail(ctx, UZ_ERR_INVALID, "unknown keyframe handle");
    if (cam < 0 || cam >= (int32_t)ctx->kfs[handle].cams.size()) return fail(ctx, UZ_ERR_INVALID, "camera index out of range");
    const Cam& c = ctx->kfs[handle].cams[cam];
    if (n_out) *n_out = c.n;
    if (desc_bytes_out) *desc_bytes_out = c.dbytes;
    if (c.n == 0) return UZ_OK;
    if (c.n > capacity) return fail(ctx, UZ_ERR_INVALID, "output capacity too small");
    if (descriptors_out) UZ_CUDA(ctx, cudaMemcpyAsync(descriptors_out, c.raw, (size_t)c.n * c.dbytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (positions_out) UZ_CUDA(ctx, cudaMemcpyAsync(positions_out, c.pos, (size_t)c.n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (valid_out) UZ_CUDA(ctx, cudaMemcpyAsync(valid_out, c.valid, (size_t)c.n, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UZ_OK;
}

}  // extern "C"
