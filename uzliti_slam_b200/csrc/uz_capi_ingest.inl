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
    const CamLayout L = cam_layout(0, n, dbytes);
    uint8_t* base = (uint8_t*)arena.alloc(std::max<size_t>(L.end, 1));
    if (!base) return fail(ctx, UZ_ERR_NOMEM, "device arena allocation failed");
    block = BlockRef{base, std::max<size_t>(L.end, 1)};
    if (n == 0) return UZ_OK;
    c.raw = (uint32_t*)(base + L.raw); c.pos = (double*)(base + L.pos); c.valid = base + L.valid;
    c.csa = (uint32_t*)(base + L.csa);
    c.e8 = dbytes == UZ_DESC_BYTES ? base + L.e8 : nullptr;
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
