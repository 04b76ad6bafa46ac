// uz_capi_places.inl — host side of the device place recogniser (included at the end of uz_capi.cu; kernels in
// uz_places.cuh).  Mirrors PlaceRecognizer's public calls (/root/reference/place_recognition/include/
// place_recognition/place_recognizer.h:37-72) for the LshSetRecognizer back end, batched: a call handles n keyframes
// with the results the reference would produce by handling them one by one, in order.

namespace {

enum PlaceMode { kSearchAndAdd = 0, kAddOnly = 1, kSearchOnly = 2 };

uz_status places_grow_places(uz_context* ctx, size_t need) {
    PlacesState& ps = ctx->places;
    if (need <= ps.place_cap) return UZ_OK;
    size_t cap = std::max<size_t>(1024, ps.place_cap);
    while (cap < need) cap *= 2;
    long long* ns = nullptr; uint8_t* nl = nullptr;
    UZ_CUDA(ctx, cudaMalloc(&ns, cap * sizeof(long long)));
    UZ_CUDA(ctx, cudaMalloc(&nl, cap));
    UZ_CUDA(ctx, cudaMemsetAsync(nl, 0, cap, ctx->stream));
    if (ps.place_cap) {
        UZ_CUDA(ctx, cudaMemcpyAsync(ns, ps.d_stamps, ps.places.size() * sizeof(long long), cudaMemcpyDeviceToDevice, ctx->stream));
        UZ_CUDA(ctx, cudaMemcpyAsync(nl, ps.d_live, ps.places.size(), cudaMemcpyDeviceToDevice, ctx->stream));
        UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ps.d_stamps); cudaFree(ps.d_live);
    }
    ps.d_stamps = ns; ps.d_live = nl; ps.place_cap = cap;
    return UZ_OK;
}

uz_status places_grow_nodes(uz_context* ctx, size_t need) {
    PlacesState& ps = ctx->places;
    if (need <= ps.node_cap) return UZ_OK;
    if (need >= 0xFFFFFFF0ull) return fail(ctx, UZ_ERR_NOMEM, "place recogniser: more than 2^32 bucket entries");
    size_t cap = std::max<size_t>((size_t)1 << 20, ps.node_cap);
    while (cap < need) cap *= 2;
    cap = std::min<size_t>(cap, 0xFFFFFFF0ull);
    PlaceNode* nn = nullptr;
    UZ_CUDA(ctx, cudaMalloc(&nn, cap * sizeof(PlaceNode)));
    if (ps.n_nodes) {
        UZ_CUDA(ctx, cudaMemcpyAsync(nn, ps.d_nodes, ps.n_nodes * sizeof(PlaceNode), cudaMemcpyDeviceToDevice, ctx->stream));
        UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (ps.d_nodes) cudaFree(ps.d_nodes);
    ps.d_nodes = nn; ps.node_cap = cap;
    return UZ_OK;
}

dim3 places_grid(uz_context* ctx, const std::vector<PlaceCam>& cams, size_t first, size_t count) {
    int max_n = 1;
    for (size_t i = first; i < first + count; ++i) max_n = std::max(max_n, cams[i].n);
    return dim3((unsigned)std::min((max_n * 8 + 255) / 256, 64), (unsigned)count, 1);
}

// keeps the load factor of the slot array below 1/2 (distinct keys <= entries); growth re-links every node
uz_status places_grow_slots(uz_context* ctx, size_t entries_after) {
    PlacesState& ps = ctx->places;
    if (ps.n_slots && entries_after * 2 <= ps.n_slots) return UZ_OK;
    size_t want = std::max<size_t>((size_t)1 << 20, ps.n_slots);
    while (want < entries_after * 4) want *= 2;           // grow to load <= 1/4 so growth is rare
    if (want > ((size_t)1 << 31)) want = (size_t)1 << 31;
    if (entries_after * 10 > want * 9) return fail(ctx, UZ_ERR_NOMEM, "place recogniser: slot table full");
    PlaceSlot* ns = nullptr;
    UZ_CUDA(ctx, cudaMalloc(&ns, want * sizeof(PlaceSlot)));
    UZ_CUDA(ctx, cudaMemsetAsync(ns, 0, want * sizeof(PlaceSlot), ctx->stream));
    if (ps.d_slots) { UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ps.d_slots); }
    ps.d_slots = ns; ps.n_slots = (uint32_t)want;
    // re-link what is already there
    const size_t n = ps.inserted.size();
    for (size_t c0 = 0; c0 < n; c0 += 32768) {
        const size_t cnt = std::min<size_t>(32768, n - c0);
        UZ_CUDA(ctx, ps.d_cams.ensure(cnt * sizeof(PlaceCam)));
        UZ_CUDA(ctx, cudaMemcpyAsync(ps.d_cams.p, ps.inserted.data() + c0, cnt * sizeof(PlaceCam), cudaMemcpyHostToDevice, ctx->stream));
        places_relink_kernel<<<places_grid(ctx, ps.inserted, c0, cnt), 256, 0, ctx->stream>>>(
            (const PlaceCam*)ps.d_cams.p, ps.d_slots, ps.n_slots - 1, ps.d_nodes);
        ctx->launches++;
        UZ_CUDA(ctx, cudaGetLastError());
        UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // d_cams is reused by the next slice
    }
    return UZ_OK;
}

void places_release(uz_context* ctx) {
    PlacesState& ps = ctx->places;
    if (ps.d_slots) cudaFree(ps.d_slots);
    if (ps.d_nodes) cudaFree(ps.d_nodes);
    if (ps.d_stamps) cudaFree(ps.d_stamps);
    if (ps.d_live) cudaFree(ps.d_live);
    ps.d_slots = nullptr; ps.d_nodes = nullptr; ps.d_stamps = nullptr; ps.d_live = nullptr;
    ps.n_slots = 0; ps.node_cap = 0; ps.n_nodes = 0; ps.place_cap = 0; ps.live_entries = 0;
    ps.d_cams.release(); ps.d_votes.release(); ps.d_out.release(); ps.d_out_votes.release();
    ps.places.clear(); ps.by_handle.clear(); ps.checked.clear(); ps.inserted.clear(); ps.retired.clear();
}

uz_status places_reset(uz_context* ctx) {
    PlacesState& ps = ctx->places;
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ps.d_slots) UZ_CUDA(ctx, cudaMemsetAsync(ps.d_slots, 0, (size_t)ps.n_slots * sizeof(PlaceSlot), ctx->stream));
    if (ps.d_live) UZ_CUDA(ctx, cudaMemsetAsync(ps.d_live, 0, ps.place_cap, ctx->stream));
    ps.n_nodes = 0; ps.live_entries = 0;
    for (auto& kv : ps.retired) for (auto& b : kv.second) ctx->store_arena.free(b.p, b.bytes);
    ps.places.clear(); ps.by_handle.clear(); ps.checked.clear(); ps.inserted.clear(); ps.retired.clear();
    return UZ_OK;
}

// The keyframe behind `handle` leaves the store (uz_store_remove) and the handle will be recycled: its place dies
// (removePlace), the checked_ pairs naming it go (the next keyframe with this handle is a different node), and its rows are
// never re-linked again (their memory is about to be reused).
void places_forget_handle(uz_context* ctx, int32_t handle) {
    PlacesState& ps = ctx->places;
    auto it = ps.by_handle.find(handle);
    if (it != ps.by_handle.end()) {
        const int32_t place = it->second;
        PlaceInfo& pi = ps.places[place];
        pi.live = false;
        for (uint32_t k = 0; k < pi.ins_count; ++k) ps.inserted[pi.ins_begin + k].n = 0;
        auto rt = ps.retired.find(place);
        if (rt != ps.retired.end()) {
            for (auto& b : rt->second) ctx->store_arena.free(b.p, b.bytes);
            ps.retired.erase(rt);
        }
        ps.by_handle.erase(it);
        if (ps.d_live) cudaMemsetAsync(ps.d_live + place, 0, 1, ctx->stream);
    }
    for (auto c = ps.checked.begin(); c != ps.checked.end();) {
        if ((int32_t)(*c >> 32) == handle || (int32_t)(*c & 0xFFFFFFFFu) == handle) c = ps.checked.erase(c);
        else ++c;
    }
}

bool place_cam_usable(const Cam& c) { return c.n > 0 && c.raw != nullptr && is_binary_type(c.feature_type); }

// The one driver behind search_and_add / add / search.
uz_status places_run(uz_context* ctx, PlaceMode mode, const int32_t* handles, const int64_t* stamps_ns, int32_t n,
                     int32_t* pairs_out, int32_t capacity, int32_t* n_pairs_out) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    if (n_pairs_out) *n_pairs_out = 0;
    if (n < 0 || (n > 0 && (!handles || !stamps_ns))) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    if (mode != kAddOnly && (capacity < 0 || (capacity > 0 && !pairs_out) || !n_pairs_out)) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    if (n == 0) return UZ_OK;
    PlacesState& ps = ctx->places;
    const uz_place_params prm = ps.params;
    const int32_t nk = (int32_t)ctx->kfs.size();
    for (int i = 0; i < n; ++i)
        if (handles[i] < 0 || handles[i] >= nk || !ctx->kfs[handles[i]].live) return fail(ctx, UZ_ERR_INVALID, "unknown keyframe handle");
    if (mode == kSearchOnly && ps.by_handle.empty()) return UZ_OK;       // place_recognizer.cpp:157-160

    // 1. place indices, cameras to insert, cameras to query
    std::vector<int32_t> place_of((size_t)n, -1);         // -1: "tried to add existing place" (no place, no result)
    std::vector<PlaceCam> ins, qry;
    std::vector<int32_t> qry_owner;                       // index i of the keyframe a query camera belongs to
    size_t new_nodes = 0;
    const size_t places_before = ps.places.size();
    for (int i = 0; i < n; ++i) {
        const int32_t h = handles[i];
        int32_t place = -1;
        if (mode != kSearchOnly) {
            if (ps.by_handle.count(h)) continue;          // place_recognizer.cpp:80-84 / :142-144
            place = (int32_t)ps.places.size();
            ps.places.push_back(PlaceInfo{h, (long long)stamps_ns[i], true, (uint32_t)(ps.inserted.size() + ins.size()), 0u});
            ps.by_handle[h] = place;
            place_of[i] = place;
        }
        for (const Cam& c : ctx->kfs[h].cams) {
            if (!place_cam_usable(c)) continue;
            PlaceCam pc;
            memset(&pc, 0, sizeof(pc));
            pc.raw = c.raw; pc.n = c.n; pc.place = place; pc.stamp_ns = (long long)stamps_ns[i];
            pc.wide_shift = c.dbytes == UZ_DESC_BYTES ? 0 : 1;
            const bool big = c.n > prm.min_rows;          // lsh_set_recognizer.cpp:66 / :108
            if (mode != kSearchOnly && big) {
                pc.node_base = (uint32_t)(ps.n_nodes + new_nodes);
                pc.insert_filtered = mode == kSearchAndAdd;
                new_nodes += (size_t)c.n * 8;
                ins.push_back(pc);
                ps.places[place].ins_count++;
            }
            if (mode != kAddOnly) {
                pc.query_filtered = (mode == kSearchAndAdd && big) ? 1 : 0;
                pc.place_limit = mode == kSearchAndAdd ? place : (int32_t)places_before;
                qry.push_back(pc);
                qry_owner.push_back(i);
            }
        }
    }
    const int32_t n_places = (int32_t)ps.places.size();

    // 2. capacity + place table
    if ((st = places_grow_places(ctx, (size_t)n_places)) != UZ_OK) return st;
    if ((st = places_grow_nodes(ctx, ps.n_nodes + new_nodes)) != UZ_OK) return st;
    if ((st = places_grow_slots(ctx, ps.live_entries + new_nodes)) != UZ_OK) return st;
    if (n_places > (int32_t)places_before) {
        std::vector<long long> hs((size_t)n_places - places_before);
        for (size_t p = places_before; p < (size_t)n_places; ++p) hs[p - places_before] = ps.places[p].stamp_ns;
        UZ_CUDA(ctx, cudaMemcpyAsync(ps.d_stamps + places_before, hs.data(), hs.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
        UZ_CUDA(ctx, cudaMemsetAsync(ps.d_live + places_before, 1, hs.size(), ctx->stream));
        UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));        // hs is pageable and dies here
    }

    cudaEvent_t e0 = ctx->get_event(), e1 = ctx->get_event(), e2 = ctx->get_event(), e3 = ctx->get_event();
    float ins_ms = 0, vote_ms = 0, sel_ms = 0;

    // 3. insert every new camera (one launch per 32768 cameras: gridDim.y)
    cudaEventRecord(e0, ctx->stream);
    for (size_t c0 = 0; c0 < ins.size(); c0 += 32768) {
        const size_t cnt = std::min<size_t>(32768, ins.size() - c0);
        UZ_CUDA(ctx, ps.d_cams.ensure(cnt * sizeof(PlaceCam)));
        UZ_CUDA(ctx, cudaMemcpyAsync(ps.d_cams.p, ins.data() + c0, cnt * sizeof(PlaceCam), cudaMemcpyHostToDevice, ctx->stream));
        places_insert_kernel<<<places_grid(ctx, ins, c0, cnt), 256, 0, ctx->stream>>>(
            (const PlaceCam*)ps.d_cams.p, ps.d_slots, ps.n_slots - 1, ps.d_nodes, prm.min_key_bits);
        ctx->launches++;
        UZ_CUDA(ctx, cudaGetLastError());
        UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    cudaEventRecord(e1, ctx->stream);
    ps.n_nodes += new_nodes; ps.live_entries += new_nodes;
    ps.inserted.insert(ps.inserted.end(), ins.begin(), ins.end());
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!ins.empty()) cudaEventElapsedTime(&ins_ms, e0, e1);
    ctx->places_ms[0] = ins_ms; ctx->places_ms[1] = 0; ctx->places_ms[2] = 0;
    if (mode == kAddOnly) {
        ctx->event_pool.push_back(e0); ctx->event_pool.push_back(e1); ctx->event_pool.push_back(e2); ctx->event_pool.push_back(e3);
        return UZ_OK;
    }

    // 4. votes + selection, in slices of query cameras whose dense vote rows fit the budget
    const int k = std::max(prm.k_nearest_neighbors, 0);
    const size_t row_bytes = (size_t)std::max(n_places, 1) * sizeof(uint32_t);
    const size_t budget = (size_t)1 << 30;
    const size_t rows_per_slice = std::max<size_t>(1, std::min<size_t>(32768, budget / row_bytes));
    std::vector<int32_t> sel((size_t)qry.size() * std::max(k, 1), -1);
    PlaceSelectParams sp;
    const double need = 8.0 * prm.T;                      // votes / 8 >= T  <=>  votes >= 8 T (exact: /8 is a power of two)
    sp.min_votes = need <= 1.0 ? 1u : (need >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)std::ceil(need));
    sp.min_gap_ns = prm.min_gap_ns; sp.k = k; sp.n_places = n_places;
    for (size_t c0 = 0; c0 < qry.size() && k > 0; c0 += rows_per_slice) {
        const size_t cnt = std::min(rows_per_slice, qry.size() - c0);
        for (size_t r = 0; r < cnt; ++r) qry[c0 + r].row = (int32_t)r;
        UZ_CUDA(ctx, ps.d_cams.ensure(cnt * sizeof(PlaceCam)));
        UZ_CUDA(ctx, ps.d_votes.ensure(cnt * row_bytes));
        UZ_CUDA(ctx, ps.d_out.ensure(cnt * k * sizeof(int32_t)));
        UZ_CUDA(ctx, ps.d_out_votes.ensure(cnt * k * sizeof(uint32_t)));
        UZ_CUDA(ctx, cudaMemcpyAsync(ps.d_cams.p, qry.data() + c0, cnt * sizeof(PlaceCam), cudaMemcpyHostToDevice, ctx->stream));
        cudaEventRecord(e1, ctx->stream);
        UZ_CUDA(ctx, cudaMemsetAsync(ps.d_votes.p, 0, cnt * row_bytes, ctx->stream));
        places_vote_kernel<<<places_grid(ctx, qry, c0, cnt), 256, 0, ctx->stream>>>(
            (const PlaceCam*)ps.d_cams.p, ps.d_slots, ps.n_slots - 1, ps.d_nodes, (uint32_t*)ps.d_votes.p, n_places, prm.min_key_bits);
        cudaEventRecord(e2, ctx->stream);
        places_select_kernel<<<(unsigned)cnt, 256, 0, ctx->stream>>>((const PlaceCam*)ps.d_cams.p, (const uint32_t*)ps.d_votes.p,
                                                                      ps.d_stamps, ps.d_live, sp, (int32_t*)ps.d_out.p, (uint32_t*)ps.d_out_votes.p);
        cudaEventRecord(e3, ctx->stream);
        ctx->launches += 2;
        UZ_CUDA(ctx, cudaGetLastError());
        UZ_CUDA(ctx, cudaMemcpyAsync(sel.data() + c0 * k, ps.d_out.p, cnt * k * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, e1, e2); cudaEventElapsedTime(&b, e2, e3);
        vote_ms += a; sel_ms += b;
        ps.last_votes_bytes = (int64_t)(cnt * row_bytes);
    }
    ctx->places_ms[1] = vote_ms; ctx->places_ms[2] = sel_ms;
    ctx->event_pool.push_back(e0); ctx->event_pool.push_back(e1); ctx->event_pool.push_back(e2); ctx->event_pool.push_back(e3);

    // 5. per keyframe, in order: concatenate its cameras' ranked lists, first k, checked_ (place_recognizer.cpp:91-116)
    int32_t n_out = 0;
    size_t q = 0;
    for (int i = 0; i < n; ++i) {
        int taken = 0;
        bool full = false;
        std::vector<int32_t> mapped;
        while (q < qry.size() && qry_owner[q] == i) {
            for (int r = 0; r < k && !full; ++r) {
                const int32_t p = sel[q * k + r];
                if (p < 0) break;
                mapped.push_back(ps.places[p].handle);
                if (++taken >= k) full = true;
            }
            ++q;
        }
        for (int32_t from : mapped) {
            const uint64_t key = ((uint64_t)(uint32_t)from << 32) | (uint32_t)handles[i];
            if (ps.checked.insert(key).second) {
                if (n_out < capacity) { pairs_out[2 * n_out] = from; pairs_out[2 * n_out + 1] = handles[i]; }
                ++n_out;
            }
        }
    }
    *n_pairs_out = n_out;
    return UZ_OK;
}

}  // namespace

extern "C" {

void uz_default_place_params(uz_place_params* p) {
    if (!p) return;
    p->T = 2.0;                        // iti_slam_launch/yaml/slam.yaml:48
    p->k_nearest_neighbors = 20;       // slam.yaml:46
    p->min_rows = 150;                 // lsh_set_recognizer.cpp:66
    p->min_key_bits = 12;              // 3 * key_width (lsh_set_recognizer.cpp:239), key_width = 4 (:35)
    p->min_gap_ns = 5000000000LL;      // place_recognizer.cpp:94
}

uz_status uz_places_set_params(uz_context* ctx, const uz_place_params* p) {
    if (!ctx || !p) return UZ_ERR_INVALID;
    if (p->k_nearest_neighbors < 0 || p->k_nearest_neighbors > 4096 || p->min_rows < 0 || p->min_key_bits < 0 || p->min_key_bits > 32 ||
        !(p->T >= 0.0) || p->min_gap_ns < 0)
        return fail(ctx, UZ_ERR_INVALID, "place parameters out of range");
    ctx->places.params = *p;
    return UZ_OK;
}

uz_status uz_places_clear(uz_context* ctx) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    return places_reset(ctx);
}

uz_status uz_places_search_and_add(uz_context* ctx, const int32_t* handles, const int64_t* stamps_ns, int32_t n,
                                   int32_t* pairs_out, int32_t capacity, int32_t* n_pairs_out) {
    return places_run(ctx, kSearchAndAdd, handles, stamps_ns, n, pairs_out, capacity, n_pairs_out);
}

uz_status uz_places_add(uz_context* ctx, const int32_t* handles, const int64_t* stamps_ns, int32_t n) {
    return places_run(ctx, kAddOnly, handles, stamps_ns, n, nullptr, 0, nullptr);
}

uz_status uz_places_search(uz_context* ctx, const int32_t* handles, const int64_t* stamps_ns, int32_t n,
                           int32_t* pairs_out, int32_t capacity, int32_t* n_pairs_out) {
    return places_run(ctx, kSearchOnly, handles, stamps_ns, n, pairs_out, capacity, n_pairs_out);
}

uz_status uz_places_remove(uz_context* ctx, int32_t handle) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    PlacesState& ps = ctx->places;
    auto it = ps.by_handle.find(handle);
    if (it == ps.by_handle.end()) return fail(ctx, UZ_ERR_INVALID, "tried to remove a non-existing place");
    const int32_t place = it->second;
    ps.places[place].live = false;
    ps.by_handle.erase(it);
    UZ_CUDA(ctx, cudaMemsetAsync(ps.d_live + place, 0, 1, ctx->stream));
    return UZ_OK;
}

int32_t uz_places_count(const uz_context* ctx) { return ctx ? (int32_t)ctx->places.places.size() : 0; }

uz_status uz_places_votes(uz_context* ctx, int32_t handle, int32_t cam, int32_t filtered, int32_t* votes_out, int32_t capacity) {
    uz_status st = check_ctx(ctx);
    if (st != UZ_OK) return st;
    PlacesState& ps = ctx->places;
    if (handle < 0 || handle >= (int32_t)ctx->kfs.size() || !ctx->kfs[handle].live) return fail(ctx, UZ_ERR_INVALID, "unknown keyframe handle");
    if (cam < 0 || cam >= (int32_t)ctx->kfs[handle].cams.size() || !votes_out) return fail(ctx, UZ_ERR_INVALID, "bad arguments");
    const int32_t n_places = (int32_t)ps.places.size();
    const int32_t nout = std::min(capacity, n_places);
    if (nout <= 0) return UZ_OK;
    const Cam& c = ctx->kfs[handle].cams[cam];
    if (!place_cam_usable(c) || !ps.d_slots) { memset(votes_out, 0, (size_t)nout * 4); return UZ_OK; }
    PlaceCam pc;
    memset(&pc, 0, sizeof(pc));
    pc.raw = c.raw; pc.n = c.n; pc.place = -1; pc.query_filtered = filtered ? 1 : 0; pc.place_limit = n_places; pc.row = 0;
    pc.wide_shift = c.dbytes == UZ_DESC_BYTES ? 0 : 1;
    UZ_CUDA(ctx, ps.d_cams.ensure(sizeof(PlaceCam)));
    UZ_CUDA(ctx, ps.d_votes.ensure((size_t)n_places * 4));
    UZ_CUDA(ctx, cudaMemcpyAsync(ps.d_cams.p, &pc, sizeof(pc), cudaMemcpyHostToDevice, ctx->stream));
    UZ_CUDA(ctx, cudaMemsetAsync(ps.d_votes.p, 0, (size_t)n_places * 4, ctx->stream));
    std::vector<PlaceCam> one(1, pc);
    places_vote_kernel<<<places_grid(ctx, one, 0, 1), 256, 0, ctx->stream>>>((const PlaceCam*)ps.d_cams.p, ps.d_slots, ps.n_slots - 1,
                                                                              ps.d_nodes, (uint32_t*)ps.d_votes.p, n_places, ps.params.min_key_bits);
    ctx->launches++;
    UZ_CUDA(ctx, cudaGetLastError());
    UZ_CUDA(ctx, cudaMemcpyAsync(votes_out, ps.d_votes.p, (size_t)nout * 4, cudaMemcpyDeviceToHost, ctx->stream));
    UZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return UZ_OK;
}

uz_status uz_places_last_timing(uz_context* ctx, double* insert_ms, double* vote_ms, double* select_ms) {
    if (!ctx) return UZ_ERR_INVALID;
    if (insert_ms) *insert_ms = ctx->places_ms[0];
    if (vote_ms) *vote_ms = ctx->places_ms[1];
    if (select_ms) *select_ms = ctx->places_ms[2];
    return UZ_OK;
}

}  // extern "C"
