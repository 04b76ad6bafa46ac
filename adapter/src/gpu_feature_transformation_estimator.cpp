// gpu_feature_transformation_estimator.cpp — see the header.  Follows
// transformation_estimation/src/transformation_estimator.cpp:22-62 and
// transformation_estimation/src/feature_transformation_estimator.cpp:32-171,178-184,337-353.
#include <transformation_estimation/gpu_feature_transformation_estimator.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <unordered_set>

// ---------------------------------------------------------------------------------------------------------
// TransformationEstimator: queue + worker thread (transformation_estimator.cpp:22-62)
// ---------------------------------------------------------------------------------------------------------
TransformationEstimator::TransformationEstimator(boost::function<void(SlamEdge)> callback) : callback_(callback) {
    if (const char* c = std::getenv("UZ_ADAPTER_CHUNK")) { if (std::atoi(c) >= 64) delivery_chunk_ = std::atoi(c); }
}

TransformationEstimator::~TransformationEstimator() { stopThread(); }

void TransformationEstimator::startThread() {
    running_ = true;
    estimation_thread_ = std::thread(&TransformationEstimator::estimationThread, this);
    delivery_thread_ = std::thread(&TransformationEstimator::deliveryThread, this);
}

void TransformationEstimator::stopThread() {
    {
        std::lock_guard<std::mutex> lk(estimation_mutex_);
        if (!running_) return;
        running_ = false;
    }
    cv_.notify_all();
    if (estimation_thread_.joinable()) estimation_thread_.join();
    if (delivery_thread_.joinable()) delivery_thread_.join();
}

void TransformationEstimator::estimateEdge(SlamNode& from, SlamNode& to) {
    bool wake;
    {   // cheap and non-blocking: callers hold graph_mutex_ (graph_slam_node.cpp:207,448)
        std::lock_guard<std::mutex> lk(estimation_mutex_);
        wake = est_queue_.empty();                        // the worker only sleeps on an empty queue: one wake-up per burst
        est_queue_.emplace_back(from, to);                // copies the nodes; FeatureData is shared (shared_ptr)
    }
    if (wake) cv_.notify_all();
}

void TransformationEstimator::estimateEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >& pairs,
                                                std::vector<SlamEdge>& edges, std::vector<char>& ok) {
    edges.assign(pairs.size(), SlamEdge());
    ok.assign(pairs.size(), 0);
    for (size_t i = 0; i < pairs.size(); ++i) ok[i] = estimateEdgeImpl(pairs[i].first, pairs[i].second, edges[i]);
}

// Two threads: the estimation thread turns queued pairs into edges, the delivery thread fires the callbacks.  A drained queue
// is worked off in chunks, and while chunk c's edges are being delivered (edge copies, strings, the caller's callback) chunk
// c + 1 is already being estimated.  Order of delivery is the order of estimation (LIFO per drained queue, :49-50).
// An estimator may split a chunk's work in three (prepare / submit / finish): chunk c + 1 is PREPARED (host-side look-ups)
// while chunk c is still running on the device, then c is finished and c + 1 submitted.  The defaults make that the plain
// synchronous estimateEdgeBatch.
void TransformationEstimator::prepareEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >&, int) {}
void TransformationEstimator::submitEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >& pairs, std::vector<SlamEdge>& edges,
                                              std::vector<char>& ok, int) { estimateEdgeBatch(pairs, edges, ok); }
void TransformationEstimator::finishEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >&, std::vector<SlamEdge>&, std::vector<char>&, int) {}
void TransformationEstimator::endOfBurst() {}

void TransformationEstimator::estimationThread() {
    std::unique_lock<std::mutex> lk(estimation_mutex_);
    // the queue and the batch being worked on are two vectors that change roles, and the chunk / edge vectors live across
    // batches: their capacity is kept, so a steady stream of estimateEdge calls neither regrows the queue nor touches fresh pages
    std::vector<std::pair<SlamNode, SlamNode> > batch, chunk[2];
    int fill = 0;
    while (running_) {
        if (est_queue_.empty()) { cv_.wait(lk); continue; }
        batch.swap(est_queue_);                          // est_queue_ gets the (cleared) vector of the previous batch
        busy_ = true;
        lk.unlock();
        // the reference pops the NEWEST pair first (LIFO, :49-50): deliver in that order
        std::reverse(batch.begin(), batch.end());
        int inflight = -1;
        auto finish = [&](int slot) {
            finishEdgeBatch(chunk[slot], slots_[slot].edges, slots_[slot].ok, slot);
            lk.lock();
            slot_full_[slot] = true;
            lk.unlock();
            cv_.notify_all();
        };
        for (size_t at = 0; at < batch.size(); at += (size_t)delivery_chunk_) {
            const size_t end = std::min(batch.size(), at + (size_t)delivery_chunk_);
            chunk[fill].assign(std::make_move_iterator(batch.begin() + at), std::make_move_iterator(batch.begin() + end));
            prepareEdgeBatch(chunk[fill], fill);         // beside the device work of the chunk in flight
            if (inflight >= 0) finish(inflight);
            // wait for the delivery slot with nothing in flight and no estimator lock held: a callback that calls back into
            // the estimator (forgetNode, estimateEdgeDirect) must not find the worker sitting on gpuMutex()
            lk.lock();
            cv_.wait(lk, [&] { return !slot_full_[fill]; });
            lk.unlock();
            submitEdgeBatch(chunk[fill], slots_[fill].edges, slots_[fill].ok, fill);
            inflight = fill;
            fill ^= 1;
        }
        if (inflight >= 0) finish(inflight);
        endOfBurst();
        batch.clear();
        chunk[0].clear(); chunk[1].clear();
        lk.lock();
        busy_ = false;
        cv_.notify_all();
    }
}

void TransformationEstimator::deliveryThread() {
    std::unique_lock<std::mutex> lk(estimation_mutex_);
    int cur = 0;
    for (;;) {
        cv_.wait(lk, [&] { return slot_full_[cur] || (!running_ && !busy_); });
        if (!slot_full_[cur]) return;                    // stopped and nothing left to deliver
        lk.unlock();
        std::vector<SlamEdge>& edges = slots_[cur].edges;
        const std::vector<char>& ok = slots_[cur].ok;
        for (size_t i = 0; i < edges.size(); ++i) {
            if (!ok[i]) edges[i].matching_score_ = 0.;   // :53-55
            callback_(edges[i]);                         // :56 - fired even on failure
        }
        lk.lock();
        slot_full_[cur] = false;
        cv_.notify_all();
        cur ^= 1;
    }
}

void TransformationEstimator::waitIdle() {
    std::unique_lock<std::mutex> lk(estimation_mutex_);
    cv_.wait(lk, [this] { return est_queue_.empty() && !busy_ && !slot_full_[0] && !slot_full_[1]; });
}

// ---------------------------------------------------------------------------------------------------------
// GpuFeatureTransformationEstimator
// ---------------------------------------------------------------------------------------------------------
GpuFeatureTransformationEstimator::GpuFeatureTransformationEstimator(boost::function<void(SlamEdge)> callback, int device)
    : TransformationEstimator(callback) {
    init(std::vector<int>(1, device));
}

GpuFeatureTransformationEstimator::GpuFeatureTransformationEstimator(boost::function<void(SlamEdge)> callback, const std::vector<int>& devices)
    : TransformationEstimator(callback) {
    init(devices);
}

void GpuFeatureTransformationEstimator::init(const std::vector<int>& devices) {
    std::vector<int32_t> dev(devices.begin(), devices.end());
    if (uz_group_create(dev.data(), (int32_t)dev.size(), &grp_) != UZ_OK)
        throw std::runtime_error(std::string("uz_group_create failed (no CPU fallback): ") + uz_group_last_error(nullptr));
    ctx_ = uz_group_context(grp_, 0);
    setConfig(config_);
    startThread();
}

GpuFeatureTransformationEstimator::~GpuFeatureTransformationEstimator() {
    stopThread();
    uz_group_destroy(grp_);
}

const char* GpuFeatureTransformationEstimator::lastError() const {
    const char* g = uz_group_last_error(grp_);
    return (g && g[0]) ? g : uz_last_error(ctx_);
}

void GpuFeatureTransformationEstimator::setConfig(transformation_estimation::FeatureLinkEstimationConfig config) {
    std::lock_guard<std::mutex> lk(gpu_mutex_);
    config_ = config;                                     // feature_transformation_estimator.cpp:350-353
    uz_params p;
    uz_default_params(&p);
    p.ransac_threshold = config.ransac_threshold;
    p.ransac_iterations = config.ransac_iteration;
    p.break_percentage = config.ransac_break_percentage;
    // link_covariance is unused by the live reference code (:138-144 commented out); use_epnp is ignored (:130)
    if (uz_group_set_params(grp_, &p) != UZ_OK) std::fprintf(stderr, "setConfig: %s\n", lastError());
}

int GpuFeatureTransformationEstimator::internFrame(const std::string& frame) {
    auto it = frames_.find(frame);
    if (it != frames_.end()) return it->second;
    const int id = (int)frames_.size();
    frames_[frame] = id;
    return id;
}

static void feature_view(const FeatureData& f, int frame_tag, std::vector<uint8_t>& valid_bytes, uz_features* out) {
    const int n = (int)f.feature_positions_.cols();
    valid_bytes.resize((size_t)n);
    for (int i = 0; i < n; ++i) valid_bytes[i] = f.valid_3d_[i] ? 1 : 0;     // std::vector<bool> -> bytes
    out->descriptors = f.features_.data;
    out->positions = f.feature_positions_.data();
    out->valid_3d = valid_bytes.data();
    out->n = n;
    out->desc_stride = (int)f.features_.step;
    out->desc_bytes = f.features_.cols;          // 32 (ORB, BRIEF) or 64 (BRISK, FREAK)
    out->feature_type = f.feature_type_;
    out->sensor_frame = frame_tag;
}

void GpuFeatureTransformationEstimator::collectCams(const SlamNode& node, std::vector<FeatureDataPtr>& cams) const {
    cams.clear();
    for (const SensorDataPtr& d : node.sensor_data_)
        if (d->type_ == graph_slam_msgs::SensorData::SENSOR_TYPE_FEATURE) {           // :41,:43
            FeatureDataPtr f = boost::dynamic_pointer_cast<FeatureData>(d);
            if (f) cams.push_back(f);
        }
}

// true when the device copy was made from exactly the node's FeatureData objects: same objects (FeatureData is immutable
// once published, sensor_data.h:113-114), same sizes.  The common case - node resident and unchanged -
// allocates nothing.
static bool node_matches(const SlamNode& node, const std::vector<FeatureDataPtr>& have, const std::vector<int>& rows) {
    size_t k = 0;
    for (const SensorDataPtr& d : node.sensor_data_) {
        if (d->type_ != graph_slam_msgs::SensorData::SENSOR_TYPE_FEATURE) continue;
        const FeatureData* f = dynamic_cast<const FeatureData*>(d.get());
        if (!f) continue;
        if (k >= have.size() || have[k].get() != f || rows[k] != f->features_.rows) return false;
        ++k;
    }
    return k == have.size();
}

bool GpuFeatureTransformationEstimator::ensureResident(const SlamNode& node, Resident** out) {
    auto it = node.id_.empty() ? handles_.end() : handles_.find(node.id_);
    if (it != handles_.end() && node_matches(node, it->second.cams, it->second.rows)) { *out = &it->second; return true; }
    std::vector<FeatureDataPtr> cams;
    collectCams(node, cams);
    std::vector<uz_features> views(cams.size());
    std::vector<std::vector<uint8_t> > valid(cams.size());
    std::vector<int> rows(cams.size());
    for (size_t i = 0; i < cams.size(); ++i) {
        feature_view(*cams[i], internFrame(cams[i]->sensor_frame_), valid[i], &views[i]);
        rows[i] = cams[i]->features_.rows;
    }
    if (it != handles_.end()) {
        // the node's sensor data changed under its id (graph_slam_node.cpp:244, :1010-1026): the reference copies the node on
        // every estimateEdge (transformation_estimator.cpp:39) and so always matches the current data - re-upload, same handle
        if (uz_group_store_replace(grp_, it->second.handle, views.data(), (int32_t)views.size()) != UZ_OK) return false;
        ++store_epoch_;
        it->second.cams = cams; it->second.rows = rows;
        *out = &it->second;
        return true;
    }
    Resident r;
    r.cams = cams; r.rows = rows;
    if (uz_group_store_add(grp_, views.data(), (int32_t)views.size(), &r.handle) != UZ_OK) return false;
    ++store_epoch_;
    std::string key = node.id_.empty() ? "#anon" + std::to_string(anon_++) : node.id_;
    auto ins = handles_.emplace(key, r);
    *out = &ins.first->second;
    return true;
}

bool GpuFeatureTransformationEstimator::loadNodes(const std::vector<SlamNode>& nodes) {
    std::lock_guard<std::mutex> lk(gpu_mutex_);
    std::vector<const SlamNode*> ptrs;
    ptrs.reserve(nodes.size());
    for (const SlamNode& n : nodes) ptrs.push_back(&n);
    return loadNodesLocked(ptrs);
}

bool GpuFeatureTransformationEstimator::loadNodesLocked(const std::vector<const SlamNode*>& nodes) {
    std::vector<uz_features> views;
    std::vector<std::vector<uint8_t> > valid;
    std::vector<int32_t> counts;
    std::vector<Resident> res;
    std::vector<const SlamNode*> fresh;
    size_t total = 0;
    for (const SlamNode* np : nodes) {
        const SlamNode& n = *np;
        if (n.id_.empty() || handles_.count(n.id_)) continue;
        Resident r;
        collectCams(n, r.cams);
        total += r.cams.size();
        res.push_back(r);
        fresh.push_back(&n);
    }
    views.resize(total); valid.resize(total);
    size_t k = 0;
    for (Resident& r : res) {
        r.rows.resize(r.cams.size());
        for (size_t i = 0; i < r.cams.size(); ++i, ++k) {
            feature_view(*r.cams[i], internFrame(r.cams[i]->sensor_frame_), valid[k], &views[k]);
            r.rows[i] = r.cams[i]->features_.rows;
        }
        counts.push_back((int32_t)r.cams.size());
    }
    std::vector<int32_t> handles(res.size());
    if (!res.empty() && uz_group_store_add_bulk(grp_, views.data(), counts.data(), (int32_t)res.size(), handles.data()) != UZ_OK) {
        std::fprintf(stderr, "loadNodes: %s\n", lastError());
        return false;
    }
    for (size_t i = 0; i < res.size(); ++i) { res[i].handle = handles[i]; handles_.emplace(fresh[i]->id_, res[i]); }
    if (!res.empty()) ++store_epoch_;
    return true;
}

void GpuFeatureTransformationEstimator::forgetNode(const std::string& id) {
    std::lock_guard<std::mutex> lk(gpu_mutex_);
    auto it = handles_.find(id);
    if (it == handles_.end()) return;
    uz_group_store_remove(grp_, it->second.handle);
    handles_.erase(it);
    ++store_epoch_;
}

static void set_identity6(Eigen::MatrixXd& m) {
    if (m.rows() != 6 || m.cols() != 6) { m = Eigen::MatrixXd::Identity(6, 6); return; }
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) m(a, b) = a == b ? 1.0 : 0.0;
}

// the state of SlamEdge() (slam_edge.cpp:22-25 / slam_edge.h:47-93), in place
static void resetEdge(SlamEdge& e) {
    e.id_.clear(); e.id_from_.clear(); e.id_to_.clear();
    e.transform_ = Eigen::Isometry3d::Identity();
    e.displacement_from_ = Eigen::Isometry3d::Identity();
    e.displacement_to_ = Eigen::Isometry3d::Identity();
    set_identity6(e.information_);
    e.type_ = 0;
    e.sensor_from_.clear(); e.sensor_to_.clear();
    e.age_ = 0.; e.error_ = 0.; e.matching_score_ = 0.; e.valid_ = false;        // init(): slam_edge.cpp:33-48
    e.diff_time_ = ros::Duration(1);
}

void GpuFeatureTransformationEstimator::fillEdge(const uz_edge_result& r, const Resident& from, const Resident& to,
                                                 SlamEdge& edge) const {
    if (!r.ok) return;                                   // on failure the edge keeps its default-constructed fields
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) edge.transform_(a, b) = r.T[4 * a + b];               // :147
    set_identity6(edge.information_);
    if (r.consensus > 0 && r.mse > 0) {                                                     // :134-137
        edge.information_ *= r.info_scale;
        for (int a = 3; a < 6; ++a)
            for (int b = 3; b < 6; ++b) edge.information_(a, b) *= 100.;
    }
    edge.type_ = graph_slam_msgs::Edge::TYPE_3D_FULL;                                       // :149
    const FeatureData& f = *from.cams[r.cam_from];
    const FeatureData& t = *to.cams[r.cam_to];
    edge.sensor_from_ = f.sensor_frame_;                                                    // :150-153
    edge.sensor_to_ = t.sensor_frame_;
    edge.displacement_from_ = f.displacement_;
    edge.displacement_to_ = t.displacement_;
    edge.matching_score_ = r.consensus;                                                     // :155
}

// ---- the worker's three-step form ------------------------------------------------------------------------------------------
// The worker thread owns gpu_mutex_ while a chunk is in flight (submit .. finish, with the look-ups of the next chunk in
// between) and gives it up after every finished chunk; store_epoch_ tells submit whether somebody changed the store meanwhile.
GpuFeatureTransformationEstimator::Resident* GpuFeatureTransformationEstimator::lookupResident(const SlamNode& node) {
    if (node.id_.empty()) return nullptr;
    auto it = handles_.find(node.id_);
    return (it != handles_.end() && node_matches(node, it->second.cams, it->second.rows)) ? &it->second : nullptr;
}

void GpuFeatureTransformationEstimator::prepareEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >& pairs, int slot) {
    const bool own = !holding_;            // first chunk of a burst: nothing in flight, take the lock for the look-ups only
    if (own) gpu_mutex_.lock();
    Flight& F = flights_[slot];
    const size_t n = pairs.size();
    F.hf.resize(n); F.ht.resize(n); F.rf.resize(n); F.rt.resize(n);
    F.all_resident = true; F.failed = false; F.epoch = store_epoch_;
    for (size_t i = 0; i < n; ++i) {
        Resident* a = lookupResident(pairs[i].first);
        Resident* b = lookupResident(pairs[i].second);
        if (!a || !b) { F.all_resident = false; break; }
        F.rf[i] = a; F.rt[i] = b; F.hf[i] = a->handle; F.ht[i] = b->handle;
    }
    if (own) gpu_mutex_.unlock();
}

void GpuFeatureTransformationEstimator::submitEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >& pairs, std::vector<SlamEdge>& edges,
                                                        std::vector<char>& ok, int slot) {
    Flight& F = flights_[slot];
    const size_t n = pairs.size();
    edges.resize(n);
    for (SlamEdge& e : edges) resetEdge(e);               // == default constructed, without giving the 6 x 6 matrix back to the heap
    ok.assign(n, 0);
    if (!holding_) { gpu_mutex_.lock(); holding_ = true; }
    if (!F.all_resident || F.epoch != store_epoch_) {     // nodes to upload (nothing is in flight now), or the store changed during a yield
        if (!residentsOf(pairs, F.hf, F.ht, F.rf, F.rt)) { F.failed = true; std::fprintf(stderr, "estimateEdgeBatch: %s\n", lastError()); return; }
    }
    F.res.resize(n);
    if (n > 0 && uz_group_estimate_edges_begin(grp_, F.hf.data(), F.ht.data(), (int32_t)n, F.res.data()) != UZ_OK) {
        F.failed = true;
        std::fprintf(stderr, "uz_group_estimate_edges_begin: %s\n", lastError());
    }
}

void GpuFeatureTransformationEstimator::finishEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >& pairs, std::vector<SlamEdge>& edges,
                                                        std::vector<char>& ok, int slot) {
    Flight& F = flights_[slot];
    const size_t n = pairs.size();
    uz_status st = F.failed ? UZ_ERR_INVALID : uz_group_estimate_edges_end(grp_);
    if (!F.failed && st != UZ_OK) std::fprintf(stderr, "uz_group_estimate_edges_end: %s\n", lastError());
    for (size_t i = 0; i < n; ++i) {
        if (st == UZ_OK) { fillEdge(F.res[i], *F.rf[i], *F.rt[i], edges[i]); ok[i] = F.res[i].ok ? 1 : 0; }
        edges[i].id_from_ = pairs[i].first.id_;          // :168-169, set even on failure
        edges[i].id_to_ = pairs[i].second.id_;
    }
    holding_ = false;
    gpu_mutex_.unlock();                                  // nothing is in flight: the place recogniser / forgetNode get their turn
}

void GpuFeatureTransformationEstimator::endOfBurst() {
    if (holding_) { holding_ = false; gpu_mutex_.unlock(); }
}

// handles and device copies of every node of a batch (uploads what is missing or stale); gpu_mutex_ held
bool GpuFeatureTransformationEstimator::residentsOf(std::vector<std::pair<SlamNode, SlamNode> >& pairs, std::vector<int32_t>& hf,
                                                    std::vector<int32_t>& ht, std::vector<Resident*>& rf, std::vector<Resident*>& rt) {
    const size_t n = pairs.size();
    hf.resize(n); ht.resize(n); rf.resize(n); rt.resize(n);
    {   // nodes this batch sees for the first time travel in ONE bulk upload (a cold queue otherwise pays one upload call per node)
        std::vector<const SlamNode*> fresh;
        std::unordered_set<std::string> seen;
        for (size_t i = 0; i < n; ++i)
            for (const SlamNode* nd : {&pairs[i].first, &pairs[i].second})
                if (!nd->id_.empty() && !handles_.count(nd->id_) && seen.insert(nd->id_).second) fresh.push_back(nd);
        if (fresh.size() > 1 && !loadNodesLocked(fresh)) return false;
    }
    for (size_t i = 0; i < n; ++i) {
        if (!ensureResident(pairs[i].first, &rf[i]) || !ensureResident(pairs[i].second, &rt[i])) return false;
        hf[i] = rf[i]->handle;
        ht[i] = rt[i]->handle;
    }
    return true;
}

void GpuFeatureTransformationEstimator::estimateEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >& pairs,
                                                          std::vector<SlamEdge>& edges, std::vector<char>& ok) {
    std::lock_guard<std::mutex> lk(gpu_mutex_);
    const size_t n = pairs.size();
    edges.resize(n);
    for (SlamEdge& e : edges) resetEdge(e);               // == default constructed, without giving the 6 x 6 matrix back to the heap
    ok.assign(n, 0);
    std::vector<int32_t> hf, ht;
    std::vector<Resident*> rf, rt;
    if (!residentsOf(pairs, hf, ht, rf, rt)) {
        std::fprintf(stderr, "estimateEdgeBatch: %s\n", lastError());
        for (size_t k = 0; k < n; ++k) { edges[k].id_from_ = pairs[k].first.id_; edges[k].id_to_ = pairs[k].second.id_; }
        return;
    }
    std::vector<uz_edge_result> res(n);
    const uz_status st = uz_group_estimate_edges(grp_, hf.data(), ht.data(), (int32_t)n, res.data());
    for (size_t i = 0; i < n; ++i) {
        if (st == UZ_OK) { fillEdge(res[i], *rf[i], *rt[i], edges[i]); ok[i] = res[i].ok ? 1 : 0; }
        edges[i].id_from_ = pairs[i].first.id_;          // :168-169, set even on failure
        edges[i].id_to_ = pairs[i].second.id_;
    }
    if (st != UZ_OK) std::fprintf(stderr, "uz_group_estimate_edges: %s\n", lastError());
}

bool GpuFeatureTransformationEstimator::estimateEdgeImpl(SlamNode& from, SlamNode& to, SlamEdge& edge) {
    std::vector<std::pair<SlamNode, SlamNode> > one(1, std::make_pair(from, to));
    std::vector<SlamEdge> edges;
    std::vector<char> ok;
    estimateEdgeBatch(one, edges, ok);
    edge = edges[0];
    return ok[0] != 0;
}

bool GpuFeatureTransformationEstimator::estimateEdgeDirect(std::vector<SensorDataPtr> from, std::vector<SensorDataPtr> to,
                                                           SlamEdge& edge) {
    std::lock_guard<std::mutex> lk(gpu_mutex_);
    Resident rf, rt;
    std::vector<uz_features> vf, vt;
    std::vector<std::vector<uint8_t> > valf, valt;
    auto collect = [this](const std::vector<SensorDataPtr>& in, Resident& r, std::vector<uz_features>& v,
                          std::vector<std::vector<uint8_t> >& val) {
        for (const SensorDataPtr& d : in)
            if (d->type_ == graph_slam_msgs::SensorData::SENSOR_TYPE_FEATURE) {
                FeatureDataPtr f = boost::dynamic_pointer_cast<FeatureData>(d);
                if (f) r.cams.push_back(f);
            }
        v.resize(r.cams.size());
        val.resize(r.cams.size());
        for (size_t i = 0; i < r.cams.size(); ++i) feature_view(*r.cams[i], internFrame(r.cams[i]->sensor_frame_), val[i], &v[i]);
    };
    collect(from, rf, vf, valf);
    collect(to, rt, vt, valt);
    const int32_t nf = (int32_t)vf.size(), nt = (int32_t)vt.size();
    uz_edge_result r;
    if (uz_estimate_edges_host(ctx_, vf.data(), &nf, vt.data(), &nt, 1, &r) != UZ_OK) {
        std::fprintf(stderr, "estimateEdgeDirect: %s\n", uz_last_error(ctx_));
        return false;
    }
    fillEdge(r, rf, rt, edge);
    return r.ok != 0;
}

void GpuFeatureTransformationEstimator::estimateSVD(Eigen::MatrixXd P, Eigen::MatrixXd Q, Eigen::Isometry3d& T, int& consensus,
                                                    double& mse, double maxError, int iterations, double breakPercentage,
                                                    bool do_prosac) {
    std::lock_guard<std::mutex> lk(gpu_mutex_);
    double T16[16];
    int32_t c = 0;
    consensus = 0; mse = 0.; T = Eigen::Isometry3d::Identity();
    if (uz_estimate_svd(ctx_, P.data(), Q.data(), P.cols(), maxError, iterations, breakPercentage, do_prosac ? 1 : 0, nullptr,
                        T16, &c, &mse, nullptr, nullptr, nullptr) != UZ_OK) {
        std::fprintf(stderr, "estimateSVD: %s\n", uz_last_error(ctx_));
        return;
    }
    consensus = c;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) T(a, b) = T16[4 * a + b];
}

int GpuFeatureTransformationEstimator::consensus3D(Eigen::MatrixXd P, Eigen::MatrixXd Q, Eigen::Isometry3d T, double thresh,
                                                   Eigen::Array<bool, 1, Eigen::Dynamic>& consensusSet) {
    std::lock_guard<std::mutex> lk(gpu_mutex_);
    double T16[16];
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) T16[4 * a + b] = T(a, b);
    std::vector<uint8_t> set((size_t)std::max(P.cols(), 1));
    int32_t count = 0;
    consensusSet.resize(P.cols());
    if (uz_consensus3d(ctx_, P.data(), Q.data(), P.cols(), T16, thresh, set.data(), &count) != UZ_OK) {
        std::fprintf(stderr, "consensus3D: %s\n", uz_last_error(ctx_));
        return 0;
    }
    for (int i = 0; i < P.cols(); ++i) consensusSet[i] = set[i] != 0;
    return count;
}

bool GpuFeatureTransformationEstimator::residentHandle(const SlamNode& node, int32_t* handle) {
    Resident* r = nullptr;
    if (!ensureResident(node, &r)) return false;
    *handle = r->handle;
    return true;
}

void GpuFeatureTransformationEstimator::acceptEdges(const std::vector<SlamEdge>& edges, double min_matching_score,
                                                    double max_edge_distance_T, double max_edge_distance_R, std::vector<char>& accept) {
    std::lock_guard<std::mutex> lk(gpu_mutex_);
    const size_t n = edges.size();
    accept.assign(n, 0);
    if (n == 0) return;
    std::vector<uz_edge_result> rec(n);
    for (size_t i = 0; i < n; ++i) {
        std::memset(&rec[i], 0, sizeof(rec[i]));
        rec[i].ok = 1;                                                   // a failed estimate already carries score 0
        rec[i].consensus = (int32_t)edges[i].matching_score_;            // matching_score_ = consensus (:155)
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) rec[i].T[4 * a + b] = edges[i].transform_(a, b);
    }
    uz_gate_params g;
    g.min_matching_score = min_matching_score; g.max_edge_distance_T = max_edge_distance_T; g.max_edge_distance_R = max_edge_distance_R;
    std::vector<uint8_t> acc(n);
    if (uz_gate_edges(ctx_, rec.data(), (int32_t)n, &g, acc.data(), nullptr, nullptr) != UZ_OK) {
        std::fprintf(stderr, "acceptEdges: %s\n", uz_last_error(ctx_));
        return;
    }
    for (size_t i = 0; i < n; ++i) accept[i] = (char)acc[i];
}

void GpuFeatureTransformationEstimator::estimateSVDBatch(const std::vector<Eigen::MatrixXd>& P, const std::vector<Eigen::MatrixXd>& Q,
                                                         std::vector<Eigen::Isometry3d>& T, std::vector<int>& consensus,
                                                         std::vector<double>& mse, std::vector<std::vector<char> >& consensus_sets,
                                                         double maxError, int iterations, double breakPercentage, bool do_prosac) {
    std::lock_guard<std::mutex> lk(gpu_mutex_);
    const size_t n = P.size();
    T.assign(n, Eigen::Isometry3d::Identity());
    consensus.assign(n, 0); mse.assign(n, 0.); consensus_sets.assign(n, std::vector<char>());
    if (n == 0 || Q.size() != n) return;
    std::vector<int32_t> off(n + 1, 0);
    for (size_t i = 0; i < n; ++i) off[i + 1] = off[i] + P[i].cols();
    std::vector<double> p((size_t)off[n] * 3 + 3), q((size_t)off[n] * 3 + 3);
    for (size_t i = 0; i < n; ++i) {
        std::memcpy(p.data() + (size_t)off[i] * 3, P[i].data(), (size_t)P[i].cols() * 24);
        std::memcpy(q.data() + (size_t)off[i] * 3, Q[i].data(), (size_t)Q[i].cols() * 24);
    }
    std::vector<double> T16(16 * n), m(n);
    std::vector<int32_t> c(n);
    std::vector<uint8_t> mask((size_t)off[n] + 1);
    if (uz_estimate_svd_batch(ctx_, p.data(), q.data(), off.data(), (int32_t)n, maxError, iterations, breakPercentage, do_prosac ? 1 : 0,
                              T16.data(), c.data(), m.data(), mask.data()) != UZ_OK) {
        std::fprintf(stderr, "estimateSVDBatch: %s\n", uz_last_error(ctx_));
        return;
    }
    for (size_t i = 0; i < n; ++i) {
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) T[i](a, b) = T16[16 * i + 4 * a + b];
        consensus[i] = c[i]; mse[i] = m[i];
        consensus_sets[i].assign(mask.begin() + off[i], mask.begin() + off[i + 1]);
    }
}
