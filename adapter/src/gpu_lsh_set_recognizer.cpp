// gpu_lsh_set_recognizer.cpp — see the header.  Reference files are cited relative to /root/reference/.
#include <place_recognition/gpu_lsh_set_recognizer.h>

#include <chrono>
#include <cmath>
#include <cstdio>

static int64_t stamp_ns(const SlamNode& node) {
    if (node.stamps_.empty()) return 0;
#ifdef UZ_ADAPTER_REAL_HEADERS
    return (int64_t)node.stamps_.front().toNSec();                       // pr_time_map_[id] = node.stamps_.front()
#else
    return (int64_t)std::llround(node.stamps_.front().sec * 1e9);
#endif
}

GpuLshSetRecognizer::GpuLshSetRecognizer(GpuFeatureTransformationEstimator& estimator) : est_(estimator) {
    setConfig(config_);
    thread_ = std::thread(&GpuLshSetRecognizer::placeRecognitionThread, this);
}

GpuLshSetRecognizer::~GpuLshSetRecognizer() {
    {
        std::lock_guard<std::mutex> lk(res_mutex_);
        running_ = false;
    }
    cv_.notify_all();
    thread_.join();
}

void GpuLshSetRecognizer::setConfig(place_recognition::PlaceRecognizerConfig config) {      // place_recognizer.cpp:45-47
    config_ = config;
    uz_place_params p;
    uz_default_place_params(&p);
    p.T = config.T;
    p.k_nearest_neighbors = config.k_nearest_neighbors;
    std::lock_guard<std::mutex> lk(est_.gpuMutex());
    if (uz_places_set_params(est_.context(), &p) != UZ_OK) std::fprintf(stderr, "setConfig: %s\n", est_.lastError());
}

void GpuLshSetRecognizer::clear() {                                                         // place_recognizer.cpp:49-62
    std::lock_guard<std::mutex> lk(res_mutex_);
    std::lock_guard<std::mutex> gk(est_.gpuMutex());
    uz_places_clear(est_.context());
    pr_queue_.clear();
    potential_neighbors_.clear();
    id_of_handle_.clear();
}

// mode 0 = searchAndAddPlace, 1 = addPlace, 2 = searchPlace
std::vector<std::pair<std::string, std::string> > GpuLshSetRecognizer::run(int mode, const std::vector<SlamNode>& nodes) {
    std::vector<std::pair<std::string, std::string> > res;
    if (nodes.empty()) return res;
    std::lock_guard<std::mutex> gk(est_.gpuMutex());
    std::vector<int32_t> handles(nodes.size());
    std::vector<int64_t> stamps(nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) {
        if (!est_.residentHandle(nodes[i], &handles[i])) {
            std::fprintf(stderr, "place recognition: %s\n", est_.lastError());
            return res;
        }
        stamps[i] = stamp_ns(nodes[i]);
        id_of_handle_[handles[i]] = nodes[i].id_;
    }
    const int32_t cap = (int32_t)nodes.size() * std::max(config_.k_nearest_neighbors, 1);
    std::vector<int32_t> pairs((size_t)cap * 2);
    int32_t n = 0;
    uz_status st;
    if (mode == 0) st = uz_places_search_and_add(est_.context(), handles.data(), stamps.data(), (int32_t)nodes.size(), pairs.data(), cap, &n);
    else if (mode == 1) st = uz_places_add(est_.context(), handles.data(), stamps.data(), (int32_t)nodes.size());
    else st = uz_places_search(est_.context(), handles.data(), stamps.data(), (int32_t)nodes.size(), pairs.data(), cap, &n);
    if (st != UZ_OK) { std::fprintf(stderr, "place recognition: %s\n", est_.lastError()); return res; }
    for (int32_t i = 0; i < std::min(n, cap); ++i)
        res.push_back(std::make_pair(id_of_handle_[pairs[2 * i]], id_of_handle_[pairs[2 * i + 1]]));
    return res;
}

std::vector<std::pair<std::string, std::string> > GpuLshSetRecognizer::searchAndAddPlaces(const std::vector<SlamNode>& nodes) { return run(0, nodes); }
void GpuLshSetRecognizer::addPlaces(const std::vector<SlamNode>& nodes) { run(1, nodes); }

std::vector<std::pair<std::string, std::string> > GpuLshSetRecognizer::searchAndAddPlace(const SlamNode& node) {
    return run(0, std::vector<SlamNode>(1, node));
}
void GpuLshSetRecognizer::addPlace(const SlamNode& node) { run(1, std::vector<SlamNode>(1, node)); }               // :120-123
std::vector<std::pair<std::string, std::string> > GpuLshSetRecognizer::searchPlace(const SlamNode& node) {         // :149-152
    return run(2, std::vector<SlamNode>(1, node));
}

void GpuLshSetRecognizer::addNode(const SlamNode& node) {                                   // place_recognizer.cpp:64-69
    std::lock_guard<std::mutex> lk(res_mutex_);
    pr_queue_.push_back(node);
    cv_.notify_all();
}
void GpuLshSetRecognizer::addPlaceQueue(const SlamNode& node) {                             // :125-130
    std::lock_guard<std::mutex> lk(res_mutex_);
    pr_add_queue_.push_back(node);
    cv_.notify_all();
}
void GpuLshSetRecognizer::removePlaceQueue(const SlamNode& node) {                          // :192-196
    std::lock_guard<std::mutex> lk(res_mutex_);
    pr_remove_queue_.push_back(node.id_);
    cv_.notify_all();
}

void GpuLshSetRecognizer::removePlace(const std::string& id) {                              // :198-230
    std::lock_guard<std::mutex> gk(est_.gpuMutex());
    for (auto& kv : id_of_handle_)
        if (kv.second == id) {
            if (uz_places_remove(est_.context(), kv.first) != UZ_OK)
                std::fprintf(stderr, "tried to remove a non-existing place: %s\n", id.c_str());
            return;
        }
    std::fprintf(stderr, "tried to remove a non-existing place: %s\n", id.c_str());
}

std::vector<std::pair<std::string, std::string> > GpuLshSetRecognizer::recognizedPlaces() {   // :237-244
    std::lock_guard<std::mutex> lk(res_mutex_);
    auto res = potential_neighbors_;
    potential_neighbors_.clear();
    return res;
}

bool GpuLshSetRecognizer::hasRecognizedPlaces() {                                           // :246-249
    std::lock_guard<std::mutex> lk(res_mutex_);
    return !potential_neighbors_.empty();
}

// place_recognizer.cpp:251-290, batched.  The reference pops pr_queue_ from the BACK (newest first); the whole queue is
// taken here in that same order.
void GpuLshSetRecognizer::placeRecognitionThread() {
    std::unique_lock<std::mutex> lk(res_mutex_);
    while (true) {
        cv_.wait(lk, [this] { return !running_ || !pr_queue_.empty() || !pr_add_queue_.empty() || !pr_remove_queue_.empty(); });
        if (!running_) break;
        std::vector<SlamNode> search(pr_queue_.rbegin(), pr_queue_.rend());
        std::vector<SlamNode> add(pr_add_queue_.rbegin(), pr_add_queue_.rend());
        std::vector<std::string> rem(pr_remove_queue_.rbegin(), pr_remove_queue_.rend());
        pr_queue_.clear(); pr_add_queue_.clear(); pr_remove_queue_.clear();
        busy_ = true;
        lk.unlock();
        auto neighbors = run(0, search);
        for (const std::string& id : rem) removePlace(id);
        run(1, add);
        lk.lock();
        potential_neighbors_.insert(potential_neighbors_.end(), neighbors.begin(), neighbors.end());
        busy_ = false;
        cv_.notify_all();
    }
}

void GpuLshSetRecognizer::waitIdle() {
    std::unique_lock<std::mutex> lk(res_mutex_);
    cv_.wait(lk, [this] { return pr_queue_.empty() && pr_add_queue_.empty() && pr_remove_queue_.empty() && !busy_; });
}
