// gpu_feature_transformation_estimator.h — host-side mirror of the reference's estimator interface for the
// feature-edge path, on top of the C-ABI (include/uzliti_edge.h).
//
//   class TransformationEstimator            <- transformation_estimation/include/transformation_estimation/
//                                               transformation_estimator.h:45-67 (+ .cpp:22-62)
//   class GpuFeatureTransformationEstimator  <- .../feature_transformation_estimator.h:33-58
//
// Same class shape, method names, argument meaning and failure convention (estimateEdgeImpl returns bool;
// on false the worker zeroes matching_score_ and STILL fires the callback, transformation_estimator.cpp:53-56).
// What changes is the execution model: the reference pops ONE pair per 1 ms tick on its worker thread; here
// the worker drains the whole queue and hands it to the GPU as one batch (estimateEdgeBatch), with the
// nodes' FeatureData cached in the device-resident keyframe store.  Callbacks are fired from a delivery thread (in
// estimation order, while the next chunk is being estimated), never while the caller of estimateEdge() is on the stack
// (the reference's callers hold graph_mutex_).
#pragma once
#ifdef UZ_ADAPTER_REAL_HEADERS
#include <graph_slam_common/slam_node.h>
#include <transformation_estimation/FeatureLinkEstimationConfig.h>
#else
#include <graph_slam_common/shim_types.h>
#endif
#include <condition_variable>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "uzliti_edge.h"

class TransformationEstimator {
public:
    TransformationEstimator(boost::function<void(SlamEdge)> callback);
    virtual ~TransformationEstimator();

    void estimateEdge(SlamNode& from, SlamNode& to);                                   // transformation_estimator.h:52
    virtual bool estimateEdgeImpl(SlamNode& from, SlamNode& to, SlamEdge& edge) = 0;   // :54
    // batch hook: default = one estimateEdgeImpl per pair; the GPU estimator overrides it
    virtual void estimateEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >& pairs, std::vector<SlamEdge>& edges,
                                   std::vector<char>& ok);
    // the worker's form of the same, in three steps per chunk of a drained queue (slot = 0 / 1, two chunks alternate):
    // prepare(c + 1) runs while chunk c is still in flight, then finish(c), then submit(c + 1); endOfBurst() after the last
    // finish.  Defaults: submit = estimateEdgeBatch, the rest nothing.
    virtual void prepareEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >& pairs, int slot);
    virtual void submitEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >& pairs, std::vector<SlamEdge>& edges, std::vector<char>& ok, int slot);
    virtual void finishEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >& pairs, std::vector<SlamEdge>& edges, std::vector<char>& ok, int slot);
    virtual void endOfBurst();

    std::map<std::string, Eigen::Isometry3d> sensor_transforms_;                       // :56 (unused by this path)

    void waitIdle();                                  // test helper: block until the queue is drained and delivered

protected:
    void startThread();                               // called by the most-derived constructor
    void stopThread();
    void estimationThread();
    void deliveryThread();

    // edges of one chunk on their way to the callback; two slots, filled and delivered in turn
    struct Delivery { std::vector<SlamEdge> edges; std::vector<char> ok; };
    int delivery_chunk_ = 4096;                       // pairs per estimateEdgeBatch call of a drained queue (UZ_ADAPTER_CHUNK)
    Delivery slots_[2];
    bool slot_full_[2] = {false, false};
    std::thread delivery_thread_;
    std::thread estimation_thread_;
    std::mutex estimation_mutex_;
    std::condition_variable cv_;
    bool running_ = false;
    bool busy_ = false;
    std::vector<std::pair<SlamNode, SlamNode> > est_queue_;
    boost::function<void(SlamEdge)> callback_;
};

class GpuFeatureTransformationEstimator : public TransformationEstimator {
public:
    GpuFeatureTransformationEstimator(boost::function<void(SlamEdge)> callback, int device = 0);
    // Several GPUs of one box behind the same class (uz_group): the node store is replicated on every device (one upload
    // over PCIe, the other devices pull from the first one's HBM over NVLink), a batch of queued pairs is cut into
    // contiguous shards, and every device's solve kernel writes its edge records straight into one result buffer.  Edges
    // are byte-identical to the single-device estimator's.
    GpuFeatureTransformationEstimator(boost::function<void(SlamEdge)> callback, const std::vector<int>& devices);
    ~GpuFeatureTransformationEstimator();

    bool estimateEdgeImpl(SlamNode& from, SlamNode& to, SlamEdge& edge);                                  // :38
    void setConfig(transformation_estimation::FeatureLinkEstimationConfig config);                         // :40
    bool estimateEdgeDirect(std::vector<SensorDataPtr> from, std::vector<SensorDataPtr> to, SlamEdge& edge);   // :42
    void estimateSVD(Eigen::MatrixXd P, Eigen::MatrixXd Q, Eigen::Isometry3d& T, int& consensus, double& mse,
                     double maxError, int iterations, double breakPercentage, bool do_prosac = true);      // :45
    int consensus3D(Eigen::MatrixXd P, Eigen::MatrixXd Q, Eigen::Isometry3d T, double thresh,
                    Eigen::Array<bool, 1, Eigen::Dynamic>& consensusSet);                                   // :53
    void estimateEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >& pairs, std::vector<SlamEdge>& edges,
                           std::vector<char>& ok);
    // look-ups of chunk c + 1 beside the device work of chunk c (uz_group_estimate_edges_begin / _end)
    void prepareEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >& pairs, int slot);
    void submitEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >& pairs, std::vector<SlamEdge>& edges, std::vector<char>& ok, int slot);
    void finishEdgeBatch(std::vector<std::pair<SlamNode, SlamNode> >& pairs, std::vector<SlamEdge>& edges, std::vector<char>& ok, int slot);
    void endOfBurst();

    // the steps after the path, batched
    // GraphSlamNode::newEdgeCallback's numeric gate (graph_slam_node.cpp:798-804) for many edges at once
    void acceptEdges(const std::vector<SlamEdge>& edges, double min_matching_score, double max_edge_distance_T,
                     double max_edge_distance_R, std::vector<char>& accept);
    // TransformationFilter::calcValidEdges' per-cluster estimateSVD + consensus3D (transformation_filter.cpp:266-275)
    void estimateSVDBatch(const std::vector<Eigen::MatrixXd>& P, const std::vector<Eigen::MatrixXd>& Q,
                          std::vector<Eigen::Isometry3d>& T, std::vector<int>& consensus, std::vector<double>& mse,
                          std::vector<std::vector<char> >& consensus_sets, double maxError = 0.3, int iterations = 200,
                          double breakPercentage = 1.0, bool do_prosac = false);

    // shared with the place recogniser (adapter/include/place_recognition/gpu_lsh_set_recognizer.h)
    uz_context* context() { return ctx_; }
    std::mutex& gpuMutex() { return gpu_mutex_; }
    bool residentHandle(const SlamNode& node, int32_t* handle);       // uploads the node if needed; gpuMutex() must be held

    // device-resident keyframe store: nodes are uploaded once and addressed by id_ afterwards.  Every call re-validates the
    // cached copy against the node it is handed (the FEATURE sensor data of a node changes under a fixed id: late sensor
    // arrivals graph_slam_node.cpp:244, node merge :1010-1026) and re-uploads it under the same handle when it differs.
    // forgetNode gives the node's device memory back; call it where the reference removes a node (graph.removeNode,
    // graph_slam_node.cpp:1050) - see INTEGRATION.md.
    void forgetNode(const std::string& id);
    // Resume (GraphSlamNode::load, graph_slam_node.cpp:875-888: every stored node is re-added): all nodes in ONE bulk upload.
    bool loadNodes(const std::vector<SlamNode>& nodes);
    int devices() const { return uz_group_size(grp_); }
    int64_t storeBytes() const { return uz_store_bytes(ctx_); }
    size_t residentNodes() const { return handles_.size(); }
    const char* lastError() const;

protected:
    struct Resident { int32_t handle = -1; std::vector<FeatureDataPtr> cams; std::vector<int> rows; };
    bool ensureResident(const SlamNode& node, Resident** out);
    Resident* lookupResident(const SlamNode& node);                       // resident and unchanged, or nullptr; no upload
    bool residentsOf(std::vector<std::pair<SlamNode, SlamNode> >& pairs, std::vector<int32_t>& hf, std::vector<int32_t>& ht,
                     std::vector<Resident*>& rf, std::vector<Resident*>& rt);
    struct Flight {                                                       // one chunk between prepare and finish
        std::vector<int32_t> hf, ht;
        std::vector<Resident*> rf, rt;
        std::vector<uz_edge_result> res;
        bool all_resident = false, failed = false;
        unsigned long long epoch = 0;
    };
    Flight flights_[2];
    bool holding_ = false;                // the worker thread holds gpu_mutex_ (prepare .. endOfBurst)
    unsigned long long store_epoch_ = 0;  // bumped by everything that changes handles_
    bool loadNodesLocked(const std::vector<const SlamNode*>& nodes);      // gpuMutex() held
    void fillEdge(const uz_edge_result& r, const Resident& from, const Resident& to, SlamEdge& edge) const;
    int internFrame(const std::string& frame);

    void init(const std::vector<int>& devices);
    void collectCams(const SlamNode& node, std::vector<FeatureDataPtr>& cams) const;

    uz_group* grp_ = nullptr;
    uz_context* ctx_ = nullptr;           // the group's first device: stage entry points and the place recogniser run there
    transformation_estimation::FeatureLinkEstimationConfig config_;
    std::unordered_map<std::string, Resident> handles_;
    std::unordered_map<std::string, int> frames_;
    size_t anon_ = 0;
    std::mutex gpu_mutex_;
};
