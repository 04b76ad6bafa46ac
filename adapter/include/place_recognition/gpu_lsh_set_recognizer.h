// gpu_lsh_set_recognizer.h — host-side mirror of the reference's place-recognition interface for the LSH back end,
// on top of the C-ABI (include/uzliti_edge.h, uz_places_*).
//
//   class GpuLshSetRecognizer  <- place_recognition/include/place_recognition/place_recognizer.h:37-107 (public calls of
//                                 PlaceRecognizer) + lsh_set_recognizer.h:77-92 (LshSetRecognizer)
//
// Same method names, argument meaning and result shape (pairs of node ids: (recognised older node, new node)).  What
// changes is the execution model: the reference's thread pops ONE queued node per 1 ms tick (place_recognizer.cpp:
// 248-290); here the worker drains the whole queue and hands it to the GPU as one batch — the device recogniser
// guarantees the batch equals node-by-node processing in queue order.  Descriptors are not copied a second time: the
// recogniser votes over the estimator's device-resident keyframe store, so it is constructed on an estimator.
#pragma once
#include <transformation_estimation/gpu_feature_transformation_estimator.h>

#include <deque>

#ifndef UZ_ADAPTER_REAL_HEADERS
namespace place_recognition {
struct PlaceRecognizerConfig {   // cfg/PlaceRecognizer.cfg:10-12 (generated struct)
    int k_nearest_neighbors = 10;
    int history = 10;
    double T = 10;
};
}  // namespace place_recognition
#else
#include <place_recognition/PlaceRecognizerConfig.h>
#endif

class GpuLshSetRecognizer {
public:
    explicit GpuLshSetRecognizer(GpuFeatureTransformationEstimator& estimator);
    ~GpuLshSetRecognizer();

    void setConfig(place_recognition::PlaceRecognizerConfig config);                              // place_recognizer.h:43
    void clear();                                                                                 // :45
    void addNode(const SlamNode& node);                                                           // :47 (queued search-and-add)
    std::vector<std::pair<std::string, std::string> > searchAndAddPlace(const SlamNode& node);    // :49 (by node: needs the stamp)
    void addPlace(const SlamNode& node);                                                          // :51
    void addPlaceQueue(const SlamNode& node);                                                     // :53
    std::vector<std::pair<std::string, std::string> > searchPlace(const SlamNode& node);          // :57
    void removePlaceQueue(const SlamNode& node);                                                  // :61
    void removePlace(const std::string& id);                                                      // :63
    std::vector<std::pair<std::string, std::string> > recognizedPlaces();                         // :67
    bool hasRecognizedPlaces();                                                                   // :69

    // batch forms (what the worker uses)
    std::vector<std::pair<std::string, std::string> > searchAndAddPlaces(const std::vector<SlamNode>& nodes);
    void addPlaces(const std::vector<SlamNode>& nodes);
    void waitIdle();                                      // test helper: block until the queues are drained

protected:
    void placeRecognitionThread();
    std::vector<std::pair<std::string, std::string> > run(int mode, const std::vector<SlamNode>& nodes);

    GpuFeatureTransformationEstimator& est_;
    place_recognition::PlaceRecognizerConfig config_;
    std::thread thread_;
    std::mutex res_mutex_;
    std::condition_variable cv_;
    bool running_ = true, busy_ = false;
    std::deque<SlamNode> pr_queue_;
    std::vector<SlamNode> pr_add_queue_;
    std::vector<std::string> pr_remove_queue_;
    std::vector<std::pair<std::string, std::string> > potential_neighbors_;
    std::unordered_map<int32_t, std::string> id_of_handle_;
};
