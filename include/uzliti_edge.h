/* uzliti_edge.h — C-ABI of the B200 feature-edge estimation path.
 *
 * This is the drop-in boundary for ONE path of jan-frost/uzliti_slam: what
 *   transformation_estimation/include/transformation_estimation/feature_transformation_estimator.h:33-58
 * exposes (estimateEdgeImpl / estimateEdgeDirect / estimateSVD / consensus3D / setConfig) and what
 *   transformation_estimation/include/transformation_estimation/transformation_estimator.h:45-67
 * queues (estimateEdge), restated as plain C entry points: borrowed read-only host pointers in,
 * caller-allocated result arrays out, no C++/torch/ROS types.  The C++ adapter in adapter/ maps
 * SlamNode / FeatureData / SlamEdge onto these calls; INTEGRATION.md shows the reference-side binding.
 *
 * All functions return UZ_OK (0) or a negative uz_status.  A per-pair failure (no comparable camera
 * pair, < 3 depth-valid matches) is NOT an error: it is reported in-band as result.ok == 0 with
 * consensus == 0, mirroring transformation_estimator.cpp:53-55 (matching_score_ = 0, callback still fires).
 * There is no CPU fallback: every compute entry point fails with UZ_ERR_CUDA when no sm_100 device
 * is usable.
 */
#ifndef UZLITI_EDGE_H
#define UZLITI_EDGE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UZ_DESC_BYTES 32          /* ORB/BRIEF-256: feature_extraction/external/aorb/aorb.h:54 (kBytes = 32) */
#define UZ_MAX_DESC_BYTES 64      /* BRISK/FREAK-512: cv::BRISK / cv::FREAK rows, feature_extraction_core.cpp:69-77 */
#define UZ_MAX_FEATURES 4096      /* per camera (solve kernel keeps a pair on-chip); reference front-end caps at 400 (cfg/FeatureExtraction.cfg:11) */
#define UZ_MAX_ITERATIONS 4096    /* cfg/FeatureLinkEstimation.cfg:11 allows 1..1000 */

typedef enum {
    UZ_OK = 0,
    UZ_ERR_INVALID = -1,      /* bad argument (null pointer, size out of range, unknown keyframe handle) */
    UZ_ERR_CUDA = -2,         /* CUDA runtime error or no usable device; see uz_last_error() */
    UZ_ERR_NOMEM = -3,
    UZ_ERR_UNSUPPORTED = -4   /* e.g. descriptor width other than 32 or 64 bytes */
} uz_status;

/* graph_slam_msgs/msg/Features.msg:1-6 */
enum { UZ_FEATURE_BRIEF = 1, UZ_FEATURE_ORB = 2, UZ_FEATURE_BRISK = 3, UZ_FEATURE_FREAK = 4,
       UZ_FEATURE_SURF = 5, UZ_FEATURE_SIFT = 6 };

typedef struct uz_context uz_context;

/* Parameters of the path.  Mirrors transformation_estimation/cfg/FeatureLinkEstimation.cfg:9-13 plus the
 * constants hard-coded in feature_transformation_estimator.cpp (:47 min 7 keypoints, :67 ratio 0.99,
 * :118 min 3 matches).  uz_default_params() fills the reference's production values
 * (iti_slam_launch/yaml/slam.yaml:35-36: threshold 0.1, 100 iterations; cfg default break 0.6). */
typedef struct {
    double  ransac_threshold;      /* config_.ransac_threshold  (consensus3D maxError, metres)   */
    double  break_percentage;      /* config_.ransac_break_percentage                            */
    int32_t ransac_iterations;     /* config_.ransac_iteration                                   */
    int32_t do_prosac;             /* estimateSVD(..., do_prosac): 1 = growing-prefix shuffle     */
    int32_t ratio_num, ratio_den;  /* keep best iff ratio_den*d0 < ratio_num*d1   (99/100 == :67) */
    int32_t min_keypoints;         /* :47 (7)                                                    */
    int32_t cross_check;           /* opt-in (reference: none, 0): keep a ratio survivor (q,t) only if q is also the
                                      nearest query row of t, lowest index on ties == cv::BFMatcher(crossCheck=true);
                                      the match kernel tracks the per-train-row minima in the same pass (+10-15 %)  */
} uz_params;

/* One FeatureData (graph_slam_common/include/graph_slam_common/sensor_data.h:49-70) as borrowed POD. */
typedef struct {
    const uint8_t* descriptors;    /* features_: n rows x desc_bytes, row stride desc_stride (cv::Mat CV_8U) */
    const double*  positions;      /* feature_positions_: 3 x n column-major doubles (Eigen::MatrixXd)      */
    const uint8_t* valid_3d;       /* valid_3d_: n bytes, non-zero = has depth                               */
    int32_t n;
    int32_t desc_stride;           /* bytes between descriptor rows (>= desc_bytes; features_.step)          */
    int32_t desc_bytes;            /* features_.cols: 32 (ORB, BRIEF) or 64 (BRISK, FREAK); 0 means 32.  Cameras of
                                      different widths are never compared (cv::BFMatcher would throw): such a camera
                                      pair is skipped like one with different feature_type_                   */
    int32_t feature_type;          /* feature_type_                                                          */
    int32_t sensor_frame;          /* sensor_frame_ interned by the caller (equal strings <=> equal tags)    */
} uz_features;

/* The SlamEdge fields estimateEdgeDirect fills (feature_transformation_estimator.cpp:147-156). */
typedef struct {
    int32_t ok;                    /* estimateEdgeImpl's bool                                            */
    int32_t cam_from, cam_to;      /* index of the winning FeatureData in from/to sensor lists; -1 none  */
    int32_t n_ratio_matches;       /* score of the winning camera pair (:78)                             */
    int32_t n_matches;             /* M: matches left after the valid_3d filter (:115)                   */
    int32_t consensus;             /* matching_score_ (:155); 0 when !ok                                 */
    int32_t best_iteration;        /* index of the winning hypothesis, -1 if none                        */
    int32_t iterations_run;        /* hypotheses evaluated before the early break (:239)                 */
    double  mse;                   /* mean inlier residual norm (:285-290)                               */
    double  info_scale;            /* information_ = I6 * info_scale, rotation block x100 more (:133-137)*/
    double  T[16];                 /* transform_: row-major 4x4, maps to-frame points into the from-frame */
} uz_edge_result;

/* ---- context ------------------------------------------------------------------------------- */
uz_status uz_create(int32_t device, uz_context** out);
void      uz_destroy(uz_context* ctx);
const char* uz_last_error(const uz_context* ctx);          /* never NULL */
void      uz_default_params(uz_params* p);
/* setConfig (feature_transformation_estimator.cpp:350-353).  Snapshotted per batch. */
uz_status uz_set_params(uz_context* ctx, const uz_params* p);
uz_status uz_get_params(const uz_context* ctx, uz_params* p);
/* Run all work of this context on an existing CUDA stream (cudaStream_t as void*); NULL = own stream. */
uz_status uz_set_stream(uz_context* ctx, void* cuda_stream);

/* ---- device-resident keyframe store --------------------------------------------------------- */
/* Adds one SlamNode's FEATURE sensor data (slam_node.h:93 sensor_data_) and returns a dense handle.
 * Host buffers are copied before return.  Replaces what estimateEdge's by-value SlamNode copy carried
 * (transformation_estimator.cpp:39). */
uz_status uz_store_add(uz_context* ctx, const uz_features* cams, int32_t n_cams, int32_t* handle_out);
uz_status uz_store_add_bulk(uz_context* ctx, const uz_features* cams, const int32_t* cams_per_keyframe,
                            int32_t n_keyframes, int32_t* handles_out);
/* The node's sensor data changed under a fixed id (late sensor arrivals graph_slam_node.cpp:244, node merge :1010-1026; the
 * reference copies the whole SlamNode on every estimateEdge, transformation_estimator.cpp:39, so it always sees the current
 * data): new cameras under the SAME handle.  A place that was built from the old rows keeps reading them, as the reference's
 * recogniser keeps the copy it was given. */
uz_status uz_store_replace(uz_context* ctx, int32_t handle, const uz_features* cams, int32_t n_cams);
/* Gives the keyframe's device memory back to the store (uz_store_bytes drops) and recycles the handle; a place and the
 * checked_ pairs of that handle are forgotten with it. */
uz_status uz_store_remove(uz_context* ctx, int32_t handle);
uz_status uz_store_clear(uz_context* ctx);
int32_t   uz_store_size(const uz_context* ctx);            /* live keyframes */
int64_t   uz_store_bytes(const uz_context* ctx);           /* device bytes held by the store */

/* ---- ingestion on the device ---------------------------------------------- */
/* Pinhole model + depth gate of FeatureExtractionCore::extract3dFeatures
 * (feature_extraction/src/feature_extraction_core.cpp:254-295; max_depth 0 = no limit, slam.yaml:7 uses 7 m). */
typedef struct {
    double  fx, fy, cx, cy;
    double  max_depth;
    int32_t width, height;         /* depth image size in pixels */
} uz_camera;
/* extract3dFeatures on the device: (u, v) clamped into the image, depth read as float; valid iff depth != 0, not NaN
 * and <= max_depth; invalid keypoints get (0, 0, -1).  depth: height rows of float32, depth_stride_bytes apart.
 * reverse != 0 writes feature i to slot n-1-i, the order the reference produces (it walks its input back to front). */
uz_status uz_backproject(uz_context* ctx, const int32_t* u, const int32_t* v, int32_t n, const float* depth,
                         int32_t depth_stride_bytes, const uz_camera* cam, int32_t reverse, double* positions_out,
                         uint8_t* valid_out);
/* Same, but the keyframe goes straight into the store (positions never exist on the host): descriptors n x desc_bytes
 * (32 or 64; row stride desc_stride) + pixels + the depth image in, a handle out. */
uz_status uz_store_add_rgbd(uz_context* ctx, const uint8_t* descriptors, int32_t desc_stride, int32_t desc_bytes,
                            const int32_t* u, const int32_t* v,
                            int32_t n, const float* depth, int32_t depth_stride_bytes, const uz_camera* cam, int32_t feature_type,
                            int32_t sensor_frame, int32_t reverse, int32_t* handle_out);
/* FeatureData::fromMsg (graph_slam_common/src/sensor_data.cpp:124-171) applied on the device to the ROS1-serialised
 * graph_slam_msgs/Feature[] field (graph_slam_msgs/msg/Feature.msg: int32 u, int32 v, bool is_3d, float32
 * keypoint_strength, float32[] descriptor, geometry_msgs/Point keypoint_position; little endian, unpadded): blob =
 * uint32 count followed by the elements, as it sits in a SensorData message or a RosbagStorage record
 * (graph_slam_common/src/rosbag_storage.cpp:135-211).  Descriptor floats are narrowed to bytes as the reference does. */
uz_status uz_store_add_wire(uz_context* ctx, const uint8_t* features_blob, size_t blob_bytes, int32_t feature_type,
                            int32_t sensor_frame, int32_t* handle_out);
/* Resume (GraphSlamNode::load, graph_slam_node.cpp:875-888: RosbagStorage::loadGraph reads every stored Node,
 * rosbag_storage.cpp:135-211, and every node is re-added): many serialised Feature[] fields in ONE decode launch.  blobs are
 * concatenated per keyframe (cams_per_keyframe[k] of them); each blob is what uz_store_add_wire takes. */
uz_status uz_store_add_wire_bulk(uz_context* ctx, const uint8_t* const* blobs, const size_t* blob_bytes,
                                 const int32_t* feature_types, const int32_t* sensor_frames,
                                 const int32_t* cams_per_keyframe, int32_t n_keyframes, int32_t* handles_out);
/* Where a sensor's fields sit inside a ROS1-serialised message (byte offsets from the start of the message walked). */
typedef struct {
    int32_t sensor_type;           /* SensorData.msg: 1 = SENSOR_TYPE_FEATURE                                        */
    int32_t descriptor_type;       /* Features.msg descriptor_type (feature_type_)                                   */
    int32_t n_features;
    int32_t sensor_frame_len;
    size_t  sensor_frame_offset;   /* characters of SensorData.sensor_frame (not NUL terminated)                     */
    size_t  displacement_offset;   /* geometry_msgs/Pose: 7 float64 (position xyz, orientation xyzw)                 */
    size_t  features_offset;       /* the Feature[] field: uint32 count + elements = the blob of uz_store_add_wire*  */
    size_t  features_bytes;
} uz_wire_sensor;
/* Host-only walks (no device, no context): one graph_slam_msgs/SensorData message (what /sensor_data carries,
 * graph_slam_msgs/msg/SensorData.msg) ... */
uz_status uz_wire_walk_sensor_data(const uint8_t* msg, size_t bytes, uz_wire_sensor* sensor_out, size_t* consumed_out);
/* ... and one graph_slam_msgs/Node message (what a RosbagStorage node record holds, graph_slam_msgs/msg/Node.msg): every
 * SensorData of its SensorDataArray in order (n_sensors_out may exceed capacity), and where the node's id string sits. */
uz_status uz_wire_walk_node(const uint8_t* msg, size_t bytes, uz_wire_sensor* sensors_out, int32_t capacity,
                            int32_t* n_sensors_out, size_t* id_offset_out, int32_t* id_len_out);
/* The decode alone, results back on the host (uv_out optional: n x 2 int32).  The descriptor width is what the
 * elements carry (32 or 64, all equal): descriptors_out needs capacity x UZ_MAX_DESC_BYTES bytes and receives n packed
 * rows of *desc_bytes_out bytes. */
uz_status uz_wire_decode(uz_context* ctx, const uint8_t* features_blob, size_t blob_bytes, int32_t capacity, int32_t* n_out,
                         int32_t* desc_bytes_out, uint8_t* descriptors_out, double* positions_out, uint8_t* valid_out,
                         int32_t* uv_out);
/* FeatureData::toMsg (graph_slam_common/src/sensor_data.cpp:77-122) for one stored camera: the serialised
 * graph_slam_msgs/Feature[] field (uint32 count + elements) as uz_store_add_wire / uz_wire_decode read it - what a
 * SensorData message or a RosbagStorage record carries.  uv (optional, n x 2 int32) supplies feature_positions_2d_, which
 * the store does not keep (zeros if NULL); keypoint_strength is -1 as in the reference.  *bytes_out = bytes written. */
uz_status uz_wire_encode(uz_context* ctx, int32_t handle, int32_t cam, const int32_t* uv, uint8_t* blob_out, size_t capacity,
                         size_t* bytes_out);
/* Read one stored camera back (descriptors n x *desc_bytes_out as given, positions 3 x n, valid n); any output may be
 * NULL; descriptors_out needs capacity x UZ_MAX_DESC_BYTES bytes. */
uz_status uz_store_read(uz_context* ctx, int32_t handle, int32_t cam, int32_t capacity, int32_t* n_out,
                        int32_t* desc_bytes_out, uint8_t* descriptors_out, double* positions_out, uint8_t* valid_out);

/* ---- stage entry points (parity + direct callers) -------------------------------------------- */
/* cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, 2) (feature_transformation_estimator.cpp:38,58).
 * desc_bytes: 32 or 64 (both matrices).  idx/dist: nq x 2 int32, ordered by (distance, trainIdx); missing
 * neighbours (nt < 2) are -1. */
uz_status uz_match_knn2(uz_context* ctx, int32_t desc_bytes, const uint8_t* query, int32_t nq, int32_t q_stride,
                        const uint8_t* train, int32_t nt, int32_t t_stride,
                        int32_t* idx_out, int32_t* dist_out);

/* estimateSVD (feature_transformation_estimator.cpp:178-184; second caller transformation_filter.cpp:272).
 * P, Q: 3 x M column-major.  samples: optional iterations x 3 int32 sample list (NULL = the built-in
 * replay of std::random_shuffle over rand() seed 1).  inlier_mask (optional): M bytes. */
uz_status uz_estimate_svd(uz_context* ctx, const double* P, const double* Q, int32_t M,
                          double max_error, int32_t iterations, double break_percentage, int32_t do_prosac,
                          const int32_t* samples, double* T16_out, int32_t* consensus_out, double* mse_out,
                          uint8_t* inlier_mask_out, int32_t* best_iteration_out, int32_t* iterations_run_out);

/* consensus3D (feature_transformation_estimator.cpp:337-347). */
uz_status uz_consensus3d(uz_context* ctx, const double* P, const double* Q, int32_t M, const double* T16,
                         double thresh, uint8_t* set_out, int32_t* count_out);

/* The sample list the built-in generator produces for (M, iterations, do_prosac): iterations x 3. */
uz_status uz_sample_list(uz_context* ctx, int32_t M, int32_t iterations, int32_t do_prosac, int32_t* out);

/* ---- the batched path ------------------------------------------------------------------------ */
/* estimateEdge x n_pairs on keyframes already in the store (handles from uz_store_add).
 * results: n_pairs records in HOST memory.  Blocks until done. */
uz_status uz_estimate_edges(uz_context* ctx, const int32_t* from_handles, const int32_t* to_handles,
                            int32_t n_pairs, uz_edge_result* results);

/* Same, results written to DEVICE memory (n_pairs records), asynchronous on the context's stream; the
 * caller synchronises.  This is the form the multi-GPU harness gathers over NCCL. */
uz_status uz_estimate_edges_device(uz_context* ctx, const int32_t* from_handles, const int32_t* to_handles,
                                   int32_t n_pairs, void* results_device);

/* estimateEdgeDirect x n_pairs straight from host FeatureData (feature_transformation_estimator.cpp:32):
 * from_cams/to_cams are concatenated per pair, n_from[i]/n_to[i] cameras each.  Uploads, runs, downloads. */
uz_status uz_estimate_edges_host(uz_context* ctx, const uz_features* from_cams, const int32_t* n_from,
                                 const uz_features* to_cams, const int32_t* n_to, int32_t n_pairs,
                                 uz_edge_result* results);

/* Parity taps for the LAST uz_estimate_edges / uz_estimate_edges_host call (valid until the next call on this context):
 * the sorted final_matches (:114) as (queryIdx, trainIdx, distance) triples and the final inlier mask.
 * Enable with uz_set_debug(ctx, 1) BEFORE the call; capacity is in matches. Returns M through n_out. */
uz_status uz_set_debug(uz_context* ctx, int32_t enable);
uz_status uz_debug_pair(uz_context* ctx, int32_t pair_index, int32_t* matches_out, uint8_t* inlier_mask_out,
                        int32_t capacity, int32_t* n_out);

/* Consensus count of every hypothesis the last call evaluated for that pair (iteration order; -1 = not
 * evaluated because of the early break). */
uz_status uz_debug_counts(uz_context* ctx, int32_t pair_index, int32_t* counts_out, int32_t capacity,
                          int32_t* n_out);

/* clock64() of the pair's solve CTA at its 8 phase boundaries (start, keys built, sorted, gathered,
 * hypotheses solved, winner known, refit done, end); profiling tap, debug mode only. */
uz_status uz_debug_phases(uz_context* ctx, int32_t pair_index, int64_t* clocks8_out);

/* ---- after the path -------------------------------------------------------------- */
/* TransformationFilter::calcValidEdges (transformation_estimation/src/transformation_filter.cpp:216-285): one
 * estimateSVD(P, Q, T, consensus, mse, 0.3, 200, 1.0, do_prosac=false) + consensus3D per edge cluster, all clusters in
 * one launch.  P, Q: 3 x offsets[n_problems] column-major, problem b owns columns [offsets[b], offsets[b+1]).
 * T16_out: n_problems x 16; inlier_mask_out (optional): offsets[n_problems] bytes, the consensus3D set of the final T.
 * Problems with < 3 points return T = I, consensus 0 (the reference would index out of range). */
uz_status uz_estimate_svd_batch(uz_context* ctx, const double* P, const double* Q, const int32_t* offsets, int32_t n_problems,
                                double max_error, int32_t iterations, double break_percentage, int32_t do_prosac,
                                double* T16_out, int32_t* consensus_out, double* mse_out, uint8_t* inlier_mask_out);

/* GraphSlamNode::newEdgeCallback's numeric gate (graph_slam/src/graph_slam_node.cpp:798-804): an edge is linked only if
 * matching_score_ >= min_matching_score, |translation| <= max_edge_distance_T (m) and the rotation angle
 * (Eigen::AngleAxisd(transform_.linear()).angle(), degrees) <= max_edge_distance_R.  The graph-state checks around it
 * (isMerged, existsEdge, checkEdgeHeuristic) stay with the graph. */
typedef struct {
    double min_matching_score;     /* graph_slam/cfg/GraphSlam.cfg: min_matching_score   */
    double max_edge_distance_T;    /* max_edge_distance_T                                */
    double max_edge_distance_R;    /* max_edge_distance_R (degrees)                      */
} uz_gate_params;
void      uz_default_gate_params(uz_gate_params* g);      /* iti_slam_launch/yaml/slam.yaml:25-27: 1.5 m, 30 deg, score 20 */
uz_status uz_gate_edges(uz_context* ctx, const uz_edge_result* results, int32_t n, const uz_gate_params* gate,
                        uint8_t* accept_out, double* translation_norm_out, double* rotation_deg_out);
/* Same on records that are still on the device (straight behind uz_estimate_edges_device); asynchronous. */
uz_status uz_gate_edges_device(uz_context* ctx, const void* results_device, int32_t n, const uz_gate_params* gate,
                               void* accept_device, void* translation_norm_device, void* rotation_deg_device);

/* ---- candidate generation (the step before the path) --------------------------- */
/* place_recognition: LshSetRecognizer (place_recognition/src/lsh_set_recognizer.cpp:46-94,96-165,188-305: 8 hash tables
 * over descriptor bytes [4k,4k+4), one vote per bucket entry) behind PlaceRecognizer's filters
 * (place_recognition/src/place_recognizer.cpp:64-118,154-190: live place, |dt| > 5 s, first k, checked_ set), on the
 * device over keyframes already in the store.  A place's id IS its store handle; stamps are ros::Time in nanoseconds.
 * Results are (from = recognised older keyframe, to = new keyframe) handle pairs, exactly the pairs the reference feeds
 * to estimateEdge (graph_slam_node.cpp:512-530), so they can go straight into uz_estimate_edges. */
typedef struct {
    double  T;                     /* cfg/PlaceRecognizer.cfg:12; similarity = votes / 8 tables >= T      */
    int32_t k_nearest_neighbors;   /* cfg/PlaceRecognizer.cfg:10                                          */
    int32_t min_rows;              /* 150: only keyframes with more descriptors are inserted (:66,:108)   */
    int32_t min_key_bits;          /* 12 = 3*key_width: matchAndAdd skips keys with popcount <= this (:239) */
    int64_t min_gap_ns;            /* 5 s: neighbours closer in time are dropped (place_recognizer.cpp:94) */
} uz_place_params;
void      uz_default_place_params(uz_place_params* p);     /* iti_slam_launch/yaml/slam.yaml:45-48: k = 20, T = 2 */
uz_status uz_places_set_params(uz_context* ctx, const uz_place_params* p);
/* PlaceRecognizer::clear (place_recognizer.cpp:49-62); also forgets the checked_ pairs (handles are recycled). */
uz_status uz_places_clear(uz_context* ctx);
/* addNode + searchAndAddPlace for handles[0..n) IN ORDER (keyframe i sees every earlier place, including
 * handles[0..i)), as one batch.  pairs_out: capacity x 2 int32 (from, to); *n_pairs_out = pairs found (may exceed
 * capacity: only the first `capacity` are written).  pair_owner_out (optional, capacity): index i of the keyframe. */
uz_status uz_places_search_and_add(uz_context* ctx, const int32_t* handles, const int64_t* stamps_ns, int32_t n,
                                   int32_t* pairs_out, int32_t capacity, int32_t* n_pairs_out);
/* addPlace (place_recognizer.cpp:120-147; resume path graph_slam_node.cpp:148-151): insert without searching,
 * no popcount filter on the keys. */
uz_status uz_places_add(uz_context* ctx, const int32_t* handles, const int64_t* stamps_ns, int32_t n);
/* searchPlace (place_recognizer.cpp:154-190): query without inserting; each query sees every place. */
uz_status uz_places_search(uz_context* ctx, const int32_t* handles, const int64_t* stamps_ns, int32_t n,
                           int32_t* pairs_out, int32_t capacity, int32_t* n_pairs_out);
/* removePlace (place_recognizer.cpp:198-206). */
uz_status uz_places_remove(uz_context* ctx, int32_t handle);
int32_t   uz_places_count(const uz_context* ctx);          /* place_count_ (removed places included) */
/* Parity tap: FastLshSet::match votes of one stored camera against every place (votes_out: uz_places_count ints);
 * filtered != 0 applies matchAndAdd's popcount filter to the query keys. */
uz_status uz_places_votes(uz_context* ctx, int32_t handle, int32_t cam, int32_t filtered, int32_t* votes_out,
                          int32_t capacity);
/* Device time (ms) of the insert / vote / select kernels of the last uz_places_* call and its bucket-entry visits. */
uz_status uz_places_last_timing(uz_context* ctx, double* insert_ms, double* vote_ms, double* select_ms);

/* ---- several GPUs of one box (SURVEY.md 8b(1) "device list", 8e) -------------------------------------------- */
/* The reference owns one worker thread per estimator (transformation_estimator.cpp:26) and handles one pair per tick; pairs
 * are independent, so a group owns one context per device in ONE process (one host worker thread each):
 *   - the keyframe store is replicated: uz_group_store_add* uploads once over PCIe to devices[0]; every other device pulls the
 *     keyframe from that device's HBM over NVLink (peer memory) and derives its layouts locally; handles are identical
 *     on all devices;
 *   - uz_group_estimate_edges* cuts the pair list into contiguous shards (rank r of n takes [lo, hi) with
 *     lo = r * (N / n) + min(r, N % n)); records land in ONE result array at the pair's batch-wide index: in devices[0]'s
 *     memory every device's solve kernel writes them itself through a peer-mapped pointer over NVLink (the gather is fused
 *     into the solve), on the host every device delivers its shard over its own PCIe link.  No collective runs.
 * Results are byte-identical to uz_estimate_edges on one device.  Devices need peer access to devices[0]. */
typedef struct uz_group uz_group;
uz_status uz_group_create(const int32_t* devices, int32_t n_devices, uz_group** out);
void      uz_group_destroy(uz_group* g);
const char* uz_group_last_error(const uz_group* g);           /* never NULL; NULL group: why the last create failed */
int32_t   uz_group_size(const uz_group* g);
/* Context of one device of the group (borrowed): stage entry points, parity taps and place recognition run on rank 0. */
uz_context* uz_group_context(uz_group* g, int32_t rank);
uz_status uz_group_set_params(uz_group* g, const uz_params* p);
uz_status uz_group_store_add(uz_group* g, const uz_features* cams, int32_t n_cams, int32_t* handle_out);
uz_status uz_group_store_add_bulk(uz_group* g, const uz_features* cams, const int32_t* cams_per_keyframe,
                                  int32_t n_keyframes, int32_t* handles_out);
uz_status uz_group_store_replace(uz_group* g, int32_t handle, const uz_features* cams, int32_t n_cams);
uz_status uz_group_store_remove(uz_group* g, int32_t handle);
uz_status uz_group_store_clear(uz_group* g);
int32_t   uz_group_store_size(const uz_group* g);
/* estimateEdge x n_pairs over all devices; results: n_pairs records in HOST memory, pair order.  Blocks until done. */
uz_status uz_group_estimate_edges(uz_group* g, const int32_t* from_handles, const int32_t* to_handles,
                                  int32_t n_pairs, uz_edge_result* results);
/* Same, records land in devices[0]'s memory (n_pairs records), e.g. for uz_gate_edges_device.  Blocks until done. */
/* The same in two halves, for a host that has work of its own to do meanwhile (the adapter prepares the next chunk of its
 * queue): _begin hands the batch to the group's device workers and returns, _end waits for them and returns the first failure.
 * from_handles, to_handles and results must stay valid until _end; no other call on the group or its contexts in between
 * (they answer UZ_ERR_INVALID).  _end without a batch in flight is a no-op. */
uz_status uz_group_estimate_edges_begin(uz_group* g, const int32_t* from_handles, const int32_t* to_handles,
                                        int32_t n_pairs, uz_edge_result* results);
uz_status uz_group_estimate_edges_end(uz_group* g);
uz_status uz_group_estimate_edges_device(uz_group* g, const int32_t* from_handles, const int32_t* to_handles,
                                         int32_t n_pairs, void* results_on_first_device);
/* Host results only.  1 (default): records stay local and every device copies its shard into the pinned result array over its
 * own PCIe link.  0: the solve kernels write through a host-mapped pointer (the measured alternative: slower).  Device
 * results (uz_group_estimate_edges_device) are always written by the solve kernels through the peer-mapped pointer. */
uz_status uz_group_set_gather(uz_group* g, int32_t mode);
/* Device time (ms, CUDA events per device) of the last uz_group_estimate_edges* call, one value per device. */
uz_status uz_group_last_timing(const uz_group* g, double* ms_per_device, int32_t capacity);

/* ---- execution form ------------------------------------------------------------------------------ */
/* How the solve (K2..K5) of a large batch is scheduled.  ctas_per_sm > 0 (default 1): a persistent solve grid of that
 * many CTAs per SM runs BESIDE the match kernel on a high-priority stream and consumes pairs as the match kernel
 * publishes them; it never blocks the match kernel and falls back to a launch behind it if the two cannot share the
 * device.  0: one solve CTA per pair, launched behind the match kernel.  Results are identical in both forms.  (The
 * reference has no such knob: its estimator owns one worker thread, transformation_estimator.cpp:26.) */
uz_status uz_set_stream_solve(uz_context* ctx, int32_t ctas_per_sm);

/* ---- introspection for the bench harness ----------------------------------------------------- */
/* Kernel launches issued by this context since creation (the bench's gpu_launches claim). */
int64_t   uz_launch_count(const uz_context* ctx);
/* Device time (ms, CUDA events on the context stream) of the matching / solve kernels accumulated since
 * the last uz_reset_timers(); timing is off unless enabled (events add launch overhead). */
uz_status uz_enable_timers(uz_context* ctx, int32_t enable);
uz_status uz_reset_timers(uz_context* ctx);
uz_status uz_get_timers(uz_context* ctx, double* match_ms, double* solve_ms, int64_t* match_launches,
                        int64_t* solve_launches, int64_t* descriptor_compares);
/* Integer-pipe microbenchmark used for the roofline denominator: op 0 = POPC.b32, 1 = LOP3.b32,
 * 2 = IMAD.u32, 3 = VIMNMX.u32, 4 = the match kernel's own compare mix.  Returns giga-ops/s. */
uz_status uz_microbench(uz_context* ctx, int32_t op, double* gops_out);

const char* uz_version(void);

#ifdef __cplusplus
}
#endif
#endif /* UZLITI_EDGE_H */
