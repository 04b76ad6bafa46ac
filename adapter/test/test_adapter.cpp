// test_adapter.cpp — drives the adapter the way GraphSlamNode drives the reference estimator
// (graph_slam/src/graph_slam_node.cpp:48 construction with a callback, :266/:284 estimateEdge per candidate,
// :779-829 newEdgeCallback).  Needs a B200; run by tests/test_adapter.py (-m gpu).
#include <place_recognition/gpu_lsh_set_recognizer.h>
#include <transformation_estimation/gpu_feature_transformation_estimator.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static double uni() { return (rnd() >> 11) * (1.0 / 9007199254740992.0); }

struct Pose { double R[9], t[3]; };
static Pose small_pose(double ang, double tr) {
    Pose p;
    double ax[3] = {uni() - .5, uni() - .5, uni() - .5};
    double n = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    for (double& a : ax) a /= n;
    double c = std::cos(ang), s = std::sin(ang), C = 1 - c;
    double R[9] = {c + ax[0] * ax[0] * C, ax[0] * ax[1] * C - ax[2] * s, ax[0] * ax[2] * C + ax[1] * s,
                   ax[1] * ax[0] * C + ax[2] * s, c + ax[1] * ax[1] * C, ax[1] * ax[2] * C - ax[0] * s,
                   ax[2] * ax[0] * C - ax[1] * s, ax[2] * ax[1] * C + ax[0] * s, c + ax[2] * ax[2] * C};
    for (int i = 0; i < 9; ++i) p.R[i] = R[i];
    for (int i = 0; i < 3; ++i) p.t[i] = (uni() - .5) * tr;
    return p;
}

// n landmarks shared by both nodes (+ n fresh each); node B observes them through pose G: p_b = G * x
static void make_pair_nodes(int n, const Pose& G, SlamNode& a, SlamNode& b, const std::string& ida, const std::string& idb,
                            const std::string& frame, int cols = 32, int ftype = graph_slam_msgs::Features::ORB) {
    FeatureDataPtr fa(new FeatureData()), fb(new FeatureData());
    for (FeatureDataPtr f : {fa, fb}) {
        f->feature_type_ = ftype;
        f->sensor_frame_ = frame;
        f->features_.create(2 * n, cols, CV_8U);
        f->feature_positions_.resize(3, 2 * n);
        f->valid_3d_.assign(2 * n, true);
    }
    for (int i = 0; i < 2 * n; ++i) {
        double z = 0.5 + 6.5 * uni(), x = (uni() * 640 - 319.5) * z / 525., y = (uni() * 480 - 239.5) * z / 525.;
        for (int k = 0; k < cols; ++k) fa->features_.at<unsigned char>(i, k) = (unsigned char)rnd();
        fa->feature_positions_(0, i) = x; fa->feature_positions_(1, i) = y; fa->feature_positions_(2, i) = z;
        if (i < n) {           // shared landmark: same descriptor with a few flipped bits, transformed position
            for (int k = 0; k < cols; ++k) {
                unsigned char m = (unsigned char)(rnd() & rnd() & rnd() & rnd());
                fb->features_.at<unsigned char>(i, k) = fa->features_.at<unsigned char>(i, k) ^ m;
            }
            fb->feature_positions_(0, i) = G.R[0] * x + G.R[1] * y + G.R[2] * z + G.t[0] + (uni() - .5) * 0.004;
            fb->feature_positions_(1, i) = G.R[3] * x + G.R[4] * y + G.R[5] * z + G.t[1] + (uni() - .5) * 0.004;
            fb->feature_positions_(2, i) = G.R[6] * x + G.R[7] * y + G.R[8] * z + G.t[2] + (uni() - .5) * 0.004;
        } else {
            for (int k = 0; k < cols; ++k) fb->features_.at<unsigned char>(i, k) = (unsigned char)rnd();
            double z2 = 0.5 + 6.5 * uni();
            fb->feature_positions_(0, i) = (uni() * 640 - 319.5) * z2 / 525.;
            fb->feature_positions_(1, i) = (uni() * 480 - 239.5) * z2 / 525.;
            fb->feature_positions_(2, i) = z2;
        }
        if (i % 9 == 0) { fb->valid_3d_[i] = false; fb->feature_positions_(2, i) = -1; }
    }
    a = SlamNode(); b = SlamNode();
    a.id_ = ida; b.id_ = idb;
    a.addSensorData(fa); b.addSensorData(fb);
}

#define CHECK(c) do { if (!(c)) { std::printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
    std::mutex m;
    std::map<std::string, SlamEdge> got;          // keyed by id_from|id_to
    // UZ_TEST_DEVICES="0,1,...": the same class on several GPUs (uz_group); default: device 0
    std::vector<int> devices;
    if (const char* dv = std::getenv("UZ_TEST_DEVICES")) {
        for (const char* p = dv; *p;) { devices.push_back(std::atoi(p)); while (*p && *p != ',') ++p; if (*p == ',') ++p; }
    }
    if (devices.empty()) devices.push_back(0);
    GpuFeatureTransformationEstimator est([&](SlamEdge e) { std::lock_guard<std::mutex> lk(m); got[e.id_from_ + "|" + e.id_to_] = e; }, devices);
    CHECK(est.devices() == (int)devices.size());
    transformation_estimation::FeatureLinkEstimationConfig cfg;
    cfg.ransac_threshold = 0.1; cfg.ransac_iteration = 100; cfg.ransac_break_percentage = 0.6;    // slam.yaml:35-36
    est.setConfig(cfg);

    // 1. queue interface: many pairs, callbacks for all of them, same answers as the direct call
    const int NP = 24;
    std::vector<SlamNode> A(NP), B(NP);
    std::vector<Pose> G(NP);
    for (int i = 0; i < NP; ++i) {
        G[i] = small_pose(0.3 * uni(), 1.0);
        make_pair_nodes(250, G[i], A[i], B[i], "a" + std::to_string(i), "b" + std::to_string(i), "/camera_rgb_optical_frame");
        est.estimateEdge(A[i], B[i]);
    }
    est.waitIdle();
    CHECK((int)got.size() == NP);
    CHECK(est.residentNodes() == (size_t)2 * NP);
    for (int i = 0; i < NP; ++i) {
        const SlamEdge& e = got["a" + std::to_string(i) + "|b" + std::to_string(i)];
        CHECK(e.matching_score_ >= 100);                                  // ~250 planted correspondences, 1/9 without depth
        CHECK(e.type_ == graph_slam_msgs::Edge::TYPE_3D_FULL);
        CHECK(e.sensor_from_ == "/camera_rgb_optical_frame");
        // transform_ maps to-frame (B) points into the from-frame (A): inverse of G
        double err = 0;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) err = std::fmax(err, std::fabs(e.transform_(r, c) - G[i].R[3 * c + r]));
        CHECK(err < 2e-3);
        CHECK(e.information_(0, 0) > 1.0 && std::fabs(e.information_(3, 3) - 100. * e.information_(0, 0)) < 1e-6 * e.information_(3, 3));
        SlamEdge d;
        CHECK(est.estimateEdgeDirect(A[i].sensor_data_, B[i].sensor_data_, d));
        CHECK(d.matching_score_ == e.matching_score_);
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) CHECK(d.transform_(r, c) == e.transform_(r, c));
    }

    // 1c. a callback that calls back into the estimator (forgetNode of an unknown id and estimateEdgeDirect both take
    //     gpuMutex()) while a burst of many small chunks is being worked off: every edge arrives, nothing deadlocks, and the
    //     answers are the ones of section 1
    {
        setenv("UZ_ADAPTER_CHUNK", "64", 1);
        std::mutex m2;
        std::map<std::string, SlamEdge> got2;
        GpuFeatureTransformationEstimator* self = nullptr;
        int direct_ok = 0;
        GpuFeatureTransformationEstimator est2([&](SlamEdge e) {
            self->forgetNode("no such node");
            if (e.id_from_ == "a0") { SlamEdge d; direct_ok += self->estimateEdgeDirect(A[1].sensor_data_, B[1].sensor_data_, d) ? 1 : 0; }
            std::lock_guard<std::mutex> lk(m2);
            got2[e.id_from_ + "|" + e.id_to_ + "#" + std::to_string(got2.size())] = e;
        }, devices);
        unsetenv("UZ_ADAPTER_CHUNK");
        self = &est2;
        est2.setConfig(cfg);
        const int reps = 40;                                   // 40 x 24 pairs = 15 chunks of 64
        for (int r = 0; r < reps; ++r)
            for (int i = 0; i < NP; ++i) est2.estimateEdge(A[i], B[i]);
        est2.waitIdle();
        CHECK((int)got2.size() == reps * NP);
        CHECK(direct_ok >= 1);
        for (auto& kv : got2) {
            const SlamEdge& e = kv.second;
            const SlamEdge& want = got[e.id_from_ + "|" + e.id_to_];
            CHECK(e.matching_score_ == want.matching_score_);
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) CHECK(e.transform_(r, c) == want.transform_(r, c));
        }
    }

    // 1b. BRISK keyframes (64-byte rows; cv::BRISK is FeatureExtractionCore's default, feature_extraction_core.cpp:46-49)
    //     through the same queue, and one BRISK-vs-ORB pair: never compared (cv::BFMatcher would throw), score 0
    {
        got.clear();
        std::vector<SlamNode> WA(6), WB(6);
        for (int i = 0; i < 6; ++i) {
            make_pair_nodes(200, G[i], WA[i], WB[i], "wa" + std::to_string(i), "wb" + std::to_string(i), "/camera_rgb_optical_frame",
                            64, graph_slam_msgs::Features::BRISK);
            est.estimateEdge(WA[i], WB[i]);
        }
        boost::dynamic_pointer_cast<FeatureData>(A[0].sensor_data_[0])->feature_type_ = graph_slam_msgs::Features::BRISK;
        SlamNode A0 = A[0];
        A0.id_ = "a0_as_brisk";
        est.estimateEdge(WA[0], A0);
        est.waitIdle();
        boost::dynamic_pointer_cast<FeatureData>(A[0].sensor_data_[0])->feature_type_ = graph_slam_msgs::Features::ORB;
        CHECK((int)got.size() == 7);
        for (int i = 0; i < 6; ++i) {
            const SlamEdge& e = got["wa" + std::to_string(i) + "|wb" + std::to_string(i)];
            CHECK(e.matching_score_ >= 80);
            double err = 0;
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) err = std::fmax(err, std::fabs(e.transform_(r, c) - G[i].R[3 * c + r]));
            CHECK(err < 2e-3);
        }
        CHECK(got["wa0|a0_as_brisk"].matching_score_ == 0.);
    }

    // 2. failure convention: different sensor frames -> no comparable pair -> score 0, callback still fires
    got.clear();
    SlamNode X, Y;
    make_pair_nodes(100, G[0], X, Y, "x", "y", "/cam0");
    boost::dynamic_pointer_cast<FeatureData>(Y.sensor_data_[0])->sensor_frame_ = "/cam1";
    est.estimateEdge(X, Y);
    est.waitIdle();
    CHECK(got.size() == 1 && got["x|y"].matching_score_ == 0. && got["x|y"].type_ == 0);
    SlamEdge fe;
    CHECK(!est.estimateEdgeImpl(X, Y, fe) && fe.id_from_ == "x" && fe.id_to_ == "y");

    // 2b. a node's sensor data changes under its id (late sensor arrival graph_slam_node.cpp:244, node merge :1010-1026): the
    //     reference copies the node on every estimateEdge and matches the CURRENT data; a cached device copy must follow
    {
        got.clear();
        SlamNode U, V, X2, Y2;
        make_pair_nodes(200, G[2], U, V, "u_new", "v", "/cam0");
        make_pair_nodes(100, G[3], X2, Y2, "x2", "y2", "/cam1");
        SlamNode Uold;
        Uold.id_ = "u";
        Uold.addSensorData(X2.sensor_data_[0]);                      // only a /cam1 camera so far: nothing comparable with v
        est.estimateEdge(Uold, V);
        est.waitIdle();
        CHECK(got["u|v"].matching_score_ == 0.);
        const size_t resident = est.residentNodes();
        SlamNode Unew = Uold;
        Unew.addSensorData(U.sensor_data_[0]);                       // the /cam0 camera arrives under the same id
        est.estimateEdge(Unew, V);
        est.waitIdle();
        CHECK(got["u|v"].matching_score_ >= 80);
        CHECK(est.residentNodes() == resident);                      // re-uploaded under the same handle, not added
        SlamEdge d;
        CHECK(est.estimateEdgeDirect(Unew.sensor_data_, V.sensor_data_, d));
        CHECK(d.matching_score_ == got["u|v"].matching_score_);
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) CHECK(d.transform_(r, c) == got["u|v"].transform_(r, c));
        CHECK(got["u|v"].sensor_from_ == "/cam0");
        est.estimateEdge(Uold, V);                                   // and back: the old copy of the node is handed in again
        est.waitIdle();
        CHECK(got["u|v"].matching_score_ == 0.);
    }

    // 2c. removed nodes give their device memory back (graph.removeNode, graph_slam_node.cpp:1050 -> forgetNode); resume
    //     re-adds every stored node in one bulk upload (GraphSlamNode::load, :875-888 -> loadNodes)
    {
        const int64_t bytes0 = est.storeBytes();
        const size_t resident0 = est.residentNodes();
        for (int k = 0; k < 12; ++k) {
            SlamNode P1, P2;
            SlamEdge e;
            make_pair_nodes(150 + 10 * k, G[k % NP], P1, P2, "tmp_a" + std::to_string(k), "tmp_b" + std::to_string(k), "/cam0");
            CHECK(est.estimateEdgeImpl(P1, P2, e) && e.matching_score_ >= 50);
            est.forgetNode(P1.id_); est.forgetNode(P2.id_);
        }
        CHECK(est.storeBytes() == bytes0 && est.residentNodes() == resident0);
        std::vector<SlamNode> stored;
        for (int k = 0; k < 5; ++k) {
            SlamNode P1, P2;
            make_pair_nodes(120, G[k], P1, P2, "ld_a" + std::to_string(k), "ld_b" + std::to_string(k), "/cam0");
            stored.push_back(P1); stored.push_back(P2);
        }
        CHECK(est.loadNodes(stored));
        CHECK(est.residentNodes() == resident0 + 10);
        SlamEdge e, d;
        CHECK(est.estimateEdgeImpl(stored[4], stored[5], e));
        CHECK(est.estimateEdgeDirect(stored[4].sensor_data_, stored[5].sensor_data_, d));
        CHECK(e.matching_score_ == d.matching_score_ && e.matching_score_ >= 50);
        CHECK(est.residentNodes() == resident0 + 10);
        for (const SlamNode& n : stored) est.forgetNode(n.id_);
        CHECK(est.storeBytes() == bytes0);
    }

    // 3. estimateSVD / consensus3D, the TransformationFilter call shape (transformation_filter.cpp:272,275)
    Eigen::MatrixXd P(3, 80), Q(3, 80);
    for (int i = 0; i < 80; ++i) {
        double x = uni() * 4, y = uni() * 4, z = uni() * 4;
        P(0, i) = x; P(1, i) = y; P(2, i) = z;
        Q(0, i) = G[1].R[0] * x + G[1].R[1] * y + G[1].R[2] * z + G[1].t[0];
        Q(1, i) = G[1].R[3] * x + G[1].R[4] * y + G[1].R[5] * z + G[1].t[1];
        Q(2, i) = G[1].R[6] * x + G[1].R[7] * y + G[1].R[8] * z + G[1].t[2];
        if (i % 4 == 0) Q(0, i) += 2.0;
    }
    Eigen::Isometry3d T;
    int consensus = 0;
    double mse = 0;
    est.estimateSVD(P, Q, T, consensus, mse, 0.3, 200, 1.0, false);
    CHECK(consensus == 60 && mse < 1e-3);
    Eigen::Array<bool, 1, Eigen::Dynamic> set;
    CHECK(est.consensus3D(P, Q, T, 0.3, set) == 60 && set.count() == 60 && !set[0] && set[1]);

    // 4. the steps either side of the path through the mirrored interfaces
    //    (a) place recognition: a revisit sequence; node B[i] was observed from the same place as A[i]
    {
        GpuLshSetRecognizer pr(est);
        place_recognition::PlaceRecognizerConfig pc;
        pc.k_nearest_neighbors = 5; pc.T = 2;                                  // global_slam.yaml:31,33
        pr.setConfig(pc);
        for (int i = 0; i < NP; ++i) { A[i].stamps_.assign(1, ros::Time()); A[i].stamps_[0].sec = 10.0 * i; }
        for (int i = 0; i < NP; ++i) { B[i].stamps_.assign(1, ros::Time()); B[i].stamps_[0].sec = 1000.0 + 10.0 * i; }
        for (int i = 0; i < NP; ++i) pr.addNode(A[i]);
        pr.waitIdle();
        CHECK(!pr.hasRecognizedPlaces());                                       // unrelated places: nothing recognised
        for (int i = 0; i < NP; ++i) pr.addNode(B[i]);
        pr.waitIdle();
        CHECK(pr.hasRecognizedPlaces());
        auto nb = pr.recognizedPlaces();
        CHECK((int)nb.size() == NP && !pr.hasRecognizedPlaces());
        std::map<std::string, std::string> from_of;
        for (auto& p : nb) from_of[p.second] = p.first;
        for (int i = 0; i < NP; ++i) CHECK(from_of["b" + std::to_string(i)] == "a" + std::to_string(i));
        // the recognised pairs are exactly what estimateEdge takes next (graph_slam_node.cpp:512-530)
        auto again = pr.searchPlace(B[3]);
        CHECK(again.empty());                                                   // checked_: already reported
        pr.removePlace("a5");
        pr.clear();
    }
    //    (b) acceptance gate and cluster RANSAC
    {
        std::vector<SlamEdge> es;
        for (auto& kv : got) es.push_back(kv.second);
        for (int i = 0; i < NP; ++i) es.push_back(SlamEdge());
        std::vector<char> acc;
        est.acceptEdges(es, 20., 1.5, 30., acc);
        CHECK(acc.size() == es.size());
        CHECK(!acc[0]);                                                          // the failed x|y edge: score 0
        for (size_t i = es.size() - NP; i < es.size(); ++i) CHECK(!acc[i]);      // default edges: score 0
        SlamEdge good;
        CHECK(est.estimateEdgeDirect(A[1].sensor_data_, B[1].sensor_data_, good));
        std::vector<SlamEdge> one(1, good);
        est.acceptEdges(one, 20., 1.5, 30., acc);
        CHECK(acc.size() == 1 && acc[0]);                                        // |t| <= 0.87 m, angle <= 17 deg, score >= 100
        est.acceptEdges(one, 20., 1e-6, 30., acc);
        CHECK(!acc[0]);
        std::vector<Eigen::MatrixXd> Ps(3, P), Qs(3, Q);
        std::vector<Eigen::Isometry3d> Ts; std::vector<int> cs; std::vector<double> ms; std::vector<std::vector<char> > sets;
        est.estimateSVDBatch(Ps, Qs, Ts, cs, ms, sets);
        CHECK(cs.size() == 3 && cs[0] == 60 && cs[2] == 60 && (int)sets[1].size() == 80 && !sets[1][0] && sets[1][1]);
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) CHECK(Ts[1](r, c) == T(r, c));
    }

    const size_t before = est.residentNodes();
    est.forgetNode("a0");
    CHECK(est.residentNodes() == before - 1);
    std::printf("ADAPTER OK: %d queued pairs on %d device(s), callbacks delivered, direct/batch identical\n", NP, est.devices());
    return 0;
}
