// bench_adapter.cpp - the benchmark's keyframe map (include/uz_synth.h: the generator bench.py uses, same bytes) driven
// through the adapter exactly as GraphSlamNode drives the reference estimator: nodes with FeatureData in pageable cv::Mat /
// Eigen memory, one estimateEdge(from, to) per candidate pair (graph_slam_node.cpp:266,284), edges delivered by callback.
// Prints one JSON line.  usage: bench_adapter [n_keyframes=10000] [max_pairs=25000] [repeats=3]; UZ_TEST_DEVICES="0,1,.."
#include <transformation_estimation/gpu_feature_transformation_estimator.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "uz_synth.h"

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    const int n_kf = argc > 1 ? std::atoi(argv[1]) : 10000;
    const long max_pairs = argc > 2 ? std::atol(argv[2]) : 25000;
    const int repeats = argc > 3 ? std::atoi(argv[3]) : 3;
    std::vector<int> devices;
    if (const char* dv = std::getenv("UZ_TEST_DEVICES"))
        for (const char* p = dv; *p;) { devices.push_back(std::atoi(p)); while (*p && *p != ',') ++p; if (*p == ',') ++p; }
    if (devices.empty()) devices.push_back(0);

    // 1. the map, generated here (8 host threads; keyframes are independent streams)
    uz_synth_cfg cfg = uz_synth_c4(n_kf);
    const size_t N = (size_t)cfg.n_features;
    std::vector<uint8_t> desc((size_t)n_kf * N * 32), valid((size_t)n_kf * N);
    std::vector<double> pos((size_t)n_kf * N * 3);
    std::vector<int32_t> pairs((size_t)n_kf * cfg.k_candidates * 2);
    double t0 = now_s();
    {
        std::vector<std::thread> th;
        std::atomic<int> bad{0};
        for (int t = 0; t < 8; ++t)
            th.emplace_back([&, t] {
                for (int i = t; i < n_kf; i += 8)
                    if (uz_synth_keyframe(&cfg, i, desc.data() + (size_t)i * N * 32, pos.data() + (size_t)i * N * 3, valid.data() + (size_t)i * N)) bad = 1;
            });
        for (auto& x : th) x.join();
        if (bad) { std::printf("{\"error\": \"generator failed\"}\n"); return 1; }
    }
    long n_pairs = 0;
    for (int i = 0; i < n_kf; ++i) n_pairs += uz_synth_candidates(&cfg, i, pairs.data() + 2 * n_pairs);
    if (n_pairs > max_pairs) n_pairs = max_pairs;
    const double gen_s = now_s() - t0;
    const unsigned long long ck_desc = uz_synth_checksum(desc.data(), desc.size()), ck_pos = uz_synth_checksum(pos.data(), pos.size() * 8);

    // 2. nodes as the SLAM graph holds them
    std::vector<SlamNode> nodes((size_t)n_kf);
    for (int i = 0; i < n_kf; ++i) {
        FeatureDataPtr f(new FeatureData());
        f->feature_type_ = graph_slam_msgs::Features::ORB;
        f->sensor_frame_ = "/camera_rgb_optical_frame";
        f->features_.create((int)N, 32, CV_8U);
        std::memcpy(f->features_.data, desc.data() + (size_t)i * N * 32, N * 32);
        f->feature_positions_.resize(3, (int)N);
        std::memcpy(f->feature_positions_.data(), pos.data() + (size_t)i * N * 3, N * 24);
        f->valid_3d_.resize(N);
        for (size_t k = 0; k < N; ++k) f->valid_3d_[k] = valid[(size_t)i * N + k] != 0;
        nodes[i].id_ = "n" + std::to_string(i);
        nodes[i].addSensorData(f);
    }

    transformation_estimation::FeatureLinkEstimationConfig ec;
    ec.ransac_threshold = 0.1; ec.ransac_iteration = 100; ec.ransac_break_percentage = 0.6;    // slam.yaml:35-36

    std::atomic<long> delivered{0}, accepted{0};
    std::atomic<long long> score_sum{0};
    auto cb = [&](SlamEdge e) { ++delivered; if (e.matching_score_ >= 15) ++accepted; score_sum += (long long)e.matching_score_; };
    auto run_queue = [&](GpuFeatureTransformationEstimator& est) {
        const double t = now_s();
        for (long p = 0; p < n_pairs; ++p) est.estimateEdge(nodes[pairs[2 * p]], nodes[pairs[2 * p + 1]]);
        est.waitIdle();
        return now_s() - t;
    };

    // 3. cold: nothing resident, every node travels from pageable memory inside the timed region
    double first_s, cold_s, load_s, best = 1e30;
    long long score_cold, score_warm = 0;
    {
        GpuFeatureTransformationEstimator est(cb, devices);
        est.setConfig(ec);
        {   // warm-up of the context (first launches, pinned staging) on a few pairs, then forget those nodes
            for (long p = 0; p < std::min<long>(n_pairs, 64); ++p) est.estimateEdge(nodes[pairs[2 * p]], nodes[pairs[2 * p + 1]]);
            est.waitIdle();
            for (long p = 0; p < std::min<long>(n_pairs, 64); ++p) { est.forgetNode(nodes[pairs[2 * p]].id_); est.forgetNode(nodes[pairs[2 * p + 1]].id_); }
            delivered = 0; accepted = 0; score_sum = 0;
        }
        first_s = run_queue(est);               // first use: the store arena and the pinned staging ring grow inside this pass
        if (delivered != n_pairs) { std::printf("{\"error\": \"%ld of %ld edges delivered\"}\n", (long)delivered, n_pairs); return 1; }
        for (int i = 0; i < n_kf; ++i) est.forgetNode(nodes[i].id_);
        delivered = 0; accepted = 0; score_sum = 0;
        cold_s = run_queue(est);                // cold data, warm allocator: every node travels again
        score_cold = score_sum;
        if (delivered != n_pairs) { std::printf("{\"error\": \"%ld of %ld edges delivered\"}\n", (long)delivered, n_pairs); return 1; }
    }
    // 4. resume + steady state: loadNodes (GraphSlamNode::load re-adds every node), then the same queue on resident nodes
    long acc_warm = 0;
    {
        GpuFeatureTransformationEstimator est(cb, devices);
        est.setConfig(ec);
        t0 = now_s();
        if (!est.loadNodes(nodes)) { std::printf("{\"error\": \"loadNodes: %s\"}\n", est.lastError()); return 1; }
        load_s = now_s() - t0;
        for (int r = 0; r < repeats + 1; ++r) {
            delivered = 0; accepted = 0; score_sum = 0;
            const double s = run_queue(est);
            if (r > 0 && s < best) best = s;           // first pass warms the launch shapes
            score_warm = score_sum; acc_warm = accepted;
            if (delivered != n_pairs) { std::printf("{\"error\": \"%ld of %ld edges delivered\"}\n", (long)delivered, n_pairs); return 1; }
        }
    }
    std::printf("{\"bench\": \"adapter queue, C++ host\", \"devices\": %d, \"keyframes\": %d, \"pairs\": %ld, \"generator_s\": %.2f, "
                "\"map_checksum_desc\": %llu, \"map_checksum_pos\": %llu, \"first_use_ms\": %.3f, \"edges_per_s_cold_pageable\": %.1f, \"cold_ms\": %.3f, "
                "\"load_nodes_s\": %.3f, \"keyframes_per_s_load\": %.1f, \"edges_per_s_resident\": %.1f, \"resident_ms\": %.3f, "
                "\"edges_score_ge_15\": %ld, \"score_sum_cold\": %lld, \"score_sum_resident\": %lld, \"same_edges\": %s}\n",
                (int)devices.size(), n_kf, n_pairs, gen_s, ck_desc, ck_pos, first_s * 1e3, n_pairs / cold_s, cold_s * 1e3, load_s, n_kf / load_s,
                n_pairs / best, best * 1e3, acc_warm, score_cold, score_warm, score_cold == score_warm ? "true" : "false");
    return 0;
}
