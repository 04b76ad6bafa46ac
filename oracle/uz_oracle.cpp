// uz_oracle.cpp — CPU ORACLE for the feature-edge estimation hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product
// (uzliti_slam_b200/csrc) never includes, links or calls anything in oracle/.
//
// It is a plain restatement of the reference algorithm, with the arithmetic of the
// un-vendored third-party libraries written out:
//   /root/reference/transformation_estimation/src/feature_transformation_estimator.cpp
//     :32-159  estimateEdgeDirect  (camera-pair loop, kNN-2, ratio, valid filter, sort, gather, edge)
//     :178-184 estimateSVD         (binds pose + consensus into prosac)
//     :186-297 prosac              (growing-prefix shuffle, strict '>' update, early break, refit, mse)
//     :299-314 estimatePoseSVD     (pcl::TransformationFromCorrespondences, float32, weight==1 bug)
//     :337-347 consensus3D         (double, strict '<')
//   /root/reference/transformation_estimation/src/transformation_estimator.cpp:53-55 (score 0 on failure)
// of the data formats in front of the store:
//   /root/reference/feature_extraction/src/feature_extraction_core.cpp:254-295 (extract3dFeatures)
//   /root/reference/graph_slam_common/src/sensor_data.cpp:124-171             (FeatureData::fromMsg)
// and of the step before the path (candidate generation):
//   /root/reference/place_recognition/src/lsh_set_recognizer.cpp:46-94,96-165,188-305 (LshSetRecognizer, FastLshSet/Table)
//   /root/reference/place_recognition/src/place_recognizer.cpp:73-118,140-190        (searchAndAddPlace / searchPlace filters)
//
// Third-party arithmetic restated here (sources are NOT under /root/reference):
//   * OpenCV cv::BFMatcher(NORM_HAMMING).knnMatch(k=2)  [ROS Indigo => OpenCV 2.4.8; checked live
//     against python cv2 4.13 in tests/test_oracle_matching.py]: per query row the two train rows with
//     the smallest Hamming distance, ordered by (distance, trainIdx) ascending.
//   * PCL 1.7 pcl::TransformationFromCorrespondences::{add,getTransformation}
//     (common/impl/transformation_from_correspondences.hpp): float32 incremental mean/covariance,
//     then Eigen::JacobiSVD<Matrix3f>(FullU|FullV), R = U*diag(1,1,±1)*V^T, t = mean2 - R*mean1.
//   * Eigen 3.2.0 JacobiSVD (src/SVD/JacobiSVD.h: compute, real_2x2_jacobi_svd) and
//     JacobiRotation (src/Jacobi/Jacobi.h: makeJacobi, operator*, applyOnTheLeft/Right), restated
//     from the published algorithm; square 3x3 => no QR preconditioner.
//   * libstdc++ std::random_shuffle (bits/stl_algo.h) over glibc rand(): the REAL functions are
//     called here (they exist in this container), seed restarted to 1 per prosac() call.
//
// PARITY STATUS: matching stage pinned against cv2 (live).  RANSAC stage: "parity unpinned" — the
// reference has no tests/golden vectors and PCL/Eigen are not installable here, so the float32
// solve is pinned only by (a) recovering planted ground-truth motions and (b) agreement with an
// independent float64 Kabsch solve (numpy) within float32 round-off (tests/test_oracle_ransac.py).
//
// Build: see oracle/Makefile (g++ -O3 -march=native -ffp-contract=off -std=c++14).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <set>
#include <unordered_map>
#include <utility>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// Stage 1: brute-force Hamming kNN-2  (feature_transformation_estimator.cpp:38,58)
// ---------------------------------------------------------------------------------------------
inline int hamming(const uint8_t* a, const uint8_t* b, int nbytes) {
    int d = 0, i = 0;
    for (; i + 8 <= nbytes; i += 8) {
        uint64_t x, y;
        std::memcpy(&x, a + i, 8);
        std::memcpy(&y, b + i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    for (; i < nbytes; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

// idx/dist are nq x 2; missing neighbours (nt < 2) are (-1, -1).
void knn2(const uint8_t* q, int nq, int qstride, const uint8_t* t, int nt, int tstride, int nbytes,
          int32_t* idx, int32_t* dist) {
    for (int i = 0; i < nq; ++i) {
        int d0 = INT32_MAX, d1 = INT32_MAX, i0 = -1, i1 = -1;
        const uint8_t* qi = q + (size_t)i * qstride;
        for (int j = 0; j < nt; ++j) {
            int d = hamming(qi, t + (size_t)j * tstride, nbytes);
            // strict '<' while scanning j ascending == order by (distance, trainIdx)
            if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
            else if (d < d1) { d1 = d; i1 = j; }
        }
        idx[2 * i] = i0; idx[2 * i + 1] = i1;
        dist[2 * i] = i0 < 0 ? -1 : d0; dist[2 * i + 1] = i1 < 0 ? -1 : d1;
    }
}

// ---------------------------------------------------------------------------------------------
// Stage 2: ratio test (:65-71) — literal float/double form of the reference.
// ---------------------------------------------------------------------------------------------
inline bool ratio_pass(int d0, int d1) {
    float f0 = (float)d0, f1 = (float)d1;      // cv::DMatch::distance is float
    return f0 < 0.99 * f1;                      // float < double*float, as written at :67
}

struct Match { int q, t, d; };

// Opt-in cross-check (uz_params.cross_check; NOT in the reference, which never enables it): what
// cv::BFMatcher(normType, crossCheck=true).match(query, train) returns — mutual nearest neighbours: (q, t) with
// t the nearest train row of q and q the nearest query row of t, both with lowest-index tie breaking.
// idx/dist: nq entries, -1 = query unmatched.  Pinned against live cv2 4.13 in tests/test_oracle_matching.py.
void cross_match(const uint8_t* q, int nq, int qstride, const uint8_t* t, int nt, int tstride, int nbytes,
                 int32_t* idx, int32_t* dist) {
    std::vector<int> best_q((size_t)nt, -1), best_qd((size_t)nt, INT32_MAX);
    for (int i = 0; i < nq; ++i) { idx[i] = -1; dist[i] = INT32_MAX; }
    for (int i = 0; i < nq; ++i) {
        const uint8_t* qi = q + (size_t)i * qstride;
        for (int j = 0; j < nt; ++j) {
            int d = hamming(qi, t + (size_t)j * tstride, nbytes);
            if (d < dist[i]) { dist[i] = d; idx[i] = j; }               // nearest train of q (lowest j on ties)
            if (d < best_qd[j]) { best_qd[j] = d; best_q[j] = i; }      // nearest query of t (lowest i on ties)
        }
    }
    for (int i = 0; i < nq; ++i)
        if (idx[i] < 0 || best_q[idx[i]] != i) { idx[i] = -1; dist[i] = -1; }
}

// ---------------------------------------------------------------------------------------------
// Stage 3: pose from correspondences (:299-314) — PCL add()/getTransformation() in float32.
// ---------------------------------------------------------------------------------------------
struct Rot { float c, s; };   // Eigen::JacobiRotation<float>

// JacobiRotation::makeJacobi(x, y, z)  (Jacobi.h)
inline Rot make_jacobi(float x, float y, float z) {
    Rot r;
    if (y == 0.f) { r.c = 1.f; r.s = 0.f; return r; }
    float tau = (x - z) / (2.f * std::fabs(y));
    float w = std::sqrt(tau * tau + 1.f);
    float t = (tau > 0.f) ? 1.f / (tau + w) : 1.f / (tau - w);
    float sign_t = t > 0.f ? 1.f : -1.f;
    float n = 1.f / std::sqrt(t * t + 1.f);
    r.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
    r.c = n;
    return r;
}

// rows p,q of a 3x3 (row-major m[r][c]): x' = c*x + s*y ; y' = -s*x + c*y   (applyOnTheLeft)
inline void rot_rows(float m[3][3], int p, int q, Rot j) {
    for (int k = 0; k < 3; ++k) {
        float x = m[p][k], y = m[q][k];
        m[p][k] = j.c * x + j.s * y;
        m[q][k] = -j.s * x + j.c * y;
    }
}
// columns p,q with rotation j applied "in the plane" (x' = c*x + s*y ; y' = -s*x + c*y)
inline void rot_cols(float m[3][3], int p, int q, Rot j) {
    for (int k = 0; k < 3; ++k) {
        float x = m[k][p], y = m[k][q];
        m[k][p] = j.c * x + j.s * y;
        m[k][q] = -j.s * x + j.c * y;
    }
}

// Eigen::JacobiSVD<Matrix3f>(A, ComputeFullU|ComputeFullV): A = U * diag(sv) * V^T
void jacobi_svd3(const float A[3][3], float U[3][3], float V[3][3], float sv[3]) {
    const float precision = 2.f * std::numeric_limits<float>::epsilon();
    const float considerAsZero = 2.f * std::numeric_limits<float>::denorm_min();
    float W[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) { W[r][c] = A[r][c]; U[r][c] = V[r][c] = (r == c) ? 1.f : 0.f; }

    bool finished = false;
    while (!finished) {
        finished = true;
        for (int p = 1; p < 3; ++p) {
            for (int q = 0; q < p; ++q) {
                float threshold = std::max(considerAsZero,
                                           precision * std::max(std::fabs(W[p][p]), std::fabs(W[q][q])));
                if (std::max(std::fabs(W[p][q]), std::fabs(W[q][p])) > threshold) {
                    finished = false;
                    // real_2x2_jacobi_svd(W, p, q, &j_left, &j_right)
                    float m00 = W[p][p], m01 = W[p][q], m10 = W[q][p], m11 = W[q][q];
                    Rot rot1;
                    float t = m00 + m11;
                    float d = m10 - m01;
                    if (t == 0.f) {
                        rot1.c = 0.f;
                        rot1.s = d > 0.f ? 1.f : -1.f;
                    } else {
                        float u = d / t;
                        rot1.c = 1.f / std::sqrt(1.f + u * u);
                        rot1.s = rot1.c * u;
                    }
                    // m.applyOnTheLeft(0,1,rot1)
                    float n00 = rot1.c * m00 + rot1.s * m10;
                    float n01 = rot1.c * m01 + rot1.s * m11;
                    float n11 = -rot1.s * m01 + rot1.c * m11;
                    Rot jr = make_jacobi(n00, n01, n11);
                    // j_left = rot1 * j_right.transpose()
                    Rot jrt = {jr.c, -jr.s};
                    Rot jl = {rot1.c * jrt.c - rot1.s * jrt.s, rot1.c * jrt.s + rot1.s * jrt.c};
                    rot_rows(W, p, q, jl);                 // m_workMatrix.applyOnTheLeft(p,q,j_left)
                    rot_cols(U, p, q, jl);                 // m_matrixU.applyOnTheRight(p,q,j_left.transpose())
                    rot_cols(W, p, q, jrt);                // m_workMatrix.applyOnTheRight(p,q,j_right)
                    rot_cols(V, p, q, jrt);                // m_matrixV.applyOnTheRight(p,q,j_right)
                }
            }
        }
    }
    // step 3: make the diagonal positive
    for (int i = 0; i < 3; ++i) {
        float a = std::fabs(W[i][i]);
        sv[i] = a;
        if (a != 0.f) {
            float sgn = W[i][i] / a;
            for (int r = 0; r < 3; ++r) U[r][i] *= sgn;
        }
    }
    // step 4: sort descending
    for (int i = 0; i < 3; ++i) {
        int pos = 0;
        float best = sv[i];
        for (int k = i + 1; k < 3; ++k)
            if (sv[k] > best) { best = sv[k]; pos = k - i; }
        if (best == 0.f) break;
        if (pos) {
            pos += i;
            std::swap(sv[i], sv[pos]);
            for (int r = 0; r < 3; ++r) { std::swap(U[r][i], U[r][pos]); std::swap(V[r][i], V[r][pos]); }
        }
    }
}

inline float det3(const float m[3][3]) {      // Eigen bruteforce_det3_helper order
    return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1])
         - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0])
         + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}

// P,Q: 3 x k column-major double (P.col(i) = P[3*i..3*i+2]).  T: 4x4 row-major double, T*p ~= q.
void pose_svd(const double* P, const double* Q, int k, double T[16]) {
    int n = 0;
    float acc = 0.f;
    float mean1[3] = {0, 0, 0}, mean2[3] = {0, 0, 0};
    float cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < k; ++i) {
        float p[3] = {(float)P[3 * i], (float)P[3 * i + 1], (float)P[3 * i + 2]};
        float q[3] = {(float)Q[3 * i], (float)Q[3 * i + 1], (float)Q[3 * i + 2]};
        // :305-309 — inverse_weight is computed, but weight = 1./weight with weight == 1: always 1.
        float inverse_weight = p[2] * p[2] + q[2] * q[2];
        float weight = 1;
        if (inverse_weight > 0) weight = 1. / weight;
        // pcl::TransformationFromCorrespondences::add
        if (weight == 0.0f) continue;
        ++n;
        acc += weight;
        float alpha = weight / acc;
        float d1[3] = {p[0] - mean1[0], p[1] - mean1[1], p[2] - mean1[2]};
        float d2[3] = {q[0] - mean2[0], q[1] - mean2[1], q[2] - mean2[2]};
        float oma = 1.0f - alpha;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                cov[r][c] = oma * (cov[r][c] + alpha * (d2[r] * d1[c]));
        for (int r = 0; r < 3; ++r) { mean1[r] += alpha * d1[r]; mean2[r] += alpha * d2[r]; }
    }
    (void)n;
    float U[3][3], V[3][3], sv[3];
    jacobi_svd3(cov, U, V, sv);
    float sgn = (det3(U) * det3(V) < 0.0f) ? -1.0f : 1.0f;
    float R[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            R[r][c] = U[r][0] * V[c][0] + U[r][1] * V[c][1] + (U[r][2] * sgn) * V[c][2];
    float t[3];
    for (int r = 0; r < 3; ++r)
        t[r] = mean2[r] - (R[r][0] * mean1[0] + R[r][1] * mean1[1] + R[r][2] * mean1[2]);
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) T[4 * r + c] = (double)R[r][c];
        T[4 * r + 3] = (double)t[r];
    }
    T[12] = T[13] = T[14] = 0.0; T[15] = 1.0;
}

// ---------------------------------------------------------------------------------------------
// Stage 4: consensus3D (:337-347) — double, ((r0*x + r1*y) + r2*z) + t, strict '<'.
// ---------------------------------------------------------------------------------------------
inline double residual(const double T[16], const double* p, const double* q) {
    double x = ((T[0] * p[0] + T[1] * p[1]) + T[2] * p[2]) + T[3];
    double y = ((T[4] * p[0] + T[5] * p[1]) + T[6] * p[2]) + T[7];
    double z = ((T[8] * p[0] + T[9] * p[1]) + T[10] * p[2]) + T[11];
    double dx = x - q[0], dy = y - q[1], dz = z - q[2];
    return std::sqrt((dx * dx + dy * dy) + dz * dz);
}

int consensus3d(const double* P, const double* Q, int M, const double T[16], double thr, uint8_t* set) {
    int count = 0;
    for (int i = 0; i < M; ++i) {
        bool in = residual(T, P + 3 * i, Q + 3 * i) < thr;
        set[i] = in;
        count += in;
    }
    return count;
}

// ---------------------------------------------------------------------------------------------
// Stage 5: the sample-index list (:198-225).  Data independent: depends on (M, I, do_prosac) and
// the rand() stream only.  Uses the REAL std::random_shuffle + glibc rand(), seed restarted to 1.
// ---------------------------------------------------------------------------------------------
void sample_list(int M, int iterations, int do_prosac, int32_t* out /* iterations x 3 */) {
    std::vector<int> idx;
    for (int i = 0; i < M; i++) idx.push_back(i);
    std::srand(1);
    for (int i = 0; i < iterations; i++) {
        if (do_prosac) {
            std::random_shuffle(idx.begin(),
                                idx.begin() + std::min((int)std::ceil(((i + 3.) / iterations) * M), (int)M));
        } else {
            std::random_shuffle(idx.begin(), idx.end());
        }
        for (int j = 0; j < 3; ++j) out[3 * i + j] = idx[j];
    }
}

struct ProsacOut {
    double T[16];
    int consensus;
    double mse;
    int best_iteration;   // index of the winning hypothesis, -1 if none
    int iterations_run;   // hypotheses evaluated before break / exhaustion
};

// prosac (:186-297) with minCorrespondenceCount = 3, on an explicit sample list.
void prosac(const double* P, const double* Q, int M, const int32_t* samples, int iterations,
            double thr, double breakPercentage, ProsacOut* out, uint8_t* inlier_mask /* M */,
            int32_t* counts_out = nullptr /* iterations; -1 = not evaluated */) {
    if (counts_out) for (int c = 0; c < iterations; ++c) counts_out[c] = -1;
    const int minCount = 3;
    int maxConsensus = 0;
    std::vector<uint8_t> set(M), maxSet(M, 0);
    double T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    double Ttemp[16];
    out->best_iteration = -1;
    int i = 0;
    for (; i < iterations; i++) {
        double Pt[9], Qt[9];
        for (int j = 0; j < minCount; ++j) {
            int s = samples[3 * i + j];
            for (int r = 0; r < 3; ++r) { Pt[3 * j + r] = P[3 * s + r]; Qt[3 * j + r] = Q[3 * s + r]; }
        }
        pose_svd(Pt, Qt, minCount, Ttemp);
        int consensus = consensus3d(P, Q, M, Ttemp, thr, set.data());
        if (counts_out) counts_out[i] = consensus;
        if (consensus > maxConsensus) {
            maxConsensus = consensus;
            maxSet = set;
            std::memcpy(T, Ttemp, sizeof(T));
            out->best_iteration = i;
            if (maxConsensus >= minCount && maxConsensus > breakPercentage * M) { ++i; break; }
        }
    }
    out->iterations_run = i;
    double mse = 0.;
    if (maxConsensus >= minCount) {
        std::vector<double> Pf(3 * (size_t)maxConsensus), Qf(3 * (size_t)maxConsensus);
        int k = 0;
        for (int c = 0; c < M; ++c)
            if (maxSet[c]) {
                for (int r = 0; r < 3; ++r) { Pf[3 * k + r] = P[3 * c + r]; Qf[3 * k + r] = Q[3 * c + r]; }
                k++;
            }
        pose_svd(Pf.data(), Qf.data(), maxConsensus, T);
        maxConsensus = consensus3d(P, Q, M, T, thr, maxSet.data());
        for (int c = 0; c < M; ++c)
            if (maxSet[c]) mse += residual(T, P + 3 * c, Q + 3 * c);
        mse /= maxConsensus;                       // NaN if the refit recount dropped to 0 (as :290)
    } else {
        maxConsensus = 0;
        static const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        std::memcpy(T, I4, sizeof(T));
        std::fill(maxSet.begin(), maxSet.end(), 0);
    }
    std::memcpy(out->T, T, sizeof(T));
    out->consensus = maxConsensus;
    out->mse = mse;
    if (inlier_mask) std::memcpy(inlier_mask, maxSet.data(), M);
}

}  // namespace

// =================================================================================================
// extern "C" surface (ctypes)
// =================================================================================================
extern "C" {

// One camera's FeatureData (graph_slam_common/include/graph_slam_common/sensor_data.h:49-70) as POD.
struct uzo_features {
    const uint8_t* descriptors;   // n x desc_bytes, row stride desc_stride   (cv::Mat features_)
    const double* positions;      // 3 x n column-major                       (feature_positions_)
    const uint8_t* valid_3d;      // n bytes                                   (valid_3d_)
    int32_t n;
    int32_t desc_bytes;
    int32_t desc_stride;
    int32_t feature_type;         // graph_slam_msgs/Features: BRIEF=1 ORB=2 BRISK=3 FREAK=4 SURF=5 SIFT=6
    int32_t sensor_frame;         // interned sensor_frame_ string
};

struct uzo_edge {
    int32_t ok;                   // estimateEdgeDirect return value
    int32_t cam_from, cam_to;     // index of the winning FeatureData in each list (-1: none)
    int32_t n_ratio_matches;      // score at :78
    int32_t n_matches;            // M after the valid_3d filter (:115)
    int32_t consensus;            // matching_score_ (:155); 0 when !ok (transformation_estimator.cpp:54)
    int32_t best_iteration;
    int32_t iterations_run;
    double mse;
    double info_scale;            // information_ = I6*info_scale, rotation block additionally x100 (:133-137)
    double T[16];                 // transform_ row-major 4x4
};

int uzo_knn2(const uint8_t* q, int nq, int qstride, const uint8_t* t, int nt, int tstride, int nbytes,
             int32_t* idx, int32_t* dist) {
    knn2(q, nq, qstride, t, nt, tstride, nbytes, idx, dist);
    return 0;
}

int uzo_cross_match(const uint8_t* q, int nq, int qstride, const uint8_t* t, int nt, int tstride, int nbytes,
                    int32_t* idx, int32_t* dist) {
    cross_match(q, nq, qstride, t, nt, tstride, nbytes, idx, dist);
    return 0;
}

int uzo_ratio_pass(int d0, int d1) { return ratio_pass(d0, d1) ? 1 : 0; }

void uzo_sample_list(int M, int iterations, int do_prosac, int32_t* out) {
    sample_list(M, iterations, do_prosac, out);
}

// first n outputs of glibc rand() after srand(seed) — used to pin the product's own generator.
void uzo_glibc_rand(unsigned seed, int n, int32_t* out) {
    std::srand(seed);
    for (int i = 0; i < n; ++i) out[i] = std::rand();
}

void uzo_pose_svd(const double* P, const double* Q, int k, double* T16) { pose_svd(P, Q, k, T16); }

int uzo_consensus3d(const double* P, const double* Q, int M, const double* T16, double thr, uint8_t* set) {
    return consensus3d(P, Q, M, T16, thr, set);
}

// estimateSVD (:178-184).  samples may be NULL (generated from rand() seed 1 as the reference would).
void uzo_estimate_svd(const double* P, const double* Q, int M, double thr, int iterations, double bp,
                      int do_prosac, const int32_t* samples, double* T16, int32_t* consensus, double* mse,
                      uint8_t* inlier_mask, int32_t* best_iteration, int32_t* iterations_run) {
    std::vector<int32_t> own;
    if (!samples) {
        own.resize(3 * (size_t)iterations);
        if (M >= 3) sample_list(M, iterations, do_prosac, own.data());
        samples = own.data();
    }
    ProsacOut po;
    if (M < 3) {   // the reference would index idx[0..2] out of range; callers guard with :118
        static const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        std::memcpy(T16, I4, sizeof(I4));
        *consensus = 0; *mse = 0;
        if (inlier_mask) std::memset(inlier_mask, 0, M);
        if (best_iteration) *best_iteration = -1;
        if (iterations_run) *iterations_run = 0;
        return;
    }
    prosac(P, Q, M, samples, iterations, thr, bp, &po, inlier_mask);
    std::memcpy(T16, po.T, sizeof(po.T));
    *consensus = po.consensus;
    *mse = po.mse;
    if (best_iteration) *best_iteration = po.best_iteration;
    if (iterations_run) *iterations_run = po.iterations_run;
}

// estimateEdgeDirect (:32-159) + the failure convention of transformation_estimator.cpp:53-55.
// Optional parity outputs (may be NULL): matches_out (capacity max_matches x 3 int32: queryIdx,
// trainIdx, distance — the sorted final_matches of :114) and inlier_mask (capacity max_matches).
void uzo_estimate_edge(const uzo_features* from, int n_from, const uzo_features* to, int n_to,
                       double thr, int iterations, double bp, int do_prosac, int min_keypoints,
                       uzo_edge* edge, int32_t* matches_out, uint8_t* inlier_mask, int max_matches,
                       int32_t* counts_out, int cross_check) {
    static const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    std::memset(edge, 0, sizeof(*edge));
    std::memcpy(edge->T, I4, sizeof(I4));
    edge->info_scale = 1.0;
    edge->cam_from = edge->cam_to = -1;
    edge->best_iteration = -1;

    double best_matching_score = -1;
    int best_f = -1, best_t = -1;
    std::vector<Match> potential;
    for (int a = 0; a < n_from; ++a) {
        for (int b = 0; b < n_to; ++b) {
            const uzo_features& F = from[a];
            const uzo_features& Tt = to[b];
            if (F.n >= min_keypoints && Tt.n >= min_keypoints && F.feature_type == Tt.feature_type &&
                F.sensor_frame == Tt.sensor_frame) {
                std::vector<Match> matches;
                bool binary = F.feature_type >= 1 && F.feature_type <= 4;   // :54-57
                if (binary) {
                    std::vector<int32_t> idx(2 * (size_t)Tt.n), dist(2 * (size_t)Tt.n);
                    knn2(Tt.descriptors, Tt.n, Tt.desc_stride, F.descriptors, F.n, F.desc_stride,
                         F.desc_bytes, idx.data(), dist.data());
                    std::vector<int32_t> xidx, xdist;
                    if (cross_check) {     // opt-in, not in the reference: keep (q, t) only if the crossCheck matcher returns it
                        xidx.resize((size_t)Tt.n); xdist.resize((size_t)Tt.n);
                        cross_match(Tt.descriptors, Tt.n, Tt.desc_stride, F.descriptors, F.n, F.desc_stride,
                                    F.desc_bytes, xidx.data(), xdist.data());
                    }
                    for (int q = 0; q < Tt.n; ++q) {
                        if (idx[2 * q] >= 0 && idx[2 * q + 1] >= 0) {                 // size() == 2
                            if (ratio_pass(dist[2 * q], dist[2 * q + 1]) && (!cross_check || xidx[q] == idx[2 * q]))
                                matches.push_back(Match{q, idx[2 * q], dist[2 * q]});
                        }
                    }
                }
                double score = (double)matches.size();
                if (score > best_matching_score) {
                    best_matching_score = score;
                    best_f = a; best_t = b;
                    potential = matches;
                }
            }
        }
    }
    if (best_matching_score == -1) return;     // :93-95
    edge->cam_from = best_f; edge->cam_to = best_t;
    edge->n_ratio_matches = (int)potential.size();

    const uzo_features& F = from[best_f];
    const uzo_features& Tt = to[best_t];
    std::vector<Match> fin;
    for (const Match& m : potential)
        if (F.valid_3d[m.t] && Tt.valid_3d[m.q]) fin.push_back(m);
    // :114 std::sort by distance is UNSTABLE in the reference (order among equal distances is
    // unspecified); the build fixes the total order (distance, queryIdx) on both sides.
    std::stable_sort(fin.begin(), fin.end(), [](const Match& a, const Match& b) {
        return a.d != b.d ? a.d < b.d : a.q < b.q;
    });
    edge->n_matches = (int)fin.size();
    if (matches_out)
        for (int i = 0; i < (int)fin.size() && i < max_matches; ++i) {
            matches_out[3 * i] = fin[i].q; matches_out[3 * i + 1] = fin[i].t; matches_out[3 * i + 2] = fin[i].d;
        }
    if (fin.size() >= 3) {
        int M = (int)fin.size();
        std::vector<double> Xd(3 * (size_t)M), Pd(3 * (size_t)M);
        for (int i = 0; i < M; ++i)
            for (int r = 0; r < 3; ++r) {
                Xd[3 * i + r] = F.positions[3 * (size_t)fin[i].t + r];
                Pd[3 * i + r] = Tt.positions[3 * (size_t)fin[i].q + r];
            }
        std::vector<int32_t> samples(3 * (size_t)iterations);
        sample_list(M, iterations, do_prosac, samples.data());
        ProsacOut po;
        std::vector<uint8_t> mask(M);
        prosac(Pd.data(), Xd.data(), M, samples.data(), iterations, thr, bp, &po, mask.data(), counts_out);   // :130
        if (inlier_mask) std::memcpy(inlier_mask, mask.data(), std::min(M, max_matches));
        if (po.consensus > 0 && po.mse > 0) edge->info_scale = 0.1 * po.consensus / po.mse;      // :134-135
        std::memcpy(edge->T, po.T, sizeof(po.T));
        edge->consensus = po.consensus;
        edge->mse = po.mse;
        edge->best_iteration = po.best_iteration;
        edge->iterations_run = po.iterations_run;
        edge->ok = 1;
    }
}

}  // extern "C"


// =================================================================================================
// 8f-1: candidate generation — LshSetRecognizer + PlaceRecognizer filters, sequential restatement.
// =================================================================================================
namespace {

// FastLshSet(4): 8 tables keyed by descriptor bytes [4k, 4k+4) (lsh_set_recognizer.cpp:262-268, :190-200);
// a bucket is the list of place indices, one entry per descriptor row that carried the key (duplicates kept).
struct PlacesOracle {
    static const int kTables = 8, kKeyWidth = 4;
    std::unordered_map<uint32_t, std::vector<int>> tables[kTables];
    int place_count = 0;                                   // place_recognizer.h: place_count_
    std::map<int, long long> left;                         // place_id_map_.left : place index -> id (live places)
    std::map<long long, int> right;                        // place_id_map_.right: id -> place index
    std::map<long long, long long> time_ns;                // pr_time_map_
    std::set<std::pair<long long, long long>> checked;     // checked_

    static uint32_t key_of(const uint8_t* d, int k) {
        uint32_t v;
        std::memcpy(&v, d + kKeyWidth * k, 4);             // index.b[i] = descriptor[start_byte_ + i] (little endian)
        return v;
    }
    // FastLshSet::add (:276-284): no popcount filter
    void add(const uint8_t* desc, int n, int stride, int id) {
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < kTables; ++k) tables[k][key_of(desc + (size_t)i * stride, k)].push_back(id);
    }
    // FastLshSet::match (:286-294)
    void match(const uint8_t* desc, int n, int stride, std::vector<int>& votes) {
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < kTables; ++k) {
                auto it = tables[k].find(key_of(desc + (size_t)i * stride, k));
                if (it != tables[k].end())
                    for (int ind : it->second) votes[ind]++;
            }
    }
    // FastLshSet::matchAndAdd (:296-305) with FastLshTable::matchAndAdd's popcount filter (:231-246)
    void match_and_add(const uint8_t* desc, int n, int stride, int id, std::vector<int>& votes) {
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < kTables; ++k) {
                const uint32_t key = key_of(desc + (size_t)i * stride, k);
                if (__builtin_popcount(key) > 3 * kKeyWidth) {
                    auto& bucket = tables[k][key];
                    for (int ind : bucket) votes[ind]++;
                    bucket.push_back(id);
                }
            }
    }
    // "Get all matches that have more than T similarity" + sort (:72-90).  std::sort is unstable in the reference;
    // the build fixes the order among equal similarities to ascending place index (stable sort of the ascending scan).
    static void rank(const std::vector<int>& votes, double T, std::vector<int>& res) {
        std::vector<std::pair<int, float>> matches;
        for (int i = 0; i < (int)votes.size(); i++)
            if (votes[i] > 0) {
                float similarity = (float)votes[i] / (float)kTables;
                if (similarity >= T) matches.push_back(std::make_pair(i, similarity));
            }
        std::stable_sort(matches.begin(), matches.end(),
                         [](const std::pair<int, float>& l, const std::pair<int, float>& r) { return l.second > r.second; });
        for (auto& m : matches) res.push_back(m.first);
    }
    // place_recognizer.cpp:91-116: drop unknown/removed places and |dt| <= 5 s, take k, then the checked_ filter
    int filter(const std::vector<int>& neighbors, long long id, int k_nn, long long* pairs_out, int cap) {
        std::vector<long long> mapped;
        int pr_count = 0;
        for (int nb : neighbors) {
            auto it = left.find(nb);
            if (it != left.end()) {
                const long long dt = time_ns[it->second] - time_ns[id];
                if ((dt < 0 ? -dt : dt) > 5000000000LL) {
                    mapped.push_back(it->second);
                    pr_count++;
                    if (pr_count >= k_nn) break;
                }
            }
        }
        int n = 0;
        for (long long m : mapped) {
            auto pr = std::make_pair(m, id);
            if (checked.find(pr) == checked.end()) {
                if (n < cap) { pairs_out[2 * n] = m; pairs_out[2 * n + 1] = id; }
                ++n;
                checked.insert(pr);
            }
        }
        return n;
    }
};

}  // namespace

extern "C" {

void* uzo_places_new() { return new PlacesOracle(); }
void uzo_places_free(void* h) { delete (PlacesOracle*)h; }

// PlaceRecognizer::clear (place_recognizer.cpp:49-62).  Like the product, this also forgets checked_ (the reference
// keeps it; with recycled integer ids that would alias unrelated keyframes — recorded as a deliberate deviation).
void uzo_places_clear(void* h) {
    PlacesOracle* o = (PlacesOracle*)h;
    for (auto& t : o->tables) t.clear();
    o->place_count = 0; o->left.clear(); o->right.clear(); o->time_ns.clear(); o->checked.clear();
}

// addNode's pr_time_map_ entry + searchAndAddPlace (place_recognizer.cpp:64-118) over LshSetRecognizer::
// searchAndAddPlaceImpl (lsh_set_recognizer.cpp:46-94).  cams: the node's FEATURE sensors.  Returns the number of
// (neighbor id, id) pairs; at most cap are written.
int uzo_places_search_and_add(void* h, long long id, long long stamp_ns, const uzo_features* cams, int n_cams,
                              double T, int k_nn, long long* pairs_out, int cap) {
    PlacesOracle* o = (PlacesOracle*)h;
    o->time_ns[id] = stamp_ns;
    if (o->right.find(id) != o->right.end()) return 0;                    // "tried to add existing place"
    std::vector<int> neighbors;
    for (int c = 0; c < n_cams; ++c) {
        std::vector<int> votes(o->place_count + 1, 0);
        if (cams[c].n > 150) o->match_and_add(cams[c].descriptors, cams[c].n, cams[c].desc_stride, o->place_count, votes);
        else o->match(cams[c].descriptors, cams[c].n, cams[c].desc_stride, votes);
        PlacesOracle::rank(votes, T, neighbors);
    }
    o->left[o->place_count] = id; o->right[id] = o->place_count;
    o->place_count++;
    return o->filter(neighbors, id, k_nn, pairs_out, cap);
}

// addPlace (place_recognizer.cpp:120-147) over addPlaceImpl (lsh_set_recognizer.cpp:96-119)
void uzo_places_add(void* h, long long id, long long stamp_ns, const uzo_features* cams, int n_cams) {
    PlacesOracle* o = (PlacesOracle*)h;
    o->time_ns[id] = stamp_ns;
    if (o->right.find(id) != o->right.end()) return;
    for (int c = 0; c < n_cams; ++c)
        if (cams[c].n > 150) o->add(cams[c].descriptors, cams[c].n, cams[c].desc_stride, o->place_count);
    o->left[o->place_count] = id; o->right[id] = o->place_count;
    o->place_count++;
}

// searchPlace (place_recognizer.cpp:154-190) over searchImpl (lsh_set_recognizer.cpp:121-165).  The caller's stamp is
// used for the time filter without being recorded (searchPlace reads pr_time_map_[id]; callers set it beforehand).
int uzo_places_search(void* h, long long id, long long stamp_ns, const uzo_features* cams, int n_cams, double T, int k_nn,
                      long long* pairs_out, int cap) {
    PlacesOracle* o = (PlacesOracle*)h;
    if (o->left.empty()) return 0;
    o->time_ns[id] = stamp_ns;
    std::vector<int> neighbors;
    for (int c = 0; c < n_cams; ++c) {
        std::vector<int> votes(o->place_count, 0);
        o->match(cams[c].descriptors, cams[c].n, cams[c].desc_stride, votes);
        PlacesOracle::rank(votes, T, neighbors);
    }
    return o->filter(neighbors, id, k_nn, pairs_out, cap);
}

// removePlace (place_recognizer.cpp:198-206).  The bucket entries are left in place: a removed place index is never
// reported again because filter() looks it up in place_id_map_.left first, which is all removePlaceImpl's erase achieves.
void uzo_places_remove(void* h, long long id) {
    PlacesOracle* o = (PlacesOracle*)h;
    auto it = o->right.find(id);
    if (it == o->right.end()) return;
    o->left.erase(it->second);
    o->right.erase(it);
    o->time_ns.erase(id);
}

// votes of one descriptor set against the current tables (FastLshSet::match), for parity taps
void uzo_places_votes(void* h, const uzo_features* cam, int filtered, int32_t* votes_out /* place_count */) {
    PlacesOracle* o = (PlacesOracle*)h;
    std::vector<int> votes(o->place_count + 1, 0);
    if (filtered) {
        for (int i = 0; i < cam->n; ++i)
            for (int k = 0; k < 8; ++k) {
                const uint32_t key = PlacesOracle::key_of(cam->descriptors + (size_t)i * cam->desc_stride, k);
                if (__builtin_popcount(key) > 12) {
                    auto it = o->tables[k].find(key);
                    if (it != o->tables[k].end()) for (int ind : it->second) votes[ind]++;
                }
            }
    } else {
        o->match(cam->descriptors, cam->n, cam->desc_stride, votes);
    }
    for (int i = 0; i < o->place_count; ++i) votes_out[i] = votes[i];
}

}  // extern "C"


// =================================================================================================
// 8f-2: edge acceptance gate of GraphSlamNode::newEdgeCallback (graph_slam/src/graph_slam_node.cpp:798-804).
// Eigen 3.2 arithmetic written out: Quaternion(Matrix3d) (Quaternion.h, Shoemake) then AngleAxis(Quaternion)
// (AngleAxis.h: angle = 2 acos(clamp(w)), 0 if |vec|^2 < dummy_precision^2).
// =================================================================================================
extern "C" int uzo_gate_edge(const double* T16, int ok, int consensus, double min_score, double max_T, double max_R,
                             double* tnorm_out, double* rot_deg_out) {
    double m[3][3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m[r][c] = T16[4 * r + c];
    double q[3], w;
    double t = (m[0][0] + m[1][1]) + m[2][2];
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        w = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[2][1] - m[1][2]) * t; q[1] = (m[0][2] - m[2][0]) * t; q[2] = (m[1][0] - m[0][1]) * t;
    } else {
        int i = 0;
        if (m[1][1] > m[0][0]) i = 1;
        if (m[2][2] > m[i][i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        w = (m[k][j] - m[j][k]) * t;
        q[j] = (m[j][i] + m[i][j]) * t;
        q[k] = (m[k][i] + m[i][k]) * t;
    }
    double n2 = (q[0] * q[0] + q[1] * q[1]) + q[2] * q[2];
    double angle = 0.0;
    if (!(n2 < 1e-12 * 1e-12)) angle = 2.0 * std::acos(std::min(std::max(-1.0, w), 1.0));
    double rot = std::fabs(angle) * 180 / M_PI;
    double tx = T16[3], ty = T16[7], tz = T16[11];
    double tn = std::sqrt((tx * tx + ty * ty) + tz * tz);
    if (tnorm_out) *tnorm_out = tn;
    if (rot_deg_out) *rot_deg_out = rot;
    double score = ok ? (double)consensus : 0.0;           // transformation_estimator.cpp:53-55
    return (score >= min_score && tn <= max_T && rot <= max_R) ? 1 : 0;
}


// =================================================================================================
// 8f-3: extract3dFeatures (feature_extraction_core.cpp:254-295), literal.  positions: 3 x n column-major.
// =================================================================================================
extern "C" void uzo_backproject(const int32_t* u_in, const int32_t* v_in, int n, const float* depth, int stride_floats,
                                int width, int height, double fx, double fy, double cx, double cy, double max_depth,
                                int reverse, double* positions, uint8_t* valid) {
    int out = 0;
    for (int i = n - 1; i >= 0; i--) {                       // :263 walks back to front
        int u = (int)round((double)u_in[i]);
        if (u < 0) u = 0; else if (u >= width) u = width - 1;
        int v = (int)round((double)v_in[i]);
        if (v < 0) v = 0; else if (v >= height) v = height - 1;
        double d = ((double)depth[(size_t)v * stride_floats + u]);
        const int o = reverse ? out : i;                     // reverse: the reference's push_back order
        if (d != 0 && !std::isnan(d) && (max_depth == 0. || d <= max_depth)) {
            positions[3 * o + 2] = d;
            positions[3 * o + 0] = (u - cx) * d / fx;
            positions[3 * o + 1] = (v - cy) * d / fy;
            valid[o] = 1;
        } else {
            positions[3 * o + 2] = -1; positions[3 * o + 0] = 0; positions[3 * o + 1] = 0;
            valid[o] = 0;
        }
        ++out;
    }
}

// =================================================================================================
// 8f-4, the other direction: FeatureData::toMsg (sensor_data.cpp:77-122) as the ROS1-serialised graph_slam_msgs/Feature[]
// field.  u, v come from feature_positions_2d_ (the caller's n x 2 int32, zeros if absent), keypoint_strength is -1 (:92),
// every descriptor byte becomes a float32 (:101), keypoint_position the three doubles (:113-115), is_3d = valid_3d_ (:116).
// Returns the byte count, or -1 if the buffer is too small.
// =================================================================================================
extern "C" long uzo_wire_encode(const uint8_t* desc, int n, int cols, int desc_stride, const double* positions,
                                const uint8_t* valid, const int32_t* uv, uint8_t* blob, size_t capacity) {
    const size_t need = 4 + (size_t)n * (17 + 4 * (size_t)cols + 24);
    if (need > capacity) return -1;
    uint32_t cnt = (uint32_t)n;
    std::memcpy(blob, &cnt, 4);
    size_t off = 4;
    for (int i = 0; i < n; ++i) {
        int32_t u = uv ? uv[2 * i] : 0, v = uv ? uv[2 * i + 1] : 0;
        float strength = -1.0f;
        uint32_t len = (uint32_t)cols;
        std::memcpy(blob + off, &u, 4); std::memcpy(blob + off + 4, &v, 4);
        blob[off + 8] = valid[i] ? 1 : 0;
        std::memcpy(blob + off + 9, &strength, 4);
        std::memcpy(blob + off + 13, &len, 4);
        off += 17;
        for (int j = 0; j < cols; ++j) {
            float val = desc[(size_t)i * desc_stride + j];               // :101  val = features_.at<unsigned char>(i,j)
            std::memcpy(blob + off + 4 * j, &val, 4);
        }
        off += 4 * (size_t)cols;
        std::memcpy(blob + off, positions + 3 * (size_t)i, 24);
        off += 24;
    }
    return (long)off;
}

// =================================================================================================
// 8f-4: FeatureData::fromMsg (sensor_data.cpp:124-171) on a ROS1-serialised graph_slam_msgs/Feature[] field.
// ROS1 wire format: little endian, fields in .msg order, no padding, variable arrays behind a uint32 length, bool = 1 byte.
// Returns the feature count, or -1 if the blob is malformed / a descriptor length differs from the first one.
// =================================================================================================
extern "C" int uzo_wire_decode(const uint8_t* blob, size_t bytes, int capacity, uint8_t* desc /* n x cols */, int* cols_out,
                               double* positions, uint8_t* valid, int32_t* uv) {
    if (bytes < 4) return -1;
    uint32_t n;
    std::memcpy(&n, blob, 4);
    size_t off = 4;
    int cols = -1;
    if ((int)n > capacity) return -1;
    for (uint32_t i = 0; i < n; ++i) {
        if (off + 17 > bytes) return -1;
        int32_t u, v;
        uint8_t is3d;
        uint32_t len;
        std::memcpy(&u, blob + off, 4); std::memcpy(&v, blob + off + 4, 4);
        is3d = blob[off + 8];
        std::memcpy(&len, blob + off + 13, 4);
        off += 17;
        if (cols < 0) cols = (int)len;
        if ((int)len != cols || off + 4 * (size_t)len + 24 > bytes) return -1;
        for (uint32_t j = 0; j < len; ++j) {
            float val;
            std::memcpy(&val, blob + off + 4 * j, 4);
            desc[(size_t)i * cols + j] = (unsigned char)val;             // :140
        }
        off += 4 * (size_t)len;
        std::memcpy(positions + 3 * (size_t)i, blob + off, 24);          // :159-161
        off += 24;
        valid[i] = is3d ? 1 : 0;                                         // :164
        if (uv) { uv[2 * i] = u; uv[2 * i + 1] = v; }
    }
    if (cols_out) *cols_out = cols < 0 ? 0 : cols;
    return (int)n;
}
