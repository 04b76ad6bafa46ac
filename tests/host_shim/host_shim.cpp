// host_shim.cpp — TEST-ONLY: compiles the product's arithmetic spec (csrc/uz_arith.cuh) and sample-list
// generator (csrc/uz_samples.h) for the HOST so the CPU test-suite can compare them bit-for-bit with the
// independently written oracle before any GPU time is spent.  Not part of the product; never shipped.
#include <cstring>
#include <cstdint>
#include "../../uzliti_slam_b200/csrc/uz_arith.cuh"
#include "../../uzliti_slam_b200/csrc/uz_samples.h"

extern "C" {
void hs_pose(const double* P, const double* Q, int k, double* T16) {
    uz::PoseAcc acc;
    uz::pose_reset(acc);
    for (int i = 0; i < k; ++i)
        uz::pose_add(acc, (float)P[3 * i], (float)P[3 * i + 1], (float)P[3 * i + 2], (float)Q[3 * i],
                     (float)Q[3 * i + 1], (float)Q[3 * i + 2]);
    double T12[12];
    uz::pose_finish(acc, T12);
    std::memcpy(T16, T12, sizeof(T12));
    T16[12] = T16[13] = T16[14] = 0.0; T16[15] = 1.0;
}
double hs_residual_sq(const double* T16, const double* p, const double* q) {
    return uz::residual_sq(T16, p[0], p[1], p[2], q[0], q[1], q[2]);
}
void hs_glibc_rand(unsigned seed, int n, int32_t* out) {
    std::vector<uint32_t> r;
    uz::glibc_rand_stream(seed, (size_t)n, r);
    for (int i = 0; i < n; ++i) out[i] = (int32_t)r[i];
}
void hs_sample_table(int iterations, int do_prosac, int m_cap, uint16_t* out) {
    std::vector<uint16_t> t;
    uz::build_sample_table(iterations, do_prosac != 0, m_cap, t);
    std::memcpy(out, t.data(), t.size() * sizeof(uint16_t));
}
}
