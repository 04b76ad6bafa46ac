// uz_samples.h — the RANSAC sample-index list, shared with the reference.
//
// prosac() (/root/reference/transformation_estimation/src/feature_transformation_estimator.cpp:198-225)
// draws hypothesis i from idx[0..2] after std::random_shuffle of a growing prefix of a PERSISTENT index
// vector.  The draw depends only on (M, iterations, do_prosac) and the rand() stream, never on data, so
// it is replayed here on the host into a table indexed by M and shared with the device.
//   * std::random_shuffle (libstdc++ bits/stl_algo.h): for k = 1..n-1: swap(a[k], a[rand() % (k+1)])
//   * rand(): glibc TYPE_3 additive-feedback generator (r[i] = r[i-31] + r[i-3], output >> 1), default
//     seed 1 because the reference never calls srand(); the stream is restarted per pair so that pairs
//     are order-independent inside a batch.
// The product must not touch the process-global rand() state of its host application, hence the
// re-implementation; tests pin it against the real rand()/random_shuffle through the oracle.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

namespace uz {

// first `n` outputs of glibc rand() after srand(seed)
inline void glibc_rand_stream(uint32_t seed, size_t n, std::vector<uint32_t>& out) {
    out.resize(n);
    if (seed == 0) seed = 1;
    std::vector<uint32_t> r(344 + n);
    r[0] = seed;
    for (int i = 1; i < 31; ++i) {
        const int64_t hi = (int32_t)r[i - 1] / 127773, lo = (int32_t)r[i - 1] % 127773;
        int64_t word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        r[i] = (uint32_t)word;
    }
    for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
    for (size_t i = 34; i < 344 + n; ++i) r[i] = r[i - 31] + r[i - 3];
    for (size_t k = 0; k < n; ++k) out[k] = r[344 + k] >> 1;
}

inline int prosac_prefix(int i, int iterations, int M) {
    // std::min((int)std::ceil(((i + 3.) / iterations) * P.cols()), (int)P.cols())   (:217)
    return std::min((int)std::ceil(((i + 3.) / iterations) * M), M);
}

// rows[M][iterations][3] for M in [m_lo, m_hi); rows with M < 3 are left zero.
inline void build_sample_rows(int iterations, bool do_prosac, int m_lo, int m_hi,
                              const std::vector<uint32_t>& rnd, uint16_t* table /* base of M = 0 */) {
    std::vector<uint16_t> idx;
    for (int M = std::max(m_lo, 3); M < m_hi; ++M) {
        idx.resize(M);
        for (int i = 0; i < M; ++i) idx[i] = (uint16_t)i;
        size_t c = 0;
        uint16_t* row = table + (size_t)M * iterations * 3;
        for (int i = 0; i < iterations; ++i) {
            const int n = do_prosac ? prosac_prefix(i, iterations, M) : M;
            for (int k = 1; k < n; ++k) {
                const uint32_t j = rnd[c++] % (uint32_t)(k + 1);
                const uint16_t t = idx[k]; idx[k] = idx[j]; idx[j] = t;
            }
            row[3 * i] = idx[0]; row[3 * i + 1] = idx[1]; row[3 * i + 2] = idx[2];
        }
    }
}

// Whole table for M in [0, m_cap], multi-threaded over M (cost ~ iterations * m_cap^2 / 4 swaps).
inline void build_sample_table(int iterations, bool do_prosac, int m_cap, std::vector<uint16_t>& table) {
    table.assign((size_t)(m_cap + 1) * iterations * 3, 0);
    std::vector<uint32_t> rnd;
    glibc_rand_stream(1, (size_t)iterations * (size_t)std::max(m_cap, 1), rnd);
    unsigned nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if ((int64_t)iterations * m_cap * m_cap < (int64_t)4e7) nthreads = 1;
    if (nthreads == 1) { build_sample_rows(iterations, do_prosac, 0, m_cap + 1, rnd, table.data()); return; }
    // interleave M across threads so every thread gets a similar share of the quadratic cost
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; ++t)
        th.emplace_back([&, t]() {
            for (int M = 3 + (int)t; M <= m_cap; M += (int)nthreads)
                build_sample_rows(iterations, do_prosac, M, M + 1, rnd, table.data());
        });
    for (auto& x : th) x.join();
}

}  // namespace uz
