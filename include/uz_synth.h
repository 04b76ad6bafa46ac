/* uz_synth.h - the synthetic keyframe map of the benchmark, header-only C99/C++.
 *
 * One generator, two languages: this file and uzliti_slam_b200/synth_splitmix.py produce the SAME BYTES for the same
 * configuration (tests/test_synth_splitmix.py compiles this header and compares), so a C++ host - the adapter's test and
 * bench programs - sees exactly the keyframes bench.py measures on.
 *
 * Randomness is counter based: draw(s, j) is the j-th output of a splitmix64 generator seeded with s, and every entity
 * (cluster pool, keyframe pose, keyframe noise ...) owns a child stream sub(parent, k).  Gaussians are Irwin-Hall sums of
 * 12 uniforms, rotations come from a unit quaternion: only IEEE +, -, *, / and sqrt are used, evaluated in the written
 * order, so numpy and a C compiler agree to the last bit.  Compile WITHOUT -ffast-math and with -ffp-contract=off (or for a
 * target without FMA): a fused multiply-add would change the low bit.
 *
 * Scene (the shapes of the reference's sensor: 640 x 480 Kinect, fx = fy = 525, cx = 319.5, cy = 239.5 -
 * graph_slam_common/src/transformation/feature_transformation_estimator.cpp:37; depth in [0.5, 7] m - feature_max_depth,
 * iti_slam_launch/yaml/slam.yaml:7; missing depth is (0, 0, -1) with valid = 0 - feature_extraction_core.cpp:286-289):
 * keyframes come in clusters that share a pool of landmarks; a keyframe observes n_shared pool landmarks from its own pose
 * (within 15 degrees / 0.75 m of the cluster frame) plus fresh ones, with Kinect-like noise, descriptor bits flipped with
 * p = 1/16, invalid_frac of the rows without depth, rows shuffled.  Every keyframe gets k_candidates partners:
 * k_candidates - cross_cluster from its own cluster, the rest from other clusters (place-recognition false positives).
 */
#ifndef UZ_SYNTH_H
#define UZ_SYNTH_H

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct uz_synth_cfg {
    int32_t n_keyframes, n_features, cluster, pool, n_shared, k_candidates, cross_cluster, desc_bytes; /* desc_bytes: 32 or 64 */
    double invalid_frac;
    uint64_t seed;
} uz_synth_cfg;

static inline uz_synth_cfg uz_synth_c4(int32_t n_keyframes) { /* the map bench.py builds (BASELINE.json C3 / C4) */
    uz_synth_cfg c;
    c.n_keyframes = n_keyframes; c.n_features = 1000; c.cluster = 25; c.pool = 1000; c.n_shared = 600;
    c.k_candidates = 20; c.cross_cluster = 4; c.desc_bytes = 32; c.invalid_frac = 0.15; c.seed = 4;
    return c;
}

#define UZS_GAMMA 0x9E3779B97F4A7C15ull
#define UZS_FX 525.0
#define UZS_FY 525.0
#define UZS_CX 319.5
#define UZS_CY 239.5
#define UZS_COS_HALF_MAX 0.9914448613738104 /* cos(15 deg / 2) */
#define UZS_T_HALF 0.43                     /* translation cube half width: |t| <= 0.745 m */

static inline uint64_t uzs_mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t uzs_draw(uint64_t s, uint64_t j) { return uzs_mix(s + (j + 1) * UZS_GAMMA); }
static inline uint64_t uzs_sub(uint64_t s, uint64_t k) { return uzs_mix(s ^ uzs_mix(k + 0x632BE59BD9B4E019ull)); }
static inline double uzs_u01(uint64_t x) { return (double)(x >> 11) * 1.1102230246251565e-16; /* 2^-53 */ }
static inline double uzs_gauss(uint64_t s, uint64_t j) {
    double acc = uzs_u01(uzs_draw(s, 12 * j));
    for (int t = 1; t < 12; ++t) acc = acc + uzs_u01(uzs_draw(s, 12 * j + (uint64_t)t));
    return acc - 6.0;
}
static inline void uzs_landmark(uint64_t s, uint64_t j, double* X) {
    const double u = 640.0 * uzs_u01(uzs_draw(s, 3 * j));
    const double v = 480.0 * uzs_u01(uzs_draw(s, 3 * j + 1));
    const double z = 0.5 + 6.5 * uzs_u01(uzs_draw(s, 3 * j + 2));
    X[0] = ((u - UZS_CX) * z) / UZS_FX;
    X[1] = ((v - UZS_CY) * z) / UZS_FY;
    X[2] = z;
}

typedef struct { uint64_t key; int32_t idx; } uzs_keyed;
static int uzs_cmp(const void* a, const void* b) {
    const uzs_keyed* x = (const uzs_keyed*)a; const uzs_keyed* y = (const uzs_keyed*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}
/* order[] = indices 0..n-1 sorted by (draw(s, i), i) */
static inline void uzs_order(uint64_t s, int32_t n, uzs_keyed* tmp, int32_t* order) {
    for (int32_t i = 0; i < n; ++i) { tmp[i].key = uzs_draw(s, (uint64_t)i); tmp[i].idx = i; }
    qsort(tmp, (size_t)n, sizeof(uzs_keyed), uzs_cmp);
    for (int32_t i = 0; i < n; ++i) order[i] = tmp[i].idx;
}

static inline uint64_t uzs_root(const uz_synth_cfg* c) { return uzs_mix(0xC4000000ull + c->seed); }
static inline uint64_t uzs_keyframe_stream(const uz_synth_cfg* c, int32_t i) { return uzs_sub(uzs_root(c), 0x10000000ull + (uint64_t)i); }

/* cluster -> keyframe pose of keyframe i as a row-major 3 x 4 matrix [R | t] */
static inline void uz_synth_pose(const uz_synth_cfg* c, int32_t i, double* T12) {
    const uint64_t sp = uzs_sub(uzs_keyframe_stream(c, i), 0);
    const double g0 = uzs_gauss(sp, 0), g1 = uzs_gauss(sp, 1), g2 = uzs_gauss(sp, 2);
    const double n = sqrt((g0 * g0 + g1 * g1) + g2 * g2);
    double a0 = 1.0, a1 = 0.0, a2 = 0.0;
    if (n > 1e-12) { a0 = g0 / n; a1 = g1 / n; a2 = g2 / n; }
    const double w = 1.0 - uzs_u01(uzs_draw(sp, 100)) * (1.0 - UZS_COS_HALF_MAX);
    const double s2 = sqrt(1.0 - w * w);
    const double x = s2 * a0, y = s2 * a1, z = s2 * a2;
    T12[0] = 1.0 - 2.0 * ((y * y) + (z * z)); T12[1] = 2.0 * ((x * y) - (w * z));       T12[2] = 2.0 * ((x * z) + (w * y));
    T12[4] = 2.0 * ((x * y) + (w * z));       T12[5] = 1.0 - 2.0 * ((x * x) + (z * z)); T12[6] = 2.0 * ((y * z) - (w * x));
    T12[8] = 2.0 * ((x * z) - (w * y));       T12[9] = 2.0 * ((y * z) + (w * x));       T12[10] = 1.0 - 2.0 * ((x * x) + (y * y));
    for (int k = 0; k < 3; ++k) T12[4 * k + 3] = UZS_T_HALF * (2.0 * uzs_u01(uzs_draw(sp, 101 + (uint64_t)k)) - 1.0);
}

/* One keyframe: desc[n_features * desc_bytes], pos[n_features * 3] (row = one point: the memory of an Eigen 3 x N
 * column-major matrix), valid[n_features].  Returns 0, or -1 on a bad configuration / allocation failure. */
static inline int uz_synth_keyframe(const uz_synth_cfg* c, int32_t i, uint8_t* desc, double* pos, uint8_t* valid) {
    const int32_t N = c->n_features, P = c->pool, S = c->n_shared, W = c->desc_bytes / 8;
    if (N <= 0 || P <= 0 || S < 0 || S > P || S > N || (c->desc_bytes != 32 && c->desc_bytes != 64) || c->cluster <= 0) return -1;
    const uint64_t sc = uzs_sub(uzs_root(c), (uint64_t)(i / c->cluster));
    const uint64_t sk = uzs_keyframe_stream(c, i);
    const uint64_t s_px = uzs_sub(sc, 0), s_pd = uzs_sub(sc, 1);
    const uint64_t s_sel = uzs_sub(sk, 1), s_fresh = uzs_sub(sk, 2), s_flip = uzs_sub(sk, 3), s_fd = uzs_sub(sk, 4),
                   s_noise = uzs_sub(sk, 5), s_bad = uzs_sub(sk, 6), s_perm = uzs_sub(sk, 7);
    const int32_t big = P > N ? P : N;
    uzs_keyed* tmp = (uzs_keyed*)malloc((size_t)big * sizeof(uzs_keyed));
    int32_t* sel = (int32_t*)malloc((size_t)P * sizeof(int32_t));
    int32_t* perm = (int32_t*)malloc((size_t)N * sizeof(int32_t));
    double* X = (double*)malloc((size_t)N * 3 * sizeof(double));
    uint64_t* D = (uint64_t*)malloc((size_t)N * (size_t)W * sizeof(uint64_t));
    if (!tmp || !sel || !perm || !X || !D) { free(tmp); free(sel); free(perm); free(X); free(D); return -1; }
    double T[12];
    uz_synth_pose(c, i, T);
    uzs_order(s_sel, P, tmp, sel);
    for (int32_t r = 0; r < N; ++r) {
        double L[3];
        if (r < S) {
            double Xp[3];
            uzs_landmark(s_px, (uint64_t)sel[r], Xp);
            for (int k = 0; k < 3; ++k) L[k] = ((T[4 * k] * Xp[0] + T[4 * k + 1] * Xp[1]) + T[4 * k + 2] * Xp[2]) + T[4 * k + 3];
            for (int32_t w = 0; w < W; ++w) {
                const uint64_t at = ((uint64_t)r * (uint64_t)W + (uint64_t)w) * 4;
                const uint64_t m = uzs_draw(s_flip, at) & uzs_draw(s_flip, at + 1) & uzs_draw(s_flip, at + 2) & uzs_draw(s_flip, at + 3);
                D[(size_t)r * W + w] = uzs_draw(s_pd, (uint64_t)sel[r] * (uint64_t)W + (uint64_t)w) ^ m;
            }
        } else {
            uzs_landmark(s_fresh, (uint64_t)(r - S), L);
            for (int32_t w = 0; w < W; ++w) D[(size_t)r * W + w] = uzs_draw(s_fd, (uint64_t)(r - S) * (uint64_t)W + (uint64_t)w);
        }
        /* Kinect-like noise: sigma_z = 0.0012 z^2, lateral noise through the pinhole (0.5 px) */
        const double z = L[2];
        const double az = fabs(z), zs = az > 0.3 ? az : 0.3;
        const double sz = (0.0012 * zs) * zs;
        const double zn = z + uzs_gauss(s_noise, 3 * (uint64_t)r) * sz;
        const double scale = zn / (az > 1e-9 ? z : 1.0);
        X[3 * r] = L[0] * scale + ((uzs_gauss(s_noise, 3 * (uint64_t)r + 1) * 0.5) * zs) / UZS_FX;
        X[3 * r + 1] = L[1] * scale + ((uzs_gauss(s_noise, 3 * (uint64_t)r + 2) * 0.5) * zs) / UZS_FY;
        X[3 * r + 2] = zn;
    }
    uzs_order(s_perm, N, tmp, perm);
    for (int32_t p = 0; p < N; ++p) {
        const int32_t r = perm[p];
        memcpy(desc + (size_t)p * c->desc_bytes, D + (size_t)r * W, (size_t)c->desc_bytes);   /* little-endian words */
        if (uzs_u01(uzs_draw(s_bad, (uint64_t)r)) < c->invalid_frac) {
            pos[3 * p] = 0.0; pos[3 * p + 1] = 0.0; pos[3 * p + 2] = -1.0; valid[p] = 0;
        } else {
            pos[3 * p] = X[3 * r]; pos[3 * p + 1] = X[3 * r + 1]; pos[3 * p + 2] = X[3 * r + 2]; valid[p] = 1;
        }
    }
    free(tmp); free(sel); free(perm); free(X); free(D);
    return 0;
}

/* candidate partners of keyframe i: writes k <= k_candidates (from, to) pairs, returns k */
static inline int32_t uz_synth_candidates(const uz_synth_cfg* c, int32_t i, int32_t* pairs) {
    const uint64_t sk = uzs_keyframe_stream(c, i);
    const int32_t cl = i / c->cluster, c0 = cl * c->cluster;
    const int32_t nk = (c0 + c->cluster <= c->n_keyframes ? c->cluster : c->n_keyframes - c0);
    uzs_keyed* tmp = (uzs_keyed*)malloc((size_t)(nk > 0 ? nk : 1) * sizeof(uzs_keyed));
    int32_t* ord = (int32_t*)malloc((size_t)(nk > 0 ? nk : 1) * sizeof(int32_t));
    if (!tmp || !ord) { free(tmp); free(ord); return -1; }
    uzs_order(uzs_sub(sk, 8), nk, tmp, ord);
    int32_t want_own = c->k_candidates - c->cross_cluster, k = 0;
    if (want_own > nk - 1) want_own = nk - 1;
    for (int32_t j = 0; j < nk && k < want_own; ++j) {
        if (c0 + ord[j] == i) continue;
        pairs[2 * k] = i; pairs[2 * k + 1] = c0 + ord[j]; ++k;
    }
    if (c->n_keyframes > c->cluster) {
        const uint64_t sx = uzs_sub(sk, 9);
        for (uint64_t t = 0; k < c->k_candidates; ++t) {
            const int32_t j = (int32_t)(uzs_draw(sx, t) % (uint64_t)c->n_keyframes);
            if (j / c->cluster == cl) continue;
            pairs[2 * k] = i; pairs[2 * k + 1] = j; ++k;
        }
    }
    free(tmp); free(ord);
    return k;
}

/* whole map into caller-provided arrays (desc n_kf * n_feat * desc_bytes, pos n_kf * n_feat * 3, valid n_kf * n_feat,
 * pairs n_kf * k_candidates * 2); returns the number of pairs or -1 */
static inline int64_t uz_synth_map(const uz_synth_cfg* c, uint8_t* desc, double* pos, uint8_t* valid, int32_t* pairs) {
    int64_t n_pairs = 0;
    for (int32_t i = 0; i < c->n_keyframes; ++i) {
        const size_t at = (size_t)i * (size_t)c->n_features;
        if (uz_synth_keyframe(c, i, desc + at * (size_t)c->desc_bytes, pos + at * 3, valid + at) != 0) return -1;
    }
    for (int32_t i = 0; i < c->n_keyframes; ++i) {
        const int32_t k = uz_synth_candidates(c, i, pairs + 2 * n_pairs);
        if (k < 0) return -1;
        n_pairs += k;
    }
    return n_pairs;
}

/* checksum of a byte range (sum over the little-endian 8-byte words w_i, the last one zero padded, of mix(w_i + (i + 1) * GAMMA)):
 * what the C++ and Python sides print to show they hold the same map; vectorises in numpy */
static inline uint64_t uz_synth_checksum(const void* p, size_t n) {
    const uint8_t* b = (const uint8_t*)p;
    uint64_t h = 0;
    size_t i = 0;
    for (; 8 * i + 8 <= n; ++i) { uint64_t w; memcpy(&w, b + 8 * i, 8); h += uzs_mix(w + (uint64_t)(i + 1) * UZS_GAMMA); }
    if (8 * i < n) { uint64_t w = 0; memcpy(&w, b + 8 * i, n - 8 * i); h += uzs_mix(w + (uint64_t)(i + 1) * UZS_GAMMA); }
    return h;
}

#ifdef __cplusplus
}
#endif
#endif
