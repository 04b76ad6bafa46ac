// uz_solve.cuh — K2..K5: everything after matching, one CTA per keyframe pair, the pair kept on-chip.
//
// Follows /root/reference/transformation_estimation/src/feature_transformation_estimator.cpp:
//   K2  :65-71 ratio test, :74-86 best camera pair (score = #ratio survivors, first strict max),
//       :103-112 valid_3d filter, :114 sort (total order (distance, queryIdx); done as a
//       stable counting sort by distance over the query-ordered matches),
//       :118-124 gather of Pd (to) / Xd (from)
//   K3  :214-227 hypotheses from the shared sample-index list, estimatePoseSVD :299-314 (uz_arith.cuh)
//   K4  :230-241 consensus3D :337-347 per hypothesis, strict-'>' running maximum, early break
//   K5  :246-258 refit on the winner's inliers + recount, :285-290 mse, :133-137 information scale,
//       transformation_estimator.cpp:53-55 failure convention.
// Sequential semantics are reproduced exactly: hypotheses are evaluated in chunks of THREADS, the
// chunk's counts are scanned in iteration order by one thread, and evaluation stops at the first
// chunk that contains the reference's break iteration.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/uzliti_edge.h"
#include "uz_arith.cuh"
#include "uz_knn2.cuh"

namespace uz {

struct SolveParams {
    double thr;              // ransac_threshold
    double thr_sq_star;      // smallest double s with sqrt_rn(s) >= thr:  sqrt(s) < thr  <=>  s < thr_sq_star
    double break_pct;
    int32_t iterations;
    int32_t ratio_num, ratio_den;
    int32_t cap;             // power of two >= max nq of the launch: capacity of the on-chip arrays
    const uint16_t* samples; // sample table [M][iterations][3] (samples_by_m) or one list [iterations][3]
    int32_t samples_by_m;
    // direct mode (uz_estimate_svd / uz_consensus3d): P,Q given, stages K2 skipped
    const double* direct_P; const double* direct_Q; int32_t direct_M;
    const int32_t* direct_offsets;   // batched direct mode: problem b owns points [offsets[b], offsets[b+1]) of P,Q; null = one problem
    // parity taps (may be null)
    int32_t* dbg_matches;    // [pair][cap][3]
    uint8_t* dbg_mask;       // [pair][cap]
    int32_t* dbg_counts;     // [pair][iterations] consensus count of every evaluated hypothesis, -1 = not run
    long long* dbg_phase;    // [pair][8] clock64() at the phase boundaries of the CTA (profiling tap)
    int32_t pair_base;       // index of this launch's first pair inside the batch (chunked launches)
    int32_t dbg_skip;        // UZ_STREAM_PROBE=1 (measurement only): the streaming grid draws pairs but does not solve them
};

constexpr int kSolveThreads = 128;
constexpr int kHistPerLane = 17;                   // counting sort by distance: 0..512 (64-byte descriptors) in 32 x 17 bins
constexpr int kHistBins = 32 * kHistPerLane;
static_assert(kHistBins > 8 * UZ_MAX_DESC_BYTES, "one bin per possible Hamming distance");
static_assert((kSolveThreads / 32 + 1) * kHistBins * 4 <= kSolveThreads * 12 * 8, "the histograms live in the hypothesis buffer");
#ifndef UZ_SOLVE_H
#define UZ_SOLVE_H 2
#endif
#ifndef UZ_STREAM_H
#define UZ_STREAM_H 2
#endif
#ifndef UZ_SOLVE_MINB
#define UZ_SOLVE_MINB 5
#endif

// A wide CTA (kSolveThreadsWide) serves launches of at most one pair per SM - the online case - where the chip is
// otherwise idle and only the latency of the one CTA per pair counts: 16 warps score hypotheses instead of 4.
constexpr int kSolveThreadsWide = 512;
__host__ __device__ constexpr size_t solve_smem_bytes(int cap, int threads = kSolveThreads) {
    return (size_t)cap * 24 /*float32 copies of P,Q (also: unsorted keys, later the residual norms)*/ +
           (size_t)cap * 4 /*sorted keys, then packed (trainIdx<<16)|queryIdx*/ + (size_t)cap * 2 /*ordered inlier list*/ +
           (size_t)threads * 12 * 8 /*hypothesis transforms*/ + (size_t)threads * 4 /*counts*/ +
           256 /*Tbest, Tfin*/;
}

static_assert(solve_smem_bytes(UZ_MAX_FEATURES) + 64 <= 232448, "solve kernel exceeds the 227 KB per-CTA shared memory of sm_100");
static_assert(solve_smem_bytes(UZ_MAX_FEATURES, kSolveThreadsWide) + 256 <= 232448, "wide solve CTA exceeds the 227 KB per-CTA shared memory");

__device__ __forceinline__ void write_identity(double* T16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) T16[i] = (i % 5 == 0) ? 1.0 : 0.0;
}

// ---- K4 helper: float32 pre-screen of the consensus test ---------------------------------------------
// The reference test is  sqrt_d(s_d) < thr  with s_d evaluated in double (uz::residual_sq).  Evaluating the
// same residual in float32 (FMA allowed) from float-rounded points gives s_f with, per component,
//   |d_f - d| <= u (4 L1(p) + 3 |t|_inf + |q|_inf + |d|),  u = 2^-24,
// hence | sqrt(s_f) - ||d|| | <= m := K (4.5 L1(p) + 1.5 L1(q) + 3.5 L1(t)),  K = 1.2e-7 > sqrt(3) u (1 + slack);
// the kernel uses ONE margin per hypothesis chunk (max over the pair's points + max over the chunk's
// translations), so the loop compares against two CTA-uniform thresholds.
// A point is a certain inlier if sqrt(s_f) < thr(1-1e-6) - m and a certain outlier if sqrt(s_f) > thr(1+1e-6) + m.
// Per hypothesis the loop counts the certain inliers and the not-certainly-outside points; when the two counts differ
// (a borderline evaluation: about 1e-6 of them) the hypothesis is recounted exactly in double.  A NaN is outside in
// both views, exactly as in the double evaluation.  The result is therefore bit-identical to the double-only
// evaluation at a fraction of its pipe time.
constexpr float kScreenK = 1.2e-7f;

// K4 runs on packed float32 pairs (FFMA2 / FADD2 / FMUL2, sm_100): a lane scores TWO points per instruction.  The
// float32 copies are therefore stored pair-interleaved: logical point i lives at pidx(i), so that the 8-byte word at
// [64 b + 2 l] holds points 64 b + l and 64 b + 32 + l - one conflict-free LDS.64 per lane and coordinate.  The arrays
// are padded to a multiple of 64 with a point that is outside for every finite transform (p = 0, q = 1e18).
__device__ __forceinline__ int pidx(int i) { return (i & ~63) | ((i & 31) << 1) | ((i >> 5) & 1); }
constexpr float kPadQ = 1.0e18f;

// Counting without the ALU pipe (which the match kernel running beside this one saturates): for a power of two BIG,
//   fma.rn.sat(s, -BIG, c * BIG)  ==  1.0f if s < c, else 0.0f     (c * BIG exact; NaN and -inf saturate to 0)
// as long as one ulp of c times BIG is >= 1 and c * BIG is finite - solve_pair checks that and otherwise sends every
// hypothesis through the exact path.
constexpr float kSatBig = 1.2676506e30f;      // 2^100
__device__ __forceinline__ float sat_less(float s, float c_big) {
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(s), "f"(-kSatBig), "f"(c_big));
    return r;
}

#ifdef UZ_K4_SCALAR      // measured alternative: the same arithmetic as scalar FFMA / FADD / FMUL (twice the instructions):
                         // solve alone 2.88 -> 3.07 ms, streaming step 26.20 -> 26.44 ms
__device__ __forceinline__ float2 k4_fma(float2 a, float2 b, float2 c) { return make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 k4_add(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 k4_mul(float2 a, float2 b) { return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
#else
__device__ __forceinline__ float2 k4_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 k4_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 k4_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
#endif

// Rare path of the pre-screen, deliberately not inlined: it must not drag the double-precision transforms
// into registers inside the float loop.
__device__ __noinline__ int exact_inlier(const double* T12, const double* gP, const double* gQ, uint32_t tq, double thr_sq_star) {
    const double* p = gP + 3 * (tq & 0xFFFFu);
    const double* q = gQ + 3 * (tq >> 16);
    return residual_sq(T12, p[0], p[1], p[2], q[0], q[1], q[2]) < thr_sq_star;
}

// ---- edge acceptance gate ----------------------------------------------------------------
// GraphSlamNode::newEdgeCallback (/root/reference/graph_slam/src/graph_slam_node.cpp:798-804): an edge is linked only
// if matching_score_ >= min_matching_score, |t| <= max_edge_distance_T and the rotation angle of transform_.linear()
// (Eigen::AngleAxisd(R).angle(), degrees) <= max_edge_distance_R.  Eigen 3.2 goes matrix -> quaternion (Shoemake,
// Quaternion.h quaternionbase_assign_impl<Other,3,3>) -> angle = 2 acos(clamp(w)) (AngleAxis.h), 0 when |vec|^2 is
// below dummy_precision^2.
struct GateParams { double min_matching_score, max_edge_distance_T, max_edge_distance_R; };

__device__ __forceinline__ double rotation_angle_deg(const double* T /* row-major 4x4 */) {
    const double m00 = T[0], m01 = T[1], m02 = T[2], m10 = T[4], m11 = T[5], m12 = T[6], m20 = T[8], m21 = T[9], m22 = T[10];
    double w, x, y, z;
    double t = UZ_DADD(UZ_DADD(m00, m11), m22);
    if (t > 0.0) {
        t = UZ_DSQRT(UZ_DADD(t, 1.0));
        w = UZ_DMUL(0.5, t);
        t = UZ_DDIV(0.5, t);
        x = UZ_DMUL(UZ_DSUB(m21, m12), t); y = UZ_DMUL(UZ_DSUB(m02, m20), t); z = UZ_DMUL(UZ_DSUB(m10, m01), t);
    } else {
        const double m[3][3] = {{m00, m01, m02}, {m10, m11, m12}, {m20, m21, m22}};
        int i = 0;
        if (m11 > m00) i = 1;
        if (m22 > m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = UZ_DSQRT(UZ_DADD(UZ_DSUB(UZ_DSUB(m[i][i], m[j][j]), m[k][k]), 1.0));
        double q[3];
        q[i] = UZ_DMUL(0.5, t);
        t = UZ_DDIV(0.5, t);
        w = UZ_DMUL(UZ_DSUB(m[k][j], m[j][k]), t);
        q[j] = UZ_DMUL(UZ_DADD(m[j][i], m[i][j]), t);
        q[k] = UZ_DMUL(UZ_DADD(m[k][i], m[i][k]), t);
        x = q[0]; y = q[1]; z = q[2];
    }
    const double n2 = UZ_DADD(UZ_DADD(UZ_DMUL(x, x), UZ_DMUL(y, y)), UZ_DMUL(z, z));
    if (n2 < 1e-12 * 1e-12) return 0.0;                           // NumTraits<double>::dummy_precision()^2
    const double wc = fmin(fmax(-1.0, w), 1.0);
    const double angle = 2.0 * acos(wc);
    return fabs(angle) * 180.0 / 3.14159265358979323846;
}

__global__ void __launch_bounds__(256) gate_edges_kernel(const uz_edge_result* __restrict__ res, int n, GateParams g,
                                                         uint8_t* __restrict__ accept, double* __restrict__ tnorm_out,
                                                         double* __restrict__ rot_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uz_edge_result* r = res + i;
    const double tx = r->T[3], ty = r->T[7], tz = r->T[11];
    const double tn = UZ_DSQRT(UZ_DADD(UZ_DADD(UZ_DMUL(tx, tx), UZ_DMUL(ty, ty)), UZ_DMUL(tz, tz)));
    const double rot = rotation_angle_deg(r->T);
    // estimateEdgeImpl false => matching_score_ = 0 (transformation_estimator.cpp:53-55); consensus IS the score (:155)
    const double score = r->ok ? (double)r->consensus : 0.0;
    accept[i] = (score >= g.min_matching_score && tn <= g.max_edge_distance_T && rot <= g.max_edge_distance_R) ? 1 : 0;
    if (tnorm_out) tnorm_out[i] = tn;
    if (rot_out) rot_out[i] = rot;
}

// block-wide maximum of a non-negative float (NaN contributions are ignored by fmaxf)
template <int THREADS>
__device__ __forceinline__ float block_max(float v, float* s_red /* [THREADS/32] */, int tid) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    float r = s_red[0];
#pragma unroll
    for (int w = 1; w < THREADS / 32; ++w) r = fmaxf(r, s_red[w]);
    return r;
}

// Ratio test (:65-71) of query row q, plus the opt-in cross-check (uz_params.cross_check; the reference has none):
// the survivor (q, t = best train of q) is kept only if q is also the best query of t in the reversed matching,
// ties by lowest index - i.e. (q, t) is what cv::BFMatcher(NORM_HAMMING, crossCheck=true).match() returns for q.
// Keys are always loaded with ld.global.cg (L2 only): in streaming mode they were written, moments ago, by the match
// kernel running beside this one, so neither the non-coherent path nor a stale L1 sector may serve them.
__device__ __forceinline__ uint2 load_key(const uint2* p) { return __ldcg(p); }

__device__ __forceinline__ bool match_survives(const uint2 m, int q, int ratio_num, int ratio_den,
                                               const uint2* keys, uint32_t rev_key_off) {
    bool pass = (m.y != kNoKey) && ((int)(m.x >> 16) * ratio_den < (int)(m.y >> 16) * ratio_num);
    if (pass && rev_key_off != kNoRev) pass = (int)(load_key(keys + rev_key_off + (m.x & 0xFFFFu)).x & 0xFFFFu) == q;
    return pass;
}

// Everything after matching for ONE pair (or one direct problem); called by all THREADS threads of a CTA.  Every
// early return below is CTA-uniform.
template <int THREADS, int H /* hypotheses scored per pass over the points (each lane: 2 points x H hypotheses) */>
__device__ __forceinline__ void solve_pair(const MatchTask* __restrict__ tasks, const int2* __restrict__ pair_tasks,
                                           const uint2* keys, const SolveParams& prm,
                                           uz_edge_result* __restrict__ results, const int pair) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int cap = prm.cap;
    // On-chip state of the pair.  Only float32 copies of the matched 3-D points live here (the scoring loop
    // and the float32 pose recurrences read nothing else); the float64 originals stay in the keyframe store
    // and are fetched through the packed (trainIdx<<16)|queryIdx list where exactness needs them.
    double* Th = reinterpret_cast<double*>(smem_raw);                 // [THREADS][12] hypothesis transforms
    double* Tbest = Th + THREADS * 12;                                // 12
    double* Tfin = Tbest + 12;                                        // 12
    float* pf = reinterpret_cast<float*>(Tfin + 12);                  // [6][cap]: px py pz (to) qx qy qz (from)
    float* pxf = pf; float* pyf = pf + cap; float* pzf = pf + 2 * cap;
    float* qxf = pf + 3 * cap; float* qyf = pf + 4 * cap; float* qzf = pf + 5 * cap;
    double* norms = reinterpret_cast<double*>(pf);                    // residual norms reuse pf after the last pass
    uint32_t* skeys = reinterpret_cast<uint32_t*>(pf + 6 * cap);      // [cap] sorted keys -> (t<<16)|q -> inlier list
    uint32_t* tq = skeys;
    int32_t* counts = reinterpret_cast<int32_t*>(skeys + cap);        // [THREADS]
    uint16_t* ilist = reinterpret_cast<uint16_t*>(counts + THREADS);  // [cap] winner's inliers in index order
    const double* __restrict__ gP = nullptr;                          // float64 positions of the to-camera (by queryIdx)
    const double* __restrict__ gQ = nullptr;                          // float64 positions of the from-camera (by trainIdx)
    __shared__ int s_best, s_maxc, s_break, s_run, s_nvalid, s_nratio;
    constexpr int NW = THREADS / 32;
    __shared__ int s_wcnt[NW];
    __shared__ float s_red[NW];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    // The record is put together in shared memory by one thread and leaves as eleven 16-byte stores: `results` may be peer
    // memory (NVLink) or mapped host memory (PCIe) in group mode, where 22 separate 8-byte stores per record are what costs.
    __shared__ __align__(16) uz_edge_result s_rec;
    uz_edge_result* const out_rec = results + pair;    // batch-wide index; pair_tasks/results are indexed by it
    uz_edge_result* res = &s_rec;
    static_assert(sizeof(uz_edge_result) == 11 * 16, "the record leaves as eleven uint4");
#define UZ_PUBLISH_RECORD()                                                                                             \
    do {                                                                                                                \
        __syncthreads();                                                                                                \
        if (tid < 11) reinterpret_cast<uint4*>(out_rec)[tid] = reinterpret_cast<const uint4*>(&s_rec)[tid];            \
    } while (0)
#define UZ_PHASE(k) do { if (prm.dbg_phase && tid == 0) prm.dbg_phase[(size_t)pair * 8 + (k)] = clock64(); } while (0)
    UZ_PHASE(0);
    int M = 0;
    int n_ratio = 0, cam_from = -1, cam_to = -1;
    float kp_local = 0.f;        // per-thread max of the point part of the pre-screen error bound

    if (prm.direct_P == nullptr) {
        // ---------------- K2: best camera pair, filter, sort, gather -------------------------------
        const int2 pt = pair_tasks[pair];
        int best = pt.y == 1 ? 0 : -1;
        if (pt.y > 1) {          // rigs: score every same-frame camera pair first (:74-86)
            int best_score = -1;
            for (int t = 0; t < pt.y; ++t) {
                const MatchTask* tk = tasks + pt.x + t;
                const uint2* k = keys + tk->key_off;
                int cnt = 0;
                for (int base = 0; base < tk->nq; base += THREADS) {
                    const int q = base + tid;
                    bool pass = false;
                    if (q < tk->nq) pass = match_survives(load_key(k + q), q, prm.ratio_num, prm.ratio_den, keys, tk->rev_key_off);
                    cnt += __syncthreads_count(pass);
                }
                if (cnt > best_score) { best_score = cnt; best = t; }   // :81 strict '>' keeps the first
            }
        }
        if (best < 0) {          // :93-95 no comparable camera pair
            if (tid == 0) {
                res->ok = 0; res->cam_from = -1; res->cam_to = -1; res->n_ratio_matches = 0; res->n_matches = 0;
                res->consensus = 0; res->best_iteration = -1; res->iterations_run = 0; res->mse = 0.0;
                res->info_scale = 1.0; write_identity(res->T);
            }
            UZ_PUBLISH_RECORD();
            return;
        }
        const MatchTask* tk = tasks + pt.x + best;
        const uint2* k = keys + tk->key_off;
        const int nq = tk->nq;
        cam_from = tk->cam_from; cam_to = tk->cam_to;
        const uint8_t* __restrict__ vq = tk->q_valid;
        const uint8_t* __restrict__ vt = tk->t_valid;
        // ratio test (:65-71) + valid_3d filter (:103-112) + the sort of :114 as a STABLE COUNTING SORT by
        // distance: matches are produced in query order, so equal distances keep ascending queryIdx, which
        // is exactly the (distance, queryIdx) order.  Pass 1 (all warps): per-query distance + histogram.
        // The histogram is kept per PARTITION of the query rows (one partition of `part` consecutive rows per warp), so that
        // the placement pass below runs on all warps at once: hist[w][d], then tot[d]; Th is not used before K3.
        int* hist = reinterpret_cast<int*>(Th);                 // [NW][kHistBins] + [kHistBins]
        int* tot = hist + NW * kHistBins;
        static_assert((NW + 1) * kHistBins * 4 <= THREADS * 12 * 8, "the histograms live in the hypothesis buffer");
        uint16_t* dq = reinterpret_cast<uint16_t*>(pf);         // [cap] distance of query i, 0xFFFF = dropped
        const int part = (((nq + NW - 1) / NW) + 31) & ~31;     // rows per partition, a multiple of 32
        for (int bidx = tid; bidx < (NW + 1) * kHistBins; bidx += THREADS) hist[bidx] = 0;
        if (tid == 0) { s_nvalid = 0; s_nratio = 0; }
        __syncthreads();
        // Pass 1 (all warps): ratio test, valid flags, distance + histogram.  The loads of a thread's rows go out together,
        // level by level (row keys + query flags, then the dependent cross-check keys and train flags): two global-memory
        // latencies per four rows instead of two per row.
        {
            constexpr int U = 4;
            int my_ratio = 0;
            for (int base = 0; base < nq; base += THREADS * U) {
                uint2 m[U];
                uint8_t a[U], b[U];
                uint32_t rk[U];
                bool pass[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = base + u * THREADS + tid;
                    m[u] = i < nq ? load_key(k + i) : make_uint2(kNoKey, kNoKey);
                    a[u] = i < nq ? vq[i] : (uint8_t)0;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = base + u * THREADS + tid;
                    pass[u] = i < nq && m[u].y != kNoKey && ((int)(m[u].x >> 16) * prm.ratio_den < (int)(m[u].y >> 16) * prm.ratio_num);
                    rk[u] = (uint32_t)i;
                    b[u] = 0;
                    if (pass[u]) {
                        if (tk->rev_key_off != kNoRev) rk[u] = load_key(keys + tk->rev_key_off + (m[u].x & 0xFFFFu)).x & 0xFFFFu;
                        if (a[u]) b[u] = vt[m[u].x & 0xFFFFu];
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = base + u * THREADS + tid;
                    if (i >= nq) continue;
                    uint16_t d = 0xFFFFu;
                    if (pass[u] && (int)rk[u] == i) {           // (without the cross-check rk is i itself)
                        ++my_ratio;
                        if (a[u] && b[u]) { d = (uint16_t)(m[u].x >> 16); atomicAdd(&hist[(i / part) * kHistBins + d], 1); }
                    }
                    dq[i] = d;
                }
            }
            if (my_ratio) atomicAdd(&s_nratio, my_ratio);
        }
        __syncthreads();
        n_ratio = s_nratio;
        UZ_PHASE(1);
        // per bin: where each partition starts inside the bin, and the bin's total
        for (int bidx = tid; bidx < kHistBins; bidx += THREADS) {
            int run = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) { const int c = hist[w * kHistBins + bidx]; hist[w * kHistBins + bidx] = run; run += c; }
            tot[bidx] = run;
        }
        __syncthreads();
        if (warp == 0) {
            // exclusive prefix over the 513 bins: kHistPerLane consecutive bins per lane
            int loc[kHistPerLane], sum = 0;
#pragma unroll
            for (int j = 0; j < kHistPerLane; ++j) { loc[j] = tot[lane * kHistPerLane + j]; sum += loc[j]; }
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
            int run = inc - sum;
#pragma unroll
            for (int j = 0; j < kHistPerLane; ++j) { tot[lane * kHistPerLane + j] = run; run += loc[j]; }
            if (lane == 31) s_nvalid = inc;
        }
        __syncthreads();
        for (int bidx = tid; bidx < NW * kHistBins; bidx += THREADS) hist[bidx] += tot[bidx % kHistBins];
        __syncthreads();
        {
            // pass 2 (every warp its partition, in query order): slot = bin offset + rank among equal distances seen so far
            int* myhist = hist + warp * kHistBins;
            const int q_end = min(nq, (warp + 1) * part);
            for (int base = warp * part; base < q_end; base += 32) {
                const int q = base + lane;
                const unsigned d = q < q_end ? dq[q] : 0xFFFFu;
                const unsigned peers = __match_any_sync(0xffffffffu, d);
                const int rank = __popc(peers & ((1u << lane) - 1u));
                if (d != 0xFFFFu) skeys[myhist[d] + rank] = (d << 16) | (unsigned)q;
                __syncwarp();
                if (d != 0xFFFFu && rank == 0) myhist[d] += __popc(peers);
                __syncwarp();
            }
        }
        __syncthreads();
        M = s_nvalid;
        UZ_PHASE(2);
        gP = tk->q_pos;
        gQ = tk->t_pos;
        {
            // gather of the matched 3-D points, two matches per thread at a time: train indices first, then all 12 doubles
            constexpr int U = 2;
            for (int base = 0; base < M; base += THREADS * U) {
                uint32_t key[U];
                int tt[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = base + u * THREADS + tid;
                    key[u] = i < M ? skeys[i] : 0u;
                    tt[u] = i < M ? (int)(load_key(k + (key[u] & 0xFFFFu)).x & 0xFFFFu) : 0;
                }
                double pv[U][6];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = base + u * THREADS + tid;
                    const int q = key[u] & 0xFFFFu;
                    if (i < M) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) { pv[u][c] = gP[3 * q + c]; pv[u][3 + c] = gQ[3 * tt[u] + c]; }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = base + u * THREADS + tid;
                    if (i >= M) continue;
                    const int q = key[u] & 0xFFFFu, t = tt[u];
                    const float x = (float)pv[u][0], y = (float)pv[u][1], z = (float)pv[u][2];
                    const float uu = (float)pv[u][3], v = (float)pv[u][4], w = (float)pv[u][5];
                    const int pi = pidx(i);
                    pxf[pi] = x; pyf[pi] = y; pzf[pi] = z; qxf[pi] = uu; qyf[pi] = v; qzf[pi] = w;
                    kp_local = fmaxf(kp_local, 4.5f * (fabsf(x) + fabsf(y) + fabsf(z)) + 1.5f * (fabsf(uu) + fabsf(v) + fabsf(w)));
                    tq[i] = ((uint32_t)t << 16) | (uint32_t)q;
                    if (prm.dbg_matches) {
                        int32_t* d = prm.dbg_matches + ((size_t)pair * cap + i) * 3;
                        d[0] = q; d[1] = t; d[2] = (int)(key[u] >> 16);
                    }
                }
            }
        }
    } else {
        M = prm.direct_M;
        gP = prm.direct_P;
        gQ = prm.direct_Q;
        if (prm.direct_offsets) {                     // calcValidEdges-style batch: one CTA per cluster
            const int o0 = prm.direct_offsets[pair], o1 = prm.direct_offsets[pair + 1];
            M = o1 - o0; gP += 3 * (size_t)o0; gQ += 3 * (size_t)o0;
        }
        for (int i = tid; i < M; i += THREADS) {
            const float x = (float)gP[3 * i], y = (float)gP[3 * i + 1], z = (float)gP[3 * i + 2];
            const float u = (float)gQ[3 * i], v = (float)gQ[3 * i + 1], w = (float)gQ[3 * i + 2];
            const int pi = pidx(i);
            pxf[pi] = x; pyf[pi] = y; pzf[pi] = z; qxf[pi] = u; qyf[pi] = v; qzf[pi] = w;
            kp_local = fmaxf(kp_local, 4.5f * (fabsf(x) + fabsf(y) + fabsf(z)) + 1.5f * (fabsf(u) + fabsf(v) + fabsf(w)));
            tq[i] = ((uint32_t)i << 16) | (uint32_t)i;
        }
    }
    for (int i = M + tid; i < ((M + 63) & ~63); i += THREADS) {      // padding of the last 64-point block
        const int pi = pidx(i);
        pxf[pi] = 0.f; pyf[pi] = 0.f; pzf[pi] = 0.f; qxf[pi] = kPadQ; qyf[pi] = kPadQ; qzf[pi] = kPadQ;
    }
    if (tid == 0) { s_best = -1; s_maxc = 0; s_break = 0; s_run = 0; }
    const float kp_max = block_max<THREADS>(kp_local, s_red, tid);     // (also the barrier after the gather)
    UZ_PHASE(3);
    if (prm.dbg_skip == 2) return;

    if (M < 3) {                 // :118/:158 not enough depth-valid matches
        if (tid == 0) {
            res->ok = 0; res->cam_from = cam_from; res->cam_to = cam_to; res->n_ratio_matches = n_ratio;
            res->n_matches = M; res->consensus = 0; res->best_iteration = -1; res->iterations_run = 0;
            res->mse = 0.0; res->info_scale = 1.0; write_identity(res->T);
        }
        UZ_PUBLISH_RECORD();
        if (prm.dbg_mask) for (int i = tid; i < M; i += THREADS) prm.dbg_mask[(size_t)pair * cap + i] = 0;
        return;
    }

    // ---------------- K3/K4: hypotheses, consensus counts, sequential-semantics winner -------------
    const int I = prm.iterations;
    const uint16_t* __restrict__ samp = prm.samples + (prm.samples_by_m ? (size_t)M * I * 3 : 0);
    const float thr_dn = __double2float_rd(prm.thr * (1.0 - 1e-6));
    const float thr_up = __double2float_ru(prm.thr * (1.0 + 1e-6));
    const int M64 = (M + 63) & ~63;
    for (int h0 = 0; h0 < I; h0 += THREADS) {
        const int nh = min(THREADS, I - h0);
        float kt_local = 0.f;
        if (tid < nh) {
            PoseAcc acc;
            pose_reset(acc);
#pragma unroll 1
            for (int j = 0; j < 3; ++j) {
                const int s = pidx(samp[(size_t)(h0 + tid) * 3 + j]);
                pose_add(acc, pxf[s], pyf[s], pzf[s], qxf[s], qyf[s], qzf[s]);       // == (float) of the doubles (:303-304)
            }
            pose_finish(acc, Th + tid * 12);
            kt_local = 3.5f * (fabsf((float)Th[tid * 12 + 3]) + fabsf((float)Th[tid * 12 + 7]) + fabsf((float)Th[tid * 12 + 11]));
        }
        // one screening margin for the whole chunk: the largest point bound plus the largest translation bound
        const float m_scr = kScreenK * (kp_max + block_max<THREADS>(kt_local, s_red, tid));     // (barrier inside)
        const float lo = thr_dn - m_scr, hi = thr_up + m_scr;
        const float lo2 = lo > 0.f ? lo * lo : -1.f;      // sf < lo2  =>  certainly inside  (never if lo <= 0)
        const float hi2 = nextafterf(hi * hi, INFINITY);  // sf >= hi2 =>  certainly outside
        // the saturating-FMA counters need lo2 * 2^100 exact and finite with ulp(lo2) * 2^100 >= 1; outside that range
        // of thresholds (below ~1e-11 m or above ~1 km), or with a non-finite margin, every hypothesis is recounted exactly
        const bool screen_ok = lo2 > 1.4e-23f && hi2 < 1.0e6f;
        const float lo2_big = lo2 * kSatBig, hi2_big = hi2 * kSatBig;
        if (h0 == 0) UZ_PHASE(4);
        if (prm.dbg_skip == 4) return;
        for (int hb = warp * H; hb < nh; hb += NW * H) {
            float2 T[H][12];
            float2 in_c[H], in_b[H];      // per lane and point slot: certain inliers / not certainly outside
            // -T (exact: the doubles hold float32 values), so that d = q - T p needs no negation of q.  The double -> float
            // conversions run on the XU pipe, which the match kernel beside this one saturates: lane l converts ONE of the
            // 12 H elements and the warp shares them by shuffle (1 conversion per pass instead of 12 H).
            static_assert(12 * H <= 32, "one transform element per lane");
            float t_lane;
            {
                const int a_l = min(lane / 12, H - 1), e_l = lane % 12;
                t_lane = -(float)Th[min(hb + a_l, nh - 1) * 12 + e_l];          // tail: duplicates, their counts are discarded
            }
#pragma unroll
            for (int a = 0; a < H; ++a) {
#pragma unroll
                for (int e = 0; e < 12; ++e) { const float t = __shfl_sync(0xffffffffu, t_lane, a * 12 + e); T[a][e] = make_float2(t, t); }
                in_c[a] = make_float2(0.f, 0.f); in_b[a] = make_float2(0.f, 0.f);
            }
            for (int base = 0; base < M64; base += 64) {      // branch-free: 64 points x H hypotheses per warp step
                const int o = base + 2 * lane;
                const float2 x = *reinterpret_cast<const float2*>(pxf + o), y = *reinterpret_cast<const float2*>(pyf + o);
                const float2 z = *reinterpret_cast<const float2*>(pzf + o), u = *reinterpret_cast<const float2*>(qxf + o);
                const float2 v = *reinterpret_cast<const float2*>(qyf + o), w = *reinterpret_cast<const float2*>(qzf + o);
#pragma unroll
                for (int a = 0; a < H; ++a) {
                    const float2 dx = k4_add(k4_fma(T[a][0], x, k4_fma(T[a][1], y, k4_fma(T[a][2], z, T[a][3]))), u);
                    const float2 dy = k4_add(k4_fma(T[a][4], x, k4_fma(T[a][5], y, k4_fma(T[a][6], z, T[a][7]))), v);
                    const float2 dz = k4_add(k4_fma(T[a][8], x, k4_fma(T[a][9], y, k4_fma(T[a][10], z, T[a][11]))), w);
                    const float2 sf = k4_fma(dx, dx, k4_fma(dy, dy, k4_mul(dz, dz)));
                    in_c[a] = k4_add(in_c[a], make_float2(sat_less(sf.x, lo2_big), sat_less(sf.y, lo2_big)));
                    in_b[a] = k4_add(in_b[a], make_float2(sat_less(sf.x, hi2_big), sat_less(sf.y, hi2_big)));
                }
            }
#pragma unroll
            for (int a = 0; a < H; ++a) {
                // both counts in one register (each <= 4096 over the warp): certain | (not certainly outside) << 16.
                // float -> int without the XU pipe: the counts are small integers, so the low mantissa bits of
                // (count + 2^23) ARE the count
                const int cc = __float_as_int(__fadd_rn(__fadd_rn(in_c[a].x, in_c[a].y), 8388608.0f)) & 0x7FFFFF;
                const int cb = __float_as_int(__fadd_rn(__fadd_rn(in_b[a].x, in_b[a].y), 8388608.0f)) & 0x7FFFFF;
                int c = cc | (cb << 16);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                int cnt = c & 0xFFFF;
                if (!screen_ok || cnt != (c >> 16)) {        // a borderline evaluation somewhere: recount exactly, as the reference
                    cnt = 0;
                    const double* Texact = Th + min(hb + a, nh - 1) * 12;
                    for (int i = lane; i < M; i += 32) cnt += exact_inlier(Texact, gP, gQ, tq[i], prm.thr_sq_star);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                }
                if (lane == 0 && hb + a < nh) counts[hb + a] = cnt;
            }
        }
        __syncthreads();
        if (tid == 0) {
            int maxc = s_maxc, best = s_best, brk = 0, run = h0 + nh;
            for (int h = 0; h < nh; ++h) {
                const int c = counts[h];
                if (c > maxc) {                                           // :233 strict '>'
                    maxc = c; best = h0 + h;
#pragma unroll
                    for (int e = 0; e < 12; ++e) Tbest[e] = Th[h * 12 + e];
                    if (maxc >= 3 && (double)maxc > prm.break_pct * (double)M) { brk = 1; run = h0 + h + 1; break; }   // :239
                }
            }
            s_maxc = maxc; s_best = best; s_break = brk; s_run = run;
        }
        __syncthreads();
        if (prm.dbg_counts && tid < nh && h0 + tid < s_run) prm.dbg_counts[(size_t)pair * I + h0 + tid] = counts[tid];
        if (s_break) break;
    }

    const int maxc = s_maxc;
    if (prm.dbg_skip == 3) return;
    if (maxc < 3) {              // :291-294 no hypothesis reached 3 inliers: T = I, consensus 0, still "true"
        if (tid == 0) {
            res->ok = 1; res->cam_from = cam_from; res->cam_to = cam_to;
            res->n_ratio_matches = n_ratio; res->n_matches = M; res->consensus = 0;
            res->best_iteration = s_best; res->iterations_run = s_run; res->mse = 0.0; res->info_scale = 1.0;
            write_identity(res->T);
        }
        UZ_PUBLISH_RECORD();
        if (prm.dbg_mask) for (int i = tid; i < M; i += THREADS) prm.dbg_mask[(size_t)pair * cap + i] = 0;
        return;
    }

    // ---------------- K5: refit on the winner's inliers, recount, mse -------------------------------
    UZ_PHASE(5);
    // (a) the winner's consensus set, compacted in index order (the refit recurrence is order dependent)
    int n_in = 0;
    for (int base = 0; base < M; base += THREADS) {
        const int i = base + tid;
        bool in = false;
        if (i < M) {
            const uint32_t e = tq[i];
            const double* p = gP + 3 * (e & 0xFFFFu);
            const double* q = gQ + 3 * (e >> 16);
            in = residual_sq(Tbest, p[0], p[1], p[2], q[0], q[1], q[2]) < prm.thr_sq_star;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = n_in, tot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) { if (w < warp) off += s_wcnt[w]; tot += s_wcnt[w]; }
        if (in) ilist[off + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)i;
        n_in += tot;
        __syncthreads();
    }
    // (b) pcl::TransformationFromCorrespondences::add over the inliers in order (:247-257).  The float32
    // recurrence is sequential in the points but its 15 state scalars are independent of each other:
    // lane l < 9 carries covariance element (r,c) = (l/3, l%3) together with the two means it needs, every
    // lane executing exactly the scalar operation sequence of uz::pose_add for its element.  The operands
    // (float-rounded points, alpha = 1/n, 1-alpha) are staged by all threads, chunk by chunk, into the
    // hypothesis buffer (dead by now) so that the sequential loop is loads at known addresses + 10 flops.
    {
        float* fb = reinterpret_cast<float*>(Th);
        constexpr int CH = THREADS * 12 * 8 / (8 * 4);       // floats per staged array
        const int e9 = lane % 9, r = e9 / 3, c = e9 % 3;
        float m1 = 0.f, m2 = 0.f, cv = 0.f;
        for (int c0 = 0; c0 < n_in; c0 += CH) {
            const int nch = min(CH, n_in - c0);
            for (int k = tid; k < nch; k += THREADS) {
                const int i = pidx(ilist[c0 + k]);
                fb[0 * CH + k] = pxf[i]; fb[1 * CH + k] = pyf[i]; fb[2 * CH + k] = pzf[i];
                fb[3 * CH + k] = qxf[i]; fb[4 * CH + k] = qyf[i]; fb[5 * CH + k] = qzf[i];
                const float alpha = UZ_FDIV(1.0f, (float)(c0 + k + 1));     // accumulated weight == n exactly
                fb[6 * CH + k] = alpha;
                fb[7 * CH + k] = UZ_FSUB(1.0f, alpha);
            }
            __syncthreads();
            if (warp == 0) {
                const float* __restrict__ Pc = fb + c * CH;
                const float* __restrict__ Qr = fb + (3 + r) * CH;
#pragma unroll 4
                for (int k = 0; k < nch; ++k) {
                    const float alpha = fb[6 * CH + k], oma = fb[7 * CH + k];
                    const float d1 = UZ_FSUB(Pc[k], m1), d2 = UZ_FSUB(Qr[k], m2);
                    cv = UZ_FMUL(oma, UZ_FADD(cv, UZ_FMUL(alpha, UZ_FMUL(d2, d1))));
                    m1 = UZ_FADD(m1, UZ_FMUL(alpha, d1));
                    m2 = UZ_FADD(m2, UZ_FMUL(alpha, d2));
                }
            }
            __syncthreads();
        }
        if (warp == 0) {
            PoseAcc A;
            A.acc = (float)n_in;
#pragma unroll
            for (int k = 0; k < 9; ++k) A.c[k] = __shfl_sync(0xffffffffu, cv, k);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                A.m1[k] = __shfl_sync(0xffffffffu, m1, k);          // lane k: (r,c) = (0,k)
                A.m2[k] = __shfl_sync(0xffffffffu, m2, 3 * k);      // lane 3k: (r,c) = (k,0)
            }
            if (lane == 0) pose_finish(A, Tfin);
        }
    }
    __syncthreads();
    UZ_PHASE(6);
    int consensus = 0;
    for (int base = 0; base < M; base += THREADS) {
        const int i = base + tid;
        bool in = false;
        double nrm = 0.0;
        if (i < M) {
            const uint32_t e = tq[i];
            const double* p = gP + 3 * (e & 0xFFFFu);
            const double* q = gQ + 3 * (e >> 16);
            const double s = residual_sq(Tfin, p[0], p[1], p[2], q[0], q[1], q[2]);
            in = s < prm.thr_sq_star;
            nrm = in ? UZ_DSQRT(s) : 0.0;
            if (prm.dbg_mask) prm.dbg_mask[(size_t)pair * cap + i] = in;
        }
        consensus += __syncthreads_count(in);
        if (i < M) norms[i] = nrm;              // pf is dead by now (the refit staged its operands already)
    }
    __syncthreads();
    if (tid == 0) {
        double mse = 0.0;
#pragma unroll 4
        for (int i = 0; i < M; ++i) mse = UZ_DADD(mse, norms[i]);          // :285-289 in index order (+0.0 is exact)
        mse = UZ_DDIV(mse, (double)consensus);                            // :290 (NaN when consensus == 0)
        double info = 1.0;
        if (consensus > 0 && mse > 0) info = UZ_DDIV(UZ_DMUL(0.1, (double)consensus), mse);   // :134-135
        res->ok = 1; res->cam_from = cam_from; res->cam_to = cam_to; res->n_ratio_matches = n_ratio;
        res->n_matches = M; res->consensus = consensus; res->best_iteration = s_best;
        res->iterations_run = s_run; res->mse = mse; res->info_scale = info;
#pragma unroll
        for (int e = 0; e < 12; ++e) res->T[e] = Tfin[e];
        res->T[12] = 0.0; res->T[13] = 0.0; res->T[14] = 0.0; res->T[15] = 1.0;
    }
    UZ_PUBLISH_RECORD();
    UZ_PHASE(7);
#undef UZ_PHASE
#undef UZ_PUBLISH_RECORD
}

// One CTA per pair: the launch form of small batches, of the direct entry points (uz_estimate_svd, cluster RANSAC) and of
// everything that runs under a tool that serialises kernels.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == kSolveThreads ? UZ_SOLVE_MINB : 1) solve_kernel(const MatchTask* __restrict__ tasks,
                                                        const int2* __restrict__ pair_tasks,
                                                        const uint2* keys, SolveParams prm,
                                                        uz_edge_result* __restrict__ results) {
    solve_pair<THREADS, UZ_SOLVE_H>(tasks, pair_tasks, keys, prm, results, prm.pair_base + (int)blockIdx.x);
}

// Streaming form: a small persistent grid (a CTA or two per SM, 96 registers so that it fits beside five match CTAs)
// that runs BESIDE the match kernel of the same batch on a high-priority stream.  CTAs draw pairs in batch order from
// a ticket counter and wait until the match kernel has published every tile of the pair (pair_pending[pair] == 0,
// written by knn2_kernel with fence + atomic, read here with ld.acquire).  The latency-bound phases of the solve
// (sort, sequential refit, scans) then fill issue slots the POPC-bound match warps leave idle instead of costing
// their own 3 ms behind the match kernel.
//
// Forward progress does not depend on the two kernels being co-resident: a CTA whose pair makes it wait while the match
// kernel shows no progress at all for kStallNs (ctl->progress counts finished match tiles) puts the pair on the
// deferred list, raises ctl->gave_up and exits, and so does every other CTA at its next look; that frees the SMs.
// A second launch of the same kernel in CLEANUP form, stream-ordered behind the match kernel, then solves the deferred
// pairs and whatever the ticket counter had not handed out (normally nothing: it exits at once).  Under a tool that
// serialises kernels (ncu, compute-sanitizer) the match kernel simply runs first and nobody waits.
struct StreamCtl {
    unsigned int ticket;       // next pair in batch order
    unsigned int progress;     // match tiles finished so far (knn2_kernel)
    unsigned int gave_up;      // a streaming CTA starved: everyone defers and leaves
    unsigned int n_deferred;   // entries of the deferred list
    unsigned int ticket2;      // next deferred entry (cleanup form)
};
constexpr unsigned long long kStallNs = 20ull * 1000ull * 1000ull;

__device__ __forceinline__ int ld_acquire_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Test hook (UZ_STREAM_PROBE=9): holds the match kernel back for a while so that the streaming grid really starves and
// the give-up + cleanup path above is exercised (tests/test_streaming.py).
__global__ void delay_kernel(unsigned long long ns) {
    const unsigned long long t0 = global_timer_ns();
    while (global_timer_ns() - t0 < ns) __nanosleep(1000);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 5) solve_stream_kernel(const MatchTask* __restrict__ tasks,
                                                               const int2* __restrict__ pair_tasks,
                                                               const uint2* keys, SolveParams prm,
                                                               uz_edge_result* __restrict__ results, int n_pairs,
                                                               const int* pair_pending, StreamCtl* ctl,
                                                               int* deferred, int cleanup) {
    __shared__ int s_next;
    for (;;) {
        if (threadIdx.x == 0) {
            int p = -1;
            if (cleanup) {
                const unsigned int k = atomicAdd(&ctl->ticket2, 1u);
                if (k < ld_relaxed_u32(&ctl->n_deferred)) p = deferred[k];
            }
            if (p < 0 && (cleanup || ld_relaxed_u32(&ctl->gave_up) == 0)) {
                const unsigned int t = atomicAdd(&ctl->ticket, 1u);
                if (t < (unsigned int)n_pairs) p = (int)t;
            }
            if (p >= 0 && !cleanup && ld_acquire_s32(pair_pending + p) > 0) {
                unsigned int seen = ld_relaxed_u32(&ctl->progress), spins = 0;
                unsigned long long t_seen = global_timer_ns();
                while (ld_acquire_s32(pair_pending + p) > 0) {
                    __nanosleep(spins < 32 ? 200 : 1000);
                    if ((++spins & 31u) != 0) continue;
                    const unsigned int now = ld_relaxed_u32(&ctl->progress);
                    const unsigned long long t_now = global_timer_ns();
                    if (now != seen) { seen = now; t_seen = t_now; continue; }
                    if (t_now - t_seen > kStallNs || ld_relaxed_u32(&ctl->gave_up) != 0) {
                        atomicExch(&ctl->gave_up, 1u);
                        deferred[atomicAdd(&ctl->n_deferred, 1u)] = p;
                        p = -1;
                        break;
                    }
                }
            }
            s_next = p;
        }
        __syncthreads();
        const int pair = s_next;
        if (pair < 0) return;
        if (prm.dbg_skip != 1) solve_pair<THREADS, UZ_STREAM_H>(tasks, pair_tasks, keys, prm, results, prm.pair_base + pair);
        __syncthreads();             // the pair's shared-memory state (and s_next) is reused by the next one
    }
}

}  // namespace uz
