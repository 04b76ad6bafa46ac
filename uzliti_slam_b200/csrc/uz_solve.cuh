// uz_solve.cuh — K2..K5: everything after matching, one CTA per keyframe pair, the pair kept on-chip.
//
// Follows /root/reference/transformation_estimation/src/feature_transformation_estimator.cpp:
//   K2  :65-71 ratio test, :74-86 best camera pair (score = #ratio survivors, first strict max),
//       :103-112 valid_3d filter, :114 sort (total order (distance, queryIdx)),
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
    // parity taps (may be null)
    int32_t* dbg_matches;    // [pair][cap][3]
    uint8_t* dbg_mask;       // [pair][cap]
    int32_t* dbg_counts;     // [pair][iterations] consensus count of every evaluated hypothesis, -1 = not run
    long long* dbg_phase;    // [pair][8] clock64() at the phase boundaries of the CTA (profiling tap)
};

constexpr int kSolveThreads = 128;

__host__ __device__ constexpr size_t solve_smem_bytes(int cap) {
    return (size_t)cap * 48 /*P,Q (px is reused for the residual norms)*/ + (size_t)cap * 4 /*sort keys*/ +
           (size_t)kSolveThreads * 12 * 8 /*hypothesis transforms*/ + (size_t)kSolveThreads * 4 /*counts*/ +
           (size_t)cap /*mask*/ + 256 /*scalars*/;
}

static_assert(solve_smem_bytes(UZ_MAX_FEATURES) + 64 <= 232448, "solve kernel exceeds the 227 KB per-CTA shared memory of sm_100");

__device__ __forceinline__ void write_identity(double* T16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) T16[i] = (i % 5 == 0) ? 1.0 : 0.0;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) solve_kernel(const MatchTask* __restrict__ tasks,
                                                        const int2* __restrict__ pair_tasks,
                                                        const uint2* __restrict__ keys, SolveParams prm,
                                                        uz_edge_result* __restrict__ results) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int cap = prm.cap;
    double* px = reinterpret_cast<double*>(smem_raw);
    double* py = px + cap; double* pz = py + cap;
    double* qx = pz + cap; double* qy = qx + cap; double* qz = qy + cap;
    double* Th = qz + cap;                                            // [THREADS][12]
    double* norms = px;                                               // norms[i] overwrites px[i] in place (K5)
    double* Tbest = Th + THREADS * 12;                                // 12
    double* Tfin = Tbest + 12;                                        // 12
    int32_t* counts = reinterpret_cast<int32_t*>(Tfin + 12);          // [THREADS]
    uint32_t* skeys = reinterpret_cast<uint32_t*>(counts + THREADS);  // [cap]
    uint8_t* mask = reinterpret_cast<uint8_t*>(skeys + cap);          // [cap]
    __shared__ int s_best, s_maxc, s_break, s_run;

    const int tid = threadIdx.x;
    const int pair = blockIdx.x;
    uz_edge_result* res = results + pair;
#define UZ_PHASE(k) do { if (prm.dbg_phase && tid == 0) prm.dbg_phase[(size_t)pair * 8 + (k)] = clock64(); } while (0)
    UZ_PHASE(0);
    int M = 0;
    int n_ratio = 0, cam_from = -1, cam_to = -1;

    if (prm.direct_P == nullptr) {
        // ---------------- K2: best camera pair, filter, sort, gather -------------------------------
        const int2 pt = pair_tasks[pair];
        int best = -1, best_score = -1;
        for (int t = 0; t < pt.y; ++t) {
            const MatchTask* tk = tasks + pt.x + t;
            const uint2* k = keys + tk->key_off;
            int cnt = 0;
            for (int base = 0; base < tk->nq; base += THREADS) {
                const int q = base + tid;
                bool pass = false;
                if (q < tk->nq) {
                    const uint2 m = k[q];
                    pass = (m.y != kNoKey) && ((int)(m.x >> 16) * prm.ratio_den < (int)(m.y >> 16) * prm.ratio_num);
                }
                cnt += __syncthreads_count(pass);
            }
            if (cnt > best_score) { best_score = cnt; best = t; }   // :81 strict '>' keeps the first
        }
        if (best < 0) {          // :93-95 no comparable camera pair
            if (tid == 0) {
                res->ok = 0; res->cam_from = -1; res->cam_to = -1; res->n_ratio_matches = 0; res->n_matches = 0;
                res->consensus = 0; res->best_iteration = -1; res->iterations_run = 0; res->mse = 0.0;
                res->info_scale = 1.0; write_identity(res->T);
            }
            return;
        }
        const MatchTask* tk = tasks + pt.x + best;
        const uint2* k = keys + tk->key_off;
        const int nq = tk->nq;
        n_ratio = best_score; cam_from = tk->cam_from; cam_to = tk->cam_to;
        const uint8_t* __restrict__ vq = tk->q_valid;
        const uint8_t* __restrict__ vt = tk->t_valid;
        for (int i = tid; i < cap; i += THREADS) {
            uint32_t key = kNoKey;
            if (i < nq) {
                const uint2 m = k[i];
                const bool pass = (m.y != kNoKey) && ((int)(m.x >> 16) * prm.ratio_den < (int)(m.y >> 16) * prm.ratio_num);
                if (pass && vq[i] && vt[m.x & 0xFFFFu]) key = (m.x & 0xFFFF0000u) | (uint32_t)i;
            }
            skeys[i] = key;
        }
        __syncthreads();
        UZ_PHASE(1);
        // bitonic sort ascending: (distance, queryIdx); kNoKey sinks to the end
        for (int kk = 2; kk <= cap; kk <<= 1) {
            for (int j = kk >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < cap; i += THREADS) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const uint32_t a = skeys[i], b = skeys[ixj];
                        const bool asc = (i & kk) == 0;
                        if ((a > b) == asc) { skeys[i] = b; skeys[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        }
        UZ_PHASE(2);
        for (int base = 0; base < cap; base += THREADS)
            M += __syncthreads_count(skeys[base + tid] != kNoKey);
        const double* __restrict__ pq = tk->q_pos;
        const double* __restrict__ pt3 = tk->t_pos;
        for (int i = tid; i < M; i += THREADS) {
            const uint32_t key = skeys[i];
            const int q = key & 0xFFFFu;
            const int t = k[q].x & 0xFFFFu;
            px[i] = pq[3 * q]; py[i] = pq[3 * q + 1]; pz[i] = pq[3 * q + 2];
            qx[i] = pt3[3 * t]; qy[i] = pt3[3 * t + 1]; qz[i] = pt3[3 * t + 2];
            if (prm.dbg_matches) {
                int32_t* d = prm.dbg_matches + ((size_t)pair * cap + i) * 3;
                d[0] = q; d[1] = t; d[2] = (int)(key >> 16);
            }
        }
    } else {
        M = prm.direct_M;
        for (int i = tid; i < M; i += THREADS) {
            px[i] = prm.direct_P[3 * i]; py[i] = prm.direct_P[3 * i + 1]; pz[i] = prm.direct_P[3 * i + 2];
            qx[i] = prm.direct_Q[3 * i]; qy[i] = prm.direct_Q[3 * i + 1]; qz[i] = prm.direct_Q[3 * i + 2];
        }
    }
    if (tid == 0) { s_best = -1; s_maxc = 0; s_break = 0; s_run = 0; }
    __syncthreads();
    UZ_PHASE(3);

    if (M < 3) {                 // :118/:158 not enough depth-valid matches
        if (tid == 0) {
            res->ok = 0; res->cam_from = cam_from; res->cam_to = cam_to; res->n_ratio_matches = n_ratio;
            res->n_matches = M; res->consensus = 0; res->best_iteration = -1; res->iterations_run = 0;
            res->mse = 0.0; res->info_scale = 1.0; write_identity(res->T);
        }
        if (prm.dbg_mask) for (int i = tid; i < M; i += THREADS) prm.dbg_mask[(size_t)pair * cap + i] = 0;
        return;
    }

    // ---------------- K3/K4: hypotheses, consensus counts, sequential-semantics winner -------------
    const int I = prm.iterations;
    const uint16_t* __restrict__ samp = prm.samples + (prm.samples_by_m ? (size_t)M * I * 3 : 0);
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;
    for (int h0 = 0; h0 < I; h0 += THREADS) {
        const int nh = min(THREADS, I - h0);
        if (tid < nh) {
            PoseAcc acc;
            pose_reset(acc);
#pragma unroll 1
            for (int j = 0; j < 3; ++j) {
                const int s = samp[(size_t)(h0 + tid) * 3 + j];
                pose_add(acc, (float)px[s], (float)py[s], (float)pz[s], (float)qx[s], (float)qy[s], (float)qz[s]);
            }
            pose_finish(acc, Th + tid * 12);
        }
        __syncthreads();
        if (h0 == 0) UZ_PHASE(4);
        // each warp scores two hypotheses per pass over the points
        for (int h = warp * 2; h < nh; h += NW * 2) {
            const bool two = (h + 1) < nh;
            double Ta[12], Tb[12];
#pragma unroll
            for (int e = 0; e < 12; ++e) { Ta[e] = Th[h * 12 + e]; Tb[e] = Th[(two ? h + 1 : h) * 12 + e]; }
            int ca = 0, cb = 0;
            for (int i = lane; i < M; i += 32) {
                const double x = px[i], y = py[i], z = pz[i], u = qx[i], v = qy[i], w = qz[i];
                ca += residual_sq(Ta, x, y, z, u, v, w) < prm.thr_sq_star;
                cb += residual_sq(Tb, x, y, z, u, v, w) < prm.thr_sq_star;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ca += __shfl_xor_sync(0xffffffffu, ca, o);
                cb += __shfl_xor_sync(0xffffffffu, cb, o);
            }
            if (lane == 0) { counts[h] = ca; if (two) counts[h + 1] = cb; }
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
    if (maxc < 3) {              // :291-294 no hypothesis reached 3 inliers: T = I, consensus 0, still "true"
        if (tid == 0) {
            res->ok = 1; res->cam_from = cam_from; res->cam_to = cam_to;
            res->n_ratio_matches = n_ratio; res->n_matches = M; res->consensus = 0;
            res->best_iteration = s_best; res->iterations_run = s_run; res->mse = 0.0; res->info_scale = 1.0;
            write_identity(res->T);
        }
        if (prm.dbg_mask) for (int i = tid; i < M; i += THREADS) prm.dbg_mask[(size_t)pair * cap + i] = 0;
        return;
    }

    // ---------------- K5: refit on the winner's inliers, recount, mse -------------------------------
    UZ_PHASE(5);
    // (a) the winner's consensus set, compacted in index order (the refit recurrence is order dependent)
    uint16_t* ilist = reinterpret_cast<uint16_t*>(skeys);
    __shared__ int s_wcnt[NW];
    int n_in = 0;
    for (int base = 0; base < M; base += THREADS) {
        const int i = base + tid;
        const bool in = (i < M) && (residual_sq(Tbest, px[i], py[i], pz[i], qx[i], qy[i], qz[i]) < prm.thr_sq_star);
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
    // lane executing exactly the scalar operation sequence of uz::pose_add for its element.
    if (warp == 0) {
        const int e = lane % 9, r = e / 3, c = e % 3;
        const double* __restrict__ Pc = px + (size_t)c * cap;
        const double* __restrict__ Qr = qx + (size_t)r * cap;
        float acc = 0.f, m1 = 0.f, m2 = 0.f, cv = 0.f;
#pragma unroll 2
        for (int k = 0; k < n_in; ++k) {
            const int i = ilist[k];
            const float p = (float)Pc[i], q = (float)Qr[i];
            acc = UZ_FADD(acc, 1.0f);
            const float alpha = UZ_FDIV(1.0f, acc);
            const float oma = UZ_FSUB(1.0f, alpha);
            const float d1 = UZ_FSUB(p, m1), d2 = UZ_FSUB(q, m2);
            cv = UZ_FMUL(oma, UZ_FADD(cv, UZ_FMUL(alpha, UZ_FMUL(d2, d1))));
            m1 = UZ_FADD(m1, UZ_FMUL(alpha, d1));
            m2 = UZ_FADD(m2, UZ_FMUL(alpha, d2));
        }
        PoseAcc A;
        A.acc = acc;
#pragma unroll
        for (int k = 0; k < 9; ++k) A.c[k] = __shfl_sync(0xffffffffu, cv, k);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            A.m1[k] = __shfl_sync(0xffffffffu, m1, k);          // lane k: (r,c) = (0,k)
            A.m2[k] = __shfl_sync(0xffffffffu, m2, 3 * k);      // lane 3k: (r,c) = (k,0)
        }
        if (lane == 0) pose_finish(A, Tfin);
    }
    __syncthreads();
    UZ_PHASE(6);
    int consensus = 0;
    for (int base = 0; base < M; base += THREADS) {
        const int i = base + tid;
        bool in = false;
        if (i < M) {
            const double s = residual_sq(Tfin, px[i], py[i], pz[i], qx[i], qy[i], qz[i]);
            in = s < prm.thr_sq_star;
            norms[i] = in ? UZ_DSQRT(s) : 0.0;      // overwrites px[i]: P is dead after this pass
            if (prm.dbg_mask) prm.dbg_mask[(size_t)pair * cap + i] = in;
        }
        consensus += __syncthreads_count(in);
    }
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
    UZ_PHASE(7);
#undef UZ_PHASE
}

}  // namespace uz
