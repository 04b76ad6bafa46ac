// uz_arith.cuh — the arithmetic specification of the RANSAC stage, for the device.
//
// "Identical inlier sets" needs every floating-point operation of the pose solve and of the
// consensus test to round exactly as the CPU reference does.  The reference computes
//   * the 3-point / refit pose in FLOAT32: pcl::TransformationFromCorrespondences::add (incremental
//     mean + covariance) and getTransformation (Eigen::JacobiSVD<Matrix3f>, R = U diag(1,1,±1) V^T,
//     t = mean2 - R mean1)          — feature_transformation_estimator.cpp:299-314
//   * the residual ‖T p − q‖ in FLOAT64 with a strict '<'   — feature_transformation_estimator.cpp:337-347
// on x86 without FMA contraction (transformation_estimation/CMakeLists.txt:9: -msse2 -msse3 -mssse3).
// So every operation here is an explicit round-to-nearest intrinsic (never contracted into an FMA by
// nvcc), evaluated in the fixed order written below, with IEEE division and square root.
//
// The functions are UZ_HD so that tests can also compile this header for the HOST (g++,
// -ffp-contract=off) and compare it bit-for-bit with the independently written oracle on CPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define UZ_HD __host__ __device__ __forceinline__
#else
#define UZ_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define UZ_FMUL(a, b) __fmul_rn((a), (b))
#define UZ_FADD(a, b) __fadd_rn((a), (b))
#define UZ_FSUB(a, b) __fsub_rn((a), (b))
#define UZ_FDIV(a, b) __fdiv_rn((a), (b))
#define UZ_FSQRT(a) __fsqrt_rn((a))
#define UZ_DMUL(a, b) __dmul_rn((a), (b))
#define UZ_DADD(a, b) __dadd_rn((a), (b))
#define UZ_DSUB(a, b) __dsub_rn((a), (b))
#define UZ_DDIV(a, b) __ddiv_rn((a), (b))
#define UZ_DSQRT(a) __dsqrt_rn((a))
#else
#include <cmath>
#define UZ_FMUL(a, b) ((float)(a) * (float)(b))
#define UZ_FADD(a, b) ((float)(a) + (float)(b))
#define UZ_FSUB(a, b) ((float)(a) - (float)(b))
#define UZ_FDIV(a, b) ((float)(a) / (float)(b))
#define UZ_FSQRT(a) (sqrtf((a)))
#define UZ_DMUL(a, b) ((double)(a) * (double)(b))
#define UZ_DADD(a, b) ((double)(a) + (double)(b))
#define UZ_DSUB(a, b) ((double)(a) - (double)(b))
#define UZ_DDIV(a, b) ((double)(a) / (double)(b))
#define UZ_DSQRT(a) (sqrt((a)))
#endif

namespace uz {

UZ_HD float f_abs(float x) { return fabsf(x); }                                // |x|, -0 -> +0 (an operand modifier on the device)
UZ_HD float f_max(float a, float b) { return a < b ? b : a; }                  // std::max

// Running state of pcl::TransformationFromCorrespondences (all float32).
struct PoseAcc {
    float acc;          // accumulated_weight_
    float m1[3], m2[3]; // mean1_, mean2_
    float c[9];         // covariance_ row-major
};

UZ_HD void pose_reset(PoseAcc& s) {
    s.acc = 0.f;
    for (int i = 0; i < 3; ++i) s.m1[i] = s.m2[i] = 0.f;
    for (int i = 0; i < 9; ++i) s.c[i] = 0.f;
}

// add(point p, corresponding_point q, weight 1)  — the weight is always 1 (reference :305-309).
UZ_HD void pose_add(PoseAcc& s, float px, float py, float pz, float qx, float qy, float qz) {
    s.acc = UZ_FADD(s.acc, 1.0f);
    const float alpha = UZ_FDIV(1.0f, s.acc);
    const float d1[3] = {UZ_FSUB(px, s.m1[0]), UZ_FSUB(py, s.m1[1]), UZ_FSUB(pz, s.m1[2])};
    const float d2[3] = {UZ_FSUB(qx, s.m2[0]), UZ_FSUB(qy, s.m2[1]), UZ_FSUB(qz, s.m2[2])};
    const float oma = UZ_FSUB(1.0f, alpha);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            s.c[3 * r + c] = UZ_FMUL(oma, UZ_FADD(s.c[3 * r + c], UZ_FMUL(alpha, UZ_FMUL(d2[r], d1[c]))));
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        s.m1[r] = UZ_FADD(s.m1[r], UZ_FMUL(alpha, d1[r]));
        s.m2[r] = UZ_FADD(s.m2[r], UZ_FMUL(alpha, d2[r]));
    }
}

struct Rot { float c, s; };

// Eigen::JacobiRotation<float>::makeJacobi(x, y, z)
UZ_HD Rot make_jacobi(float x, float y, float z) {
    Rot r;
    if (y == 0.f) { r.c = 1.f; r.s = 0.f; return r; }
    const float ay = f_abs(y);
    const float tau = UZ_FDIV(UZ_FSUB(x, z), UZ_FMUL(2.f, ay));
    const float w = UZ_FSQRT(UZ_FADD(UZ_FMUL(tau, tau), 1.f));
    const float t = (tau > 0.f) ? UZ_FDIV(1.f, UZ_FADD(tau, w)) : UZ_FDIV(1.f, UZ_FSUB(tau, w));
    const float sign_t = t > 0.f ? 1.f : -1.f;
    const float n = UZ_FDIV(1.f, UZ_FSQRT(UZ_FADD(UZ_FMUL(t, t), 1.f)));
    r.s = UZ_FMUL(UZ_FMUL(UZ_FMUL(-sign_t, UZ_FDIV(y, ay)), f_abs(t)), n);
    r.c = n;
    return r;
}

// x' = c*x + s*y ; y' = -s*x + c*y on rows p,q (stride 3 between row elements is 1) of a row-major 3x3
template <int p, int q>
UZ_HD void rot_rows(float* m, Rot j) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float x = m[3 * p + k], y = m[3 * q + k];
        m[3 * p + k] = UZ_FADD(UZ_FMUL(j.c, x), UZ_FMUL(j.s, y));
        m[3 * q + k] = UZ_FADD(UZ_FMUL(-j.s, x), UZ_FMUL(j.c, y));
    }
}
template <int p, int q>
UZ_HD void rot_cols(float* m, Rot j) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float x = m[3 * k + p], y = m[3 * k + q];
        m[3 * k + p] = UZ_FADD(UZ_FMUL(j.c, x), UZ_FMUL(j.s, y));
        m[3 * k + q] = UZ_FADD(UZ_FMUL(-j.s, x), UZ_FMUL(j.c, y));
    }
}

// One (p,q) step of the JacobiSVD sweep; returns true if a rotation was applied.
template <int p, int q>
UZ_HD bool jacobi_step(float* W, float* U, float* V) {
    const float precision = 2.f * 1.1920928955078125e-07f;      // 2 * FLT_EPSILON
    const float considerAsZero = 2.f * 1.401298464324817e-45f;  // 2 * denorm_min
    const float wpp = W[3 * p + p], wqq = W[3 * q + q], wpq = W[3 * p + q], wqp = W[3 * q + p];
    const float threshold = f_max(considerAsZero, UZ_FMUL(precision, f_max(f_abs(wpp), f_abs(wqq))));
    if (!(f_max(f_abs(wpq), f_abs(wqp)) > threshold)) return false;
    Rot rot1;
    const float t = UZ_FADD(wpp, wqq);
    const float d = UZ_FSUB(wqp, wpq);
    if (t == 0.f) {
        rot1.c = 0.f;
        rot1.s = d > 0.f ? 1.f : -1.f;
    } else {
        const float u = UZ_FDIV(d, t);
        rot1.c = UZ_FDIV(1.f, UZ_FSQRT(UZ_FADD(1.f, UZ_FMUL(u, u))));
        rot1.s = UZ_FMUL(rot1.c, u);
    }
    const float n00 = UZ_FADD(UZ_FMUL(rot1.c, wpp), UZ_FMUL(rot1.s, wqp));
    const float n01 = UZ_FADD(UZ_FMUL(rot1.c, wpq), UZ_FMUL(rot1.s, wqq));
    const float n11 = UZ_FADD(UZ_FMUL(-rot1.s, wpq), UZ_FMUL(rot1.c, wqq));
    const Rot jr = make_jacobi(n00, n01, n11);
    Rot jrt; jrt.c = jr.c; jrt.s = -jr.s;
    Rot jl;
    jl.c = UZ_FSUB(UZ_FMUL(rot1.c, jrt.c), UZ_FMUL(rot1.s, jrt.s));
    jl.s = UZ_FADD(UZ_FMUL(rot1.c, jrt.s), UZ_FMUL(rot1.s, jrt.c));
    rot_rows<p, q>(W, jl);
    rot_cols<p, q>(U, jl);
    rot_cols<p, q>(W, jrt);
    rot_cols<p, q>(V, jrt);
    return true;
}

UZ_HD float det3(const float* m) {
    const float a = UZ_FMUL(m[0], UZ_FSUB(UZ_FMUL(m[4], m[8]), UZ_FMUL(m[5], m[7])));
    const float b = UZ_FMUL(m[1], UZ_FSUB(UZ_FMUL(m[3], m[8]), UZ_FMUL(m[5], m[6])));
    const float c = UZ_FMUL(m[2], UZ_FSUB(UZ_FMUL(m[3], m[7]), UZ_FMUL(m[4], m[6])));
    return UZ_FADD(UZ_FSUB(a, b), c);
}

UZ_HD void swap_cols(float* m, int a, int b) {   // a, b are compile-time at every call site after inlining
#pragma unroll
    for (int r = 0; r < 3; ++r) { const float t = m[3 * r + a]; m[3 * r + a] = m[3 * r + b]; m[3 * r + b] = t; }
}

// getTransformation(): T (row-major 3x4 as 12 doubles: r00 r01 r02 tx / r10 ...) from the accumulator.
// Eigen's sweep loop has no iteration cap; 64 sweeps is never reached on finite input (3x3 Jacobi
// converges quadratically) and only guards the device against a hang.
UZ_HD void pose_finish(const PoseAcc& s, double* T12) {
    float W[9], U[9], V[9], sv[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) { W[i] = s.c[i]; U[i] = V[i] = 0.f; }
    U[0] = U[4] = U[8] = 1.f;
    V[0] = V[4] = V[8] = 1.f;
    for (int sweep = 0; sweep < 64; ++sweep) {
        bool any = false;
        any |= jacobi_step<1, 0>(W, U, V);
        any |= jacobi_step<2, 0>(W, U, V);
        any |= jacobi_step<2, 1>(W, U, V);
        if (!any) break;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float w = W[4 * i];
        const float a = f_abs(w);
        sv[i] = a;
        if (a != 0.f) {
            const float sg = UZ_FDIV(w, a);
#pragma unroll
            for (int r = 0; r < 3; ++r) U[3 * r + i] = UZ_FMUL(U[3 * r + i], sg);
        }
    }
    // sort singular values descending (first maximum wins), permuting U and V columns
    {
        int pos = 0; float best = sv[0];
        if (sv[1] > best) { best = sv[1]; pos = 1; }
        if (sv[2] > best) { best = sv[2]; pos = 2; }
        if (best != 0.f) {
            if (pos == 1) { const float t = sv[0]; sv[0] = sv[1]; sv[1] = t; swap_cols(U, 0, 1); swap_cols(V, 0, 1); }
            else if (pos == 2) { const float t = sv[0]; sv[0] = sv[2]; sv[2] = t; swap_cols(U, 0, 2); swap_cols(V, 0, 2); }
            if (sv[2] > sv[1]) {
                const float t = sv[1]; sv[1] = sv[2]; sv[2] = t; swap_cols(U, 1, 2); swap_cols(V, 1, 2);
            }
            // (i = 1 with best == 0 and i = 2 are no-ops)
        }
    }
    const float sgn = (UZ_FMUL(det3(U), det3(V)) < 0.0f) ? -1.0f : 1.0f;
    float R[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            R[3 * r + c] = UZ_FADD(UZ_FADD(UZ_FMUL(U[3 * r], V[3 * c]), UZ_FMUL(U[3 * r + 1], V[3 * c + 1])),
                                   UZ_FMUL(UZ_FMUL(U[3 * r + 2], sgn), V[3 * c + 2]));
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float rm = UZ_FADD(UZ_FADD(UZ_FMUL(R[3 * r], s.m1[0]), UZ_FMUL(R[3 * r + 1], s.m1[1])),
                                 UZ_FMUL(R[3 * r + 2], s.m1[2]));
        T12[4 * r + 0] = (double)R[3 * r];
        T12[4 * r + 1] = (double)R[3 * r + 1];
        T12[4 * r + 2] = (double)R[3 * r + 2];
        T12[4 * r + 3] = (double)UZ_FSUB(s.m2[r], rm);
    }
}

// squared residual ‖T p − q‖² in double: ((r0 x + r1 y) + r2 z) + t, then (dx²+dy²)+dz².
UZ_HD double residual_sq(const double* T12, double px, double py, double pz, double qx, double qy, double qz) {
    const double x = UZ_DADD(UZ_DADD(UZ_DADD(UZ_DMUL(T12[0], px), UZ_DMUL(T12[1], py)), UZ_DMUL(T12[2], pz)), T12[3]);
    const double y = UZ_DADD(UZ_DADD(UZ_DADD(UZ_DMUL(T12[4], px), UZ_DMUL(T12[5], py)), UZ_DMUL(T12[6], pz)), T12[7]);
    const double z = UZ_DADD(UZ_DADD(UZ_DADD(UZ_DMUL(T12[8], px), UZ_DMUL(T12[9], py)), UZ_DMUL(T12[10], pz)), T12[11]);
    const double dx = UZ_DSUB(x, qx), dy = UZ_DSUB(y, qy), dz = UZ_DSUB(z, qz);
    return UZ_DADD(UZ_DADD(UZ_DMUL(dx, dx), UZ_DMUL(dy, dy)), UZ_DMUL(dz, dz));
}

}  // namespace uz
