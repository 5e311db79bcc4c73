// small_math.cuh -- fixed-size linear algebra run by one device thread per scan (K5).
//
// float routines follow the reference's Kabsch step op by op
// (localization/src/icp_point_to_point.cpp:112-159; Eigen::JacobiSVD<Matrix3f> restated from
// its published two-sided Jacobi algorithm).  The library is compiled with -fmad=false, and
// sqrtf / division are IEEE-rounded (nvcc defaults), so these produce the same bits as the
// same sequence of float operations on an SSE2 host.
// double routines serve the Gauss-Newton and Open3D-flow modes.
// All matrices are column-major: M(r,c) = m[c*rows + r].
#pragma once
#include <cfloat>

#include "common.cuh"

namespace ssf {

#define SM3(m, r, c) (m)[(c)*3 + (r)]
#define SM4(m, r, c) (m)[(c)*4 + (r)]

__device__ inline void mat4_mul_f(const float *A, const float *B, float *C)
{
    float t[16];
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            float s = SM4(A, r, 0) * SM4(B, 0, c);
            s += SM4(A, r, 1) * SM4(B, 1, c);
            s += SM4(A, r, 2) * SM4(B, 2, c);
            s += SM4(A, r, 3) * SM4(B, 3, c);
            t[c * 4 + r] = s;
        }
    for (int i = 0; i < 16; ++i) C[i] = t[i];
}

__device__ inline void rot_left3(float *W, int p, int q, float c, float s)
{
    for (int k = 0; k < 3; ++k) {
        float x = SM3(W, p, k), y = SM3(W, q, k);
        SM3(W, p, k) = c * x + s * y;
        SM3(W, q, k) = -s * x + c * y;
    }
}
__device__ inline void rot_right3(float *W, int p, int q, float c, float s)
{
    for (int k = 0; k < 3; ++k) {
        float x = SM3(W, k, p), y = SM3(W, k, q);
        SM3(W, k, p) = c * x - s * y;
        SM3(W, k, q) = s * x + c * y;
    }
}

// H = U diag(S) V^T, S descending (two-sided Jacobi, real 2x2 kernel)
__device__ inline void jacobi_svd3_f(const float *H, float *U, float *S, float *V)
{
    const float precision = 2.0f * FLT_EPSILON, tiny = FLT_MIN;
    float W[9];
    float scale = 0.f;
    for (int i = 0; i < 9; ++i) {
        float a = fabsf(H[i]);
        if (a > scale) scale = a;
    }
    if (!(scale > 0.f) || !isfinite(scale)) scale = 1.f;
    for (int i = 0; i < 9; ++i) {
        W[i] = H[i] / scale;
        U[i] = V[i] = (i % 4 == 0) ? 1.f : 0.f;
    }
    float maxdiag = fmaxf(fabsf(SM3(W, 0, 0)), fmaxf(fabsf(SM3(W, 1, 1)), fabsf(SM3(W, 2, 2))));
    bool finished = false;
    int sweeps = 0;
    while (!finished && sweeps++ < 64) {
        finished = true;
        for (int p = 1; p < 3; ++p)
            for (int q = 0; q < p; ++q) {
                const float thr = fmaxf(tiny, precision * maxdiag);
                if (!(fabsf(SM3(W, p, q)) > thr || fabsf(SM3(W, q, p)) > thr)) continue;
                finished = false;
                const float m00 = SM3(W, p, p), m01 = SM3(W, p, q), m10 = SM3(W, q, p), m11 = SM3(W, q, q);
                const float t = m00 + m11, d = m10 - m01;
                float c1, s1;
                if (fabsf(d) < tiny) {
                    s1 = 0.f;
                    c1 = 1.f;
                } else {
                    const float u = t / d, tmp = sqrtf(1.f + u * u);
                    s1 = 1.f / tmp;
                    c1 = u / tmp;
                }
                const float a00 = c1 * m00 + s1 * m10, a01 = c1 * m01 + s1 * m11, a11 = -s1 * m01 + c1 * m11;
                float cr, sr;
                const float deno = 2.f * fabsf(a01);
                if (deno < tiny) {
                    cr = 1.f;
                    sr = 0.f;
                } else {
                    const float tau = (a00 - a11) / deno, w = sqrtf(tau * tau + 1.f);
                    const float tt = tau > 0.f ? 1.f / (tau + w) : 1.f / (tau - w);
                    const float sign_t = tt > 0.f ? 1.f : -1.f, nn = 1.f / sqrtf(tt * tt + 1.f);
                    sr = -sign_t * (a01 / fabsf(a01)) * fabsf(tt) * nn;
                    cr = nn;
                }
                const float cl = c1 * cr - s1 * (-sr), sl = c1 * (-sr) + s1 * cr;
                rot_left3(W, p, q, cl, sl);
                rot_right3(U, p, q, cl, -sl);
                rot_right3(W, p, q, cr, sr);
                rot_right3(V, p, q, cr, sr);
                maxdiag = fmaxf(maxdiag, fmaxf(fabsf(SM3(W, p, p)), fabsf(SM3(W, q, q))));
            }
    }
    for (int i = 0; i < 3; ++i) {
        const float a = fabsf(SM3(W, i, i));
        S[i] = a;
        if (a != 0.f) {
            const float f = SM3(W, i, i) / a;
            for (int k = 0; k < 3; ++k) SM3(U, k, i) *= f;
        }
    }
    for (int i = 0; i < 3; ++i) S[i] *= scale;
    for (int i = 0; i < 3; ++i) {
        int pos = i;
        for (int k = i + 1; k < 3; ++k)
            if (S[k] > S[pos]) pos = k;
        if (S[pos] == 0.f) break;
        if (pos != i) {
            float ts = S[i]; S[i] = S[pos]; S[pos] = ts;
            for (int k = 0; k < 3; ++k) {
                float tu = SM3(U, k, i); SM3(U, k, i) = SM3(U, k, pos); SM3(U, k, pos) = tu;
                float tv = SM3(V, k, i); SM3(V, k, i) = SM3(V, k, pos); SM3(V, k, pos) = tv;
            }
        }
    }
}

__device__ inline void mat3_mul_abt_f(const float *A, const float *B, float *C)
{
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            const float s12 = SM3(A, r, 1) * SM3(B, c, 1) + SM3(A, r, 2) * SM3(B, c, 2);
            SM3(C, r, c) = SM3(A, r, 0) * SM3(B, c, 0) + s12;
        }
}

__device__ inline float det3_f(const float *R)
{
    return SM3(R, 0, 0) * (SM3(R, 1, 1) * SM3(R, 2, 2) - SM3(R, 1, 2) * SM3(R, 2, 1)) -
           SM3(R, 0, 1) * (SM3(R, 1, 0) * SM3(R, 2, 2) - SM3(R, 1, 2) * SM3(R, 2, 0)) +
           SM3(R, 0, 2) * (SM3(R, 1, 0) * SM3(R, 2, 1) - SM3(R, 1, 1) * SM3(R, 2, 0));
}

// Kabsch step from centroids and the centred cross-covariance (cpp:137-158)
__device__ inline void kabsch_from_moments_f(const float *cs, const float *ct, const float *H, float *T_step)
{
    float U[9], S[3], V[9], R[9];
    jacobi_svd3_f(H, U, S, V);
    mat3_mul_abt_f(V, U, R);
    if (det3_f(R) < 0.f) {
        for (int k = 0; k < 3; ++k) SM3(V, k, 2) *= -1.f;
        mat3_mul_abt_f(V, U, R);
    }
    float tr[3];
    for (int r = 0; r < 3; ++r) {
        const float a12 = SM3(R, r, 1) * cs[1] + SM3(R, r, 2) * cs[2];
        const float acc = SM3(R, r, 0) * cs[0] + a12;
        tr[r] = ct[r] - acc;
    }
    for (int i = 0; i < 16; ++i) T_step[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) SM4(T_step, r, c) = SM3(R, r, c);
    for (int r = 0; r < 3; ++r) SM4(T_step, r, 3) = tr[r];
}

// ---- double -----------------------------------------------------------------------------------
__device__ inline bool cholesky_solve6(const double *A_in, const double *b, double *x)
{
    double L[36];
    for (int i = 0; i < 36; ++i) L[i] = A_in[i];
    for (int j = 0; j < 6; ++j) {
        double d = L[j * 6 + j];
        for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k];
        if (!(d > 0.0)) return false;
        d = sqrt(d);
        L[j * 6 + j] = d;
        for (int i = j + 1; i < 6; ++i) {
            double s = L[i * 6 + j];
            for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
            L[i * 6 + j] = s / d;
        }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) {
        double s = b[i];
        for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * y[k];
        y[i] = s / L[i * 6 + i];
    }
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * x[k];
        x[i] = s / L[i * 6 + i];
    }
    return true;
}

// Ts (column-major 4x4) = [exp([w]x) | t], x = (w, t)
__device__ inline void se3_from_twist(const double *x, double *Ts)
{
    const double wx = x[0], wy = x[1], wz = x[2];
    const double th2 = wx * wx + wy * wy + wz * wz, th = sqrt(th2);
    double a, b;
    if (th < 1e-8) {
        a = 1.0 - th2 / 6.0;
        b = 0.5 - th2 / 24.0;
    } else {
        a = sin(th) / th;
        b = (1.0 - cos(th)) / th2;
    }
    const double K[9] = {0, wz, -wy, -wz, 0, wx, wy, -wx, 0};
    double K2[9];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += K[k * 3 + r] * K[c * 3 + k];
            K2[c * 3 + r] = s;
        }
    for (int i = 0; i < 16; ++i) Ts[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) Ts[c * 4 + r] = (r == c ? 1.0 : 0.0) + a * K[c * 3 + r] + b * K2[c * 3 + r];
    Ts[12] = x[3];
    Ts[13] = x[4];
    Ts[14] = x[5];
}

// T <- fl(Ts * T)
__device__ inline void compose_round(const double *Ts, float *T)
{
    double r[16];
    for (int c = 0; c < 4; ++c)
        for (int rr = 0; rr < 4; ++rr) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += Ts[k * 4 + rr] * (double)T[c * 4 + k];
            r[c * 4 + rr] = s;
        }
    for (int i = 0; i < 16; ++i) T[i] = (float)r[i];
}

// one-sided (Hestenes) Jacobi: H V = U diag(S), S descending; rank-deficient U completed
__device__ inline void jacobi_svd3_d(const double *H, double *U, double *S, double *V)
{
    double A[9];
    for (int i = 0; i < 9; ++i) {
        A[i] = H[i];
        V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int k = 0; k < 3; ++k) {
                    alpha += A[p * 3 + k] * A[p * 3 + k];
                    beta += A[q * 3 + k] * A[q * 3 + k];
                    gamma += A[p * 3 + k] * A[q * 3 + k];
                }
                if (gamma == 0.0) continue;
                const double lim = fabs(gamma) / sqrt(alpha * beta + 1e-300);
                if (lim > off) off = lim;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < 3; ++k) {
                    const double ap = A[p * 3 + k], aq = A[q * 3 + k];
                    A[p * 3 + k] = c * ap - s * aq;
                    A[q * 3 + k] = s * ap + c * aq;
                    const double vp = V[p * 3 + k], vq = V[q * 3 + k];
                    V[p * 3 + k] = c * vp - s * vq;
                    V[q * 3 + k] = s * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    for (int j = 0; j < 3; ++j)
        S[j] = sqrt(A[j * 3] * A[j * 3] + A[j * 3 + 1] * A[j * 3 + 1] + A[j * 3 + 2] * A[j * 3 + 2]);
    for (int i = 0; i < 3; ++i) {
        int pos = i;
        for (int k = i + 1; k < 3; ++k)
            if (S[k] > S[pos]) pos = k;
        if (pos != i) {
            double ts = S[i]; S[i] = S[pos]; S[pos] = ts;
            for (int k = 0; k < 3; ++k) {
                double ta = A[i * 3 + k]; A[i * 3 + k] = A[pos * 3 + k]; A[pos * 3 + k] = ta;
                double tv = V[i * 3 + k]; V[i * 3 + k] = V[pos * 3 + k]; V[pos * 3 + k] = tv;
            }
        }
    }
    for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) U[j * 3 + k] = S[j] > 1e-300 ? A[j * 3 + k] / S[j] : 0.0;
    if (!(S[1] > S[0] * 1e-14)) {
        double *u0 = U, *u1 = U + 3;
        const int k = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2) : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
        double e[3] = {0, 0, 0};
        e[k] = 1.0;
        const double d = e[0] * u0[0] + e[1] * u0[1] + e[2] * u0[2];
        for (int i = 0; i < 3; ++i) u1[i] = e[i] - d * u0[i];
        const double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
        for (int i = 0; i < 3; ++i) u1[i] /= n1;
    }
    if (!(S[2] > S[0] * 1e-14)) {
        double *u0 = U, *u1 = U + 3, *u2 = U + 6;
        u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
        u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
        u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
    }
}

__device__ inline double det3_d(const double *R)
{
    return R[0] * (R[4] * R[8] - R[7] * R[5]) - R[3] * (R[1] * R[8] - R[7] * R[2]) + R[6] * (R[1] * R[5] - R[4] * R[2]);
}

// Kabsch (Umeyama without scale) in double: sp, sq centroids, H = sum (p - sp)(q - sq)^T
__device__ inline void kabsch_from_moments_d(const double *sp, const double *sq, const double *H, double *Ts)
{
    double U[9], S[3], V[9], R[9];
    jacobi_svd3_d(H, U, S, V);
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += V[k * 3 + r] * U[k * 3 + c];
            R[c * 3 + r] = s;
        }
    if (det3_d(R) < 0) {
        for (int k = 0; k < 3; ++k) V[6 + k] = -V[6 + k];
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += V[k * 3 + r] * U[k * 3 + c];
                R[c * 3 + r] = s;
            }
    }
    for (int i = 0; i < 16; ++i) Ts[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) Ts[c * 4 + r] = R[c * 3 + r];
    for (int r = 0; r < 3; ++r) Ts[12 + r] = sq[r] - (R[0 * 3 + r] * sp[0] + R[1 * 3 + r] * sp[1] + R[2 * 3 + r] * sp[2]);
}

}  // namespace ssf
