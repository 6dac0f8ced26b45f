// 2x2 / 3x3 singular value decompositions for the constitutive models.
//
// Replaces wgebra::svd2 / svd3 (dimforge/wgmath @ 6d17942b, not vendored; call sites
// linear_elasticity.wgsl:15,29; drucker_prager.wgsl:80,139; particle_update.wgsl:103,109).
// Contract: F = U diag(S) V^T with U, V proper rotations; the sign of
// det(F) is carried by the last singular value. Every use on the path is invariant to the
// ordering / sign convention.
#pragma once

#include "common.cuh"

namespace b2 {

// ---- 2x2, closed form. m = [a b; c d] column-major input {a, c, b, d}. -----------------------
__host__ __device__ inline void svd2(const float* F, float* U, float* S, float* V) {
    float a = F[0], c = F[1], b = F[2], d = F[3];
    float e = (a + d) * 0.5f, f = (a - d) * 0.5f, g = (c + b) * 0.5f, h = (c - b) * 0.5f;
    float q = sqrtf(e * e + h * h), r = sqrtf(f * f + g * g);
    // q - 1 = (e^2 + h^2 - 1) / (q + 1), the numerator formed in f64 (same motivation as svd3 below)
    const double ed = ((double)a + (double)d) * 0.5, hd = ((double)c - (double)b) * 0.5;
    float qm1 = (float)(ed * ed + hd * hd - 1.0) / (q + 1.0f);
    S[0] = 1.0f + (qm1 + r);
    S[1] = 1.0f + (qm1 - r); // signed: carries det(F)
    float a1 = atan2f(g, f), a2 = atan2f(h, e);
    float theta = (a2 - a1) * 0.5f, phi = (a2 + a1) * 0.5f;
    float sp, cp, st, ct;
    sincosf(phi, &sp, &cp);
    sincosf(theta, &st, &ct);
    // U = rot(phi), V^T = rot(theta)  =>  V = rot(-theta)
    U[0] = cp;
    U[1] = sp;
    U[2] = -sp;
    U[3] = cp;
    V[0] = ct;
    V[1] = -st;
    V[2] = st;
    V[3] = ct;
}

// One Jacobi rotation on the symmetric matrix (app, aqq, apq, and the two other off-diagonals
// apr, aqr), accumulated into columns p, q of V.
__host__ __device__ inline void jacobi_rot(float& app, float& aqq, float& apq, float& apr, float& aqr, float* vp,
                                           float* vq) {
    float c = 1.0f, s = 0.0f;
    float absq = fabsf(apq);
    if (absq > 1e-30f) {
        // The rotation angle only steers convergence; (c, s) is renormalised below, so the
        // approximate division / square root of the device path cost no accuracy in V.
#if defined(__CUDA_ARCH__)
        float tau = __fdividef(aqq - app, 2.0f * apq);
        float x = fmaf(tau, tau, 1.0f);
        float t = __fdividef(copysignf(1.0f, tau), fabsf(tau) + x * rsqrtf(x));
        if (fabsf(tau) > 1e18f) t = 0.0f; // tau^2 overflowed / __fdividef out of range: no rotation needed
#else
        float tau = (aqq - app) / (2.0f * apq);
        float t = copysignf(1.0f, tau) / (fabsf(tau) + sqrtf(1.0f + tau * tau));
#endif
        c = rsqrtf(1.0f + t * t);
        s = t * c;
        // A' = J^T A J with J = [c s; -s c]
        float t_apq = t * apq;
        app -= t_apq;
        aqq += t_apq;
        apq = 0.0f;
        float npr = c * apr - s * aqr;
        float nqr = s * apr + c * aqr;
        apr = npr;
        aqr = nqr;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float a = vp[k], b = vq[k];
            vp[k] = c * a - s * b;
            vq[k] = s * a + c * b;
        }
    }
}

#if !defined(__CUDA_ARCH__)
static inline float rsqrtf_host(float x) { return 1.0f / sqrtf(x); }
#endif

// ---- 3x3: cyclic Jacobi for V, then U and S from B = F V. Column-major. ------------------------------
// The iteration runs on the SHIFTED matrix M = F^T F - I: Jacobi rotations are invariant under the
// shift, and the eigenvalues mu_i of M give
//   sigma_i - 1 = mu_i / (1 + sqrt(1 + mu_i))
// with an error relative to |sigma_i - 1| instead of relative to 1. Stiff materials multiply
// (sigma - 1) by ~1e7..1e9 (linear_elasticity.wgsl:32-35), so an f32 SVD whose singular values are only
// good to 1 ulp of 1.0 injects force noise; here sigma is correctly rounded in all but rare cases, which
// is what the exact-arithmetic statement of the reference's formula evaluates to.
template <int SWEEPS = 4>
__host__ __device__ inline void svd3(const float* F, float* U, float* S, float* V) {
    // Products of two f32 are exact in f64, so M is formed to ~1e-16 and then rounded ONCE to f32:
    // the result is accurate relative to |M| even when F is a large rotation times a small stretch
    // (E = F - I is then O(1) and an f32 assembly would cancel catastrophically). 18 DFMA per particle.
    const double f0 = F[0], f1 = F[1], f2 = F[2], f3 = F[3], f4 = F[4], f5 = F[5], f6 = F[6], f7 = F[7], f8 = F[8];
    float a00 = (float)(f0 * f0 + f1 * f1 + f2 * f2 - 1.0);
    float a11 = (float)(f3 * f3 + f4 * f4 + f5 * f5 - 1.0);
    float a22 = (float)(f6 * f6 + f7 * f7 + f8 * f8 - 1.0);
    float a01 = (float)(f0 * f3 + f1 * f4 + f2 * f5);
    float a02 = (float)(f0 * f6 + f1 * f7 + f2 * f8);
    float a12 = (float)(f3 * f6 + f4 * f7 + f5 * f8);
    float v0[3] = {1, 0, 0}, v1[3] = {0, 1, 0}, v2[3] = {0, 0, 1};
#pragma unroll
    for (int sweep = 0; sweep < SWEEPS; ++sweep) {
        jacobi_rot(a00, a11, a01, a02, a12, v0, v1);
        jacobi_rot(a00, a22, a02, a01, a12, v0, v2);
        jacobi_rot(a11, a22, a12, a01, a02, v1, v2);
    }
    // Sort eigenpairs by decreasing eigenvalue; a swap with one negation keeps det(V) = +1.
#define B2_CSWAP(na, nb, va, vb)          \
    if (na < nb) {                        \
        float tn = na;                    \
        na = nb;                          \
        nb = tn;                          \
        _Pragma("unroll") for (int k = 0; k < 3; ++k) { \
            float tv = va[k];             \
            va[k] = vb[k];                \
            vb[k] = -tv;                  \
        }                                 \
    }
    B2_CSWAP(a00, a11, v0, v1)
    B2_CSWAP(a00, a22, v0, v2)
    B2_CSWAP(a11, a22, v1, v2)
#undef B2_CSWAP
    // B = F V
    float b0[3], b1[3], b2[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        b0[r] = F[r] * v0[0] + F[3 + r] * v0[1] + F[6 + r] * v0[2];
        b1[r] = F[r] * v1[0] + F[3 + r] * v1[1] + F[6 + r] * v1[2];
        b2[r] = F[r] * v2[0] + F[3 + r] * v2[1] + F[6 + r] * v2[2];
    }
    // sigma_i = sqrt(1 + mu_i), sigma_i - 1 = mu_i / (1 + sigma_i); for strongly compressed directions
    // (sigma < ~0.7) 1 + mu_i cancels, and |F v_i| is the accurate expression instead.
    float s0, s1, s2;
    {
        float nb0 = sqrtf(b0[0] * b0[0] + b0[1] * b0[1] + b0[2] * b0[2]);
        float nb1 = sqrtf(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
        float nb2 = sqrtf(b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]);
        s0 = (a00 > -0.5f) ? 1.0f + a00 / (1.0f + sqrtf(1.0f + a00)) : nb0;
        s1 = (a11 > -0.5f) ? 1.0f + a11 / (1.0f + sqrtf(1.0f + a11)) : nb1;
        s2 = (a22 > -0.5f) ? 1.0f + a22 / (1.0f + sqrtf(1.0f + a22)) : nb2;
    }
    float u0[3], u1[3], u2[3];
    float n0 = sqrtf(b0[0] * b0[0] + b0[1] * b0[1] + b0[2] * b0[2]);
    if (n0 > 1e-30f) {
        float inv = 1.0f / n0;
        u0[0] = b0[0] * inv;
        u0[1] = b0[1] * inv;
        u0[2] = b0[2] * inv;
    } else {
        u0[0] = 1.0f;
        u0[1] = 0.0f;
        u0[2] = 0.0f;
    }
    // u1: b1 orthogonalised against u0
    float d01 = u0[0] * b1[0] + u0[1] * b1[1] + u0[2] * b1[2];
    float w[3] = {b1[0] - d01 * u0[0], b1[1] - d01 * u0[1], b1[2] - d01 * u0[2]};
    float nw = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    if (nw > 1e-6f * n0 && nw > 1e-30f) {
        float inv = 1.0f / nw;
        u1[0] = w[0] * inv;
        u1[1] = w[1] * inv;
        u1[2] = w[2] * inv;
    } else { // rank <= 1: any unit vector orthogonal to u0
        float ax = fabsf(u0[0]), ay = fabsf(u0[1]), az = fabsf(u0[2]);
        float e[3] = {0, 0, 0};
        if (ax <= ay && ax <= az) e[0] = 1.0f;
        else if (ay <= az) e[1] = 1.0f;
        else e[2] = 1.0f;
        float de = u0[0] * e[0] + u0[1] * e[1] + u0[2] * e[2];
        float t[3] = {e[0] - de * u0[0], e[1] - de * u0[1], e[2] - de * u0[2]};
        float inv = rsqrtf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
        u1[0] = t[0] * inv;
        u1[1] = t[1] * inv;
        u1[2] = t[2] * inv;
    }
    u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
    u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
    u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
    // det(F) < 0: the last singular value carries the sign.
    if (u2[0] * b2[0] + u2[1] * b2[1] + u2[2] * b2[2] < 0.0f) s2 = -s2;
    S[0] = s0;
    S[1] = s1;
    S[2] = s2;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        U[k] = u0[k];
        U[3 + k] = u1[k];
        U[6 + k] = u2[k];
        V[k] = v0[k];
        V[3 + k] = v1[k];
        V[6 + k] = v2[k];
    }
}

#if defined(__CUDACC__)
// ---- device fast path for the 3x3 SVD of a non-inverted, moderately strained F ---------------------------
// Same mathematics as svd3 (Jacobi on the f64-formed M = F^T F - I), restructured for instruction count:
//   * branch-free rotation  t = sign(d) 2 a_pq / (|d| + sqrt(d^2 + 4 a_pq^2)),  d = a_qq - a_pp  (one MUFU.RSQ,
//     one MUFU.RCP; (c, s) renormalised, so the approximate intrinsics only steer convergence);
//   * rolled sweep loop that stops once |off(M)| <= 1e-6 |diag(M)| (3 sweeps for almost every matrix, f32
//     round-off floor is ~2e-7) - a rolled body also keeps the kernel inside the instruction cache;
//   * no sorting (every use on the path is symmetric in the singular triplets) and no Gram-Schmidt:
//     with det F > 0 and all sigma > 0.2,  u_i = F v_i / sigma_i  directly.
// Also returns sigma_i - 1 (accurate relative to itself), which the strain / stress formulas consume.
// Returns false - caller falls back to svd3 - for det F <= 0 or any sigma <= 0.2 (there sigma - 1 from mu has
// lost more than ~1e-6 of relative accuracy to the cancellation in 1 + mu).
__device__ __forceinline__ void jacobi_rot_fast(float& app, float& aqq, float& apq, float& apr, float& aqr, float* vp,
                                                float* vq) {
    const float d = aqq - app;
    const float two = apq + apq;
    const float w = fmaf(d, d, two * two);
    const float r = w * rsqrtf(fmaxf(w, 1e-37f)); // sqrt(w); 0 for w == 0
    const float den = fmaxf(fabsf(d) + r, 1e-37f);
    const float num = __int_as_float(__float_as_int(two) ^ (__float_as_int(d) & 0x80000000));
    const float t = __fdividef(num, den);
    const float c = rsqrtf(fmaf(t, t, 1.0f));
    const float s = t * c;
    const float t_apq = t * apq;
    app -= t_apq;
    aqq += t_apq;
    apq = 0.0f;
    const float npr = c * apr - s * aqr;
    const float nqr = s * apr + c * aqr;
    apr = npr;
    aqr = nqr;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float a = vp[k], b = vq[k];
        vp[k] = c * a - s * b;
        vq[k] = s * a + c * b;
    }
}

// M = F^T F - I formed in f64 (products of two f32 are exact there) and rounded ONCE to f32, plus det F.
struct ShiftedGram3 {
    float a00, a11, a22, a01, a02, a12, detF;
};
__device__ __forceinline__ ShiftedGram3 shifted_gram3(const float* F) {
    ShiftedGram3 g;
    g.detF = F[0] * (F[4] * F[8] - F[7] * F[5]) - F[3] * (F[1] * F[8] - F[7] * F[2]) + F[6] * (F[1] * F[5] - F[4] * F[2]);
    const double f0 = F[0], f1 = F[1], f2 = F[2], f3 = F[3], f4 = F[4], f5 = F[5], f6 = F[6], f7 = F[7], f8 = F[8];
    g.a00 = (float)(f0 * f0 + f1 * f1 + f2 * f2 - 1.0);
    g.a11 = (float)(f3 * f3 + f4 * f4 + f5 * f5 - 1.0);
    g.a22 = (float)(f6 * f6 + f7 * f7 + f8 * f8 - 1.0);
    g.a01 = (float)(f0 * f3 + f1 * f4 + f2 * f5);
    g.a02 = (float)(f0 * f6 + f1 * f7 + f2 * f8);
    g.a12 = (float)(f3 * f6 + f4 * f7 + f5 * f8);
    return g;
}

__device__ __forceinline__ bool svd3_fast(const float* F, const ShiftedGram3& g, float* U, float* S, float* Sm1, float* V) {
    if (!(g.detF > 0.0f)) return false;
    float a00 = g.a00, a11 = g.a11, a22 = g.a22, a01 = g.a01, a02 = g.a02, a12 = g.a12;
    float v0[3] = {1, 0, 0}, v1[3] = {0, 1, 0}, v2[3] = {0, 0, 1};
#pragma unroll 1
    for (int sweep = 0; sweep < 6; ++sweep) {
        jacobi_rot_fast(a00, a11, a01, a02, a12, v0, v1);
        jacobi_rot_fast(a00, a22, a02, a01, a12, v0, v2);
        jacobi_rot_fast(a11, a22, a12, a01, a02, v1, v2);
        const float off2 = a01 * a01 + a02 * a02; // a12 was just annihilated
        const float dg2 = a00 * a00 + a11 * a11 + a22 * a22;
        if (off2 <= 1e-12f * dg2) break;
    }
    if (!(fminf(a00, fminf(a11, a22)) > -0.96f)) return false; // sigma <= 0.2: 1 + mu cancels, svd3 uses |F v|
    auto finish = [&](int i, float mu, const float* v) {
        const float x = 1.0f + mu;
        const float sm1 = __fdividef(mu, 1.0f + x * rsqrtf(x));
        const float sig = 1.0f + sm1;
        const float inv = __fdividef(1.0f, sig);
        Sm1[i] = sm1;
        S[i] = sig;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float b = F[r] * v[0] + F[3 + r] * v[1] + F[6 + r] * v[2];
            U[3 * i + r] = b * inv;
            V[3 * i + r] = v[r];
        }
    };
    finish(0, a00, v0);
    finish(1, a11, v1);
    finish(2, a22, v2);
    return true;
}
#endif

} // namespace b2
