// Constitutive models: corotated ("linear") elasticity, Neo-Hookean elasticity,
// Drucker-Prager return mapping. One SVD per particle serves the stretch test, the plastic
// projection and the stress (the reference does up to three).
#pragma once

#include "common.cuh"
#include "svd.cuh"

namespace b2 {

// DruckerPrager::alpha (drucker_prager.wgsl:25-29). The friction angle is O(1) rad and -h2 q <= 0, so the
// SFU approximations (abs. error ~2^-21) are well inside the f32 noise of the reference's own evaluation.
__device__ __forceinline__ float dp_alpha(const Material& m, float q) {
    float angle = m.dp_h0 + (m.dp_h1 * q - m.dp_h3) * __expf(-m.dp_h2 * q);
    float s = __sinf(angle);
    return 0.81649658092772603f * __fdividef(2.0f * s, 3.0f - s);
}

// log(1 + x) from an x that is accurate relative to itself (sigma - 1 out of the shifted SVD):
// 2 atanh(x / (2 + x)), series to z^9 for |x| < 1/4 (remainder < 4e-10 relative), logf beyond.
__device__ __forceinline__ float log1p_strain(float x, float one_plus_x) {
    if (fabsf(x) < 0.25f) {
        const float z = __fdividef(x, 2.0f + x), z2 = z * z;
        float p = fmaf(z2, 1.0f / 9.0f, 1.0f / 7.0f);
        p = fmaf(z2, p, 0.2f);
        p = fmaf(z2, p, 1.0f / 3.0f);
        p = fmaf(z2, p, 1.0f);
        return 2.0f * z * p;
    }
    return logf(one_plus_x);
}

// exp(x) - 1, Taylor to x^7 for |x| < 1/4 (remainder < 2e-9 relative), expf beyond.
__device__ __forceinline__ float expm1_strain(float x) {
    if (fabsf(x) < 0.25f) {
        float p = fmaf(x, 1.0f / 5040.0f, 1.0f / 720.0f);
        p = fmaf(x, p, 1.0f / 120.0f);
        p = fmaf(x, p, 1.0f / 24.0f);
        p = fmaf(x, p, 1.0f / 6.0f);
        p = fmaf(x, p, 0.5f);
        p = fmaf(x, p, 1.0f);
        return x * p;
    }
    return expf(x) - 1.0f;
}

// project_deformation_gradient (drucker_prager.wgsl:43-64 2D, 112-133 3D). Returns false when the
// projection is invalid (gamma <= 0: state and F are left untouched by the caller).
//   sv / svm1: singular values and (sigma - 1); outputs new_sv / new_svm1, the hardening increment, and
//   log(prod sv) - log(prod new_sv) (the log-volume the projection removed, drucker_prager.wgsl:152).
template <int D>
__device__ __forceinline__ bool dp_project_sv(const Material& m, const float* sv, const float* svm1, float log_vol_gain,
                                              float alpha, float* new_sv, float* new_svm1, float& hardening,
                                              float& log_det_ratio) {
    const float d = (float)D;
    float strain[D], dev[D];
    float trace = 0.0f, log_det = 0.0f;
    const float shift = log_vol_gain * (1.0f / d);
#pragma unroll
    for (int i = 0; i < D; ++i) {
        const float l = log1p_strain(svm1[i], sv[i]);
        log_det = (i == 0) ? l : log_det + l;
        strain[i] = l + shift;
        trace = (i == 0) ? strain[0] : trace + strain[i];
    }
    bool all_zero = true;
    float dev2 = 0.0f, strain2 = 0.0f;
    const float mean = trace * (1.0f / d);
#pragma unroll
    for (int i = 0; i < D; ++i) {
        dev[i] = strain[i] - mean;
        all_zero = all_zero && (dev[i] == 0.0f);
        dev2 += dev[i] * dev[i];
        strain2 += strain[i] * strain[i];
    }
    if (trace > 0.0f || all_zero) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            new_sv[i] = 1.0f;
            new_svm1[i] = 0.0f;
        }
        hardening = sqrtf(strain2);
        log_det_ratio = log_det;
        return true;
    }
    float dev_norm = sqrtf(dev2);
    float gamma = dev_norm + m.dp_ratio * trace * alpha;
    if (gamma <= 0.0f) return false;
    float k = __fdividef(gamma, dev_norm);
    float new_log_det = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        const float e = strain[i] - dev[i] * k;
        new_log_det += e;
        new_svm1[i] = expm1_strain(e);
        new_sv[i] = 1.0f + new_svm1[i];
    }
    hardening = gamma;
    log_det_ratio = log_det - new_log_det;
    return true;
}

// Corotated Kirchhoff stress (linear_elasticity.wgsl:28-41) WITHOUT an SVD, for strains up to ~10 %.
//   U (Sigma - I) V^T F^T = (F - R) F^T with R = U V^T = F S^-1, S = (F^T F)^(1/2), and J = det F, so with
//   M = F^T F - I (formed exactly in f64, rounded once):
//     F - R = F X,   X = I - (I + M)^(-1/2) = M/2 - 3 M^2/8 + 5 M^3/16 - 35 M^4/128 + 63 M^5/256 - 231 M^6/1024
//     J^2 - 1 = det(I + M) - 1 = tr M + ((tr M)^2 - tr M^2) / 2 + det M,   J - 1 = (J^2 - 1) / (1 + J)
//   Everything is a polynomial in the SMALL matrix M, so (Sigma - 1) and (J - 1) keep their relative accuracy
//   (the stiffness multiplies them by 1e7..1e9), at ~1/3 of the instructions of the Jacobi SVD.
// Valid while small_strain3() holds (|M|_F <= 0.1, F not inverted); the caller takes the SVD path otherwise.
__device__ __forceinline__ bool small_strain3(const ShiftedGram3& g) {
    const float tr2 = g.a00 * g.a00 + g.a11 * g.a11 + g.a22 * g.a22 + 2.0f * (g.a01 * g.a01 + g.a02 * g.a02 + g.a12 * g.a12);
    return (tr2 <= 0.01f) && (g.detF > 0.0f);
}
__device__ __forceinline__ void corotated_stress_small_strain3(const Material& m, const float* F, const ShiftedGram3& g, float* tau) {
    const float m00 = g.a00, m11 = g.a11, m22 = g.a22, m01 = g.a01, m02 = g.a02, m12 = g.a12;
    const float tr2 = m00 * m00 + m11 * m11 + m22 * m22 + 2.0f * (m01 * m01 + m02 * m02 + m12 * m12); // tr M^2 = |M|_F^2
    // Horner: X = M (c1 + M (c2 + M (c3 + M (c4 + M (c5 + c6 M))))), symmetric 3x3 as (00, 11, 22, 01, 02, 12)
    float q00 = -0.2255859375f * m00 + 0.24609375f, q11 = -0.2255859375f * m11 + 0.24609375f,
          q22 = -0.2255859375f * m22 + 0.24609375f;
    float q01 = -0.2255859375f * m01, q02 = -0.2255859375f * m02, q12 = -0.2255859375f * m12;
    const float coef[4] = {-0.2734375f, 0.3125f, -0.375f, 0.5f};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        // P = M * Q (M and Q commute, P is symmetric), then Q = c_k I + P (no shift after the last product)
        float p00 = m00 * q00 + m01 * q01 + m02 * q02;
        float p01 = m00 * q01 + m01 * q11 + m02 * q12;
        float p02 = m00 * q02 + m01 * q12 + m02 * q22;
        float p11 = m01 * q01 + m11 * q11 + m12 * q12;
        float p12 = m01 * q02 + m11 * q12 + m12 * q22;
        float p22 = m02 * q02 + m12 * q12 + m22 * q22;
        const float c = (k < 4) ? coef[k] : 0.0f;
        q00 = p00 + c, q11 = p11 + c, q22 = p22 + c;
        q01 = p01, q02 = p02, q12 = p12;
    }
    // (q is X now.)  J - 1
    const float trm = m00 + m11 + m22;
    const float detm = m00 * (m11 * m22 - m12 * m12) - m01 * (m01 * m22 - m12 * m02) + m02 * (m01 * m12 - m11 * m02);
    const float dd = trm + 0.5f * (trm * trm - tr2) + detm;
    const float J = sqrtf(1.0f + dd);
    const float diag = m.lambda * (dd / (1.0f + J)) * J;
    // T = F X ; tau = 2 mu T F^T + diag I (symmetric)
    float T[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float a = F[r], b = F[3 + r], c = F[6 + r];
        T[r] = a * q00 + b * q01 + c * q02;
        T[3 + r] = a * q01 + b * q11 + c * q12;
        T[6 + r] = a * q02 + b * q12 + c * q22;
    }
    const float mu2 = 2.0f * m.mu;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (r > c) continue;
            float s = T[r] * F[c] + T[3 + r] * F[3 + c] + T[6 + r] * F[6 + c]; // (T F^T)[r][c]
            s = mu2 * s + ((r == c) ? diag : 0.0f);
            tau[c * 3 + r] = s;
            tau[r * 3 + c] = s;
        }
}

// Steps (6)-(8) of the particle update (particle_update.wgsl:95-127): phase/stretch test,
// Drucker-Prager projection, Kirchhoff stress. F is column-major DxD, updated in place.
//   flags: in/out FLAG_PHASE_BROKEN;  plastic: (det, hardening, log_vol_gain, -) in/out.
template <int D, bool PLASTIC>
__device__ __forceinline__ void constitutive_update(const Material& m, uint32_t& flags, float* F, float4& plastic,
                                                    float* tau) {
    const bool neo = (m.model == B200MPM_MODEL_NEO_HOOKEAN);
    ShiftedGram3 gram;
    if (D == 3) gram = shifted_gram3(F);
    if (D == 3 && !PLASTIC && !neo) {
        // The choice is made per WARP: a warp with even one strongly strained particle would execute both paths
        // one after the other, so it takes the SVD path for everybody instead (same result, one pass).
        const bool small = small_strain3(gram);
        if (__ballot_sync(__activemask(), !small) == 0u) {
            corotated_stress_small_strain3(m, F, gram, tau);
            return;
        }
    }
    float U[D * D], S[D], Sm1[D], V[D * D];
    float phase = (flags & FLAG_PHASE_BROKEN) ? 0.0f : m.phase;
    const bool may_stretch = PLASTIC && phase > 0.0f && m.max_stretch > 0.0f;
    const bool need_svd = !neo || (PLASTIC && (may_stretch || phase == 0.0f));
    if (need_svd) {
        bool have = false;
        if (D == 3) have = svd3_fast(F, gram, U, S, Sm1, V);
        if (!have) {
            if (D == 2) svd2(F, U, S, V);
            else svd3<4>(F, U, S, V);
#pragma unroll
            for (int i = 0; i < D; ++i) Sm1[i] = S[i] - 1.0f;
        }
    }
    if (PLASTIC) {
        if (may_stretch) { // particle_update.wgsl:101-115
            bool broken = false;
#pragma unroll
            for (int i = 0; i < D; ++i) broken = broken || (S[i] > m.max_stretch);
            if (broken) {
                flags |= FLAG_PHASE_BROKEN;
                phase = 0.0f;
            }
        }
        if (phase == 0.0f && m.dp_lambda != 0.0f) { // particle_update.wgsl:118-122, drucker_prager.wgsl:134-158
            float alpha = dp_alpha(m, plastic.y);
            float nsv[D], nsm1[D], hard, log_ratio;
            if (dp_project_sv<D>(m, S, Sm1, plastic.z, alpha, nsv, nsm1, hard, log_ratio)) {
                float prev_det = S[0], new_det = nsv[0];
#pragma unroll
                for (int i = 1; i < D; ++i) {
                    prev_det *= S[i];
                    new_det *= nsv[i];
                }
                plastic.x = plastic.x * __fdividef(prev_det, new_det);
                plastic.z = plastic.z + log_ratio;
                plastic.y = plastic.y + hard;
                // F <- U diag(S') V^T, written as F + U diag(S' - S) V^T: the plastic correction is small, so the
                // rounding errors of U, V and S' only touch the correction and F keeps ~1 ulp accuracy.
                float UdS[D * D];
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const float ds = nsm1[k] - Sm1[k];
#pragma unroll
                    for (int r = 0; r < D; ++r) UdS[k * D + r] = U[k * D + r] * ds;
                    S[k] = nsv[k];
                    Sm1[k] = nsm1[k];
                }
#pragma unroll
                for (int c = 0; c < D; ++c)
#pragma unroll
                    for (int r = 0; r < D; ++r) {
                        float s = F[c * D + r];
#pragma unroll
                        for (int k = 0; k < D; ++k) s += UdS[k * D + r] * V[k * D + c];
                        F[c * D + r] = s;
                    }
            }
        }
    }
    if (neo) { // neo_hookean_elasticity.wgsl:11-26
        float det;
        if (D == 2) det = F[0] * F[3] - F[2] * F[1];
        else
            det = F[0] * (F[4] * F[8] - F[7] * F[5]) - F[3] * (F[1] * F[8] - F[7] * F[2]) + F[6] * (F[1] * F[5] - F[4] * F[2]);
        float j = fmaxf(det, 1.0e-10f);
        float diag = m.lambda * logf(j) - m.mu;
#pragma unroll
        for (int c = 0; c < D; ++c)
#pragma unroll
            for (int r = 0; r < D; ++r) {
                float s = 0.0f;
#pragma unroll
                for (int k = 0; k < D; ++k) s += F[k * D + r] * F[k * D + c]; // (F F^T)[r][c]
                tau[c * D + r] = m.mu * s + ((r == c) ? diag : 0.0f);
            }
    } else { // corotated: linear_elasticity.wgsl:14-41
        // 2 mu U (S - I) V^T F^T + lambda (J - 1) J I, with F^T = V S U^T:
        //   = U diag(2 mu (S - 1) S + lambda (J - 1) J) U^T
        // J - 1 = prod(1 + (S_i - 1)) - 1 expanded, so it keeps the relative accuracy of (S_i - 1)
        float jm1 = Sm1[0];
#pragma unroll
        for (int i = 1; i < D; ++i) jm1 = fmaf(jm1, Sm1[i], jm1 + Sm1[i]);
        float diag = m.lambda * jm1 * (1.0f + jm1);
        float Ue[D * D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const float e = 2.0f * m.mu * Sm1[i] * S[i] + diag;
#pragma unroll
            for (int r = 0; r < D; ++r) Ue[i * D + r] = U[i * D + r] * e;
        }
#pragma unroll
        for (int c = 0; c < D; ++c)
#pragma unroll
            for (int r = 0; r < D; ++r) {
                if (r > c) continue;
                float s = 0.0f;
#pragma unroll
                for (int k = 0; k < D; ++k) s += Ue[k * D + r] * U[k * D + c];
                tau[c * D + r] = s;
                tau[r * D + c] = s;
            }
    }
}

} // namespace b2
