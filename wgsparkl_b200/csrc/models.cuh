// Constitutive models: corotated ("linear") elasticity, Neo-Hookean elasticity,
// Drucker-Prager return mapping. One SVD per particle serves the stretch test, the plastic
// projection and the stress (the reference does up to three).
#pragma once

#include "common.cuh"
#include "svd.cuh"

namespace b2 {

// DruckerPrager::alpha (drucker_prager.wgsl:25-29)
__device__ __forceinline__ float dp_alpha(const Material& m, float q) {
    float angle = m.dp_h0 + (m.dp_h1 * q - m.dp_h3) * expf(-m.dp_h2 * q);
    float s = sinf(angle);
    return sqrtf(2.0f / 3.0f) * (2.0f * s) / (3.0f - s);
}

// project_deformation_gradient (drucker_prager.wgsl:43-64 2D, 112-133 3D). Returns false when the
// projection is invalid (gamma <= 0: state and F are left untouched by the caller).
template <int D>
__device__ __forceinline__ bool dp_project_sv(const Material& m, const float* sv, float log_vol_gain, float alpha,
                                              float* new_sv, float& hardening) {
    const float d = (float)D;
    float strain[D], dev[D];
    float trace = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        strain[i] = logf(sv[i]) + log_vol_gain / d;
        trace = (i == 0) ? strain[0] : trace + strain[i];
    }
    bool all_zero = true;
    float dev2 = 0.0f, strain2 = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        dev[i] = strain[i] - trace / d;
        all_zero = all_zero && (dev[i] == 0.0f);
        dev2 += dev[i] * dev[i];
        strain2 += strain[i] * strain[i];
    }
    if (trace > 0.0f || all_zero) {
#pragma unroll
        for (int i = 0; i < D; ++i) new_sv[i] = 1.0f;
        hardening = sqrtf(strain2);
        return true;
    }
    float dev_norm = sqrtf(dev2);
    float gamma = dev_norm + (d * m.dp_lambda + 2.0f * m.dp_mu) / (2.0f * m.dp_mu) * trace * alpha;
    if (gamma <= 0.0f) return false;
    float k = gamma / dev_norm;
#pragma unroll
    for (int i = 0; i < D; ++i) new_sv[i] = expf(strain[i] - dev[i] * k);
    hardening = gamma;
    return true;
}

// Corotated Kirchhoff stress (linear_elasticity.wgsl:28-41) WITHOUT an SVD, for strains up to ~10 %.
//   U (Sigma - I) V^T F^T = (F - R) F^T with R = U V^T = F S^-1, S = (F^T F)^(1/2), and J = det F, so with
//   M = F^T F - I (formed exactly in f64, rounded once):
//     F - R = F X,   X = I - (I + M)^(-1/2) = M/2 - 3 M^2/8 + 5 M^3/16 - 35 M^4/128 + 63 M^5/256 - 231 M^6/1024
//     J^2 - 1 = det(I + M) - 1 = tr M + ((tr M)^2 - tr M^2) / 2 + det M,   J - 1 = (J^2 - 1) / (1 + J)
//   Everything is a polynomial in the SMALL matrix M, so (Sigma - 1) and (J - 1) keep their relative accuracy
//   (the stiffness multiplies them by 1e7..1e9), at ~1/3 of the instructions of the Jacobi SVD.
// Returns false (caller falls back to the SVD) when the strain is too large or F is inverted.
__device__ __forceinline__ bool corotated_stress_small_strain3(const Material& m, const float* F, float* tau) {
    const double f0 = F[0], f1 = F[1], f2 = F[2], f3 = F[3], f4 = F[4], f5 = F[5], f6 = F[6], f7 = F[7], f8 = F[8];
    const float m00 = (float)(f0 * f0 + f1 * f1 + f2 * f2 - 1.0);
    const float m11 = (float)(f3 * f3 + f4 * f4 + f5 * f5 - 1.0);
    const float m22 = (float)(f6 * f6 + f7 * f7 + f8 * f8 - 1.0);
    const float m01 = (float)(f0 * f3 + f1 * f4 + f2 * f5);
    const float m02 = (float)(f0 * f6 + f1 * f7 + f2 * f8);
    const float m12 = (float)(f3 * f6 + f4 * f7 + f5 * f8);
    const float tr2 = m00 * m00 + m11 * m11 + m22 * m22 + 2.0f * (m01 * m01 + m02 * m02 + m12 * m12); // tr M^2 = |M|_F^2
    const float detF = F[0] * (F[4] * F[8] - F[7] * F[5]) - F[3] * (F[1] * F[8] - F[7] * F[2]) + F[6] * (F[1] * F[5] - F[4] * F[2]);
    if (!(tr2 <= 0.01f) || !(detF > 0.0f)) return false;
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
    return true;
}

// Steps (6)-(8) of the particle update (particle_update.wgsl:95-127): phase/stretch test,
// Drucker-Prager projection, Kirchhoff stress. F is column-major DxD, updated in place.
//   flags: in/out FLAG_PHASE_BROKEN;  plastic: (det, hardening, log_vol_gain, -) in/out.
template <int D, bool PLASTIC>
__device__ __forceinline__ void constitutive_update(const Material& m, uint32_t& flags, float* F, float4& plastic,
                                                    float* tau) {
    const bool neo = (m.model == B200MPM_MODEL_NEO_HOOKEAN);
    if (D == 3 && !PLASTIC && !neo) {
        if (corotated_stress_small_strain3(m, F, tau)) return;
    }
    float U[D * D], S[D], V[D * D];
    float phase = (flags & FLAG_PHASE_BROKEN) ? 0.0f : m.phase;
    const bool may_stretch = PLASTIC && phase > 0.0f && m.max_stretch > 0.0f;
    const bool need_svd = !neo || (PLASTIC && (may_stretch || phase == 0.0f));
    if (need_svd) {
        if (D == 2) svd2(F, U, S, V);
        else svd3<4>(F, U, S, V);
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
            float nsv[D], hard;
            if (dp_project_sv<D>(m, S, plastic.z, alpha, nsv, hard)) {
                float prev_det = S[0], new_det = nsv[0];
#pragma unroll
                for (int i = 1; i < D; ++i) {
                    prev_det *= S[i];
                    new_det *= nsv[i];
                }
                plastic.x = plastic.x * prev_det / new_det;
                plastic.z = plastic.z + logf(prev_det) - logf(new_det);
                plastic.y = plastic.y + hard;
#pragma unroll
                for (int i = 0; i < D; ++i) S[i] = nsv[i];
                // F = U diag(S) V^T
#pragma unroll
                for (int c = 0; c < D; ++c)
#pragma unroll
                    for (int r = 0; r < D; ++r) {
                        float s = 0.0f;
#pragma unroll
                        for (int k = 0; k < D; ++k) s += U[k * D + r] * S[k] * V[k * D + c];
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
        float j = S[0];
#pragma unroll
        for (int i = 1; i < D; ++i) j *= S[i];
        float diag = m.lambda * (j - 1.0f) * j;
        float e[D];
#pragma unroll
        for (int i = 0; i < D; ++i) e[i] = 2.0f * m.mu * (S[i] - 1.0f) * S[i] + diag;
#pragma unroll
        for (int c = 0; c < D; ++c)
#pragma unroll
            for (int r = 0; r < D; ++r) {
                float s = 0.0f;
#pragma unroll
                for (int k = 0; k < D; ++k) s += U[k * D + r] * e[k] * U[k * D + c];
                tau[c * D + r] = s;
            }
    }
}

} // namespace b2
