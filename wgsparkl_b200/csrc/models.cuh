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

// Steps (6)-(8) of the particle update (particle_update.wgsl:95-127): phase/stretch test,
// Drucker-Prager projection, Kirchhoff stress. F is column-major DxD, updated in place.
//   flags: in/out FLAG_PHASE_BROKEN;  plastic: (det, hardening, log_vol_gain, -) in/out.
template <int D, bool PLASTIC>
__device__ __forceinline__ void constitutive_update(const Material& m, uint32_t& flags, float* F, float4& plastic,
                                                    float* tau) {
    float U[D * D], S[D], V[D * D];
    float phase = (flags & FLAG_PHASE_BROKEN) ? 0.0f : m.phase;
    const bool neo = (m.model == B200MPM_MODEL_NEO_HOOKEAN);
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
