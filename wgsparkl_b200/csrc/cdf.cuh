// "g2p_cdf" pass: per-particle CPIC colour (affinity + sign bits) and MLS-reconstructed signed distance / normal to the
// colliders. Reference: src/solver/g2p_cdf.wgsl:39-63, 124-250.
//
// The reference runs it as a dispatch of its own over every particle. Here it is a device function that the P2G
// kernel calls for the particles of collider-side blocks while they sit in its staging buffers (p2g.cu): the node
// colours it needs are the (BLOCK+2)^D tile that kernel stages anyway, the result feeds the compatibility tests of
// the same work item, and no separate latency-bound kernel sits between the sort and P2G. Particles of blocks whose
// tile holds no collider keep the zero affinity word k_scatter wrote - a zero word IS the default cdf
// (g2p_cdf.wgsl:233-249) - so normal / distance are only stored for the few particles next to a collider.
#pragma once

#include "common.cuh"

namespace b2 {

template <int Q>
struct SmallMat {
    float m[Q][Q];
};
__device__ __forceinline__ float det3(const float a[3][3]) {
    return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
           a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
}
__device__ __forceinline__ float minor4(const SmallMat<4>& a, int r, int c) {
    float s[3][3];
    int ii = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i == r) continue;
        int jj = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j == c) continue;
            s[ii][jj++] = a.m[i][j];
        }
        ++ii;
    }
    return det3(s);
}
// Solves M x = b through the cofactor inverse (Inv::inv3 / inv4, g2p_cdf.wgsl:236,242); returns det(M).
__device__ __forceinline__ float det_and_solve(const SmallMat<3>& a, const float* b, float* x) {
    float d = det3(a.m);
    float id = 1.0f / d;
    float inv[3][3];
    inv[0][0] = (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) * id;
    inv[0][1] = (a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2]) * id;
    inv[0][2] = (a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1]) * id;
    inv[1][0] = (a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2]) * id;
    inv[1][1] = (a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0]) * id;
    inv[1][2] = (a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2]) * id;
    inv[2][0] = (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]) * id;
    inv[2][1] = (a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1]) * id;
    inv[2][2] = (a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0]) * id;
#pragma unroll
    for (int i = 0; i < 3; ++i) x[i] = inv[i][0] * b[0] + inv[i][1] * b[1] + inv[i][2] * b[2];
    return d;
}
__device__ __forceinline__ float det_and_solve(const SmallMat<4>& a, const float* b, float* x) {
    float cof0[4];
    float d = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        cof0[j] = minor4(a, 0, j);
        d += ((j & 1) ? -1.0f : 1.0f) * a.m[0][j] * cof0[j];
    }
    float id = 1.0f / d;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float cof = (j == 0) ? cof0[i] : minor4(a, j, i); // inv[i][j] = (-1)^(i+j) minor(j,i) / det
            float inv = (((i + j) & 1) ? -1.0f : 1.0f) * cof * id;
            s = (j == 0) ? inv * b[0] : s + inv * b[j];
        }
        x[i] = s;
    }
    return d;
}

// Does any node of the particle's stencil see a collider? (the OR of the 3^D node affinity masks, g2p_cdf.wgsl:152-165) -
// the cheap first half of the colouring: particles for which it is zero keep the default colour.
template <int D, class AffTile>
__device__ __forceinline__ uint32_t cdf_stencil_affinity(const float* pp, float h, float inv_h, const AffTile& t_aff) {
    constexpr int B = Dim<D>::BLOCK, T = Dim<D>::TILE;
    int tb = 0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        const int c = (int)(round_div(pp[a], h, inv_h) - 1.0f);
        tb += (c & (B - 1)) * ((a == 0) ? 1 : (a == 1) ? T : T * T);
    }
    uint32_t m = 0u;
#pragma unroll
    for (int sz = 0; sz < (D == 3 ? 3 : 1); ++sz)
#pragma unroll
        for (int sy = 0; sy < 3; ++sy)
#pragma unroll
            for (int sx = 0; sx < 3; ++sx) m |= t_aff(tb + sx + T * sy + T * T * sz) & 0xffffu;
    return m;
}

// Colour of one particle. t_aff / t_dist: the (BLOCK+2)^D tile of node affinities / unsigned distances of the particle's
// block (flat index lx + T ly + T^2 lz). Returns the affinity word; nd = (normal, signed distance) is only meaningful
// when the word is non-zero.
template <int D, class AffTile, class DistTile>
__device__ __forceinline__ uint32_t cdf_colour_particle(const float* pp, uint32_t prev_affinity, uint32_t num_bodies, float h,
                                                        float inv_h, const AffTile& t_aff, const DistTile& t_dist, float4& nd) {
    constexpr int B = Dim<D>::BLOCK, T = Dim<D>::TILE;
    constexpr int Q = D + 1;
    float d0[D], w[D][3];
    int tb = 0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        float cf = round_div(pp[a], h, inv_h) - 1.0f;
        int c = (int)cf;
        int l = c & (B - 1);
        tb += l * ((a == 0) ? 1 : (a == 1) ? T : T * T);
        d0[a] = cf * h - pp[a];
        bspline(-d0[a] * inv_h, w[a][0], w[a][1], w[a][2]);
    }
    // Affinity mask + sign bits (Eqn. 21; g2p_cdf.wgsl:152-190)
    uint32_t particle_affinity = 0u;
#pragma unroll
    for (int sz = 0; sz < (D == 3 ? 3 : 1); ++sz)
#pragma unroll
        for (int sy = 0; sy < 3; ++sy)
#pragma unroll
            for (int sx = 0; sx < 3; ++sx) particle_affinity |= t_aff(tb + sx + T * sy + T * T * sz) & 0xffffu;
    // No node of the stencil sees a collider: every MLS term is masked out (combined == 0 below), the system is
    // singular and the result is Particle::default_cdf() whatever the sticky signs say.
    if (particle_affinity == 0u) return 0u;
    for (uint32_t ic = 0; ic < num_bodies; ++ic) {
        const uint32_t bit = 1u << ic, mask = 1u << (ic + 16);
        if ((prev_affinity & bit) != 0u) { // sticky sign, even if the affinity bit is gone (g2p_cdf.wgsl:186-188)
            particle_affinity |= prev_affinity & mask;
            continue;
        }
        if ((particle_affinity & bit) == 0u) continue; // no node sees this collider: the sum is 0, no sign
        float sgn = 0.0f;
#pragma unroll
        for (int sz = 0; sz < (D == 3 ? 3 : 1); ++sz)
#pragma unroll
            for (int sy = 0; sy < 3; ++sy)
#pragma unroll
                for (int sx = 0; sx < 3; ++sx) {
                    const int idx = tb + sx + T * sy + T * T * sz;
                    const uint32_t ca = t_aff(idx);
                    const float weight = w[0][sx] * w[1][sy] * ((D == 3) ? w[D - 1][sz] : 1.0f);
                    const float compatible = (ca & bit) ? 1.0f : 0.0f;
                    const float sign = (ca & mask) ? -1.0f : 1.0f;
                    sgn += compatible * weight * sign * t_dist(idx);
                }
        if (sgn < 0.0f) particle_affinity |= mask;
    }
    // MLS reconstruction (Eq. 4; g2p_cdf.wgsl:192-232)
    SmallMat<Q> qtq;
    float qtu[Q];
#pragma unroll
    for (int a = 0; a < Q; ++a) {
        qtu[a] = 0.0f;
#pragma unroll
        for (int c = 0; c < Q; ++c) qtq.m[a][c] = 0.0f;
    }
#pragma unroll
    for (int sz = 0; sz < (D == 3 ? 3 : 1); ++sz)
#pragma unroll
        for (int sy = 0; sy < 3; ++sy)
#pragma unroll
            for (int sx = 0; sx < 3; ++sx) {
                const int idx = tb + sx + T * sy + T * T * sz;
                const uint32_t ca = t_aff(idx);
                const uint32_t combined = ca & particle_affinity & 0xffffu;
                if (combined == 0u) continue;
                const uint32_t sign_diff = ((ca >> 16) ^ (particle_affinity >> 16)) & combined;
                const float weight = w[0][sx] * w[1][sy] * ((D == 3) ? w[D - 1][sz] : 1.0f);
                const float dist = (sign_diff == 0u) ? t_dist(idx) : -t_dist(idx);
                float p[Q];
                p[0] = d0[0] + (float)sx * h;
                p[1] = d0[1] + (float)sy * h;
                if (D == 3) p[D - 1] = d0[D - 1] + (float)sz * h;
                p[D] = 1.0f;
#pragma unroll
                for (int a = 0; a < Q; ++a) {
#pragma unroll
                    for (int c = 0; c < Q; ++c) qtq.m[a][c] += (p[a] * p[c]) * weight;
                    qtu[a] += p[a] * weight * dist;
                }
            }
    float res[Q];
    float det = det_and_solve(qtq, qtu, res);
    if (!(det > 1.0e-8f)) return 0u; // Particle::default_cdf()
    float len = 0.0f;
#pragma unroll
    for (int a = 0; a < D; ++a) len = (a == 0) ? res[0] * res[0] : len + res[a] * res[a];
    len = sqrtf(len);
    float nrm[3] = {0.0f, 0.0f, 0.0f};
    if (D == 2) {
        if (len > 1.0e-6f) {
            nrm[0] = res[0] / len;
            nrm[1] = res[1] / len;
        }
    } else {
#pragma unroll
        for (int a = 0; a < D; ++a) nrm[a] = res[a] / len;
    }
    nd = make_float4(nrm[0], nrm[1], nrm[2], res[D]);
    return particle_affinity;
}

} // namespace b2
