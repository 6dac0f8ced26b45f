// "g2p_cdf" pass: per-particle CPIC colour (affinity + sign bits) and MLS-reconstructed signed
// distance / normal to the colliders. Reference: src/solver/g2p_cdf.wgsl:39-63, 124-250.
//
// One CTA per active block; the (BLOCK+2)^D node-cdf tile is staged in shared memory through the
// per-block neighbour table (the reference does 8 hash probes per thread, g2p_cdf.wgsl:65-122).
// Blocks whose tile holds no collider at all write affinity 0 and skip everything else: a zero
// affinity word IS the default cdf (g2p_cdf.wgsl:233-249), so normal / distance are only stored
// for the few particles next to a collider.
#include "launch.h"

namespace b2 {

// Work item = a quarter of a collider-side block (<= ~128 particles) on a CTA of 128 threads: the flagged blocks are
// few (the contact layer), so the kernel is a latency chain per item - more, shorter items finish sooner.
#ifndef CDF_PARTS
#define CDF_PARTS 4u
#endif
constexpr int CDF_THREADS = 128;

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

template <int D>
__global__ void __launch_bounds__(CDF_THREADS) k_g2p_cdf(DeviceData d, int cur) {
    constexpr int B = Dim<D>::BLOCK, LB = Dim<D>::LOG_BLOCK, T = Dim<D>::TILE, TC = Dim<D>::TILE_CELLS;
    constexpr int NA = Dim<D>::NASSOC;
    constexpr int Q = D + 1;
    __shared__ float t_dist[TC];
    __shared__ uint32_t t_aff[TC];
    __shared__ uint32_t s_nbr[NA];
    __shared__ uint32_t s_next;
    const int t = threadIdx.x;
    TL_BEGIN(d, B200MPM_KERNEL_G2P_CDF);
    const uint32_t ncpic = d.counters->num_cpic_blocks;
    const uint32_t num_bodies = d.sim->num_bodies;
    const float h = d.sim->cell_width;
    const float inv_h = 1.0f / h;
    const float4* __restrict__ pos4 = d.pos4[cur];
    const uint32_t* __restrict__ prev_aff = d.cdf_aff[cur];
    uint32_t* __restrict__ new_aff = d.cdf_aff[cur ^ 1];

    while (true) {
        __syncthreads();
        if (t == 0) s_next = atomicAdd(&d.counters->work_cdf, 1u);
        __syncthreads();
        if (s_next >= ncpic * CDF_PARTS) {
            TL_END(d, B200MPM_KERNEL_G2P_CDF);
            break;
        }
        const uint32_t b = d.cpic_list[s_next / CDF_PARTS], part = s_next % CDF_PARTS;
        uint32_t first = d.cell_start[b * CELLS_PER_BLOCK];
        uint32_t last = d.cell_start[(b + 1) * CELLS_PER_BLOCK];
        {
            const uint32_t per = (last - first + CDF_PARTS - 1u) / CDF_PARTS;
            first = min(first + part * per, last);
            last = min(first + per, last);
        }
        if (first == last) continue;
        if (t < NA) s_nbr[t] = d.nbr[b * NA + t];
        __syncthreads();
        int mine = 0;
        for (int n = t; n < TC; n += CDF_THREADS) { // g2p_cdf.wgsl:65-122
            int x = n % T, y = (n / T) % T, z = n / (T * T);
            int ox = x >= B, oy = y >= B, oz = z >= B;
            uint32_t hn = s_nbr[ox + 2 * oy + 4 * oz];
            float dist = 0.0f;
            uint32_t aff = 0u;
            if (hn != NONE) {
                uint4 g = d.node_cdf[hn * CELLS_PER_BLOCK + (x - ox * B) + (y - oy * B) * B + (z - oz * B) * B * B];
                dist = __uint_as_float(g.y);
                aff = g.z;
            }
            t_dist[n] = dist;
            t_aff[n] = aff;
            mine |= (aff != 0u);
        }
        const int any_cdf = __syncthreads_or(mine);
        for (uint32_t k = first + t; k < last; k += CDF_THREADS) {
            if (!any_cdf) {
                new_aff[k] = 0u;
                continue;
            }
            const uint32_t id = d.sorted_ids[k];
            const float4 p4 = pos4[id];
            const uint32_t prev_affinity = prev_aff[id];
            const float pp[3] = {p4.x, p4.y, p4.z};
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
                    for (int sx = 0; sx < 3; ++sx) particle_affinity |= t_aff[tb + sx + T * sy + T * T * sz] & 0xffffu;
            if (particle_affinity == 0u) {
                // No node of the stencil sees a collider: every MLS term is masked out (combined == 0 below), the
                // system is singular and the result is Particle::default_cdf() whatever the sticky signs say.
                new_aff[k] = 0u;
                continue;
            }
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
                            const uint32_t ca = t_aff[idx];
                            const float weight = w[0][sx] * w[1][sy] * ((D == 3) ? w[D - 1][sz] : 1.0f);
                            const float compatible = (ca & bit) ? 1.0f : 0.0f;
                            const float sign = (ca & mask) ? -1.0f : 1.0f;
                            sgn += compatible * weight * sign * t_dist[idx];
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
                        const uint32_t ca = t_aff[idx];
                        const uint32_t combined = ca & particle_affinity & 0xffffu;
                        if (combined == 0u) continue;
                        const uint32_t sign_diff = ((ca >> 16) ^ (particle_affinity >> 16)) & combined;
                        const float weight = w[0][sx] * w[1][sy] * ((D == 3) ? w[D - 1][sz] : 1.0f);
                        const float dist = (sign_diff == 0u) ? t_dist[idx] : -t_dist[idx];
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
            if (det > 1.0e-8f) {
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
                new_aff[k] = particle_affinity;
                if (particle_affinity != 0u) d.cdf_nd[k] = make_float4(nrm[0], nrm[1], nrm[2], res[D]);
            } else {
                new_aff[k] = 0u; // Particle::default_cdf()
            }
        }
    }
}

void launch_g2p_cdf(const LaunchCfg& c, const DeviceData& d, int cur) {
    if (!d.has_bodies || d.n == 0) return;
    const int grid = c.num_sms * 8;
    if (c.dim == 2) k_g2p_cdf<2><<<grid, CDF_THREADS, 0, c.stream>>>(d, cur);
    else k_g2p_cdf<3><<<grid, CDF_THREADS, 0, c.stream>>>(d, cur);
    ++*c.launch_counter;
}

} // namespace b2
