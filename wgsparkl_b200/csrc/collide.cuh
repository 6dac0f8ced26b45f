// Analytic collider projection and the per-node collision-detection field.
// Replaces wgparry's Shape::projectPointOnBoundary (not vendored; collide.wgsl:39-41) for
// balls, cuboids and capsules, and collide() (src/collision/collide.wgsl:23-55).
#pragma once

#include "common.cuh"

namespace b2 {

struct NodeCdf { // grid.wgsl:233-240
    float distance;
    uint32_t affinities;
    uint32_t closest_id;
};

// parry's project_local_point(pt, solid = false) on the shape boundary, local frame.
template <int D>
__host__ __device__ inline bool project_local_point_on_boundary(const BodyDev& b, const float* pt, float* out) {
    if (b.shape_type == B200MPM_SHAPE_BALL) {
        float d2 = 0.0f;
#pragma unroll
        for (int i = 0; i < D; ++i) d2 += pt[i] * pt[i];
        bool inside = d2 <= b.radius * b.radius;
        if (d2 == 0.0f) {
#pragma unroll
            for (int i = 0; i < D; ++i) out[i] = 0.0f;
            out[1] = b.radius;
        } else {
            float s = b.radius / sqrtf(d2);
#pragma unroll
            for (int i = 0; i < D; ++i) out[i] = pt[i] * s;
        }
        return inside;
    }
    if (b.shape_type == B200MPM_SHAPE_CUBOID) {
        float mins_pt[D], pt_maxs[D];
        bool inside = true;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            mins_pt[i] = -b.shape_a[i] - pt[i];
            pt_maxs[i] = pt[i] - b.shape_a[i];
            float shift = fmaxf(mins_pt[i], 0.0f) - fmaxf(pt_maxs[i], 0.0f);
            out[i] = pt[i] + shift;
            if (shift != 0.0f) inside = false;
        }
        if (!inside) return false;
        float best = -3.402823466e38f;
        bool is_mins = false;
        int best_id = 0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            if (mins_pt[i] < pt_maxs[i]) {
                if (pt_maxs[i] > best) {
                    best_id = i;
                    is_mins = false;
                    best = pt_maxs[i];
                }
            } else if (mins_pt[i] > best) {
                best_id = i;
                is_mins = true;
                best = mins_pt[i];
            }
        }
#pragma unroll
        for (int i = 0; i < D; ++i) out[i] = (i == best_id) ? pt[i] + (is_mins ? best : -best) : pt[i];
        return true;
    }
    // capsule
    float ab[D], ap[D];
    float ab_ap = 0.0f, sqnab = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        ab[i] = b.shape_b[i] - b.shape_a[i];
        ap[i] = pt[i] - b.shape_a[i];
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
        ab_ap = (i == 0) ? ab[0] * ap[0] : ab_ap + ab[i] * ap[i];
        sqnab = (i == 0) ? ab[0] * ab[0] : sqnab + ab[i] * ab[i];
    }
    float seg[D];
    if (ab_ap <= 0.0f) {
#pragma unroll
        for (int i = 0; i < D; ++i) seg[i] = b.shape_a[i];
    } else if (ab_ap >= sqnab) {
#pragma unroll
        for (int i = 0; i < D; ++i) seg[i] = b.shape_b[i];
    } else {
        float t = ab_ap / sqnab;
#pragma unroll
        for (int i = 0; i < D; ++i) seg[i] = b.shape_a[i] + ab[i] * t;
    }
    float dp[D], dist2 = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        dp[i] = pt[i] - seg[i];
        dist2 = (i == 0) ? dp[0] * dp[0] : dist2 + dp[i] * dp[i];
    }
    float dist = sqrtf(dist2);
    if (dist > 1.1920929e-7f) {
        float s = b.radius / dist;
#pragma unroll
        for (int i = 0; i < D; ++i) out[i] = seg[i] + dp[i] * s;
        return dist <= b.radius;
    }
    float dir[3] = {0.0f, 0.0f, 0.0f};
    float n = sqrtf(sqnab);
    if (n > 0.0f) {
        if (D == 2) {
            dir[0] = -ab[1] / n;
            dir[1] = ab[0] / n;
        } else {
            float ux = ab[0] / n, uy = ab[1] / n, uz = ab[D - 1] / n;
            float ex = (fabsf(ux) < 0.9f) ? 1.0f : 0.0f, ey = 1.0f - ex;
            float ox = uy * 0.0f - uz * ey, oy = uz * ex - ux * 0.0f, oz = ux * ey - uy * ex;
            float l = sqrtf(ox * ox + oy * oy + oz * oz);
            dir[0] = ox / l;
            dir[1] = oy / l;
            dir[2] = oz / l;
        }
    } else {
        dir[1] = 1.0f;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) out[i] = seg[i] + dir[i] * b.radius;
    return true;
}

// World point -> body frame -> boundary projection -> world: the vector from `point` to its projection on body b's
// boundary, and whether the point is inside (collide.wgsl:39-45).
template <int D>
__host__ __device__ inline bool project_on_body(const BodyDev& b, const float* point, float* dpt) {
    float d[D], loc[D], lp[D];
#pragma unroll
    for (int k = 0; k < D; ++k) d[k] = point[k] - b.trans[k];
#pragma unroll
    for (int r = 0; r < D; ++r) { // local = R^T (point - t)
        float s = b.rot[r * D + 0] * d[0];
#pragma unroll
        for (int k = 1; k < D; ++k) s = s + b.rot[r * D + k] * d[k];
        loc[r] = s;
    }
    const bool inside = project_local_point_on_boundary<D>(b, loc, lp);
#pragma unroll
    for (int r = 0; r < D; ++r) { // world = R lp + t
        float s = b.rot[r] * lp[0];
#pragma unroll
        for (int k = 1; k < D; ++k) s = s + b.rot[k * D + r] * lp[k];
        dpt[r] = (s + b.trans[r]) - point[r];
    }
    return inside;
}

// Can body b colour ANY node of the block whose first node sits at `origin`? Conservative: the block's nodes lie
// within R = (BLOCK - 1) / 2 * h * sqrt(D) of its centre; if the centre is outside the body and farther than
// R + 1.5 h sqrt(D) from its boundary, every node is outside too and farther than 1.5 h sqrt(D) from it, i.e. at
// least one component of its projection vector exceeds the 1.5 h cap of collide(): the body leaves no trace on the
// block and its per-node evaluation can be skipped (most blocks see no body at all).
template <int D>
__host__ __device__ inline bool body_may_touch_block(const BodyDev& b, float cell_width, const float* origin) {
    if (b.shape_type == B200MPM_SHAPE_TRIMESH || b.shape_type == B200MPM_SHAPE_POLYLINE) return false; // collide.wgsl:41
    const float half = 0.5f * (float)(Dim<D>::BLOCK - 1) * cell_width;
    float centre[D], dpt[D];
#pragma unroll
    for (int k = 0; k < D; ++k) centre[k] = origin[k] + half;
    if (project_on_body<D>(b, centre, dpt)) return true;
    float dist2 = 0.0f;
#pragma unroll
    for (int k = 0; k < D; ++k) dist2 += dpt[k] * dpt[k];
    const float root_d = (D == 2) ? 1.41421356f : 1.73205081f;
    const float reach = (half + 1.5f * cell_width) * root_d * 1.001f; // (+ 0.1 % for the rounding of both sides)
    return !(dist2 > reach * reach); // (NaN: keep the body)
}

// collide() (collision/collide.wgsl:23-55): closest collider, distance and affinity/sign bits
// of a grid node at world position `point`. `body_mask`: the bodies to look at (see body_may_touch_block).
template <int D>
__host__ __device__ inline NodeCdf collide(const BodyDev* __restrict__ bodies, uint32_t body_mask, float cell_width,
                                  const float* point) {
    NodeCdf cdf{1.0e10f, 0u, NONE};
    const float dist_cap = cell_width * 1.5f;
    for (uint32_t m = body_mask; m != 0u; m &= m - 1u) {
#ifdef __CUDA_ARCH__
        const uint32_t i = (uint32_t)__ffs((int)m) - 1u;
#else
        const uint32_t i = (uint32_t)__builtin_ffs((int)m) - 1u; // (host build of the culling property test)
#endif
        const BodyDev& b = bodies[i];
        if (b.shape_type == B200MPM_SHAPE_TRIMESH || b.shape_type == B200MPM_SHAPE_POLYLINE) continue; // collide.wgsl:41
        float dpt[D];
        const bool inside = project_on_body<D>(b, point, dpt);
        bool all_le = true;
        float dist2 = 0.0f;
#pragma unroll
        for (int r = 0; r < D; ++r) {
            all_le = all_le && (fabsf(dpt[r]) <= dist_cap);
            dist2 = (r == 0) ? dpt[r] * dpt[r] : dist2 + dpt[r] * dpt[r];
        }
        if (inside || all_le) {
            float dist = sqrtf(dist2);
            cdf.closest_id = (dist < cdf.distance) ? i : cdf.closest_id;
            cdf.distance = fminf(cdf.distance, dist);
            cdf.affinities |= (inside ? 0x00010001u : 0x00000001u) << i;
        }
    }
    return cdf;
}

} // namespace b2
