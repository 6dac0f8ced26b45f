// Rigid-body state kept on the device: world-space mass properties and the per-substep integration. Shared by
// misc.cu (the stand-alone kernels) and sort.cu (k_touch carries the deferred integration of the previous substep).
#pragma once
#include "common.cuh"

namespace b2 {

// ---- update_world_mass_properties (rigid_impulses.wgsl:139-150) -------------------------------------------------
// wgrapier Body::updateMprops: com = pose * local_com, inv_inertia_world = R I^-1 R^T; plus the
// needs_impulse flag. The reference recomputes them at the top of every substep; here whoever CHANGES a pose or a
// velocity (k_integrate_bodies, the host writes, data creation) refreshes them, so no kernel of the substep's
// critical path is spent on <= 16 bodies.
template <int D>
__device__ inline void refresh_body(BodyDev& b) {
    {
        float any = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k) any += fabsf(b.local_inv_mass[k]) + fabsf(b.linvel[k]) + fabsf(b.angvel[k]);
#pragma unroll
        for (int k = 0; k < 9; ++k) any += fabsf(b.local_inv_inertia[k]);
        b.needs_impulse = (any != 0.0f || any != any) ? 1u : 0u;
    }
#pragma unroll
    for (int r = 0; r < D; ++r) {
        float s = b.rot[r] * b.local_com[0];
#pragma unroll
        for (int k = 1; k < D; ++k) s = s + b.rot[k * D + r] * b.local_com[k];
        b.com[r] = s + b.trans[r];
    }
    if (D == 2) {
        b.inv_inertia[0] = b.local_inv_inertia[0];
    } else {
        // W = R * I * R^T (all column-major 3x3)
        float RI[9];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                float s = 0.0f;
#pragma unroll
                for (int k = 0; k < 3; ++k) s += b.rot[k * 3 + r] * b.local_inv_inertia[c * 3 + k];
                RI[c * 3 + r] = s;
            }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                float s = 0.0f;
#pragma unroll
                for (int k = 0; k < 3; ++k) s += RI[k * 3 + r] * b.rot[k * 3 + c]; // R^T[k][c] = R[c][k]
                b.inv_inertia[c * 3 + r] = s;
            }
    }
}
__device__ inline void quat_to_rot(const float* q, float* R) { // column-major
    float i = q[0], j = q[1], k = q[2], w = q[3];
    R[0] = 1.0f - 2.0f * (j * j + k * k);
    R[1] = 2.0f * (i * j + k * w);
    R[2] = 2.0f * (i * k - j * w);
    R[3] = 2.0f * (i * j - k * w);
    R[4] = 1.0f - 2.0f * (i * i + k * k);
    R[5] = 2.0f * (j * k + i * w);
    R[6] = 2.0f * (i * k + j * w);
    R[7] = 2.0f * (j * k - i * w);
    R[8] = 1.0f - 2.0f * (i * i + j * j);
}
__device__ inline void complex_to_rot(const float* c, float* R) {
    R[0] = c[0];
    R[1] = c[1];
    R[2] = -c[1];
    R[3] = c[0];
}

// ---- update (rigid_impulses.wgsl:94-137) ---------------------------------------------------------------
template <int D>
__device__ inline void integrate_body(BodyDev& b, const SimState& sim) {
    const float dt = sim.dt, h = sim.cell_width;
    float il[3] = {0, 0, 0}, ia[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; ++k) { // int2flt (rigid_impulses.wgsl:56-58)
        il[k] = (float)b.imp_lin[k] / 1e5f;
        ia[k] = (float)b.imp_ang[k] / 1e5f;
        b.imp_lin[k] = 0;
        b.imp_ang[k] = 0;
    }
    // Body::applyImpulse: lin += inv_mass (.) imp.lin ; ang += I^-1_world imp.ang
    float lin[3] = {b.linvel[0], b.linvel[1], b.linvel[2]};
    float ang[3] = {b.angvel[0], b.angvel[1], b.angvel[2]};
#pragma unroll
    for (int k = 0; k < D; ++k) lin[k] = lin[k] + b.local_inv_mass[k] * il[k];
    float imp_ang_norm;
    if (D == 2) {
        ang[0] += b.inv_inertia[0] * ia[0];
        imp_ang_norm = fabsf(ia[0]);
    } else {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float s = b.inv_inertia[r] * ia[0];
            s = s + b.inv_inertia[3 + r] * ia[1];
            s = s + b.inv_inertia[6 + r] * ia[2];
            ang[r] = ang[r] + s;
        }
        imp_ang_norm = sqrtf(ia[0] * ia[0] + ia[1] * ia[1] + ia[2] * ia[2]);
    }
    float linvel_norm = 0.0f, imp_lin_norm = 0.0f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        linvel_norm = (k == 0) ? lin[0] * lin[0] : linvel_norm + lin[k] * lin[k];
        imp_lin_norm = (k == 0) ? il[0] * il[0] : imp_lin_norm + il[k] * il[k];
    }
    linvel_norm = sqrtf(linvel_norm);
    imp_lin_norm = sqrtf(imp_lin_norm);
    float angvel_norm = (D == 2) ? fabsf(ang[0]) : sqrtf(ang[0] * ang[0] + ang[1] * ang[1] + ang[2] * ang[2]);
    const float lin_limit = 0.1f * h / dt, ang_limit = 1.0f;
    if (imp_lin_norm != 0.0f || imp_ang_norm != 0.0f) {
        if (linvel_norm > lin_limit) {
            float s = lin_limit / linvel_norm;
#pragma unroll
            for (int k = 0; k < D; ++k) lin[k] *= s;
        }
        if (angvel_norm > ang_limit) {
            float s = ang_limit / angvel_norm;
#pragma unroll
            for (int k = 0; k < 3; ++k) ang[k] *= s;
        }
    }
    // Body::integrateVelocity: rotate about the world COM by exp(ang dt), translate by lin dt.
    float com[3] = {0, 0, 0};
#pragma unroll
    for (int r = 0; r < D; ++r) {
        float s = b.rot[r] * b.local_com[0];
#pragma unroll
        for (int k = 1; k < D; ++k) s = s + b.rot[k * D + r] * b.local_com[k];
        com[r] = s + b.trans[r];
    }
    float dR[9];
    if (D == 2) {
        float a = ang[0] * dt;
        float cs = cosf(a), sn = sinf(a);
        float re = cs * b.rot_raw[0] - sn * b.rot_raw[1], im = sn * b.rot_raw[0] + cs * b.rot_raw[1];
        float n = sqrtf(re * re + im * im);
        b.rot_raw[0] = re / n;
        b.rot_raw[1] = im / n;
        float dc[2] = {cs, sn};
        complex_to_rot(dc, dR);
    } else {
        float ax = ang[0] * dt, ay = ang[1] * dt, az = ang[2] * dt;
        float angle = sqrtf(ax * ax + ay * ay + az * az);
        float dq[4];
        if (angle > 0.0f) {
            float s = sinf(angle * 0.5f) / angle;
            dq[0] = ax * s;
            dq[1] = ay * s;
            dq[2] = az * s;
            dq[3] = cosf(angle * 0.5f);
        } else {
            dq[0] = dq[1] = dq[2] = 0.0f;
            dq[3] = 1.0f;
        }
        float qi = b.rot_raw[0], qj = b.rot_raw[1], qk = b.rot_raw[2], qw = b.rot_raw[3];
        float ni = dq[3] * qi + dq[0] * qw + dq[1] * qk - dq[2] * qj;
        float nj = dq[3] * qj - dq[0] * qk + dq[1] * qw + dq[2] * qi;
        float nk = dq[3] * qk + dq[0] * qj - dq[1] * qi + dq[2] * qw;
        float nw = dq[3] * qw - dq[0] * qi - dq[1] * qj - dq[2] * qk;
        float n = sqrtf(ni * ni + nj * nj + nk * nk + nw * nw);
        b.rot_raw[0] = ni / n;
        b.rot_raw[1] = nj / n;
        b.rot_raw[2] = nk / n;
        b.rot_raw[3] = nw / n;
        quat_to_rot(dq, dR);
    }
    float nt[3] = {0, 0, 0};
    float rel[3] = {b.trans[0] - com[0], b.trans[1] - com[1], b.trans[2] - com[2]};
#pragma unroll
    for (int r = 0; r < D; ++r) {
        float s = dR[r] * rel[0];
#pragma unroll
        for (int k = 1; k < D; ++k) s = s + dR[k * D + r] * rel[k];
        nt[r] = s + lin[r] * dt + com[r];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) b.trans[k] = nt[k];
    if (D == 2) complex_to_rot(b.rot_raw, b.rot);
    else quat_to_rot(b.rot_raw, b.rot);
    // gravity on bodies with non-zero inverse mass (rigid_impulses.wgsl:130-132)
#pragma unroll
    for (int k = 0; k < D; ++k) lin[k] += sim.gravity[k] * ((b.local_inv_mass[k] != 0.0f) ? 1.0f : 0.0f) * dt;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        b.linvel[k] = lin[k];
        b.angvel[k] = ang[k];
    }
    refresh_body<D>(b); // world-space mass properties of the new pose, for the next substep
}

// Runs once per substep - at its end, or (single-GPU graph) deferred into the next substep's k_touch - and is a no-op
// when nothing is pending (the flush at the end of b200mpm_step after an in-place integrate). ONE warp, lane = body.
// The bodies are staged through shared memory (`stage`: room for B200MPM_MAX_BODIES BodyDev): on the global structs
// the update is a chain of ~30 dependent L2 round trips (every store may alias the next load), ~20 us on a busy SM.
template <int D>
__device__ inline void integrate_bodies_warp(const DeviceData& d, uint32_t lane, BodyDev* stage) {
    const bool pending = d.counters->integrate_pending != 0u;
    __syncwarp();
    if (lane == 0) d.counters->integrate_pending = 0u;
    if (!pending) return;
    const SimState sim = *d.sim;
    const uint32_t nb = min(sim.num_bodies, B200MPM_MAX_BODIES);
    constexpr uint32_t WORDS = sizeof(BodyDev) / 4;
    static_assert(sizeof(BodyDev) % 4 == 0, "BodyDev is copied word by word");
    uint32_t* g = reinterpret_cast<uint32_t*>(d.bodies);
    uint32_t* s = reinterpret_cast<uint32_t*>(stage);
#pragma unroll 8
    for (uint32_t i = lane; i < nb * WORDS; i += 32) s[i] = g[i];
    __syncwarp();
    if (lane < nb) integrate_body<D>(stage[lane], sim);
    __syncwarp();
#pragma unroll 8
    for (uint32_t i = lane; i < nb * WORDS; i += 32) g[i] = s[i];
}

} // namespace b2
