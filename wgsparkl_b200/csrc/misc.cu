// Rigid-body kernels ("update rigid particles" and "integrate_bodies" passes), per-substep counter
// bookkeeping, and the readback / host-write helpers behind the C ABI.
#include "launch.h"

#include <cstdlib>
#include "svd.cuh"

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
template <int D>
__global__ void k_refresh_bodies(DeviceData d) {
    if (threadIdx.x < d.sim->num_bodies) refresh_body<D>(d.bodies[threadIdx.x]);
}

// ---- reset_hmap (grid.wgsl:186-203) + clearing of the last sort's per-cell bins -------------------------------------
// Leaves the sparse grid ready for the next k_touch. It does not sit at the top of a substep: nothing after P2G (and
// the halo exchange of sharded runs) reads the hash map or the bins any more - G2P works from its item list - so the
// substep graph runs it on a side branch next to k_g2p, off the critical path (api.cu, finish_substep).
__global__ void __launch_bounds__(256) k_begin_substep(DeviceData d) {
    TL_BEGIN(d, B200MPM_KERNEL_BEGIN);
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = tid; i < d.capacity; i += stride) d.hkeys[i] = NONE;
    // prev_active_blocks was published by the last k_scatter (0 before the first sort: the arrays are
    // zero-initialised at creation)
    const uint32_t prev = min(d.counters->prev_active_blocks, d.capacity);
    const uint32_t nbins = prev * CELLS_PER_BLOCK + 1;
    for (uint32_t i = tid; i < nbins; i += stride) d.cell_start[i] = 0;
    if (tid == 0) d.counters->num_active_blocks = 0;
    TL_END(d, B200MPM_KERNEL_BEGIN);
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
__global__ void k_integrate_bodies(DeviceData d) {
    pdl_start();
    TL_BEGIN(d, B200MPM_KERNEL_INTEGRATE_BODIES);
    const uint32_t id = threadIdx.x;
    // Runs once per substep - at its end, or (single-GPU graph) deferred to the start of the next one, beside the
    // sort - and is a no-op when nothing is pending (the flush at the end of b200mpm_step after an in-place integrate).
    const bool pending = d.counters->integrate_pending != 0u;
    __syncwarp();
    if (id == 0) d.counters->integrate_pending = 0u;
    if (!pending || id >= d.sim->num_bodies) return;
    BodyDev& b = d.bodies[id];
    const float dt = d.sim->dt, h = d.sim->cell_width;
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
    for (int k = 0; k < D; ++k) lin[k] += d.sim->gravity[k] * ((b.local_inv_mass[k] != 0.0f) ? 1.0f : 0.0f) * dt;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        b.linvel[k] = lin[k];
        b.angvel[k] = ang[k];
    }
    refresh_body<D>(b); // world-space mass properties of the new pose, for the next substep
    TL_END(d, B200MPM_KERNEL_INTEGRATE_BODIES);
}

// ---- host writes / reads of body state (src_testbed/step.rs:79-119,175-176) ------------------------------
template <int D>
__global__ void k_write_poses(DeviceData d, const b200mpm_pose* poses, uint32_t n) {
    uint32_t id = threadIdx.x;
    if (id >= n || id >= d.sim->num_bodies) return;
    BodyDev& b = d.bodies[id];
    for (int k = 0; k < 3; ++k) b.trans[k] = poses[id].translation[k];
    for (int k = 0; k < 4; ++k) b.rot_raw[k] = poses[id].rotation[k];
    if (D == 2) complex_to_rot(b.rot_raw, b.rot);
    else quat_to_rot(b.rot_raw, b.rot);
    refresh_body<D>(b);
}
template <int D>
__global__ void k_write_vels(DeviceData d, const b200mpm_velocity* vels, uint32_t n) {
    uint32_t id = threadIdx.x;
    if (id >= n || id >= d.sim->num_bodies) return;
    BodyDev& b = d.bodies[id];
    for (int k = 0; k < 3; ++k) {
        b.linvel[k] = vels[id].linear[k];
        b.angvel[k] = vels[id].angular[k];
    }
    refresh_body<D>(b);
}
__global__ void k_read_poses(DeviceData d, b200mpm_pose* poses, b200mpm_velocity* vels, uint32_t n) {
    uint32_t id = threadIdx.x;
    if (id >= n || id >= d.sim->num_bodies) return;
    const BodyDev& b = d.bodies[id];
    if (poses) {
        for (int k = 0; k < 3; ++k) poses[id].translation[k] = b.trans[k];
        for (int k = 0; k < 4; ++k) poses[id].rotation[k] = b.rot_raw[k];
    }
    if (vels) {
        for (int k = 0; k < 3; ++k) {
            vels[id].linear[k] = b.linvel[k];
            vels[id].angular[k] = b.angvel[k];
        }
    }
}

// ---- particle / grid readback in the caller's layout ---------------------------------------------------------
__global__ void k_gather_positions(DeviceData d, int cur, float4* out, int unordered) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.counters->n_live) {
        // unordered: the spare capacity is marked "no particle", so that a fixed-size copy needs no live count
        if (unordered && i < d.n) out[i] = make_float4(0.f, 0.f, 0.f, __uint_as_float(NONE));
        return;
    }
    float4 p = d.pos4[cur][i];
    uint32_t orig = __float_as_uint(d.vel4[cur][i].w);
    if (unordered) { // device order; w carries the particle id (NONE for an emigrated particle)
        uint32_t id = (__float_as_uint(p.w) & FLAG_DEAD) ? NONE : orig;
        out[i] = make_float4(p.x, p.y, p.z, __uint_as_float(id));
    } else {
        out[orig] = make_float4(p.x, p.y, p.z, 0.0f);
    }
}

template <int D>
__global__ void k_gather_particles(DeviceData d, int cur, b200mpm_particle* out, uint32_t* ids) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.counters->n_live) return;
    float4 p = d.pos4[cur][i];
    float4 v = d.vel4[cur][i];
    uint32_t orig = __float_as_uint(v.w);
    uint32_t mbits = __float_as_uint(p.w);
    const Material m = d.materials[mbits & MAT_ID_MASK];
    b200mpm_particle o;
    memset(&o, 0, sizeof(o));
    o.position[0] = p.x;
    o.position[1] = p.y;
    o.position[2] = (D == 3) ? p.z : 0.0f;
    o.velocity[0] = v.x;
    o.velocity[1] = v.y;
    o.velocity[2] = (D == 3) ? v.z : 0.0f;
    float F[9], C[9];
    float4 fa = d.Fa[cur][i], ca = d.Ca[cur][i];
    F[0] = fa.x, F[1] = fa.y, F[2] = fa.z, F[3] = fa.w;
    C[0] = ca.x, C[1] = ca.y, C[2] = ca.z, C[3] = ca.w;
    if (D == 3) {
        float4 fb = d.Fb[cur][i], cb = d.Cb[cur][i];
        F[4] = fb.x, F[5] = fb.y, F[6] = fb.z, F[7] = fb.w, F[8] = d.Fc[cur][i];
        C[4] = cb.x, C[5] = cb.y, C[6] = cb.z, C[7] = cb.w, C[8] = d.Cc[cur][i];
    }
    for (int k = 0; k < D * D; ++k) {
        o.def_grad[k] = F[k];
        o.affine[k] = C[k];
    }
    if (d.has_bodies) {
        uint32_t aff = d.cdf_aff[cur][i];
        o.cdf_affinity = aff;
        if (aff != 0u) { // invariant: affinity == 0 <=> default cdf (g2p_cdf.wgsl:233-249)
            float4 nd = d.cdf_nd[i], rv = d.cdf_rv[i];
            o.cdf_normal[0] = nd.x, o.cdf_normal[1] = nd.y, o.cdf_normal[2] = (D == 3) ? nd.z : 0.0f;
            o.cdf_signed_distance = nd.w;
            o.cdf_rigid_vel[0] = rv.x, o.cdf_rigid_vel[1] = rv.y, o.cdf_rigid_vel[2] = (D == 3) ? rv.z : 0.0f;
        }
    }
    o.init_volume = m.init_volume;
    o.init_radius = m.init_radius;
    o.mass = m.mass;
    o.lambda = m.lambda;
    o.mu = m.mu;
    o.dp_h0 = m.dp_h0, o.dp_h1 = m.dp_h1, o.dp_h2 = m.dp_h2, o.dp_h3 = m.dp_h3;
    o.dp_lambda = m.dp_lambda, o.dp_mu = m.dp_mu;
    if (d.has_plastic) {
        float4 ps = d.plastic[cur][i];
        o.plastic_det = ps.x, o.plastic_hardening = ps.y, o.plastic_log_vol_gain = ps.z;
    } else {
        o.plastic_det = 1.0f, o.plastic_hardening = 1.0f, o.plastic_log_vol_gain = 0.0f;
    }
    o.phase = (mbits & FLAG_PHASE_BROKEN) ? 0.0f : m.phase;
    o.max_stretch = m.max_stretch;
    o.model = m.model;
    if (ids) { // device order + ids (sharded runs)
        out[i] = o;
        ids[i] = (mbits & FLAG_DEAD) ? NONE : orig;
    } else {
        out[orig] = o;
    }
}

// ---- prep_vertex_buffer (src_testbed/prep_vertex_buffer3d.wgsl:40-95, 2d: 39-94) -----------------------
template <int D>
__global__ void k_prep_vertex_buffer(DeviceData d, int cur, b200mpm_instance* inst, uint32_t mode) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.counters->n_live) return;
    const float4 p = d.pos4[cur][i];
    const float4 v = d.vel4[cur][i];
    b200mpm_instance& o = inst[__float_as_uint(v.w)]; // instances are indexed by the original particle id
    float F[9];
    const float4 fa = d.Fa[cur][i];
    F[0] = fa.x, F[1] = fa.y, F[2] = fa.z, F[3] = fa.w;
    if (D == 3) {
        const float4 fb = d.Fb[cur][i];
        F[4] = fb.x, F[5] = fb.y, F[6] = fb.z, F[7] = fb.w, F[8] = d.Fc[cur][i];
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) o.deformation[4 * c + r] = F[3 * c + r];
        o.position[0] = p.x, o.position[1] = p.y, o.position[2] = p.z;
    } else {
        o.deformation[0] = F[0], o.deformation[1] = F[1], o.deformation[2] = 0.0f;
        o.deformation[4] = F[2], o.deformation[5] = F[3], o.deformation[6] = 0.0f;
        o.deformation[8] = 0.0f, o.deformation[9] = 0.0f, o.deformation[10] = 1.0f;
        o.position[0] = p.x, o.position[1] = p.y, o.position[2] = 0.0f;
    }
    const float base[4] = {o.base_color[0], o.base_color[1], o.base_color[2], o.base_color[3]};
    float col[4] = {base[0], base[1], base[2], base[3]};
    const float h = d.sim->cell_width, dt = d.sim->dt;
    uint32_t aff = 0u;
    float4 nd = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d.has_bodies) {
        aff = d.cdf_aff[cur][i];
        if (aff != 0u) nd = d.cdf_nd[i]; // affinity == 0 <=> default cdf (zero normal, zero distance)
    }
    if (mode == B200MPM_RENDER_VELOCITY) {
        const float vel[3] = {v.x, v.y, v.z};
        for (int k = 0; k < D; ++k) col[k] = fabsf(vel[k]) * dt * 100.0f + 0.2f;
    } else if (mode == B200MPM_RENDER_VOLUME) {
        float U[9], S[3], V[9];
        if (D == 3) svd3<4>(F, U, S, V);
        else svd2(F, U, S, V);
        for (int k = 0; k < D; ++k) col[k] = (1.0f - S[k]) / 0.005f + 0.2f;
    } else if (mode == B200MPM_RENDER_CDF_NORMALS) {
        const float n[3] = {nd.x, nd.y, (D == 3) ? nd.z : 0.0f};
        const bool zero = n[0] == 0.0f && n[1] == 0.0f && n[2] == 0.0f;
        for (int k = 0; k < 3; ++k) col[k] = (zero || k >= D) ? 0.0f : (n[k] + 1.0f) / 2.0f;
    } else if (mode == B200MPM_RENDER_CDF_DISTANCES) {
        const float dd = nd.w / (h * 1.5f);
        col[0] = (dd > 0.0f) ? 0.0f : fabsf(dd);
        col[1] = (dd > 0.0f) ? fabsf(dd) : 0.0f;
        col[2] = 0.0f;
    } else if (mode == B200MPM_RENDER_CDF_SIGNS) {
        const uint32_t a = (aff >> 16) & (aff & 0x0000ffffu);
        col[0] = (aff != 0u && a != 0u) ? 1.0f : 0.0f;
        col[1] = (aff != 0u && a == 0u) ? 1.0f : 0.0f;
        col[2] = 0.0f;
    }
    for (int k = 0; k < 4; ++k) o.color[k] = col[k];
}

__global__ void k_gather_sorted_ids(DeviceData d, int cur, int indirect, uint32_t* out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= d.counters->n_live) return;
    // After a full substep the particle buffers ARE in sorted order; after a sort-only pass the
    // order is given by sorted_ids.
    uint32_t id = indirect ? d.sorted_ids[k] : k;
    out[k] = __float_as_uint(d.vel4[cur][id].w);
}

// Grid in reference form: velocities after the grid update (grid_update.wgsl:20-64) + node cdf.
template <int D>
__global__ void k_gather_grid(DeviceData d, b200mpm_block_info* blocks, b200mpm_node* nodes, uint32_t max_blocks) {
    const uint32_t nb = min(min(d.counters->prev_active_blocks, d.capacity), max_blocks);
    const uint32_t t = threadIdx.x;
    const float dt = d.sim->dt, h = d.sim->cell_width;
    for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
        if (t == 0) {
            int4 vid = d.block_vid[b];
            blocks[b].vid[0] = vid.x, blocks[b].vid[1] = vid.y, blocks[b].vid[2] = vid.z;
            const uint2 range = d.block_range[b];
            blocks[b].first_particle = range.x;
            blocks[b].num_particles = range.y;
        }
        float4 mv = d.node_mv[b * CELLS_PER_BLOCK + t];
        float mass = (D == 3) ? mv.w : mv.z;
        float inv_mass = (mass > 0.0f) ? 1.0f / mass : 0.0f;
        float lim = h / dt;
        float v[3] = {mv.x, mv.y, mv.z};
        b200mpm_node o;
        for (int k = 0; k < 4; ++k) o.momentum_velocity_mass[k] = 0.0f;
        for (int k = 0; k < D; ++k) {
            float vel = (v[k] + mass * d.sim->gravity[k] * dt) * inv_mass;
            o.momentum_velocity_mass[k] = fminf(fmaxf(vel, -lim), lim);
        }
        o.momentum_velocity_mass[D] = mass;
        if (d.has_bodies) {
            uint4 c = d.node_cdf[b * CELLS_PER_BLOCK + t];
            o.cdf_closest_id = c.x;
            o.cdf_distance = __uint_as_float(c.y);
            o.cdf_affinities = c.z;
        } else {
            o.cdf_distance = 1.0e10f; // collide() with no shapes (collide.wgsl:24-25)
            o.cdf_affinities = 0u;
            o.cdf_closest_id = NONE;
        }
        nodes[b * CELLS_PER_BLOCK + t] = o;
    }
}

// ---- launch wrappers ------------------------------------------------------------------------------------------
static inline int div_up(uint64_t a, uint64_t b) { return (int)((a + b - 1) / b); }

bool pdl_enabled() {
    static const bool on = []() {
        const char* e = getenv("B200MPM_PDL");
        return e && e[0] && e[0] != '0';
    }();
    return on;
}

void launch_begin_substep(const LaunchCfg& c, const DeviceData& d) {
    k_begin_substep<<<c.num_sms * 2, 256, 0, c.stream>>>(d);
    ++*c.launch_counter;
}
void launch_refresh_bodies(const LaunchCfg& c, const DeviceData& d) {
    if (c.dim == 2) k_refresh_bodies<2><<<1, 32, 0, c.stream>>>(d);
    else k_refresh_bodies<3><<<1, 32, 0, c.stream>>>(d);
    ++*c.launch_counter;
}
void launch_integrate_bodies(const LaunchCfg& c, const DeviceData& d) {
    if (!d.has_bodies) return;
    if (c.dim == 2) launch_pdl(k_integrate_bodies<2>, 1, 32, 0, c.stream, d);
    else launch_pdl(k_integrate_bodies<3>, 1, 32, 0, c.stream, d);
    ++*c.launch_counter;
}
void launch_gather_positions(const LaunchCfg& c, const DeviceData& d, int cur, float4* out, int unordered) {
    if (d.n == 0) return;
    k_gather_positions<<<div_up(d.n, 256), 256, 0, c.stream>>>(d, cur, out, unordered);
    ++*c.launch_counter;
}
void launch_gather_particles(const LaunchCfg& c, const DeviceData& d, int cur, b200mpm_particle* out, uint32_t* ids) {
    if (d.n == 0) return;
    if (c.dim == 2) k_gather_particles<2><<<div_up(d.n, 128), 128, 0, c.stream>>>(d, cur, out, ids);
    else k_gather_particles<3><<<div_up(d.n, 128), 128, 0, c.stream>>>(d, cur, out, ids);
    ++*c.launch_counter;
}
void launch_gather_grid(const LaunchCfg& c, const DeviceData& d, b200mpm_block_info* blocks, b200mpm_node* nodes,
                        uint32_t max_blocks) {
    if (c.dim == 2) k_gather_grid<2><<<c.num_sms * 8, CELLS_PER_BLOCK, 0, c.stream>>>(d, blocks, nodes, max_blocks);
    else k_gather_grid<3><<<c.num_sms * 8, CELLS_PER_BLOCK, 0, c.stream>>>(d, blocks, nodes, max_blocks);
    ++*c.launch_counter;
}
void launch_prep_vertex_buffer(const LaunchCfg& c, const DeviceData& d, int cur, b200mpm_instance* inst, uint32_t mode) {
    if (d.n == 0) return;
    if (c.dim == 2) k_prep_vertex_buffer<2><<<div_up(d.n, 128), 128, 0, c.stream>>>(d, cur, inst, mode);
    else k_prep_vertex_buffer<3><<<div_up(d.n, 128), 128, 0, c.stream>>>(d, cur, inst, mode);
    ++*c.launch_counter;
}
void launch_gather_sorted_ids(const LaunchCfg& c, const DeviceData& d, int cur, int indirect, uint32_t* out) {
    if (d.n == 0) return;
    k_gather_sorted_ids<<<div_up(d.n, 256), 256, 0, c.stream>>>(d, cur, indirect, out);
    ++*c.launch_counter;
}
void launch_write_poses(const LaunchCfg& c, const DeviceData& d, const b200mpm_pose* poses, uint32_t n) {
    if (c.dim == 2) k_write_poses<2><<<1, 32, 0, c.stream>>>(d, poses, n);
    else k_write_poses<3><<<1, 32, 0, c.stream>>>(d, poses, n);
    ++*c.launch_counter;
}
void launch_write_vels(const LaunchCfg& c, const DeviceData& d, const b200mpm_velocity* vels, uint32_t n) {
    if (c.dim == 2) k_write_vels<2><<<1, 32, 0, c.stream>>>(d, vels, n);
    else k_write_vels<3><<<1, 32, 0, c.stream>>>(d, vels, n);
    ++*c.launch_counter;
}
void launch_read_poses(const LaunchCfg& c, const DeviceData& d, b200mpm_pose* poses, b200mpm_velocity* vels, uint32_t n) {
    k_read_poses<<<1, 32, 0, c.stream>>>(d, poses, vels, n);
    ++*c.launch_counter;
}

} // namespace b2
