// "grid_update" + "g2p" + "particles_update" passes, fused into one kernel.
//
// Reference: src/solver/grid_update.wgsl:20-64 (momentum -> velocity, gravity, clamp),
// src/solver/g2p.wgsl:44-238 (velocity + velocity-gradient gather, CPIC ghost velocities,
// rigid_vel), src/solver/particle_update.wgsl:45-141 (advection, penalty, F update, constitutive
// models, new APIC affine). In the reference these are three dispatches that round-trip the
// 176-byte AoS `Dynamics` struct through memory (the velocity gradient is parked in `affine`,
// SURVEY A.4).
//
// B200 design: one CTA per work item = up to 512 particles of one block (k_scatter's list:
// blocks without particles never appear, collider-side and densely populated blocks come first).
//   * the (BLOCK+2)^D node tile is staged in shared memory through the neighbour table; the grid
//     update is applied while staging, so node velocities never exist in HBM;
//   * one thread per particle of the block's contiguous sorted range: 16-byte vector loads of the
//     SoA particle arrays, the stencil gather from shared memory, the constitutive update with a
//     single SVD, and 16-byte vector stores to the OTHER ping-pong buffer at the particle's sorted
//     slot — the physical reordering of the particle arrays costs no extra pass.
#include "launch.h"
#include "models.cuh"

namespace b2 {

constexpr int G2P_THREADS = 128;
// CTAs per SM: 6 x 128 threads at 80 registers (elastic), 5 at 96 (plastic: the SVD + return mapping need the
// room). Measured alternatives (tools/build_variant.py): 5 / 7 elastic and 4 / 6 plastic are all slower or equal,
// 64- and 32-thread CTAs are 20 % / 60 % slower.
constexpr int G2P_MIN_CTAS_ELASTIC = 6, G2P_MIN_CTAS_PLASTIC = 5;

template <int D, bool PLASTIC, bool CPIC>
__global__ void __launch_bounds__(G2P_THREADS, PLASTIC ? G2P_MIN_CTAS_PLASTIC : G2P_MIN_CTAS_ELASTIC) k_g2p(DeviceData d, int cur) {
    constexpr int B = Dim<D>::BLOCK, T = Dim<D>::TILE, TC = Dim<D>::TILE_CELLS;
    constexpr int NA = Dim<D>::NASSOC;
    __shared__ float4 tile_v[TC];
    __shared__ uint2 tile_c[CPIC ? TC : 1];
    __shared__ uint32_t s_nbr[NA];
    __shared__ uint32_t s_next;

    const int t = threadIdx.x;
    const int nxt = cur ^ 1;
    const uint32_t nb = min(d.counters->num_active_blocks, d.capacity);
    const float h = d.sim->cell_width;
    const float dt = d.sim->dt;
    const float inv_h = 1.0f / h;
    const float inv_d = 4.0f / (h * h); // kernel.wgsl:57-59
    const float grav[3] = {d.sim->gravity[0], d.sim->gravity[1], d.sim->gravity[2]};
    const float vel_limit = h / dt;

    const float4* __restrict__ pos4 = d.pos4[cur];
    const float4* __restrict__ vel4 = d.vel4[cur];
    const float4* __restrict__ Fa = d.Fa[cur];
    const float4* __restrict__ Fb = d.Fb[cur];
    const float* __restrict__ Fc = d.Fc[cur];

    // (block, part) items of <= G2P_ITEM particles from k_scatter: collider-side blocks first (front of the list)
    const uint32_t nfront = d.counters->num_g2p_items, nitems = nfront + d.counters->num_g2p_back;
    while (true) {
        __syncthreads();
        if (t == 0) {
            const uint32_t w = atomicAdd(&d.counters->work_g2p, 1u);
            s_next = (w < nitems) ? d.g2p_list[w < nfront ? w : d.g2p_list_len - 1u - (w - nfront)] : NONE;
        }
        __syncthreads();
        const uint32_t item = s_next;
        if (item == NONE) break;
        const uint32_t b = item & 0xffffffu;
        uint32_t first = d.cell_start[b * CELLS_PER_BLOCK];
        uint32_t last = d.cell_start[(b + 1) * CELLS_PER_BLOCK];
        {
            const uint32_t parts = g2p_parts(last - first), part = item >> 24;
            const uint32_t per = (last - first + parts - 1) / parts;
            first = min(first + part * per, last);
            last = min(first + per, last);
        }
        if (t < NA) s_nbr[t] = d.nbr[b * NA + t];
        __syncthreads();
        // Stage the tile; grid_update (grid_update.wgsl:45-64) on the fly.
        const int any_cdf = CPIC ? (int)d.block_flags[b] : 0; // the tile holds a collider (k_scatter)
        for (int n = t; n < TC; n += G2P_THREADS) { // g2p.wgsl:72-132
            int x = n % T, y = (n / T) % T, z = n / (T * T);
            int ox = x >= B, oy = y >= B, oz = z >= B;
            uint32_t hn = s_nbr[ox + 2 * oy + 4 * oz];
            float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
            uint2 cdf = make_uint2(0u, NONE);
            if (hn != NONE) {
                uint32_t node = hn * CELLS_PER_BLOCK + (x - ox * B) + (y - oy * B) * B + (z - oz * B) * B * B;
                float4 mv = d.node_mv[node];
                float mass = (D == 3) ? mv.w : mv.z;
                float inv_mass = (mass > 0.0f) ? 1.0f / mass : 0.0f;
                float vx = (mv.x + mass * grav[0] * dt) * inv_mass;
                float vy = (mv.y + mass * grav[1] * dt) * inv_mass;
                out.x = fminf(fmaxf(vx, -vel_limit), vel_limit);
                out.y = fminf(fmaxf(vy, -vel_limit), vel_limit);
                if (D == 3) {
                    float vz = (mv.z + mass * grav[2] * dt) * inv_mass;
                    out.z = fminf(fmaxf(vz, -vel_limit), vel_limit);
                }
                out.w = mass;
                if (CPIC && any_cdf) {
                    uint4 g = d.node_cdf[node];
                    cdf = make_uint2(g.z, g.x); // (affinities, closest_id)
                }
            }
            tile_v[n] = out;
            if (CPIC && any_cdf) tile_c[n] = cdf;
        }
        __syncthreads();

        for (uint32_t k = first + t; k < last; k += G2P_THREADS) {
            const uint32_t id = __ldg(d.sorted_ids + k);
            const float4 p4 = __ldg(pos4 + id);
            const float4 v4 = __ldg(vel4 + id);
            float F[D * D];
            {
                float4 fa = __ldg(Fa + id);
                F[0] = fa.x, F[1] = fa.y, F[2] = fa.z, F[3] = fa.w;
                if (D == 3) {
                    float4 fb = __ldg(Fb + id);
                    F[4] = fb.x, F[5] = fb.y, F[6] = fb.z, F[7] = fb.w;
                    F[D * D - 1] = __ldg(Fc + id);
                }
            }
            uint32_t mbits = __float_as_uint(p4.w);
            const Material m = d.materials[mbits & MAT_ID_MASK];
            const float pp[3] = {p4.x, p4.y, p4.z};
            float d0[D], w[D][3];
            int tb = 0;
#pragma unroll
            for (int a = 0; a < D; ++a) {
                float cf = round_div(pp[a], h, inv_h) - 1.0f; // particle3d.wgsl:41-57
                int l = ((int)cf) & (B - 1);
                tb += l * ((a == 0) ? 1 : (a == 1) ? T : T * T);
                d0[a] = cf * h - pp[a];
                bspline(-d0[a] * inv_h, w[a][0], w[a][1], w[a][2]);
            }
            uint32_t pa = 0u;
            V3 normal = v3(0, 0, 0);
            float sd = 0.0f;
            if (CPIC) {
                if (any_cdf) pa = d.cdf_aff[nxt][k];
                if (pa != 0u) {
                    float4 nd = d.cdf_nd[k];
                    normal = v3(nd.x, nd.y, (D == 3) ? nd.z : 0.0f);
                    sd = nd.w;
                }
            }
            const V3 pvel = v3(v4.x, v4.y, (D == 3) ? v4.z : 0.0f);
            const V3 ppos = v3(pp[0], pp[1], (D == 3) ? pp[2] : 0.0f);

            // ---- G2P gather (g2p.wgsl:170-218): v = sum w v_n, grad v = sum (w inv_d) v_n (x) dpt,
            //      dpt = d0 + s h. Accumulated as v, and the first moments of w v_n over s.
            float vs[D], mom[D][D]; // mom[c][r] = sum w v_n[r] s_c
#pragma unroll
            for (int r = 0; r < D; ++r) {
                vs[r] = 0.0f;
#pragma unroll
                for (int c = 0; c < D; ++c) mom[c][r] = 0.0f;
            }
            if (!CPIC || pa == 0u) {
                // Fast path (a zero affinity word is compatible with every node): fully unrolled,
                // separable accumulation — ~4.3 FMA per node and component.
#pragma unroll
                for (int sz = 0; sz < (D == 3 ? 3 : 1); ++sz) {
                    float r0[D], rx[D], ry[D];
#pragma unroll
                    for (int r = 0; r < D; ++r) r0[r] = rx[r] = ry[r] = 0.0f;
#pragma unroll
                    for (int sy = 0; sy < 3; ++sy) {
                        float t0[D], t1[D];
#pragma unroll
                        for (int r = 0; r < D; ++r) t0[r] = t1[r] = 0.0f;
#pragma unroll
                        for (int sx = 0; sx < 3; ++sx) {
                            const float4 cell = tile_v[tb + sx + T * sy + T * T * sz];
                            const float cv[3] = {cell.x, cell.y, cell.z};
                            const float wx = w[0][sx], sxw = (float)sx * w[0][sx];
#pragma unroll
                            for (int r = 0; r < D; ++r) {
                                t0[r] = fmaf(wx, cv[r], t0[r]);
                                if (sx > 0) t1[r] = fmaf(sxw, cv[r], t1[r]);
                            }
                        }
                        const float wy = w[1][sy];
#pragma unroll
                        for (int r = 0; r < D; ++r) {
                            r0[r] += wy * t0[r];
                            rx[r] += wy * t1[r];
                            if (sy > 0) ry[r] += ((float)sy * wy) * t0[r];
                        }
                    }
                    const float wz = (D == 3) ? w[D - 1][sz] : 1.0f;
#pragma unroll
                    for (int r = 0; r < D; ++r) {
                        vs[r] += wz * r0[r];
                        mom[0][r] += wz * rx[r];
                        mom[1][r] += wz * ry[r];
                        if (D == 3 && sz > 0) mom[D - 1][r] += ((float)sz * wz) * r0[r];
                    }
                }
            } else {
                // CPIC path (particles next to a collider only; g2p.wgsl:186-207): incompatible nodes
                // contribute the particle's ghost velocity. Rolled over the (sy, sz) rows, unrolled along x: the
                // fully unrolled form triples the kernel's instruction-cache footprint, the fully rolled one spends
                // more on index arithmetic and weight selection than on the gather itself.
                // Against a body that neither moves nor can be moved the ghost velocity does not depend on the
                // node: v_body(x) = 0, ghost = project_velocity(v_p, n_p).
                const V3 ghost_static = project_velocity(pvel, normal);
#pragma unroll 1
                for (int yz = 0; yz < (D == 3 ? 9 : 3); ++yz) {
                    const int sy = yz % 3, sz = yz / 3;
                    const float wys = (sy == 0) ? w[1][0] : (sy == 1) ? w[1][1] : w[1][2];
                    const float wzs = (D == 3) ? ((sz == 0) ? w[D - 1][0] : (sz == 1) ? w[D - 1][1] : w[D - 1][2]) : 1.0f;
                    const float wyz = wys * wzs;
                    const int row = tb + T * sy + T * T * sz;
                    float t0[D], t1[D];
#pragma unroll
                    for (int r = 0; r < D; ++r) t0[r] = t1[r] = 0.0f;
#pragma unroll
                    for (int sx = 0; sx < 3; ++sx) {
                        const float4 cell = tile_v[row + sx];
                        const uint2 nc = tile_c[row + sx];
                        float cv[3] = {cell.x, cell.y, cell.z};
                        if (!affinities_are_compatible(pa, nc.x)) {
                            V3 ghost = pvel;
                            if (nc.y != NONE) {
                                const BodyDev& body = d.bodies[nc.y];
                                ghost = ghost_static;
                                if (body.needs_impulse) {
                                    V3 center = v3(d0[0] + (float)sx * h, d0[1] + (float)sy * h,
                                                   (D == 3) ? d0[D - 1] + (float)sz * h : 0.0f) + ppos;
                                    V3 bpv = velocity_at_point<D>(body, center);
                                    ghost = bpv + project_velocity(pvel - bpv, normal);
                                }
                            }
                            cv[0] = ghost.x, cv[1] = ghost.y, cv[2] = ghost.z;
                        }
                        const float wx = w[0][sx], sxw = (float)sx * w[0][sx];
#pragma unroll
                        for (int r = 0; r < D; ++r) {
                            t0[r] = fmaf(wx, cv[r], t0[r]);
                            if (sx > 0) t1[r] = fmaf(sxw, cv[r], t1[r]);
                        }
                    }
                    const float syw = (float)sy * wyz, szw = (float)sz * wyz;
#pragma unroll
                    for (int r = 0; r < D; ++r) {
                        vs[r] = fmaf(wyz, t0[r], vs[r]);
                        mom[0][r] = fmaf(wyz, t1[r], mom[0][r]);
                        mom[1][r] = fmaf(syw, t0[r], mom[1][r]);
                        if (D == 3) mom[D - 1][r] = fmaf(szw, t0[r], mom[D - 1][r]);
                    }
                }
            }
            float G[D * D]; // velocity gradient, column-major: G[c*D + r] = sum (w inv_d) v[r] dpt[c]
#pragma unroll
            for (int c = 0; c < D; ++c)
#pragma unroll
                for (int r = 0; r < D; ++r) G[c * D + r] = inv_d * (vs[r] * d0[c] + h * mom[c][r]);

            // rigid_vel (g2p.wgsl:220-227)
            V3 rigid_vel = v3(0, 0, 0);
            if (CPIC) {
                uint32_t bits = pa & 0xffffu;
                while (bits) {
                    int i = __ffs(bits) - 1;
                    bits &= bits - 1;
                    rigid_vel = rigid_vel + velocity_at_point<D>(d.bodies[i], ppos);
                }
            }

            // ---- particle update (particle_update.wgsl:60-137)
            V3 vel = v3(vs[0], vs[1], (D == 3) ? vs[D - 1] : 0.0f);
            const bool penetrating = CPIC && (sd < -0.05f * h);
            if (penetrating) vel = rigid_vel + project_velocity(vel - rigid_vel, normal);
            {
                float len = length(vel);
                if (len > vel_limit) vel = vel * (1.0f / len) * h * (1.0f / dt);
            }
            const V3 new_pos = ppos + vel * dt;
            if (penetrating) {
                float corrected = fmaxf(sd, -0.3f * h);
                vel = vel + normal * (dt * -corrected * 1.0e3f);
            }
            float Fn[D * D]; // F + (G dt) F
#pragma unroll
            for (int c = 0; c < D; ++c)
#pragma unroll
                for (int r = 0; r < D; ++r) {
                    float s = 0.0f;
#pragma unroll
                    for (int kk = 0; kk < D; ++kk) s += (G[kk * D + r] * dt) * F[c * D + kk];
                    Fn[c * D + r] = F[c * D + r] + s;
                }
            float4 plastic = make_float4(1.0f, 1.0f, 0.0f, 0.0f);
            if (PLASTIC) plastic = d.plastic[cur][id];
            float tau[D * D];
            uint32_t flags = mbits;
            constitutive_update<D, PLASTIC>(m, flags, Fn, plastic, tau);
            const float sc = m.init_volume * inv_d * dt;
            float Cn[D * D]; // affine = grad_v * mass - stress * (V0 inv_d dt)
#pragma unroll
            for (int i = 0; i < D * D; ++i) Cn[i] = G[i] * m.mass - tau[i] * sc;

            // ---- write to the other buffer at the sorted slot
            d.pos4[nxt][k] = make_float4(new_pos.x, new_pos.y, new_pos.z, __uint_as_float(flags));
            d.vel4[nxt][k] = make_float4(vel.x, vel.y, vel.z, v4.w);
            d.Fa[nxt][k] = make_float4(Fn[0], Fn[1], Fn[2], Fn[3]);
            d.Ca[nxt][k] = make_float4(Cn[0], Cn[1], Cn[2], Cn[3]);
            if (D == 3) {
                d.Fb[nxt][k] = make_float4(Fn[4], Fn[5], Fn[6], Fn[7]);
                d.Fc[nxt][k] = Fn[D * D - 1];
                d.Cb[nxt][k] = make_float4(Cn[4], Cn[5], Cn[6], Cn[7]);
                d.Cc[nxt][k] = Cn[D * D - 1];
            }
            if (PLASTIC) d.plastic[nxt][k] = plastic;
            if (CPIC) {
                if (pa != 0u) d.cdf_rv[k] = make_float4(rigid_vel.x, rigid_vel.y, rigid_vel.z, 0.0f);
            }
        }
    }

    // Particles of dropped blocks (capacity overflow only): carried over unchanged.
    const uint32_t dropped = d.counters->dropped_particles;
    if (dropped) {
        const uint32_t total = d.cell_start[nb * CELLS_PER_BLOCK];
        for (uint32_t k = total + blockIdx.x * blockDim.x + t; k < total + dropped && k < d.counters->n_live; k += gridDim.x * blockDim.x) {
            const uint32_t id = d.sorted_ids[k];
            d.pos4[nxt][k] = d.pos4[cur][id];
            d.vel4[nxt][k] = d.vel4[cur][id];
            d.Fa[nxt][k] = d.Fa[cur][id];
            d.Ca[nxt][k] = d.Ca[cur][id];
            if (D == 3) {
                d.Fb[nxt][k] = d.Fb[cur][id];
                d.Fc[nxt][k] = d.Fc[cur][id];
                d.Cb[nxt][k] = d.Cb[cur][id];
                d.Cc[nxt][k] = d.Cc[cur][id];
            }
            if (PLASTIC) d.plastic[nxt][k] = d.plastic[cur][id];
            if (CPIC) d.cdf_aff[nxt][k] = d.cdf_aff[cur][id];
        }
    }
}

template <int D>
static void launch_g2p_dim(const LaunchCfg& c, const DeviceData& d, int cur) {
    const int grid = c.num_sms * 8;
    if (d.has_plastic) {
        if (d.has_bodies) k_g2p<D, true, true><<<grid, G2P_THREADS, 0, c.stream>>>(d, cur);
        else k_g2p<D, true, false><<<grid, G2P_THREADS, 0, c.stream>>>(d, cur);
    } else {
        if (d.has_bodies) k_g2p<D, false, true><<<grid, G2P_THREADS, 0, c.stream>>>(d, cur);
        else k_g2p<D, false, false><<<grid, G2P_THREADS, 0, c.stream>>>(d, cur);
    }
}

void launch_g2p_update(const LaunchCfg& c, const DeviceData& d, int cur) {
    if (d.n == 0) return;
    if (c.dim == 2) launch_g2p_dim<2>(c, d, cur);
    else launch_g2p_dim<3>(c, d, cur);
    ++*c.launch_counter;
}

} // namespace b2
