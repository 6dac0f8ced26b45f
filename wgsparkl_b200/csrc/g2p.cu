// "grid_update" + "g2p" + "particles_update" passes, fused into one kernel.
//
// Reference: src/solver/grid_update.wgsl:20-64 (momentum -> velocity, gravity, clamp),
// src/solver/g2p.wgsl:44-238 (velocity + velocity-gradient gather, CPIC ghost velocities,
// rigid_vel), src/solver/particle_update.wgsl:45-141 (advection, penalty, F update, constitutive
// models, new APIC affine). In the reference these are three dispatches that round-trip the
// 176-byte AoS `Dynamics` struct through memory (the velocity gradient is parked in `affine`,
// SURVEY A.4).
//
// B200 design: persistent CTAs of 128 threads; a work item (k_scatter's list) is <= 128
// consecutive sorted slots of one block, one particle per thread, collider-side blocks first. The kernel is a
// software pipeline in which no global load is ever consumed in the iteration that issued it:
//   item i+3: four lanes request the 64-byte item descriptor (slot range, flags, neighbour table)   [cp.async]
//   item i+2: every thread requests the id of its particle (sorted_ids, coalesced)                  [cp.async]
//   item i+1: every thread requests its particle's 16-byte SoA records into its private shared-memory slot and,
//             when the block changes, two nodes of the (BLOCK+2)^D tile through the item's neighbour table
//             (two tile slots)                                                                       [cp.async]
//   item i  : cp.async.wait_all, the grid update (grid_update.wgsl:45-64) applied in place to the nodes the thread
//             requested itself, ONE barrier, then the stencil gather from shared memory, the constitutive update
//             with a single decomposition, and 16-byte vector stores to the OTHER ping-pong buffer at the
//             particle's sorted slot - the physical reordering of the particle arrays costs no extra pass.
// Work distribution is STATIC: CTA c takes the runs {c, c + P, c + 2P, ...} of G2P_RUN consecutive items. A deep
// prefetch pipeline and dynamic claims do not mix at this granularity (a million particles are ~11 items per
// CTA; the items a CTA would have to claim ahead of their use are a third of its work, handed out blindly -
// measured: 66 -> 85 us when the claims grow). Strided runs give every CTA the same number of items (+-1), deal
// the slow collider-side items at the front of the list out evenly, and keep consecutive parts of a block on
// one tile.
// Registers decide this kernel (measured with the compute phase repeated 1-3x per item): 96 registers x 16 warps
// beat 72 x 20 and 64 x 24 by 30-50 % because the spills land in the middle of the stencil gather - so the CTA
// count follows from the register count, not the other way round, and no warp is set aside as a producer.
#include "launch.h"
#include "models.cuh"

#include <cstdio>
#include <cstdlib>
#include <utility>

namespace b2 {

constexpr int G2P_THREADS = 128;
#ifndef G2P_CTAS_ELASTIC
#define G2P_CTAS_ELASTIC 5
#endif
#ifndef G2P_CTAS_PLASTIC
#define G2P_CTAS_PLASTIC 4
#endif
#ifndef G2P_RUN
#define G2P_RUN 2u
#endif
constexpr int G2P_DQ = 8, G2P_IQ = 4; // look-ahead rings: descriptors 3 items ahead, ids 2 items ahead
constexpr int G2P_SMEM_MATS = 16; // material-table entries mirrored in shared memory

template <int D, bool PLASTIC, bool CPIC>
struct __align__(16) G2PShared {
    static constexpr int TC = Dim<D>::TILE_CELLS;
    float4 tile_v[2][TC]; // node momentum + mass as they land; velocity + mass after the in-place grid update
    uint2 tile_c[CPIC ? 2 : 1][CPIC ? TC : 1]; // (affinities, closest_id) of NodeCdf
    float4 pos[2][G2P_THREADS], vel[2][G2P_THREADS], Fa[2][G2P_THREADS];
    float4 Fb[2][D == 3 ? G2P_THREADS : 1];
    float4 plastic[PLASTIC ? 2 : 1][PLASTIC ? G2P_THREADS : 1];
    float4 nd[CPIC ? 2 : 1][CPIC ? G2P_THREADS : 1];
    float Fc[2][D == 3 ? G2P_THREADS : 1];
    uint32_t aff[CPIC ? 2 : 1][CPIC ? G2P_THREADS : 1];
    G2PItem dq[G2P_THREADS / 32][G2P_DQ]; // per warp: item i's descriptor in slot i % G2P_DQ (every warp keeps its own
                                          // copy, so that descriptors never need a CTA barrier)
    uint32_t ids[G2P_IQ][G2P_THREADS]; // item i's particle ids in row i % G2P_IQ (each thread reads its own)
    Material mats[G2P_SMEM_MATS];
    float h, dt, grav[3], inv_h, inv_d, vel_limit;
    int slab_lo, slab_hi;
};

template <int D, bool PLASTIC, bool CPIC, int CTAS, bool SHARDED>
__global__ void __launch_bounds__(G2P_THREADS, CTAS) k_g2p(DeviceData d) {
    // The launch wrapper hands over the ping-pong arrays already swapped: index 0 = current, 1 = next, so that every
    // array base is a constant-bank operand instead of a register pair.
    constexpr int cur = 0, nxt = 1;
    constexpr int B = Dim<D>::BLOCK, T = Dim<D>::TILE, TC = Dim<D>::TILE_CELLS;
    extern __shared__ __align__(16) unsigned char g2p_smem[];
    G2PShared<D, PLASTIC, CPIC>& sm = *reinterpret_cast<G2PShared<D, PLASTIC, CPIC>*>(g2p_smem);

    // (pdl_wait() comes later: the prologue below only reads what k_scatter and the previous substep left)
    const int t = threadIdx.x;
    const bool mats_in_smem = d.num_materials <= (uint32_t)G2P_SMEM_MATS;
    if (t == 0) { // simulation constants live in shared memory, not in 8 registers per thread
        sm.h = d.sim->cell_width;
        sm.dt = d.sim->dt;
        sm.grav[0] = d.sim->gravity[0], sm.grav[1] = d.sim->gravity[1], sm.grav[2] = d.sim->gravity[2];
        sm.inv_h = 1.0f / sm.h;
        sm.inv_d = 4.0f / (sm.h * sm.h); // kernel.wgsl:57-59
        sm.vel_limit = sm.h / sm.dt;
        sm.slab_lo = d.sim->slab_lo, sm.slab_hi = d.sim->slab_hi;
    }
    if (mats_in_smem && t < (int)(d.num_materials * 4u)) ((float4*)sm.mats)[t] = ((const float4*)d.materials)[t];

    // Logical item w lives at the front of the list for w < nfront (collider-side blocks) and at the back, counted
    // from the end, otherwise.
    const uint32_t nfront = d.counters->num_g2p_items, nitems = nfront + d.counters->num_g2p_back;
    // Requests the descriptor of this CTA's i-th item into the warp's ring (an END descriptor after the last item).
    const int lane = t & 31;
    G2PItem* const dq = sm.dq[t >> 5];
    auto request_desc = [&](uint32_t i) {
        if (lane >= 16) return;
        const uint32_t idx = ((i / G2P_RUN) * gridDim.x + blockIdx.x) * G2P_RUN + (i % G2P_RUN);
        uint32_t* slot = (uint32_t*)&dq[i % G2P_DQ];
        if (idx < nitems) {
            const uint32_t phys = idx < nfront ? idx : d.g2p_items_len - 1u - (idx - nfront);
            if (lane < 4) cp_async16(slot + 4 * lane, (const uint32_t*)(d.g2p_items + phys) + 4 * lane);
        } else {
            slot[lane] = (lane == 0) ? NONE : 0u;
        }
    };
    // Requests the id of this thread's particle of item i (its descriptor has landed).
    auto request_id = [&](uint32_t i) {
        const G2PItem& it = dq[i % G2P_DQ];
        if ((uint32_t)t < it.count) cp_async4(&sm.ids[i % G2P_IQ][t], d.sorted_ids + it.first + t);
    };
    // Item j brings a new tile iff its block differs from item j-1's; the tile slot alternates with every new tile.
    auto new_tile = [&](uint32_t j) -> bool {
        return j == 0u || dq[j % G2P_DQ].block != dq[(j - 1u) % G2P_DQ].block;
    };
    int ts = 1; // slot of the current item's tile
    // Requests item j's records of this thread's particle (its id has landed) ...
    auto request_records = [&](uint32_t j) {
        const G2PItem& it = dq[j % G2P_DQ];
        if (it.block == NONE) return;
        const int s = (int)(j & 1u);
        if ((uint32_t)t < it.count) {
            const uint32_t id = sm.ids[j % G2P_IQ][t];
            cp_async16(&sm.pos[s][t], d.pos4[cur] + id);
            cp_async16(&sm.vel[s][t], d.vel4[cur] + id);
            cp_async16(&sm.Fa[s][t], d.Fa[cur] + id);
            if (D == 3) {
                cp_async16(&sm.Fb[s][t], d.Fb[cur] + id);
                cp_async4(&sm.Fc[s][t], d.Fc[cur] + id);
            }
            if (PLASTIC) cp_async16(&sm.plastic[PLASTIC ? s : 0][PLASTIC ? t : 0], d.plastic[cur] + id);
        }
    };
    // ... and what THIS substep's P2G produced: the particle's colour (collider-side blocks) and, on a block change,
    // this thread's two nodes of the tile.
    auto request_grid = [&](uint32_t j) {
        const G2PItem& it = dq[j % G2P_DQ];
        if (it.block == NONE) return;
        const int s = (int)(j & 1u);
        const bool any_cdf = CPIC && (it.flags & 1u);
        if (any_cdf && (uint32_t)t < it.count) { // by sorted slot (k_scatter's default / k_p2g's colouring)
            cp_async4(&sm.aff[CPIC ? s : 0][CPIC ? t : 0], d.cdf_aff[nxt] + it.first + t);
            cp_async16(&sm.nd[CPIC ? s : 0][CPIC ? t : 0], d.cdf_nd + it.first + t);
        }
        if (new_tile(j)) { // g2p.wgsl:72-132, through the item's neighbour table
            const int u = ts ^ 1;
            // (an opaque copy of t: the index arithmetic below is loop-invariant, and hoisted out of the main loop
            // it would sit in registers - or in spill slots - across the whole compute phase)
            int tt = t;
            asm volatile("" : "+r"(tt));
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int n = tt + G2P_THREADS * k;
                if (n < TC) {
                    const int x = n % T, y = (n / T) % T, z = n / (T * T);
                    const int ox = x >= B, oy = y >= B, oz = z >= B;
                    const uint32_t hn = it.nbr[ox + 2 * oy + 4 * oz];
                    if (hn != NONE) {
                        const uint32_t node = hn * CELLS_PER_BLOCK + (x - ox * B) + (y - oy * B) * B + (z - oz * B) * B * B;
                        cp_async16(&sm.tile_v[u][n], d.node_mv + node);
                        if (any_cdf) {
                            const uint32_t* g = (const uint32_t*)(d.node_cdf + node);
                            cp_async4(&sm.tile_c[CPIC ? u : 0][CPIC ? n : 0].x, g + 2); // affinities
                            cp_async4(&sm.tile_c[CPIC ? u : 0][CPIC ? n : 0].y, g); // closest_id
                        }
                    } else {
                        sm.tile_v[u][n] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (any_cdf) sm.tile_c[CPIC ? u : 0][CPIC ? n : 0] = make_uint2(0u, NONE);
                    }
                }
            }
        }
    };

    // ---- prologue: descriptors 0..2, ids 0..1, records of item 0 (three exposed latencies, once per CTA)
    request_desc(0);
    request_desc(1);
    request_desc(2);
    cp_async_wait_all();
    __syncthreads(); // (also: the constants and the material table)
    request_id(0);
    request_id(1);
    cp_async_wait_all();
    request_records(0);
    pdl_wait(); // k_p2g is complete: node momenta and particle colours may be read
    pdl_trigger();
    TL_BEGIN(d, B200MPM_KERNEL_G2P);
    request_grid(0);

    for (uint32_t i = 0;; ++i) {
        const bool cur_new = dq[i % G2P_DQ].block != NONE && new_tile(i);
        if (cur_new) ts ^= 1;
        cp_async_wait_all(); // everything requested one iteration ago has landed (for this thread)
        if (cur_new) { // grid_update (grid_update.wgsl:45-64), in place, each thread on the nodes it requested
            const float dt = sm.dt, vel_limit = sm.vel_limit;
            const float grav[3] = {sm.grav[0], sm.grav[1], sm.grav[2]};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int n = t + G2P_THREADS * k;
                if (n < TC) {
                    const float4 mv = sm.tile_v[ts][n];
                    const float mass = (D == 3) ? mv.w : mv.z;
                    const float inv_mass = (mass > 0.0f) ? 1.0f / mass : 0.0f;
                    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float vx = (mv.x + mass * grav[0] * dt) * inv_mass;
                    const float vy = (mv.y + mass * grav[1] * dt) * inv_mass;
                    out.x = fminf(fmaxf(vx, -vel_limit), vel_limit);
                    out.y = fminf(fmaxf(vy, -vel_limit), vel_limit);
                    if (D == 3) {
                        const float vz = (mv.z + mass * grav[2] * dt) * inv_mass;
                        out.z = fminf(fmaxf(vz, -vel_limit), vel_limit);
                    }
                    out.w = mass;
                    sm.tile_v[ts][n] = out;
                }
            }
        }
        // A CTA barrier only where the warps share something: a freshly updated tile. (It also keeps the tile slots
        // safe: nobody requests the tile after next before everybody is done with the previous one.) Records and
        // ids are thread-private, descriptors warp-private: between block changes the warps drift freely.
        if (cur_new) __syncthreads();
        else __syncwarp();
        const G2PItem& item = dq[i % G2P_DQ];
        if (item.block == NONE) break;
        request_records(i + 1);
        request_grid(i + 1);
        request_id(i + 2);
        request_desc(i + 3);

        const uint32_t count = item.count;
        const int any_cdf = CPIC ? (int)(item.flags & 1u) : 0; // the tile holds a collider (k_scatter)
#ifdef G2P_REPEAT /* diagnostic: every item is computed and stored G2P_REPEAT times (time difference = pure compute) */
#pragma unroll 1
        for (int rep = 0; rep < G2P_REPEAT; ++rep) {
        asm volatile("" ::: "memory");
#endif
        if ((uint32_t)t < count) {
            const float h = sm.h, dt = sm.dt, inv_h = sm.inv_h, inv_d = sm.inv_d, vel_limit = sm.vel_limit;
            const int s = (int)(i & 1u);
            const uint32_t k = item.first + t;
            const float4* __restrict__ tile_v = sm.tile_v[ts];
            const uint2* __restrict__ tile_c = sm.tile_c[CPIC ? ts : 0];
            const float4 p4 = sm.pos[s][t];
            const float4 v4 = sm.vel[s][t];
            float F[D * D];
            {
                const float4 fa = sm.Fa[s][t];
                F[0] = fa.x, F[1] = fa.y, F[2] = fa.z, F[3] = fa.w;
                if (D == 3) {
                    const float4 fb = sm.Fb[s][t];
                    F[4] = fb.x, F[5] = fb.y, F[6] = fb.z, F[7] = fb.w;
                    F[D * D - 1] = sm.Fc[s][t];
                }
            }
            uint32_t mbits = __float_as_uint(p4.w);
            Material m; // per-material constants: from the shared-memory copy when the table fits (it almost always does)
            {
                const uint32_t mid = mbits & MAT_ID_MASK;
                if (mats_in_smem) m = sm.mats[mid];
                else m = d.materials[mid];
            }
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
                if (any_cdf) pa = sm.aff[s][t];
                if (pa != 0u) {
                    const float4 nd = sm.nd[s][t];
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
            if (PLASTIC) plastic = sm.plastic[s][t];
            float tau[D * D];
            uint32_t flags = mbits;
            constitutive_update<D, PLASTIC>(m, flags, Fn, plastic, tau);
            const float sc = m.init_volume * inv_d * dt;
            float Cn[D * D]; // affine = grad_v * mass - stress * (V0 inv_d dt)
#pragma unroll
            for (int i = 0; i < D * D; ++i) Cn[i] = G[i] * m.mass - tau[i] * sc;

            // ---- write to the other buffer at the sorted slot
            d.pos4[nxt][k] = make_float4(new_pos.x, new_pos.y, new_pos.z, __uint_as_float(flags));
            if (SHARDED) { // whoever left the slab is listed for the next substep's migration (shard.cu)
                const int bx = assoc_cell(new_pos.x, h, inv_h) >> Dim<D>::LOG_BLOCK;
                if (bx < sm.slab_lo || bx >= sm.slab_hi) {
                    const uint32_t slot = atomicAdd(&d.counters->emig_count, 1u);
                    if (slot < d.emig_cap) d.emig_list[slot] = k;
                }
            }
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
#ifdef G2P_REPEAT
        }
#endif
    }
    cp_async_wait_all();
    TL_END(d, B200MPM_KERNEL_G2P);
    clear_sparse_grid(d, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x); // for the next substep

    // Particles of dropped blocks (capacity overflow only): carried over unchanged.
    const uint32_t dropped = d.counters->dropped_particles;
    if (dropped) {
        const uint32_t total = d.counters->sorted_total;
        for (uint32_t k = total + blockIdx.x * blockDim.x + threadIdx.x; k < total + dropped && k < d.counters->n_live; k += gridDim.x * blockDim.x) {
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

// SHARDED: peer-to-peer slab runs (DeviceData::emig_list); an instantiation of its own because the kernel sits at its
// register cap - the few extra instructions cost the unsharded path 1 % when they were merely branched over.
template <int D, bool PLASTIC, bool CPIC, bool SHARDED>
static void launch_g2p_inst(const LaunchCfg& c, const DeviceData& d, int cur) {
    constexpr int CTAS = PLASTIC ? G2P_CTAS_PLASTIC : G2P_CTAS_ELASTIC;
    auto kernel = k_g2p<D, PLASTIC, CPIC, CTAS, SHARDED>;
    constexpr size_t smem = sizeof(G2PShared<D, PLASTIC, CPIC>);
    static int resident = 0; // CTAs per SM (one process drives one device). The static work split needs the whole grid resident.
    if (!resident) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, G2P_THREADS, smem) != cudaSuccess || resident < 1) resident = 1;
        if (resident > CTAS) resident = CTAS;
        if (getenv("B200MPM_VERBOSE"))
            fprintf(stderr, "k_g2p<%d,%d,%d>: %zu B smem, %d CTAs/SM resident (wanted %d)\n", D, (int)PLASTIC, (int)CPIC, smem, resident, CTAS);
    }
    DeviceData dd = d;
    if (cur) {
        std::swap(dd.pos4[0], dd.pos4[1]);
        std::swap(dd.vel4[0], dd.vel4[1]);
        std::swap(dd.Fa[0], dd.Fa[1]);
        std::swap(dd.Fb[0], dd.Fb[1]);
        std::swap(dd.Fc[0], dd.Fc[1]);
        std::swap(dd.Ca[0], dd.Ca[1]);
        std::swap(dd.Cb[0], dd.Cb[1]);
        std::swap(dd.Cc[0], dd.Cc[1]);
        std::swap(dd.plastic[0], dd.plastic[1]);
        std::swap(dd.cdf_aff[0], dd.cdf_aff[1]);
    }
    launch_pdl(kernel, c.num_sms * resident, G2P_THREADS, smem, c.stream, dd);
}

template <int D, bool SHARDED>
static void launch_g2p_dim(const LaunchCfg& c, const DeviceData& d, int cur) {
    if (d.has_plastic) {
        if (d.has_bodies) launch_g2p_inst<D, true, true, SHARDED>(c, d, cur);
        else launch_g2p_inst<D, true, false, SHARDED>(c, d, cur);
    } else {
        if (d.has_bodies) launch_g2p_inst<D, false, true, SHARDED>(c, d, cur);
        else launch_g2p_inst<D, false, false, SHARDED>(c, d, cur);
    }
}

void launch_g2p_update(const LaunchCfg& c, const DeviceData& d, int cur) { // (+ the clearing for the next substep)
    if (d.n == 0) return launch_begin_substep(c, d);
    if (d.emig_list) {
        if (c.dim == 2) launch_g2p_dim<2, true>(c, d, cur);
        else launch_g2p_dim<3, true>(c, d, cur);
    } else {
        if (c.dim == 2) launch_g2p_dim<2, false>(c, d, cur);
        else launch_g2p_dim<3, false>(c, d, cur);
    }
    ++*c.launch_counter;
}

} // namespace b2
