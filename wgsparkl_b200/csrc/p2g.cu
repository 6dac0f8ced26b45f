// "p2g" pass: particle-to-grid transfer of mass, momentum and the APIC affine term.
//
// Reference: src/solver/p2g.wgsl:69-236 — a GATHER: one thread per node walks per-cell particle
// linked lists of the 3^D neighbouring cells (27 x max-list-length shared-memory iterations with
// two barriers each, 16 hash probes per thread).
//
// B200 design: a block-local SCATTER over cell-sorted particles.
//   * one CTA per active block, ONE THREAD PER CELL: the thread walks the contiguous run of its
//     cell's particles and reduces their 3^D stencil contributions in REGISTERS
//     (27 x 4 accumulators in 3D) — the segmented reduction over the cell's run never touches
//     memory and costs ~8 issue slots per particle-node pair, against ~25 for a lane-per-node
//     layout and ~44 for a shuffle-based segmented reduction;
//   * the per-cell partial stencils are merged into the block's (BLOCK+2)^D shared-memory tile
//     in 3^D conflict-free phases (in phase s every cell adds to node cell+s: all distinct), so the
//     shared-memory reduction needs no atomics (shared f32 atomics are CAS loops on sm_100);
//   * the tile is flushed with one vector reduction per node (RED.E.ADD.F32x4, sm_90+) into the
//     2^D blocks it overlaps, found through the per-block neighbour table instead of hash probes.
// CPIC-incompatible particle/node pairs (grid.wgsl:250-255) are skipped and turned into body
// impulses exactly like p2g.wgsl:201-226; only blocks whose tile holds a collider run that path.
#include "launch.h"

namespace b2 {

constexpr int P2G_THREADS = CELLS_PER_BLOCK;

template <int D>
struct P2GAcc {
    float a[Dim<D>::NBH][D + 1];
};

template <int D, bool CPIC>
__device__ __forceinline__ void p2g_accumulate(const DeviceData& d, int cur, uint32_t start, uint32_t end,
                                               const float* cellpos, float h, float inv_h, int tb, const uint2* tcdf,
                                               float* timp, P2GAcc<D>& acc) {
    constexpr int T = Dim<D>::TILE;
    const float4* __restrict__ pos4 = d.pos4[cur];
    const float4* __restrict__ vel4 = d.vel4[cur];
    const float4* __restrict__ Ca = d.Ca[cur];
    const float4* __restrict__ Cb = d.Cb[cur];
    const float* __restrict__ Cc = d.Cc[cur];
    for (uint32_t q = start; q < end; ++q) {
        const uint32_t id = __ldg(d.sorted_ids + q);
        const float4 p4 = __ldg(pos4 + id);
        const float4 v4 = __ldg(vel4 + id);
        float C[D * D];
        {
            float4 ca = __ldg(Ca + id);
            C[0] = ca.x, C[1] = ca.y, C[2] = ca.z, C[3] = ca.w;
            if (D == 3) {
                float4 cb = __ldg(Cb + id);
                C[4] = cb.x, C[5] = cb.y, C[6] = cb.z, C[7] = cb.w;
                C[D * D - 1] = __ldg(Cc + id);
            }
        }
        const float mass = __ldg(&d.materials[__float_as_uint(p4.w) & MAT_ID_MASK].mass);
        const float pp[3] = {p4.x, p4.y, p4.z};
        const float vv[3] = {v4.x, v4.y, v4.z};
        float d0[D], w[D][3], base[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            d0[k] = cellpos[k] - pp[k]; // dir_to_associated_grid_node (particle3d.wgsl:55-57)
            bspline(-d0[k] * inv_h, w[k][0], w[k][1], w[k][2]); // kernel.wgsl:96-104
        }
#pragma unroll
        for (int r = 0; r < D; ++r) {
            float s = mass * vv[r];
#pragma unroll
            for (int c = 0; c < D; ++c) s += C[c * D + r] * d0[c];
            base[r] = s; // affine * d0 + m v
        }
        uint32_t pa = 0;
        V3 normal = v3(0, 0, 0);
        if (CPIC) {
            pa = d.cdf_aff[cur ^ 1][q]; // this substep's affinity, written by k_g2p_cdf by sorted slot
            if (pa != 0u) {
                float4 nd = d.cdf_nd[q];
                normal = v3(nd.x, nd.y, (D == 3) ? nd.z : 0.0f);
            }
        }
        // Stencil: node s = (sx, sy, sz) in {0,1,2}^D, dpt = d0 + s h,
        // contribution w (affine dpt + m v, m) (p2g.wgsl:188-230).
#pragma unroll
        for (int sz = 0; sz < (D == 3 ? 3 : 1); ++sz) {
            float az[D];
#pragma unroll
            for (int r = 0; r < D; ++r) az[r] = (D == 3) ? base[r] + (float)sz * h * C[(D - 1) * D + r] : base[r];
            const float wz = (D == 3) ? w[D - 1][sz] : 1.0f;
#pragma unroll
            for (int sy = 0; sy < 3; ++sy) {
                float ay[D];
#pragma unroll
                for (int r = 0; r < D; ++r) ay[r] = az[r] + (float)sy * h * C[1 * D + r];
                const float wyz = w[1][sy] * wz;
#pragma unroll
                for (int sx = 0; sx < 3; ++sx) {
                    const int n = sx + 3 * sy + 9 * sz;
                    const float wt = w[0][sx] * wyz;
                    float a[D];
#pragma unroll
                    for (int r = 0; r < D; ++r) a[r] = ay[r] + (float)sx * h * C[r];
                    if (CPIC) {
                        const int idx = tb + sx + T * sy + T * T * sz;
                        const uint2 nc = tcdf[idx];
                        if (!affinities_are_compatible(nc.x, pa)) {
                            if (nc.y != NONE) { // p2g.wgsl:203-225
                                const BodyDev& body = d.bodies[nc.y];
                                V3 dpt = v3(d0[0] + (float)sx * h, d0[1] + (float)sy * h, (D == 3) ? d0[D - 1] + (float)sz * h : 0.0f);
                                V3 pv = v3(vv[0], vv[1], (D == 3) ? vv[2] : 0.0f);
                                V3 center = dpt + v3(pp[0], pp[1], (D == 3) ? pp[2] : 0.0f);
                                V3 bpv = velocity_at_point<D>(body, center);
                                V3 ghost = bpv + project_velocity(pv - bpv, normal);
                                V3 delta = (pv - ghost) * (wt * mass);
                                V3 lever = v3(body.com[0], body.com[1], (D == 3) ? body.com[2] : 0.0f) - center;
                                atomicAdd(timp + idx * 6 + 0, delta.x);
                                atomicAdd(timp + idx * 6 + 1, delta.y);
                                if (D == 3) {
                                    V3 ang = cross(delta, lever);
                                    atomicAdd(timp + idx * 6 + 2, delta.z);
                                    atomicAdd(timp + idx * 6 + 3, ang.x);
                                    atomicAdd(timp + idx * 6 + 4, ang.y);
                                    atomicAdd(timp + idx * 6 + 5, ang.z);
                                } else {
                                    atomicAdd(timp + idx * 6 + 3, delta.x * lever.y - delta.y * lever.x);
                                }
                            }
                            continue;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < D; ++r) acc.a[n][r] += wt * a[r];
                    acc.a[n][D] += wt * mass;
                }
            }
        }
    }
}

template <int D, bool CPIC>
__global__ void __launch_bounds__(P2G_THREADS) k_p2g(DeviceData d, int cur) {
    constexpr int B = Dim<D>::BLOCK, LB = Dim<D>::LOG_BLOCK, T = Dim<D>::TILE, TC = Dim<D>::TILE_CELLS;
    constexpr int NA = Dim<D>::NASSOC;
    __shared__ float4 tile[TC];
    __shared__ uint32_t s_nbr[NA];
    __shared__ uint32_t s_next;
    __shared__ uint2 tcdf[CPIC ? TC : 1];
    __shared__ float timp[CPIC ? TC * 6 : 1];

    const int t = threadIdx.x;
    const int lx = t & (B - 1), ly = (t >> LB) & (B - 1), lz = (D == 3) ? (t >> (2 * LB)) : 0;
    const int tb = lx + T * ly + T * T * lz;
    const uint32_t nb = min(d.counters->num_active_blocks, d.capacity);
    const float h = d.sim->cell_width;
    const float inv_h = 1.0f / h;

    while (true) {
        __syncthreads();
        if (t == 0) s_next = atomicAdd(&d.counters->work_p2g, 1u);
        __syncthreads();
        const uint32_t b = s_next;
        if (b >= nb) break;
        const uint32_t first = d.cell_start[b * CELLS_PER_BLOCK];
        const uint32_t last = d.cell_start[(b + 1) * CELLS_PER_BLOCK];
        if (first == last) continue; // halo block without particles: nothing to scatter
        const uint32_t start = d.cell_start[b * CELLS_PER_BLOCK + t];
        const uint32_t end = d.cell_start[b * CELLS_PER_BLOCK + t + 1];
        if (t < NA) s_nbr[t] = d.nbr[b * NA + t];
        for (int n = t; n < TC; n += P2G_THREADS) tile[n] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        int any_cdf = 0;
        if (CPIC) {
            int mine = 0;
            for (int n = t; n < TC; n += P2G_THREADS) {
                int x = n % T, y = (n / T) % T, z = n / (T * T);
                int ox = x >= B, oy = y >= B, oz = z >= B;
                uint32_t hn = s_nbr[ox + 2 * oy + 4 * oz];
                uint2 c = make_uint2(0u, NONE);
                if (hn != NONE) {
                    uint4 g = d.node_cdf[hn * CELLS_PER_BLOCK + (x - ox * B) + (y - oy * B) * B + (z - oz * B) * B * B];
                    c = make_uint2(g.y, g.z);
                }
                tcdf[n] = c;
                mine |= (c.x != 0u);
#pragma unroll
                for (int k = 0; k < 6; ++k) timp[n * 6 + k] = 0.0f;
            }
            any_cdf = __syncthreads_or(mine);
        }

        const int4 vid = d.block_vid[b];
        const float cellpos[3] = {(float)(vid.x * B + lx) * h, (float)(vid.y * B + ly) * h, (float)(vid.z * B + lz) * h};
        P2GAcc<D> acc;
#pragma unroll
        for (int n = 0; n < Dim<D>::NBH; ++n)
#pragma unroll
            for (int r = 0; r <= D; ++r) acc.a[n][r] = 0.0f;

        if (CPIC && any_cdf) p2g_accumulate<D, true>(d, cur, start, end, cellpos, h, inv_h, tb, tcdf, timp, acc);
        else p2g_accumulate<D, false>(d, cur, start, end, cellpos, h, inv_h, tb, nullptr, nullptr, acc);

        // Merge the per-cell stencils into the tile: 3^D conflict-free phases.
#pragma unroll
        for (int sz = 0; sz < (D == 3 ? 3 : 1); ++sz)
#pragma unroll
            for (int sy = 0; sy < 3; ++sy)
#pragma unroll
                for (int sx = 0; sx < 3; ++sx) {
                    const int n = sx + 3 * sy + 9 * sz;
                    const int idx = tb + sx + T * sy + T * T * sz;
                    float4 c = tile[idx];
                    c.x += acc.a[n][0];
                    c.y += acc.a[n][1];
                    c.z += acc.a[n][2];
                    if (D == 3) c.w += acc.a[n][D];
                    tile[idx] = c;
                    __syncthreads();
                }

        // Flush the tile: one 16-byte reduction per touched node.
        for (int n = t; n < TC; n += P2G_THREADS) {
            int x = n % T, y = (n / T) % T, z = n / (T * T);
            int ox = x >= B, oy = y >= B, oz = z >= B;
            uint32_t hn = s_nbr[ox + 2 * oy + 4 * oz];
            if (hn == NONE) continue; // only after a capacity overflow
            uint32_t node = hn * CELLS_PER_BLOCK + (x - ox * B) + (y - oy * B) * B + (z - oz * B) * B * B;
            float4 c = tile[n];
            if (D == 2) { // 2D stores (px, py, mass, 0)
                c.w = 0.0f;
            }
            if (c.x != 0.0f || c.y != 0.0f || c.z != 0.0f || c.w != 0.0f) atomicAdd(d.node_mv + node, c);
            if (CPIC && any_cdf) {
                uint32_t cid = tcdf[n].y;
                if (cid != NONE) { // p2g.wgsl:142-155
                    BodyDev& body = d.bodies[cid];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        float li = timp[n * 6 + k], ai = timp[n * 6 + 3 + k];
                        if (li != 0.0f) atomicAdd(&body.imp_lin[k], flt2int(li));
                        if (ai != 0.0f) atomicAdd(&body.imp_ang[k], flt2int(ai));
                    }
                }
            }
        }
    }
}

void launch_p2g(const LaunchCfg& c, const DeviceData& d, int cur) {
    if (d.n == 0) return;
    const int grid = c.num_sms * 8;
    if (c.dim == 2) {
        if (d.has_bodies) k_p2g<2, true><<<grid, P2G_THREADS, 0, c.stream>>>(d, cur);
        else k_p2g<2, false><<<grid, P2G_THREADS, 0, c.stream>>>(d, cur);
    } else {
        if (d.has_bodies) k_p2g<3, true><<<grid, P2G_THREADS, 0, c.stream>>>(d, cur);
        else k_p2g<3, false><<<grid, P2G_THREADS, 0, c.stream>>>(d, cur);
    }
    ++*c.launch_counter;
}

} // namespace b2
