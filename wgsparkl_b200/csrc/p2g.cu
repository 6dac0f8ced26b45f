// "p2g" pass: particle-to-grid transfer of mass, momentum and the APIC affine term.
//
// Reference: src/solver/p2g.wgsl:69-236 — a GATHER: one thread per node walks per-cell particle
// linked lists of the 3^D neighbouring cells (27 x max-list-length shared-memory iterations with
// two barriers each, 16 hash probes per thread).
//
// B200 design: a block-local SCATTER over cell-sorted particles.
//   * one WARP per half block (32 cells): a CTA is a single warp and an independent worker with its own
//     staging buffers, tile and position in the dynamic work queue, so there is no CTA barrier anywhere;
//     the warp's particles (a contiguous range of the sorted order, in windows of 256) are staged into
//     shared memory with cp.async (LDGSTS, 16 bytes per request, no register staging);
//   * ONE LANE PER CELL: the lane walks the contiguous run of its cell's particles and reduces
//     their 3^D stencil contributions in REGISTERS (27 x 4 accumulators in 3D) — the segmented
//     reduction over the cell's run never touches memory and costs ~8 issue slots per
//     particle-node pair, against ~25 for a lane-per-node layout and ~44 for a shuffle-based
//     segmented reduction;
//   * the per-cell partial stencils are merged into the warp's (BLOCK+2)^D shared-memory tile
//     in 3^D conflict-free phases (in phase s every cell adds to node cell+s: all distinct), so the
//     shared-memory reduction needs no atomics (shared f32 atomics are CAS loops on sm_100);
//   * the tile is flushed with one vector reduction per node (RED.E.ADD.F32x4, sm_90+) into the
//     2^D blocks it overlaps, found through the per-block neighbour table instead of hash probes.
// CPIC-incompatible particle/node pairs (grid.wgsl:250-255) are skipped and turned into body
// impulses exactly like p2g.wgsl:201-226; only the blocks whose tile holds a collider run that
// instantiation (a second pass over the staged particles accumulates the per-node impulses with
// the same register/phase scheme, so the impulse path has no floating-point atomics either).
#include "launch.h"

#include <cstdio>
#include <cstdlib>
#include <utility>

namespace b2 {

constexpr int P2G_CHUNK = 256; // particles staged per pass and warp: 32 cells x 8 (the reference's seeding density)

enum { P2G_FAST = 0, P2G_CPIC_MOMENTUM = 1, P2G_CPIC_IMPULSE = 2 };

template <int NBH, int W>
struct P2GAcc {
    float a[NBH][W];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int n = 0; n < NBH; ++n)
#pragma unroll
            for (int r = 0; r < W; ++r) a[n][r] = 0.0f;
    }
};

// Staged particle record in shared memory (SoA of float4, 64 bytes per particle):
//   sp = (x, y, z, mass)   sv = (vx, vy, vz, -)   sa = C[0..3]   sb = C[4..7]   sc = C[8]
// Slot i is stored at i ^ ((i >> 3) & 7): a cell's run starts at ~8 t for thread t, so consecutive lanes
// would otherwise hit the same bank group on every 16-byte load (8-way conflict).
__device__ __forceinline__ int p2g_swz(int i) { return i ^ ((i >> 3) & 7); }

// Walks the staged particles [lo, hi) of one cell.
//   MODE FAST / CPIC_MOMENTUM: acc[n] += w (affine dpt + m v, m)       (p2g.wgsl:188-230)
//   MODE CPIC_IMPULSE        : acc[n] += (delta_impulse, delta_ang)     (p2g.wgsl:203-225)
template <int D, int MODE, int W>
__device__ __forceinline__ bool p2g_accumulate(const DeviceData& d, int cur, uint32_t base, int lo, int hi,
                                               const float4* sp, const float4* sv, const float4* sa, const float4* sb,
                                               const float* sc, const uint32_t* s_aff, const float* cellpos, float h, float inv_h, int tb,
                                               const uint2* tcdf, P2GAcc<Dim<D>::NBH, W>& acc) {
    constexpr int T = Dim<D>::TILE;
    bool any_incompatible = false;
    for (int i = lo; i < hi; ++i) {
        const int s = p2g_swz(i);
        const float4 p4 = sp[s];
        const float4 v4 = sv[s];
        float C[D * D];
        {
            float4 ca = sa[s];
            C[0] = ca.x, C[1] = ca.y, C[2] = ca.z, C[3] = ca.w;
            if (D == 3) {
                float4 cb = sb[s];
                C[4] = cb.x, C[5] = cb.y, C[6] = cb.z, C[7] = cb.w;
                C[D * D - 1] = sc[s];
            }
        }
        const float mass = p4.w;
        const float pp[3] = {p4.x, p4.y, p4.z};
        const float vv[3] = {v4.x, v4.y, v4.z};
        float d0[D], w[D][3], bs[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            d0[k] = cellpos[k] - pp[k]; // dir_to_associated_grid_node (particle3d.wgsl:55-57)
            bspline(-d0[k] * inv_h, w[k][0], w[k][1], w[k][2]); // kernel.wgsl:96-104
        }
#pragma unroll
        for (int r = 0; r < D; ++r) {
            float t = mass * vv[r];
#pragma unroll
            for (int c = 0; c < D; ++c) t += C[c * D + r] * d0[c];
            bs[r] = t; // affine * d0 + m v
        }
        uint32_t pa = 0;
        V3 normal = v3(0, 0, 0);
        if (MODE != P2G_FAST) {
            pa = s_aff[s];
            if (MODE == P2G_CPIC_IMPULSE && pa == 0u) continue; // compatible with every node: no impulse
            if (MODE == P2G_CPIC_IMPULSE) {
                float4 nd = d.cdf_nd[base + (uint32_t)i];
                normal = v3(nd.x, nd.y, (D == 3) ? nd.z : 0.0f);
            }
        }
        // Momentum pass next to a collider: which of the 3^D nodes are CPIC-incompatible with this particle
        // (grid.wgsl:250-255). A particle away from every collider (pa == 0, the majority even in collider-side
        // blocks) is compatible with all of them and never looks at the node colours.
        uint32_t bad = 0u;
        if (MODE == P2G_CPIC_MOMENTUM && pa != 0u) {
#pragma unroll
            for (int n = 0; n < Dim<D>::NBH; ++n) {
                const uint2 nc = tcdf[tb + (n % 3) + T * ((n / 3) % 3) + T * T * (n / 9)];
                if (!affinities_are_compatible(nc.x, pa)) {
                    bad |= 1u << n;
                    if (nc.y != NONE) any_incompatible = any_incompatible || (d.bodies[nc.y].needs_impulse != 0u);
                }
            }
        }
#pragma unroll
        for (int sz = 0; sz < (D == 3 ? 3 : 1); ++sz) {
            float az[D];
#pragma unroll
            for (int r = 0; r < D; ++r) az[r] = (D == 3) ? bs[r] + (float)sz * h * C[(D - 1) * D + r] : bs[r];
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
                    if (MODE == P2G_CPIC_MOMENTUM && ((bad >> n) & 1u)) continue;
                    if (MODE == P2G_CPIC_IMPULSE) {
                        const uint2 nc = tcdf[tb + sx + T * sy + T * T * sz];
                        const bool compatible = affinities_are_compatible(nc.x, pa);
                        {
                            if (!compatible && nc.y != NONE && d.bodies[nc.y].needs_impulse) { // p2g.wgsl:203-225
                                const BodyDev& body = d.bodies[nc.y];
                                V3 dpt = v3(d0[0] + (float)sx * h, d0[1] + (float)sy * h, (D == 3) ? d0[D - 1] + (float)sz * h : 0.0f);
                                V3 pv = v3(vv[0], vv[1], (D == 3) ? vv[2] : 0.0f);
                                V3 center = dpt + v3(pp[0], pp[1], (D == 3) ? pp[2] : 0.0f);
                                V3 bpv = velocity_at_point<D>(body, center);
                                V3 ghost = bpv + project_velocity(pv - bpv, normal);
                                V3 delta = (pv - ghost) * (wt * mass);
                                V3 lever = v3(body.com[0], body.com[1], (D == 3) ? body.com[2] : 0.0f) - center;
                                acc.a[n][0] += delta.x;
                                acc.a[n][1] += delta.y;
                                if (D == 3) {
                                    V3 ang = cross(delta, lever);
                                    acc.a[n][2] += delta.z;
                                    acc.a[n][3] += ang.x;
                                    acc.a[n][4] += ang.y;
                                    acc.a[n][W - 1] += ang.z;
                                } else {
                                    acc.a[n][2] += delta.x * lever.y - delta.y * lever.x;
                                }
                            }
                            continue;
                        }
                    }
                    if (MODE != P2G_CPIC_IMPULSE) {
#pragma unroll
                        for (int r = 0; r < D; ++r) acc.a[n][r] += wt * (ay[r] + (float)sx * h * C[r]);
                        acc.a[n][D] += wt * mass;
                    }
                }
            }
        }
    }
    return any_incompatible;
}

// One WARP owns half a block (32 cells, one lane per cell) and is an independent worker: its own staging
// buffers, its own (BLOCK+2)^D tile, its own dynamic work queue position — there is no CTA-wide barrier anywhere
// in the kernel (a CTA is a single warp), so some warps stage while others compute.
//   CPIC = false: blocks whose tile holds no collider (block_flags == 0, or no bodies at all).
//   CPIC = true : the few blocks next to a collider (compact list), further split into PARTS work items.
//   IMP (with CPIC): some body can react to impulses (DeviceData::bodies_react) - without it the per-node impulse
//   accumulators (27 x 6 registers, 5 KB of shared memory) and the second pass are compiled out.
template <int D, bool CPIC, bool IMP>
__global__ void __launch_bounds__(32, CPIC ? (IMP ? 7 : 9) : 10) k_p2g(DeviceData d, int cur) {
    constexpr int B = Dim<D>::BLOCK, LB = Dim<D>::LOG_BLOCK, T = Dim<D>::TILE, TC = Dim<D>::TILE_CELLS;
    constexpr int NA = Dim<D>::NASSOC, NBH = Dim<D>::NBH;
    constexpr int WI = (D == 3) ? 6 : 3; // impulse components per node
    constexpr int CHUNK = P2G_CHUNK; // particles staged per pass = 8 per cell, the reference's seeding density
    constexpr int HALF = CELLS_PER_BLOCK / 2;
    __shared__ float4 tile[TC];
    __shared__ float4 sp[CHUNK], sv[CHUNK], sa[CHUNK];
    __shared__ float4 sb[D == 3 ? CHUNK : 1];
    __shared__ float sc[D == 3 ? CHUNK : 1];
    __shared__ uint32_t s_ids[CHUNK];
    __shared__ uint32_t s_aff[CPIC ? CHUNK : 1]; // this substep's particle affinities (k_g2p_cdf, by sorted slot)
    __shared__ uint32_t s_nbr[NA];
    __shared__ uint2 tcdf[CPIC ? TC : 1];
    __shared__ float timp[IMP ? TC * WI : 1];

    const int lane = threadIdx.x;
    // A collider-side half-block is split into PARTS work items (each takes every cell's PARTS-th share of the
    // run): these blocks are few, so their latency — not throughput — is what shows up.
    constexpr uint32_t PARTS = CPIC ? 2u : 1u;
    // CPIC = false walks p2g_list (k_scatter: blocks that hold particles and see no collider, the densely populated
    // ones first - longest items first keeps the last scheduling round short).
    const uint32_t nfront = d.counters->num_p2g_front;
    const uint32_t nwork = (CPIC ? d.counters->num_cpic_blocks * PARTS : nfront + d.counters->num_p2g_back) * 2u;
    const float h = d.sim->cell_width;
    const float inv_h = 1.0f / h;
    uint32_t* work = CPIC ? &d.counters->work_p2g_cpic : &d.counters->work_p2g;
    const float4* __restrict__ pos4 = d.pos4[cur];
    const float4* __restrict__ vel4 = d.vel4[cur];
    const float4* __restrict__ Ca = d.Ca[cur];
    const float4* __restrict__ Cb = d.Cb[cur];
    const float* __restrict__ Cc = d.Cc[cur];

    // Stage one chunk of the sorted range into shared memory (gather through sorted_ids, a
    // near-identity permutation of the current buffers).
    auto stage = [&](uint32_t base, int cn) {
        // ids go through shared memory so that the request loop below stays rolled (few live registers
        // next to the 108 accumulators); each lane only reads back the ids it wrote itself.
#pragma unroll
        for (int j = 0; j < CHUNK / 32; ++j) {
            const int i = lane + j * 32;
            if (i < cn) s_ids[i] = __ldg(d.sorted_ids + base + i);
        }
#pragma unroll 1
        for (int i = lane; i < cn; i += 32) {
            const uint32_t id = s_ids[i];
            const int s = p2g_swz(i);
            cp_async16(sp + s, pos4 + id);
            cp_async16(sv + s, vel4 + id);
            cp_async16(sa + s, Ca + id);
            if (D == 3) {
                cp_async16(sb + s, Cb + id);
                cp_async4(sc + s, Cc + id);
            }
            if (CPIC) cp_async4(s_aff + s, d.cdf_aff[cur ^ 1] + base + i);
        }
        cp_async_wait_all();
        // (x, y, z, material bits) -> (x, y, z, mass)
#pragma unroll 1
        for (int i = lane; i < cn; i += 32) {
            const int s = p2g_swz(i);
            const uint32_t mbits = __float_as_uint(sp[s].w);
            sp[s].w = __ldg(&d.materials[mbits & MAT_ID_MASK].mass);
        }
        __syncwarp();
    };

    while (true) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(work, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= nwork) break;
        const uint32_t half = w & 1u;
        const uint32_t part = (w >> 1) % PARTS;
        uint32_t b;
        if (CPIC) {
            b = d.cpic_list[w / (2u * PARTS)];
        } else {
            const uint32_t e = w >> 1;
            b = d.p2g_list[e < nfront ? e : d.capacity - 1u - (e - nfront)];
        }
        const uint32_t cell = half * HALF + lane; // this lane's cell of the block
        const uint32_t first = d.cell_start[b * CELLS_PER_BLOCK + half * HALF];
        const uint32_t last = d.cell_start[b * CELLS_PER_BLOCK + half * HALF + HALF];
        if (first == last) continue; // nothing to scatter from this half
        uint32_t start = d.cell_start[b * CELLS_PER_BLOCK + cell];
        uint32_t end = d.cell_start[b * CELLS_PER_BLOCK + cell + 1];
        if (PARTS > 1) {
            const uint32_t len = end - start;
            end = start + (len * (part + 1)) / PARTS;
            start = start + (len * part) / PARTS;
        }
        const int lx = cell & (B - 1), ly = (cell >> LB) & (B - 1), lz = (D == 3) ? (cell >> (2 * LB)) : 0;
        const int tb = lx + T * ly + T * T * lz;
        __syncwarp(); // the previous work item's tile / s_nbr are no longer read
        if (lane < NA) s_nbr[lane] = d.nbr[b * NA + lane];
        for (int n = lane; n < TC; n += 32) tile[n] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        if (CPIC) {
            for (int n = lane; n < TC; n += 32) {
                int x = n % T, y = (n / T) % T, z = n / (T * T);
                int ox = x >= B, oy = y >= B, oz = z >= B;
                uint32_t hn = s_nbr[ox + 2 * oy + 4 * oz];
                uint2 c = make_uint2(0u, NONE);
                if (hn != NONE) {
                    uint4 g = d.node_cdf[hn * CELLS_PER_BLOCK + (x - ox * B) + (y - oy * B) * B + (z - oz * B) * B * B];
                    c = make_uint2(g.z, g.x); // (affinities, closest_id)
                }
                tcdf[n] = c;
                if (IMP) {
#pragma unroll
                    for (int k = 0; k < WI; ++k) timp[n * WI + k] = 0.0f;
                }
            }
            __syncwarp();
        }

        const int4 vid = d.block_vid[b];
        const float cellpos[3] = {(float)(vid.x * B + lx) * h, (float)(vid.y * B + ly) * h, (float)(vid.z * B + lz) * h};
        bool incompatible = false;
        {
            P2GAcc<NBH, D + 1> acc;
            acc.clear();
            for (uint32_t base = first; base < last; base += CHUNK) {
                const int cn = (int)min((uint32_t)CHUNK, last - base);
                __syncwarp(); // the previous chunk is no longer in use
                stage(base, cn);
                const int lo = (int)(max(start, base) - base);
                const int hi = (int)(min(end, base + (uint32_t)cn) - base);
                incompatible |= p2g_accumulate<D, CPIC ? P2G_CPIC_MOMENTUM : P2G_FAST, D + 1>(
                    d, cur, base, lo, hi, sp, sv, sa, sb, sc, s_aff, cellpos, h, inv_h, tb, tcdf, acc);
            }
            // Merge the per-cell stencils into the tile: 3^D conflict-free phases (within a phase the 32 lanes
            // add to 32 distinct nodes).
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
                        __syncwarp();
                    }
        }
        if (IMP) {
            // Second pass, only if some particle/node pair of this work item is CPIC-incompatible with a collider
            // that can react: per-node body impulses (p2g.wgsl:201-226).
            if (__any_sync(0xffffffffu, incompatible)) {
                P2GAcc<NBH, WI> imp;
                imp.clear();
                const bool single_chunk = (last - first) <= (uint32_t)CHUNK; // still staged
                for (uint32_t base = first; base < last; base += CHUNK) {
                    const int cn = (int)min((uint32_t)CHUNK, last - base);
                    if (!single_chunk) {
                        __syncwarp();
                        stage(base, cn);
                    }
                    const int lo = (int)(max(start, base) - base);
                    const int hi = (int)(min(end, base + (uint32_t)cn) - base);
                    p2g_accumulate<D, P2G_CPIC_IMPULSE, WI>(d, cur, base, lo, hi, sp, sv, sa, sb, sc, s_aff, cellpos, h, inv_h, tb, tcdf, imp);
                }
#pragma unroll
                for (int sz = 0; sz < (D == 3 ? 3 : 1); ++sz)
#pragma unroll
                    for (int sy = 0; sy < 3; ++sy)
#pragma unroll
                        for (int sx = 0; sx < 3; ++sx) {
                            const int n = sx + 3 * sy + 9 * sz;
                            const int idx = tb + sx + T * sy + T * T * sz;
#pragma unroll
                            for (int k = 0; k < WI; ++k) timp[idx * WI + k] += imp.a[n][k];
                            __syncwarp();
                        }
            }
        }

        // Flush the tile: one 16-byte reduction per touched node.
        for (int n = lane; n < TC; n += 32) {
            int x = n % T, y = (n / T) % T, z = n / (T * T);
            int ox = x >= B, oy = y >= B, oz = z >= B;
            uint32_t hn = s_nbr[ox + 2 * oy + 4 * oz];
            if (hn == NONE) continue; // only after a capacity overflow
            uint32_t node = hn * CELLS_PER_BLOCK + (x - ox * B) + (y - oy * B) * B + (z - oz * B) * B * B;
            float4 c = tile[n];
            if (D == 2) c.w = 0.0f; // 2D stores (px, py, mass, 0)
            if (c.x != 0.0f || c.y != 0.0f || c.z != 0.0f || c.w != 0.0f) atomicAdd(d.node_mv + node, c);
            if (IMP) {
                uint32_t cid = tcdf[n].y;
                if (cid != NONE) { // p2g.wgsl:142-155: integer atomics, i32(x * 1e5)
                    BodyDev& body = d.bodies[cid];
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        float li = timp[n * WI + k];
                        if (li != 0.0f) atomicAdd(&body.imp_lin[k], flt2int(li));
                    }
#pragma unroll
                    for (int k = 0; k < WI - D; ++k) {
                        float ai = timp[n * WI + D + k];
                        if (ai != 0.0f) atomicAdd(&body.imp_ang[k], flt2int(ai));
                    }
                }
            }
        }
    }
}

// =====================================================================================================================
// k_p2g_fast: the blocks whose tile holds no collider (all of them without bodies) - the bulk of the particles.
//
// One lane per cell keeps the segmented reduction over a cell's run in registers, but 3^D x (D+1) accumulators per
// thread (168 registers) leave room for only 2-3 warps per scheduler, and every warp then spends most of its time in
// latency it cannot cover itself (ncu: each warp issues in 18-20 % of its cycles). Here the 3^D stencil is SLICED
// along the last axis over the THREE WARPS of a CTA: warp w owns the nodes with shift w along z (y in 2D), i.e.
// 9 x 4 accumulators, ~80 registers, ~20 warps per SM. All three warps walk the same staged particles of the same
// half block (32 cells, lane = cell). The CTA is a software pipeline over the stage table k_scatter builds
// (DeviceData::p2g_stages), a stage = up to P2G_K particles of every cell of one half block:
//   stage g+4: the table entry                                              [cp.async 8 B,  thread 0]
//   stage g+3: 33 cell boundaries, neighbour table, block id of its block    [cp.async,      warp 2]
//   stage g+2: ids of the stage's particles (sorted_ids)                     [cp.async 4 B,  warp 2]
//   stage g+1: their SoA records (pos + vel: warp 0, affine: warp 1, last affine word: warp 2)   [cp.async 16 B]
//   stage g  : cp.async.wait_all + one CTA barrier, then P2G_K particles x 3^(D-1) nodes of FMAs per lane and warp.
// Requests are issued TRANSPOSED - request q of lane l is particle l & 3 of cell 8 q + (l >> 2) - so that one
// warp-wide request covers 8 runs of 64 contiguous bytes instead of 32 different lines. At the end of a half block
// every warp merges its slice into its own tile in 3^(D-1) conflict-free phases, and the CTA flushes the sum of the
// three tiles with one RED.ADD.F32x4 per node. Work distribution is static and exact: the CTAs take equal
// contiguous shares of the stage table; a half block cut by a share boundary is flushed twice (the node
// reductions commute).
constexpr int P2G_SQ = 8, P2G_CQ = 4; // rings: table entries (4 stages ahead), cell ranges (3 stages ahead)
constexpr int P2G_FAST_THREADS = 96;
#ifndef P2G_FAST_CTAS
#define P2G_FAST_CTAS 7
#endif

template <int D>
struct __align__(16) P2GFastShared {
    static constexpr int TC = Dim<D>::TILE_CELLS;
    struct __align__(16) Cells {
        uint32_t start[36]; // the 33 boundaries of the half block's 32 cell runs (sorted slots)
        uint32_t nbr[8]; // header ids of the blocks the tile overlaps
        int4 vid; // block coordinates
    };
    float4 tile[3][TC]; // one per warp (slice)
    // records of stage g in [g & 1][particle j of the cell][cell]
    float4 sp[2][P2G_K][32], sv[2][P2G_K][32], sa[2][P2G_K][32];
    float4 sb[2][D == 3 ? P2G_K : 1][32];
    float sc[2][D == 3 ? P2G_K : 1][32];
    uint32_t ids[2][P2G_K][32];
    Cells cells[P2G_CQ];
    uint2 dsc[P2G_SQ];
    float mass[16]; // masses of the first 16 materials
};

template <int D>
__global__ void __launch_bounds__(P2G_FAST_THREADS, P2G_FAST_CTAS) k_p2g_fast(DeviceData d) {
    // (the launch wrapper hands over the current particle arrays at index 0)
    constexpr int B = Dim<D>::BLOCK, LB = Dim<D>::LOG_BLOCK, T = Dim<D>::TILE, TC = Dim<D>::TILE_CELLS;
    constexpr int NA = Dim<D>::NASSOC;
    constexpr int K = (int)P2G_K;
    constexpr int NS = (D == 3) ? 9 : 3; // nodes of one slice
    constexpr int TS = (D == 3) ? T * T : T; // tile stride of the sliced axis
    __shared__ P2GFastShared<D> sm;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5; // warp = slice = shift along the last axis
    const uint32_t total = min(d.counters->num_p2g_stages, d.p2g_stages_cap);
    const uint32_t share = (total + gridDim.x - 1u) / gridDim.x;
    const uint32_t f0 = blockIdx.x * share;
    if (f0 >= total) return;
    const uint32_t ng = min(share, total - f0); // this CTA's stages: table entries [f0, f0 + ng)
    const float h = d.sim->cell_width;
    const float inv_h = 1.0f / h;
    const bool mass_in_smem = d.num_materials <= 16u;
    if (mass_in_smem && t < (int)d.num_materials) sm.mass[t] = d.materials[t].mass;

    auto request_entry = [&](uint32_t g) {
        if (t != 0) return;
        if (g < ng) cp_async8(&sm.dsc[g % P2G_SQ], d.p2g_stages + f0 + g);
        else sm.dsc[g % P2G_SQ] = make_uint2(NONE, 0u);
    };
    auto request_cells = [&](uint32_t g) { // warp 2
        const uint2 e = sm.dsc[g % P2G_SQ];
        if (e.x == NONE) return;
        const uint32_t b = e.x & 0x7fffffffu, half = e.x >> 31;
        typename P2GFastShared<D>::Cells& c = sm.cells[g % P2G_CQ];
        const uint32_t* src = d.cell_start + b * CELLS_PER_BLOCK + half * 32u;
        cp_async4(&c.start[lane], src + lane);
        if (lane == 0) cp_async4(&c.start[32], src + 32);
        if (lane < NA) cp_async4(&c.nbr[lane], d.nbr + b * NA + lane);
        if (lane == 8) cp_async16(&c.vid, d.block_vid + b);
    };
    // cell c's particles of stage g: sorted slots [lo, lo + n), n <= P2G_K
    auto cell_range = [&](uint32_t g, int c, uint32_t& lo, int& n) {
        const uint32_t st = sm.dsc[g % P2G_SQ].y & 0xffffu;
        const typename P2GFastShared<D>::Cells& cl = sm.cells[g % P2G_CQ];
        lo = cl.start[c] + P2G_K * st;
        const uint32_t end = cl.start[c + 1];
        n = end > lo ? (int)min(end - lo, P2G_K) : 0;
    };
    static_assert(P2G_K == 4, "the transposed request mapping assumes 4 particles per cell and stage");
    auto request_ids = [&](uint32_t g) { // warp 2
        if (sm.dsc[g % P2G_SQ].x == NONE) return;
        const int j = lane & 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = 8 * q + (lane >> 2);
            uint32_t lo;
            int n;
            cell_range(g, c, lo, n);
            if (j < n) cp_async4(&sm.ids[g & 1u][j][c], d.sorted_ids + lo + j);
        }
    };
    auto request_records = [&](uint32_t g) { // all warps, a share of the fields each
        if (sm.dsc[g % P2G_SQ].x == NONE) return;
        const int st = (int)(g & 1u), j = lane & 3;
        uint32_t id[4];
        bool have[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = 8 * q + (lane >> 2);
            uint32_t lo;
            int n;
            cell_range(g, c, lo, n);
            have[q] = j < n;
            id[q] = sm.ids[st][j][c];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = 8 * q + (lane >> 2);
            if (have[q]) {
                if (warp == 0) {
                    cp_async16(&sm.sp[st][j][c], d.pos4[0] + id[q]);
                    cp_async16(&sm.sv[st][j][c], d.vel4[0] + id[q]);
                } else if (warp == 1) {
                    cp_async16(&sm.sa[st][j][c], d.Ca[0] + id[q]);
                    if (D == 3) cp_async16(&sm.sb[st][j][c], d.Cb[0] + id[q]);
                } else {
                    if (D == 3) cp_async4(&sm.sc[st][j][c], d.Cc[0] + id[q]);
                }
            }
        }
    };

    // ---- prologue (four exposed latencies, once per CTA)
    for (uint32_t g = 0; g < 4; ++g) request_entry(g);
    cp_async_wait_all();
    __syncthreads();
    if (warp == 2)
        for (uint32_t g = 0; g < 3; ++g) request_cells(g);
    cp_async_wait_all();
    __syncthreads();
    if (warp == 2) {
        request_ids(0);
        request_ids(1);
    }
    cp_async_wait_all();
    __syncthreads();
    request_records(0);

    float acc[NS][D + 1];
    for (uint32_t g = 0; g < ng; ++g) {
        cp_async_wait_all();
        __syncthreads(); // everything requested during the previous stage has landed and is visible to the CTA
        const uint2 e = sm.dsc[g % P2G_SQ];
        const bool first_of_half = (g == 0u) || (sm.dsc[(g - 1u) % P2G_SQ].x != e.x);
        const bool last_of_half = (g + 1u == ng) || (sm.dsc[(g + 1u) % P2G_SQ].x != e.x);
        request_records(g + 1);
        if (warp == 2) {
            request_ids(g + 2);
            request_cells(g + 3);
        }
        request_entry(g + 4);

        const uint32_t half = e.x >> 31;
        const uint32_t cell = half * 32u + (uint32_t)lane; // this lane's cell of the block
        const int lx = cell & (B - 1), ly = (cell >> LB) & (B - 1), lz = (D == 3) ? (cell >> (2 * LB)) : 0;
        const int tb = lx + T * ly + T * T * lz + TS * warp; // lowest tile node of this warp's slice of the lane's stencil
        const typename P2GFastShared<D>::Cells& cells = sm.cells[g % P2G_CQ];
        float4* const tile = sm.tile[warp];
        if (first_of_half) {
#pragma unroll
            for (int n = 0; n < NS; ++n)
#pragma unroll
                for (int r = 0; r <= D; ++r) acc[n][r] = 0.0f;
            for (int n = lane; n < TC; n += 32) tile[n] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        {
            const int4 vid = cells.vid;
            const float cellpos[3] = {(float)(vid.x * B + lx) * h, (float)(vid.y * B + ly) * h, (float)(vid.z * B + lz) * h};
            uint32_t lo;
            int np;
            cell_range(g, lane, lo, np);
            const int st = (int)(g & 1u);
            for (int j = 0; j < np; ++j) { // p2g.wgsl:188-230 for one particle and one slice: acc[n] += w (affine dpt + m v, m)
                const float4 p4 = sm.sp[st][j][lane];
                const float4 v4 = sm.sv[st][j][lane];
                float C[D * D];
                {
                    const float4 ca = sm.sa[st][j][lane];
                    C[0] = ca.x, C[1] = ca.y, C[2] = ca.z, C[3] = ca.w;
                    if (D == 3) {
                        const float4 cb = sm.sb[st][j][lane];
                        C[4] = cb.x, C[5] = cb.y, C[6] = cb.z, C[7] = cb.w;
                        C[D * D - 1] = sm.sc[st][j][lane];
                    }
                }
                const uint32_t mid = __float_as_uint(p4.w) & MAT_ID_MASK;
                const float mass = mass_in_smem ? sm.mass[mid] : __ldg(&d.materials[mid].mass);
                const float pp[3] = {p4.x, p4.y, p4.z};
                const float vv[3] = {v4.x, v4.y, v4.z};
                float d0[D], w[D][3], bs[D];
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    d0[k] = cellpos[k] - pp[k]; // dir_to_associated_grid_node (particle3d.wgsl:55-57)
                    bspline(-d0[k] * inv_h, w[k][0], w[k][1], w[k][2]); // kernel.wgsl:96-104
                }
#pragma unroll
                for (int r = 0; r < D; ++r) {
                    float s0 = mass * vv[r];
#pragma unroll
                    for (int c = 0; c < D; ++c) s0 += C[c * D + r] * d0[c];
                    bs[r] = s0; // affine * d0 + m v
                }
                // this warp's shift along the last axis: weight and affine offset
                const float ws = (warp == 0) ? w[D - 1][0] : (warp == 1) ? w[D - 1][1] : w[D - 1][2];
                float as[D];
#pragma unroll
                for (int r = 0; r < D; ++r) as[r] = bs[r] + (float)warp * h * C[(D - 1) * D + r];
                if (D == 3) {
#pragma unroll
                    for (int sy = 0; sy < 3; ++sy) {
                        float ay[D];
#pragma unroll
                        for (int r = 0; r < D; ++r) ay[r] = as[r] + (float)sy * h * C[1 * D + r];
                        const float wyz = w[1][sy] * ws;
#pragma unroll
                        for (int sx = 0; sx < 3; ++sx) {
                            const int n = sx + 3 * sy;
                            const float wt = w[0][sx] * wyz;
#pragma unroll
                            for (int r = 0; r < D; ++r) acc[n][r] += wt * (ay[r] + (float)sx * h * C[r]);
                            acc[n][D] += wt * mass;
                        }
                    }
                } else {
#pragma unroll
                    for (int sx = 0; sx < 3; ++sx) {
                        const float wt = w[0][sx] * ws;
#pragma unroll
                        for (int r = 0; r < D; ++r) acc[sx][r] += wt * (as[r] + (float)sx * h * C[r]);
                        acc[sx][D] += wt * mass;
                    }
                }
            }
        }
        if (last_of_half) {
            // Merge the lanes' slices into the warp's tile: 3^(D-1) conflict-free phases (within a phase the 32 lanes
            // add to 32 distinct nodes); then the CTA flushes the sum of the three tiles, one 16-byte reduction per node.
            __syncwarp();
#pragma unroll
            for (int n = 0; n < NS; ++n) {
                const int idx = tb + (n % 3) + ((D == 3) ? T * (n / 3) : 0);
                float4 c = tile[idx];
                c.x += acc[n][0];
                c.y += acc[n][1];
                c.z += acc[n][2];
                if (D == 3) c.w += acc[n][D];
                tile[idx] = c;
                __syncwarp();
            }
            __syncthreads();
            for (int n = t; n < TC; n += P2G_FAST_THREADS) {
                const int x = n % T, y = (n / T) % T, z = n / (T * T);
                const int ox = x >= B, oy = y >= B, oz = z >= B;
                const uint32_t hn = cells.nbr[ox + 2 * oy + 4 * oz];
                if (hn == NONE) continue; // only after a capacity overflow
                const uint32_t node = hn * CELLS_PER_BLOCK + (x - ox * B) + (y - oy * B) * B + (z - oz * B) * B * B;
                const float4 c0 = sm.tile[0][n], c1 = sm.tile[1][n], c2 = sm.tile[2][n];
                float4 c = make_float4(c0.x + c1.x + c2.x, c0.y + c1.y + c2.y, c0.z + c1.z + c2.z, c0.w + c1.w + c2.w);
                if (D == 2) c.w = 0.0f; // 2D stores (px, py, mass, 0)
                if (c.x != 0.0f || c.y != 0.0f || c.z != 0.0f || c.w != 0.0f) atomicAdd(d.node_mv + node, c);
            }
            // (the next stage's barrier separates these reads from the next half block's zeroing of the tiles)
        }
    }
    cp_async_wait_all();
}

#ifndef P2G_USE_FAST
#define P2G_USE_FAST 1
#endif
void launch_p2g(const LaunchCfg& c, const DeviceData& d, int cur) {
    if (d.n == 0) return;
#if P2G_USE_FAST
    DeviceData dd = d;
    if (cur) {
        std::swap(dd.pos4[0], dd.pos4[1]);
        std::swap(dd.vel4[0], dd.vel4[1]);
        std::swap(dd.Ca[0], dd.Ca[1]);
        std::swap(dd.Cb[0], dd.Cb[1]);
        std::swap(dd.Cc[0], dd.Cc[1]);
    }
    static int resident[2] = {0, 0}; // CTAs per SM (the static work split wants the whole grid resident)
    int& res = resident[c.dim - 2];
    if (!res) {
        if (c.dim == 2) {
            cudaFuncSetAttribute(k_p2g_fast<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, k_p2g_fast<2>, P2G_FAST_THREADS, 0);
        } else {
            cudaFuncSetAttribute(k_p2g_fast<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, k_p2g_fast<3>, P2G_FAST_THREADS, 0);
        }
        if (res < 1) res = 1;
        if (getenv("B200MPM_VERBOSE")) fprintf(stderr, "k_p2g_fast<%d>: %d CTAs/SM resident\n", c.dim, res);
    }
    const int grid = c.num_sms * res;
    if (c.dim == 2) k_p2g_fast<2><<<grid, P2G_FAST_THREADS, 0, c.stream>>>(dd);
    else k_p2g_fast<3><<<grid, P2G_FAST_THREADS, 0, c.stream>>>(dd);
#else
    const int grid = c.num_sms * 10; // 10 single-warp CTAs per SM (shared memory: 22 KB each)
    if (c.dim == 2) k_p2g<2, false, false><<<grid, 32, 0, c.stream>>>(d, cur);
    else k_p2g<3, false, false><<<grid, 32, 0, c.stream>>>(d, cur);
#endif
    ++*c.launch_counter;
}

// The blocks next to a collider (compact list built by k_scatter). Independent of launch_p2g: the two
// instantiations touch disjoint blocks and meet only in the commutative node reductions.
void launch_p2g_cpic(const LaunchCfg& c, const DeviceData& d, int cur) {
    if (d.n == 0 || !d.has_bodies) return;
    if (d.bodies_react) {
        const int grid = c.num_sms * 7;
        if (c.dim == 2) k_p2g<2, true, true><<<grid, 32, 0, c.stream>>>(d, cur);
        else k_p2g<3, true, true><<<grid, 32, 0, c.stream>>>(d, cur);
    } else { // every body is immovable and at rest: no impulse can have an effect (rigid_impulses.wgsl:94-137)
        const int grid = c.num_sms * 9;
        if (c.dim == 2) k_p2g<2, true, false><<<grid, 32, 0, c.stream>>>(d, cur);
        else k_p2g<3, true, false><<<grid, 32, 0, c.stream>>>(d, cur);
    }
    ++*c.launch_counter;
}

} // namespace b2
