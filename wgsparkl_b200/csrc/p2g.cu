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
#include "cdf.cuh"
#include "launch.h"

namespace b2 {

#ifndef P2G_CHUNK_SIZE
#define P2G_CHUNK_SIZE 288
#endif
#ifndef P2G_WARPS_PLAIN
// resident single-warp CTAs per SM: no bodies / bodies / bodies that react. Eight, i.e. two per SM sub-partition:
// ptxas may then use 254 registers (no spills beside the 108 accumulators), which measured 1-3 % faster than 9-10
// warps at 168 registers; the shared memory that frees holds 288-particle chunks (fewer two-chunk half blocks in
// compressed material).
#define P2G_WARPS_PLAIN 8
#define P2G_WARPS_CPIC 8
#define P2G_WARPS_IMP 6
#endif
constexpr int P2G_CHUNK = P2G_CHUNK_SIZE; // particles staged per pass and warp: 32 cells x 9 (the reference seeds 8 per cell)

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
//   sp = (x, y, z, mass)   sv = (vx, vy, vz, CPIC affinity word of this substep)   sa = C[0..3]   sb = C[4..7]   sc = C[8]
// Slot i is stored at i ^ ((i >> 3) & 7): a cell's run starts at ~8 t for thread t, so consecutive lanes
// would otherwise hit the same bank group on every 16-byte load (8-way conflict).
__device__ __forceinline__ int p2g_swz(int i) { return i ^ ((i >> 3) & 7); }

// Walks the staged particles [lo, hi) of one cell.
//   MODE FAST / CPIC_MOMENTUM: acc[n] += w (affine dpt + m v, m)       (p2g.wgsl:188-230)
//   MODE CPIC_IMPULSE        : acc[n] += (delta_impulse, delta_ang)     (p2g.wgsl:203-225)
template <int D, int MODE, int W>
__device__ __forceinline__ bool p2g_accumulate(const DeviceData& d, int cur, uint32_t base, int lo, int hi,
                                               const float4* sp, const float4* sv, const float4* sa, const float4* sb,
                                               const float* sc, const float* cellpos, float h, float inv_h, int tb,
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
            pa = __float_as_uint(v4.w); // this substep's colour of the particle (the colouring phase left it here)
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
//   CPIC = false: scenes without bodies - every block takes the plain path.
//   CPIC = true : scenes with bodies - ONE kernel for all blocks. The collider-side blocks (k_scatter's compact list)
//     come FIRST in the work queue (they cost several times more per particle); for them the warp also runs the
//     "g2p_cdf" pass on the staged particles (cdf.cuh: colour = affinity / sign bits + MLS normal and distance, one
//     lane per particle) before it scatters them with the compatibility tests. All other blocks take the plain path.
//     Running both kinds in one persistent kernel matters: as separate kernels on a second stream the collider-side
//     work only got SM slots when the plain kernel's warps retired, i.e. it ran AFTER it (timeline: +21 us per substep
//     on the 1M cube, +120 us on the 2M dam slab with its four walls).
//   IMP (with CPIC): some body can react to impulses (DeviceData::bodies_react) - without it the per-node impulse
//   accumulators (27 x 6 registers, 5 KB of shared memory) and the second pass are compiled out.
template <int D, bool CPIC, bool IMP>
__global__ void __launch_bounds__(32, CPIC ? (IMP ? P2G_WARPS_IMP : P2G_WARPS_CPIC) : P2G_WARPS_PLAIN) k_p2g(DeviceData d, int cur) {
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
    __shared__ uint32_t s_nbr[NA];
    __shared__ uint2 tcdf[CPIC ? TC : 1]; // node (affinities, closest_id) of the tile
    __shared__ float tdist[CPIC ? TC : 1]; // node distance
    __shared__ float timp[IMP ? TC * WI : 1];

    pdl_start();
    TL_BEGIN(d, B200MPM_KERNEL_P2G);
    const int lane = threadIdx.x;
    // Work queue: the collider-side half blocks (cpic_list) first, then p2g_list bucket by bucket (k_scatter: blocks that
    // hold particles and see no collider, by decreasing population - longest items first keeps the last round short).
    __shared__ uint32_t s_cum[P2G_BUCKETS + 1]; // first work index of every bucket
    __shared__ float s_mass[16]; // masses of the first 16 materials (saves a dependent global load per staged chunk)
    // tile node -> (which of the 2^D neighbour blocks) << 6 | node inside that block: the index arithmetic of the tile
    // staging and of the flush, done once per warp instead of once per node and work item
    __shared__ uint16_t s_node[TC];
    for (int n = lane; n < TC; n += 32) {
        const int x = n % T, y = (n / T) % T, z = n / (T * T);
        const int ox = x >= B, oy = y >= B, oz = z >= B;
        s_node[n] = (uint16_t)(((ox + 2 * oy + 4 * oz) << 6) | ((x - ox * B) + (y - oy * B) * B + (z - oz * B) * B * B));
        tile[n] = make_float4(0.f, 0.f, 0.f, 0.f); // (every flush leaves the tile zeroed for the next work item)
    }
    const bool mass_in_smem = d.num_materials <= 16u;
    if (mass_in_smem && lane < (int)d.num_materials) s_mass[lane] = d.materials[lane].mass;
    const uint32_t ncpic = CPIC ? d.counters->num_cpic_blocks * 2u : 0u;
    if (lane == 0) {
        uint32_t c = ncpic;
        for (uint32_t k = 0; k < P2G_BUCKETS; ++k) {
            s_cum[k] = c;
            c += d.counters->num_p2g[k] * 2u;
        }
        s_cum[P2G_BUCKETS] = c;
    }
    __syncwarp();
    const uint32_t nwork = s_cum[P2G_BUCKETS];
    const float h = d.sim->cell_width;
    const float inv_h = 1.0f / h;
    const uint32_t num_bodies = d.sim->num_bodies;
    uint32_t* work = &d.counters->work_p2g;
    const float4* __restrict__ pos4 = d.pos4[cur];
    const float4* __restrict__ vel4 = d.vel4[cur];
    const float4* __restrict__ Ca = d.Ca[cur];
    const float4* __restrict__ Cb = d.Cb[cur];
    const float* __restrict__ Cc = d.Cc[cur];

    // Stage one chunk of the sorted range into shared memory (gather through sorted_ids, a
    // near-identity permutation of the current buffers).
    // reload_colour: second (impulse) pass over a half block of several chunks - the colours come back from
    // cdf_aff[next], where the first pass left them.
    auto stage = [&](uint32_t base, int cn, bool reload_colour, bool ids_prefetched) {
        // ids go through shared memory so that the request loop below stays rolled (few live registers
        // next to the 108 accumulators); each lane only reads back the ids it wrote itself.
        if (ids_prefetched) {
            cp_async_wait_all(); // (requested while the previous item was flushed)
        } else {
#pragma unroll
            for (int j = 0; j < CHUNK / 32; ++j) {
                const int i = lane + j * 32;
                if (i < cn) s_ids[i] = __ldg(d.sorted_ids + base + i);
            }
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
        }
        cp_async_wait_all();
        // (x, y, z, material bits) -> (x, y, z, mass)
#pragma unroll 1
        for (int i = lane; i < cn; i += 32) {
            const int s = p2g_swz(i);
            const uint32_t mbits = __float_as_uint(sp[s].w);
            sp[s].w = mass_in_smem ? s_mass[mbits & MAT_ID_MASK] : __ldg(&d.materials[mbits & MAT_ID_MASK].mass);
            if (CPIC && reload_colour) sv[s].w = __uint_as_float(d.cdf_aff[cur ^ 1][base + i]);
        }
        __syncwarp();
    };
    // "g2p_cdf" pass (g2p_cdf.wgsl:39-63) on the staged chunk, one lane per particle: the colour goes to sv[].w for
    // the scatter below and, if it is not the default, to cdf_aff[next] / cdf_nd (by sorted slot) for G2P.
    auto colour = [&](uint32_t base, int cn, bool any_cdf) {
        // Two phases, because the full colouring (sign votes + a 4x4 MLS solve, ~1.5 k instructions) only concerns the
        // particles within reach of a collider: first every particle checks its stencil (27 loads) and the ones that
        // see a collider are COMPACTED into a list, then the warp works through the list with all lanes busy.
        uint32_t* const s_list = s_ids; // compacted chunk indices; reuses the id buffer (the requests have been issued,
                                        // and phase 2 re-reads the few ids it needs from sorted_ids)
        int count = 0;
        const auto aff_tile = [&](int n) { return tcdf[n].x; };
#pragma unroll 1
        for (int i0 = 0; i0 < cn; i0 += 32) {
            const int i = i0 + lane;
            bool near = false;
            if (i < cn) {
                const int s = p2g_swz(i);
                if (any_cdf) {
                    const float4 p4 = sp[s];
                    const float pp[3] = {p4.x, p4.y, p4.z};
                    near = cdf_stencil_affinity<D>(pp, h, inv_h, aff_tile) != 0u;
                }
                if (!near) sv[s].w = __uint_as_float(0u);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, near);
            if (near) s_list[count + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
            count += __popc(m);
        }
        __syncwarp();
#pragma unroll 1
        for (int j0 = 0; j0 < count; j0 += 32) {
            const int j = j0 + lane;
            if (j < count) {
                const int i = (int)s_list[j];
                const int s = p2g_swz(i);
                const float4 p4 = sp[s];
                const float pp[3] = {p4.x, p4.y, p4.z};
                float4 nd;
                const uint32_t prev = d.cdf_aff[cur][__ldg(d.sorted_ids + base + i)];
                const uint32_t aff = cdf_colour_particle<D>(pp, prev, num_bodies, h, inv_h, aff_tile, [&](int n) { return tdist[n]; }, nd);
                if (aff != 0u) {
                    d.cdf_aff[cur ^ 1][base + i] = aff; // (k_scatter wrote the default, 0)
                    d.cdf_nd[base + i] = nd;
                }
                sv[s].w = __uint_as_float(aff);
            }
        }
        __syncwarp();
    };

    // The work loop is software-pipelined by hand: every item starts with a chain of dependent L2 round trips (work
    // counter -> block list -> ranges / bins, ~2.5 us that a single-warp CTA cannot hide by itself), so the NEXT item's
    // chain is issued piecewise while the current item stages and scatters: its queue index is requested before the
    // staging, its block id after the staging has landed, its ranges after the first chunk.
    struct Meta {
        uint32_t first, last, start, end; // sorted range of the half block, and of this lane's cell
        int vx, vy, vz; // block coordinates
    };
    const auto lookup = [&](uint32_t w) -> uint32_t { // work index -> block
        if (CPIC && w < ncpic) return d.cpic_list[w >> 1];
        uint32_t k = 0;
#pragma unroll
        for (uint32_t q = 1; q < P2G_BUCKETS; ++q) k += (w >= s_cum[q]) ? 1u : 0u;
        return d.p2g_list[(size_t)k * d.capacity + ((w - s_cum[k]) >> 1)];
    };
    const auto load_meta = [&](uint32_t w, uint32_t b) -> Meta {
        const uint32_t half = w & 1u, cell = half * HALF + lane;
        // (the blocks' ranges follow each other in arbitrary order - k_block_prepare - so the end of the block's last
        // cell is the block's own end, not the next block's first bin)
        const uint2 range = d.block_range[b];
        const int4 vid = d.block_vid[b];
        Meta m;
        m.first = d.cell_start[b * CELLS_PER_BLOCK + half * HALF];
        m.last = half ? range.x + range.y : d.cell_start[b * CELLS_PER_BLOCK + HALF];
        m.start = d.cell_start[b * CELLS_PER_BLOCK + cell];
        m.end = (cell + 1u < (uint32_t)CELLS_PER_BLOCK) ? d.cell_start[b * CELLS_PER_BLOCK + cell + 1] : range.x + range.y;
        m.vx = vid.x, m.vy = vid.y, m.vz = vid.z;
        return m;
    };
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(work, 1u);
    w = __shfl_sync(0xffffffffu, w, 0);
    uint32_t b = (w < nwork) ? lookup(w) : 0u;
    Meta meta = {};
    if (w < nwork) meta = load_meta(w, b);
    bool ids_prefetched = false; // s_ids holds the ids of this item's first chunk already
    while (w < nwork) {
        uint32_t wn = 0; // (1) the next item's queue index: requested now, read after the staging
        if (lane == 0) wn = atomicAdd(work, 1u);
        uint32_t bn = 0;
        Meta meta_n = {};
        const uint32_t half = w & 1u; // (ncpic is even)
        const bool cpic_item = CPIC && w < ncpic; // (warp-uniform)
        const uint32_t cell = half * HALF + lane; // this lane's cell of the block
        const uint32_t first = meta.first, last = meta.last, start = meta.start, end = meta.end;
        const bool empty = first == last; // nothing to scatter from this half (warp-uniform)
        const int lx = cell & (B - 1), ly = (cell >> LB) & (B - 1), lz = (D == 3) ? (cell >> (2 * LB)) : 0;
        const int tb = lx + T * ly + T * T * lz;
        if (empty) { // (rare: just keep the pipeline going)
            wn = __shfl_sync(0xffffffffu, wn, 0);
            if (wn < nwork) {
                bn = lookup(wn);
                meta_n = load_meta(wn, bn);
            }
            w = wn, b = bn, meta = meta_n;
            ids_prefetched = false;
            continue;
        }
        __syncwarp(); // the previous work item's tile / s_nbr are no longer read
        if (lane < NA) s_nbr[lane] = d.nbr[b * NA + lane];
        __syncwarp();
        bool any_cdf = false; // the tile holds a coloured node
        if (cpic_item) {
            for (int n = lane; n < TC; n += 32) {
                const uint32_t where = s_node[n];
                uint32_t hn = s_nbr[where >> 6];
                uint2 c = make_uint2(0u, NONE);
                float dist = 0.0f;
                if (hn != NONE) {
                    uint4 g = d.node_cdf[hn * CELLS_PER_BLOCK + (where & 63u)];
                    c = make_uint2(g.z, g.x); // (affinities, closest_id)
                    dist = __uint_as_float(g.y);
                }
                tcdf[n] = c;
                tdist[n] = dist;
                any_cdf = any_cdf || (c.x != 0u);
                if (IMP) {
#pragma unroll
                    for (int k = 0; k < WI; ++k) timp[n * WI + k] = 0.0f;
                }
            }
            any_cdf = __any_sync(0xffffffffu, any_cdf);
            __syncwarp();
        }

        const float cellpos[3] = {(float)(meta.vx * B + lx) * h, (float)(meta.vy * B + ly) * h, (float)(meta.vz * B + lz) * h};
        bool incompatible = false;
        {
            P2GAcc<NBH, D + 1> acc;
            // One chunk: stage, colour (collider side only), scatter into the register accumulators. The first chunk -
            // for all but overfull half blocks the only one - is peeled so that the 3^D x (D+1) accumulators are not
            // live yet while the colouring (a 4x4 solve per particle) runs.
            auto scatter_chunk = [&](uint32_t base) {
                const int cn = (int)min((uint32_t)CHUNK, last - base);
                const int lo = (int)(max(start, base) - base);
                const int hi = (int)(min(end, base + (uint32_t)cn) - base);
                if (cpic_item && any_cdf)
                    incompatible |= p2g_accumulate<D, P2G_CPIC_MOMENTUM, D + 1>(d, cur, base, lo, hi, sp, sv, sa, sb, sc, cellpos, h,
                                                                               inv_h, tb, tcdf, acc);
                else // no coloured node in reach: every particle keeps the default colour k_scatter wrote
                    p2g_accumulate<D, P2G_FAST, D + 1>(d, cur, base, lo, hi, sp, sv, sa, sb, sc, cellpos, h, inv_h, tb, tcdf, acc);
            };
            stage(first, (int)min((uint32_t)CHUNK, last - first), false, ids_prefetched);
            wn = __shfl_sync(0xffffffffu, wn, 0); // (2) arrived while the chunk was staged: request the next block id
            if (wn < nwork) bn = lookup(wn);
            if (cpic_item && any_cdf) colour(first, (int)min((uint32_t)CHUNK, last - first), true);
            acc.clear();
            scatter_chunk(first);
            if (wn < nwork) meta_n = load_meta(wn, bn); // (3) ... and, one chunk later, its ranges
            for (uint32_t base = first + CHUNK; base < last; base += CHUNK) {
                const int cn = (int)min((uint32_t)CHUNK, last - base);
                __syncwarp(); // the previous chunk is no longer in use
                stage(base, cn, false, false);
                if (cpic_item && any_cdf) colour(base, cn, true);
                scatter_chunk(base);
            }
            // Merge the per-cell stencils into the tile: 3^D conflict-free phases (within a phase the 32 lanes
            // add to 32 distinct nodes). In 3D the half block is two cells thick in z, so the phases sz = 0 and sz = 2
            // touch disjoint node layers and go as ONE step (two independent load-add-store chains, one barrier).
            auto merge_phase = [&](int sx, int sy, int sz) {
                const int n = sx + 3 * sy + 9 * sz;
                const int idx = tb + sx + T * sy + T * T * sz;
                float4 c = tile[idx];
                c.x += acc.a[n][0];
                c.y += acc.a[n][1];
                c.z += acc.a[n][2];
                if (D == 3) c.w += acc.a[n][D];
                tile[idx] = c;
            };
#pragma unroll
            for (int sy = 0; sy < 3; ++sy)
#pragma unroll
                for (int sx = 0; sx < 3; ++sx) {
                    if (D == 3) {
                        const int n0 = sx + 3 * sy, n2 = n0 + 18;
                        const int i0 = tb + sx + T * sy, i2 = i0 + 2 * T * T;
                        float4 c0 = tile[i0], c2 = tile[i2];
                        c0.x += acc.a[n0][0], c0.y += acc.a[n0][1], c0.z += acc.a[n0][2], c0.w += acc.a[n0][D];
                        c2.x += acc.a[n2][0], c2.y += acc.a[n2][1], c2.z += acc.a[n2][2], c2.w += acc.a[n2][D];
                        tile[i0] = c0;
                        tile[i2] = c2;
                    } else {
                        merge_phase(sx, sy, 0);
                    }
                    __syncwarp();
                }
            if (D == 3) {
#pragma unroll
                for (int sy = 0; sy < 3; ++sy)
#pragma unroll
                    for (int sx = 0; sx < 3; ++sx) {
                        merge_phase(sx, sy, 1);
                        __syncwarp();
                    }
            }
        }
        if (IMP && cpic_item) {
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
                        stage(base, cn, true, false);
                    }
                    const int lo = (int)(max(start, base) - base);
                    const int hi = (int)(min(end, base + (uint32_t)cn) - base);
                    p2g_accumulate<D, P2G_CPIC_IMPULSE, WI>(d, cur, base, lo, hi, sp, sv, sa, sb, sc, cellpos, h, inv_h, tb, tcdf, imp);
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

        // (4) the next item's particle ids, requested before this item's flush: they land while the tile is written out
        // and the next tile is set up (the id buffer is free: every chunk of this item has been staged and coloured)
        const bool prefetch_ids = wn < nwork && meta_n.first != meta_n.last;
        if (prefetch_ids) {
            const int cn_n = (int)min((uint32_t)CHUNK, meta_n.last - meta_n.first);
#pragma unroll
            for (int j = 0; j < CHUNK / 32; ++j) {
                const int i = lane + j * 32;
                if (i < cn_n) cp_async4(&s_ids[i], d.sorted_ids + meta_n.first + i);
            }
        }
        // Flush the tile: one 16-byte reduction per touched node.
        for (int n = lane; n < TC; n += 32) {
            const uint32_t where = s_node[n];
            const uint32_t hn = s_nbr[where >> 6];
            float4 c = tile[n];
            tile[n] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hn == NONE) continue; // only after a capacity overflow
            const uint32_t node = hn * CELLS_PER_BLOCK + (where & 63u);
            if (D == 2) c.w = 0.0f; // 2D stores (px, py, mass, 0)
            if (c.x != 0.0f || c.y != 0.0f || c.z != 0.0f || c.w != 0.0f) atomicAdd(d.node_mv + node, c);
            if (IMP && cpic_item) {
                // p2g.wgsl:142-155 converts the NODE's impulse to fixed point, once; this work item only holds a part
                // of it, so the parts meet in float (node_imp) and k_node_impulses does the conversion.
                if (tcdf[n].y != NONE) {
                    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}; // (lin.xyz, 0, ang.xyz, 0); 2D: (lin.xy, 0, 0, ang, 0, 0, 0)
                    bool any_lin = false, any_ang = false;
#pragma unroll
                    for (int k = 0; k < D; ++k) v[k] = timp[n * WI + k], any_lin |= v[k] != 0.0f;
#pragma unroll
                    for (int k = 0; k < WI - D; ++k) v[4 + k] = timp[n * WI + D + k], any_ang |= v[4 + k] != 0.0f;
                    if (any_lin) atomicAdd(d.node_imp + 2 * node, make_float4(v[0], v[1], v[2], 0.0f));
                    if (any_ang) atomicAdd(d.node_imp + 2 * node + 1, make_float4(v[4], v[5], v[6], 0.0f));
                }
            }
        }
        w = wn, b = bn, meta = meta_n;
        ids_prefetched = prefetch_ids;
    }
    TL_END(d, B200MPM_KERNEL_P2G);
}

// Apply the impulse to the closest body (p2g.wgsl:142-155): IntegerImpulseAtomic, i32(x * 1e5) of the node's total.
// One warp per block that holds a coloured node (block_f0), two nodes per lane; the fixed-point values are summed
// per CTA in shared memory first (integer sums are exact in any order) and node_imp is left cleared.
constexpr int IMPULSE_THREADS = 256;
__global__ void __launch_bounds__(IMPULSE_THREADS) k_node_impulses(DeviceData d) {
    __shared__ int s_imp[B200MPM_MAX_BODIES][6];
    for (int i = threadIdx.x; i < (int)B200MPM_MAX_BODIES * 6; i += blockDim.x) (&s_imp[0][0])[i] = 0;
    __syncthreads();
    const uint32_t nb = min(d.counters->num_active_blocks, d.capacity);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < nb; b += warps) {
        if (!d.block_f0[b]) continue;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t node = b * CELLS_PER_BLOCK + lane + 32 * k;
            const uint32_t cid = d.node_cdf[node].x;
            if (cid >= B200MPM_MAX_BODIES) continue; // NONE: nobody scattered an impulse here
            const float4 lin = d.node_imp[2 * node], ang = d.node_imp[2 * node + 1];
            const float v[6] = {lin.x, lin.y, lin.z, ang.x, ang.y, ang.z};
            bool any = false;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                if (v[c] != 0.0f) {
                    atomicAdd(&s_imp[cid][c], flt2int(v[c]));
                    any = true;
                }
            }
            if (any) {
                d.node_imp[2 * node] = make_float4(0.f, 0.f, 0.f, 0.f);
                d.node_imp[2 * node + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (int)B200MPM_MAX_BODIES * 6; i += blockDim.x) {
        const int v = (&s_imp[0][0])[i];
        if (v == 0) continue;
        BodyDev& body = d.bodies[i / 6];
        const int c = i % 6;
        atomicAdd(c < 3 ? &body.imp_lin[c] : &body.imp_ang[c - 3], v);
    }
}

void launch_p2g(const LaunchCfg& c, const DeviceData& d, int cur) {
    if (d.n == 0) return;
    // single-warp CTAs, as many per SM as the shared memory allows (22 KB each without bodies, 24.5 KB with, 29.6 KB
    // with impulse accumulators)
    if (!d.has_bodies) {
        const int grid = c.num_sms * P2G_WARPS_PLAIN;
        if (c.dim == 2) launch_pdl(k_p2g<2, false, false>, grid, 32, 0, c.stream, d, cur);
        else launch_pdl(k_p2g<3, false, false>, grid, 32, 0, c.stream, d, cur);
    } else if (d.bodies_react) {
        const int grid = c.num_sms * P2G_WARPS_IMP;
        if (c.dim == 2) launch_pdl(k_p2g<2, true, true>, grid, 32, 0, c.stream, d, cur);
        else launch_pdl(k_p2g<3, true, true>, grid, 32, 0, c.stream, d, cur);
        k_node_impulses<<<c.num_sms * (2048 / IMPULSE_THREADS), IMPULSE_THREADS, 0, c.stream>>>(d);
        ++*c.launch_counter;
    } else { // every body is immovable and at rest: no impulse can have an effect (rigid_impulses.wgsl:94-137)
        const int grid = c.num_sms * P2G_WARPS_CPIC;
        if (c.dim == 2) launch_pdl(k_p2g<2, true, false>, grid, 32, 0, c.stream, d, cur);
        else launch_pdl(k_p2g<3, true, false>, grid, 32, 0, c.stream, d, cur);
    }
    ++*c.launch_counter;
}

} // namespace b2
