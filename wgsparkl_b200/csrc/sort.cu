// "grid sort" pass: sparse-grid block activation + counting sort of particles by (block, cell).
//
// Reference: WgGrid::queue_sort (src/grid/grid.rs:30-207) = reset_hmap, touch_particle_blocks,
// init_indirect_workgroups, update_block_particle_count, copy_particles_len_to_scan_value,
// WgPrefixSum, copy_scan_values_to_first_particles, reset, finalize_particles_sort
// (src/grid/grid.wgsl, src/grid/sort.wgsl, src/grid/prefix_sum.wgsl).
//
// B200 design: the sort key is (block header id, cell-in-block), i.e. one bin per
// grid cell, so that P2G can walk each cell's particles as a contiguous run (the reference
// keeps per-node linked lists for that, sort.wgsl:129-135). One histogram + one single-pass
// decoupled-look-back scan over B*64+1 bins (B is device-resident) replace the reference's
// capacity-length Blelloch scan; the rank returned by the histogram atomic makes the final
// scatter atomic-free. The streaming kernels carry 4 particles per thread (memory-level parallelism).
// Also here: the rigid-particle kernels of mesh colliders (transform, block activation, p2g_cdf as a scatter).
#include "launch.h"
#include "bodies.cuh"

namespace b2 {

constexpr int SORT_THREADS = 256;
// Large tiles: the scan covers ~200 k bins per million particles, so with 8192-bin tiles every tile finds all its
// predecessors in ONE look-back round of 32 (the kernel is a latency chain, not a bandwidth problem).
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

uint32_t scan_num_tiles(uint64_t len) { return (uint32_t)((len + SCAN_TILE - 1) / SCAN_TILE); }

// ---- insertion_index + mark_block_as_active (grid.wgsl:121-164, 323-334) -------------------------
// Claims the hash slot of a block. `won` reports that THIS call created the entry, in which case the caller owes
// the block its dense index (hvals[slot], block_vid[index]); see k_touch for how those are handed out.
template <int D>
__device__ __forceinline__ uint32_t claim_block(const DeviceData& d, int bx, int by, int bz, bool& won) {
    const uint32_t mask = d.capacity - 1;
    const uint32_t key = pack_key<D>(bx, by, bz);
    uint32_t slot = hash_key(key) & mask;
    won = false;
    for (uint32_t k = 0; k <= mask; ++k) {
        uint32_t cur = *((volatile uint32_t*)(d.hkeys + slot));
        if (cur == key) return slot;
        if (cur == NONE) {
            uint32_t prev = atomicCAS(d.hkeys + slot, NONE, key);
            if (prev == NONE) {
                won = true;
                return slot;
            }
            if (prev == key) return slot;
        }
        slot = (slot + 1) & mask;
    }
    d.counters->overflow = 1u; // table full: the block is dropped (silent in the reference, grid.wgsl:126-128)
    return NONE;
}

// hvals[slot] is NONE from k_begin_substep until the block's creator publishes the dense index here (k_touch waits on
// exactly that); a block beyond the capacity is published as HVAL_DROPPED, which every lookup treats as "absent".
__device__ __forceinline__ void publish_block(const DeviceData& d, uint32_t slot, uint32_t hid, int4 vid) {
    if (hid < d.capacity) d.block_vid[hid] = vid;
    *((volatile uint32_t*)(d.hvals + slot)) = (hid < d.capacity) ? hid : HVAL_DROPPED;
}

// ---- touch_particle_blocks (sort.wgsl:26-36) ------------------------------------------------------
// These three streaming kernels (touch, count, scatter) are chains of dependent long-latency operations per
// particle (load -> hash probe -> atomic -> store). With one particle per thread the 64 warps an SM can hold
// do not cover that latency (ncu: >90 % of warp time in long_scoreboard at < 1 TB/s), so every thread carries
// SORT_ITEMS particles whose loads / probes / atomics are issued back to back: a warp owns 32 * SORT_ITEMS
// CONSECUTIVE particles (item j of lane l is particle warp_base + 32 j + l, so every access stays coalesced).
//
// k_touch: consecutive particles that map to the same block (the common case: the buffers are already in
// last substep's sorted order) form a run; the run leaders' 2^D insertions of blocks_associated_to_block
// (grid.wgsl:300-320) are done by ALL lanes of the warp at once, 32 / 2^D leaders per round.
//
// Dense block indices come from ONE counter (num_active_blocks); every substep re-creates ~20 k blocks per
// million particles, so the blocks a warp creates in one round share a single atomic on it.
#ifndef TOUCH_ITEMS
#define TOUCH_ITEMS 4
#endif
#ifndef SCATTER_ITEMS
#define SCATTER_ITEMS 4
#endif
#ifndef TOUCH_MIN_CTAS
#define TOUCH_MIN_CTAS 8 // resident CTAs per SM that k_touch is compiled for (register cap)
#endif
#ifndef SCATTER_MIN_CTAS
#define SCATTER_MIN_CTAS 1
#endif
constexpr int SORT_ITEMS = TOUCH_ITEMS;
constexpr int SORT_PER_WARP = 32 * SORT_ITEMS;
constexpr int SORT_PER_CTA = SORT_THREADS * SORT_ITEMS;
constexpr int SCATTER_PER_WARP = 32 * SCATTER_ITEMS;

template <int D>
__global__ void __launch_bounds__(SORT_THREADS, TOUCH_MIN_CTAS) k_touch(DeviceData d, int cur, int integrate) {
    pdl_start();
    // `integrate`: the deferred body integration of the PREVIOUS substep (api.cu, enqueue_substep) rides along as an
    // extra first CTA - nothing in this kernel looks at the bodies, k_block_prepare is the first that does. As a
    // kernel of its own on a side branch it could not start before this kernel's single wave had drained.
    constexpr int NA = Dim<D>::NASSOC, WARPS = SORT_THREADS / 32, PER_ROUND = 32 / NA;
    __shared__ int4 s_lead[WARPS][SORT_PER_WARP]; // run leaders of a warp: block (x, y, z), then the slot in .w
    static_assert(sizeof(s_lead) >= sizeof(BodyDev) * B200MPM_MAX_BODIES, "the integration stages the bodies in s_lead");
    if (integrate) {
        if (blockIdx.x == 0) {
            if (threadIdx.x < 32) integrate_bodies_warp<D>(d, threadIdx.x, reinterpret_cast<BodyDev*>(&s_lead[0][0]));
            return;
        }
    }
    const uint32_t cta = blockIdx.x - (integrate ? 1u : 0u);
    TL_BEGIN(d, B200MPM_KERNEL_TOUCH);
    if (cta == 0 && threadIdx.x == 0) {
        // Housekeeping for the kernels that follow (nobody reads these before k_block_alloc / k_scatter, and the readers
        // of the last substep are long done): the work-list counters.
        Counters* c = d.counters;
        c->scan_ticket = 0;
        c->work_p2g = 0;
        c->work_p2g_cpic = 0;
        c->work_g2p = 0;
        c->work_cdf = 0;
        c->num_cpic_blocks = 0;
        c->num_g2p_items = 0;
        c->num_g2p_back = 0;
        for (uint32_t k = 0; k < P2G_BUCKETS; ++k) c->num_p2g[k] = 0;
        c->dropped_particles = 0;
        c->sorted_total = 0; // (bumped by k_block_alloc; the sharded tick has read the previous value already)
    }
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t warp_base = (cta * WARPS + warp) * SORT_PER_WARP;
    const uint32_t n_live = d.counters->n_live;
    const float h = d.sim->cell_width;
    const float inv_h = 1.0f / h;
    const float4* __restrict__ pos4 = d.pos4[cur];

    float4 p[SORT_ITEMS];
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        const uint32_t i = warp_base + j * 32 + lane;
        p[j] = (i < n_live) ? pos4[i] : make_float4(0.f, 0.f, 0.f, __uint_as_float(FLAG_DEAD));
    }
    uint32_t cell[SORT_ITEMS], lmask[SORT_ITEMS];
    bool alive[SORT_ITEMS];
    uint32_t before = 0; // leaders in items < j
    uint32_t last_key = 0;
    bool last_alive = false;
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        // emigrated particles (k_emigrate) and the tail beyond n_live are not sorted: parked, then dropped
        alive[j] = (__float_as_uint(p[j].w) & FLAG_DEAD) == 0u;
        const int cx = assoc_cell(p[j].x, h, inv_h), cy = assoc_cell(p[j].y, h, inv_h);
        const int cz = (D == 3) ? assoc_cell(p[j].z, h, inv_h) : 0;
        const int bx = cx >> Dim<D>::LOG_BLOCK, by = cy >> Dim<D>::LOG_BLOCK, bz = cz >> Dim<D>::LOG_BLOCK; // grid.wgsl:286
        cell[j] = (cx & (Dim<D>::BLOCK - 1)) + (cy & (Dim<D>::BLOCK - 1)) * Dim<D>::BLOCK +
                  ((D == 3) ? (cz & (Dim<D>::BLOCK - 1)) * Dim<D>::BLOCK * Dim<D>::BLOCK : 0);
        const uint32_t key = pack_key<D>(bx, by, bz);
        // run leader: alive, and the previous particle (lane - 1, or lane 31 of the previous item) is dead or elsewhere
        uint32_t pk = __shfl_up_sync(0xffffffffu, key, 1);
        bool pa = __shfl_up_sync(0xffffffffu, (int)alive[j], 1) != 0;
        if (lane == 0) pk = last_key, pa = last_alive;
        const bool leader = alive[j] && (!pa || pk != key);
        lmask[j] = __ballot_sync(0xffffffffu, leader);
        if (leader) s_lead[warp][before + __popc(lmask[j] & ((1u << lane) - 1u))] = make_int4(bx, by, bz, (int)NONE);
        before += __popc(lmask[j]);
        last_key = __shfl_sync(0xffffffffu, key, 31);
        last_alive = __shfl_sync(0xffffffffu, (int)alive[j], 31) != 0;
    }
    const uint32_t num_leaders = before;
    __syncwarp();
    for (uint32_t q0 = 0; q0 < num_leaders; q0 += PER_ROUND) {
        const uint32_t q = q0 + lane / NA, o = lane % NA;
        if (q < num_leaders) {
            const int4 L = s_lead[warp][q];
            const int bx = L.x + (int)(o & 1), by = L.y + (int)((o >> 1) & 1), bz = L.z + ((D == 3) ? (int)((o >> 2) & 1) : 0);
            bool won;
            const uint32_t slot = claim_block<D>(d, bx, by, bz, won);
            // the blocks this round created take their dense indices with ONE atomic per warp
            const uint32_t active_lanes = __activemask();
            const uint32_t wmask = __ballot_sync(active_lanes, won);
            if (wmask) {
                const int first = __ffs(wmask) - 1;
                uint32_t base = 0;
                if ((int)lane == first) base = atomicAdd(&d.counters->num_active_blocks, (uint32_t)__popc(wmask));
                base = __shfl_sync(active_lanes, base, first);
                if (won) publish_block(d, slot, base + __popc(wmask & ((1u << lane) - 1u)), make_int4(bx, by, bz, 0));
            }
            if (o == 0) s_lead[warp][q].w = (int)slot; // offset (0,..,0): the slot that identifies the run's block
        }
    }
    __syncwarp();
    // update_block_particle_count (sort.wgsl:89-99), one bin per cell, in the same kernel: the run's dense block index
    // is there as soon as the block's creator has published it. Every publish of THIS warp is behind us, so the wait
    // below only ever depends on warps that are already running and never wait before their own publishes.
    for (uint32_t q = lane; q < num_leaders; q += 32) {
        const uint32_t slot = (uint32_t)s_lead[warp][q].w;
        uint32_t hid = NONE;
        if (slot != NONE) {
            const volatile uint32_t* hv = d.hvals + slot;
            while ((hid = *hv) == NONE) {}
            if (hid >= d.capacity) hid = NONE;
        }
        s_lead[warp][q].w = (int)hid;
    }
    __syncwarp();
    uint32_t ck[SORT_ITEMS];
    before = 0;
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        const uint32_t upto = before + __popc(lmask[j] & (0xffffffffu >> (31 - lane))); // leaders at or before me
        before += __popc(lmask[j]);
        ck[j] = NONE;
        if (alive[j]) {
            const uint32_t hid = (uint32_t)s_lead[warp][upto - 1].w;
            if (hid != NONE) ck[j] = hid * CELLS_PER_BLOCK + cell[j];
        }
    }
    // Runs of consecutive lanes in the same cell (the buffers are nearly sorted already) share one atomic.
    uint32_t base[SORT_ITEMS];
    int lead[SORT_ITEMS];
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        const uint32_t prev = __shfl_up_sync(0xffffffffu, ck[j], 1);
        const bool leader = (lane == 0) || (prev != ck[j]);
        const uint32_t leaders = __ballot_sync(0xffffffffu, leader);
        const uint32_t below = leaders & (0xffffffffu >> (31 - lane));
        const int L = 31 - __clz(below);
        const uint32_t above = (L == 31) ? 0u : (leaders >> (L + 1));
        const int run = above ? __ffs(above) : (32 - L);
        lead[j] = L;
        base[j] = (leader && ck[j] != NONE) ? atomicAdd(d.cell_start + ck[j], (uint32_t)run) : 0u;
    }
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        const uint32_t i = warp_base + j * 32 + lane;
        const uint32_t b = __shfl_sync(0xffffffffu, base[j], lead[j]);
        if (i < n_live) {
            d.pkey[i] = ck[j];
            d.rank[i] = b + (lane - (uint32_t)lead[j]);
        }
    }
    TL_END(d, B200MPM_KERNEL_TOUCH);
}

// ---- rigid particles (sample points of trimesh / polyline colliders) ---------------------------------------
// transform_sample_points + transform_shape_points (rigid_particle_update.wgsl:26-50): world = R local + t with
// the pose of the point's collider.
template <int D>
__global__ void k_transform_rigid(DeviceData d) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool vertex = i < d.num_mesh_verts;
    const uint32_t j = vertex ? i : i - d.num_mesh_verts;
    if (!vertex && j >= d.num_rigid) return;
    const float4 l = vertex ? d.mv_local[j] : d.rp_local[j];
    const BodyDev& b = d.bodies[vertex ? d.mv_body[j] : d.rp_ids[j].w];
    const float lp[3] = {l.x, l.y, l.z};
    float w[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < D; ++r) { // (contraction-free: the vertices feed the integer-valued colouring of k_p2g_cdf)
        float s = __fmul_rn(b.rot[r], lp[0]);
#pragma unroll
        for (int k = 1; k < D; ++k) s = __fadd_rn(s, __fmul_rn(b.rot[k * D + r], lp[k]));
        w[r] = __fadd_rn(s, b.trans[r]);
    }
    (vertex ? d.mv_world : d.rp_world)[j] = make_float4(w[0], w[1], w[2], 0.f);
}

template <int D>
__device__ __forceinline__ void rigid_block(const DeviceData& d, uint32_t i, int& bx, int& by, int& bz, uint32_t& cell) {
    const float h = d.sim->cell_width, inv_h = 1.0f / h;
    const float4 p = d.rp_world[i];
    const int cx = assoc_cell(p.x, h, inv_h), cy = assoc_cell(p.y, h, inv_h), cz = (D == 3) ? assoc_cell(p.z, h, inv_h) : 0;
    bx = cx >> Dim<D>::LOG_BLOCK, by = cy >> Dim<D>::LOG_BLOCK, bz = cz >> Dim<D>::LOG_BLOCK;
    cell = (cx & (Dim<D>::BLOCK - 1)) + (cy & (Dim<D>::BLOCK - 1)) * Dim<D>::BLOCK +
           ((D == 3) ? (cz & (Dim<D>::BLOCK - 1)) * Dim<D>::BLOCK * Dim<D>::BLOCK : 0);
}

// mark_rigid_particles_needing_block (sort.wgsl:54-86): a sample point asks for its own block iff that block is
// missing while another block its stencil reaches exists. Evaluated against the table as touch_particle_blocks
// left it - hence a kernel of its own, before any sample point inserts anything.
template <int D>
__global__ void k_mark_rigid(DeviceData d) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.num_rigid) return;
    int bx, by, bz;
    uint32_t cell;
    rigid_block<D>(d, i, bx, by, bz, cell);
    const uint32_t mask = d.capacity - 1;
    uint32_t needs = 0u;
    if (find_block(d.hkeys, d.hvals, mask, pack_key<D>(bx, by, bz)) == NONE) {
        for (int o = 1; o < Dim<D>::NASSOC && !needs; ++o) {
            const int ox = o & 1, oy = (o >> 1) & 1, oz = (D == 3) ? (o >> 2) & 1 : 0;
            needs = find_block(d.hkeys, d.hvals, mask, pack_key<D>(bx + ox, by + oy, bz + oz)) != NONE;
        }
    }
    d.rp_needs_block[i] = needs;
}

// touch_rigid_particle_blocks (sort.wgsl:38-52)
template <int D>
__global__ void k_touch_rigid(DeviceData d) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.num_rigid || !d.rp_needs_block[i]) return;
    int bx, by, bz;
    uint32_t cell;
    rigid_block<D>(d, i, bx, by, bz, cell);
    bool won;
    const uint32_t slot = claim_block<D>(d, bx, by, bz, won);
    if (won) publish_block(d, slot, atomicAdd(&d.counters->num_active_blocks, 1u), make_int4(bx, by, bz, 0));
}

// p2g_cdf (p2g_cdf.wgsl:51-190). The reference sorts the sample points into per-cell linked lists and lets every
// node GATHER the primitives of the 3^D cells around it. Here every sample point SCATTERS: it projects the 3^D nodes
// of its stencil onto its own primitive and, where the projection falls inside the primitive, ORs the collider's
// affinity / sign bits into the node and lowers the node's (distance, closest collider) pair with one 64-bit
// atomicMin - no lists, no sort of the sample points, and the result does not depend on any order (the reference
// breaks distance ties by list order; here the smaller collider index wins).
template <int D>
__global__ void k_p2g_cdf(DeviceData d) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.num_rigid) return;
    int bx, by, bz;
    uint32_t cell;
    rigid_block<D>(d, i, bx, by, bz, cell);
    const uint32_t mask = d.capacity - 1;
    // a sample point whose own block is inactive is in no list (sort.wgsl:149-152)
    uint32_t hid[Dim<D>::NASSOC];
    hid[0] = find_block(d.hkeys, d.hvals, mask, pack_key<D>(bx, by, bz));
    if (hid[0] == NONE || hid[0] >= d.capacity) return;
#pragma unroll
    for (int o = 1; o < Dim<D>::NASSOC; ++o) {
        const int ox = o & 1, oy = (o >> 1) & 1, oz = (D == 3) ? (o >> 2) & 1 : 0;
        hid[o] = find_block(d.hkeys, d.hvals, mask, pack_key<D>(bx + ox, by + oy, bz + oz));
        if (hid[o] >= d.capacity) hid[o] = NONE;
    }
    const float h = d.sim->cell_width;
    const uint4 ids = d.rp_ids[i];
    const uint32_t collider = ids.w;
    const float4 A = d.mv_world[ids.x], Bv = d.mv_world[ids.y];
    const float4 C = (D == 3) ? d.mv_world[ids.z] : make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int B = Dim<D>::BLOCK, LB = Dim<D>::LOG_BLOCK;
    const int lx = cell & (B - 1), ly = (cell >> LB) & (B - 1), lz = (D == 3) ? (cell >> (2 * LB)) : 0;
    for (int n = 0; n < Dim<D>::NBH; ++n) {
        const int sx = n % 3, sy = (n / 3) % 3, sz = n / 9;
        const int nx = lx + sx, ny = ly + sy, nz = lz + sz; // node, in cells relative to the own block's origin
        const int o = (nx >= B) + 2 * (ny >= B) + ((D == 3) ? 4 * (nz >= B) : 0);
        if (hid[o] == NONE) continue;
        const uint32_t node = hid[o] * CELLS_PER_BLOCK + (nx & (B - 1)) + (ny & (B - 1)) * B + ((D == 3) ? (nz & (B - 1)) * B * B : 0);
        const float px = (float)(bx * B + nx) * h, py = (float)(by * B + ny) * h, pz = (float)(bz * B + nz) * h;
        float distance;
        bool sign;
        if (D == 2) {
            // wgparry Segment::projectLocalPoint, then p2g_cdf.wgsl:143-158
            // (contraction-free arithmetic: the outcome is a set of affinity / sign BITS)
            const float abx = Bv.x - A.x, aby = Bv.y - A.y, apx = px - A.x, apy = py - A.y;
            const float ab_ap = __fadd_rn(__fmul_rn(abx, apx), __fmul_rn(aby, apy));
            const float sqnab = __fadd_rn(__fmul_rn(abx, abx), __fmul_rn(aby, aby));
            if (ab_ap <= 0.0f || ab_ap >= sqnab) continue; // projects on an end point
            const float t = __fdiv_rn(ab_ap, sqnab);
            const float qx = __fadd_rn(A.x, __fmul_rn(abx, t)), qy = __fadd_rn(A.y, __fmul_rn(aby, t));
            if ((qx == A.x && qy == A.y) || (qx == Bv.x && qy == Bv.y)) continue;
            const float dx = px - qx, dy = py - qy;
            distance = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
            sign = __fadd_rn(__fmul_rn(dx, -aby), __fmul_rn(dy, abx)) < 0.0f;
        } else {
            // p2g_cdf.wgsl:160-188: projection on the face interior only
            const V3 a = v3(A.x, A.y, A.z), b = v3(Bv.x, Bv.y, Bv.z), c = v3(C.x, C.y, C.z), pt = v3(px, py, pz);
            const V3 ap = pt - a, bp = pt - b, cp = pt - c, ab = b - a, ac = c - a, bc = c - b;
            // (contraction-free arithmetic: the outcome is a set of affinity / sign BITS)
            const V3 nrm = cross_rn(ab, ac);
            const float n_length = sqrtf(dot_rn(nrm, nrm));
            if (!(n_length != 0.0f && dot_rn(cross_rn(ab, nrm), ap) <= 0.0f && dot_rn(cross_rn(bc, nrm), bp) <= 0.0f &&
                  dot_rn(cross_rn(ac, nrm), cp) >= 0.0f))
                continue;
            const float signed_dist = __fdiv_rn(dot_rn(nrm, ap), n_length);
            sign = signed_dist < 0.0f;
            distance = fabsf(signed_dist);
        }
        uint4* cdf = d.node_cdf + node;
        atomicOr(&cdf->z, (1u << collider) | ((uint32_t)sign << (collider + 16)));
        atomicMin((unsigned long long*)cdf, ((unsigned long long)__float_as_uint(distance) << 32) | collider);
        d.block_f0[hid[o]] = 1; // the block holds a coloured node (benign race: everybody stores 1)
    }
}

// ---- exclusive scan, single pass with decoupled look-back (replaces prefix_sum.wgsl) -----------------
// Same result as WgPrefixSum::eval_cpu (prefix_sum.rs:71-83): out[i] = sum_{j<i} in[j].
__global__ void __launch_bounds__(SCAN_THREADS) k_scan(uint32_t* __restrict__ data, uint32_t len_value,
                                                       const Counters* __restrict__ counters, uint32_t capacity,
                                                       uint64_t* state, uint32_t* ticket) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ uint32_t s_prefix;
    pdl_start();
    const uint32_t len = counters ? (min(counters->num_active_blocks, capacity) * CELLS_PER_BLOCK + 1) : len_value;
    // The grid is sized for the worst case (capacity); CTAs beyond the live tiles leave without a ticket, so
    // the tickets handed out are exactly 0 .. ntiles-1 and the same-address atomic stays cheap.
    if (blockIdx.x * SCAN_TILE >= len) return;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * SCAN_TILE;
    if (base >= len) return;
    const uint32_t t0 = base + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
    if (t0 + SCAN_ITEMS <= len && (reinterpret_cast<uintptr_t>(data) & 15u) == 0u) { // (t0 is a multiple of SCAN_ITEMS)
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; j += 4) {
            const uint4 q = *reinterpret_cast<const uint4*>(data + t0 + j);
            v[j] = q.x, v[j + 1] = q.y, v[j + 2] = q.z, v[j + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) v[j] = (t0 + j < len) ? data[t0 + j] : 0u;
    }
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) sum += v[j];
    // inclusive warp scan of the per-thread sums
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t warp_off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        uint32_t x = s_warp[w];
        if (w < warp) warp_off += x;
        total += x;
    }
    if (warp == 0) { // decoupled look-back, 32 predecessors per round
        const uint64_t FLAG_A = 1ull << 62, FLAG_P = 2ull << 62;
        uint32_t running = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch((unsigned long long*)(state + 0), FLAG_P | (uint64_t)total);
        } else {
            if (lane == 0) atomicExch((unsigned long long*)(state + tile), FLAG_A | (uint64_t)total);
            int j0 = (int)tile - 1; // lane l looks at tile j0 - l
            while (true) {
                const int j = j0 - (int)lane;
                uint64_t s = (j >= 0) ? *((volatile uint64_t*)(state + j)) : FLAG_P; // virtual tile -1: prefix 0
                uint64_t flag = s >> 62;
                // wait until every inspected tile has published something
                if (__any_sync(0xffffffffu, flag == 0)) continue;
                const uint32_t pmask = __ballot_sync(0xffffffffu, flag == 2);
                const int stop = pmask ? (__ffs(pmask) - 1) : 32; // nearest tile with an inclusive prefix
                uint32_t v = ((int)lane <= stop) ? (uint32_t)s : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                running += v;
                if (pmask) break;
                j0 -= 32;
            }
            if (lane == 0) atomicExch((unsigned long long*)(state + tile), FLAG_P | (uint64_t)(running + total));
        }
        if (lane == 0) s_prefix = running;
    }
    __syncthreads();
    uint32_t excl = s_prefix + warp_off + (inc - sum);
    if (t0 + SCAN_ITEMS <= len && (reinterpret_cast<uintptr_t>(data) & 15u) == 0u) {
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; j += 4) {
            uint4 q;
            q.x = excl, excl += v[j];
            q.y = excl, excl += v[j + 1];
            q.z = excl, excl += v[j + 2];
            q.w = excl, excl += v[j + 3];
            *reinterpret_cast<uint4*>(data + t0 + j) = q;
        }
    } else {
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) {
            if (t0 + j < len) data[t0 + j] = excl;
            excl += v[j];
        }
    }
}

// ---- per-block preparation: neighbour table + reset (grid.wgsl:362-379) + grid_update_cdf ----------
}
#include "collide.cuh"
namespace b2 {

// One WARP per block (lane l owns cells / nodes l and l + 32), one block per warp in a single wave: everything a block
// needs between k_touch and k_scatter.
//  * copy_particles_len_to_scan_value + prefix sum + copy_scan_values_to_first_particles (sort.wgsl:101-115). The
//    reference scans the per-block counts over the whole capacity. All the sort needs, though, is that every block
//    owns a contiguous range of the sorted array - in WHICH order the blocks follow each other is immaterial (the
//    reference's own order is the atomic order of its header ids). So there is no scan across blocks at all: the warp
//    sums the block's 64 per-cell counts, takes the block's range with ONE atomicAdd on a bump counter (sorted_total)
//    and turns the counts into the cells' first slots in place. No look-back chain, no tile descriptors (k_scan
//    remains as the stand-alone b200mpm_prefix_sum_u32).
//  * the neighbour table, the node reset (grid.wgsl:362-379) and grid_update_cdf (grid_update_cdf.wgsl:16-39).
constexpr int PREPARE_THREADS = 256;
template <int D>
__global__ void __launch_bounds__(PREPARE_THREADS) k_block_prepare(DeviceData d) {
    pdl_start();
    TL_BEGIN(d, B200MPM_KERNEL_BLOCK_PREPARE);
    const uint32_t nb = min(d.counters->num_active_blocks, d.capacity);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const float h = d.sim->cell_width;
    const uint32_t num_bodies = d.sim->num_bodies;
    for (uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < nb; b += warps) {
        const int4 vid = d.block_vid[b];
        uint32_t* bins = d.cell_start + b * CELLS_PER_BLOCK;
        const uint32_t c0 = bins[lane], c1 = bins[32 + lane];
        if (lane < (uint32_t)Dim<D>::NASSOC) {
            const int ox = lane & 1, oy = (lane >> 1) & 1, oz = (D == 3) ? (lane >> 2) & 1 : 0;
            const uint32_t key = pack_key<D>(vid.x + ox, vid.y + oy, vid.z + oz);
            d.nbr[b * Dim<D>::NASSOC + lane] = (lane == 0) ? b : find_block(d.hkeys, d.hvals, d.capacity - 1, key);
        }
        uint32_t i0 = c0, i1 = c1; // inclusive scans of the two halves
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y0 = __shfl_up_sync(0xffffffffu, i0, o), y1 = __shfl_up_sync(0xffffffffu, i1, o);
            if ((int)lane >= o) i0 += y0, i1 += y1;
        }
        const uint32_t t0 = __shfl_sync(0xffffffffu, i0, 31), t1 = __shfl_sync(0xffffffffu, i1, 31);
        uint32_t base = 0;
        if (lane == 0 && t0 + t1 != 0u) base = atomicAdd(&d.counters->sorted_total, t0 + t1);
        base = __shfl_sync(0xffffffffu, base, 0);
        bins[lane] = base + i0 - c0;
        bins[32 + lane] = base + t0 + i1 - c1;
        if (lane == 0) d.block_range[b] = make_uint2(base, t0 + t1);
        d.node_mv[b * CELLS_PER_BLOCK + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        d.node_mv[b * CELLS_PER_BLOCK + 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d.has_bodies) { // grid_update_cdf.wgsl:16-39
            // lane i looks at body i once for the whole block; the nodes then only visit the bodies within reach
            const float origin[3] = {(float)(vid.x * Dim<D>::BLOCK) * h, (float)(vid.y * Dim<D>::BLOCK) * h,
                                     (float)(vid.z * Dim<D>::BLOCK) * h};
            const uint32_t body_mask =
                __ballot_sync(0xffffffffu, lane < num_bodies && body_may_touch_block<D>(d.bodies[lane], h, origin));
            bool any = false;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t t = lane + 32 * k;
                const int lx = t & (Dim<D>::BLOCK - 1), ly = (t >> Dim<D>::LOG_BLOCK) & (Dim<D>::BLOCK - 1),
                          lz = (D == 3) ? (t >> (2 * Dim<D>::LOG_BLOCK)) : 0;
                float pt[3] = {(float)(vid.x * Dim<D>::BLOCK + lx) * h, (float)(vid.y * Dim<D>::BLOCK + ly) * h,
                               (float)(vid.z * Dim<D>::BLOCK + lz) * h};
                const NodeCdf c = collide<D>(d.bodies, body_mask, h, pt);
                d.node_cdf[b * CELLS_PER_BLOCK + t] = make_uint4(c.closest_id, __float_as_uint(c.distance), c.affinities, 0u);
                any |= c.affinities != 0u;
            }
            any = __any_sync(0xffffffffu, any);
            if (lane == 0) d.block_f0[b] = any ? 1 : 0;
        }
    }
    TL_END(d, B200MPM_KERNEL_BLOCK_PREPARE);
}

// ---- finalize_particles_sort (sort.wgsl:117-137): atomic-free scatter ------------------------------
template <int D>
__global__ void __launch_bounds__(SORT_THREADS, SCATTER_MIN_CTAS) k_scatter(DeviceData d, int cur) {
    pdl_start();
    TL_BEGIN(d, B200MPM_KERNEL_SCATTER);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        d.counters->prev_active_blocks = min(d.counters->num_active_blocks, d.capacity); // for the host and the next clear
        d.counters->integrate_pending = 1u; // this substep's body impulses / poses are still to be integrated
    }
    {
        // Per-block work lists (i doubles as a block index here; the grid covers capacity blocks).
        const uint32_t nb = min(d.counters->num_active_blocks, d.capacity);
        uint32_t np = 0, first = 0;
        if (i < nb) {
            const uint2 range = d.block_range[i]; // (k_block_alloc)
            first = range.x, np = range.y;
        }
        uint32_t nbr[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) nbr[o] = NONE;
        int flag = 0;
        if (i < nb) {
#pragma unroll
            for (int o = 0; o < Dim<D>::NASSOC; ++o) nbr[o] = d.nbr[i * Dim<D>::NASSOC + o];
        }
        if (d.has_bodies && i < nb) {
            // CPIC: a block runs the collider-aware paths iff one of the 2^D blocks its tile overlaps has a node
            // near / inside a collider.
#pragma unroll
            for (int o = 0; o < Dim<D>::NASSOC; ++o)
                if (nbr[o] != NONE) flag |= d.block_f0[nbr[o]];
            d.block_flags[i] = (uint32_t)flag;
            if (flag && np != 0u) d.cpic_list[atomicAdd(&d.counters->num_cpic_blocks, 1u)] = i;
        }
        // Two-ended lists, one atomic per warp and list end: positions [0, front) grow upwards, the back grows
        // downwards from the end; the consumers walk front first, then the back from the end - slow items first.
        const uint32_t lane = threadIdx.x & 31;
        auto reserve = [&](uint32_t mine, uint32_t* counter) -> uint32_t {
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += v;
            }
            const uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t base = 0;
            if (warp_total != 0u && lane == 31) base = atomicAdd(counter, warp_total);
            return __shfl_sync(0xffffffffu, base, 31) + incl - mine;
        };
        // G2P items: blocks that hold particles, in parts of <= G2P_ITEM particles (one per thread of a k_g2p CTA).
        // Collider-side blocks cost more per particle (ghost-velocity gather), so they go to the FRONT of the
        // list and are scheduled first; everything else fills the list from the back.
        const uint32_t parts = g2p_parts(np);
        const bool slow = flag != 0;
        const uint32_t base_f = reserve(slow ? parts : 0u, &d.counters->num_g2p_items);
        const uint32_t base_b = reserve(slow ? 0u : parts, &d.counters->num_g2p_back);
        if (parts) {
            const uint32_t per = (np + parts - 1) / parts;
            for (uint32_t p = 0; p < parts; ++p) {
                const uint32_t pos = slow ? base_f + p : d.g2p_items_len - 1u - (base_b + p);
                uint4* dst = (uint4*)(d.g2p_items + pos);
                dst[0] = make_uint4(i, first + p * per, min(per, np - p * per), (uint32_t)flag);
                dst[1] = make_uint4(nbr[0], nbr[1], nbr[2], nbr[3]);
                dst[2] = make_uint4(nbr[4], nbr[5], nbr[6], nbr[7]);
            }
        }
        // P2G list: blocks that hold particles and whose tile holds no collider, bucketed by population (common.cuh).
        const bool p2g_mine = np != 0u && flag == 0;
        const uint32_t my_bucket = p2g_bucket(np);
#pragma unroll
        for (uint32_t k = 0; k < P2G_BUCKETS; ++k) {
            const bool in = p2g_mine && my_bucket == k;
            if (!__any_sync(0xffffffffu, in)) continue; // (warp-uniform)
            const uint32_t pos = reserve(in ? 1u : 0u, &d.counters->num_p2g[k]);
            if (in) d.p2g_list[(size_t)k * d.capacity + pos] = i;
        }
    }
    // particles: SCATTER_ITEMS per thread, see k_touch
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t warp_base = (blockIdx.x * (SORT_THREADS / 32) + warp) * SCATTER_PER_WARP;
    const uint32_t n_live = d.counters->n_live;
    if (warp_base >= n_live) {
        TL_END(d, B200MPM_KERNEL_SCATTER);
        return;
    }
    uint32_t ck[SCATTER_ITEMS], rk[SCATTER_ITEMS], dest[SCATTER_ITEMS];
#pragma unroll
    for (int j = 0; j < SCATTER_ITEMS; ++j) {
        const uint32_t p = warp_base + j * 32 + lane;
        ck[j] = (p < n_live) ? d.pkey[p] : NONE;
        rk[j] = (p < n_live) ? d.rank[p] : 0u;
    }
#pragma unroll
    for (int j = 0; j < SCATTER_ITEMS; ++j) dest[j] = (ck[j] != NONE) ? d.cell_start[ck[j]] + rk[j] : NONE;
#pragma unroll
    for (int j = 0; j < SCATTER_ITEMS; ++j) {
        const uint32_t p = warp_base + j * 32 + lane;
        if (p >= n_live) continue;
        if (ck[j] != NONE) {
            d.sorted_ids[dest[j]] = p;
            if (d.has_bodies) d.cdf_aff[cur ^ 1][dest[j]] = 0u; // default cdf; k_g2p_cdf overwrites the flagged blocks
        } else {
            // Particle of a dropped block (capacity overflow): parked after the sorted range so that
            // its state survives the ping-pong (see k_g2p tail).
            uint32_t total = d.counters->sorted_total;
            uint32_t k = atomicAdd(&d.counters->dropped_particles, 1u);
            d.sorted_ids[total + k] = p;
            // a LIVE particle without a block (capacity / hash overflow): sharded runs compact the parked tail away
            // together with the emigrants, so the loss is reported (bit 2 of the overflow word)
            if ((__float_as_uint(d.pos4[cur][p].w) & FLAG_DEAD) == 0u) atomicOr(&d.counters->overflow, 4u);
        }
    }
    TL_END(d, B200MPM_KERNEL_SCATTER);
}

// ---- launch wrappers ------------------------------------------------------------------------------------
static inline int div_up(uint64_t a, uint64_t b) { return (int)((a + b - 1) / b); }

void launch_touch(const LaunchCfg& c, const DeviceData& d, int cur, bool integrate_first) {
    const int integrate = (integrate_first && d.has_bodies) ? 1 : 0;
    if (d.n == 0) {
        if (integrate) launch_integrate_bodies(c, d);
        return;
    }
    const int ctas = div_up(d.n, SORT_PER_CTA) + integrate;
    if (c.dim == 2) launch_pdl(k_touch<2>, ctas, SORT_THREADS, 0, c.stream, d, cur, integrate);
    else launch_pdl(k_touch<3>, ctas, SORT_THREADS, 0, c.stream, d, cur, integrate);
    ++*c.launch_counter;
}
void launch_transform_rigid(const LaunchCfg& c, const DeviceData& d) {
    const uint32_t n = d.num_rigid + d.num_mesh_verts;
    if (n == 0) return;
    if (c.dim == 2) k_transform_rigid<2><<<div_up(n, 256), 256, 0, c.stream>>>(d);
    else k_transform_rigid<3><<<div_up(n, 256), 256, 0, c.stream>>>(d);
    ++*c.launch_counter;
}
// mark + touch of the sample points' blocks; must follow launch_touch on the same stream.
void launch_touch_rigid(const LaunchCfg& c, const DeviceData& d) {
    if (d.num_rigid == 0) return;
    if (c.dim == 2) {
        k_mark_rigid<2><<<div_up(d.num_rigid, 256), 256, 0, c.stream>>>(d);
        k_touch_rigid<2><<<div_up(d.num_rigid, 256), 256, 0, c.stream>>>(d);
    } else {
        k_mark_rigid<3><<<div_up(d.num_rigid, 256), 256, 0, c.stream>>>(d);
        k_touch_rigid<3><<<div_up(d.num_rigid, 256), 256, 0, c.stream>>>(d);
    }
    *c.launch_counter += 2;
}
// Must follow launch_block_prepare (collide() initialises the node cdf) on the same stream.
void launch_p2g_cdf(const LaunchCfg& c, const DeviceData& d) {
    if (d.num_rigid == 0) return;
    if (c.dim == 2) k_p2g_cdf<2><<<div_up(d.num_rigid, 128), 128, 0, c.stream>>>(d);
    else k_p2g_cdf<3><<<div_up(d.num_rigid, 128), 128, 0, c.stream>>>(d);
    ++*c.launch_counter;
}
void launch_block_prepare(const LaunchCfg& c, const DeviceData& d) {
    const int grid = c.num_sms * (2048 / PREPARE_THREADS);
    if (c.dim == 2) launch_pdl(k_block_prepare<2>, grid, PREPARE_THREADS, 0, c.stream, d);
    else launch_pdl(k_block_prepare<3>, grid, PREPARE_THREADS, 0, c.stream, d);
    ++*c.launch_counter;
}
void launch_scatter(const LaunchCfg& c, const DeviceData& d, int cur) {
    if (d.n == 0) return;
    // (covers both the particles and, for the CPIC work list, the active blocks: blocks <= 2^D * particles,
    // and in practice far fewer; the grid is sized for whichever is larger)
    uint64_t max_blocks = (uint64_t)d.n * (c.dim == 2 ? 4 : 8);
    if (max_blocks > d.capacity) max_blocks = d.capacity;
    uint64_t ctas_p = div_up(d.n, SORT_THREADS * SCATTER_ITEMS), ctas_b = div_up(max_blocks, SORT_THREADS);
    int ctas = (int)(ctas_p > ctas_b ? ctas_p : ctas_b);
    if (c.dim == 2) launch_pdl(k_scatter<2>, ctas, SORT_THREADS, 0, c.stream, d, cur);
    else launch_pdl(k_scatter<3>, ctas, SORT_THREADS, 0, c.stream, d, cur);
    ++*c.launch_counter;
}
void launch_exclusive_scan_u32(const LaunchCfg& c, uint32_t* data, uint32_t len, uint64_t* scan_state,
                               uint32_t* ticket) {
    if (len == 0) return;
    uint32_t tiles = scan_num_tiles(len);
    cudaMemsetAsync(scan_state, 0, sizeof(uint64_t) * (tiles + 1), c.stream);
    cudaMemsetAsync(ticket, 0, sizeof(uint32_t), c.stream);
    k_scan<<<tiles, SCAN_THREADS, 0, c.stream>>>(data, len, nullptr, 0u, scan_state, ticket);
    ++*c.launch_counter;
}

} // namespace b2
