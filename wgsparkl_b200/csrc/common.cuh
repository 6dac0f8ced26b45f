// Shared device-side types and helpers for the B200 MPM substep.
// Reference citations are file:line relative to the wgsparkl source tree.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200mpm.h"

namespace b2 {

constexpr uint32_t NONE = 0xffffffffu; // grid.wgsl:80
constexpr uint32_t HVAL_DROPPED = 0xfffffffeu; // hvals entry of a block created beyond the grid capacity
constexpr int CELLS_PER_BLOCK = 64; // grid.wgsl:43
constexpr uint32_t MAT_ID_MASK = 0x0fffffffu;
constexpr uint32_t FLAG_DEAD = 0x40000000u; // the particle emigrated to a neighbour slab; dropped at the end of the substep
constexpr uint32_t FLAG_PHASE_BROKEN = 0x80000000u; // phases[i].phase was set to 0 (particle_update.wgsl:106,112)

template <int D>
struct Dim;
template <>
struct Dim<2> {
    static constexpr int BLOCK = 8; // 8x8 cells per block (particle2d.wgsl:46)
    static constexpr int TILE = 10; // p2g.wgsl:31
    static constexpr int TILE_CELLS = 100;
    static constexpr int NBH = 9; // kernel.wgsl:6
    static constexpr int NASSOC = 4; // grid.wgsl:209
    static constexpr int LOG_BLOCK = 3;
};
template <>
struct Dim<3> {
    static constexpr int BLOCK = 4; // 4x4x4 cells per block (particle3d.wgsl:43)
    static constexpr int TILE = 6; // p2g.wgsl:38
    static constexpr int TILE_CELLS = 216;
    static constexpr int NBH = 27; // kernel.wgsl:22
    static constexpr int NASSOC = 8; // grid.wgsl:211
    static constexpr int LOG_BLOCK = 2;
};

// ---- per-material constants (dedup of the reference's per-particle model buffers) -------
// ParticleDynamics.{init_volume,init_radius,mass} (particle3d.rs:23-25), ElasticCoefficients
// (models/mod.rs:63-68), DruckerPrager (drucker_prager.rs:6-15), ParticlePhase
// (particle_update.rs:37-42). 64 bytes.
struct __align__(16) Material {
    float mass, init_volume, init_radius, lambda;
    float mu, dp_h0, dp_h1, dp_h2;
    float dp_h3, dp_lambda, dp_mu, phase;
    float max_stretch;
    uint32_t model;
    float dp_ratio; // (d dp_lambda + 2 dp_mu) / (2 dp_mu) (drucker_prager.wgsl:56,125), filled by the host
    uint32_t pad1;
};

// ---- rigid bodies (wgrapier GpuBodySet + rigid_impulses.wgsl state), <= 16 ---------------
struct BodyDev {
    uint32_t shape_type;
    float shape_a[3];
    float shape_b[3];
    float radius;
    float rot[9]; // rotation matrix, column-major (2D: 2x2 in rot[0..3])
    float trans[3];
    float rot_raw[4]; // quaternion (i,j,k,w) / unit complex, what the host reads back
    float linvel[3];
    float angvel[3];
    float local_inv_mass[3];
    float local_inv_inertia[9];
    float local_com[3];
    // world-space mass properties (update_world_mass_properties, rigid_impulses.wgsl:139-150)
    float com[3];
    float inv_inertia[9];
    // IntegerImpulseAtomic (rigid_impulses.wgsl:27-47)
    int imp_lin[3];
    int imp_ang[3];
    // 0 if an impulse cannot change anything observable: infinite mass AND zero velocity (then applyImpulse is
    // a no-op and the velocity caps of rigid_impulses.wgsl:118-126 cannot trigger). Refreshed every substep.
    uint32_t needs_impulse;
};

struct SimState {
    float gravity[3];
    float dt;
    float cell_width;
    uint32_t num_bodies;
    int slab_lo, slab_hi; // this rank owns particles whose block x-index is in [slab_lo, slab_hi) (sharded runs)
};

// P2G work list: the blocks are bucketed by population (bucket 0: >= 7 * 96 particles ... bucket 7: < 96) and the
// persistent warps walk the buckets in order - longest items first, so that the last scheduling round (a warp gets
// ~3 half blocks per million particles) consists of the cheapest items instead of arbitrary ones.
constexpr uint32_t P2G_BUCKETS = 8, P2G_BUCKET_STEP = 96;
__host__ __device__ inline uint32_t p2g_bucket(uint32_t np) {
    const uint32_t q = np / P2G_BUCKET_STEP;
    return q >= P2G_BUCKETS - 1u ? 0u : P2G_BUCKETS - 1u - q;
}

// ---- device-resident counters --------------------------------------------------------------
struct Counters {
    uint32_t num_active_blocks; // Grid.num_active_blocks (grid.wgsl:223)
    uint32_t prev_active_blocks; // active count of the last sort (published by k_scatter): what the host reads, and
                                 // what k_begin_substep has to clear - num_active_blocks itself is already back
                                 // to zero when a substep ends
    uint32_t overflow; // sticky: the block capacity / hash map was exceeded at least once
    uint32_t scan_ticket; // dynamic tile id for the single-pass scan
    uint32_t work_p2g; // dynamic block schedulers
    uint32_t work_p2g_cpic;
    uint32_t work_g2p;
    uint32_t work_cdf;
    uint32_t num_cpic_blocks; // blocks (with particles) whose tile holds a collider this substep
    uint32_t num_g2p_items; // entries at the FRONT of g2p_list this substep (collider-side blocks: slow items first)
    uint32_t num_g2p_back; // entries at the BACK of g2p_list (all other blocks), filled downwards from the end
    uint32_t dropped_particles; // particles whose block was dropped (overflow) or that left the slab
    uint32_t n_live; // particles currently held (== n unless the data is a slab of a sharded run)
    uint32_t send_count[2]; // emigrants packed for the -x / +x neighbour (sharded runs)
    uint32_t halo_count[2]; // blocks packed for the -x / +x neighbour (sharded runs)
    uint32_t shard_seq; // substep sequence number of the peer-to-peer exchange flags
    uint32_t n_base; // live count at the start of the substep: where the immigrants are appended
    uint32_t emig_count; // entries of emig_list (k_g2p appends, k_emigrate_listed consumes and resets)
    uint32_t halo_ticket; // CTAs of k_halo_pack that are done (the last one publishes)
    uint32_t num_p2g[P2G_BUCKETS]; // entries of every population bucket of p2g_list (bucket 0 = most particles)
    uint32_t sorted_total; // particles in the sorted range this substep (== cell_start[num_active_blocks * 64])
    uint32_t integrate_pending; // a substep ran since the last k_integrate_bodies (which may be deferred, api.cu)
};

// G2P work items: a block's sorted range in parts of at most G2P_ITEM particles = one particle per thread of a
// k_g2p CTA (k_scatter builds the list; blocks without particles are never visited). An item carries everything the
// kernel needs to start fetching - slot range, collider flag, neighbour table - so that the chain
// "work counter -> list -> cell_start -> nbr -> nodes" of dependent loads collapses into ONE 64-byte load that is
// issued three items ahead of its use.
constexpr uint32_t G2P_ITEM = 128;
struct __align__(16) G2PItem {
    uint32_t block; // header id (NONE: end of work)
    uint32_t first; // first sorted slot
    uint32_t count; // <= G2P_ITEM
    uint32_t flags; // bit 0: the block's (BLOCK+2)^D tile holds a collider (block_flags)
    uint32_t nbr[8]; // header ids of the blocks vid + {0,1}^D (2D: the first four)
    uint32_t pad[4];
};
static_assert(sizeof(G2PItem) == 64, "G2PItem is fetched as 16 words by half a warp");
__host__ __device__ inline uint32_t g2p_parts(uint32_t n) { return (n + G2P_ITEM - 1) / G2P_ITEM; }

// ---- all device pointers of one MpmData --------------------------------------------------
struct DeviceData {
    uint32_t n; // particles
    uint32_t capacity; // block capacity == hash capacity (power of two; grid.rs:283)
    uint32_t num_materials;
    int has_plastic; // any particle can reach phase == 0
    int has_bodies; // CPIC on
    int bodies_react; // some body has mass or motion, i.e. impulses can matter (host-maintained, api.cu)

    // particle state, ping-pong (index 0/1). Layout: see DESIGN.md "Data layout in HBM".
    float4* pos4[2]; // x y z | bits(material id | flags)
    float4* vel4[2]; // vx vy vz | bits(original particle id)
    float4* Fa[2]; // F column-major elements 0..3
    float4* Fb[2]; // 4..7   (3D only)
    float* Fc[2]; // 8        (3D only)
    float4* Ca[2]; // APIC affine (momentum form, particle_update.wgsl:132), elements 0..3
    float4* Cb[2];
    float* Cc[2];
    float4* plastic[2]; // det, hardening, log_vol_gain, - (only if has_plastic)
    uint32_t* cdf_aff[2]; // particle CPIC affinity (persists across substeps; only if has_bodies)
    float4* cdf_nd; // normal xyz + signed distance, by sorted slot (only if has_bodies)
    float4* cdf_rv; // rigid_vel, by sorted slot (only if has_bodies)
    const Material* materials;

    // sort scratch
    uint32_t* pkey; // hash slot * 64 + cell-in-block, then (hid * 64 + cell)
    uint32_t* rank; // rank of the particle inside its cell
    uint32_t* sorted_ids; // sorted slot -> index into the current particle buffers

    // sparse grid
    uint32_t* hkeys; // packed block key or NONE (GridHashMapEntry.state, grid.wgsl:110-117)
    uint32_t* hvals; // block header id
    int4* block_vid; // ActiveBlockHeader.virtual_id (grid.wgsl:215-219)
    uint32_t* cell_start; // capacity*64 + 1: per-cell particle counts, scanned in place
    uint32_t* nbr; // capacity * NASSOC: header ids of blocks vid + {0,1}^D
    float4* node_mv; // capacity*64: momentum xyz + mass (Node.momentum_velocity_mass, grid.wgsl:257-267)
    // capacity*64: closest_id, bits(distance), affinities, - (NodeCdf, grid.wgsl:233-240). The first two words form
    // one 64-bit key (distance in the high half; distances are >= 0, so the key orders like the distance) that the
    // mesh-collider scatter lowers with a single atomicMin (k_p2g_cdf).
    uint4* node_cdf;
    // capacity*64*2 (with bodies): the node's body impulse (linear, angular) in float, summed over every block that
    // scatters to it and converted to fixed point ONCE per node (p2g.wgsl:142-155) by k_node_impulses, which also
    // clears it again: all zero outside launch_p2g.
    float4* node_imp;
    uint64_t* scan_state; // single-pass scan tile descriptors
    uint8_t* block_f0; // capacity: 1 if one of the block's own nodes is near / inside a collider (k_block_prepare)
    uint32_t* block_flags; // capacity: 1 if the block's (BLOCK+2)^D tile holds a collider (k_scatter)
    uint32_t* cpic_list; // capacity: compact list of flagged blocks that hold particles
    G2PItem* g2p_items; // g2p_items_len work items of <= G2P_ITEM particles (collider-side blocks at the front)
    uint32_t g2p_items_len; // capacity + n / G2P_ITEM + 1
    uint2* block_range; // capacity: (first sorted slot, particle count) of every block (k_scatter; read-back helpers)
    uint32_t* p2g_list; // P2G_BUCKETS x capacity: blocks that hold particles and whose tile holds no collider, by bucket

    BodyDev* bodies;
    // Rigid particles = sample points of trimesh / polyline colliders (GpuRigidParticles, particle3d.rs:82-88) and
    // the collider vertices their primitives refer to; world-space copies are refreshed every substep.
    uint32_t num_rigid, num_mesh_verts;
    float4* rp_local;
    float4* rp_world;
    uint4* rp_ids; // (vertex a, b, c, collider index)
    uint32_t* rp_needs_block; // one word per sample point (sort.wgsl:54-86)
    float4* mv_local;
    float4* mv_world;
    uint32_t* mv_body;
    SimState* sim;
    Counters* counters;
    // Debug timeline (only while timeline_on, b200mpm_debug_timeline): [2k] = earliest start, [2k+1] = latest end of
    // kernel k (B200MPM_KERNEL_*) in %globaltimer nanoseconds, stamped by thread 0 of every CTA.
    unsigned long long* timeline;
    int timeline_on;
    // Peer-to-peer sharded runs: slots (in the buffers k_g2p writes) of the particles whose new position left the slab,
    // so that the migration packs a few hundred particles instead of scanning all of them. Null otherwise.
    uint32_t* emig_list;
    uint32_t emig_cap;
};

// ---- small math --------------------------------------------------------------------------
__host__ __device__ inline uint32_t pack_key2(int x, int y) { // grid.wgsl:83-86
    return ((uint32_t)(x + 0x00007fff) & 0x0000ffffu) | (((uint32_t)(y + 0x00007fff) & 0x0000ffffu) << 16);
}
__host__ __device__ inline uint32_t pack_key3(int x, int y, int z) { // grid.wgsl:88-95
    return ((uint32_t)(x + 0x000003ff) & 0x000007ffu) | (((uint32_t)(y + 0x000001ff) & 0x000003ffu) << 11) |
           (((uint32_t)(z + 0x000003ff) & 0x000007ffu) << 21);
}
template <int D>
__host__ __device__ inline uint32_t pack_key(int x, int y, int z) {
    if (D == 2) return pack_key2(x, y);
    return pack_key3(x, y, z);
}
__host__ __device__ inline uint32_t hash_key(uint32_t k) { // grid.wgsl:98-105 (murmur3 finaliser step)
    k *= 0xcc9e2d51u;
    k = (k << 15) | (k >> 17);
    k *= 0x1b873593u;
    return k;
}

// round(p / h) with round = ties-to-even and a true IEEE division, bit-exact with grid.wgsl:285 /
// particle3d.wgsl:42 (never roundf) - without paying for the division: q = p * (1/h) is within
// 1.8e-7 |q| of the correctly rounded quotient, so rint(q) can only differ from rint(p / h) when a half-integer
// lies that close to q. Those (rare) lanes take the IEEE division; everyone else a multiply and a compare.
__host__ __device__ __forceinline__ float round_div(float p, float h, float inv_h) {
    const float q = p * inv_h;
    const float r = rintf(q);
    // (|q| >= 2^22: q - r == 0 and 4e-7 |q| > 0.5, so the test sends those to the division as well; NaN / inf come
    // out of rintf(q) as they would out of the division)
#ifdef __CUDA_ARCH__
    if (0.5f - fabsf(q - r) <= 4e-7f * fabsf(q)) return rintf(__fdiv_rn(p, h));
#else
    if (0.5f - fabsf(q - r) <= 4e-7f * fabsf(q)) return rintf(p / h); // (host build: tests/cpp/round_div_host.cu)
#endif
    return r;
}

// Associated cell of a coordinate: round(p / h) - 1. inv_h must be the IEEE 1.0f / h.
__device__ __forceinline__ int assoc_cell(float p, float h, float inv_h) { return (int)(round_div(p, h, inv_h) - 1.0f); }

__device__ __forceinline__ uint32_t find_block(const uint32_t* __restrict__ hkeys, const uint32_t* __restrict__ hvals,
                                               uint32_t cap_mask, uint32_t packed) { // grid.wgsl:167-184
    uint32_t slot = hash_key(packed) & cap_mask;
    for (uint32_t k = 0; k <= cap_mask; ++k) {
        uint32_t s = __ldg(hkeys + slot);
        if (s == packed) {
            const uint32_t hid = __ldg(hvals + slot);
            return hid == HVAL_DROPPED ? NONE : hid;
        }
        if (s == NONE) return NONE;
        slot = (slot + 1) & cap_mask;
    }
    return NONE;
}

// CPIC affinity words (grid.wgsl:230-255)
__device__ __forceinline__ bool affinities_are_compatible(uint32_t a1, uint32_t a2) {
    uint32_t common = a1 & a2 & 0xffffu;
    return (((a1 >> 16) ^ (a2 >> 16)) & common) == 0u;
}

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }

// Contraction-free variants (every product and sum rounded on its own, like WGSL without fused multiply-add and
// like the oracle's -ffp-contract=off): for code whose RESULT IS AN INTEGER DECISION - the affinity / sign bits of
// the mesh-collider colouring must match bit for bit, and one fused multiply-add next to a triangle edge flips them.
__device__ __forceinline__ float dot_rn(V3 a, V3 b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ V3 cross_rn(V3 a, V3 b) {
    return V3{__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
              __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x))};
}

// project_velocity (grid.wgsl:390-404). For 2D pass z = 0.
__device__ __forceinline__ V3 project_velocity(V3 vel, V3 n) {
    float normal_vel = dot(vel, n);
    if (normal_vel < 0.0f) {
        const float friction = 20.0f;
        V3 t = vel - n * normal_vel;
        float len = length(t);
        V3 dir = (len > 1.0e-8f) ? t * (1.0f / len) : v3(0, 0, 0);
        return dir * fmaxf(0.0f, len + friction * normal_vel);
    }
    return vel;
}

// Body::velocity_at_point (wgrapier body.wgsl). 2D: angular in angvel[0].
template <int D>
__device__ __forceinline__ V3 velocity_at_point(const BodyDev& b, V3 pt) {
    V3 d = pt - v3(b.com[0], b.com[1], b.com[2]);
    if (D == 2) return v3(b.linvel[0] - d.y * b.angvel[0], b.linvel[1] + d.x * b.angvel[0], 0.0f);
    V3 w = v3(b.angvel[0], b.angvel[1], b.angvel[2]);
    return v3(b.linvel[0], b.linvel[1], b.linvel[2]) + cross(w, d);
}

// Quadratic B-spline weights (kernel.wgsl:61-67); x in [0.5, 1.5].
__device__ __forceinline__ void bspline(float x, float& w0, float& w1, float& w2) {
    float a = 1.5f - x, b = x - 1.0f, c = x - 0.5f;
    w0 = 0.5f * a * a;
    w1 = 0.75f - b * b;
    w2 = 0.5f * c * c;
}

// flt2int (rigid_impulses.wgsl:52-54): i32(x * 1e5), truncating and saturating like WGSL.
__device__ __forceinline__ int flt2int(float f) {
    float x = f * 1e5f;
    if (!(x == x)) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (-2147483647 - 1);
    return (int)x; // cvt.rzi
}

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned long long tl_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TL_BEGIN(d, k) do { if ((d).timeline_on && threadIdx.x == 0) atomicMin((d).timeline + 2 * (k), tl_now()); } while (0)
#define TL_END(d, k) do { if ((d).timeline_on && threadIdx.x == 0) atomicMax((d).timeline + 2 * (k) + 1, tl_now()); } while (0)
// ---- reset_hmap (grid.wgsl:186-203) + clearing of the last sort's per-cell bins -------------------------------------
// Nothing after P2G (and the halo exchange of sharded runs) reads the hash map or the bins any more - G2P works from
// its item list - so the clearing for the NEXT substep does not sit at the top of the critical path: the CTAs of k_g2p
// do it as they retire (as a kernel on a side branch it either ran in front of k_g2p, delaying part of its
// statically scheduled grid, or waited for its first CTAs to retire anyway).
__device__ __forceinline__ void clear_sparse_grid(const DeviceData& d, uint32_t tid, uint32_t stride) {
    for (uint32_t i = tid; i < d.capacity; i += stride) d.hkeys[i] = NONE, d.hvals[i] = NONE; // (hvals: see publish_block)
    // prev_active_blocks was published by the last k_scatter (0 before the first sort: the arrays are
    // zero-initialised at creation)
    const uint32_t prev = min(d.counters->prev_active_blocks, d.capacity);
    const uint32_t nbins = prev * CELLS_PER_BLOCK + 1;
    for (uint32_t i = tid; i < nbins; i += stride) d.cell_start[i] = 0;
    if (tid == 0) d.counters->num_active_blocks = 0;
}

// ---- cp.async (LDGSTS): global -> shared without register staging -------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
// ---- mbarrier (shared-memory transaction barriers): producer / consumer pipelines without CTA-wide barriers --------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
// arrive (release): everything this thread wrote before is visible to a thread that observes the phase completion
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// arrive once all cp.async requests this thread has issued so far have landed (counts against the barrier's expected arrivals)
__device__ __forceinline__ void mbar_arrive_on_cp_async(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// wait (acquire) for the completion of the phase with the given parity
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }\n"
                     : "=r"(ok)
                     : "r"(a), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void prefetch_l1(const void* gmem) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(gmem)); }
__device__ __forceinline__ void prefetch_l2(const void* gmem) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(gmem)); }
#endif

} // namespace b2
