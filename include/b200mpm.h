/*
 * b200mpm.h — C ABI of the B200-native MPM substep (drop-in for wgsparkl's hot path).
 *
 * Every entry point replaces one piece of the reference's Rust/wgpu surface; the
 * reference location is cited as (file:line) relative to the wgsparkl source tree.
 * Plain pointers and sizes only; no torch / CUDA types cross this boundary.
 *
 *   - all functions return 0 (B200MPM_OK) or a negative b200mpm_status;
 *   - b200mpm_last_error() returns a static, thread-local, human-readable message;
 *   - handles are NOT thread-safe; distinct handles may be used from distinct threads;
 *   - input pointers are borrowed for the duration of the call, outputs are
 *     caller-allocated HOST buffers unless the name says `_device`.
 *
 * There is no CPU fallback behind this ABI: if no sm_100 device is present,
 * b200mpm_pipeline_create() fails with B200MPM_ERR_NO_DEVICE.
 */
#ifndef B200MPM_H
#define B200MPM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MPM_MAX_BODIES 16u /* rigid_impulses.rs:42, collide.wgsl:36 (CPIC bitmask width) */
#define B200MPM_NONE 0xffffffffu /* grid.wgsl:80 */

typedef enum b200mpm_status {
    B200MPM_OK = 0,
    B200MPM_ERR_INVALID_ARGUMENT = -1,
    B200MPM_ERR_NO_DEVICE = -2,
    B200MPM_ERR_CUDA = -3,
    B200MPM_ERR_OUT_OF_MEMORY = -4,
    B200MPM_ERR_GRID_OVERFLOW = -5, /* reported by b200mpm_data_status only; stepping never aborts */
    B200MPM_ERR_COMM = -6
} b200mpm_status;

/* ---- plain-old-data mirrors of the reference's public structs ------------------ */

/* SimulationParams (src/solver/params.rs:6-16). 2D uses gravity[0..2]. */
typedef struct b200mpm_sim_params {
    float gravity[3];
    float dt;
} b200mpm_sim_params;

/* Constitutive model selector (additive; the reference hard-wires corotated:
 * src/solver/particle_update.wgsl:7-8). */
enum { B200MPM_MODEL_COROTATED = 0, B200MPM_MODEL_NEO_HOOKEAN = 1 };

/*
 * One MPM particle = Particle (src/solver/particle3d.rs:53-60 / particle2d.rs:49-56)
 * flattened: position + ParticleDynamics (particle3d.rs:16-26) + Cdf (43-50) +
 * ElasticCoefficients (models/mod.rs:63-68) + DruckerPrager (models/drucker_prager.rs:6-15)
 * + DruckerPragerPlasticState (36-42) + ParticlePhase (solver/particle_update.rs:37-42).
 * Matrices are column-major DIMxDIM packed in the first DIM*DIM floats (nalgebra order).
 * `Option::None` must be resolved by the caller exactly like GpuModels::from_particles
 * (models/mod.rs:20-36): plasticity None -> DruckerPrager::new(-1,-1); phase None -> {0,-1}.
 */
typedef struct b200mpm_particle {
    float position[3];
    float velocity[3];
    float def_grad[9];
    float affine[9];
    float cdf_normal[3];
    float cdf_rigid_vel[3];
    float cdf_signed_distance;
    uint32_t cdf_affinity;
    float init_volume;
    float init_radius;
    float mass;
    float lambda; /* ElasticCoefficients */
    float mu;
    float dp_h0, dp_h1, dp_h2, dp_h3, dp_lambda, dp_mu; /* DruckerPrager */
    float plastic_det, plastic_hardening, plastic_log_vol_gain; /* DruckerPragerPlasticState */
    float phase, max_stretch; /* ParticlePhase */
    uint32_t model; /* B200MPM_MODEL_* */
} b200mpm_particle;

/* Analytic collider shapes handled by collide() (src/collision/collide.wgsl:23-55). */
enum {
    B200MPM_SHAPE_BALL = 0,
    B200MPM_SHAPE_CUBOID = 1,
    B200MPM_SHAPE_CAPSULE = 2,
    /* mesh colliders: no analytic projection (collide.wgsl:41 skips them); they act through their sample points,
     * see b200mpm_data_set_rigid_particles */
    B200MPM_SHAPE_TRIMESH = 3, /* 3D; also what a heightfield is converted to (particle3d.rs:123-132) */
    B200MPM_SHAPE_POLYLINE = 4 /* 2D */
};

/*
 * One coupled collider + its parent body = one BodyCouplingEntry (src/pipeline.rs:107-117);
 * replaces what GpuBodySet::from_rapier uploads (src/pipeline.rs:141). Index in the array is
 * the CPIC collider index (bit position in the affinity masks).
 *   shape_a: ball -> unused; cuboid -> half extents; capsule -> segment endpoint a (local)
 *   shape_b: capsule -> segment endpoint b (local)
 *   rotation: 3D unit quaternion (i,j,k,w); 2D unit complex (re,im,0,0)
 *   inv_inertia: 3D local-frame inverse inertia tensor, column-major 3x3; 2D scalar in [0]
 *   inv_mass: per-axis inverse mass (0 for fixed / kinematic bodies)
 */
typedef struct b200mpm_body {
    uint32_t shape_type;
    float shape_a[3];
    float shape_b[3];
    float radius;
    float translation[3];
    float rotation[4];
    float linvel[3];
    float angvel[3];
    float inv_mass[3];
    float inv_inertia[9];
    float local_com[3];
    uint32_t two_ways; /* BodyCoupling::TwoWays (pipeline.rs:114); 0 = the MPM never moves the body */
} b200mpm_body;

/* GpuSim pose as written by the testbed each frame (src_testbed/step.rs:79-96). */
typedef struct b200mpm_pose {
    float translation[3];
    float rotation[4];
} b200mpm_pose;

/* GpuVelocity (src_testbed/step.rs:98-119). */
typedef struct b200mpm_velocity {
    float linear[3];
    float angular[3];
} b200mpm_velocity;

/* One active grid block in canonical form (debug / parity tooling). */
typedef struct b200mpm_block_info {
    int32_t vid[3]; /* BlockVirtualId (grid.wgsl:55-61) */
    uint32_t first_particle; /* ActiveBlockHeader.first_particle (grid.wgsl:215-219) */
    uint32_t num_particles;
} b200mpm_block_info;

/* Grid node (grid.wgsl:257-267): momentum/velocity + mass, then NodeCdf (233-240). */
typedef struct b200mpm_node {
    float momentum_velocity_mass[4]; /* 2D: xy + mass in [2] */
    float cdf_distance;
    uint32_t cdf_affinities;
    uint32_t cdf_closest_id;
} b200mpm_node;

/* The reference's ten timed passes (src/pipeline.rs:201-271, src_testbed/lib.rs:133-146). */
enum {
    B200MPM_PASS_UPDATE_RIGID_PARTICLES = 0,
    B200MPM_PASS_GRID_SORT = 1,
    B200MPM_PASS_GRID_UPDATE_CDF = 2,
    B200MPM_PASS_P2G_CDF = 3,
    B200MPM_PASS_G2P_CDF = 4,
    B200MPM_PASS_P2G = 5,
    B200MPM_PASS_GRID_UPDATE = 6,
    B200MPM_PASS_G2P = 7,
    B200MPM_PASS_PARTICLES_UPDATE = 8,
    B200MPM_PASS_INTEGRATE_BODIES = 9,
    B200MPM_NUM_PASSES = 10
};

/* Kernel-level timers of this implementation (b200mpm_get_kernel_timings): finer than the reference's pass names,
 * e.g. the two P2G instantiations, which the "p2g" pass sums. Used by bench.py's per-kernel roofline. */
enum {
    B200MPM_KERNEL_TOUCH = 0, /* touch_particle_blocks + update_block_particle_count */
    B200MPM_KERNEL_COUNT = 1, /* (0: part of TOUCH) */
    B200MPM_KERNEL_SCAN = 2, /* (0: the blocks' ranges are taken in BLOCK_PREPARE) */
    B200MPM_KERNEL_BLOCK_PREPARE = 3, /* ranges + neighbour table + node reset + grid_update_cdf */
    B200MPM_KERNEL_SCATTER = 4,
    B200MPM_KERNEL_G2P_CDF = 5, /* (0: inside P2G) */
    B200MPM_KERNEL_P2G_CPIC = 6, /* (0: inside P2G) */
    B200MPM_KERNEL_P2G = 7, /* particle colouring + collider-side blocks + all other blocks */
    B200MPM_KERNEL_BEGIN = 8, /* reset_hmap for the next substep (0 inside a substep: the tail of G2P) */
    B200MPM_KERNEL_G2P = 9, /* grid_update + g2p + particles_update */
    B200MPM_KERNEL_INTEGRATE_BODIES = 10,
    B200MPM_KERNEL_RIGID = 11, /* mesh-collider kernels (transform, mark / touch, p2g_cdf) */
    B200MPM_KERNEL_SHARD_MIGRATE = 12, /* sharded runs: emigrate .. immigrate */
    B200MPM_KERNEL_SHARD_HALO = 13, /* sharded runs: node-halo pack .. add */
    B200MPM_KERNEL_SHARD_END = 14, /* sharded runs: live count of the next substep */
    B200MPM_NUM_KERNELS = 15
};

typedef struct b200mpm_pipeline b200mpm_pipeline; /* MpmPipeline (src/pipeline.rs:24-39) */
typedef struct b200mpm_data b200mpm_data; /* MpmData (src/pipeline.rs:84-95) */

/* ---- pipeline ------------------------------------------------------------------ */

/* MpmPipeline::new(&Device) (src/pipeline.rs:176-193). `device` is the CUDA ordinal, dim is 2 or 3. */
int b200mpm_pipeline_create(int device, int dim, b200mpm_pipeline** out);
void b200mpm_pipeline_destroy(b200mpm_pipeline* p);
/* Run on a caller-owned CUDA stream (a `cudaStream_t` passed as void*; NULL restores the pipeline's
 * own stream). Plays the role of the caller-provided wgpu Queue / CommandEncoder
 * (src_testbed/step.rs:74-76): the host decides which queue the substeps are submitted to. */
int b200mpm_pipeline_set_stream(b200mpm_pipeline* p, void* cuda_stream);
const char* b200mpm_last_error(void);
/* Kernel launches issued by this pipeline since creation (bench bookkeeping). */
uint64_t b200mpm_pipeline_launch_count(const b200mpm_pipeline* p);

/* ---- data ---------------------------------------------------------------------- */

/* MpmData::with_select_coupling (src/pipeline.rs:130-168): copies particles and bodies to
 * the device. grid_capacity is rounded up to a power of two (src/grid/grid.rs:283). */
int b200mpm_data_create(b200mpm_pipeline* p, const b200mpm_sim_params* params,
                        const b200mpm_particle* particles, size_t num_particles,
                        const b200mpm_body* bodies, size_t num_bodies, float cell_width,
                        uint32_t grid_capacity, b200mpm_data** out);
void b200mpm_data_destroy(b200mpm_data* d);
size_t b200mpm_data_num_particles(const b200mpm_data* d);
size_t b200mpm_data_num_bodies(const b200mpm_data* d);

/* ---- stepping ------------------------------------------------------------------ */

/* MpmPipeline::queue_step + `for _ in 0..num_substeps { queue.encode }` + submit
 * (src/pipeline.rs:195-281, src_testbed/step.rs:122-128,169): enqueues `num_substeps`
 * substeps on the pipeline's stream and returns without waiting. */
int b200mpm_step(b200mpm_pipeline* p, b200mpm_data* d, uint32_t num_substeps);
/* device.poll(Maintain::Wait) (src/pipeline.rs:339). */
int b200mpm_sync(b200mpm_pipeline* p);
/* Toggle per-pass CUDA-event timestamps (queue.compute_pass(name, add_timestamps),
 * src/pipeline.rs:201); off by default. */
int b200mpm_set_timestamps(b200mpm_pipeline* p, int enabled);
/* Accumulated ms per reference pass since the last call (src_testbed/step.rs:219-254). Syncs. */
int b200mpm_get_timings(b200mpm_pipeline* p, double ms[B200MPM_NUM_PASSES]);
/* The same accumulation per kernel of this implementation (no reference counterpart; timestamps mode only). Syncs. */
int b200mpm_get_kernel_timings(b200mpm_pipeline* p, double ms[B200MPM_NUM_KERNELS]);
/* Debug: [2k] earliest start / [2k+1] latest end (GPU %globaltimer, ns; all ones / zero for a kernel that did not run)
 * of kernel k since the last call, then reset - what ran when INSIDE the graph replay, where events cannot look
 * (tools/timeline.py, bench.py's in-graph kernel durations). The first call switches the recording on (thread 0 of
 * every CTA stamps its start and end: two atomics) and returns the reset values; ns = NULL switches it off again.
 * Either switch re-captures the substep graphs. Syncs. */
int b200mpm_debug_timeline(b200mpm_data* d, uint64_t ns[2 * B200MPM_NUM_KERNELS]);

/* ---- per-frame host writes / reads (src_testbed/step.rs:79-119,175-176, ui.rs:98-103) -- */
int b200mpm_write_sim_params(b200mpm_data* d, const b200mpm_sim_params* params);
int b200mpm_write_body_poses(b200mpm_data* d, const b200mpm_pose* poses, size_t n);
int b200mpm_write_body_vels(b200mpm_data* d, const b200mpm_velocity* vels, size_t n);
int b200mpm_read_body_poses(b200mpm_data* d, b200mpm_pose* poses, size_t n);
int b200mpm_read_body_vels(b200mpm_data* d, b200mpm_velocity* vels, size_t n);

/* ---- readback of particle / grid state (tests, renderer hand-off) ------------------ */

/* particles.positions (src/solver/particle3d.rs:177): vec4 per particle (2D: xy00), in the
 * caller's original particle order. `out` holds 4*num_particles floats. */
int b200mpm_read_positions(b200mpm_data* d, float* out);
/* Same result, asynchronous: the positions as of the work enqueued so far are gathered on the pipeline's stream
 * and copied into `out` (which should be page-locked host memory, or the copy serialises) on a separate copy
 * stream, overlapping with whatever is enqueued next; `out` is valid after b200mpm_sync. At most two readbacks
 * are in flight (the third waits for the first). This is the wgpu staging-buffer + map_async pattern of the
 * reference's readbacks (src_testbed/step.rs:131-176). */
int b200mpm_read_positions_async(b200mpm_data* d, float* out);
/* Full particle state, original order. */
int b200mpm_read_particles(b200mpm_data* d, b200mpm_particle* out);
/* Sparse grid after the last substep, in the device's own block order.
 * blocks: capacity entries max; nodes: 64 per block. Returns the count in *num_blocks. */
int b200mpm_read_grid(b200mpm_data* d, b200mpm_block_info* blocks, b200mpm_node* nodes,
                      size_t capacity, size_t* num_blocks);
/* sorted_particle_ids (src/solver/particle3d.rs:179) expressed in ORIGINAL particle ids. */
int b200mpm_read_sorted_ids(b200mpm_data* d, uint32_t* out);
/* Number of active blocks after the last substep; B200MPM_ERR_GRID_OVERFLOW if the
 * block capacity was exceeded at any point (the reference drops blocks silently,
 * grid.wgsl:126-128). */
int b200mpm_data_status(b200mpm_data* d, uint32_t* num_active_blocks);

/* ---- render hand-off: the step on the far side of the path (src_testbed/prep_vertex_buffer{2d,3d}.wgsl:40-108,
 * src_testbed/prep_vertex_buffer.rs:11-57) -----------------------------------------------------------------------
 * One InstanceData per particle, in the caller's ORIGINAL particle order, in the testbed's vertex-buffer layout
 * (src_testbed/instancing3d.rs:66-73: three vec4 columns of the deformation gradient, vec4 position, base_color,
 * color; 96 bytes). Like the shader it writes xyz of the deformation columns and of the position plus `color`,
 * and reads (never writes) `base_color`; the w lanes are left as they are. */
typedef struct b200mpm_instance {
    float deformation[12]; /* 3 columns, each padded to 4 floats (2D: (F.x, 0), (F.y, 0), (0, 0, 1)) */
    float position[4];
    float base_color[4];
    float color[4];
} b200mpm_instance;
enum { /* RenderMode (prep_vertex_buffer.rs:11-18, prep_vertex_buffer3d.wgsl:25-30) */
    B200MPM_RENDER_DEFAULT = 0,
    B200MPM_RENDER_VOLUME = 1,
    B200MPM_RENDER_VELOCITY = 2,
    B200MPM_RENDER_CDF_NORMALS = 3,
    B200MPM_RENDER_CDF_DISTANCES = 4,
    B200MPM_RENDER_CDF_SIGNS = 5
};
/* `dev_instances` is DEVICE memory with room for num_particles instances - typically the renderer's vertex buffer
 * imported into CUDA (cudaImportExternalMemory on a Vulkan/D3D12 allocation), so that drawing needs no host round
 * trip. Asynchronous on the pipeline's stream. Not available for sharded data (a slab holds a changing subset). */
int b200mpm_prep_vertex_buffer(b200mpm_pipeline* p, b200mpm_data* d, b200mpm_instance* dev_instances, uint32_t mode);

/* ---- mesh colliders: trimesh / heightfield (3D) and polyline (2D) coupling -----------------------------------
 * GpuRigidParticles::from_rapier (src/solver/particle3d.rs:101-160, particle2d.rs:80-140): the host samples every
 * mesh collider (sample_mesh / sample_polyline, sampling step = cell_width, pipeline.rs:140) and hands over, in the
 * colliders' LOCAL frames,
 *   vertices          3 floats per collider vertex (2D: z = 0),      vertex_colliders  its collider index,
 *   samples           3 floats per sample point,                      sample_ids        4 uint32 per sample point:
 *                     the vertex ids of its triangle (2D: segment, third id ignored) and its collider index.
 * Collider index = position in the `bodies` array of b200mpm_data_create, whose shape_type must be
 * B200MPM_SHAPE_TRIMESH / _POLYLINE for these colliders. Every substep then runs the reference's
 * transform_sample_points / transform_shape_points, mark + touch_rigid_particle_blocks and p2g_cdf
 * (rigid_particle_update.wgsl:26-50, sort.wgsl:38-86, p2g_cdf.wgsl:51-190). Call once, before stepping. */
int b200mpm_data_set_rigid_particles(b200mpm_data* d, const float* vertices, const uint32_t* vertex_colliders,
                                     size_t num_vertices, const float* samples, const uint32_t* sample_ids,
                                     size_t num_samples);

/* Block-capacity growth - the part the reference leaves as a stub ("TODO: handle grid buffer resizing",
 * src/grid/grid.rs:43-118). b200mpm_data_reserve_grid replaces the capacity-sized arrays by larger ones
 * (capacity rounded up to a power of two like grid.rs:283; never shrinks; synchronises; the grid readback is
 * empty until the next substep). b200mpm_data_set_auto_grow(d, max_load) with max_load in (0, 1] makes every
 * b200mpm_step / b200mpm_shard_step call check the active block count left by the previous substeps (one stream
 * synchronisation per call) and double the capacity while count > max_load * capacity; 0 (the default) keeps
 * the reference's fixed capacity. */
int b200mpm_data_reserve_grid(b200mpm_data* d, uint32_t grid_capacity);
int b200mpm_data_set_auto_grow(b200mpm_data* d, float max_load);

/* Run only the "grid sort" pass (WgGrid::queue_sort, src/grid/grid.rs:30-207) — mirrors the
 * reference's gpu_grid_sort smoke test (grid.rs:347-402). */
int b200mpm_sort_only(b200mpm_pipeline* p, b200mpm_data* d);

/* WgPrefixSum::queue on a host vector (src/grid/prefix_sum.rs:20-69): in-place exclusive scan
 * ("as if a 0 was prepended", prefix_sum.rs:7-8) computed on the device. */
int b200mpm_prefix_sum_u32(b200mpm_pipeline* p, uint32_t* data, size_t len);

/* ---- multi-GPU slab sharding (new: the reference is single-device) ------------------------
 *
 * One process per GPU; rank r owns the particles whose block x-index lies in [x_lo, x_hi). The library packs
 * and unpacks DEVICE buffers; the caller moves them between ranks (NCCL send/recv and all-reduce — through
 * torch.distributed in wgsparkl_b200/sharded.py, directly in a Rust host). Per substep:
 *
 *   b200mpm_shard_emigrate   -> exchange records with the -x / +x neighbours -> b200mpm_shard_immigrate
 *   b200mpm_shard_step_begin    (sort ... P2G on the rank's own particles)
 *   b200mpm_shard_halo_pack  -> exchange the shared block columns             -> b200mpm_shard_halo_add
 *   b200mpm_shard_impulses(read) -> all-reduce(sum, int32[16*6])              -> b200mpm_shard_impulses(write)
 *   b200mpm_shard_step_end      (G2P + particle update, body integration)
 */
#define B200MPM_PARTICLE_RECORD_BYTES 128u /* one migrating particle */
#define B200MPM_HALO_BLOCK_BYTES 1040u /* one halo block: virtual id + 64 x (momentum xyz, mass) */
#define B200MPM_SHARD_HEADER_BYTES 16u

/* b200mpm_data_create with room for `particle_capacity` >= num_particles particles (immigrants) and
 * caller-chosen particle ids (NULL = 0..n-1; a sharded run passes global ids so that migrated particles stay
 * identifiable). */
int b200mpm_data_create_ex(b200mpm_pipeline* p, const b200mpm_sim_params* params, const b200mpm_particle* particles,
                           size_t num_particles, const uint32_t* particle_ids, size_t particle_capacity,
                           const b200mpm_body* bodies, size_t num_bodies, float cell_width, uint32_t grid_capacity,
                           b200mpm_data** out);
int b200mpm_slab_configure(b200mpm_data* d, int32_t x_lo, int32_t x_hi);
/* Particles currently held by this data (== num_particles unless sharded). Synchronises. */
int b200mpm_data_num_live(b200mpm_data* d, size_t* num_live);
/* Exchange buffers are DEVICE memory: a 16-byte header (uint32 record count + padding, written and read on the
 * device: no host round trip) followed by `cap` records. Fixed-size buffers travel between the ranks.
 *
 * Packs the particles that left the slab into the -x / +x buffers and flags them dead (they are dropped at the
 * end of the substep). Buffer size: 16 + cap_records * B200MPM_PARTICLE_RECORD_BYTES. Asynchronous. */
int b200mpm_shard_emigrate(b200mpm_pipeline* p, b200mpm_data* d, void* dev_left, void* dev_right, uint32_t cap_records);
/* Appends the records of a received migration buffer to the live particles. Asynchronous. */
int b200mpm_shard_immigrate(b200mpm_pipeline* p, b200mpm_data* d, const void* dev_buffer, uint32_t cap_records);
int b200mpm_shard_step_begin(b200mpm_pipeline* p, b200mpm_data* d);
/* Packs the node momenta of the shared block columns (x == x_lo -> dev_left, x == x_hi -> dev_right).
 * Buffer size: 16 + cap_blocks * B200MPM_HALO_BLOCK_BYTES. Asynchronous. */
int b200mpm_shard_halo_pack(b200mpm_pipeline* p, b200mpm_data* d, void* dev_left, void* dev_right, uint32_t cap_blocks);
/* Adds a received halo buffer to the blocks this rank also holds. Asynchronous. */
int b200mpm_shard_halo_add(b200mpm_pipeline* p, b200mpm_data* d, const void* dev_buffer, uint32_t cap_blocks);
/* write == 0: body impulses -> dev_buf (int32[16*6]); write != 0: dev_buf -> body impulses. */
int b200mpm_shard_impulses(b200mpm_pipeline* p, b200mpm_data* d, int32_t* dev_buf, int write);
int b200mpm_shard_step_end(b200mpm_pipeline* p, b200mpm_data* d);
/* Native transport: the library owns an NCCL communicator (libnccl.so.2 is resolved with dlopen, i.e. the copy
 * the host process already uses) and the exchange buffers, and b200mpm_shard_step enqueues WHOLE sharded substeps
 * — kernels, the two neighbour send/recv groups and the impulse all-reduce — on the pipeline's stream, replayed
 * from one captured CUDA graph per substep. rank 0 calls b200mpm_nccl_unique_id and distributes the 128 bytes
 * (any out-of-band channel: torch.distributed in the Python host), then every rank calls
 * b200mpm_shard_comm_init after b200mpm_slab_configure. */
int b200mpm_nccl_unique_id(void* out, size_t bytes);
int b200mpm_shard_comm_init(b200mpm_pipeline* p, b200mpm_data* d, int rank, int world, const void* unique_id,
                            uint32_t migration_cap_records, uint32_t halo_cap_blocks);
int b200mpm_shard_step(b200mpm_pipeline* p, b200mpm_data* d, uint32_t num_substeps);
/* Optional, after b200mpm_shard_comm_init: replace the two NCCL send/recv groups of every substep by direct
 * NVLink stores. Each rank exports a CUDA-IPC handle (64 bytes) of its receive arena, the handles of ALL ranks are
 * gathered out of band, and each rank maps its two neighbours. The pack kernels then write the migration records
 * and the halo blocks straight into the neighbour's memory and raise a sequence flag there; the unpack side spins
 * on its local flag. (The impulse all-reduce stays on NCCL.) */
int b200mpm_shard_p2p_export(b200mpm_pipeline* p, b200mpm_data* d, void* handle_out, size_t bytes);
int b200mpm_shard_p2p_connect(b200mpm_pipeline* p, b200mpm_data* d, const void* handles, size_t num_handles);
/* Positions of the live particles in device order: xyz + the particle id's bits in w (B200MPM_NONE for a
 * particle that has emigrated). `out` holds 4 * capacity floats. */
int b200mpm_read_positions_unordered(b200mpm_data* d, float* out, size_t capacity, size_t* count);
/* Same, with the copy on the copy stream (see b200mpm_read_positions_async) and WITHOUT any host synchronisation: the
 * copy covers all particle_capacity slots (so `capacity` must be at least that; *count returns it), and the slots that
 * hold no particle - spare capacity, emigrated particles - carry the id NONE (0xffffffff) in w. `out` is valid after
 * b200mpm_sync. */
int b200mpm_read_positions_unordered_async(b200mpm_data* d, float* out, size_t capacity, size_t* count);
/* Live particles in device order with their ids (no un-permutation). */
int b200mpm_read_particles_unordered(b200mpm_data* d, b200mpm_particle* out, uint32_t* ids, size_t capacity,
                                     size_t* count);

#ifdef __cplusplus
}
#endif
#endif /* B200MPM_H */
