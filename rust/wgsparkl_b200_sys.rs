//! Raw FFI declarations for libb200mpm.so (include/b200mpm.h). `-sys` style: no logic.
//! NOT compiled in this repository's image (no rustc); kept as the binding a wgsparkl maintainer adds.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const B200MPM_MAX_BODIES: usize = 16;
pub const B200MPM_NUM_PASSES: usize = 10;
/// Length of the arrays of `b200mpm_get_kernel_timings` (f64) and, doubled, `b200mpm_debug_timeline` (u64).
pub const B200MPM_NUM_KERNELS: usize = 15;

#[repr(C)]
#[derive(Copy, Clone, Debug, Default)]
pub struct b200mpm_sim_params {
    pub gravity: [f32; 3],
    pub dt: f32,
}

#[repr(C)]
#[derive(Copy, Clone, Debug)]
pub struct b200mpm_particle {
    pub position: [f32; 3],
    pub velocity: [f32; 3],
    pub def_grad: [f32; 9],
    pub affine: [f32; 9],
    pub cdf_normal: [f32; 3],
    pub cdf_rigid_vel: [f32; 3],
    pub cdf_signed_distance: f32,
    pub cdf_affinity: u32,
    pub init_volume: f32,
    pub init_radius: f32,
    pub mass: f32,
    pub lambda: f32,
    pub mu: f32,
    pub dp_h0: f32,
    pub dp_h1: f32,
    pub dp_h2: f32,
    pub dp_h3: f32,
    pub dp_lambda: f32,
    pub dp_mu: f32,
    pub plastic_det: f32,
    pub plastic_hardening: f32,
    pub plastic_log_vol_gain: f32,
    pub phase: f32,
    pub max_stretch: f32,
    pub model: u32,
}

#[repr(C)]
#[derive(Copy, Clone, Debug)]
pub struct b200mpm_body {
    pub shape_type: u32,
    pub shape_a: [f32; 3],
    pub shape_b: [f32; 3],
    pub radius: f32,
    pub translation: [f32; 3],
    pub rotation: [f32; 4],
    pub linvel: [f32; 3],
    pub angvel: [f32; 3],
    pub inv_mass: [f32; 3],
    pub inv_inertia: [f32; 9],
    pub local_com: [f32; 3],
    pub two_ways: u32,
}

#[repr(C)]
#[derive(Copy, Clone, Debug, Default)]
pub struct b200mpm_pose {
    pub translation: [f32; 3],
    pub rotation: [f32; 4],
}

#[repr(C)]
#[derive(Copy, Clone, Debug, Default)]
pub struct b200mpm_velocity {
    pub linear: [f32; 3],
    pub angular: [f32; 3],
}

pub enum b200mpm_pipeline {}
pub enum b200mpm_data {}

#[link(name = "b200mpm")]
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200mpm_block_info {
    pub vid: [i32; 3],
    pub first_particle: u32,
    pub num_particles: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200mpm_node {
    pub momentum_velocity_mass: [f32; 4],
    pub cdf_distance: f32,
    pub cdf_affinities: u32,
    pub cdf_closest_id: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200mpm_instance {
    pub deformation: [f32; 12],
    pub position: [f32; 4],
    pub base_color: [f32; 4],
    pub color: [f32; 4],
}

pub const B200MPM_PARTICLE_RECORD_BYTES: u32 = 128;
pub const B200MPM_HALO_BLOCK_BYTES: u32 = 1040;
pub const B200MPM_SHARD_HEADER_BYTES: u32 = 16;

extern "C" {
    pub fn b200mpm_last_error() -> *const c_char;
    pub fn b200mpm_pipeline_create(device: c_int, dim: c_int, out: *mut *mut b200mpm_pipeline) -> c_int;
    pub fn b200mpm_pipeline_destroy(p: *mut b200mpm_pipeline);
    pub fn b200mpm_pipeline_set_stream(p: *mut b200mpm_pipeline, cuda_stream: *mut c_void) -> c_int;
    pub fn b200mpm_pipeline_launch_count(p: *const b200mpm_pipeline) -> u64;
    pub fn b200mpm_data_create(
        p: *mut b200mpm_pipeline,
        params: *const b200mpm_sim_params,
        particles: *const b200mpm_particle,
        num_particles: usize,
        bodies: *const b200mpm_body,
        num_bodies: usize,
        cell_width: f32,
        grid_capacity: u32,
        out: *mut *mut b200mpm_data,
    ) -> c_int;
    pub fn b200mpm_data_destroy(d: *mut b200mpm_data);
    pub fn b200mpm_data_num_particles(d: *const b200mpm_data) -> usize;
    pub fn b200mpm_data_num_bodies(d: *const b200mpm_data) -> usize;
    pub fn b200mpm_step(p: *mut b200mpm_pipeline, d: *mut b200mpm_data, num_substeps: u32) -> c_int;
    pub fn b200mpm_sync(p: *mut b200mpm_pipeline) -> c_int;
    pub fn b200mpm_set_timestamps(p: *mut b200mpm_pipeline, enabled: c_int) -> c_int;
    pub fn b200mpm_get_timings(p: *mut b200mpm_pipeline, ms: *mut f64) -> c_int;
    pub fn b200mpm_get_kernel_timings(p: *mut b200mpm_pipeline, ms: *mut f64) -> c_int;
    pub fn b200mpm_debug_timeline(d: *mut b200mpm_data, ns: *mut u64) -> c_int;
    pub fn b200mpm_write_sim_params(d: *mut b200mpm_data, params: *const b200mpm_sim_params) -> c_int;
    pub fn b200mpm_write_body_poses(d: *mut b200mpm_data, poses: *const b200mpm_pose, n: usize) -> c_int;
    pub fn b200mpm_write_body_vels(d: *mut b200mpm_data, vels: *const b200mpm_velocity, n: usize) -> c_int;
    pub fn b200mpm_read_body_poses(d: *mut b200mpm_data, poses: *mut b200mpm_pose, n: usize) -> c_int;
    pub fn b200mpm_read_body_vels(d: *mut b200mpm_data, vels: *mut b200mpm_velocity, n: usize) -> c_int;
    pub fn b200mpm_read_positions(d: *mut b200mpm_data, out: *mut f32) -> c_int;
    pub fn b200mpm_read_positions_async(d: *mut b200mpm_data, out: *mut f32) -> c_int;
    pub fn b200mpm_read_particles(d: *mut b200mpm_data, out: *mut b200mpm_particle) -> c_int;
    pub fn b200mpm_data_status(d: *mut b200mpm_data, num_active_blocks: *mut u32) -> c_int;
    pub fn b200mpm_data_set_rigid_particles(
        d: *mut b200mpm_data,
        vertices: *const f32,
        vertex_colliders: *const u32,
        num_vertices: usize,
        samples: *const f32,
        sample_ids: *const u32,
        num_samples: usize,
    ) -> c_int;
    pub fn b200mpm_data_reserve_grid(d: *mut b200mpm_data, grid_capacity: u32) -> c_int;
    pub fn b200mpm_data_set_auto_grow(d: *mut b200mpm_data, max_load: f32) -> c_int;
    pub fn b200mpm_prep_vertex_buffer(p: *mut b200mpm_pipeline, d: *mut b200mpm_data, dev_instances: *mut b200mpm_instance, mode: u32) -> c_int;
    pub fn b200mpm_sort_only(p: *mut b200mpm_pipeline, d: *mut b200mpm_data) -> c_int;
    pub fn b200mpm_prefix_sum_u32(p: *mut b200mpm_pipeline, data: *mut u32, len: usize) -> c_int;
    pub fn b200mpm_read_grid(
        d: *mut b200mpm_data,
        blocks: *mut b200mpm_block_info,
        nodes: *mut b200mpm_node,
        capacity: usize,
        num_blocks: *mut usize,
    ) -> c_int;
    pub fn b200mpm_read_sorted_ids(d: *mut b200mpm_data, out: *mut u32) -> c_int;

    // ---- multi-GPU slab sharding (one process per GPU) ----
    pub fn b200mpm_data_create_ex(
        p: *mut b200mpm_pipeline,
        params: *const b200mpm_sim_params,
        particles: *const b200mpm_particle,
        num_particles: usize,
        particle_ids: *const u32,
        particle_capacity: usize,
        bodies: *const b200mpm_body,
        num_bodies: usize,
        cell_width: f32,
        grid_capacity: u32,
        out: *mut *mut b200mpm_data,
    ) -> c_int;
    pub fn b200mpm_slab_configure(d: *mut b200mpm_data, x_lo: i32, x_hi: i32) -> c_int;
    pub fn b200mpm_data_num_live(d: *mut b200mpm_data, num_live: *mut usize) -> c_int;
    pub fn b200mpm_shard_emigrate(p: *mut b200mpm_pipeline, d: *mut b200mpm_data, dev_left: *mut c_void, dev_right: *mut c_void, cap_records: u32) -> c_int;
    pub fn b200mpm_shard_immigrate(p: *mut b200mpm_pipeline, d: *mut b200mpm_data, dev_buffer: *const c_void, cap_records: u32) -> c_int;
    pub fn b200mpm_shard_step_begin(p: *mut b200mpm_pipeline, d: *mut b200mpm_data) -> c_int;
    pub fn b200mpm_shard_halo_pack(p: *mut b200mpm_pipeline, d: *mut b200mpm_data, dev_left: *mut c_void, dev_right: *mut c_void, cap_blocks: u32) -> c_int;
    pub fn b200mpm_shard_halo_add(p: *mut b200mpm_pipeline, d: *mut b200mpm_data, dev_buffer: *const c_void, cap_blocks: u32) -> c_int;
    pub fn b200mpm_shard_impulses(p: *mut b200mpm_pipeline, d: *mut b200mpm_data, dev_buf: *mut i32, write: c_int) -> c_int;
    pub fn b200mpm_shard_step_end(p: *mut b200mpm_pipeline, d: *mut b200mpm_data) -> c_int;
    pub fn b200mpm_nccl_unique_id(out: *mut c_void, bytes: usize) -> c_int;
    pub fn b200mpm_shard_comm_init(
        p: *mut b200mpm_pipeline,
        d: *mut b200mpm_data,
        rank: c_int,
        world: c_int,
        unique_id: *const c_void,
        migration_cap_records: u32,
        halo_cap_blocks: u32,
    ) -> c_int;
    pub fn b200mpm_shard_step(p: *mut b200mpm_pipeline, d: *mut b200mpm_data, num_substeps: u32) -> c_int;
    pub fn b200mpm_shard_p2p_export(p: *mut b200mpm_pipeline, d: *mut b200mpm_data, handle_out: *mut c_void, bytes: usize) -> c_int;
    pub fn b200mpm_shard_p2p_connect(p: *mut b200mpm_pipeline, d: *mut b200mpm_data, handles: *const c_void, num_handles: usize) -> c_int;
    pub fn b200mpm_read_positions_unordered(d: *mut b200mpm_data, out: *mut f32, capacity: usize, count: *mut usize) -> c_int;
    pub fn b200mpm_read_positions_unordered_async(d: *mut b200mpm_data, out: *mut f32, capacity: usize, count: *mut usize) -> c_int;
    pub fn b200mpm_read_particles_unordered(d: *mut b200mpm_data, out: *mut b200mpm_particle, ids: *mut u32, capacity: usize, count: *mut usize) -> c_int;
}
