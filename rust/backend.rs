//! Drop-in replacement of `wgsparkl::pipeline::{MpmPipeline, MpmData}` (src/pipeline.rs) over the B200 C ABI.
//! Same type names, same constructor arguments (minus the wgpu `Device`), same per-frame protocol as
//! src_testbed/step.rs. NOT compiled in this repository's image (no rustc).
//!
//! In wgsparkl this file would live at `src/pipeline_b200.rs` behind `#[cfg(feature = "b200")]`, with
//! `pub use pipeline_b200::{MpmData, MpmPipeline};` replacing the wgpu versions in `src/lib.rs`.
use crate::models::DruckerPrager;
use crate::solver::{Particle, ParticlePhase, SimulationParams};
use rapier::dynamics::RigidBodySet;
use rapier::geometry::{ColliderSet, ShapeType};
use std::ffi::CStr;
use std::ptr;
use wgrapier::dynamics::body::{BodyCoupling, BodyCouplingEntry};
use wgsparkl_b200_sys as sys;

#[derive(Debug)]
pub struct B200Error(pub i32, pub String);

fn check(code: i32) -> Result<(), B200Error> {
    if code == 0 {
        Ok(())
    } else {
        let msg = unsafe { CStr::from_ptr(sys::b200mpm_last_error()) };
        Err(B200Error(code, msg.to_string_lossy().into_owned()))
    }
}

#[cfg(feature = "dim2")]
const DIM: i32 = 2;
#[cfg(feature = "dim3")]
const DIM: i32 = 3;

pub struct MpmPipeline {
    raw: *mut sys::b200mpm_pipeline,
}

pub struct MpmData {
    raw: *mut sys::b200mpm_data,
    coupling: Vec<BodyCouplingEntry>,
}

impl MpmPipeline {
    /// `MpmPipeline::new(&Device) -> Result<Self, ComposerError>` (src/pipeline.rs:176-193).
    pub fn new(cuda_device: i32) -> Result<Self, B200Error> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::b200mpm_pipeline_create(cuda_device, DIM, &mut raw) })?;
        Ok(Self { raw })
    }

    /// `queue_step` + `for _ in 0..num_substeps { queue.encode(..) }` + `submit`
    /// (src/pipeline.rs:195-281; src_testbed/step.rs:122-128,169). Asynchronous.
    pub fn queue_step(&self, data: &mut MpmData, num_substeps: u32, add_timestamps: bool) -> Result<(), B200Error> {
        check(unsafe { sys::b200mpm_set_timestamps(self.raw, add_timestamps as i32) })?;
        check(unsafe { sys::b200mpm_step(self.raw, data.raw, num_substeps) })
    }

    /// Per-pass milliseconds in the order of `Timestamps` (src_testbed/lib.rs:133-146).
    pub fn timings_ms(&self) -> Result<[f64; sys::B200MPM_NUM_PASSES], B200Error> {
        let mut ms = [0.0; sys::B200MPM_NUM_PASSES];
        check(unsafe { sys::b200mpm_get_timings(self.raw, ms.as_mut_ptr()) })?;
        Ok(ms)
    }
}

impl Drop for MpmPipeline {
    fn drop(&mut self) {
        unsafe { sys::b200mpm_pipeline_destroy(self.raw) }
    }
}

fn flatten_particle(p: &Particle) -> sys::b200mpm_particle {
    // GpuParticles::from_particles / GpuModels::from_particles (particle3d.rs:192-210, models/mod.rs:20-49)
    let mut o: sys::b200mpm_particle = unsafe { std::mem::zeroed() };
    o.position[..p.position.len()].copy_from_slice(p.position.as_slice());
    o.velocity[..p.dynamics.velocity.len()].copy_from_slice(p.dynamics.velocity.as_slice());
    o.def_grad[..p.dynamics.def_grad.len()].copy_from_slice(p.dynamics.def_grad.as_slice()); // column-major
    o.affine[..p.dynamics.affine.len()].copy_from_slice(p.dynamics.affine.as_slice());
    o.cdf_normal[..p.dynamics.cdf.normal.len()].copy_from_slice(p.dynamics.cdf.normal.as_slice());
    o.cdf_rigid_vel[..p.dynamics.cdf.rigid_vel.len()].copy_from_slice(p.dynamics.cdf.rigid_vel.as_slice());
    o.cdf_signed_distance = p.dynamics.cdf.signed_distance;
    o.cdf_affinity = p.dynamics.cdf.affinity;
    o.init_volume = p.dynamics.init_volume;
    o.init_radius = p.dynamics.init_radius;
    o.mass = p.dynamics.mass;
    o.lambda = p.model.lambda;
    o.mu = p.model.mu;
    let dp = p.plasticity.unwrap_or(DruckerPrager::new(-1.0, -1.0));
    (o.dp_h0, o.dp_h1, o.dp_h2, o.dp_h3, o.dp_lambda, o.dp_mu) = (dp.h0, dp.h1, dp.h2, dp.h3, dp.lambda, dp.mu);
    (o.plastic_det, o.plastic_hardening, o.plastic_log_vol_gain) = (1.0, 1.0, 0.0);
    let ph = p.phase.unwrap_or(ParticlePhase { phase: 0.0, max_stretch: -1.0 });
    (o.phase, o.max_stretch) = (ph.phase, ph.max_stretch);
    o.model = 0; // corotated, the reference's hard-wired model (particle_update.wgsl:7-8)
    o
}

fn flatten_body(bodies: &RigidBodySet, colliders: &ColliderSet, c: &BodyCouplingEntry) -> sys::b200mpm_body {
    // What GpuBodySet::from_rapier uploads per coupled collider (src/pipeline.rs:141).
    let rb = &bodies[c.body];
    let co = &colliders[c.collider];
    let mut o: sys::b200mpm_body = unsafe { std::mem::zeroed() };
    match co.shape().shape_type() {
        ShapeType::Ball => {
            o.shape_type = 0;
            o.radius = co.shape().as_ball().unwrap().radius;
        }
        ShapeType::Cuboid => {
            o.shape_type = 1;
            let he = co.shape().as_cuboid().unwrap().half_extents;
            o.shape_a[..he.len()].copy_from_slice(he.as_slice());
        }
        ShapeType::Capsule => {
            o.shape_type = 2;
            let cap = co.shape().as_capsule().unwrap();
            o.shape_a[..cap.segment.a.coords.len()].copy_from_slice(cap.segment.a.coords.as_slice());
            o.shape_b[..cap.segment.b.coords.len()].copy_from_slice(cap.segment.b.coords.as_slice());
            o.radius = cap.radius;
        }
        // Mesh colliders have no analytic projection: they act through their sample points (rigid_particles()).
        #[cfg(feature = "dim3")]
        ShapeType::TriMesh | ShapeType::HeightField => o.shape_type = 3,
        #[cfg(feature = "dim2")]
        ShapeType::Polyline => o.shape_type = 4,
        other => panic!("collider shape {:?} is not on the B200 path", other),
    }
    let pos = co.position();
    o.translation[..pos.translation.vector.len()].copy_from_slice(pos.translation.vector.as_slice());
    #[cfg(feature = "dim3")]
    o.rotation.copy_from_slice(pos.rotation.coords.as_slice()); // (i, j, k, w)
    #[cfg(feature = "dim2")]
    {
        o.rotation[0] = pos.rotation.re;
        o.rotation[1] = pos.rotation.im;
    }
    o.linvel[..rb.linvel().len()].copy_from_slice(rb.linvel().as_slice());
    #[cfg(feature = "dim3")]
    o.angvel.copy_from_slice(rb.angvel().as_slice());
    #[cfg(feature = "dim2")]
    {
        o.angvel[0] = rb.angvel();
    }
    let mp = rb.mass_properties();
    let inv_mass = mp.effective_inv_mass; // zero for fixed / kinematic bodies
    o.inv_mass[..inv_mass.len()].copy_from_slice(inv_mass.as_slice());
    #[cfg(feature = "dim3")]
    {
        // local-frame inverse inertia tensor, column-major
        let inv = mp.local_mprops.reconstruct_inverse_inertia_matrix();
        o.inv_inertia.copy_from_slice(inv.as_slice());
        if !rb.is_dynamic() {
            o.inv_inertia = [0.0; 9];
        }
    }
    #[cfg(feature = "dim2")]
    {
        let s = mp.local_mprops.inv_principal_inertia_sqrt;
        o.inv_inertia[0] = if rb.is_dynamic() { s * s } else { 0.0 };
    }
    let com = mp.local_mprops.local_com;
    o.local_com[..com.coords.len()].copy_from_slice(com.coords.as_slice());
    o.two_ways = matches!(c.mode, BodyCoupling::TwoWays) as u32;
    o
}

impl MpmData {
    /// `MpmData::new` (src/pipeline.rs:98-128): every collider with a parent is coupled, TwoWays.
    pub fn new(
        pipeline: &MpmPipeline,
        params: SimulationParams,
        particles: &[Particle],
        bodies: &RigidBodySet,
        colliders: &ColliderSet,
        cell_width: f32,
        grid_capacity: u32,
    ) -> Result<Self, B200Error> {
        let coupling: Vec<_> = colliders
            .iter()
            .filter_map(|(co_handle, co)| {
                let rb_handle = co.parent()?;
                Some(BodyCouplingEntry { body: rb_handle, collider: co_handle, mode: BodyCoupling::TwoWays })
            })
            .collect();
        Self::with_select_coupling(pipeline, params, particles, bodies, colliders, coupling, cell_width, grid_capacity)
    }

    /// `MpmData::with_select_coupling` (src/pipeline.rs:130-168).
    pub fn with_select_coupling(
        pipeline: &MpmPipeline,
        params: SimulationParams,
        particles: &[Particle],
        bodies: &RigidBodySet,
        colliders: &ColliderSet,
        coupling: Vec<BodyCouplingEntry>,
        cell_width: f32,
        grid_capacity: u32,
    ) -> Result<Self, B200Error> {
        let flat: Vec<_> = particles.iter().map(flatten_particle).collect();
        let flat_bodies: Vec<_> = coupling.iter().map(|c| flatten_body(bodies, colliders, c)).collect();
        let mut sp = sys::b200mpm_sim_params::default();
        sp.gravity[..params.gravity.len()].copy_from_slice(params.gravity.as_slice());
        sp.dt = params.dt;
        let mut raw = ptr::null_mut();
        check(unsafe {
            sys::b200mpm_data_create(
                pipeline.raw,
                &sp,
                flat.as_ptr(),
                flat.len(),
                flat_bodies.as_ptr(),
                flat_bodies.len(),
                cell_width,
                grid_capacity,
                &mut raw,
            )
        })?;
        // GpuRigidParticles::from_rapier (src/solver/particle3d.rs:101-160 / particle2d.rs:80-140): the crate's own
        // CPU sampling (sample_mesh / sample_polyline with sampling_step = cell_width, pipeline.rs:140) stays as it
        // is; only its output is flattened - vertices + collider index per vertex, sample points + (primitive
        // vertex ids, collider index) per sample.
        let rp = crate::solver::sample_rigid_particles(colliders, &coupling, cell_width);
        if !rp.samples.is_empty() {
            check(unsafe {
                sys::b200mpm_data_set_rigid_particles(
                    raw,
                    rp.vertices.as_ptr() as *const f32,
                    rp.vertex_colliders.as_ptr(),
                    rp.vertex_colliders.len(),
                    rp.samples.as_ptr() as *const f32,
                    rp.sample_ids.as_ptr() as *const u32,
                    rp.sample_ids.len(),
                )
            })?;
        }
        Ok(Self { raw, coupling })
    }

    pub fn coupling(&self) -> &[BodyCouplingEntry] {
        &self.coupling
    }

    /// `queue.write_buffer(bodies.poses(), ..)` (src_testbed/step.rs:92-96).
    pub fn write_body_poses(&mut self, poses: &[sys::b200mpm_pose]) -> Result<(), B200Error> {
        check(unsafe { sys::b200mpm_write_body_poses(self.raw, poses.as_ptr(), poses.len()) })
    }
    /// `queue.write_buffer(bodies.vels(), ..)` (src_testbed/step.rs:98-119).
    pub fn write_body_vels(&mut self, vels: &[sys::b200mpm_velocity]) -> Result<(), B200Error> {
        check(unsafe { sys::b200mpm_write_body_vels(self.raw, vels.as_ptr(), vels.len()) })
    }
    /// `block_on(poses_staging.read(device))` (src_testbed/step.rs:175-176).
    pub fn read_body_poses(&mut self) -> Result<Vec<sys::b200mpm_pose>, B200Error> {
        let n = unsafe { sys::b200mpm_data_num_bodies(self.raw) };
        let mut out = vec![sys::b200mpm_pose::default(); n];
        check(unsafe { sys::b200mpm_read_body_poses(self.raw, out.as_mut_ptr(), n) })?;
        Ok(out)
    }
    /// `sim_params.params` rewrite from the UI sliders (src_testbed/ui.rs:98-103).
    pub fn write_sim_params(&mut self, params: SimulationParams) -> Result<(), B200Error> {
        let mut sp = sys::b200mpm_sim_params::default();
        sp.gravity[..params.gravity.len()].copy_from_slice(params.gravity.as_slice());
        sp.dt = params.dt;
        check(unsafe { sys::b200mpm_write_sim_params(self.raw, &sp) })
    }
}

impl Drop for MpmData {
    fn drop(&mut self) {
        unsafe { sys::b200mpm_data_destroy(self.raw) }
    }
}
