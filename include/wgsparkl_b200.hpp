// wgsparkl_b200.hpp — C++ host-side mirror of wgsparkl's public interface for the MPM substep hot path,
// layered on the C ABI of include/b200mpm.h (header-only, no CUDA / torch types).
//
// The reference's host code is Rust; there is no Rust toolchain in this image, so the typed host layer
// above the C ABI is written in C++ with the reference's names, argument meaning and error behaviour:
//
//   wgsparkl::models::ElasticCoefficients::from_young_modulus   (src/models/mod.rs:70-75)
//   wgsparkl::models::DruckerPrager::new_                        (src/models/drucker_prager.rs:18-33)
//   wgsparkl::solver::ParticleDynamics::with_density             (src/solver/particle3d.rs:28-42)
//   wgsparkl::solver::{Particle, ParticlePhase, SimulationParams}
//   wgsparkl::pipeline::MpmPipeline::{new_, queue_step}          (src/pipeline.rs:176-281)
//   wgsparkl::pipeline::MpmData::{new_, with_select_coupling}    (src/pipeline.rs:98-172)
//
// The Rust binding a wgsparkl maintainer would add is rust/wgsparkl_b200_sys.rs + rust/backend.rs
// (INTEGRATION.md); it has the same shape as this file.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <optional>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <utility>
#include <vector>

#include "b200mpm.h"

namespace wgsparkl {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int code) {
    if (code != B200MPM_OK) throw Error(code, b200mpm_last_error());
}

namespace models {

// lame_lambda_mu (src/models/mod.rs:52-61), f32 arithmetic like the original.
inline void lame_lambda_mu(float young_modulus, float poisson_ratio, float& lambda, float& mu) {
    lambda = young_modulus * poisson_ratio / ((1.0f + poisson_ratio) * (1.0f - 2.0f * poisson_ratio));
    mu = young_modulus / (2.0f * (1.0f + poisson_ratio));
}

struct ElasticCoefficients { // src/models/mod.rs:63-75
    float lambda = 0.0f, mu = 0.0f;
    static ElasticCoefficients from_young_modulus(float young_modulus, float poisson_ratio) {
        ElasticCoefficients e;
        lame_lambda_mu(young_modulus, poisson_ratio, e.lambda, e.mu);
        return e;
    }
};

struct DruckerPrager { // src/models/drucker_prager.rs:6-33
    float h0, h1, h2, h3, lambda, mu;
    static DruckerPrager new_(float young_modulus, float poisson_ratio) {
        DruckerPrager d;
        if (young_modulus > 0.0f) lame_lambda_mu(young_modulus, poisson_ratio, d.lambda, d.mu);
        else d.lambda = d.mu = -1.0f;
        const float to_rad = 3.14159265358979323846f / 180.0f;
        d.h0 = 35.0f * to_rad;
        d.h1 = 9.0f * to_rad;
        d.h2 = 0.2f;
        d.h3 = 10.0f * to_rad;
        return d;
    }
};

struct DruckerPragerPlasticState { // src/models/drucker_prager.rs:36-52
    float plastic_deformation_gradient_det = 1.0f, plastic_hardening = 1.0f, log_vol_gain = 0.0f;
};

} // namespace models

namespace solver {

template <int DIM>
struct SimulationParamsT { // src/solver/params.rs:6-16
    float gravity[DIM];
    float dt;
    b200mpm_sim_params to_abi() const {
        b200mpm_sim_params p{};
        for (int i = 0; i < DIM; ++i) p.gravity[i] = gravity[i];
        p.dt = dt;
        return p;
    }
};

struct ParticlePhase { // src/solver/particle_update.rs:37-42
    float phase, max_stretch;
};

template <int DIM>
struct Cdf { // src/solver/particle3d.rs:43-50
    float normal[DIM] = {}, rigid_vel[DIM] = {};
    float signed_distance = 0.0f;
    uint32_t affinity = 0;
};

template <int DIM>
struct ParticleDynamics { // src/solver/particle3d.rs:16-42
    float velocity[DIM] = {};
    float def_grad[DIM * DIM] = {}; // column-major
    float affine[DIM * DIM] = {};
    Cdf<DIM> cdf;
    float init_volume = 0.0f, init_radius = 0.0f, mass = 0.0f;
    static ParticleDynamics with_density(float radius, float density) {
        ParticleDynamics d;
        float v = radius * 2.0f, vol = 1.0f;
        for (int i = 0; i < DIM; ++i) vol *= v; // powi(exponent): the particles are square-ish
        for (int i = 0; i < DIM; ++i) d.def_grad[i * DIM + i] = 1.0f;
        d.init_volume = vol;
        d.init_radius = radius;
        d.mass = vol * density;
        return d;
    }
};

template <int DIM>
struct Particle { // src/solver/particle3d.rs:53-60
    float position[DIM] = {};
    ParticleDynamics<DIM> dynamics;
    models::ElasticCoefficients model;
    std::optional<models::DruckerPrager> plasticity;
    std::optional<ParticlePhase> phase;
    uint32_t model_kind = B200MPM_MODEL_COROTATED; // additive selector (b200mpm.h)

    // GpuParticles::from_particles + GpuModels::from_particles (particle3d.rs:192-210, models/mod.rs:20-49)
    b200mpm_particle to_abi() const {
        b200mpm_particle o;
        std::memset(&o, 0, sizeof(o));
        for (int i = 0; i < DIM; ++i) {
            o.position[i] = position[i];
            o.velocity[i] = dynamics.velocity[i];
            o.cdf_normal[i] = dynamics.cdf.normal[i];
            o.cdf_rigid_vel[i] = dynamics.cdf.rigid_vel[i];
        }
        for (int i = 0; i < DIM * DIM; ++i) {
            o.def_grad[i] = dynamics.def_grad[i];
            o.affine[i] = dynamics.affine[i];
        }
        o.cdf_signed_distance = dynamics.cdf.signed_distance;
        o.cdf_affinity = dynamics.cdf.affinity;
        o.init_volume = dynamics.init_volume;
        o.init_radius = dynamics.init_radius;
        o.mass = dynamics.mass;
        o.lambda = model.lambda;
        o.mu = model.mu;
        const models::DruckerPrager dp = plasticity.value_or(models::DruckerPrager::new_(-1.0f, -1.0f));
        o.dp_h0 = dp.h0, o.dp_h1 = dp.h1, o.dp_h2 = dp.h2, o.dp_h3 = dp.h3, o.dp_lambda = dp.lambda, o.dp_mu = dp.mu;
        const models::DruckerPragerPlasticState st;
        o.plastic_det = st.plastic_deformation_gradient_det;
        o.plastic_hardening = st.plastic_hardening;
        o.plastic_log_vol_gain = st.log_vol_gain;
        const ParticlePhase ph = phase.value_or(ParticlePhase{0.0f, -1.0f});
        o.phase = ph.phase;
        o.max_stretch = ph.max_stretch;
        o.model = model_kind;
        return o;
    }
};

// ---- rigid particles of mesh colliders (GpuRigidParticles::from_rapier, src/solver/particle3d.rs:101-160) --------
// CPU sampling of triangle meshes: sample_mesh / sample_triangle / sample_edge (particle3d.rs:251-428) and of
// polylines (particle2d.rs:206-234). The four arrays are what b200mpm_data_set_rigid_particles takes.
struct RigidParticles {
    std::vector<float> vertices; // 3 per collider vertex, local frame (2D: z = 0)
    std::vector<uint32_t> vertex_colliders;
    std::vector<float> samples; // 3 per sample point, local frame
    std::vector<uint32_t> sample_ids; // 4 per sample point: primitive vertex ids (global) + collider index
};

namespace detail {
struct V3f {
    float x, y, z;
};
inline V3f operator+(V3f a, V3f b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3f operator-(V3f a, V3f b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3f operator*(V3f a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(V3f a, V3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float norm(V3f a) { return std::sqrt(dot(a, a)); }
constexpr float kSamplingEps = 1.0e-5f; // particle3d.rs:241

template <class Emit>
inline void sample_edge(V3f a, V3f b, float xy_spacing, Emit&& emit) { // particle3d.rs:300-320
    const V3f ab = b - a;
    const float edge_length = norm(ab);
    if (edge_length > kSamplingEps) {
        const V3f edge_dir = ab * (1.0f / edge_length);
        const float spacing = xy_spacing / std::sqrt(2.0f);
        const size_t nsteps = (size_t)std::ceil(edge_length / spacing);
        for (size_t i = 1; i < nsteps; ++i) emit(a + edge_dir * (spacing * (float)i));
    }
}
template <class Emit>
inline void sample_triangle(V3f a, V3f b, V3f c, float xy_spacing, Emit&& emit) { // particle3d.rs:336-428
    const float d_ab = norm(b - a), d_bc = norm(c - b), d_ca = norm(a - c);
    const float mx = std::max(d_ab, std::max(d_bc, d_ca));
    if (mx == d_bc) {
        const V3f t = a;
        a = b, b = c, c = t;
    } else if (mx == d_ca) {
        const V3f t = c;
        c = b, b = a, a = t;
    }
    const V3f ac = c - a, base = b - a;
    const float base_length = norm(base);
    if (!(base_length > 0.0f)) return;
    const V3f base_dir = base * (1.0f / base_length);
    const float spacing = xy_spacing / std::sqrt(2.0f);
    const float base_step_count = std::ceil(base_length / spacing);
    const V3f base_step = base_dir * spacing;
    const float ac_offset_length = dot(ac, base_dir);
    const float bc_offset_length = base_length - ac_offset_length;
    if (ac_offset_length < kSamplingEps || bc_offset_length < kSamplingEps || base_length < kSamplingEps) return;
    const V3f height = ac - base_dir * ac_offset_length;
    const float height_length = norm(height);
    const V3f height_dir = height * (1.0f / height_length);
    const float tan_alpha = height_length / ac_offset_length, tan_beta = height_length / bc_offset_length;
    for (uint32_t i = 1; i < (uint32_t)base_step_count; ++i) {
        const V3f base_position = a + base_step * (float)i;
        const float hl = std::min(tan_alpha * norm(base_position - a), tan_beta * norm(base_position - b));
        const float height_step_count = std::ceil(hl / spacing);
        const V3f height_step = height_dir * spacing;
        for (uint32_t j = 1; j < (uint32_t)height_step_count; ++j) {
            const V3f p = base_position + height_step * (float)j;
            if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z)) emit(p);
        }
    }
}
} // namespace detail

// Appends the trimesh collider `collider_id` (local-frame vertices, 3 floats each; triangles, 3 indices each) with
// the reference's sampling step (= cell_width, pipeline.rs:140).
inline void sample_trimesh(RigidParticles& out, uint32_t collider_id, const std::vector<float>& vertices,
                           const std::vector<uint32_t>& triangles, float sampling_step) {
    using detail::V3f;
    const uint32_t base_vid = (uint32_t)out.vertex_colliders.size();
    const size_t nv = vertices.size() / 3;
    for (size_t i = 0; i < nv; ++i) {
        for (int k = 0; k < 3; ++k) out.vertices.push_back(vertices[3 * i + k]);
        out.vertex_colliders.push_back(collider_id);
    }
    auto vtx = [&](uint32_t i) { return V3f{vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]}; };
    std::vector<std::pair<uint32_t, uint32_t>> visited; // edges already sampled (particle3d.rs:262-270)
    auto edge_needs_sampling = [&](uint32_t ia, uint32_t ib) {
        if (ib > ia) std::swap(ia, ib);
        const std::pair<uint32_t, uint32_t> key(ia, ib);
        if (std::find(visited.begin(), visited.end(), key) != visited.end()) return false;
        visited.push_back(key);
        return true;
    };
    for (size_t t = 0; t + 2 < triangles.size(); t += 3) {
        const uint32_t ia = triangles[t], ib = triangles[t + 1], ic = triangles[t + 2];
        auto emit = [&](V3f p) {
            out.samples.push_back(p.x), out.samples.push_back(p.y), out.samples.push_back(p.z);
            out.sample_ids.push_back(base_vid + ia), out.sample_ids.push_back(base_vid + ib);
            out.sample_ids.push_back(base_vid + ic), out.sample_ids.push_back(collider_id);
        };
        detail::sample_triangle(vtx(ia), vtx(ib), vtx(ic), sampling_step, emit);
        if (edge_needs_sampling(ia, ib)) detail::sample_edge(vtx(ia), vtx(ib), sampling_step, emit);
        if (edge_needs_sampling(ib, ic)) detail::sample_edge(vtx(ib), vtx(ic), sampling_step, emit);
        if (edge_needs_sampling(ic, ia)) detail::sample_edge(vtx(ic), vtx(ia), sampling_step, emit);
    }
}

} // namespace solver

namespace pipeline {

template <int DIM>
class MpmData;

// MpmPipeline (src/pipeline.rs:24-39). `new_` can fail (the reference returns Result<_, ComposerError>).
template <int DIM>
class MpmPipeline {
public:
    static MpmPipeline new_(int cuda_device = 0) {
        MpmPipeline p;
        check(b200mpm_pipeline_create(cuda_device, DIM, &p.h_));
        return p;
    }
    MpmPipeline(MpmPipeline&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    MpmPipeline& operator=(MpmPipeline&& o) noexcept {
        std::swap(h_, o.h_);
        return *this;
    }
    MpmPipeline(const MpmPipeline&) = delete;
    ~MpmPipeline() { b200mpm_pipeline_destroy(h_); }

    // queue_step(&self, &mut MpmData, &mut KernelInvocationQueue, add_timestamps) followed by
    // `for _ in 0..num_substeps { queue.encode(..) }` and `submit` (pipeline.rs:195-281, step.rs:122-128,169).
    void queue_step(MpmData<DIM>& data, uint32_t num_substeps, bool add_timestamps = false);
    void sync() { check(b200mpm_sync(h_)); }
    void set_stream(void* cuda_stream) { check(b200mpm_pipeline_set_stream(h_, cuda_stream)); }
    std::vector<double> timings_ms() {
        std::vector<double> ms(B200MPM_NUM_PASSES);
        check(b200mpm_get_timings(h_, ms.data()));
        return ms;
    }
    // WgPrepVertexBuffer::queue (src_testbed/prep_vertex_buffer.rs:81-113): fills the renderer's instance buffer
    // (device memory, num_particles x b200mpm_instance) from the device state.
    void prep_vertex_buffer(MpmData<DIM>& data, b200mpm_instance* dev_instances, uint32_t mode = B200MPM_RENDER_DEFAULT);
    b200mpm_pipeline* raw() { return h_; }

private:
    MpmPipeline() = default;
    b200mpm_pipeline* h_ = nullptr;
};

// MpmData (src/pipeline.rs:84-172).
template <int DIM>
class MpmData {
public:
    // MpmData::with_select_coupling: `bodies` is what GpuBodySet::from_rapier extracts per BodyCouplingEntry.
    static MpmData with_select_coupling(MpmPipeline<DIM>& pipeline, const solver::SimulationParamsT<DIM>& params,
                                        const std::vector<solver::Particle<DIM>>& particles,
                                        const std::vector<b200mpm_body>& bodies, float cell_width, uint32_t grid_capacity) {
        std::vector<b200mpm_particle> flat;
        flat.reserve(particles.size());
        for (const auto& p : particles) flat.push_back(p.to_abi());
        MpmData d;
        const b200mpm_sim_params sp = params.to_abi();
        check(b200mpm_data_create(pipeline.raw(), &sp, flat.data(), flat.size(), bodies.data(), bodies.size(), cell_width,
                                  grid_capacity, &d.h_));
        return d;
    }
    MpmData(MpmData&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    MpmData(const MpmData&) = delete;
    ~MpmData() { b200mpm_data_destroy(h_); }

    size_t num_particles() const { return b200mpm_data_num_particles(h_); }
    void write_sim_params(const solver::SimulationParamsT<DIM>& p) { // ui.rs:98-103
        const b200mpm_sim_params sp = p.to_abi();
        check(b200mpm_write_sim_params(h_, &sp));
    }
    void write_body_poses(const std::vector<b200mpm_pose>& poses) { check(b200mpm_write_body_poses(h_, poses.data(), poses.size())); }
    void write_body_vels(const std::vector<b200mpm_velocity>& vels) { check(b200mpm_write_body_vels(h_, vels.data(), vels.size())); }
    std::vector<b200mpm_pose> read_body_poses() { // poses_staging.read (step.rs:175-176)
        std::vector<b200mpm_pose> out(b200mpm_data_num_bodies(h_));
        check(b200mpm_read_body_poses(h_, out.data(), out.size()));
        return out;
    }
    std::vector<b200mpm_particle> read_particles() {
        std::vector<b200mpm_particle> out(num_particles());
        check(b200mpm_read_particles(h_, out.data()));
        return out;
    }
    // particles.positions (particle3d.rs:177) as vec4 per particle; blocking / staged like map_async (valid after
    // MpmPipeline::sync; `out_pinned` should be page-locked).
    std::vector<float> read_positions() {
        std::vector<float> out(4 * num_particles());
        check(b200mpm_read_positions(h_, out.data()));
        return out;
    }
    void read_positions_async(float* out_pinned) { check(b200mpm_read_positions_async(h_, out_pinned)); }
    // GpuRigidParticles::from_rapier (particle3d.rs:101-160): sample points of the trimesh / polyline colliders,
    // local frames; ids = (vertex a, b, c, collider index) per sample point. Once, before stepping.
    void set_rigid_particles(const std::vector<float>& vertices, const std::vector<uint32_t>& vertex_colliders,
                             const std::vector<float>& samples, const std::vector<uint32_t>& sample_ids) {
        check(b200mpm_data_set_rigid_particles(h_, vertices.data(), vertex_colliders.data(), vertex_colliders.size(),
                                               samples.data(), sample_ids.data(), sample_ids.size() / 4));
    }
    // The resize the reference leaves as a stub (grid.rs:43-118).
    void reserve_grid(uint32_t grid_capacity) { check(b200mpm_data_reserve_grid(h_, grid_capacity)); }
    void set_auto_grow(float max_load) { check(b200mpm_data_set_auto_grow(h_, max_load)); }
    b200mpm_data* raw() { return h_; }

private:
    MpmData() = default;
    b200mpm_data* h_ = nullptr;
};

template <int DIM>
inline void MpmPipeline<DIM>::prep_vertex_buffer(MpmData<DIM>& data, b200mpm_instance* dev_instances, uint32_t mode) {
    check(b200mpm_prep_vertex_buffer(h_, data.raw(), dev_instances, mode));
}

template <int DIM>
inline void MpmPipeline<DIM>::queue_step(MpmData<DIM>& data, uint32_t num_substeps, bool add_timestamps) {
    check(b200mpm_set_timestamps(h_, add_timestamps ? 1 : 0));
    check(b200mpm_step(h_, data.raw(), num_substeps));
}

} // namespace pipeline
} // namespace wgsparkl
