// ORACLE — TEST INFRASTRUCTURE ONLY (see mpm_oracle.hpp). C entry points for ctypes.
// Parity status: UNPINNED except the prefix sum (see header of mpm_oracle.hpp).
#include <cstdio>
#include <cstdlib>

#include "mpm_oracle.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace oracle;

namespace {

template <int D>
Mat<D> mat_from(const float* p) {
    Mat<D> m;
    for (int c = 0; c < D; ++c)
        for (int r = 0; r < D; ++r) m.c[c].v[r] = p[c * D + r];
    return m;
}
template <int D>
void mat_to(const Mat<D>& m, float* p) {
    for (int i = 0; i < 9; ++i) p[i] = 0.0f;
    for (int c = 0; c < D; ++c)
        for (int r = 0; r < D; ++r) p[c * D + r] = m.c[c].v[r];
}
template <int D>
Vec<D> vec_from(const float* p) {
    Vec<D> v;
    for (int i = 0; i < D; ++i) v[i] = p[i];
    return v;
}
template <int D>
void vec_to(const Vec<D>& v, float* p) {
    for (int i = 0; i < 3; ++i) p[i] = 0.0f;
    for (int i = 0; i < D; ++i) p[i] = v[i];
}

struct Handle {
    int dim;
    Sim<2>* s2 = nullptr;
    Sim<3>* s3 = nullptr;
};

template <int D>
void set_body(Sim<D>& s, size_t i, const b200mpm_body& b) {
    Shape<D> sh;
    sh.type = b.shape_type;
    sh.a = vec_from<D>(b.shape_a);
    sh.b = vec_from<D>(b.shape_b);
    sh.radius = b.radius;
    s.collision_shapes[i] = sh;
    Pose<D> p;
    p.R = rotation_from(b.rotation, std::integral_constant<int, D>{});
    p.t = vec_from<D>(b.translation);
    s.poses[i] = p;
    for (int k = 0; k < 4; ++k) s.pose_rot_raw[i * 4 + k] = b.rotation[k];
    Velocity<D> v;
    v.linear = vec_from<D>(b.linvel);
    if constexpr (D == 2) v.angular = b.angvel[0];
    else v.angular = vec_from<3>(b.angvel);
    s.body_vels[i] = v;
    MassProperties<D> mp;
    mp.com = vec_from<D>(b.local_com);
    mp.inv_mass = vec_from<D>(b.inv_mass);
    if constexpr (D == 2) {
        mp.inv_inertia = Mat<2>::zero();
        mp.inv_inertia.at(0, 0) = b.inv_inertia[0];
    } else {
        mp.inv_inertia = mat_from<3>(b.inv_inertia);
    }
    if (!b.two_ways) {
        mp.inv_mass = Vec<D>::zero();
        mp.inv_inertia = Mat<D>::zero();
    }
    s.local_mprops[i] = mp;
    s.mprops[i] = mp;
    IntegerImpulse<D> imp;
    imp.com = Vec<D>::zero();
    for (int k = 0; k < D; ++k) imp.linear[k] = 0;
    for (int k = 0; k < 3; ++k) imp.angular[k] = 0;
    s.body_impulses[i] = imp;
}

template <int D>
Sim<D>* build(const b200mpm_sim_params* params, const b200mpm_particle* parts, size_t n, const b200mpm_body* bodies,
              size_t nb, float cell_width, uint32_t grid_capacity) {
    auto* s = new Sim<D>();
    s->gravity = vec_from<D>(params->gravity);
    s->dt = params->dt;
    s->init_grid(grid_capacity, cell_width);
    s->particles_pos.resize(n);
    s->particles_dyn.resize(n);
    s->constitutive_model.resize(n);
    s->plasticity.resize(n);
    s->plastic_state.resize(n);
    s->phases.resize(n);
    s->model_kind.resize(n);
    s->sorted_particle_ids.assign(n, 0);
    s->particle_node_linked_lists.assign(n, NONE);
    for (size_t i = 0; i < n; ++i) {
        const b200mpm_particle& p = parts[i];
        s->particles_pos[i] = vec_from<D>(p.position);
        Dynamics<D>& d = s->particles_dyn[i];
        d.velocity = vec_from<D>(p.velocity);
        d.def_grad = mat_from<D>(p.def_grad);
        d.affine = mat_from<D>(p.affine);
        d.cdf.normal = vec_from<D>(p.cdf_normal);
        d.cdf.rigid_vel = vec_from<D>(p.cdf_rigid_vel);
        d.cdf.signed_distance = p.cdf_signed_distance;
        d.cdf.affinity = p.cdf_affinity;
        d.init_volume = p.init_volume;
        d.init_radius = p.init_radius;
        d.mass = p.mass;
        s->constitutive_model[i] = ElasticCoefficients{p.lambda, p.mu};
        s->plasticity[i] = Plasticity{p.dp_h0, p.dp_h1, p.dp_h2, p.dp_h3, p.dp_lambda, p.dp_mu};
        s->plastic_state[i] = PlasticState{p.plastic_det, p.plastic_hardening, p.plastic_log_vol_gain};
        s->phases[i] = Phase{p.phase, p.max_stretch};
        s->model_kind[i] = p.model;
    }
    s->collision_shapes.resize(nb);
    s->poses.resize(nb);
    s->pose_rot_raw.resize(nb * 4);
    s->body_vels.resize(nb);
    s->local_mprops.resize(nb);
    s->mprops.resize(nb);
    s->body_impulses.resize(nb);
    for (size_t i = 0; i < nb; ++i) set_body<D>(*s, i, bodies[i]);
    return s;
}

template <int D>
void read_particles(const Sim<D>& s, b200mpm_particle* out) {
    for (size_t i = 0; i < s.num_particles(); ++i) {
        b200mpm_particle& p = out[i];
        std::memset(&p, 0, sizeof(p));
        const Dynamics<D>& d = s.particles_dyn[i];
        vec_to<D>(s.particles_pos[i], p.position);
        vec_to<D>(d.velocity, p.velocity);
        mat_to<D>(d.def_grad, p.def_grad);
        mat_to<D>(d.affine, p.affine);
        vec_to<D>(d.cdf.normal, p.cdf_normal);
        vec_to<D>(d.cdf.rigid_vel, p.cdf_rigid_vel);
        p.cdf_signed_distance = d.cdf.signed_distance;
        p.cdf_affinity = d.cdf.affinity;
        p.init_volume = d.init_volume;
        p.init_radius = d.init_radius;
        p.mass = d.mass;
        p.lambda = s.constitutive_model[i].lambda;
        p.mu = s.constitutive_model[i].mu;
        const Plasticity& pl = s.plasticity[i];
        p.dp_h0 = pl.ha;
        p.dp_h1 = pl.hb;
        p.dp_h2 = pl.hc;
        p.dp_h3 = pl.hd;
        p.dp_lambda = pl.lambda;
        p.dp_mu = pl.mu;
        p.plastic_det = s.plastic_state[i].plastic_deformation_gradient_det;
        p.plastic_hardening = s.plastic_state[i].plastic_hardening;
        p.plastic_log_vol_gain = s.plastic_state[i].log_vol_gain;
        p.phase = s.phases[i].phase;
        p.max_stretch = s.phases[i].max_stretch;
        p.model = s.model_kind[i];
    }
}

template <int D>
size_t read_grid(const Sim<D>& s, b200mpm_block_info* blocks, b200mpm_node* nodes, size_t cap) {
    size_t nbk = std::min<size_t>(s.num_active_blocks, cap);
    for (size_t b = 0; b < nbk; ++b) {
        const ActiveBlockHeader<D>& ab = s.active_blocks[b];
        for (int k = 0; k < 3; ++k) blocks[b].vid[k] = (k < D) ? ab.virtual_id.id[k] : 0;
        blocks[b].first_particle = ab.first_particle;
        blocks[b].num_particles = ab.num_particles;
        for (size_t c = 0; c < 64; ++c) {
            const Node<D>& n = s.nodes[b * 64 + c];
            b200mpm_node& o = nodes[b * 64 + c];
            for (int k = 0; k < 4; ++k) o.momentum_velocity_mass[k] = 0.0f;
            for (int k = 0; k < D; ++k) o.momentum_velocity_mass[k] = n.momentum_velocity[k];
            o.momentum_velocity_mass[D] = n.mass;
            o.cdf_distance = n.cdf.distance;
            o.cdf_affinities = n.cdf.affinities;
            o.cdf_closest_id = n.cdf.closest_id;
        }
    }
    return nbk;
}

template <int D>
void stage(Sim<D>& s, int id) {
    switch (id) {
    case B200MPM_PASS_UPDATE_RIGID_PARTICLES:
        s.update_world_mass_properties();
        s.transform_rigid_points();
        break;
    case B200MPM_PASS_GRID_SORT: s.queue_sort(); break;
    case B200MPM_PASS_GRID_UPDATE_CDF: s.grid_update_cdf(); break;
    case B200MPM_PASS_P2G_CDF: s.p2g_cdf(); break;
    case B200MPM_PASS_G2P_CDF: s.g2p_cdf(); break;
    case B200MPM_PASS_P2G: s.p2g(); break;
    case B200MPM_PASS_GRID_UPDATE: s.grid_update(); break;
    case B200MPM_PASS_G2P: s.g2p(); break;
    case B200MPM_PASS_PARTICLES_UPDATE: s.particles_update(); break;
    case B200MPM_PASS_INTEGRATE_BODIES: s.integrate_bodies(); break;
    default: break;
    }
}

template <int D>
void read_poses(const Sim<D>& s, b200mpm_pose* out, size_t n) {
    for (size_t i = 0; i < std::min(n, s.num_bodies()); ++i) {
        vec_to<D>(s.poses[i].t, out[i].translation);
        for (int k = 0; k < 4; ++k) out[i].rotation[k] = s.pose_rot_raw[i * 4 + k];
    }
}
template <int D>
void read_vels(const Sim<D>& s, b200mpm_velocity* out, size_t n) {
    for (size_t i = 0; i < std::min(n, s.num_bodies()); ++i) {
        vec_to<D>(s.body_vels[i].linear, out[i].linear);
        out[i].angular[0] = out[i].angular[1] = out[i].angular[2] = 0.0f;
        if constexpr (D == 2) out[i].angular[0] = s.body_vels[i].angular;
        else vec_to<3>(s.body_vels[i].angular, out[i].angular);
    }
}
template <int D>
void write_poses(Sim<D>& s, const b200mpm_pose* in, size_t n) {
    for (size_t i = 0; i < std::min(n, s.num_bodies()); ++i) {
        s.poses[i].R = rotation_from(in[i].rotation, std::integral_constant<int, D>{});
        s.poses[i].t = vec_from<D>(in[i].translation);
        for (int k = 0; k < 4; ++k) s.pose_rot_raw[i * 4 + k] = in[i].rotation[k];
    }
}
template <int D>
void write_vels(Sim<D>& s, const b200mpm_velocity* in, size_t n) {
    for (size_t i = 0; i < std::min(n, s.num_bodies()); ++i) {
        s.body_vels[i].linear = vec_from<D>(in[i].linear);
        if constexpr (D == 2) s.body_vels[i].angular = in[i].angular[0];
        else s.body_vels[i].angular = vec_from<3>(in[i].angular);
    }
}

} // namespace

#define DISPATCH(h, expr2, expr3)          \
    do {                                   \
        if ((h)->dim == 2) {               \
            auto& s = *(h)->s2;            \
            (void)s;                       \
            expr2;                         \
        } else {                           \
            auto& s = *(h)->s3;            \
            (void)s;                       \
            expr3;                         \
        }                                  \
    } while (0)

extern "C" {

void* oracle_create(int dim, const b200mpm_sim_params* params, const b200mpm_particle* parts, size_t n,
                    const b200mpm_body* bodies, size_t nb, float cell_width, uint32_t grid_capacity) {
    if ((dim != 2 && dim != 3) || nb > B200MPM_MAX_BODIES) return nullptr;
    auto* h = new Handle();
    h->dim = dim;
    if (dim == 2) h->s2 = build<2>(params, parts, n, bodies, nb, cell_width, grid_capacity);
    else h->s3 = build<3>(params, parts, n, bodies, nb, cell_width, grid_capacity);
    return h;
}
void oracle_destroy(void* hv) {
    auto* h = (Handle*)hv;
    if (!h) return;
    delete h->s2;
    delete h->s3;
    delete h;
}
void oracle_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_substep(void* hv, uint32_t n) {
    auto* h = (Handle*)hv;
    for (uint32_t i = 0; i < n; ++i) DISPATCH(h, s.substep(), s.substep());
}
void oracle_stage(void* hv, int id) {
    auto* h = (Handle*)hv;
    DISPATCH(h, stage<2>(s, id), stage<3>(s, id));
}
void oracle_read_particles(void* hv, b200mpm_particle* out) {
    auto* h = (Handle*)hv;
    DISPATCH(h, read_particles<2>(s, out), read_particles<3>(s, out));
}
size_t oracle_read_grid(void* hv, b200mpm_block_info* blocks, b200mpm_node* nodes, size_t cap) {
    auto* h = (Handle*)hv;
    size_t r = 0;
    DISPATCH(h, r = read_grid<2>(s, blocks, nodes, cap), r = read_grid<3>(s, blocks, nodes, cap));
    return r;
}
uint32_t oracle_num_active_blocks(void* hv) {
    auto* h = (Handle*)hv;
    uint32_t r = 0;
    DISPATCH(h, r = s.num_active_blocks, r = s.num_active_blocks);
    return r;
}
int oracle_overflowed(void* hv) {
    auto* h = (Handle*)hv;
    int r = 0;
    DISPATCH(h, r = s.overflowed, r = s.overflowed);
    return r;
}
void oracle_read_sorted_ids(void* hv, uint32_t* out) {
    auto* h = (Handle*)hv;
    DISPATCH(h, std::memcpy(out, s.sorted_particle_ids.data(), s.num_particles() * 4),
             std::memcpy(out, s.sorted_particle_ids.data(), s.num_particles() * 4));
}
void oracle_read_body_poses(void* hv, b200mpm_pose* out, size_t n) {
    auto* h = (Handle*)hv;
    DISPATCH(h, read_poses<2>(s, out, n), read_poses<3>(s, out, n));
}
void oracle_read_body_vels(void* hv, b200mpm_velocity* out, size_t n) {
    auto* h = (Handle*)hv;
    DISPATCH(h, read_vels<2>(s, out, n), read_vels<3>(s, out, n));
}
void oracle_write_body_poses(void* hv, const b200mpm_pose* in, size_t n) {
    auto* h = (Handle*)hv;
    DISPATCH(h, write_poses<2>(s, in, n), write_poses<3>(s, in, n));
}
void oracle_write_body_vels(void* hv, const b200mpm_velocity* in, size_t n) {
    auto* h = (Handle*)hv;
    DISPATCH(h, write_vels<2>(s, in, n), write_vels<3>(s, in, n));
}
void oracle_write_sim_params(void* hv, const b200mpm_sim_params* p) {
    auto* h = (Handle*)hv;
    DISPATCH(h, (s.gravity = vec_from<2>(p->gravity), s.dt = p->dt), (s.gravity = vec_from<3>(p->gravity), s.dt = p->dt));
}
// Integer body impulses accumulated by the last p2g (linear xyz, angular xyz per body).
void oracle_read_impulses(void* hv, int32_t* out) {
    auto* h = (Handle*)hv;
    DISPATCH(h,
             for (size_t b = 0; b < s.num_bodies(); ++b) {
                 for (int k = 0; k < 3; ++k) out[b * 6 + k] = (k < 2) ? s.body_impulses[b].linear[k] : 0;
                 for (int k = 0; k < 3; ++k) out[b * 6 + 3 + k] = s.body_impulses[b].angular[k];
             },
             for (size_t b = 0; b < s.num_bodies(); ++b) {
                 for (int k = 0; k < 3; ++k) out[b * 6 + k] = s.body_impulses[b].linear[k];
                 for (int k = 0; k < 3; ++k) out[b * 6 + 3 + k] = s.body_impulses[b].angular[k];
             });
}

// WgPrefixSum::eval_cpu (src/grid/prefix_sum.rs:71-83).
void oracle_prefix_sum(uint32_t* v, size_t len) {
    std::vector<uint32_t> tmp(v, v + len);
    Sim<3>::prefix_sum(tmp);
    std::memcpy(v, tmp.data(), len * 4);
}

// Unit-level hooks (column-major DxD matrices).
void oracle_svd(int dim, const float* F, float* U, float* S, float* Vt) {
    if (dim == 2) {
        Svd<2> r = svd(mat_from<2>(F));
        mat_to<2>(r.U, U);
        mat_to<2>(r.Vt, Vt);
        S[0] = r.S[0];
        S[1] = r.S[1];
        S[2] = 0;
    } else {
        Svd<3> r = svd(mat_from<3>(F));
        mat_to<3>(r.U, U);
        mat_to<3>(r.Vt, Vt);
        for (int i = 0; i < 3; ++i) S[i] = r.S[i];
    }
}
void oracle_kirchoff_stress(int dim, int model, float lambda, float mu, const float* F, float* out) {
    ElasticCoefficients m{lambda, mu};
    if (dim == 2) {
        Mat<2> r = model == B200MPM_MODEL_NEO_HOOKEAN ? Sim<2>::kirchoff_stress_neo_hookean(m, mat_from<2>(F))
                                                       : Sim<2>::kirchoff_stress_corotated(m, mat_from<2>(F));
        mat_to<2>(r, out);
    } else {
        Mat<3> r = model == B200MPM_MODEL_NEO_HOOKEAN ? Sim<3>::kirchoff_stress_neo_hookean(m, mat_from<3>(F))
                                                       : Sim<3>::kirchoff_stress_corotated(m, mat_from<3>(F));
        mat_to<3>(r, out);
    }
}
// plasticity: 6 floats (h0..h3, lambda, mu); state: 3 floats in/out; F in/out.
void oracle_dp_project(int dim, const float* plasticity, float* state, float* F) {
    Plasticity p{plasticity[0], plasticity[1], plasticity[2], plasticity[3], plasticity[4], plasticity[5]};
    PlasticState st{state[0], state[1], state[2]};
    if (dim == 2) {
        Mat<2> f = mat_from<2>(F);
        Sim<2>::dp_project(p, st, f);
        mat_to<2>(f, F);
    } else {
        Mat<3> f = mat_from<3>(F);
        Sim<3>::dp_project(p, st, f);
        mat_to<3>(f, F);
    }
    state[0] = st.plastic_deformation_gradient_det;
    state[1] = st.plastic_hardening;
    state[2] = st.log_vol_gain;
}
// Shape::projectPointOnBoundary for one body description.
int oracle_project_point(int dim, const b200mpm_body* b, const float* pt, float* out) {
    if (dim == 2) {
        Shape<2> sh{b->shape_type, vec_from<2>(b->shape_a), vec_from<2>(b->shape_b), b->radius};
        Pose<2> p{rotation_from(b->rotation, std::integral_constant<int, 2>{}), vec_from<2>(b->translation)};
        auto r = project_point_on_boundary(sh, p, vec_from<2>(pt));
        vec_to<2>(r.point, out);
        return r.is_inside;
    }
    Shape<3> sh{b->shape_type, vec_from<3>(b->shape_a), vec_from<3>(b->shape_b), b->radius};
    Pose<3> p{rotation_from(b->rotation, std::integral_constant<int, 3>{}), vec_from<3>(b->translation)};
    auto r = project_point_on_boundary(sh, p, vec_from<3>(pt));
    vec_to<3>(r.point, out);
    return r.is_inside;
}
void oracle_set_rigid_particles(void* hv, const float* vertices, const uint32_t* vertex_colliders, size_t nv,
                                const float* samples, const uint32_t* ids4, size_t ns) {
    auto* h = (Handle*)hv;
    DISPATCH(h, s.set_rigid_particles(vertices, vertex_colliders, nv, samples, ids4, ns),
             s.set_rigid_particles(vertices, vertex_colliders, nv, samples, ids4, ns));
}

void oracle_prep_vertex_buffer(void* hv, uint32_t mode, float* inst) {
    auto* h = (Handle*)hv;
    DISPATCH(h, s.prep_vertex_buffer(mode, inst), s.prep_vertex_buffer(mode, inst));
}

// Test diagnostic: see exact_sigma_mode() in mpm_oracle.hpp. Process-wide.
void oracle_set_exact_sigma(int on) { exact_sigma_mode() = (on != 0); }

} // extern "C"
