// Compile / behaviour test of the C++ host mirror (include/wgsparkl_b200.hpp). Mirrors the reference's
// pipeline_queue_step test (src/pipeline.rs:296-343): 10^3 lattice, E = 1e5, nu = 0.33, three substeps.
#include <cstdio>

#include "wgsparkl_b200.hpp"

using namespace wgsparkl;

int main() {
    const float cell_width = 1.0f;
    std::vector<solver::Particle<3>> particles;
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 10; ++j)
            for (int k = 0; k < 10; ++k) {
                solver::Particle<3> p;
                p.position[0] = i / cell_width / 2.0f;
                p.position[1] = j / cell_width / 2.0f;
                p.position[2] = k / cell_width / 2.0f;
                p.dynamics = solver::ParticleDynamics<3>::with_density(cell_width / 4.0f, 1.0f);
                p.model = models::ElasticCoefficients::from_young_modulus(100000.0f, 0.33f);
                particles.push_back(p);
            }
    if (std::fabs(particles[0].dynamics.mass - 0.125f) > 1e-7f) return 2;
    if (std::fabs(models::DruckerPrager::new_(2.0e9f, 0.2f).lambda - 555555520.0f) > 64.0f) return 3;
    {
        // sample_trimesh (particle3d.rs:214-292) on two triangles sharing an edge; compared with the Python
        // restatement by tests/test_abi.py
        solver::RigidParticles rp;
        solver::sample_trimesh(rp, 3, {0, 0, 0, 10, 0, 0, 0, 8, 0, 10, 8, 3}, {0, 1, 2, 1, 3, 2}, 1.0f);
        double cx = 0, cy = 0, cz = 0;
        const size_t n = rp.sample_ids.size() / 4;
        for (size_t i = 0; i < n; ++i) cx += rp.samples[3 * i], cy += rp.samples[3 * i + 1], cz += rp.samples[3 * i + 2];
        std::printf("samples %zu centroid %.5f %.5f %.5f vertices %zu collider %u\n", n, cx / n, cy / n, cz / n,
                    rp.vertex_colliders.size(), rp.sample_ids[3]);
    }
    solver::SimulationParamsT<3> params{{0.0f, -9.81f, 0.0f}, (1.0f / 60.0f) / 10.0f};
    try {
        auto pipeline = pipeline::MpmPipeline<3>::new_(0);
        auto data = pipeline::MpmData<3>::with_select_coupling(pipeline, params, particles, {}, cell_width, 100000);
        for (int it = 0; it < 3; ++it) pipeline.queue_step(data, 1);
        pipeline.sync();
        auto out = data.read_particles();
        const float vy = out[500].velocity[1];
        if (!(vy < -0.048f && vy > -0.050f)) return 4;
        std::printf("stepped: vy = %g\n", vy);
    } catch (const Error& e) {
        if (e.code != B200MPM_ERR_NO_DEVICE) {
            std::printf("unexpected error %d: %s\n", e.code, e.what());
            return 5;
        }
        std::printf("no device: %s\n", e.what());
    }
    return 0;
}
