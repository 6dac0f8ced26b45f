"""CPU tests of the host-side mirror of the reference's data structs (wgsparkl_b200/{models,solver,rapier}.py)."""
import numpy as np

from wgsparkl_b200 import abi, scenes
from wgsparkl_b200.models import DruckerPrager, ElasticCoefficients
from wgsparkl_b200.rapier import ColliderBuilder, ColliderSet, RigidBodyBuilder, RigidBodySet, bodies_to_abi
from wgsparkl_b200.solver import Particle, ParticleDynamics, ParticlePhase, make_particles, particles_to_abi


def test_lame_parameters_in_f32():
    e = ElasticCoefficients.from_young_modulus(100_000.0, 0.33)  # models/mod.rs:52-75
    assert np.isclose(e.lambda_, 72976.56, rtol=1e-6) and np.isclose(e.mu, 37593.984, rtol=1e-6)
    dp = DruckerPrager.new(-1.0, -1.0)  # disabled plasticity quirk: lambda = mu = -1 (drucker_prager.rs:19-23)
    assert dp.lambda_ == -1.0 and dp.mu == -1.0
    assert np.isclose(dp.h0, np.deg2rad(35.0)) and np.isclose(dp.h3, np.deg2rad(10.0))


def test_with_density():
    d3 = ParticleDynamics.with_density(0.25, 2700.0, 3)  # particle3d.rs:28-42
    assert np.isclose(d3.init_volume, 0.125) and np.isclose(d3.mass, 337.5)
    d2 = ParticleDynamics.with_density(0.05, 1000.0, 2)
    assert np.isclose(d2.init_volume, 0.01) and np.isclose(d2.mass, 10.0)


def test_option_defaults_match_gpu_models_from_particles():
    """models/mod.rs:20-36: plasticity None -> DruckerPrager::new(-1,-1); phase None -> {0, -1}."""
    p = Particle([0.0, 1.0, 2.0], ParticleDynamics.with_density(0.25, 1.0), ElasticCoefficients.from_young_modulus(1e5, 0.33))
    flat = particles_to_abi([p], 3)[0]
    assert flat["dp_lambda"] == -1.0 and flat["phase"] == 0.0 and flat["max_stretch"] == -1.0
    assert flat["plastic_det"] == 1.0 and flat["plastic_hardening"] == 1.0 and flat["plastic_log_vol_gain"] == 0.0
    assert np.array_equal(flat["def_grad"], np.eye(3, dtype=np.float32).reshape(-1))
    q = Particle([0.0, 1.0], ParticleDynamics.with_density(0.25, 1.0, 2), ElasticCoefficients.from_young_modulus(1e5, 0.33),
                 phase=ParticlePhase(1.0, 3.0e38))
    flat2 = particles_to_abi([q], 2)[0]
    assert np.array_equal(flat2["def_grad"][:4], [1, 0, 0, 1]) and flat2["phase"] == 1.0
    vec = make_particles(np.float32([[0.0, 1.0, 2.0]]), 3, 0.25, 1.0, ElasticCoefficients.from_young_modulus(1e5, 0.33))
    for f in abi.particle_dtype.names:
        assert np.array_equal(vec[0][f], flat[f]), f


def test_coupling_order_and_mass_properties():
    bodies, colliders = RigidBodySet(), ColliderSet()
    fixed = bodies.insert(RigidBodyBuilder.fixed().translation([0.0, -4.0, 0.0]))
    colliders.insert_with_parent(ColliderBuilder.cuboid(100.0, 4.0, 100.0), fixed, bodies)
    colliders.insert(ColliderBuilder.ball(1.0))  # no parent: not coupled (pipeline.rs:107-117)
    dyn = bodies.insert(RigidBodyBuilder.dynamic().translation([0.0, 5.0, 0.0]))
    colliders.insert_with_parent(ColliderBuilder.cuboid(1.0, 2.0, 3.0).density(10.0), dyn, bodies)
    out = bodies_to_abi(bodies, colliders, 3)
    assert len(out) == 2 and out[0]["shape_type"] == abi.SHAPE_CUBOID
    assert np.all(out[0]["inv_mass"] == 0.0)
    m = 10.0 * 8 * 1 * 2 * 3
    assert np.allclose(out[1]["inv_mass"], 1.0 / m)
    ixx = m * (2.0**2 + 3.0**2) / 3.0
    assert np.isclose(out[1]["inv_inertia"][0], 1.0 / ixx)
    assert np.allclose(out[1]["rotation"], [0, 0, 0, 1])


def test_scene_sizes():
    s = scenes.elastic_cube_3d(10)
    assert len(s["particles"]) == 1000 and len(s["bodies"]) == 1 and s["dim"] == 3
    s = scenes.reference_test_lattice()
    assert len(s["particles"]) == 1000 and s["grid_capacity"] == 100_000
    assert np.all((s["particles"]["position"] * 2) % 1 == 0)  # every coordinate on a round() tie
    s = scenes.mixed_coupled_3d(6, 6, 6)
    assert set(np.unique(s["particles"]["model"])) == {abi.MODEL_COROTATED, abi.MODEL_NEO_HOOKEAN}


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--n-side", "12"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "particle-substeps/sec" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["unit"] == "particle-substeps/s" and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]
    # ranks other than 0 of a multi-process launch print nothing and exit 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_bench_own_arm_fails_loudly_without_a_gpu():
    """No CPU fallback anywhere: on a machine without an sm_100 device `bench.py` must fail, not measure something else."""
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--n-side", "8", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode != 0
    assert res.stdout.strip() == "", "no JSON line may be printed"
    assert "device" in res.stderr.lower() or "cuda" in res.stderr.lower()
