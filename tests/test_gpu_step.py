"""GPU parity of whole substeps (MpmPipeline::queue_step, src/pipeline.rs:195-281) against the oracle.

Primary gate: ONE substep from an identical, developed state, fields within
1e-5 relative (norm-wise, relative to the largest magnitude of the field). Secondary: positions within
1e-4 after 100 substeps on an elastic scene. Tolerances are f32 summation-order / SVD-conditioning
bounds, written next to each assertion. All calls go through the C ABI.
"""
import numpy as np
import pytest

import parity
from wgsparkl_b200 import abi, scenes
from wgsparkl_b200.pipeline import MpmData

pytestmark = pytest.mark.gpu

TOL = 1e-5  # BASELINE.md §5 primary gate


def developed_state(oracle_mod, scene, substeps):
    """Advance the scene with the oracle so that F, affine, velocities and cdf are non-trivial, and
    return the particle array to be used as the common initial state."""
    sim = oracle_mod.OracleSim(scene["dim"], scene["params"], scene["particles"], scene["bodies"], scene["cell_width"],
                               scene["grid_capacity"])
    sim.step(substeps)
    parts = sim.read_particles()
    poses, vels = sim.read_body_poses(), sim.read_body_vels()
    sim.close()
    return parts, poses, vels


def one_substep_both(oracle_mod, pipe, scene, particles, n=1):
    data = MpmData(pipe, scene["params"], particles, scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    sim = oracle_mod.OracleSim(scene["dim"], scene["params"], particles, scene["bodies"], scene["cell_width"],
                               scene["grid_capacity"])
    pipe.queue_step(data, n)
    pipe.sync()
    sim.step(n)
    return data, sim


def test_reference_pipeline_queue_step_scene(pipe3, oracle_mod):
    """The reference's pipeline_queue_step smoke test (pipeline.rs:296-343): 10^3 lattice on round() ties,
    plasticity None / phase None => Drucker-Prager with lambda = mu = -1. 3 substeps."""
    scene = scenes.reference_test_lattice()
    data, sim = one_substep_both(oracle_mod, pipe3, scene, scene["particles"], n=3)
    g, o = data.read_particles(), sim.read_particles()
    # With lambda = mu = -1 the return mapping runs on F = I + O(eps): every singular value it produces
    # is 1 +- a few ulps, and the corotated stress turns each ulp into 2 mu eps of force. The velocity
    # tolerance is that conditioning bound (4 ulps * 2 mu * V0 * inv_d * dt / mass * h per substep ~ 1e-4
    # absolute on |v| = 0.05), not a property of either implementation.
    parity.assert_particles_close(g, o, TOL, fields=("position", "def_grad"), tols={"position": 2e-6})
    assert parity.field_rel_err(g["velocity"], o["velocity"]) <= 5e-3
    parity.assert_affine_close(g, o, 3, scene["cell_width"], float(scene["params"].dt))
    assert data.status()[0] == sim.num_active_blocks()
    for f in ("plastic_det", "plastic_hardening"):
        assert parity.field_rel_err(g[f], o[f]) <= 1e-5
    data.close()


@pytest.mark.parametrize("ground", [False, True])
def test_one_substep_elastic_cube(pipe3, oracle_mod, ground):
    """Config-2 physics at test size: corotated elasticity, cube resting on / bouncing off the ground cuboid."""
    scene = scenes.elastic_cube_3d(16, y_offset=-6.0 if ground else 3.0, ground=ground)
    scene["particles"]["velocity"][:, 1] = -4.0
    parts, _, _ = developed_state(oracle_mod, scene, 60)
    if ground:
        # Push part of the contact layer through the ground surface (y = -3): their colour is latched
        # "outside" (g2p_cdf.wgsl:179-188), so the next substep sees negative signed distances and runs
        # the CPIC projection + penalty branches (particle_update.wgsl:64-83).
        low = (parts["cdf_affinity"] != 0) & (parts["position"][:, 1] < -2.8) & (parts["position"][:, 0] > 0.0)
        assert low.sum() > 20
        parts["position"][low, 1] -= 0.25
    data, sim = one_substep_both(oracle_mod, pipe3, scene, parts)
    g, o = data.read_particles(), sim.read_particles()
    gb, gn = data.read_grid()
    ob, on = sim.read_grid()
    parity.assert_grid_close(gb, gn, ob, on, TOL)
    # position: 2e-6 of the domain extent = a few f32 ulps; affine: f32 conditioning bound of the
    # stress term (parity.affine_abs_bound, DESIGN.md "Tolerances")
    parity.assert_particles_close(g, o, TOL, fields=("position", "velocity", "def_grad"), tols={"position": 2e-6})
    parity.assert_affine_close(g, o, 3, scene["cell_width"], float(scene["params"].dt))
    if ground:
        assert np.array_equal(g["cdf_affinity"], o["cdf_affinity"])
        assert (o["cdf_affinity"] != 0).sum() > 100, "test scene must exercise CPIC"
        assert (o["cdf_signed_distance"] < -0.05).sum() > 20, "test scene must exercise penetration"
        for f in ("cdf_normal", "cdf_signed_distance", "cdf_rigid_vel"):
            assert parity.field_rel_err(g[f], o[f]) <= 1e-4, f
    data.close()


def test_one_substep_sand(pipe3, oracle_mod):
    """Config-3 physics at test size: Drucker-Prager sand on the ground cuboid, developed for 60 substeps."""
    scene = scenes.sand_column_3d(12, 24, 12, y_offset=-5.0)
    parts, _, _ = developed_state(oracle_mod, scene, 60)
    data, sim = one_substep_both(oracle_mod, pipe3, scene, parts)
    g, o = data.read_particles(), sim.read_particles()
    errs = parity.assert_particles_close(g, o, TOL, fields=("position", "velocity", "def_grad"), tols={"position": 2e-6})
    parity.assert_affine_close(g, o, 3, scene["cell_width"], float(scene["params"].dt))
    # plastic state: log_vol_gain accumulates log(det) differences of singular values 1 +- 1e-3 in f32:
    # absolute 1e-7 on values of order 1e-4
    assert parity.field_rel_err(g["plastic_det"], o["plastic_det"]) <= 1e-5, errs
    assert parity.field_rel_err(g["plastic_hardening"], o["plastic_hardening"]) <= 1e-5, errs
    assert parity.field_rel_err(g["plastic_log_vol_gain"], o["plastic_log_vol_gain"]) <= 5e-3
    assert np.array_equal(g["cdf_affinity"], o["cdf_affinity"])
    data.close()


def test_hundred_substeps_elastic(pipe3, oracle_mod):
    """Secondary gate: positions within 1e-4 (relative to the domain extent) after 100 substeps."""
    scene = scenes.elastic_cube_3d(12, y_offset=-5.0)
    data, sim = one_substep_both(oracle_mod, pipe3, scene, scene["particles"], n=100)
    g, o = data.read_particles(), sim.read_particles()
    assert parity.field_rel_err(g["position"], o["position"]) <= 1e-4
    assert parity.field_rel_err(g["velocity"], o["velocity"]) <= 1e-4
    data.close()


def test_hundred_substeps_sand_reported(pipe3, oracle_mod):
    """Sand is chaotic (hard branches in drucker_prager.wgsl:118,125 amplify summation-order noise,
    SURVEY §7): positions still have to stay within the 1e-4 bound, velocities are only sanity-checked."""
    scene = scenes.sand_column_3d(12, 24, 12, y_offset=-5.0)
    data, sim = one_substep_both(oracle_mod, pipe3, scene, scene["particles"], n=100)
    g, o = data.read_particles(), sim.read_particles()
    assert parity.field_rel_err(g["position"], o["position"]) <= 1e-4
    assert parity.field_rel_err(g["velocity"], o["velocity"]) <= 5e-2
    data.close()


def test_one_substep_2d(pipe2, oracle_mod):
    """Config 1 (2D elastic block on a cuboid) at test size."""
    scene = scenes.elastic_block_2d(40)
    scene["particles"]["position"][:, 1] -= 9.9  # start in contact with the ground
    parts, _, _ = developed_state(oracle_mod, scene, 40)
    data, sim = one_substep_both(oracle_mod, pipe2, scene, parts)
    g, o = data.read_particles(), sim.read_particles()
    gb, gn = data.read_grid()
    ob, on = sim.read_grid()
    parity.assert_grid_close(gb, gn, ob, on, TOL)
    parity.assert_particles_close(g, o, TOL, fields=("position", "velocity", "def_grad"), tols={"position": 2e-6})
    parity.assert_affine_close(g, o, 2, scene["cell_width"], float(scene["params"].dt))
    assert np.array_equal(g["cdf_affinity"], o["cdf_affinity"])
    data.close()


def test_two_way_coupling_bodies(pipe3, oracle_mod):
    """Config-4 physics at test size: sand + Neo-Hookean solids, kinematic rotating cuboid and dynamic
    cuboids; body poses / velocities integrated on the device (rigid_impulses.wgsl:94-137)."""
    scene = scenes.mixed_coupled_3d(12, 12, 12, n_dynamic=2)
    # drop the dynamic bodies right onto the material so that impulses flow within the test window
    scene["bodies"]["translation"][2:, 1] = 12.0
    parts, poses, vels = developed_state(oracle_mod, scene, 30)
    scene2 = dict(scene)
    bodies = scene["bodies"].copy()
    bodies["translation"] = poses["translation"]
    bodies["rotation"] = poses["rotation"]
    bodies["linvel"] = vels["linear"]
    bodies["angvel"] = vels["angular"]
    scene2["bodies"] = bodies
    data, sim = one_substep_both(oracle_mod, pipe3, scene2, parts, n=1)
    g, o = data.read_particles(), sim.read_particles()
    parity.assert_particles_close(g, o, TOL, fields=("position", "velocity", "def_grad"), tols={"position": 2e-6})
    parity.assert_affine_close(g, o, 3, scene["cell_width"], float(scene["params"].dt))
    assert np.array_equal(g["cdf_affinity"], o["cdf_affinity"])
    gp, op = data.read_body_poses(), sim.read_body_poses()
    gv, ov = data.read_body_vels(), sim.read_body_vels()
    assert parity.field_rel_err(gp["translation"], op["translation"]) <= 1e-6
    assert parity.field_rel_err(gp["rotation"], op["rotation"]) <= 1e-6
    # impulses are accumulated as i32(x * 1e5) per node (p2g.wgsl:142-155, rigid_impulses.wgsl:50-58) on both sides;
    # what is left is the summation order of the node's float total in front of the truncation
    el, ea = parity.field_rel_err(gv["linear"], ov["linear"]), parity.field_rel_err(gv["angular"], ov["angular"])
    assert el <= 2e-6 and ea <= 2e-6, (el, ea)
    data.close()


def test_host_writes_take_effect(pipe3):
    scene = scenes.elastic_cube_3d(8, y_offset=3.0)
    data = MpmData(pipe3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    poses = data.read_body_poses()
    poses["translation"][0] = [1.0, -5.0, 2.0]
    data.write_body_poses(poses)
    vels = data.read_body_vels()
    vels["linear"][0] = [0.5, 0.0, 0.0]
    data.write_body_vels(vels)
    assert np.allclose(data.read_body_poses()["translation"][0], [1.0, -5.0, 2.0])
    assert np.allclose(data.read_body_vels()["linear"][0], [0.5, 0.0, 0.0])
    from wgsparkl_b200 import SimulationParams

    data.write_sim_params(SimulationParams([0.0, 0.0, 0.0], 1e-3))
    pipe3.queue_step(data, 5)
    pipe3.sync()
    out = data.read_particles()
    assert np.abs(out["velocity"]).max() < 1e-3  # gravity switched off
    data.close()


def test_timings_use_reference_pass_names(pipe3):
    scene = scenes.elastic_cube_3d(8, y_offset=3.0)
    data = MpmData(pipe3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    pipe3.set_timestamps(True)
    pipe3.queue_step(data, 4)
    t = pipe3.timings_ms()
    pipe3.set_timestamps(False)
    assert tuple(t.keys()) == abi.PASS_NAMES
    assert t["p2g"] > 0.0 and t["g2p"] > 0.0 and t["grid sort"] > 0.0
    data.close()


def test_kernel_timings_and_in_graph_timeline(pipe3):
    """Per-kernel event timers (booked on the reference's pass names as well) and the %globaltimer timeline of the
    graph replay: the substep is a chain touch -> block_prepare -> scatter -> p2g -> g2p."""
    scene = scenes.elastic_cube_3d(16, y_offset=-1.0)
    data = MpmData(pipe3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    pipe3.timings_ms(), pipe3.kernel_timings_ms()  # (both accumulate since their last call: reset)
    pipe3.set_timestamps(True)
    pipe3.queue_step(data, 4)
    passes, kernels = pipe3.timings_ms(), pipe3.kernel_timings_ms()
    pipe3.set_timestamps(False)
    assert tuple(kernels.keys()) == abi.KERNEL_NAMES
    for k in ("touch", "block_prepare", "scatter", "p2g", "g2p"):
        assert kernels[k] > 0.0, k
    assert abs(passes["p2g"] - kernels["p2g"]) <= 1e-6 and abs(passes["g2p"] - kernels["g2p"]) <= 1e-6
    assert abs(sum(passes.values()) - sum(kernels.values())) <= 1e-4  # every kernel timer is booked on one pass
    first = data.debug_timeline()  # switches the recording on
    assert all(v is None for v in first.values())
    ref = data.read_positions()
    pipe3.queue_step(data, 1)
    pipe3.sync()
    tl = data.debug_timeline()
    chain = [tl[k] for k in ("touch", "block_prepare", "scatter", "p2g", "g2p")]
    assert all(c is not None and c[1] >= c[0] for c in chain)
    for a, b in zip(chain, chain[1:]):
        assert b[0] >= a[1] - 2000, (a, b)  # (the stamps are taken by thread 0 of a CTA: allow 2 us of skew)
    assert chain[-1][1] - chain[0][0] < 5_000_000
    data.debug_timeline(enable=False)
    pipe3.queue_step(data, 1)
    pipe3.sync()
    assert np.isfinite(data.read_positions()).all() and not np.array_equal(ref, data.read_positions())
    data.close()


def test_async_position_readback_matches_blocking(pipe3):
    """b200mpm_read_positions_async: snapshots taken between steps land (after sync) with exactly the values
    the blocking read returns at the same points (up to atomic-order noise), also with three readbacks enqueued back to back."""
    import torch

    scene = scenes.elastic_cube_3d(12, y_offset=-5.0)
    n = len(scene["particles"])

    def fresh():
        return MpmData(pipe3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])

    data = fresh()
    expected = []
    for _ in range(3):
        pipe3.queue_step(data, 5)
        expected.append(data.read_positions().copy())
    data.close()

    data = fresh()
    outs = [torch.empty((n, 4), dtype=torch.float32).pin_memory().numpy() for _ in range(3)]
    for k in range(3):
        pipe3.queue_step(data, 5)
        data.read_positions_async(outs[k])
    pipe3.sync()
    # (two runs are not bit-identical: the node reductions are floating-point atomics)
    for k in range(3):
        assert np.allclose(outs[k], expected[k], rtol=0.0, atol=2e-5)
    assert np.abs(outs[1] - outs[0]).max() > 2e-4 and np.abs(outs[2] - outs[1]).max() > 2e-4, "snapshots must differ"
    data.close()


@pytest.mark.parametrize("plastic", [False, True])
def test_dense_cells_multi_part_blocks(pipe3, oracle_mod, plastic):
    """32 particles per cell (4x the seeding density, what a strongly compressed material reaches): blocks hold
    up to 2048 particles, so G2P work items are split into parts (g2p_list), P2G stages several windows per
    half-block, and for sand the large-strain SVD path runs. One substep from a moving state against the oracle."""
    scene = scenes.sand_column_3d(8, 8, 8, y_offset=-5.0) if plastic else scenes.elastic_cube_3d(8, y_offset=-5.0)
    n_side = (16, 32, 8)  # two blocks along x, each completely filled: cells [0, 4) <=> coordinates [0.5, 4.5)
    i, j, k = np.meshgrid(np.arange(n_side[0]), np.arange(n_side[1]), np.arange(n_side[2]), indexing="ij")
    rng = np.random.default_rng(11)
    pos = np.stack([(i.ravel() + 0.5) * 0.5 + 0.5, (j.ravel() + 0.5) * 0.125 + 0.5, (k.ravel() + 0.5) * 0.5 + 0.5], axis=1)
    pos += rng.uniform(-0.02, 0.02, size=pos.shape)
    parts = np.repeat(scene["particles"][:1], len(pos)).copy()
    parts["position"][:, :3] = pos.astype(np.float32)
    parts["velocity"][:, :3] = rng.normal(0.0, 0.5, size=(len(pos), 3)).astype(np.float32)
    parts["velocity"][:, 1] -= 3.0
    F = np.tile(np.eye(3, dtype=np.float32).reshape(1, 9), (len(pos), 1))
    F += rng.normal(0.0, 0.02 if plastic else 0.15, size=F.shape).astype(np.float32)  # elastic: beyond the polynomial path
    parts["def_grad"][:, :9] = F
    data, sim = one_substep_both(oracle_mod, pipe3, scene, parts)
    gb, gn = data.read_grid()
    ob, on = sim.read_grid()
    assert gb["num_particles"].max() > 1024, "test scene must produce multi-part blocks"
    parity.assert_grid_close(gb, gn, ob, on, TOL)
    g, o = data.read_particles(), sim.read_particles()
    parity.assert_particles_close(g, o, TOL, fields=("position", "velocity", "def_grad"), tols={"position": 2e-6})
    parity.assert_affine_close(g, o, 3, scene["cell_width"], float(scene["params"].dt))
    data.close()


def test_grid_capacity_growth(pipe3):
    """b200mpm_data_reserve_grid / set_auto_grow (the reference's stubbed resize, grid.rs:43-118): a run that
    starts with a block capacity too small for where the particles are going must, with auto-growth on, end
    like a run that had plenty of capacity from the start - and without it, report the overflow."""
    scene = scenes.sand_column_3d(12, 12, 12, y_offset=10.0)  # cohesionless: an outward velocity field disperses it
    pos = scene["particles"]["position"][:, :3]
    radial = pos - pos.mean(axis=0)
    scene["particles"]["velocity"][:, :3] = (40.0 * radial / np.abs(radial).max()).astype(np.float32)

    def run(capacity, auto):
        data = MpmData(pipe3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], capacity)
        if auto:
            data.set_auto_grow(0.5)
        for _ in range(48):
            pipe3.queue_step(data, 5)
        out = data.read_particles(), data.status()
        data.close()
        return out

    (ref, (nb_ref, over_ref)) = run(8192, False)
    assert not over_ref and nb_ref > 256, "the cloud must spread over more blocks than the small capacity"
    (_, (_, over_small)) = run(64, False)
    assert over_small, "without growth the small capacity must overflow (and say so)"
    (got, (nb, over)) = run(64, True)
    assert not over and nb == nb_ref
    assert parity.field_rel_err(got["position"], ref["position"]) <= 1e-5  # two runs: atomic-order noise only
    assert parity.field_rel_err(got["velocity"], ref["velocity"]) <= 5e-3


def test_reserve_grid_between_steps(pipe3):
    scene = scenes.elastic_cube_3d(12, y_offset=-5.0)
    a = MpmData(pipe3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], 1024)
    b = MpmData(pipe3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], 1024)
    pipe3.queue_step(a, 10)
    pipe3.queue_step(b, 10)
    b.reserve_grid(5000)  # -> 8192, graphs re-captured
    pipe3.queue_step(a, 10)
    pipe3.queue_step(b, 10)
    pa, pb = a.read_particles(), b.read_particles()
    assert parity.field_rel_err(pb["position"], pa["position"]) <= 2e-6
    assert parity.field_rel_err(pb["velocity"], pa["velocity"]) <= 1e-4
    assert a.status()[0] == b.status()[0]
    a.close()
    b.close()


@pytest.mark.parametrize("dim", [2, 3])
def test_prep_vertex_buffer_all_render_modes(pipe2, pipe3, oracle_mod, dim):
    """Render hand-off (src_testbed/prep_vertex_buffer{2d,3d}.wgsl): the instance buffer written on the device
    from the device state, against the oracle's restatement, for every RenderMode."""
    import torch

    pipe = pipe3 if dim == 3 else pipe2
    if dim == 3:
        scene = scenes.elastic_cube_3d(10, y_offset=-5.0)
    else:
        scene = scenes.elastic_block_2d(24)
        scene["particles"]["position"][:, 1] -= 9.9
    scene["particles"]["velocity"][:, 1] = -3.0
    parts, _, _ = developed_state(oracle_mod, scene, 40)
    data, sim = one_substep_both(oracle_mod, pipe, scene, parts, n=3)
    n = len(parts)
    rng = np.random.default_rng(2)
    init = np.zeros(n, dtype=abi.instance_dtype)
    init["base_color"] = rng.uniform(0.0, 1.0, size=(n, 4)).astype(np.float32)
    init["deformation"][:, :, 3] = 7.0  # padding lanes: must survive
    init["position"][:, 3] = 9.0
    o = sim.read_particles()
    assert (o["cdf_affinity"] != 0).sum() > 20, "test scene must have collider-side particles"
    for mode in range(6):
        dev = torch.from_numpy(init.view(np.float32).reshape(n, 24).copy()).cuda()
        pipe.prep_vertex_buffer(data, dev.data_ptr(), mode)
        pipe.sync()
        got = dev.cpu().numpy().reshape(-1).view(abi.instance_dtype)
        ref = sim.prep_vertex_buffer(init.copy(), mode)
        assert np.array_equal(got["base_color"], init["base_color"])
        assert np.all(got["deformation"][:, :, 3] == 7.0) and np.all(got["position"][:, 3] == 9.0)
        assert parity.field_rel_err(got["position"][:, :3], ref["position"][:, :3]) <= 2e-6
        assert parity.field_rel_err(got["deformation"][:, :, :3], ref["deformation"][:, :, :3]) <= 1e-5
        # VOLUME divides 1 - sigma by 0.005: an ulp of sigma is 2e-5 of colour; CDF modes follow normals /
        # distances that agree to 1e-4 (test_one_substep_elastic_cube); the rest is exact arithmetic on equal inputs
        tol = {abi.RENDER_VOLUME: 2e-3, abi.RENDER_VELOCITY: 1e-4, abi.RENDER_CDF_NORMALS: 1e-4, abi.RENDER_CDF_DISTANCES: 1e-4}.get(mode, 0.0)
        assert np.abs(got["color"] - ref["color"]).max() <= tol * max(1.0, np.abs(ref["color"]).max()), mode
    data.close()


def test_sixteen_colliders_of_every_shape(pipe3, oracle_mod):
    """The cap of 16 coupled colliders (rigid_impulses.rs:42, collide.wgsl:36): fixed balls and rotating kinematic
    capsules embedded in a block of sand, dynamic cuboids hovering just above it - affinity bits 0..15 and sign
    bits are all in use. (Dynamic bodies embedded in E = 2e9 sand receive impulses beyond the i32(x 1e5) range of
    rigid_impulses.wgsl:52-58, where the result is an overflow artefact in the reference too; impulses on dynamic
    bodies are covered by test_two_way_coupling_bodies.) Three substeps from a developed state against the oracle."""
    from wgsparkl_b200.rapier import ColliderBuilder, ColliderSet, RigidBodyBuilder, RigidBodySet, bodies_to_abi

    scene = scenes.sand_column_3d(16, 12, 16, y_offset=-4.0)  # y in [-1.75, 3.75]
    bodies, colliders = RigidBodySet(), ColliderSet()
    rb = bodies.insert(RigidBodyBuilder.fixed().translation([0.0, -4.0, 0.0]))
    colliders.insert_with_parent(ColliderBuilder.cuboid(100.0, 1.0, 100.0), rb, bodies)
    k = 0
    for ix in range(-2, 3):
        for iz in (-1.5, 0.0, 1.5):
            kind = k % 3
            if kind == 0:
                rb = bodies.insert(RigidBodyBuilder.fixed().translation([1.6 * ix, 0.2 + 0.4 * (k % 2), 1.7 * iz]))
                shape = ColliderBuilder.ball(0.55)
            elif kind == 1:
                rb = bodies.insert(RigidBodyBuilder.kinematic_velocity_based().translation([1.6 * ix, 1.0, 1.7 * iz])
                                   .rotation([0.3, 0.0, 0.2]).angvel([0.0, 2.0, 0.0]))
                shape = ColliderBuilder.capsule_y(0.5, 0.35)
            else:
                rb = bodies.insert(RigidBodyBuilder.dynamic().translation([1.6 * ix, 4.45, 1.7 * iz]).rotation([0.0, 0.4, 0.1]))
                shape = ColliderBuilder.cuboid(0.5, 0.3, 0.4).density(50.0)
            colliders.insert_with_parent(shape, rb, bodies)
            k += 1
    scene["bodies"] = bodies_to_abi(bodies, colliders, 3)
    assert len(scene["bodies"]) == 16
    parts, poses, vels = developed_state(oracle_mod, scene, 12)
    scene["bodies"]["translation"] = poses["translation"]
    scene["bodies"]["rotation"] = poses["rotation"]
    scene["bodies"]["linvel"] = vels["linear"]
    scene["bodies"]["angvel"] = vels["angular"]
    data, sim = one_substep_both(oracle_mod, pipe3, scene, parts, n=3)
    g, o = data.read_particles(), sim.read_particles()
    gb, gn = data.read_grid()
    ob, on = sim.read_grid()
    parity.assert_grid_close(gb, gn, ob, on, 1e-4)
    used = np.bitwise_or.reduce(on["cdf_affinities"].ravel())
    assert used & 0xFFFF == 0xFFFF and (used >> 16) != 0, "every collider must colour some node, some from inside"
    assert (o["cdf_affinity"] != 0).mean() > 0.2, "a good part of the particles must be collider-side"
    assert np.array_equal(g["cdf_affinity"], o["cdf_affinity"])
    parity.assert_particles_close(g, o, TOL, fields=("position", "def_grad"), tols={"position": 2e-6})
    assert parity.field_rel_err(g["velocity"], o["velocity"]) <= 5e-3  # stiff sand
    assert parity.field_rel_err(data.read_body_poses()["translation"], sim.read_body_poses()["translation"]) <= 1e-5
    assert parity.field_rel_err(data.read_body_vels()["linear"], sim.read_body_vels()["linear"]) <= 2e-3
    data.close()


def test_2d_sand_with_ball_and_capsule_colliders(pipe2, oracle_mod):
    """The reference's 2D sand scene at test size (sand2.rs:28-156): 2D Drucker-Prager (svd2 + the 2D return mapping,
    drucker_prager.wgsl:43-64) next to kinematic BALL and CAPSULE colliders, a rotating cuboid and a dynamic cuboid."""
    scene = scenes.sand_2d(60, 60)
    parts, poses, vels = developed_state(oracle_mod, scene, 120)  # the block has fallen onto the colliders
    assert (parts["cdf_affinity"] & 0b0110).any(), "the developed state must touch the ball and the capsule"
    bodies = scene["bodies"].copy()
    bodies["translation"] = poses["translation"]
    bodies["rotation"] = poses["rotation"]
    bodies["linvel"] = vels["linear"]
    bodies["angvel"] = vels["angular"]
    scene2 = dict(scene)
    scene2["bodies"] = bodies
    data, sim = one_substep_both(oracle_mod, pipe2, scene2, parts)
    g, o = data.read_particles(), sim.read_particles()
    gb, gn = data.read_grid()
    ob, on = sim.read_grid()
    parity.assert_grid_close(gb, gn, ob, on, TOL)
    parity.assert_particles_close(g, o, TOL, fields=("position", "velocity", "def_grad"), tols={"position": 2e-6})
    parity.assert_affine_close(g, o, 2, scene["cell_width"], float(scene["params"].dt))
    assert np.array_equal(g["cdf_affinity"], o["cdf_affinity"])
    for f in ("plastic_det", "plastic_hardening"):
        assert parity.field_rel_err(g[f], o[f]) <= 1e-5, f
    assert parity.field_rel_err(data.read_body_poses()["translation"], sim.read_body_poses()["translation"]) <= 1e-6
    data.close()
    # and 40 more substeps: positions stay within the sand bound
    data, sim = one_substep_both(oracle_mod, pipe2, scene2, parts, n=40)
    g, o = data.read_particles(), sim.read_particles()
    assert parity.field_rel_err(g["position"], o["position"]) <= 1e-4
    data.close()
