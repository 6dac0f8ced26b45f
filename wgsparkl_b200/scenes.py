"""Synthetic scenes of BASELINE.json's configs, built the way the reference's scene
constructors do (crates/wgsparkl3d/examples/{sand3,elastic_cut3}.rs,
crates/wgsparkl2d/examples/elasticity2.rs), plus the reference's own test lattice
(src/pipeline.rs:296-331, src/grid/grid.rs:355-373).

Every scene is a dict: dim, params, particles (abi.particle_dtype array), bodies
(abi.body_dtype array), cell_width, grid_capacity, name, substeps_per_frame.
"""
import numpy as np

from . import abi
from .models import DruckerPrager, ElasticCoefficients
from .rapier import rigid_particles_to_abi  # noqa: F401  (re-exported for scene users)
from .rapier import ColliderBuilder, ColliderSet, RigidBodyBuilder, RigidBodySet, bodies_to_abi
from .solver import F32_MAX, ParticlePhase, SimulationParams, make_particles

SEED = 20261017


def _jitter(pos, h, jitter, seed):
    if jitter:
        rng = np.random.default_rng(seed)
        pos = pos + rng.uniform(-0.1 * h, 0.1 * h, size=pos.shape)
    return pos.astype(np.float32)


def _lattice3(nx, ny, nz, offset, h):
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    pos = np.stack([i.ravel() + 0.5 + offset[0], j.ravel() + 0.5 + offset[1], k.ravel() + 0.5 + offset[2]], axis=1)
    return pos * (h / 2.0)


def reference_test_lattice(n=10):
    """The 1000-particle lattice of the reference's only pipeline tests: positions (i,j,k)/2 lie exactly on
    round() ties; plasticity None / phase None exercises the lambda = mu = -1 quirk."""
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    pos = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float32) / 1.0 / 2.0
    parts = make_particles(pos, 3, 1.0 / 4.0, 1.0, ElasticCoefficients.from_young_modulus(100_000.0, 0.33))
    return dict(name="reference_test_lattice", dim=3, params=SimulationParams([0.0, -9.81, 0.0], (1.0 / 60.0) / 10.0),
                particles=parts, bodies=np.zeros(0, dtype=abi.body_dtype), cell_width=1.0, grid_capacity=100_000,
                substeps_per_frame=10)


def elastic_block_2d(n_side=100, jitter=True, seed=SEED):
    """Config 1: 2D elastic block drop (down-scaled elasticity2.rs:29-97)."""
    h = 0.2
    i, j = np.meshgrid(np.arange(n_side), np.arange(n_side), indexing="ij")
    pos = np.stack([i.ravel() + 0.5, j.ravel() + 0.5], axis=1) * (h / 2.0) + np.array([0.0, 10.0])
    pos = _jitter(pos, h, jitter, seed)
    parts = make_particles(pos, 2, h / 4.0, 1000.0, ElasticCoefficients.from_young_modulus(5_000_000.0, 0.2),
                           phase=ParticlePhase(1.0, F32_MAX))
    bodies, colliders = RigidBodySet(), ColliderSet()
    rb = bodies.insert(RigidBodyBuilder.fixed().translation([0.0, -1.0]))
    colliders.insert_with_parent(ColliderBuilder.cuboid(1000.0, 1.0), rb, bodies)
    return dict(name="2d_elastic_block_%d" % (n_side * n_side), dim=2,
                params=SimulationParams([0.0, -9.81 * 2.0], (1.0 / 60.0) / 15.0), particles=parts,
                bodies=bodies_to_abi(bodies, colliders, 2), cell_width=h, grid_capacity=60_000,
                substeps_per_frame=15)


def elastic_cube_3d(n_side=100, y_offset=60.0, jitter=True, seed=SEED, grid_capacity=60_000, ground=True, nx=None):
    """Config 2: 3D corotated-elastic cube dropped on a static ground cuboid (elastic_cut3.rs:28-71
    scaled, trimeshes removed). `nx` (default n_side) stretches the block along x for weak-scaling runs."""
    h = 1.0
    nx = nx or n_side
    pos = _lattice3(nx, n_side, n_side, (-nx / 2.0, y_offset, -n_side / 2.0), h)
    pos = _jitter(pos, h, jitter, seed)
    parts = make_particles(pos, 3, h / 4.0, 2700.0, ElasticCoefficients.from_young_modulus(10_000_000.0, 0.2),
                           phase=ParticlePhase(1.0, F32_MAX))
    bodies, colliders = RigidBodySet(), ColliderSet()
    if ground:
        rb = bodies.insert(RigidBodyBuilder.fixed().translation([0.0, -4.0, 0.0]))
        colliders.insert_with_parent(ColliderBuilder.cuboid(max(100.0, nx * 1.0), 1.0, max(100.0, n_side * 1.0)), rb, bodies)
    return dict(name="3d_elastic_cube_%d" % (nx * n_side * n_side), dim=3,
                params=SimulationParams([0.0, -9.81 * 4.0, 0.0], (1.0 / 60.0) / 20.0), particles=parts,
                bodies=bodies_to_abi(bodies, colliders, 3), cell_width=h, grid_capacity=grid_capacity,
                substeps_per_frame=20)


def sand_column_3d(nx=100, ny=400, nz=100, y_offset=10.0, jitter=True, seed=SEED, grid_capacity=60_000, walls=False):
    """Config 3: 3D Drucker-Prager sand column collapse (sand3.rs:28-68)."""
    h = 1.0
    pos = _lattice3(nx, ny, nz, (-nx / 2.0, y_offset, -nz / 2.0), h)
    pos = _jitter(pos, h, jitter, seed)
    parts = make_particles(pos, 3, h / 4.0, 2700.0, ElasticCoefficients.from_young_modulus(2_000_000_000.0, 0.2),
                           plasticity=DruckerPrager.new(2_000_000_000.0, 0.2), phase=None)
    bodies, colliders = RigidBodySet(), ColliderSet()
    rb = bodies.insert(RigidBodyBuilder.fixed().translation([0.0, -4.0, 0.0]))
    ext = max(100.0, nx * 2.0, nz * 2.0)
    colliders.insert_with_parent(ColliderBuilder.cuboid(ext, 4.0, ext), rb, bodies)
    if walls:  # sand3.rs:70-93
        for t, he in (([0.0, 5.0, -35.0], [35.0, 5.0, 0.5]), ([0.0, 5.0, 35.0], [35.0, 5.0, 0.5]),
                      ([-35.0, 5.0, 0.0], [0.5, 5.0, 35.0]), ([35.0, 5.0, 0.0], [0.5, 5.0, 35.0])):
            rb = bodies.insert(RigidBodyBuilder.fixed().translation(t))
            colliders.insert_with_parent(ColliderBuilder.cuboid(*he), rb, bodies)
    return dict(name="3d_sand_column_%d" % (nx * ny * nz), dim=3,
                params=SimulationParams([0.0, -9.81, 0.0], (1.0 / 60.0) / 20.0), particles=parts,
                bodies=bodies_to_abi(bodies, colliders, 3), cell_width=h, grid_capacity=grid_capacity,
                substeps_per_frame=20)


def mixed_coupled_3d(nx=40, ny=40, nz=40, jitter=True, seed=SEED, grid_capacity=60_000, n_dynamic=2):
    """Config 4 (scaled by the caller): half sand / half elastic (Neo-Hookean selector) with a kinematic
    rotating cuboid (sand3.rs:95-103) and dynamic cuboids (pattern of sand2.rs:149-156), two-way coupled."""
    h = 1.0
    pos = _lattice3(nx, ny, nz, (-nx / 2.0, 10.0, -nz / 2.0), h)
    pos = _jitter(pos, h, jitter, seed)
    sand = make_particles(pos, 3, h / 4.0, 2700.0, ElasticCoefficients.from_young_modulus(2_000_000_000.0, 0.2),
                          plasticity=DruckerPrager.new(2_000_000_000.0, 0.2), phase=None)
    solid = make_particles(pos, 3, h / 4.0, 2700.0, ElasticCoefficients.from_young_modulus(10_000_000.0, 0.2),
                           phase=ParticlePhase(1.0, F32_MAX), model_kind=abi.MODEL_NEO_HOOKEAN)
    parts = np.where((pos[:, 0] < 0.0)[:, None].repeat(1, axis=1).ravel(), sand, solid)
    bodies, colliders = RigidBodySet(), ColliderSet()
    rb = bodies.insert(RigidBodyBuilder.fixed().translation([0.0, -4.0, 0.0]))
    colliders.insert_with_parent(ColliderBuilder.cuboid(100.0, 4.0, 100.0), rb, bodies)
    rb = bodies.insert(RigidBodyBuilder.kinematic_velocity_based().translation([0.0, 2.0, 0.0])
                       .rotation([0.0, 0.0, -0.5]).angvel([0.0, -1.0, 0.0]))
    colliders.insert_with_parent(ColliderBuilder.cuboid(0.5, 2.0, min(30.0, nz / 4.0)), rb, bodies)
    top = (ny + 10.0) * h / 2.0
    for k in range(n_dynamic):
        rb = bodies.insert(RigidBodyBuilder.dynamic().translation([3.0 * k - 2.0, top + 3.0 + 3.0 * k, 0.0]))
        colliders.insert_with_parent(ColliderBuilder.cuboid(2.0, 0.5, 2.0).density(10.0 + 100.0 * k), rb, bodies)
    return dict(name="3d_mixed_coupled_%d" % (nx * ny * nz), dim=3,
                params=SimulationParams([0.0, -9.81, 0.0], (1.0 / 60.0) / 20.0), particles=parts,
                bodies=bodies_to_abi(bodies, colliders, 3), cell_width=h, grid_capacity=grid_capacity,
                substeps_per_frame=20)


def sand_dam_3d(nx=400, ny=200, nz=200, jitter=True, seed=SEED, grid_capacity=262_144):
    """Config 5: 3D sand dam break against the -x wall inside a 4-wall box (sand3.rs:70-93 scaled)."""
    h = 1.0
    half_x, half_z = nx * 0.5, nz * 0.5  # box half extents in world units (a bit larger than the dam)
    pos = _lattice3(nx, ny, nz, (-half_x * 2.0 + 2.0, 2.0, -nz / 2.0), h)
    pos = _jitter(pos, h, jitter, seed)
    parts = make_particles(pos, 3, h / 4.0, 2700.0, ElasticCoefficients.from_young_modulus(2_000_000_000.0, 0.2),
                           plasticity=DruckerPrager.new(2_000_000_000.0, 0.2), phase=None)
    bodies, colliders = RigidBodySet(), ColliderSet()
    wall_h = ny * 0.5
    specs = [([0.0, -4.0, 0.0], [half_x * 2.0 + 20.0, 4.0, half_z + 20.0]),
             ([0.0, wall_h, -half_z - 1.0], [half_x * 2.0, wall_h, 1.0]),
             ([0.0, wall_h, half_z + 1.0], [half_x * 2.0, wall_h, 1.0]),
             ([-half_x - 1.0, wall_h, 0.0], [1.0, wall_h, half_z]),
             ([half_x * 2.0, wall_h, 0.0], [1.0, wall_h, half_z])]
    for t, he in specs:
        rb = bodies.insert(RigidBodyBuilder.fixed().translation(t))
        colliders.insert_with_parent(ColliderBuilder.cuboid(*he), rb, bodies)
    return dict(name="3d_sand_dam_%d" % (nx * ny * nz), dim=3,
                params=SimulationParams([0.0, -9.81, 0.0], (1.0 / 60.0) / 20.0), particles=parts,
                bodies=bodies_to_abi(bodies, colliders, 3), cell_width=h, grid_capacity=grid_capacity,
                substeps_per_frame=20)


def box_trimesh(half_extents):
    """Closed, outward-oriented triangle mesh of an axis-aligned box (12 triangles)."""
    x, y, z = half_extents
    v = np.array([[-x, -y, -z], [x, -y, -z], [x, y, -z], [-x, y, -z], [-x, -y, z], [x, -y, z], [x, y, z], [-x, y, z]], dtype=np.float32)
    f = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [2, 3, 7], [2, 7, 6], [1, 2, 6], [1, 6, 5],
                  [0, 4, 7], [0, 7, 3]], dtype=np.uint32)
    return v, f


def elastic_cube_on_trimesh_3d(n_side=12, tilt=0.15, moving=True):
    """an elastic cube dropped on TRIANGLE-MESH colliders (the elastic_cut3 / heightfield3 pattern):
    a tilted fixed slab as ground and a kinematic, slowly turning bar cutting into the cube. The meshes are placed off
    the grid lines so that no node sits exactly on a triangle edge. Returns the scene with `rigid_particles`."""
    s = elastic_cube_3d(n_side, y_offset=-5.0, ground=False)
    bodies, colliders = RigidBodySet(), ColliderSet()
    rb = bodies.insert(RigidBodyBuilder.fixed().translation([0.137, -4.163, -0.071]).rotation([tilt, 0.0, 0.6 * tilt]))
    colliders.insert_with_parent(ColliderBuilder.trimesh(*box_trimesh((9.3, 1.1, 9.7))), rb, bodies)
    top = (n_side - 5.0) * 0.5
    rb = bodies.insert(RigidBodyBuilder.kinematic_velocity_based().translation([0.21, top + 0.35, 0.13]).rotation([0.0, 0.3, 0.2])
                       .linvel([0.0, -3.0 if moving else 0.0, 0.0]).angvel([0.0, 1.5 if moving else 0.0, 0.0]))
    colliders.insert_with_parent(ColliderBuilder.trimesh(*box_trimesh((0.45, 0.6, 4.3))), rb, bodies)
    s["bodies"] = bodies_to_abi(bodies, colliders, 3)
    s["rigid_particles"] = rigid_particles_to_abi(bodies, colliders, 3, s["cell_width"])
    s["name"] = "3d_elastic_cube_on_trimesh_%d" % len(s["particles"])
    return s


def elastic_block_on_polyline_2d(n_side=24):
    """2D counterpart: an elastic block on a polyline ground with a kink (the elastic_cut2 pattern)."""
    s = elastic_block_2d(n_side)
    s["particles"]["position"][:, 1] -= 9.9
    bodies, colliders = RigidBodySet(), ColliderSet()
    rb = bodies.insert(RigidBodyBuilder.fixed().translation([0.013, -0.237]))
    colliders.insert_with_parent(ColliderBuilder.polyline([[-30.0, 0.9], [-4.1, 0.23], [3.7, 0.31], [30.0, 1.4]]), rb, bodies)
    s["bodies"] = bodies_to_abi(bodies, colliders, 2)
    s["rigid_particles"] = rigid_particles_to_abi(bodies, colliders, 2, s["cell_width"])
    s["name"] = "2d_elastic_block_on_polyline_%d" % len(s["particles"])
    return s


def sand_2d(nx=60, ny=60, jitter=True, seed=SEED, n_dynamic=1):
    """The reference's 2D sand scene (crates/wgsparkl2d/examples/sand2.rs:28-156) at test size: Drucker-Prager sand
    (phase None: plasticity on) falling on a static platform, past kinematic rotating colliders of the three analytic
    shapes - cuboid, BALL and CAPSULE (sand2.rs:124-136) - and under dynamic cuboids (sand2.rs:149-156)."""
    h = 0.2
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    pos = np.stack([i.ravel() + 0.5, j.ravel() + 0.5], axis=1) * (h / 2.0) + np.array([0.0, 2.0])
    pos = _jitter(pos, h, jitter, seed)
    parts = make_particles(pos, 2, h / 4.0, 1000.0, ElasticCoefficients.from_young_modulus(10_000_000.0, 0.2),
                           plasticity=DruckerPrager.new(10_000_000.0, 0.2), phase=None)
    bodies, colliders = RigidBodySet(), ColliderSet()
    w = nx * h / 2.0
    rb = bodies.insert(RigidBodyBuilder.fixed().translation([w / 2.0, -1.0]))
    colliders.insert_with_parent(ColliderBuilder.cuboid(42.0, 1.0), rb, bodies)
    rb = bodies.insert(RigidBodyBuilder.kinematic_velocity_based().translation([0.25 * w, 1.2]).angvel(-1.0))
    colliders.insert_with_parent(ColliderBuilder.ball(0.7), rb, bodies)
    rb = bodies.insert(RigidBodyBuilder.kinematic_velocity_based().translation([0.75 * w, 1.2]).angvel(-1.0))
    colliders.insert_with_parent(ColliderBuilder.capsule_y(0.5, 0.3), rb, bodies)
    rb = bodies.insert(RigidBodyBuilder.kinematic_velocity_based().translation([0.5 * w, 1.0]).angvel(1.0))
    colliders.insert_with_parent(ColliderBuilder.cuboid(0.1, 0.8), rb, bodies)
    top = 2.0 + ny * h / 2.0
    for k in range(n_dynamic):
        rb = bodies.insert(RigidBodyBuilder.dynamic().translation([0.5 * w + 0.6 * k, top + 0.4]))
        colliders.insert_with_parent(ColliderBuilder.cuboid(0.5, 0.1).density(10.0 + 100.0 * k), rb, bodies)
    return dict(name="2d_sand_%d" % (nx * ny), dim=2, params=SimulationParams([0.0, -9.81], (1.0 / 60.0) / 10.0),
                particles=parts, bodies=bodies_to_abi(bodies, colliders, 2), cell_width=h, grid_capacity=60_000,
                substeps_per_frame=10)
