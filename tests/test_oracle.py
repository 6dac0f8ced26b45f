"""CPU tests of the oracle (oracle/): the reference's only known-answer test (prefix sum), the restated
un-vendored arithmetic against numpy in float64, physical invariants, and the committed golden fixtures."""
import os

import numpy as np
import pytest

from wgsparkl_b200 import abi, scenes
from wgsparkl_b200.models import ElasticCoefficients
from wgsparkl_b200.solver import SimulationParams, make_particles

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_prefix_sum_known_answer(oracle_mod):
    """src/grid/prefix_sum.rs:180-230: LEN 15071, ones / iota / random % 10000; expected = eval_cpu
    (prefix_sum.rs:71-83: inclusive scan shifted right by one, v[0] = 0)."""
    LEN = 15071
    rng = np.random.default_rng(0)
    for v in (np.ones(LEN, np.uint32), np.arange(LEN, dtype=np.uint32),
              (rng.integers(0, 2**32, LEN, dtype=np.uint64) % 10_000).astype(np.uint32)):
        ref = v.copy()
        for i in range(LEN - 1):  # literal eval_cpu
            ref[i + 1] += ref[i]
        ref[1:] = ref[:-1].copy()
        ref[0] = 0
        assert np.array_equal(oracle_mod.prefix_sum(v), ref)


@pytest.mark.parametrize("dim", [2, 3])
def test_svd_against_numpy(oracle_mod, dim):
    rng = np.random.default_rng(1)
    for k in range(200):
        scale = [1.0, 0.3, 1e-3][k % 3]
        F = (np.eye(dim) + scale * rng.uniform(-1, 1, (dim, dim))).astype(np.float32)
        if k % 7 == 0:
            F[:, 0] *= -1.0  # det < 0: sign carried by the last singular value
        U, S, Vt = oracle_mod.svd(F)
        assert np.allclose(U @ np.diag(S) @ Vt, F, atol=2e-6 * max(1.0, np.abs(F).max()))
        assert np.allclose(U @ U.T, np.eye(dim), atol=1e-6) and np.allclose(Vt @ Vt.T, np.eye(dim), atol=1e-6)
        assert np.linalg.det(U.astype(np.float64)) > 0 and np.linalg.det(Vt.astype(np.float64)) > 0
        ref = np.linalg.svd(F.astype(np.float64), compute_uv=False)
        assert np.allclose(np.sort(np.abs(S))[::-1], ref, rtol=2e-7, atol=1e-7)
        assert np.sign(np.prod(S)) == np.sign(np.linalg.det(F.astype(np.float64)))


@pytest.mark.parametrize("dim", [2, 3])
def test_kirchhoff_stress_against_float64(oracle_mod, dim):
    """linear_elasticity.wgsl:14-41 and neo_hookean_elasticity.wgsl:11-26 restated in numpy float64."""
    rng = np.random.default_rng(2)
    lam, mu = 2.0e5, 3.0e5
    for _ in range(50):
        F = (np.eye(dim) + 0.2 * rng.uniform(-1, 1, (dim, dim))).astype(np.float32)
        F64 = F.astype(np.float64)
        U, S, Vt = np.linalg.svd(F64)
        J = np.prod(S)
        tau = 2 * mu * (U @ np.diag(S - 1) @ Vt) @ F64.T + lam * (J - 1) * J * np.eye(dim)
        got = oracle_mod.kirchoff_stress(F, lam, mu, abi.MODEL_COROTATED)
        assert np.allclose(got, tau, rtol=2e-4, atol=2e-4 * mu * 1e-2)
        Jd = max(np.linalg.det(F64), 1e-10)
        tau_nh = mu * F64 @ F64.T + (lam * np.log(Jd) - mu) * np.eye(dim)
        got = oracle_mod.kirchoff_stress(F, lam, mu, abi.MODEL_NEO_HOOKEAN)
        assert np.allclose(got, tau_nh, rtol=1e-5, atol=1e-5 * mu)


def test_drucker_prager_invariants(oracle_mod):
    """drucker_prager.wgsl:112-158: lambda == 0 disables; tensile states snap to the identity stretch;
    compressed states are projected onto the cone (volume preserved in log space, deviator shrunk)."""
    from wgsparkl_b200.models import DruckerPrager

    dp = DruckerPrager.new(2.0e9, 0.2)
    pl = [dp.h0, dp.h1, dp.h2, dp.h3, dp.lambda_, dp.mu]
    F = np.diag([1.1, 1.05, 1.2]).astype(np.float32)  # expansion: trace(strain) > 0
    Fp, st = oracle_mod.dp_project(F, pl, [1.0, 1.0, 0.0])
    assert np.allclose(np.linalg.svd(Fp.astype(np.float64), compute_uv=False), 1.0, atol=1e-6)
    assert st[1] > 1.0 and np.isclose(st[2], np.log(np.linalg.det(F.astype(np.float64))), atol=1e-5)
    F = np.diag([0.9, 0.99, 0.97]).astype(np.float32)  # compression with shear
    Fp, st = oracle_mod.dp_project(F, pl, [1.0, 1.0, 0.0])
    e0, e1 = np.log(np.diag(F.astype(np.float64))), np.log(np.diag(Fp.astype(np.float64)))
    assert np.isclose(e0.sum(), e1.sum(), atol=1e-5)  # return mapping is deviatoric
    assert np.linalg.norm(e1 - e1.mean()) <= np.linalg.norm(e0 - e0.mean()) + 1e-7
    pl0 = list(pl)
    pl0[4] = 0.0
    Fp, st = oracle_mod.dp_project(F, pl0, [1.0, 1.0, 0.0])
    assert np.array_equal(Fp, F) and np.array_equal(st, np.float32([1.0, 1.0, 0.0]))


def _body(shape_type, a=(0, 0, 0), b=(0, 0, 0), radius=0.0, t=(0, 0, 0), rot=(0, 0, 0, 1)):
    o = np.zeros((), dtype=abi.body_dtype)
    o["shape_type"], o["shape_a"], o["shape_b"], o["radius"], o["translation"], o["rotation"] = shape_type, a, b, radius, t, rot
    return o


def test_shape_projection(oracle_mod):
    """wgparry Shape::projectPointOnBoundary contract for cuboid / ball / capsule."""
    cub = _body(abi.SHAPE_CUBOID, a=(2, 1, 3), t=(1, 0, 0))
    p, inside = oracle_mod.project_point(3, cub, (1.0 + 5.0, 0.5, 0.0))
    assert not inside and np.allclose(p, (3.0, 0.5, 0.0))
    p, inside = oracle_mod.project_point(3, cub, (1.0 + 0.5, 0.8, 0.0))  # inside: nearest face is +y
    assert inside and np.allclose(p, (1.5, 1.0, 0.0))
    ball = _body(abi.SHAPE_BALL, radius=2.0)
    p, inside = oracle_mod.project_point(3, ball, (0.0, 0.5, 0.0))
    assert inside and np.allclose(p, (0, 2, 0))
    p, inside = oracle_mod.project_point(3, ball, (3.0, 4.0, 0.0))
    assert not inside and np.allclose(p, (1.2, 1.6, 0.0), atol=1e-6)
    cap = _body(abi.SHAPE_CAPSULE, a=(0, -1, 0), b=(0, 1, 0), radius=0.5)
    p, inside = oracle_mod.project_point(3, cap, (2.0, 0.3, 0.0))
    assert not inside and np.allclose(p, (0.5, 0.3, 0.0), atol=1e-6)
    p, inside = oracle_mod.project_point(3, cap, (0.0, 3.0, 0.0))
    assert not inside and np.allclose(p, (0.0, 1.5, 0.0), atol=1e-6)
    # rotated cuboid, 2D: rotation by 90 degrees swaps the half extents
    c2 = _body(abi.SHAPE_CUBOID, a=(2, 1, 0), rot=(0, 1, 0, 0))
    p, inside = oracle_mod.project_point(2, c2, (0.0, 5.0))
    assert not inside and np.allclose(p, (0.0, 2.0), atol=1e-6)


def test_mass_and_momentum_conservation(oracle_mod):
    """P2G weights are a partition of unity: grid mass == particle mass and, with no forces, grid momentum ==
    particle momentum (p2g.wgsl:188-230 with APIC affine = 0)."""
    rng = np.random.default_rng(3)
    pos = rng.uniform(-3, 3, (500, 3)).astype(np.float32)
    vel = rng.uniform(-1, 1, (500, 3)).astype(np.float32)
    parts = make_particles(pos, 3, 0.25, 3.0, ElasticCoefficients.from_young_modulus(1e5, 0.3), velocity=vel)
    sim = oracle_mod.OracleSim(3, SimulationParams([0, 0, 0], 1e-3), parts, None, 1.0, 4096)
    for stage in (0, 1, 2, 4, 5):
        sim.stage(stage)
    _, nodes = sim.read_grid()
    mvm = nodes["momentum_velocity_mass"].astype(np.float64)
    assert np.isclose(mvm[..., 3].sum(), parts["mass"].astype(np.float64).sum(), rtol=1e-6)
    mom = (parts["mass"][:, None] * vel).astype(np.float64).sum(0)
    assert np.allclose(mvm[..., :3].sum((0, 1)), mom, rtol=1e-5, atol=1e-5)


def test_free_fall_matches_gravity(oracle_mod):
    scene = scenes.elastic_cube_3d(6, y_offset=30.0, ground=False, jitter=False)
    sim = oracle_mod.OracleSim(3, scene["params"], scene["particles"], scene["bodies"], 1.0, 4096)
    sim.step(10)
    out = sim.read_particles()
    g, dt = scene["params"].gravity[1], scene["params"].dt
    assert np.allclose(out["velocity"][:, 1], g * dt * 10, rtol=1e-4)
    assert np.allclose(out["velocity"][:, [0, 2]], 0.0, atol=1e-5)


def test_block_activation_includes_empty_neighbours(oracle_mod):
    """touch_particle_blocks activates the particle's block and its {0,1}^3 neighbours (grid.wgsl:300-320)."""
    parts = make_particles(np.float32([[0.4, 0.4, 0.4]]), 3, 0.25, 1.0, ElasticCoefficients.from_young_modulus(1e5, 0.3))
    sim = oracle_mod.OracleSim(3, SimulationParams([0, 0, 0], 1e-3), parts, None, 1.0, 64)
    sim.sort_only()
    blocks, _ = sim.read_grid()
    assert len(blocks) == 8 and blocks["num_particles"].sum() == 1
    vids = {tuple(v) for v in blocks["vid"]}
    assert vids == {(-1 + a, -1 + b, -1 + c) for a in (0, 1) for b in (0, 1) for c in (0, 1)}


def test_capacity_overflow_drops_blocks_silently(oracle_mod):
    scene = scenes.elastic_cube_3d(12, y_offset=3.0)
    sim = oracle_mod.OracleSim(3, scene["params"], scene["particles"], scene["bodies"], 1.0, 8)
    sim.sort_only()
    assert sim.overflowed() and sim.num_active_blocks() == 8


@pytest.mark.parametrize("name", ["elastic3d", "sand3d", "elastic2d", "coupled3d", "trimesh3d"])
def test_golden_fixtures(oracle_mod, name):
    """The oracle reproduces its committed outputs (tests/golden/make_golden.py): pins it against drift."""
    from golden.make_golden import CASES, run_case

    path = os.path.join(GOLDEN, name + ".npz")
    ref = np.load(path)
    got = run_case(oracle_mod, *CASES[name])
    for key in ref.files:
        a, b = got[key], ref[key]
        if a.dtype.kind in "iu":
            assert np.array_equal(a, b), key
        else:
            assert np.allclose(a, b, rtol=1e-5, atol=1e-6 * max(1.0, np.abs(b).max())), key


def test_exact_sigma_mode_bounds_reference_rounding_noise(oracle_mod):
    """The exact-sigma diagnostic mode only removes the f32 rounding of the singular values: one substep from an
    undeformed state is identical, and after 10 substeps of stiff sand the two modes differ by the noise level
    the GPU tolerances are built on, not more."""
    import parity

    def run(mode, n):
        s = scenes.sand_column_3d(6, 10, 6, y_offset=-5.0)
        oracle_mod.lib().oracle_set_exact_sigma(mode)
        try:
            sim = oracle_mod.OracleSim(s["dim"], s["params"], s["particles"], s["bodies"], s["cell_width"], s["grid_capacity"])
            sim.step(n)
            out = sim.read_particles()
            sim.close()
        finally:
            oracle_mod.lib().oracle_set_exact_sigma(0)
        return out

    a, b = run(0, 1), run(1, 1)
    assert np.array_equal(a["velocity"], b["velocity"]) and np.array_equal(a["def_grad"], b["def_grad"])
    a, b = run(0, 10), run(1, 10)
    assert 0.0 < parity.field_rel_err(a["velocity"], b["velocity"]) <= 5e-3
    assert parity.field_rel_err(a["position"], b["position"]) <= 2e-6


def test_mesh_sampling_restatement():
    """sample_triangle / sample_edge / sample_mesh (particle3d.rs:251-428): samples lie in their triangle, edges are
    sampled once, and a cell_width-spaced grid over a big triangle has no unsampled cell."""
    from wgsparkl_b200.rapier import sample_mesh, sample_polyline

    v = np.array([[0, 0, 0], [10, 0, 0], [0, 8, 0], [10, 8, 3]], dtype=np.float32)
    f = np.array([[0, 1, 2], [1, 3, 2]], dtype=np.uint32)
    pts = sample_mesh(v, f, 1.0)
    assert len(pts) > 100
    for p, t in pts:
        a, b, c = (v[int(i)].astype(np.float64) for i in f[t])
        n = np.cross(b - a, c - a)
        assert abs(np.dot(n, p - a)) / np.linalg.norm(n) < 1e-4  # in the plane
        l = np.linalg.solve(np.stack([b - a, c - a, n], axis=1), p - a)
        assert l[0] > -1e-4 and l[1] > -1e-4 and l[0] + l[1] < 1 + 1e-4  # in the triangle
    cells = {(int(np.floor(p[0])), int(np.floor(p[1]))) for p, t in pts if t == 0}
    inside = {(i, j) for i in range(10) for j in range(8) if (i + 1) / 10 + (j + 1) / 8 < 1.0}
    assert inside <= cells, "every grid cell fully inside the triangle holds a sample"
    edge_pts = [p for p, t in pts if abs(p[1] - (8 - 0.8 * p[0])) < 1e-4 and abs(p[2]) < 1e-6]
    assert 5 < len(edge_pts) < 25  # the shared edge (1, 2) is sampled once, not twice
    poly = sample_polyline(np.array([[0, 0], [3, 0], [3, 4]], dtype=np.float32), np.array([[0, 1], [1, 2]]), 1.0)
    assert len(poly) == (1 + 4 + 1) + (1 + 5 + 1)  # a, shifts 0..floor(len), b per segment (particle2d.rs:206-234)


def test_trimesh_box_agrees_with_analytic_cuboid(oracle_mod):
    """p2g_cdf restatement: a box given as a triangle mesh colours the nodes near its faces with the same distance,
    and (away from the surface plane itself) the same inside/outside sign, as the analytic cuboid of collide()."""
    from wgsparkl_b200.rapier import ColliderBuilder, ColliderSet, RigidBodyBuilder, RigidBodySet, bodies_to_abi, rigid_particles_to_abi

    he = (12.0, 3.0, 12.0)
    grids = {}
    for mesh in (False, True):
        scene = scenes.elastic_cube_3d(10, y_offset=-5.0)
        bodies, colliders = RigidBodySet(), ColliderSet()
        rb = bodies.insert(RigidBodyBuilder.fixed().translation([0.0, -6.2, 0.0]))
        shape = ColliderBuilder.trimesh(*scenes.box_trimesh(he)) if mesh else ColliderBuilder.cuboid(*he)
        colliders.insert_with_parent(shape, rb, bodies)
        scene["bodies"] = bodies_to_abi(bodies, colliders, 3)
        sim = oracle_mod.OracleSim(3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
        if mesh:
            sim.set_rigid_particles(*rigid_particles_to_abi(bodies, colliders, 3, scene["cell_width"]))
        for st in range(4):
            sim.stage(st)
        blocks, nodes = sim.read_grid()
        grids[mesh] = {tuple(b["vid"]): nodes[i] for i, b in enumerate(blocks)}
        sim.close()
    assert len(grids[True]) > len(grids[False]), "sample points activate blocks below the particles' own"
    n_both = 0
    for vid, na in grids[False].items():
        nb = grids[True][vid]
        both = (na["cdf_affinities"] != 0) & (nb["cdf_affinities"] != 0)
        n_both += int(both.sum())
        assert np.allclose(na["cdf_distance"][both], nb["cdf_distance"][both], atol=1e-5)
        assert np.array_equal(na["cdf_affinities"][both], nb["cdf_affinities"][both])
    assert n_both > 300


def test_heightfield_collider_holds_sand(oracle_mod):
    """heightfield3-style scene: sand dropped on a heightfield (converted to a trimesh on the host, as
    particle3d.rs:123-132 does) comes to rest ON it: no particle ends up below the surface."""
    from wgsparkl_b200.rapier import ColliderBuilder, ColliderSet, RigidBodyBuilder, RigidBodySet, bodies_to_abi, rigid_particles_to_abi

    scene = scenes.sand_column_3d(8, 8, 8, y_offset=2.0)
    n = 9
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    heights = 0.25 * np.sin(0.9 * ii) * np.cos(0.7 * jj)
    bodies, colliders = RigidBodySet(), ColliderSet()
    rb = bodies.insert(RigidBodyBuilder.fixed().translation([0.03, 0.41, -0.02]))
    colliders.insert_with_parent(ColliderBuilder.heightfield(heights, (16.0, 1.0, 16.0)), rb, bodies)
    scene["bodies"] = bodies_to_abi(bodies, colliders, 3)
    rp = rigid_particles_to_abi(bodies, colliders, 3, scene["cell_width"])
    assert len(rp[2]) > 500 and len(rp[0]) == n * n
    sim = oracle_mod.OracleSim(3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    sim.set_rigid_particles(*rp)
    sim.step(400)
    p = sim.read_particles()
    sim.close()
    assert np.all(np.isfinite(p["position"]))
    assert (p["cdf_affinity"] != 0).sum() > 30, "the bottom layer must feel the heightfield"
    assert p["position"][:, 1].min() > 0.41 - 0.25 - 0.3, "no particle falls through the surface"
    assert np.abs(p["velocity"][:, 1]).mean() < 2.0, "the pile is held (free fall would be ~3.3 by now)"


def _numpy_mls_mpm_substep(pos, vel, F, C, mass, vol0, lam, mu, h, dt, gravity):
    """An INDEPENDENT float64 statement of one collider-free substep, written from the formulas of SURVEY §8
    (a1, a8, a9, a10, a13, a15, a16) with dense numpy arrays instead of blocks / tiles / lists: quadratic B-spline
    weights around the associated cell c = round(x/h) - 1, P2G of (affine dpt + m v, m), grid velocity with gravity
    and the +-h/dt clamp, G2P of v and grad v = inv_d sum w v_n (x) dpt, advection, F update, corotated Kirchhoff
    stress from numpy's SVD, new affine = grad v m - tau V0 inv_d dt."""
    n, d = pos.shape
    inv_d = 4.0 / (h * h)
    c = np.rint(pos / h) - 1.0  # ties-to-even, like WGSL round
    x = (pos - c * h) / h  # in [0.5, 1.5]
    w = np.stack([0.5 * (1.5 - x) ** 2, 0.75 - (x - 1.0) ** 2, 0.5 * (x - 0.5) ** 2], axis=0)  # [shift][particle][axis]
    ci = c.astype(np.int64)
    lo = ci.min(axis=0)
    dims = ci.max(axis=0) - lo + 3
    gm = np.zeros(tuple(dims))
    gp = np.zeros(tuple(dims) + (d,))
    import itertools

    shifts = list(itertools.product(range(3), repeat=d))

    def weight(s):
        wt = w[s[0], :, 0] * w[s[1], :, 1]
        return wt * w[s[2], :, 2] if d == 3 else wt

    for s in shifts:
        sv = np.array(s)
        wt = weight(s)
        dpt = (c + sv) * h - pos
        contrib = np.einsum("nij,nj->ni", C, dpt) + mass[:, None] * vel
        idx = tuple((ci + sv - lo).T)
        np.add.at(gm, idx, wt * mass)
        np.add.at(gp, idx, wt[:, None] * contrib)
    with np.errstate(divide="ignore", invalid="ignore"):
        gv = np.where(gm[..., None] > 0, (gp + gm[..., None] * gravity * dt) / gm[..., None], 0.0)
    gv = np.clip(gv, -h / dt, h / dt)
    v_new = np.zeros_like(vel)
    grad = np.zeros((n, d, d))
    for s in shifts:
        sv = np.array(s)
        wt = weight(s)
        dpt = (c + sv) * h - pos
        vn = gv[tuple((ci + sv - lo).T)]
        v_new += wt[:, None] * vn
        grad += (wt * inv_d)[:, None, None] * np.einsum("ni,nj->nij", vn, dpt)
    speed = np.linalg.norm(v_new, axis=1)
    v_new = np.where((speed > h / dt)[:, None], v_new / np.maximum(speed, 1e-300)[:, None] * h / dt, v_new)
    pos_new = pos + v_new * dt
    F_new = F + np.einsum("nij,njk->nik", grad * dt, F)
    U, S, Vt = np.linalg.svd(F_new)
    J = np.linalg.det(F_new)
    R = np.einsum("nij,njk->nik", U, Vt)
    tau = 2.0 * mu[:, None, None] * np.einsum("nij,nkj->nik", F_new - R, F_new) + (lam * (J - 1.0) * J)[:, None, None] * np.eye(d)
    C_new = grad * mass[:, None, None] - tau * (vol0 * inv_d * dt)[:, None, None]
    return pos_new, v_new, F_new, C_new


def test_oracle_against_independent_numpy_mpm(oracle_mod):
    """The C++ oracle (blocks, hash map, tiles, linked lists, f32) against a dense float64 numpy statement of the same
    physics, 5 substeps of a falling, spinning, pre-strained elastic cube without colliders: two independent
    transcriptions of the reference's formulas must agree to f32 accuracy."""
    scene = scenes.elastic_cube_3d(8, y_offset=3.0, ground=False)
    p = scene["particles"]
    rng = np.random.default_rng(9)
    ctr = p["position"].mean(axis=0)
    p["velocity"][:, 0] = -3.0 * (p["position"][:, 2] - ctr[2]) + 1.0
    p["velocity"][:, 2] = 3.0 * (p["position"][:, 0] - ctr[0])
    p["velocity"][:, 1] = -2.0
    p["def_grad"][:, :9] += rng.normal(0.0, 0.01, size=(len(p), 9)).astype(np.float32)
    sim = oracle_mod.OracleSim(3, scene["params"], p, scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    h, dt = float(scene["cell_width"]), float(scene["params"].dt)
    g = np.array(scene["params"].gravity, dtype=np.float64)
    pos, vel = p["position"].astype(np.float64), p["velocity"].astype(np.float64)
    F = p["def_grad"].astype(np.float64).reshape(-1, 3, 3).transpose(0, 2, 1)  # column-major storage -> [row][col]
    C = p["affine"].astype(np.float64).reshape(-1, 3, 3).transpose(0, 2, 1)
    for step in range(5):
        pos, vel, F, C = _numpy_mls_mpm_substep(pos, vel, F, C, p["mass"].astype(np.float64), p["init_volume"].astype(np.float64),
                                                p["lambda"].astype(np.float64), p["mu"].astype(np.float64), h, dt, g)
        sim.step(1)
        o = sim.read_particles()
        oF = o["def_grad"].astype(np.float64).reshape(-1, 3, 3).transpose(0, 2, 1)
        assert np.abs(o["position"] - pos).max() <= 2e-6 * np.abs(pos).max(), step
        assert np.abs(o["velocity"] - vel).max() <= 2e-5 * np.abs(vel).max(), step
        assert np.abs(oF - F).max() <= 1e-5, step
    sim.close()


def test_oracle_sand_against_independent_numpy_mpm(oracle_mod):
    """Same cross-check with the Drucker-Prager return mapping (drucker_prager.wgsl:112-158) in the loop: a loose
    sand block (E = 1e6 so that f32 rounding of the singular values stays small) sheared and dropped."""
    from wgsparkl_b200.models import ElasticCoefficients

    scene = scenes.sand_column_3d(8, 8, 8, y_offset=6.0)
    scene["bodies"] = scene["bodies"][:0]
    p = scene["particles"]
    el = ElasticCoefficients.from_young_modulus(1.0e6, 0.2)
    for f in ("lambda", "dp_lambda"):
        p[f] = el.lambda_
    for f in ("mu", "dp_mu"):
        p[f] = el.mu
    ctr = p["position"].mean(axis=0)
    p["velocity"][:, 0] = 6.0 * (p["position"][:, 1] - ctr[1])  # shear: yields
    p["velocity"][:, 1] = -1.0
    sim = oracle_mod.OracleSim(3, scene["params"], p, scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    h, dt = float(scene["cell_width"]), float(scene["params"].dt)
    g = np.array(scene["params"].gravity, dtype=np.float64)
    f64 = lambda name: p[name].astype(np.float64)  # noqa: E731
    pos, vel = f64("position"), f64("velocity")
    F = f64("def_grad").reshape(-1, 3, 3).transpose(0, 2, 1)
    C = f64("affine").reshape(-1, 3, 3).transpose(0, 2, 1)
    mass, vol0, lam, mu = f64("mass"), f64("init_volume"), f64("lambda"), f64("mu")
    h0, h1, h2, h3 = f64("dp_h0"), f64("dp_h1"), f64("dp_h2"), f64("dp_h3")
    q, lvg = f64("plastic_hardening"), f64("plastic_log_vol_gain")
    projected_any = 0
    for step in range(8):
        # gather / scatter / F update exactly as in the elastic statement, then project F and redo the stress
        pos, vel, Ftrial, _ = _numpy_mls_mpm_substep(pos, vel, F, C, mass, vol0, lam, mu, h, dt, g)
        # grad v * dt = (Ftrial - F) F^-1  (F update: Ftrial = F + dt grad F)
        gradv = np.einsum("nij,njk->nik", Ftrial - F, np.linalg.inv(F)) / dt
        U, S, Vt = np.linalg.svd(Ftrial)
        angle = h0 + (h1 * q - h3) * np.exp(-h2 * q)
        alpha = np.sqrt(2.0 / 3.0) * 2.0 * np.sin(angle) / (3.0 - np.sin(angle))
        eps = np.log(S) + (lvg / 3.0)[:, None]
        tr = eps.sum(axis=1)
        dev = eps - (tr / 3.0)[:, None]
        devn = np.linalg.norm(dev, axis=1)
        tension = (tr > 0.0) | (devn == 0.0)
        gamma = devn + (3.0 * lam + 2.0 * mu) / (2.0 * mu) * tr * alpha
        yielding = (~tension) & (gamma > 0.0)
        S_new = S.copy()
        S_new[tension] = 1.0
        with np.errstate(divide="ignore", invalid="ignore"):
            S_y = np.exp(eps - dev * (gamma / devn)[:, None])
        S_new[yielding] = S_y[yielding]
        changed = tension | yielding
        q = q + np.where(tension, np.linalg.norm(eps, axis=1), np.where(yielding, gamma, 0.0))
        lvg = lvg + np.where(changed, np.log(S.prod(axis=1)) - np.log(S_new.prod(axis=1)), 0.0)
        projected_any += int(yielding.sum())
        F = np.einsum("nij,nj,njk->nik", U, S_new, Vt)
        Un, Sn, Vtn = np.linalg.svd(F)
        R = np.einsum("nij,njk->nik", Un, Vtn)
        J = np.linalg.det(F)
        tau = 2.0 * mu[:, None, None] * np.einsum("nij,nkj->nik", F - R, F) + (lam * (J - 1.0) * J)[:, None, None] * np.eye(3)
        C = gradv * mass[:, None, None] - tau * (vol0 * (4.0 / (h * h)) * dt)[:, None, None]
        sim.step(1)
        o = sim.read_particles()
        oF = o["def_grad"].astype(np.float64).reshape(-1, 3, 3).transpose(0, 2, 1)
        assert np.abs(o["position"] - pos).max() <= 2e-6 * np.abs(pos).max(), step
        assert np.abs(o["velocity"] - vel).max() <= 2e-4 * np.abs(vel).max(), step
        assert np.abs(oF - F).max() <= 2e-5, step
        assert np.abs(o["plastic_hardening"] - q).max() <= 1e-4 * max(1.0, np.abs(q).max()), step
    assert projected_any > 100, "the shear must drive particles onto the yield surface"
    sim.close()


def test_cpic_reconstruction_recovers_plane_and_sphere(oracle_mod):
    """g2p_cdf (g2p_cdf.wgsl:124-250) reconstructs, per particle, the collider's signed distance and normal by an
    MLS fit through the coloured nodes. For a flat cuboid face and for a ball the exact answers are known."""
    from wgsparkl_b200.rapier import ColliderBuilder, ColliderSet, RigidBodyBuilder, RigidBodySet, bodies_to_abi

    # plane: ground cuboid with its top face at y = -3.2, particles from y = -2.25 upwards
    scene = scenes.elastic_cube_3d(10, y_offset=-5.0)
    scene["bodies"]["translation"][0, 1] = -4.2
    sim = oracle_mod.OracleSim(3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    for st in range(5):
        sim.stage(st)
    o = sim.read_particles()
    sim.close()
    near = o["cdf_affinity"] != 0
    assert near.sum() > 150
    assert np.all(o["cdf_normal"][near][:, 1] > 0.9999)
    assert np.abs(o["cdf_signed_distance"][near] - (o["position"][near][:, 1] + 3.2)).max() < 1e-3
    # ball of radius 2.3 below the block: normals are radial, distances are |x - c| - r
    scene = scenes.elastic_cube_3d(10, y_offset=-5.0, ground=False)
    bodies, colliders = RigidBodySet(), ColliderSet()
    centre = np.array([0.13, -4.4, -0.21])
    rb = bodies.insert(RigidBodyBuilder.fixed().translation(centre))
    colliders.insert_with_parent(ColliderBuilder.ball(2.3), rb, bodies)
    scene["bodies"] = bodies_to_abi(bodies, colliders, 3)
    sim = oracle_mod.OracleSim(3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    for st in range(5):
        sim.stage(st)
    o = sim.read_particles()
    sim.close()
    near = o["cdf_affinity"] != 0
    assert near.sum() > 20
    r = o["position"][near].astype(np.float64) - centre
    dist = np.linalg.norm(r, axis=1)
    # a linear fit through samples of a curved distance field (second-order in h / radius), and only the nodes within
    # 1.5 h of the surface are coloured: judge the particles whose whole stencil is supported, i.e. within one cell
    close = (dist - 2.3) < 1.0
    assert close.sum() > 10
    # the field is reconstructed RELATIVE to the particle's own colour (g2p_cdf.wgsl:205-214): a particle that starts
    # coloured "inside" (sign bit of collider 0; decided by the weighted node signs, g2p_cdf.wgsl:160-188, so it can
    # differ from the geometric side right at the surface) sees positive distances inside and an inward normal
    side = np.where((o["cdf_affinity"][near] >> 16) & 1, -1.0, 1.0)
    assert np.abs(o["cdf_signed_distance"][near][close] - (side * (dist - 2.3))[close]).max() < 0.25  # ~ h^2 / (2 r)
    assert np.all(side[close] * np.einsum("ni,ni->n", o["cdf_normal"][near][close], (r / dist[:, None])[close]) > 0.97)


def test_oracle_2d_against_independent_numpy_mpm(oracle_mod):
    """The 2D instantiation (8x8 blocks, 10x10 tiles, svd2) against the same dense float64 statement."""
    scene = scenes.elastic_block_2d(20)
    scene["bodies"] = scene["bodies"][:0]
    p = scene["particles"]
    rng = np.random.default_rng(4)
    ctr = p["position"].mean(axis=0)
    p["velocity"][:, 0] = -4.0 * (p["position"][:, 1] - ctr[1]) + 0.5
    p["velocity"][:, 1] = 4.0 * (p["position"][:, 0] - ctr[0]) - 1.0
    Fp = p["def_grad"].reshape(len(p), -1)
    Fp[:, :4] += rng.normal(0.0, 0.01, size=(len(p), 4)).astype(np.float32)
    sim = oracle_mod.OracleSim(2, scene["params"], p, scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    h, dt = float(scene["cell_width"]), float(scene["params"].dt)
    g = np.array(scene["params"].gravity, dtype=np.float64)[:2]
    pos, vel = p["position"][:, :2].astype(np.float64), p["velocity"][:, :2].astype(np.float64)
    F = p["def_grad"].reshape(len(p), -1)[:, :4].astype(np.float64).reshape(-1, 2, 2).transpose(0, 2, 1)
    C = p["affine"].reshape(len(p), -1)[:, :4].astype(np.float64).reshape(-1, 2, 2).transpose(0, 2, 1)
    for step in range(5):
        pos, vel, F, C = _numpy_mls_mpm_substep(pos, vel, F, C, p["mass"].astype(np.float64), p["init_volume"].astype(np.float64),
                                                p["lambda"].astype(np.float64), p["mu"].astype(np.float64), h, dt, g)
        sim.step(1)
        o = sim.read_particles()
        oF = o["def_grad"].reshape(len(p), -1)[:, :4].astype(np.float64).reshape(-1, 2, 2).transpose(0, 2, 1)
        assert np.abs(o["position"][:, :2] - pos).max() <= 2e-6 * np.abs(pos).max(), step
        assert np.abs(o["velocity"][:, :2] - vel).max() <= 5e-5 * np.abs(vel).max(), step
        assert np.abs(oF - F).max() <= 1e-5, step
    sim.close()


def test_polyline_box_agrees_with_analytic_cuboid_2d(oracle_mod):
    """2D p2g_cdf: a rectangle given as a closed polyline against the analytic cuboid of collide(). The polyline runs
    CLOCKWISE: then (-ab.y, ab.x) points outwards and p2g_cdf.wgsl:150 sets the sign bit for interior nodes, like
    collide() does (0x00010001 when inside). CPIC only compares particle and node signs, so the other orientation
    is equally valid - it just puts the bit on the other side."""
    from wgsparkl_b200.rapier import ColliderBuilder, ColliderSet, RigidBodyBuilder, RigidBodySet, bodies_to_abi, rigid_particles_to_abi

    hx, hy = 6.0, 0.6
    grids = {}
    for mesh in (False, True):
        scene = scenes.elastic_block_2d(30)
        scene["particles"]["position"][:, 1] -= 9.9
        bodies, colliders = RigidBodySet(), ColliderSet()
        rb = bodies.insert(RigidBodyBuilder.fixed().translation([1.51, -0.73]))
        if mesh:
            loop = np.array([[-hx, -hy], [-hx, hy], [hx, hy], [hx, -hy], [-hx, -hy]], dtype=np.float32)
            shape = ColliderBuilder.polyline(loop)
        else:
            shape = ColliderBuilder.cuboid(hx, hy)
        colliders.insert_with_parent(shape, rb, bodies)
        scene["bodies"] = bodies_to_abi(bodies, colliders, 2)
        sim = oracle_mod.OracleSim(2, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
        if mesh:
            sim.set_rigid_particles(*rigid_particles_to_abi(bodies, colliders, 2, scene["cell_width"]))
        for st in range(4):
            sim.stage(st)
        blocks, nodes = sim.read_grid()
        grids[mesh] = {tuple(b["vid"][:2]): nodes[i] for i, b in enumerate(blocks)}
        sim.close()
    n_both = n_same = 0
    for vid, na in grids[False].items():
        nb = grids[True].get(vid)
        if nb is None:
            continue
        both = (na["cdf_affinities"] != 0) & (nb["cdf_affinities"] != 0)
        n_both += int(both.sum())
        assert np.allclose(na["cdf_distance"][both], nb["cdf_distance"][both], atol=1e-5)
        n_same += int((na["cdf_affinities"][both] == nb["cdf_affinities"][both]).sum())
    assert n_both > 25 and n_same == n_both, (n_both, n_same)
