"""A minimal stand-in for the parts of rapier{2,3}d that wgsparkl's scene constructors use
(RigidBodyBuilder / ColliderBuilder / RigidBodySet / ColliderSet; e.g.
crates/wgsparkl3d/examples/sand3.rs:62-103) and for what GpuBodySet::from_rapier extracts
from them (src/pipeline.rs:141): shape, pose, velocity and mass properties per coupled
collider. In a Rust host the real rapier sets fill `b200mpm_body` instead."""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import abi

f32 = np.float32

FIXED, DYNAMIC, KINEMATIC_VELOCITY_BASED = 0, 1, 2


def quat_from_scaled_axis(v):
    v = np.asarray(v, dtype=np.float64)
    a = np.linalg.norm(v)
    if a == 0.0:
        return np.array([0, 0, 0, 1], dtype=np.float64)
    s = np.sin(a / 2) / a
    return np.array([v[0] * s, v[1] * s, v[2] * s, np.cos(a / 2)])


def quat_to_matrix(q):
    i, j, k, w = [float(x) for x in q]
    return np.array(
        [
            [1 - 2 * (j * j + k * k), 2 * (i * j - k * w), 2 * (i * k + j * w)],
            [2 * (i * j + k * w), 1 - 2 * (i * i + k * k), 2 * (j * k - i * w)],
            [2 * (i * k - j * w), 2 * (j * k + i * w), 1 - 2 * (i * i + j * j)],
        ]
    )


@dataclass
class RigidBody:
    body_type: int = FIXED
    translation: np.ndarray = field(default_factory=lambda: np.zeros(3))
    rotation: object = None  # 3D: scaled-axis vector; 2D: angle
    linvel: np.ndarray = field(default_factory=lambda: np.zeros(3))
    angvel: object = None

    def is_dynamic(self):
        return self.body_type == DYNAMIC


class RigidBodyBuilder:
    def __init__(self, body_type):
        self._rb = RigidBody(body_type=body_type)

    @staticmethod
    def fixed():
        return RigidBodyBuilder(FIXED)

    @staticmethod
    def dynamic():
        return RigidBodyBuilder(DYNAMIC)

    @staticmethod
    def kinematic_velocity_based():
        return RigidBodyBuilder(KINEMATIC_VELOCITY_BASED)

    def translation(self, t):
        tt = np.zeros(3)
        tt[: len(t)] = t
        self._rb.translation = tt
        return self

    def rotation(self, r):
        self._rb.rotation = r
        return self

    def linvel(self, v):
        vv = np.zeros(3)
        vv[: len(v)] = v
        self._rb.linvel = vv
        return self

    def angvel(self, w):
        self._rb.angvel = w
        return self

    def build(self):
        return self._rb


@dataclass
class Collider:
    shape_type: int
    shape_a: np.ndarray
    shape_b: np.ndarray
    radius: float
    density: float = 1.0
    parent: Optional[int] = None
    vertices: Optional[np.ndarray] = None  # trimesh / polyline colliders: (nv, dim) local vertices
    indices: Optional[np.ndarray] = None  # (nt, 3) triangles / (ns, 2) segments


class ColliderBuilder:
    def __init__(self, c):
        self._c = c

    @staticmethod
    def cuboid(*half_extents):
        a = np.zeros(3)
        a[: len(half_extents)] = half_extents
        return ColliderBuilder(Collider(abi.SHAPE_CUBOID, a, np.zeros(3), 0.0))

    @staticmethod
    def ball(radius):
        return ColliderBuilder(Collider(abi.SHAPE_BALL, np.zeros(3), np.zeros(3), float(radius)))

    @staticmethod
    def capsule_y(half_height, radius):
        return ColliderBuilder(
            Collider(abi.SHAPE_CAPSULE, np.array([0.0, -half_height, 0.0]), np.array([0.0, half_height, 0.0]), float(radius))
        )

    @staticmethod
    def trimesh(vertices, indices):
        """3D triangle mesh (massless here: use it on fixed / kinematic bodies, like the reference's examples)."""
        c = Collider(abi.SHAPE_TRIMESH, np.zeros(3), np.zeros(3), 0.0, density=0.0)
        c.vertices = np.asarray(vertices, dtype=np.float32).reshape(-1, 3)
        c.indices = np.asarray(indices, dtype=np.uint32).reshape(-1, 3)
        return ColliderBuilder(c)

    @staticmethod
    def heightfield(heights, scale):
        """3D heightfield, handed to the MPM side as the triangle mesh parry's `HeightField::to_trimesh` yields
        (particle3d.rs:123-132): heights[i, j] over a regular (rows along z, columns along x) grid centred on the
        origin and spanning scale.x by scale.z, heights multiplied by scale.y; two triangles per cell."""
        hgt = np.asarray(heights, dtype=np.float32)
        nrows, ncols = hgt.shape
        sx, sy, sz = (float(v) for v in scale)
        xs = (np.arange(ncols, dtype=np.float32) / np.float32(ncols - 1) - np.float32(0.5)) * np.float32(sx)
        zs = (np.arange(nrows, dtype=np.float32) / np.float32(nrows - 1) - np.float32(0.5)) * np.float32(sz)
        verts = np.stack([np.tile(xs, nrows), (hgt * np.float32(sy)).ravel(), np.repeat(zs, ncols)], axis=1).astype(np.float32)
        tris = []
        for i in range(nrows - 1):
            for j in range(ncols - 1):
                a, b = i * ncols + j, i * ncols + j + 1
                c, d = (i + 1) * ncols + j, (i + 1) * ncols + j + 1
                tris.append([a, c, b])  # normals point towards +y
                tris.append([b, c, d])
        return ColliderBuilder.trimesh(verts, np.array(tris, dtype=np.uint32))

    @staticmethod
    def polyline(vertices, indices=None):
        """2D polyline; indices default to consecutive vertices."""
        c = Collider(abi.SHAPE_POLYLINE, np.zeros(3), np.zeros(3), 0.0, density=0.0)
        c.vertices = np.asarray(vertices, dtype=np.float32).reshape(-1, 2)
        if indices is None:
            n = len(c.vertices)
            indices = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1)
        c.indices = np.asarray(indices, dtype=np.uint32).reshape(-1, 2)
        return ColliderBuilder(c)

    def density(self, d):
        self._c.density = float(d)
        return self

    def build(self):
        return self._c


class RigidBodySet:
    def __init__(self):
        self.bodies: List[RigidBody] = []

    def insert(self, rb):
        if isinstance(rb, RigidBodyBuilder):
            rb = rb.build()
        self.bodies.append(rb)
        return len(self.bodies) - 1

    def __getitem__(self, h):
        return self.bodies[h]

    def __len__(self):
        return len(self.bodies)


class ColliderSet:
    def __init__(self):
        self.colliders: List[Collider] = []

    def insert_with_parent(self, co, parent, bodies=None):
        if isinstance(co, ColliderBuilder):
            co = co.build()
        co.parent = parent
        self.colliders.append(co)
        return len(self.colliders) - 1

    def insert(self, co):
        if isinstance(co, ColliderBuilder):
            co = co.build()
        self.colliders.append(co)
        return len(self.colliders) - 1

    def __iter__(self):
        return iter(enumerate(self.colliders))

    def __getitem__(self, h):
        return self.colliders[h]

    def __len__(self):
        return len(self.colliders)


def _mass_properties(co: Collider, dim: int):
    """(mass, local_com, local inertia tensor) of a collider, as parry computes them."""
    rho = co.density
    if co.shape_type in (abi.SHAPE_TRIMESH, abi.SHAPE_POLYLINE):
        # parry gives a surface mesh zero mass; the reference only hangs meshes on fixed / kinematic bodies
        return 0.0, np.zeros(3), np.eye(3 if dim == 3 else 1)
    if co.shape_type == abi.SHAPE_CUBOID:
        he = co.shape_a
        if dim == 3:
            m = rho * 8.0 * he[0] * he[1] * he[2]
            inertia = np.diag([he[1] ** 2 + he[2] ** 2, he[0] ** 2 + he[2] ** 2, he[0] ** 2 + he[1] ** 2]) * m / 3.0
        else:
            m = rho * 4.0 * he[0] * he[1]
            inertia = np.array([[m * (he[0] ** 2 + he[1] ** 2) / 3.0]])
        return m, np.zeros(3), inertia
    if co.shape_type == abi.SHAPE_BALL:
        r = co.radius
        if dim == 3:
            m = rho * 4.0 / 3.0 * np.pi * r**3
            inertia = np.eye(3) * (2.0 / 5.0 * m * r * r)
        else:
            m = rho * np.pi * r * r
            inertia = np.array([[m * r * r / 2.0]])
        return m, np.zeros(3), inertia
    # capsule along local y: cylinder/rectangle + two half balls
    hh = abs(co.shape_b[1] - co.shape_a[1]) / 2.0
    r = co.radius
    if dim == 3:
        m_cyl = rho * np.pi * r * r * 2 * hh
        m_sph = rho * 4.0 / 3.0 * np.pi * r**3
        iy = m_cyl * r * r / 2.0 + m_sph * 2.0 / 5.0 * r * r
        ix = (
            m_cyl * (3 * r * r + 4 * hh * hh) / 12.0
            + m_sph * (2.0 / 5.0 * r * r + hh * hh + 3.0 / 8.0 * 2 * hh * r)
        )
        return m_cyl + m_sph, np.zeros(3), np.diag([ix, iy, ix])
    m_rect = rho * 2 * r * 2 * hh
    m_circ = rho * np.pi * r * r
    i = m_rect * ((2 * r) ** 2 + (2 * hh) ** 2) / 12.0 + m_circ * (r * r / 2.0 + hh * hh)
    return m_rect + m_circ, np.zeros(3), np.array([[i]])


def bodies_to_abi(bodies: RigidBodySet, colliders: ColliderSet, dim: int, coupling=None) -> np.ndarray:
    """What MpmData::new derives from the rapier sets (src/pipeline.rs:107-117,141): one entry per
    collider that has a parent body, in collider-set order, BodyCoupling::TwoWays."""
    if coupling is None:
        coupling = [(co.parent, h, True) for h, co in colliders if co.parent is not None]
    if len(coupling) > abi.MAX_BODIES:
        raise ValueError("CPIC supports at most 16 coupled colliders (rigid_impulses.rs:42)")
    out = np.zeros(len(coupling), dtype=abi.body_dtype)
    for i, (bh, ch, two_ways) in enumerate(coupling):
        rb, co = bodies[bh], colliders[ch]
        o = out[i]
        o["shape_type"] = co.shape_type
        o["shape_a"] = co.shape_a.astype(f32)
        o["shape_b"] = co.shape_b.astype(f32)
        o["radius"] = co.radius
        o["translation"] = rb.translation.astype(f32)
        if dim == 3:
            q = quat_from_scaled_axis(rb.rotation if rb.rotation is not None else np.zeros(3))
            o["rotation"] = q.astype(f32)
            w = np.zeros(3)
            if rb.angvel is not None:
                w[:] = rb.angvel
            o["angvel"] = w.astype(f32)
        else:
            a = float(rb.rotation) if rb.rotation is not None else 0.0
            o["rotation"] = np.array([np.cos(a), np.sin(a), 0, 0], dtype=f32)
            o["angvel"] = np.array([float(rb.angvel) if rb.angvel is not None else 0.0, 0, 0], dtype=f32)
        o["linvel"] = rb.linvel.astype(f32)
        m, com, inertia = _mass_properties(co, dim)
        if rb.body_type == DYNAMIC and m > 0:
            o["inv_mass"][:dim] = f32(1.0 / m)
            inv = np.zeros((3, 3))
            if dim == 3:
                inv = np.linalg.inv(inertia)
                o["inv_inertia"] = inv.T.reshape(-1).astype(f32)  # column-major
            else:
                o["inv_inertia"][0] = f32(1.0 / inertia[0, 0])
        o["local_com"] = com.astype(f32)
        o["two_ways"] = 1 if two_ways else 0
    return out


# ---- rigid particles: sample points of mesh colliders (GpuRigidParticles::from_rapier) -------------------------
_EPS = np.float32(1.0e-5)  # particle3d.rs:241


def sample_edge(a, b, xy_spacing, triangle_id, out):
    """particle3d.rs:300-320: points strictly after `a`, spacing / sqrt(2) apart."""
    a, b = a.astype(f32), b.astype(f32)
    ab = b - a
    edge_length = f32(np.linalg.norm(ab))
    if edge_length > _EPS:
        edge_dir = ab / edge_length
        spacing = f32(xy_spacing) / f32(np.sqrt(f32(2.0)))
        nsteps = int(np.ceil(edge_length / spacing))
        for i in range(1, nsteps):
            out.append((a + edge_dir * (spacing * f32(i)), triangle_id))


def sample_triangle(a, b, c, xy_spacing, triangle_id, out):
    """particle3d.rs:336-428: a grid oriented along the longest edge (base) and the height of the triangle."""
    a, b, c = a.astype(f32), b.astype(f32), c.astype(f32)
    d_ab, d_bc, d_ca = f32(np.linalg.norm(b - a)), f32(np.linalg.norm(c - b)), f32(np.linalg.norm(a - c))
    mx = max(d_ab, d_bc, d_ca)
    if mx == d_bc:
        a, b, c = b, c, a
    elif mx == d_ca:
        a, b, c = c, a, b
    ac = c - a
    base = b - a
    base_length = f32(np.linalg.norm(base))
    if not base_length > 0:
        return
    base_dir = base / base_length
    spacing = f32(xy_spacing) / f32(np.sqrt(f32(2.0)))
    base_step_count = np.ceil(base_length / spacing)
    base_step = base_dir * spacing
    ac_offset_length = f32(ac.dot(base_dir))
    bc_offset_length = base_length - ac_offset_length
    if ac_offset_length < _EPS or bc_offset_length < _EPS or base_length < _EPS:
        return
    height = ac - base_dir * ac_offset_length
    height_length = f32(np.linalg.norm(height))
    height_dir = height / height_length
    tan_alpha = height_length / ac_offset_length
    tan_beta = height_length / bc_offset_length
    for i in range(1, int(base_step_count)):
        base_position = a + f32(i) * base_step
        height_ac = tan_alpha * f32(np.linalg.norm(base_position - a))
        height_bc = tan_beta * f32(np.linalg.norm(base_position - b))
        hl = min(height_ac, height_bc)
        height_step_count = np.ceil(hl / spacing)
        height_step = height_dir * spacing
        for j in range(1, int(height_step_count)):
            p = base_position + f32(j) * height_step
            if np.all(np.isfinite(p)):
                out.append((p.astype(f32), triangle_id))


def sample_mesh(vertices, indices, xy_spacing):
    """particle3d.rs:251-292: every triangle's interior, and every edge once."""
    out, visited = [], set()
    for tri_id, idx in enumerate(indices):
        va, vb, vc = (vertices[int(idx[k])] for k in range(3))
        sample_triangle(va, vb, vc, xy_spacing, tri_id, out)
        for ia, ib in ((idx[0], idx[1]), (idx[1], idx[2]), (idx[2], idx[0])):
            key = (max(int(ia), int(ib)), min(int(ia), int(ib)))
            if key not in visited:
                visited.add(key)
                sample_edge(vertices[int(ia)], vertices[int(ib)], xy_spacing, tri_id, out)
    return out


def sample_polyline(vertices, indices, sampling_step):
    """particle2d.rs:206-234 (the end points and the i = 0 sample are duplicated, as in the reference)."""
    out = []
    for seg_id, idx in enumerate(indices):
        a, b = vertices[int(idx[0])].astype(f32), vertices[int(idx[1])].astype(f32)
        out.append((a, seg_id))
        ab = b - a
        length = f32(np.linalg.norm(ab))
        if length > 0:  # parry: direction() is None for a degenerate segment
            d = ab / length
            i = 0
            while True:
                shift = f32(i) * f32(sampling_step)
                if shift > length:
                    break
                out.append((a + d * shift, seg_id))
                i += 1
            out.append((b, seg_id))
    return out


def rigid_particles_to_abi(bodies: RigidBodySet, colliders: ColliderSet, dim: int, cell_width: float, coupling=None):
    """GpuRigidParticles::from_rapier (particle3d.rs:101-160, particle2d.rs:80-140) with the reference's sampling step
    (= cell_width, pipeline.rs:140): (vertices (nv,3) f32 local, vertex_colliders (nv,) u32, samples (ns,3) f32 local,
    ids (ns,4) u32 = vertex ids of the primitive + collider index). Collider index = position in the coupling list."""
    if coupling is None:
        coupling = [(co.parent, h, True) for h, co in colliders if co.parent is not None]
    verts, vcol, samples, ids = [], [], [], []
    for collider_id, (_, ch, _) in enumerate(coupling):
        co = colliders[ch]
        if co.vertices is None:
            continue
        base_vid = len(verts)
        for v in co.vertices:
            vv = np.zeros(3, dtype=f32)
            vv[:dim] = v[:dim]
            verts.append(vv)
            vcol.append(collider_id)
        if dim == 3:
            pts = sample_mesh(co.vertices, co.indices, cell_width)
        else:
            pts = sample_polyline(co.vertices, co.indices, cell_width)
        for p, prim in pts:
            pp = np.zeros(3, dtype=f32)
            pp[:dim] = p[:dim]
            samples.append(pp)
            idx = co.indices[prim]
            ids.append([base_vid + int(idx[0]), base_vid + int(idx[1]), base_vid + int(idx[2]) if dim == 3 else 0, collider_id])
    return (np.array(verts, dtype=f32).reshape(-1, 3), np.array(vcol, dtype=np.uint32),
            np.array(samples, dtype=f32).reshape(-1, 3), np.array(ids, dtype=np.uint32).reshape(-1, 4))
