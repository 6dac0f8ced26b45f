"""numpy / ctypes mirrors of the POD structs in include/b200mpm.h.

Every dtype here is packed 4-byte fields, so `arr.ctypes.data` can be handed to the C ABI
directly. Field names follow the reference's Rust structs (src/solver/particle3d.rs:16-60,
src/models/mod.rs:63-68, src/models/drucker_prager.rs:6-42, src/solver/particle_update.rs:37-42).
"""
import ctypes

import numpy as np

MAX_BODIES = 16  # rigid_impulses.rs:42
NONE = 0xFFFFFFFF  # grid.wgsl:80

MODEL_COROTATED = 0
MODEL_NEO_HOOKEAN = 1

SHAPE_BALL = 0
SHAPE_CUBOID = 1
SHAPE_CAPSULE = 2
SHAPE_TRIMESH = 3  # 3D mesh colliders act through their sample points (MpmData.set_rigid_particles)
SHAPE_POLYLINE = 4  # 2D

PASS_NAMES = (  # src/pipeline.rs:201-271
    "update rigid particles",
    "grid sort",
    "grid_update_cdf",
    "p2g_cdf",
    "g2p_cdf",
    "p2g",
    "grid_update",
    "g2p",
    "particles_update",
    "integrate_bodies",
)

KERNEL_NAMES = (  # include/b200mpm.h B200MPM_KERNEL_*
    "touch", "count", "scan", "block_prepare", "scatter", "g2p_cdf", "p2g_cpic", "p2g", "begin", "g2p",
    "integrate_bodies", "rigid", "shard_migrate", "shard_halo", "shard_end",
)

sim_params_dtype = np.dtype([("gravity", "<f4", 3), ("dt", "<f4")])

particle_dtype = np.dtype(
    [
        ("position", "<f4", 3),
        ("velocity", "<f4", 3),
        ("def_grad", "<f4", 9),
        ("affine", "<f4", 9),
        ("cdf_normal", "<f4", 3),
        ("cdf_rigid_vel", "<f4", 3),
        ("cdf_signed_distance", "<f4"),
        ("cdf_affinity", "<u4"),
        ("init_volume", "<f4"),
        ("init_radius", "<f4"),
        ("mass", "<f4"),
        ("lambda", "<f4"),
        ("mu", "<f4"),
        ("dp_h0", "<f4"),
        ("dp_h1", "<f4"),
        ("dp_h2", "<f4"),
        ("dp_h3", "<f4"),
        ("dp_lambda", "<f4"),
        ("dp_mu", "<f4"),
        ("plastic_det", "<f4"),
        ("plastic_hardening", "<f4"),
        ("plastic_log_vol_gain", "<f4"),
        ("phase", "<f4"),
        ("max_stretch", "<f4"),
        ("model", "<u4"),
    ]
)
assert particle_dtype.itemsize == 196

body_dtype = np.dtype(
    [
        ("shape_type", "<u4"),
        ("shape_a", "<f4", 3),
        ("shape_b", "<f4", 3),
        ("radius", "<f4"),
        ("translation", "<f4", 3),
        ("rotation", "<f4", 4),
        ("linvel", "<f4", 3),
        ("angvel", "<f4", 3),
        ("inv_mass", "<f4", 3),
        ("inv_inertia", "<f4", 9),
        ("local_com", "<f4", 3),
        ("two_ways", "<u4"),
    ]
)
assert body_dtype.itemsize == 37 * 4

pose_dtype = np.dtype([("translation", "<f4", 3), ("rotation", "<f4", 4)])
velocity_dtype = np.dtype([("linear", "<f4", 3), ("angular", "<f4", 3)])
block_info_dtype = np.dtype([("vid", "<i4", 3), ("first_particle", "<u4"), ("num_particles", "<u4")])
# InstanceData of the testbed's vertex buffer (src_testbed/instancing3d.rs:66-73)
instance_dtype = np.dtype([("deformation", "<f4", (3, 4)), ("position", "<f4", 4), ("base_color", "<f4", 4), ("color", "<f4", 4)])
assert instance_dtype.itemsize == 96
RENDER_DEFAULT, RENDER_VOLUME, RENDER_VELOCITY, RENDER_CDF_NORMALS, RENDER_CDF_DISTANCES, RENDER_CDF_SIGNS = range(6)
node_dtype = np.dtype(
    [
        ("momentum_velocity_mass", "<f4", 4),
        ("cdf_distance", "<f4"),
        ("cdf_affinities", "<u4"),
        ("cdf_closest_id", "<u4"),
    ]
)


def ptr(arr):
    """void* of a C-contiguous numpy array (or None)."""
    if arr is None:
        return ctypes.c_void_p(0)
    assert arr.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(arr.ctypes.data)
