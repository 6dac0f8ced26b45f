"""ORACLE — TEST INFRASTRUCTURE ONLY (ctypes front-end of oracle/liboracle.so).

CPU restatement of wgsparkl's MPM substep (see oracle/mpm_oracle.hpp for the per-kernel
citations and the parity status: UNPINNED except for the exclusive prefix sum).
May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs only. The product package (wgsparkl_b200) never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

from wgsparkl_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile liboracle.so with the committed Makefile (g++ -O2 -ffp-contract=off -fopenmp)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "mpm_oracle.hpp")]
    srcs.append(os.path.join(_HERE, "..", "include", "b200mpm.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = ctypes.CDLL(so)
        L.oracle_create.restype = ctypes.c_void_p
        L.oracle_create.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                    ctypes.c_size_t, ctypes.c_float, ctypes.c_uint32]
        L.oracle_destroy.argtypes = [ctypes.c_void_p]
        L.oracle_substep.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.oracle_stage.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.oracle_read_particles.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_read_grid.restype = ctypes.c_size_t
        L.oracle_read_grid.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.oracle_num_active_blocks.restype = ctypes.c_uint32
        L.oracle_num_active_blocks.argtypes = [ctypes.c_void_p]
        L.oracle_overflowed.argtypes = [ctypes.c_void_p]
        L.oracle_read_sorted_ids.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        for name in ("oracle_read_body_poses", "oracle_read_body_vels", "oracle_write_body_poses",
                     "oracle_write_body_vels"):
            getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.oracle_write_sim_params.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_read_impulses.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_prefix_sum.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.oracle_svd.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 4
        L.oracle_kirchoff_stress.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                             ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_dp_project.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_project_point.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        L.oracle_set_exact_sigma.argtypes = [ctypes.c_int]
        L.oracle_prep_vertex_buffer.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
        L.oracle_set_rigid_particles.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.oracle_max_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def set_threads(n):
    lib().oracle_set_threads(int(n))


def max_threads():
    return int(lib().oracle_max_threads())


class OracleSim:
    """One simulation state (the reference's MpmData) advanced by the restated kernels."""

    def __init__(self, dim, params, particles, bodies, cell_width, grid_capacity):
        self.dim = dim
        self.n = len(particles)
        self.nb = 0 if bodies is None else len(bodies)
        particles = np.ascontiguousarray(particles, dtype=abi.particle_dtype)
        bodies = np.zeros(0, dtype=abi.body_dtype) if bodies is None else np.ascontiguousarray(bodies, dtype=abi.body_dtype)
        p = params.to_abi() if hasattr(params, "to_abi") else params
        self.capacity = 1
        while self.capacity < grid_capacity:
            self.capacity <<= 1
        self._h = lib().oracle_create(dim, abi.ptr(p), abi.ptr(particles), self.n, abi.ptr(bodies), self.nb,
                                      ctypes.c_float(cell_width), grid_capacity)
        if not self._h:
            raise ValueError("oracle_create failed (dim must be 2 or 3, at most 16 bodies)")

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def step(self, n=1):
        lib().oracle_substep(self._h, n)

    def stage(self, pass_id):
        lib().oracle_stage(self._h, pass_id)

    def sort_only(self):
        self.stage(1)

    def read_particles(self):
        out = np.zeros(self.n, dtype=abi.particle_dtype)
        lib().oracle_read_particles(self._h, abi.ptr(out))
        return out

    def set_rigid_particles(self, vertices, vertex_colliders, samples, ids):
        """GpuRigidParticles (mesh collider sample points): see wgsparkl_b200.rapier.rigid_particles_to_abi."""
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        vc = np.ascontiguousarray(vertex_colliders, dtype=np.uint32)
        sp = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1, 3)
        ii = np.ascontiguousarray(ids, dtype=np.uint32).reshape(-1, 4)
        lib().oracle_set_rigid_particles(self._h, abi.ptr(v), abi.ptr(vc), len(v), abi.ptr(sp), abi.ptr(ii), len(sp))

    def prep_vertex_buffer(self, instances, mode):
        """prep_vertex_buffer{2d,3d}.wgsl main: updates `instances` (n x abi.instance_dtype) in place."""
        assert instances.dtype == abi.instance_dtype and len(instances) == self.n
        lib().oracle_prep_vertex_buffer(self._h, int(mode), abi.ptr(instances))
        return instances

    def num_active_blocks(self):
        return int(lib().oracle_num_active_blocks(self._h))

    def overflowed(self):
        return bool(lib().oracle_overflowed(self._h))

    def read_grid(self):
        nb = self.num_active_blocks()
        blocks = np.zeros(nb, dtype=abi.block_info_dtype)
        nodes = np.zeros(nb * 64, dtype=abi.node_dtype)
        got = lib().oracle_read_grid(self._h, abi.ptr(blocks), abi.ptr(nodes), nb)
        assert got == nb
        return blocks, nodes.reshape(nb, 64)

    def read_sorted_ids(self):
        out = np.zeros(self.n, dtype=np.uint32)
        lib().oracle_read_sorted_ids(self._h, abi.ptr(out))
        return out

    def read_body_poses(self):
        out = np.zeros(self.nb, dtype=abi.pose_dtype)
        lib().oracle_read_body_poses(self._h, abi.ptr(out), self.nb)
        return out

    def read_body_vels(self):
        out = np.zeros(self.nb, dtype=abi.velocity_dtype)
        lib().oracle_read_body_vels(self._h, abi.ptr(out), self.nb)
        return out

    def write_body_poses(self, poses):
        poses = np.ascontiguousarray(poses, dtype=abi.pose_dtype)
        lib().oracle_write_body_poses(self._h, abi.ptr(poses), len(poses))

    def write_body_vels(self, vels):
        vels = np.ascontiguousarray(vels, dtype=abi.velocity_dtype)
        lib().oracle_write_body_vels(self._h, abi.ptr(vels), len(vels))

    def write_sim_params(self, params):
        p = params.to_abi() if hasattr(params, "to_abi") else params
        lib().oracle_write_sim_params(self._h, abi.ptr(p))

    def read_impulses(self):
        out = np.zeros((self.nb, 6), dtype=np.int32)
        lib().oracle_read_impulses(self._h, abi.ptr(out))
        return out


def prefix_sum(v):
    """WgPrefixSum::eval_cpu (src/grid/prefix_sum.rs:71-83)."""
    out = np.ascontiguousarray(v, dtype=np.uint32).copy()
    lib().oracle_prefix_sum(abi.ptr(out), len(out))
    return out


def svd(F):
    F = np.asarray(F, dtype=np.float32)
    d = F.shape[0]
    Fc = np.ascontiguousarray(F.T.reshape(-1))
    U = np.zeros(9, dtype=np.float32)
    S = np.zeros(3, dtype=np.float32)
    Vt = np.zeros(9, dtype=np.float32)
    lib().oracle_svd(d, abi.ptr(Fc), abi.ptr(U), abi.ptr(S), abi.ptr(Vt))
    return U[: d * d].reshape(d, d).T.copy(), S[:d].copy(), Vt[: d * d].reshape(d, d).T.copy()


def kirchoff_stress(F, lam, mu, model=abi.MODEL_COROTATED):
    F = np.asarray(F, dtype=np.float32)
    d = F.shape[0]
    Fc = np.ascontiguousarray(F.T.reshape(-1))
    out = np.zeros(9, dtype=np.float32)
    lib().oracle_kirchoff_stress(d, model, lam, mu, abi.ptr(Fc), abi.ptr(out))
    return out[: d * d].reshape(d, d).T.copy()


def dp_project(F, plasticity6, state3):
    F = np.asarray(F, dtype=np.float32)
    d = F.shape[0]
    Fc = np.zeros(9, dtype=np.float32)
    Fc[: d * d] = F.T.reshape(-1)
    pl = np.ascontiguousarray(plasticity6, dtype=np.float32)
    st = np.ascontiguousarray(state3, dtype=np.float32).copy()
    lib().oracle_dp_project(d, abi.ptr(pl), abi.ptr(st), abi.ptr(Fc))
    return Fc[: d * d].reshape(d, d).T.copy(), st


def project_point(dim, body, pt):
    b = np.ascontiguousarray(body, dtype=abi.body_dtype)
    p = np.zeros(3, dtype=np.float32)
    p[:dim] = np.asarray(pt, dtype=np.float32)[:dim]
    out = np.zeros(3, dtype=np.float32)
    inside = lib().oracle_project_point(dim, abi.ptr(b), abi.ptr(p), abi.ptr(out))
    return out[:dim].copy(), bool(inside)
