"""Host-side mirror of wgsparkl's `pipeline` module (src/pipeline.rs): `MpmPipeline` and
`MpmData`, bound to libb200mpm.so (include/b200mpm.h) through ctypes.

There is no CPU fallback: if the library has not been built, or no sm_100 CUDA device is
visible, constructing an `MpmPipeline` raises `B200MpmError`.
"""
import ctypes
import os
from typing import Optional, Sequence

import numpy as np

from . import abi
from .solver import SimulationParams, particles_to_abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200MPM_LIB: development override used by tools/build_variant.py for A/B kernel timing.
_LIB_PATH = os.environ.get("B200MPM_LIB") or os.path.join(_HERE, "libb200mpm.so")
_lib = None


class B200MpmError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("b200mpm error %d: %s" % (code, message))
        self.code = code


OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_NO_DEVICE = -2
ERR_CUDA = -3
ERR_OUT_OF_MEMORY = -4
ERR_GRID_OVERFLOW = -5

# Every symbol include/b200mpm.h declares (tests check that the library exports all of them).
EXPORTS = (
    "b200mpm_pipeline_create",
    "b200mpm_pipeline_destroy",
    "b200mpm_pipeline_set_stream",
    "b200mpm_last_error",
    "b200mpm_pipeline_launch_count",
    "b200mpm_data_create",
    "b200mpm_data_destroy",
    "b200mpm_data_num_particles",
    "b200mpm_data_num_bodies",
    "b200mpm_step",
    "b200mpm_sync",
    "b200mpm_set_timestamps",
    "b200mpm_get_timings",
    "b200mpm_get_kernel_timings",
    "b200mpm_debug_timeline",
    "b200mpm_write_sim_params",
    "b200mpm_write_body_poses",
    "b200mpm_write_body_vels",
    "b200mpm_read_body_poses",
    "b200mpm_read_body_vels",
    "b200mpm_read_positions",
    "b200mpm_read_positions_async",
    "b200mpm_read_particles",
    "b200mpm_read_grid",
    "b200mpm_read_sorted_ids",
    "b200mpm_data_status",
    "b200mpm_data_set_rigid_particles",
    "b200mpm_prep_vertex_buffer",
    "b200mpm_data_reserve_grid",
    "b200mpm_data_set_auto_grow",
    "b200mpm_sort_only",
    "b200mpm_prefix_sum_u32",
    "b200mpm_slab_configure",
    "b200mpm_data_create_ex",
    "b200mpm_data_num_live",
    "b200mpm_shard_emigrate",
    "b200mpm_shard_immigrate",
    "b200mpm_shard_step_begin",
    "b200mpm_shard_halo_pack",
    "b200mpm_shard_halo_add",
    "b200mpm_shard_impulses",
    "b200mpm_shard_step_end",
    "b200mpm_read_particles_unordered",
    "b200mpm_read_positions_unordered",
    "b200mpm_read_positions_unordered_async",
    "b200mpm_nccl_unique_id",
    "b200mpm_shard_comm_init",
    "b200mpm_shard_step",
    "b200mpm_shard_p2p_export",
    "b200mpm_shard_p2p_connect",
)

PARTICLE_RECORD_BYTES = 128  # B200MPM_PARTICLE_RECORD_BYTES
HALO_BLOCK_BYTES = 1040  # B200MPM_HALO_BLOCK_BYTES
SHARD_HEADER_BYTES = 16  # B200MPM_SHARD_HEADER_BYTES


def load_library():
    """dlopen libb200mpm.so and declare the prototypes. Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise B200MpmError(
            ERR_NO_DEVICE,
            "libb200mpm.so is not built (run `python -m wgsparkl_b200.build`); there is no CPU fallback",
        )
    L = ctypes.CDLL(_LIB_PATH)
    vp, sz, u32, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_int
    L.b200mpm_last_error.restype = ctypes.c_char_p
    L.b200mpm_pipeline_create.argtypes = [i32, i32, ctypes.POINTER(vp)]
    L.b200mpm_pipeline_destroy.argtypes = [vp]
    L.b200mpm_pipeline_destroy.restype = None
    L.b200mpm_pipeline_set_stream.argtypes = [vp, vp]
    L.b200mpm_pipeline_launch_count.argtypes = [vp]
    L.b200mpm_pipeline_launch_count.restype = ctypes.c_uint64
    L.b200mpm_data_create.argtypes = [vp, vp, vp, sz, vp, sz, ctypes.c_float, u32, ctypes.POINTER(vp)]
    L.b200mpm_data_destroy.argtypes = [vp]
    L.b200mpm_data_destroy.restype = None
    L.b200mpm_data_num_particles.argtypes = [vp]
    L.b200mpm_data_num_particles.restype = sz
    L.b200mpm_data_num_bodies.argtypes = [vp]
    L.b200mpm_data_num_bodies.restype = sz
    L.b200mpm_step.argtypes = [vp, vp, u32]
    L.b200mpm_sync.argtypes = [vp]
    L.b200mpm_set_timestamps.argtypes = [vp, i32]
    L.b200mpm_get_timings.argtypes = [vp, vp]
    L.b200mpm_get_kernel_timings.argtypes = [vp, vp]
    L.b200mpm_debug_timeline.argtypes = [vp, vp]
    L.b200mpm_write_sim_params.argtypes = [vp, vp]
    for name in ("b200mpm_write_body_poses", "b200mpm_write_body_vels", "b200mpm_read_body_poses",
                 "b200mpm_read_body_vels"):
        getattr(L, name).argtypes = [vp, vp, sz]
    L.b200mpm_read_positions.argtypes = [vp, vp]
    L.b200mpm_data_reserve_grid.argtypes = [vp, ctypes.c_uint32]
    L.b200mpm_data_set_rigid_particles.argtypes = [vp, vp, vp, sz, vp, vp, sz]
    L.b200mpm_prep_vertex_buffer.argtypes = [vp, vp, vp, ctypes.c_uint32]
    L.b200mpm_data_set_auto_grow.argtypes = [vp, ctypes.c_float]
    L.b200mpm_read_positions_async.argtypes = [vp, vp]
    L.b200mpm_read_particles.argtypes = [vp, vp]
    L.b200mpm_read_grid.argtypes = [vp, vp, vp, sz, ctypes.POINTER(sz)]
    L.b200mpm_read_sorted_ids.argtypes = [vp, vp]
    L.b200mpm_data_status.argtypes = [vp, ctypes.POINTER(u32)]
    L.b200mpm_sort_only.argtypes = [vp, vp]
    L.b200mpm_prefix_sum_u32.argtypes = [vp, vp, sz]
    L.b200mpm_slab_configure.argtypes = [vp, ctypes.c_int32, ctypes.c_int32]
    L.b200mpm_data_create_ex.argtypes = [vp, vp, vp, sz, vp, sz, vp, sz, ctypes.c_float, u32, ctypes.POINTER(vp)]
    L.b200mpm_data_num_live.argtypes = [vp, ctypes.POINTER(sz)]
    L.b200mpm_shard_emigrate.argtypes = [vp, vp, vp, vp, u32]
    L.b200mpm_shard_immigrate.argtypes = [vp, vp, vp, u32]
    L.b200mpm_shard_step_begin.argtypes = [vp, vp]
    L.b200mpm_shard_halo_pack.argtypes = [vp, vp, vp, vp, u32]
    L.b200mpm_shard_halo_add.argtypes = [vp, vp, vp, u32]
    L.b200mpm_shard_impulses.argtypes = [vp, vp, vp, i32]
    L.b200mpm_shard_step_end.argtypes = [vp, vp]
    L.b200mpm_read_particles_unordered.argtypes = [vp, vp, vp, sz, ctypes.POINTER(sz)]
    L.b200mpm_read_positions_unordered.argtypes = [vp, vp, sz, ctypes.POINTER(sz)]
    L.b200mpm_read_positions_unordered_async.argtypes = [vp, vp, sz, ctypes.POINTER(sz)]
    L.b200mpm_nccl_unique_id.argtypes = [vp, sz]
    L.b200mpm_shard_comm_init.argtypes = [vp, vp, i32, i32, vp, u32, u32]
    L.b200mpm_shard_step.argtypes = [vp, vp, u32]
    L.b200mpm_shard_p2p_export.argtypes = [vp, vp, vp, sz]
    L.b200mpm_shard_p2p_connect.argtypes = [vp, vp, vp, sz]
    _lib = L
    return L


def _check(code):
    if code != OK:
        raise B200MpmError(code, load_library().b200mpm_last_error().decode("utf-8", "replace"))


def nccl_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    _check(load_library().b200mpm_nccl_unique_id(buf, 128))
    return buf.raw


class MpmPipeline:
    """MpmPipeline (src/pipeline.rs:24-39,176-281). One CUDA stream on one device."""

    def __init__(self, device: int = 0, dim: int = 3):
        L = load_library()
        self.dim = dim
        self.device = device
        h = ctypes.c_void_p()
        _check(L.b200mpm_pipeline_create(device, dim, ctypes.byref(h)))
        self._h = h

    # MpmPipeline::new(&Device) (pipeline.rs:176)
    @staticmethod
    def new(device: int = 0, dim: int = 3) -> "MpmPipeline":
        return MpmPipeline(device, dim)

    def close(self):
        if getattr(self, "_h", None):
            load_library().b200mpm_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def queue_step(self, data: "MpmData", num_substeps: int = 1, add_timestamps: Optional[bool] = None):
        """queue_step + `for _ in 0..num_substeps { queue.encode }` + submit
        (pipeline.rs:195-281, src_testbed/step.rs:122-128,169). Asynchronous."""
        if add_timestamps is not None:
            self.set_timestamps(add_timestamps)
        _check(load_library().b200mpm_step(self._h, data._h, int(num_substeps)))

    step = queue_step

    def prep_vertex_buffer(self, data: "MpmData", dev_instances: int, mode: int = abi.RENDER_DEFAULT):
        """WgPrepVertexBuffer::queue (src_testbed/prep_vertex_buffer.rs:81-113): fills the renderer's instance
        buffer - a DEVICE pointer to num_particles x abi.instance_dtype - from the device state. Asynchronous."""
        _check(load_library().b200mpm_prep_vertex_buffer(self._h, data._h, int(dev_instances), int(mode)))

    def sort_only(self, data: "MpmData"):
        """WgGrid::queue_sort alone (grid.rs:30-207), as in the gpu_grid_sort test (grid.rs:347-402)."""
        _check(load_library().b200mpm_sort_only(self._h, data._h))

    def sync(self):
        _check(load_library().b200mpm_sync(self._h))

    def set_stream(self, cuda_stream):
        """Submit to a caller-owned CUDA stream (an integer `cudaStream_t`, e.g.
        `torch.cuda.current_stream().cuda_stream`); None restores the pipeline's own stream."""
        _check(load_library().b200mpm_pipeline_set_stream(self._h, ctypes.c_void_p(cuda_stream or 0)))

    def set_timestamps(self, enabled: bool):
        _check(load_library().b200mpm_set_timestamps(self._h, 1 if enabled else 0))

    def timings_ms(self):
        """Accumulated milliseconds per reference pass name (src_testbed/lib.rs:133-146)."""
        out = np.zeros(len(abi.PASS_NAMES), dtype=np.float64)
        _check(load_library().b200mpm_get_timings(self._h, abi.ptr(out)))
        return dict(zip(abi.PASS_NAMES, out.tolist()))

    def kernel_timings_ms(self):
        """Accumulated milliseconds per kernel of this implementation (timestamps mode; abi.KERNEL_NAMES)."""
        out = np.zeros(len(abi.KERNEL_NAMES), dtype=np.float64)
        _check(load_library().b200mpm_get_kernel_timings(self._h, abi.ptr(out)))
        return dict(zip(abi.KERNEL_NAMES, out.tolist()))

    def launch_count(self) -> int:
        return int(load_library().b200mpm_pipeline_launch_count(self._h))

    def prefix_sum(self, v) -> np.ndarray:
        """WgPrefixSum::queue (prefix_sum.rs:20-69) on a host vector; returns the scanned copy."""
        out = np.ascontiguousarray(v, dtype=np.uint32).copy()
        _check(load_library().b200mpm_prefix_sum_u32(self._h, abi.ptr(out), out.size))
        return out


class MpmData:
    """MpmData (src/pipeline.rs:84-172): owns every device buffer of one simulation."""

    def __init__(self, pipeline: MpmPipeline, params: SimulationParams, particles, bodies=None,
                 cell_width: float = 1.0, grid_capacity: int = 60_000, particle_ids=None, particle_capacity=None):
        L = load_library()
        self.pipeline = pipeline
        self.dim = pipeline.dim
        if not isinstance(particles, np.ndarray):
            particles = particles_to_abi(particles, self.dim)
        particles = np.ascontiguousarray(particles, dtype=abi.particle_dtype)
        if bodies is None:
            bodies = np.zeros(0, dtype=abi.body_dtype)
        bodies = np.ascontiguousarray(bodies, dtype=abi.body_dtype)
        p = params.to_abi() if hasattr(params, "to_abi") else params
        self.num_particles = int(particles.shape[0])
        self.num_bodies = int(bodies.shape[0])
        self.capacity = 1
        while self.capacity < grid_capacity:
            self.capacity <<= 1
        h = ctypes.c_void_p()
        if particle_ids is None and particle_capacity is None:
            _check(L.b200mpm_data_create(pipeline._h, abi.ptr(p), abi.ptr(particles), self.num_particles,
                                         abi.ptr(bodies), self.num_bodies, ctypes.c_float(cell_width),
                                         int(grid_capacity), ctypes.byref(h)))
            self.particle_capacity = self.num_particles
        else:
            ids = None if particle_ids is None else np.ascontiguousarray(particle_ids, dtype=np.uint32)
            self.particle_capacity = int(particle_capacity or self.num_particles)
            _check(L.b200mpm_data_create_ex(pipeline._h, abi.ptr(p), abi.ptr(particles), self.num_particles, abi.ptr(ids),
                                            self.particle_capacity, abi.ptr(bodies), self.num_bodies,
                                            ctypes.c_float(cell_width), int(grid_capacity), ctypes.byref(h)))
        self._h = h

    @staticmethod
    def new(pipeline, params, particles, bodies, colliders, cell_width, grid_capacity):
        """MpmData::new(device, params, particles, &RigidBodySet, &ColliderSet, cell_width, capacity)
        (pipeline.rs:98-128): every collider with a parent body is coupled, TwoWays."""
        from .rapier import bodies_to_abi

        return MpmData(pipeline, params, particles, bodies_to_abi(bodies, colliders, pipeline.dim), cell_width,
                       grid_capacity)

    def close(self):
        if getattr(self, "_h", None):
            load_library().b200mpm_data_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- per-frame host writes / reads (src_testbed/step.rs:79-119,175-176; ui.rs:98-103)
    def write_sim_params(self, params):
        p = params.to_abi() if hasattr(params, "to_abi") else params
        _check(load_library().b200mpm_write_sim_params(self._h, abi.ptr(p)))

    def write_body_poses(self, poses):
        poses = np.ascontiguousarray(poses, dtype=abi.pose_dtype)
        _check(load_library().b200mpm_write_body_poses(self._h, abi.ptr(poses), poses.shape[0]))

    def write_body_vels(self, vels):
        vels = np.ascontiguousarray(vels, dtype=abi.velocity_dtype)
        _check(load_library().b200mpm_write_body_vels(self._h, abi.ptr(vels), vels.shape[0]))

    def read_body_poses(self):
        out = np.zeros(self.num_bodies, dtype=abi.pose_dtype)
        _check(load_library().b200mpm_read_body_poses(self._h, abi.ptr(out), self.num_bodies))
        return out

    def read_body_vels(self):
        out = np.zeros(self.num_bodies, dtype=abi.velocity_dtype)
        _check(load_library().b200mpm_read_body_vels(self._h, abi.ptr(out), self.num_bodies))
        return out

    # ---- state readback
    def read_positions(self, out: Optional[np.ndarray] = None):
        if out is None:
            out = np.zeros((self.num_particles, 4), dtype=np.float32)
        _check(load_library().b200mpm_read_positions(self._h, abi.ptr(out)))
        return out

    def read_positions_async(self, out: np.ndarray):
        """Asynchronous variant: `out` ((n, 4) float32, page-locked for a real overlap) is valid after
        `MpmPipeline.sync()`; the copy runs on its own stream while the next substeps execute."""
        if out.dtype != np.float32 or out.size < 4 * self.num_particles or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous float32 array with room for 4 * num_particles values")
        _check(load_library().b200mpm_read_positions_async(self._h, abi.ptr(out)))
        return out

    def read_particles(self):
        out = np.zeros(self.num_particles, dtype=abi.particle_dtype)
        _check(load_library().b200mpm_read_particles(self._h, abi.ptr(out)))
        return out

    def set_rigid_particles(self, vertices, vertex_colliders, samples, ids):
        """GpuRigidParticles::from_rapier: the sample points of the mesh colliders
        (wgsparkl_b200.rapier.rigid_particles_to_abi builds the four arrays). Once, before stepping."""
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        vc = np.ascontiguousarray(vertex_colliders, dtype=np.uint32)
        sp = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1, 3)
        ii = np.ascontiguousarray(ids, dtype=np.uint32).reshape(-1, 4)
        _check(load_library().b200mpm_data_set_rigid_particles(self._h, abi.ptr(v), abi.ptr(vc), len(v), abi.ptr(sp), abi.ptr(ii), len(sp)))

    def reserve_grid(self, grid_capacity: int):
        """Grow the block capacity (the reference's stubbed resize, grid.rs:43-118)."""
        _check(load_library().b200mpm_data_reserve_grid(self._h, int(grid_capacity)))

    def set_auto_grow(self, max_load: float):
        """Double the block capacity whenever more than `max_load` of it is active at a step call (0 = off)."""
        _check(load_library().b200mpm_data_set_auto_grow(self._h, float(max_load)))

    def status(self):
        """(num_active_blocks, overflowed)."""
        nb = ctypes.c_uint32(0)
        code = load_library().b200mpm_data_status(self._h, ctypes.byref(nb))
        if code not in (OK, ERR_GRID_OVERFLOW):
            _check(code)
        return int(nb.value), code == ERR_GRID_OVERFLOW

    def debug_timeline(self, enable=True):
        """{kernel: (first start, last end)} in GPU nanoseconds since the last call (None for kernels that did not
        run). The first call switches the recording on and returns nothing but None; enable=False switches it off."""
        if not enable:
            _check(load_library().b200mpm_debug_timeline(self._h, None))
            return {}
        raw = np.zeros(2 * len(abi.KERNEL_NAMES), dtype=np.uint64)
        _check(load_library().b200mpm_debug_timeline(self._h, abi.ptr(raw)))
        out = {}
        for k, name in enumerate(abi.KERNEL_NAMES):
            a, b = int(raw[2 * k]), int(raw[2 * k + 1])
            out[name] = None if b == 0 else (a, b)
        return out

    def read_grid(self):
        nb, _ = self.status()
        blocks = np.zeros(nb, dtype=abi.block_info_dtype)
        nodes = np.zeros(nb * 64, dtype=abi.node_dtype)
        got = ctypes.c_size_t(0)
        _check(load_library().b200mpm_read_grid(self._h, abi.ptr(blocks), abi.ptr(nodes), nb, ctypes.byref(got)))
        return blocks[: got.value], nodes.reshape(-1, 64)[: got.value]

    # ---- slab sharding (include/b200mpm.h "multi-GPU slab sharding"); buffers are raw device pointers (ints)
    def slab_configure(self, x_lo: int, x_hi: int):
        _check(load_library().b200mpm_slab_configure(self._h, int(x_lo), int(x_hi)))

    def num_live(self) -> int:
        n = ctypes.c_size_t(0)
        _check(load_library().b200mpm_data_num_live(self._h, ctypes.byref(n)))
        return int(n.value)

    def shard_emigrate(self, dev_left: int, dev_right: int, cap_records: int):
        _check(load_library().b200mpm_shard_emigrate(self.pipeline._h, self._h, ctypes.c_void_p(dev_left),
                                                     ctypes.c_void_p(dev_right), cap_records))

    def shard_immigrate(self, dev_buffer: int, cap_records: int):
        _check(load_library().b200mpm_shard_immigrate(self.pipeline._h, self._h, ctypes.c_void_p(dev_buffer), cap_records))

    def shard_step_begin(self):
        _check(load_library().b200mpm_shard_step_begin(self.pipeline._h, self._h))

    def shard_halo_pack(self, dev_left: int, dev_right: int, cap_blocks: int):
        _check(load_library().b200mpm_shard_halo_pack(self.pipeline._h, self._h, ctypes.c_void_p(dev_left),
                                                      ctypes.c_void_p(dev_right), cap_blocks))

    def shard_halo_add(self, dev_buffer: int, cap_blocks: int):
        _check(load_library().b200mpm_shard_halo_add(self.pipeline._h, self._h, ctypes.c_void_p(dev_buffer), cap_blocks))

    def shard_impulses(self, dev_buf: int, write: bool):
        _check(load_library().b200mpm_shard_impulses(self.pipeline._h, self._h, ctypes.c_void_p(dev_buf), 1 if write else 0))

    def shard_step_end(self):
        _check(load_library().b200mpm_shard_step_end(self.pipeline._h, self._h))

    def shard_comm_init(self, rank: int, world: int, unique_id: bytes, migration_cap: int, halo_cap: int):
        buf = ctypes.create_string_buffer(unique_id, 128)
        _check(load_library().b200mpm_shard_comm_init(self.pipeline._h, self._h, rank, world, buf, migration_cap, halo_cap))

    def shard_p2p_export(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        _check(load_library().b200mpm_shard_p2p_export(self.pipeline._h, self._h, buf, 64))
        return buf.raw

    def shard_p2p_connect(self, handles):
        blob = ctypes.create_string_buffer(b"".join(handles), 64 * len(handles))
        _check(load_library().b200mpm_shard_p2p_connect(self.pipeline._h, self._h, blob, len(handles)))

    def shard_step(self, num_substeps: int):
        """Whole sharded substeps (kernels + NCCL exchanges) from the native library, asynchronous."""
        _check(load_library().b200mpm_shard_step(self.pipeline._h, self._h, int(num_substeps)))

    def read_positions_unordered(self, out: Optional[np.ndarray] = None):
        """(n_live, 4) float32: xyz + id bits in w (abi.NONE for emigrated particles), device order."""
        if out is None:
            out = np.zeros((self.particle_capacity, 4), dtype=np.float32)
        n = ctypes.c_size_t(0)
        _check(load_library().b200mpm_read_positions_unordered(self._h, abi.ptr(out), out.shape[0], ctypes.byref(n)))
        return out[: n.value]

    def read_positions_unordered_async(self, out: np.ndarray) -> int:
        """Like read_positions_unordered, but without any host synchronisation: copies all `particle_capacity` slots
        (returns that number; `out` must hold as many rows, page-locked for a real overlap); rows whose id (w, as
        uint32) is NONE hold no particle. Valid after `MpmPipeline.sync()`."""
        n = ctypes.c_size_t(0)
        _check(load_library().b200mpm_read_positions_unordered_async(self._h, abi.ptr(out), out.shape[0], ctypes.byref(n)))
        return int(n.value)

    def read_particles_unordered(self):
        """(particles, ids) of the live particles in device order."""
        out = np.zeros(self.particle_capacity, dtype=abi.particle_dtype)
        ids = np.zeros(self.particle_capacity, dtype=np.uint32)
        n = ctypes.c_size_t(0)
        _check(load_library().b200mpm_read_particles_unordered(self._h, abi.ptr(out), abi.ptr(ids), self.particle_capacity,
                                                               ctypes.byref(n)))
        keep = ids[: n.value] != abi.NONE
        return out[: n.value][keep], ids[: n.value][keep]

    def read_sorted_ids(self):
        out = np.zeros(self.num_particles, dtype=np.uint32)
        _check(load_library().b200mpm_read_sorted_ids(self._h, abi.ptr(out)))
        return out
