"""Multi-GPU MPM: 1-D slab decomposition of the sparse grid along x, one process per GPU.

The reference is single-device; this is the B200-native extension `north_star` asks for. Device work
(packing / unpacking, the substep phases) is in libb200mpm.so (csrc/shard.cu); this module owns the
TRANSPORT: `torch.distributed` (NCCL over NVLink / NVSwitch) moves fixed-size device buffers between slab
neighbours and sums the fixed-point body impulses. No host round trip happens inside a substep: record
counts travel in the buffers' device-side headers.

Per substep (see include/b200mpm.h "multi-GPU slab sharding"):
    emigrate -> neighbour exchange -> immigrate -> step_begin (sort .. P2G)
    halo_pack -> neighbour exchange -> halo_add -> [impulses all-reduce] -> step_end (G2P + update)
"""
from typing import List, Optional, Tuple

import numpy as np

from . import abi
from .pipeline import HALO_BLOCK_BYTES, PARTICLE_RECORD_BYTES, SHARD_HEADER_BYTES, MpmData, MpmPipeline

INT_MIN, INT_MAX = -(2**31), 2**31 - 1


def particle_block_x(positions_x: np.ndarray, cell_width: float, dim: int) -> np.ndarray:
    """Block x-index of every particle, bit-identical to block_associated_to_point (grid.wgsl:269-292):
    floor((round(x / h) - 1) / BLOCK) with round = ties-to-even in float32."""
    block = 8 if dim == 2 else 4
    c = np.rint(positions_x.astype(np.float32) / np.float32(cell_width)) - np.float32(1.0)
    return np.floor(c / np.float32(block)).astype(np.int64)


def partition_slabs(block_x: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Cuts the block-x axis into `world` contiguous slabs holding ~equal particle counts. Returns [lo, hi)
    per rank; the first slab is open towards -x and the last towards +x."""
    if world == 1:
        return [(INT_MIN, INT_MAX)]
    lo_x, hi_x = int(block_x.min()), int(block_x.max())
    hist = np.bincount((block_x - lo_x).astype(np.int64), minlength=hi_x - lo_x + 1)
    cum = np.cumsum(hist)
    total = int(cum[-1])
    cuts = []
    prev = lo_x
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side="left")) + 1 + lo_x  # first column of the next slab
        k = max(k, prev + 1)  # every slab owns at least one block column
        k = min(k, hi_x + 1)
        cuts.append(k)
        prev = k
    bounds = [INT_MIN] + cuts + [INT_MAX]
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def neighbours(rank: int, world: int) -> Tuple[Optional[int], Optional[int]]:
    return (rank - 1 if rank > 0 else None, rank + 1 if rank < world - 1 else None)


def exchange_with_neighbours(dist, send_left, send_right, recv_left, recv_right, rank: int, world: int):
    """One batched neighbour exchange (a single NCCL group): send_left -> rank-1, send_right -> rank+1,
    recv_left <- rank-1, recv_right <- rank+1. Works with the gloo backend on CPU tensors too (tests)."""
    left, right = neighbours(rank, world)
    ops = []
    if left is not None:
        ops.append(dist.P2POp(dist.isend, send_left, left))
        ops.append(dist.P2POp(dist.irecv, recv_left, left))
    if right is not None:
        ops.append(dist.P2POp(dist.isend, send_right, right))
        ops.append(dist.P2POp(dist.irecv, recv_right, right))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class ShardedMpm:
    """One rank of a slab-sharded simulation. Construct on every rank with the SAME full scene (synthetic scenes
    are cheap to rebuild); each rank keeps the particles of its slab."""

    def __init__(self, scene: dict, rank: int, world: int, device: int, migration_cap: int = 16384,
                 halo_cap: Optional[int] = None, slack: float = 1.6, slabs: Optional[List[Tuple[int, int]]] = None,
                 stream=None, native: bool = True, p2p: bool = True):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank, self.world, self.dim = rank, world, scene["dim"]
        parts = scene["particles"]
        bx = particle_block_x(parts["position"][:, 0], scene["cell_width"], self.dim)
        self.slabs = slabs or partition_slabs(bx, world)
        lo, hi = self.slabs[rank]
        mine = (bx >= lo) & (bx < hi)
        ids = np.nonzero(mine)[0].astype(np.uint32)
        n_mine = int(mine.sum())
        self.pipe = MpmPipeline(device, self.dim)
        # One explicit torch stream carries the kernels of this rank AND its NCCL exchanges (the legacy default
        # stream has handle 0, which the C ABI reads as "use the pipeline's own stream").
        self.stream = stream if stream is not None else torch.cuda.Stream(device=device)
        self.pipe.set_stream(self.stream.cuda_stream)
        cap = int(n_mine * slack) + 4 * migration_cap
        self.data = MpmData(self.pipe, scene["params"], parts[mine], scene["bodies"], scene["cell_width"], scene["grid_capacity"],
                            particle_ids=ids, particle_capacity=cap)
        if scene.get("rigid_particles") is not None:
            # mesh colliders are replicated: every slab holds all sample points and applies them to its own blocks
            self.data.set_rigid_particles(*scene["rigid_particles"])
        if world > 1:
            self.data.slab_configure(lo, hi)
        self.n_global = len(parts)
        self.substeps_per_frame = scene.get("substeps_per_frame", 1)
        # Exchange buffers (device memory owned by torch; raw pointers cross the C ABI).
        if halo_cap is None:
            # blocks of one block column: bounded by the y-z extent of the particles (+ neighbours), with slack
            block = 8 if self.dim == 2 else 4
            ext = (parts["position"].max(0) - parts["position"].min(0)) / scene["cell_width"] / block + 3
            halo_cap = int(4 * ext[1] * (ext[2] if self.dim == 3 else 1)) + 64
        self.migration_cap, self.halo_cap = migration_cap, halo_cap
        mig_bytes = SHARD_HEADER_BYTES + migration_cap * PARTICLE_RECORD_BYTES
        halo_bytes = SHARD_HEADER_BYTES + halo_cap * HALO_BLOCK_BYTES
        mk = lambda nbytes: torch.zeros(nbytes, dtype=torch.uint8, device="cuda:%d" % device)
        self.mig_send = [mk(mig_bytes), mk(mig_bytes)]
        self.mig_recv = [mk(mig_bytes), mk(mig_bytes)]
        self.halo_send = [mk(halo_bytes), mk(halo_bytes)]
        self.halo_recv = [mk(halo_bytes), mk(halo_bytes)]
        self.impulses = torch.zeros(abi.MAX_BODIES * 6, dtype=torch.int32, device="cuda:%d" % device)
        b = scene["bodies"]
        # the impulse all-reduce is only needed if some body can react to an impulse
        self.needs_impulses = bool(len(b)) and bool(
            np.any(b["inv_mass"] != 0) or np.any(b["inv_inertia"] != 0) or np.any(b["linvel"] != 0) or np.any(b["angvel"] != 0))
        # Native transport: the library owns an NCCL communicator and replays whole sharded substeps (kernels +
        # send/recv groups + impulse all-reduce) from one CUDA graph; torch.distributed only ships the unique id.
        self.native = bool(native and world > 1 and dist.is_available() and dist.is_initialized())
        if self.native:
            from .pipeline import nccl_unique_id

            box = [nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            self.data.shard_comm_init(rank, world, box[0], migration_cap, halo_cap)
            self.p2p = False
            if p2p:
                # NVLink peer-to-peer stores instead of NCCL send/recv: exchange CUDA-IPC handles of the arenas.
                # (a rank that cannot export - no peer access, IPC unavailable - makes EVERY rank fall back to the
                # NCCL send/recv transport before anybody has mapped anything)
                try:
                    mine = self.data.shard_p2p_export()
                except Exception:
                    mine = None
                handles = [None] * world
                dist.all_gather_object(handles, mine)
                ok = 0
                if all(hd is not None for hd in handles):
                    try:
                        self.data.shard_p2p_connect(handles)
                        ok = 1
                    except Exception:
                        ok = 0
                flag = torch.tensor([ok], device="cuda:%d" % device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # all ranks or none
                self.p2p = bool(flag.item())
                if not self.p2p and ok:
                    raise RuntimeError("peer-to-peer mapping succeeded on some ranks only")

    def substep(self):
        with self.torch.cuda.stream(self.stream):
            self._substep()

    def _substep(self):
        d, dist = self.data, self.dist
        left, right = neighbours(self.rank, self.world)
        if self.world > 1:
            d.shard_emigrate(self.mig_send[0].data_ptr(), self.mig_send[1].data_ptr(), self.migration_cap)
            exchange_with_neighbours(dist, self.mig_send[0], self.mig_send[1], self.mig_recv[0], self.mig_recv[1],
                                     self.rank, self.world)
            if left is not None:
                d.shard_immigrate(self.mig_recv[0].data_ptr(), self.migration_cap)
            if right is not None:
                d.shard_immigrate(self.mig_recv[1].data_ptr(), self.migration_cap)
        d.shard_step_begin()
        if self.world > 1:
            d.shard_halo_pack(self.halo_send[0].data_ptr(), self.halo_send[1].data_ptr(), self.halo_cap)
            exchange_with_neighbours(dist, self.halo_send[0], self.halo_send[1], self.halo_recv[0], self.halo_recv[1],
                                     self.rank, self.world)
            if left is not None:
                d.shard_halo_add(self.halo_recv[0].data_ptr(), self.halo_cap)
            if right is not None:
                d.shard_halo_add(self.halo_recv[1].data_ptr(), self.halo_cap)
            if self.needs_impulses:
                d.shard_impulses(self.impulses.data_ptr(), False)
                dist.all_reduce(self.impulses, op=dist.ReduceOp.SUM)  # exact: fixed-point integers
                d.shard_impulses(self.impulses.data_ptr(), True)
        d.shard_step_end()

    def step(self, substeps: int):
        if self.native:
            self.data.shard_step(substeps)
            return
        for _ in range(substeps):
            self.substep()

    def sync(self):
        self.torch.cuda.synchronize()

    def num_live(self) -> int:
        if getattr(self, "p2p", False):
            # the peer-to-peer path drops the emigrated tail lazily (at the next substep): count the live ids
            return int(len(self.data.read_particles_unordered()[1]))
        return self.data.num_live()

    def gather_particles(self):
        """All particles of the simulation in their ORIGINAL order, on rank 0 (None elsewhere)."""
        torch, dist = self.torch, self.dist
        parts, ids = self.data.read_particles_unordered()
        if self.world == 1:
            out = np.zeros(self.n_global, dtype=abi.particle_dtype)
            out[ids] = parts
            return out
        payload = (parts, ids)
        gathered = [None] * self.world if self.rank == 0 else None
        dist.gather_object(payload, gathered, dst=0)
        if self.rank != 0:
            return None
        out = np.zeros(self.n_global, dtype=abi.particle_dtype)
        seen = np.zeros(self.n_global, dtype=np.int32)
        for p, i in gathered:
            out[i] = p
            seen[i] += 1
        assert np.all(seen == 1), "particles lost or duplicated by migration: %d missing, %d duplicated" % (
            int((seen == 0).sum()), int((seen > 1).sum()))
        return out

    def close(self):
        self.data.close()
        self.pipe.close()


class LocalSlabs:
    """All slabs of a sharded run inside ONE process on ONE GPU: the same device code path as `ShardedMpm`
    (emigrate / immigrate / halo pack / halo add / phased substep), with device-to-device copies in place of the
    NCCL exchange. Used to validate the sharding logic against the unsharded run on a single-GPU box."""

    def __init__(self, scene: dict, world: int, device: int = 0, **kw):
        import torch

        self.torch = torch
        self.world = world
        bx = particle_block_x(scene["particles"]["position"][:, 0], scene["cell_width"], scene["dim"])
        slabs = kw.pop("slabs", None) or partition_slabs(bx, world)
        self.stream = torch.cuda.Stream(device=device)  # all slabs share one stream: program order = data order
        self.ranks = [ShardedMpm(scene, r, world, device, slabs=slabs, stream=self.stream, native=False, **kw)
                      for r in range(world)]
        self.n_global = len(scene["particles"])

    def substep(self):
        with self.torch.cuda.stream(self.stream):
            self._substep()

    def _substep(self):
        R, W = self.ranks, self.world
        for r in R:
            r.data.shard_emigrate(r.mig_send[0].data_ptr(), r.mig_send[1].data_ptr(), r.migration_cap)
        for i, r in enumerate(R):
            if i > 0:
                R[i - 1].mig_recv[1].copy_(r.mig_send[0])
            if i < W - 1:
                R[i + 1].mig_recv[0].copy_(r.mig_send[1])
        for i, r in enumerate(R):
            if i > 0:
                r.data.shard_immigrate(r.mig_recv[0].data_ptr(), r.migration_cap)
            if i < W - 1:
                r.data.shard_immigrate(r.mig_recv[1].data_ptr(), r.migration_cap)
        for r in R:
            r.data.shard_step_begin()
        for r in R:
            r.data.shard_halo_pack(r.halo_send[0].data_ptr(), r.halo_send[1].data_ptr(), r.halo_cap)
        for i, r in enumerate(R):
            if i > 0:
                R[i - 1].halo_recv[1].copy_(r.halo_send[0])
            if i < W - 1:
                R[i + 1].halo_recv[0].copy_(r.halo_send[1])
        for i, r in enumerate(R):
            if i > 0:
                r.data.shard_halo_add(r.halo_recv[0].data_ptr(), r.halo_cap)
            if i < W - 1:
                r.data.shard_halo_add(r.halo_recv[1].data_ptr(), r.halo_cap)
        if R[0].needs_impulses and W > 1:
            for r in R:
                r.data.shard_impulses(r.impulses.data_ptr(), False)
            total = self.torch.stack([r.impulses for r in R]).sum(0, dtype=self.torch.int32)
            for r in R:
                r.impulses.copy_(total)
                r.data.shard_impulses(r.impulses.data_ptr(), True)
        for r in R:
            r.data.shard_step_end()

    def step(self, n: int):
        for _ in range(n):
            self.substep()

    def gather_particles(self):
        self.torch.cuda.synchronize()
        out = np.zeros(self.n_global, dtype=abi.particle_dtype)
        seen = np.zeros(self.n_global, dtype=np.int32)
        for r in self.ranks:
            p, i = r.data.read_particles_unordered()
            out[i] = p
            seen[i] += 1
        assert np.all(seen == 1), "particles lost or duplicated by migration: %d missing, %d duplicated" % (
            int((seen == 0).sum()), int((seen > 1).sum()))
        return out

    def live_counts(self):
        return [r.num_live() for r in self.ranks]

    def close(self):
        for r in self.ranks:
            r.close()
