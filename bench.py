#!/usr/bin/env python
"""bench.py — particle-substeps/s of the MPM substep hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n-side S]

A "step" is one testbed frame of the hot path = 20 substeps (sand3.rs:54 / elastic_cut3.rs:54) over
the synthetic scene named in `config.workload`. At N=1 that is BASELINE.json configs[1]: the 3D
corotated-elastic cube drop on a static ground cuboid, 1M particles. At N>1 it is configs[4], the 3D
Drucker-Prager sand dam break, slab-sharded along x with 2M particles per GPU (weak scaling; N=8 is the
16M-particle scene the north star names); `--workload cube` runs the stretched cube instead. Particle state is resident in
HBM when the timed region starts (it lives there in the reference too: src/pipeline.rs:130-168 uploads
once); `e2e` times the same frames through the C ABI with HOST buffers: per frame the body poses and
velocities are uploaded from host memory (src_testbed/step.rs:79-119) and the body poses plus all
particle positions are read back into host memory.

`--impl reference` times the CPU restatement of the reference (oracle/, OpenMP on all host cores) on a
bounded sample of the same workload; the reference itself (Rust + WGSL on wgpu) cannot be built in this
image (DESIGN.md "Oracle").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle-substeps/sec"
UNIT = "particle-substeps/s"

# Algorithmic bytes per particle-substep (BASELINE.md §3 / SURVEY §8d), 3D f32.
BYTES_P2G = 66.0
BYTES_G2P_ELASTIC = 162.0
BYTES_G2P_SAND = 218.0
# dram__bytes_read.sum + dram__bytes_write.sum of one k_g2p launch on the 1M-particle cube, from the committed
# `ncu --set full` capture (profiles/r01_ncu_full_1M_cube.md: 77.77 MB read + 59.39 MB written).
G2P_NCU_TRAFFIC_BYTES = 137.16e6
G2P_NCU_TRAFFIC_PARTICLES = 1_000_000


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


DAM_SLAB = (50, 200, 200)  # particles per GPU of the sand dam: 2M (x 8 GPUs = the 16M scene of configs[4])


def resolve_workload(args, world):
    if args.workload != "auto":
        return args.workload
    return "cube" if world == 1 else "dam"


def build_scene(args, world, workload):
    from wgsparkl_b200 import scenes

    if workload == "dam":
        # configs[4]: sand dam break inside a 4-wall box, 2M particles per GPU, slabs along x.
        return scenes.sand_dam_3d(DAM_SLAB[0] * world, DAM_SLAB[1], DAM_SLAB[2], grid_capacity=65_536)
    # configs[1]: 3D elastic cube drop, n_side^3 particles (stretched along x for N > 1). The cube starts just
    # above the ground cuboid so that the timed region covers the contact phase (CPIC active), the more
    # expensive regime.
    return scenes.elastic_cube_3d(args.n_side, y_offset=-5.0, grid_capacity=60_000, nx=args.n_side * world)


def workload_name(workload, world, n_total):
    if workload == "dam":
        return ("3D Drucker-Prager sand dam break in a 4-wall box (BASELINE configs[4]), %d particles = %d per GPU"
                % (n_total, n_total // world))
    return ("3D corotated-elastic cube drop on a static ground cuboid (BASELINE configs[1])"
            + ("" if world == 1 else ", stretched to %d x 1M particles along x" % world))


def frame_io_arrays(scene):
    from wgsparkl_b200 import abi

    nb = len(scene["bodies"])
    poses = np.zeros(nb, dtype=abi.pose_dtype)
    poses["translation"] = scene["bodies"]["translation"]
    poses["rotation"] = scene["bodies"]["rotation"]
    vels = np.zeros(nb, dtype=abi.velocity_dtype)
    vels["linear"] = scene["bodies"]["linvel"]
    vels["angular"] = scene["bodies"]["angvel"]
    return poses, vels


def run_ours(args):
    import torch
    import torch.distributed as dist

    from wgsparkl_b200.pipeline import MpmData, MpmPipeline

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    hbm_peak, peak_kind = load_peaks()

    # N = 1: BASELINE configs[1] (1M-particle elastic cube). N > 1: configs[4], the sand dam break at 2M
    # particles per GPU, slab-sharded over the N GPUs (weak scaling; migration + node halo every substep).
    workload = resolve_workload(args, world)
    scene = build_scene(args, world, workload)
    n_total = len(scene["particles"])
    spf = scene["substeps_per_frame"]
    stream = torch.cuda.Stream(device=local_rank)
    sharded = None
    if world == 1:
        pipe = MpmPipeline(local_rank, 3)  # raises without the CUDA library / an sm_100 device: no fallback
        pipe.set_stream(stream.cuda_stream)
        data = MpmData(pipe, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    else:
        from wgsparkl_b200.sharded import ShardedMpm

        sharded = ShardedMpm(scene, rank, world, local_rank, stream=stream)
        pipe, data = sharded.pipe, sharded.data
    poses, vels = frame_io_arrays(scene)
    n_local = data.num_particles
    host_pos = torch.empty((2, max(data.particle_capacity, 1), 4), dtype=torch.float32).pin_memory().numpy()
    e2e_slot = [0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            if finish is not None:
                finish()  # e.g. the last asynchronous readback: it must land inside the timed region
            e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def frame_device():
        if sharded is None:
            pipe.queue_step(data, spf)
        else:
            sharded.step(spf)

    def frame_e2e():
        data.write_body_poses(poses)  # H2D (src_testbed/step.rs:92-96)
        data.write_body_vels(vels)  # H2D (step.rs:98-119)
        frame_device()
        data.read_body_poses()  # D2H, blocking (step.rs:175-176)
        if sharded is None:
            # D2H of the step's result into pinned host memory on the copy stream: overlaps with the next frame,
            # like the reference's staging-buffer + map_async readbacks; completed by pipe.sync() (e2e_finish)
            data.read_positions_async(host_pos[e2e_slot[0]])
            e2e_slot[0] ^= 1
        else:
            data.read_positions_unordered_async(host_pos[e2e_slot[0]])  # D2H of this rank's slab, same pattern
            e2e_slot[0] ^= 1

    for _ in range(args.warmup):
        frame_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = pipe.launch_count()
    ms = timed(frame_device, args.steps)
    launches = pipe.launch_count() - l0
    value = n_total * spf * args.steps / (ms * 1e-3)

    # End-to-end through the C ABI with host buffers, over the SAME frames of the same trajectory as the
    # device-timed region (the cost of a frame depends on the state: the cube is being compressed): at N = 1 the
    # data object is rebuilt from the scene and warmed up again; a sharded run continues from where it is.
    if sharded is None:
        data.close()
        data = MpmData(pipe, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
        for _ in range(args.warmup):
            frame_e2e()
    else:
        frame_e2e()
    ms_e2e = timed(frame_e2e, args.steps, finish=pipe.sync)
    clocks = sampler.stop() if rank == 0 else None  # sampled (100 ms period) over both timed regions
    e2e_value = n_total * spf * args.steps / (ms_e2e * 1e-3)
    nb = len(scene["bodies"])
    h2d = nb * (poses.dtype.itemsize + vels.dtype.itemsize)
    d2h = nb * poses.dtype.itemsize + n_local * 16

    nblocks, overflow = data.status()
    # Per-kernel durations for the roofline: CUDA events around each pass, on the launching stream. At N > 1 this
    # runs AFTER both timed regions, on each rank's own slab with the exchanges off (plain substeps), rank 0 reports.
    pipe.set_timestamps(True)
    frames_prof = max(1, min(args.steps, 3))
    with torch.cuda.stream(stream):
        for _ in range(frames_prof):
            pipe.queue_step(data, spf)
    torch.cuda.synchronize()
    t = pipe.timings_ms()
    pipe.set_timestamps(False)
    launches_per_kernel = frames_prof * spf
    sand = workload == "dam"
    bytes_g2p = BYTES_G2P_SAND if sand else BYTES_G2P_ELASTIC
    n_roof = data.num_live() if sharded is not None else n_total
    g2p_ms = t["g2p"] / launches_per_kernel
    p2g_ms = t["p2g"] / launches_per_kernel
    g2p_gbs = bytes_g2p * n_roof / (g2p_ms * 1e-3) / 1e9
    p2g_gbs = BYTES_P2G * n_roof / (p2g_ms * 1e-3) / 1e9
    traffic = None
    if G2P_NCU_TRAFFIC_BYTES is not None and not sand and world == 1 and n_total == G2P_NCU_TRAFFIC_PARTICLES:
        traffic = G2P_NCU_TRAFFIC_BYTES
    roof = {"bound": "hbm", "kernel": "k_g2p (grid_update + g2p + particles_update)",
            "achieved": g2p_gbs, "peak": hbm_peak, "peak_source": peak_kind, "unit": "GB/s",
            "frac": g2p_gbs / hbm_peak, "traffic": traffic,
            "bytes_per_particle": bytes_g2p, "particles_per_launch": int(n_roof), "ms_per_launch": g2p_ms,
            "p2g": {"achieved": p2g_gbs, "frac": p2g_gbs / hbm_peak, "bytes_per_particle": BYTES_P2G,
                    "ms_per_launch": p2g_ms, "note": "both P2G instantiations, serialised (timestamps mode)"},
            "pass_ms_per_substep": {k: v / launches_per_kernel for k, v in t.items()}}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(workload, world, n_total),
                   "particles_total": n_total, "particles_per_gpu": n_total // world, "substeps_per_step": spf,
                   "cell_width": scene["cell_width"], "active_blocks_rank0": nblocks,
                   "parallelism": "1 GPU" if world == 1 else
                   "%d slabs along x (particle migration + node-halo exchange over NVLink peer stores, body "
                   "impulses all-reduced over NCCL, every substep)" % world,
                   "l2": "inputs larger than L2 (%.0f MB of particle state per GPU)" % (n_total // world * 220 / 1e6)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    out["roofline"] = roof
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, scene)
    if rank == 0:
        print(json.dumps(out))
    if sharded is not None:
        sharded.close()
    else:
        data.close()
        pipe.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, scene, budget_s=20.0):
    """Oracle (a port of the reference's kernels, OpenMP) on the host cores, bounded sample of the same scene."""
    from oracle import oracle

    cores = os.cpu_count() or 1
    oracle.set_threads(cores)
    n = len(scene["particles"])
    sim = oracle.OracleSim(3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    t0 = time.perf_counter()
    sim.step(1)
    t1 = time.perf_counter() - t0
    k = int(max(1, min(20, budget_s / max(t1, 1e-3))))
    t0 = time.perf_counter()
    sim.step(k)
    dt = time.perf_counter() - t0
    sim.close()
    return {"value": n * k / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d substeps of the full %d-particle scene after 1 warm-up substep" % (k, n)}


def run_reference(args):
    """The reference's own CPU-runnable form of the path: the oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle

    cores = os.cpu_count() or 1
    oracle.set_threads(cores)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = resolve_workload(args, world)
    # bounded sample of the arm's workload: the cube as is; of the sand dam, ONE GPU's share (2M particles) -
    # the metric is per particle-substep, and 16M particles would take minutes per substep on the host.
    scene = build_scene(args, 1, workload)
    n = len(scene["particles"])
    spf = scene["substeps_per_frame"]
    sim = oracle.OracleSim(3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    # bounded sample: each "step" advances `sample` substeps (instead of the full 20) so that the whole run
    # ends within a few minutes; the metric is per particle-substep, so it is directly comparable.
    t0 = time.perf_counter()
    sim.step(1)
    t1 = time.perf_counter() - t0
    total_budget = 120.0
    sample = int(max(1, min(spf, total_budget / max(t1, 1e-3) / max(1, args.steps + args.warmup))))
    for _ in range(args.warmup):
        sim.step(sample)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sim.step(sample)
    dt = time.perf_counter() - t0
    value = n * sample * args.steps / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(workload, world, n * world),
                   "particles_sampled": n, "substeps_per_step": sample,
                   "note": "CPU restatement of the reference's WGSL kernels (oracle/), OpenMP; the Rust/wgpu reference cannot be built here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d substeps per step x %d steps of the full %d-particle scene" % (sample, args.steps, n)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-side", type=int, default=100, help="cube side in particles (100 -> 1M particles)")
    ap.add_argument("--workload", default="auto", choices=["auto", "cube", "dam"],
                    help="auto: cube (configs[1]) at N=1, sand dam (configs[4], 2M particles per GPU) at N>1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # launched by hand without torchrun: re-launch as the driver does (one process per GPU, NCCL over 127.0.0.1)
        port = os.environ.get("MASTER_PORT", "29541")
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                                   "--master-addr", "127.0.0.1", "--master-port", port, os.path.abspath(__file__)] + sys.argv[1:])
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
