#!/usr/bin/env python
"""bench.py — particle-substeps/s of the MPM substep hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload auto|cube|dam|dam-strong|column|mixed] [--no-extras] [--no-cpu-baseline]

A "step" is one testbed frame of the hot path = 20 substeps (sand3.rs:54 / elastic_cut3.rs:54) over the synthetic
scene named in `config.workload`:
  cube        BASELINE configs[1]: 3D corotated-elastic cube drop on a static ground cuboid, 1M particles (N = 1 default)
  column      configs[2]: 3D Drucker-Prager sand column collapse, 4M particles, run for >= 2000 substeps; the active
              block count B(t) and the overflow flag are reported (sparse-grid block activation churn)
  mixed       configs[3]: 4M sand + Neo-Hookean solids with a kinematic rotating cuboid and dynamic cuboids, two-way coupled
  dam         configs[4], WEAK scaling: sand dam break, 2M particles per GPU, slab-sharded along x (N > 1 default)
  dam-strong  configs[4], STRONG scaling: the fixed 16M-particle dam over N GPUs (N = 1: one GPU holds all of it)
With no `--workload` the line of the default workload also carries, measured in the same run, `also.*`: at N = 1 the
column, the mixed scene and one 2M dam slab alone (the like-for-like base of the weak-scaling lines); at N > 1
`scaling_base` (one dam slab alone on rank 0's GPU) and `also.dam_strong` (the 16M dam over the same N GPUs).

Particle state is resident in HBM when a timed region starts (it lives there in the reference too: src/pipeline.rs:
130-168 uploads once). K steps are timed five times (at N = 1 every region restarts the scene, so all five cover the same frames);
`value` is the MEDIAN region. `e2e` times the same frames through the C ABI with HOST buffers: per frame the body poses and
velocities are uploaded from host memory (src_testbed/step.rs:79-119) and the body poses plus all particle positions
are read back into pinned host memory.

`--impl reference` times the CPU restatement of the reference (oracle/, OpenMP on all host cores) on a bounded
sample of the same workload; the reference itself (Rust + WGSL on wgpu) cannot be built in this image nor on the GPU
box (no cargo / rustc / Vulkan loader: DESIGN.md "Oracle").
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

# Algorithmic bytes per particle-substep (BASELINE.md §3 / SURVEY §8d / DESIGN.md §4), 3D f32.
BYTES_P2G = 66.0
BYTES_G2P_ELASTIC = 162.0
BYTES_G2P_SAND = 218.0
# DRAM traffic per launch of the dominant kernels from the committed `ncu --set full` captures: written by
# `tools/ncu_summary.py --json` (never typed in by hand), keyed by workload and kernel.
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def load_traffic(workload, kernel, particles):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch (bytes), or None if no capture matches."""
    try:
        with open(TRAFFIC_FILE) as f:
            table = json.load(f)
        e = table[workload][kernel]
        if int(e["particles"]) != int(particles):
            return None
        return float(e["dram_bytes"])
    except Exception:
        return None


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


DAM_SLAB = (50, 200, 200)  # particles per GPU of the weak-scaling sand dam: 2M (x 8 GPUs = the 16M scene of configs[4])
DAM_FULL = (400, 200, 200)  # the 16M-particle dam (strong scaling)
CUBE_Y_OFFSET = -5.0  # SURVEY §8d drops the cube from +60; here it starts in contact so that the timed frames cover the
                      # compression phase (CPIC + large-strain stress path active) - the more expensive regime


def resolve_workload(args, world):
    if args.workload != "auto":
        return args.workload
    return "cube" if world == 1 else "dam"


def build_scene(args, world, workload):
    from wgsparkl_b200 import scenes

    if workload == "dam":
        return scenes.sand_dam_3d(DAM_SLAB[0] * world, DAM_SLAB[1], DAM_SLAB[2], grid_capacity=65_536)
    if workload == "dam-strong":
        return scenes.sand_dam_3d(*DAM_FULL, grid_capacity=524_288 if world == 1 else 262_144)
    if workload == "column":
        return scenes.sand_column_3d(100, 400, 100, grid_capacity=131_072)
    if workload == "mixed":
        return scenes.mixed_coupled_3d(160, 160, 160, grid_capacity=131_072, n_dynamic=4)
    return scenes.elastic_cube_3d(args.n_side, y_offset=CUBE_Y_OFFSET, grid_capacity=60_000, nx=args.n_side * world)


def workload_name(workload, world, n_total):
    if workload == "dam":
        return ("3D Drucker-Prager sand dam break in a 4-wall box (BASELINE configs[4], weak scaling), %d particles = %d per GPU"
                % (n_total, n_total // world))
    if workload == "dam-strong":
        return "3D Drucker-Prager sand dam break in a 4-wall box (BASELINE configs[4], strong scaling), %d particles over %d GPU(s)" % (n_total, world)
    if workload == "column":
        return "3D Drucker-Prager sand column collapse (BASELINE configs[2]), %d particles" % n_total
    if workload == "mixed":
        return ("3D mixed sand + Neo-Hookean solids with a kinematic rotating cuboid and dynamic cuboids, two-way coupled "
                "(BASELINE configs[3]), %d particles" % n_total)
    return ("3D corotated-elastic cube drop on a static ground cuboid (BASELINE configs[1]); the cube starts in contact "
            "(y offset %.0f instead of SURVEY 8d's +60) so that the timed frames cover the compression phase" % CUBE_Y_OFFSET
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


class Runner:
    """One workload on this rank's GPU (world == 1: plain MpmData; world > 1: a slab of a ShardedMpm)."""

    def __init__(self, scene, rank, world, local_rank, stream, shard):
        import torch
        from wgsparkl_b200.pipeline import MpmData, MpmPipeline

        self.torch, self.scene, self.world, self.stream = torch, scene, world, stream
        self.spf = scene["substeps_per_frame"]
        self.n_total = len(scene["particles"])
        self.sharded = None
        if not shard:
            self.pipe = MpmPipeline(local_rank, 3)  # raises without the CUDA library / an sm_100 device: no fallback
            self.pipe.set_stream(stream.cuda_stream)
            self.data = MpmData(self.pipe, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"],
                                scene["grid_capacity"])
        else:
            from wgsparkl_b200.sharded import ShardedMpm

            self.sharded = ShardedMpm(scene, rank, world, local_rank, stream=stream)
            self.pipe, self.data = self.sharded.pipe, self.sharded.data

    def rebuild(self):
        from wgsparkl_b200.pipeline import MpmData

        assert self.sharded is None
        s = self.scene
        self.data.close()
        self.data = MpmData(self.pipe, s["params"], s["particles"], s["bodies"], s["cell_width"], s["grid_capacity"])

    def frame(self):
        if self.sharded is None:
            self.pipe.queue_step(self.data, self.spf)
        else:
            self.sharded.step(self.spf)

    def close(self):
        if self.sharded is not None:
            self.sharded.close()
        else:
            self.data.close()
            self.pipe.close()


def make_timer(torch, dist, world, stream):
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        """K steps bracketed by barrier + synchronize, CUDA events on the launching stream, MAX over ranks (ms)."""
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

    return timed, None


def kernel_roofline(runner, workload, hbm_peak, peak_kind, frames):
    """Per-kernel durations from CUDA events around every kernel (timestamps mode: plain launches on the launching
    stream instead of the graph replay), and the HBM roofline of P2G and G2P from their ALGORITHMIC bytes."""
    torch, pipe, data = runner.torch, runner.pipe, runner.data
    spf = runner.spf
    pipe.set_timestamps(True)
    with torch.cuda.stream(runner.stream):
        for _ in range(frames):
            pipe.queue_step(data, spf)
    torch.cuda.synchronize()
    passes = pipe.timings_ms()
    kernels = pipe.kernel_timings_ms()
    pipe.set_timestamps(False)
    launches = frames * spf
    k_ms = {k: v / launches for k, v in kernels.items()}
    # The same kernels INSIDE the graph replay (where events cannot look): %globaltimer stamps taken by thread 0 of
    # every CTA (first start .. last end per kernel and substep), read back after every single-substep replay.
    g_ms = {}
    try:
        data.debug_timeline()  # switches the recording on (re-captures the graphs)
        with torch.cuda.stream(runner.stream):
            pipe.queue_step(data, 2)
        torch.cuda.synchronize()
        data.debug_timeline()
        acc, reps = {}, 2 * spf
        for _ in range(reps):
            with torch.cuda.stream(runner.stream):
                pipe.queue_step(data, 1)
            for k, se in data.debug_timeline().items():
                if se is not None:
                    acc[k] = acc.get(k, 0.0) + (se[1] - se[0]) * 1e-6
        g_ms = {k: v / reps for k, v in acc.items()}
        data.debug_timeline(enable=False)
    except Exception as e:  # (never at the cost of the main line)
        g_ms = {"error": repr(e)}
    sand = workload in ("dam", "dam-strong", "column", "mixed")
    bytes_g2p = BYTES_G2P_SAND if sand else BYTES_G2P_ELASTIC
    n_roof = data.num_live() if runner.sharded is not None else runner.n_total

    def entry(kernel, name, bytes_pp, ms, key):
        gbs = bytes_pp * n_roof / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        e = {"kernel": name, "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
             "bytes_per_particle": bytes_pp, "particles_per_launch": int(n_roof), "ms_per_launch": ms,
             "timing": "CUDA events around each launch (plain launches on the launching stream; includes ~3-5 us of "
                       "launch / event overhead per kernel)",
             "traffic": load_traffic(workload, kernel, n_roof)}
        gms = g_ms.get(key) if isinstance(g_ms.get(key), float) else None
        if gms:
            ggbs = bytes_pp * n_roof / (gms * 1e-3) / 1e9
            e["in_graph"] = {"ms_per_launch": gms, "achieved": ggbs, "frac": ggbs / hbm_peak,
                             "timing": "%globaltimer, first CTA start .. last CTA end inside the graph replay"}
        return e

    g2p = entry("k_g2p", "k_g2p (grid_update + g2p + particles_update + reset_hmap)", bytes_g2p, k_ms["g2p"], "g2p")
    p2g = entry("k_p2g", "k_p2g (particle colouring + collider-side blocks + all other blocks: one kernel)",
                BYTES_P2G, k_ms["p2g"] + k_ms["p2g_cpic"], "p2g")
    lead, other = (g2p, p2g) if g2p["ms_per_launch"] >= p2g["ms_per_launch"] else (p2g, g2p)
    roof = dict(lead)
    roof.update({"bound": "hbm", "peak_source": peak_kind,
                 "other": other, "substep_bytes_per_particle": 306.0 if sand else 250.0,
                 "kernel_ms_per_substep": k_ms, "kernel_ms_per_substep_in_graph": g_ms,
                 "pass_ms_per_substep": {k: v / launches for k, v in passes.items()}})
    return roof


def run_extra(args, name, workload, local_rank, stream, torch, frames_warm, frames_timed, report_blocks=False):
    """A further BASELINE configuration on one GPU, measured in the same run (device-timed, median of 3 regions)."""
    scene = build_scene(args, 1, workload)
    runner = Runner(scene, 0, 1, local_rank, stream, shard=False)
    timed, _ = make_timer(torch, None, 1, stream)
    out = {"workload": workload_name(workload, 1, runner.n_total), "particles": runner.n_total}
    for _ in range(frames_warm):
        runner.frame()
    spf = runner.spf
    if report_blocks:
        # configs[2]: >= 2000 substeps of the collapse, B(t) sampled every 10 frames (a 4-byte readback between regions)
        series, total_ms, chunk = [], 0.0, 10
        nb0, _ = runner.data.status()
        series.append([0, int(nb0)])
        done = 0
        while done < frames_timed:
            total_ms += timed(runner.frame, chunk)
            done += chunk
            nb, overflow = runner.data.status()
            series.append([done * spf, int(nb)])
        out.update({"substeps": done * spf, "value": runner.n_total * spf * done / (total_ms * 1e-3), "unit": UNIT,
                    "us_per_substep": total_ms * 1e3 / (done * spf), "active_blocks_over_substeps": series,
                    "overflow": bool(overflow)})
    else:
        samples = [timed(runner.frame, frames_timed) for _ in range(3)]
        ms = float(np.median(samples))
        nb, overflow = runner.data.status()
        out.update({"substeps": frames_timed * spf, "value": runner.n_total * spf * frames_timed / (ms * 1e-3), "unit": UNIT,
                    "us_per_substep": ms * 1e3 / (frames_timed * spf), "active_blocks": int(nb), "overflow": bool(overflow)})
    pos = runner.data.read_positions() if runner.n_total <= 5_000_000 else None
    if pos is not None:
        out["finite"] = bool(np.isfinite(pos).all())
    runner.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    hbm_peak, peak_kind = load_peaks()

    workload = resolve_workload(args, world)
    scene = build_scene(args, world, workload)
    stream = torch.cuda.Stream(device=local_rank)
    runner = Runner(scene, rank, world, local_rank, stream, shard=world > 1)
    pipe, n_total, spf = runner.pipe, runner.n_total, runner.spf
    poses, vels = frame_io_arrays(scene)
    n_local = runner.data.num_particles
    host_pos = torch.empty((2, max(runner.data.particle_capacity, 1), 4), dtype=torch.float32).pin_memory().numpy()
    e2e_slot = [0]
    timed, timed_median = make_timer(torch, dist, world, stream)

    def frame_e2e():
        data = runner.data
        data.write_body_poses(poses)  # H2D (src_testbed/step.rs:92-96)
        data.write_body_vels(vels)  # H2D (step.rs:98-119)
        runner.frame()
        data.read_body_poses()  # D2H, blocking (step.rs:175-176)
        # D2H of the step's result into pinned host memory on the copy stream: overlaps with the next frame, like
        # the reference's staging-buffer + map_async readbacks; completed by pipe.sync() (finish)
        if runner.sharded is None:
            data.read_positions_async(host_pos[e2e_slot[0]])
        else:
            data.read_positions_unordered_async(host_pos[e2e_slot[0]])  # this rank's slab, same pattern
        e2e_slot[0] ^= 1

    # The cost of a frame depends on the state (the cube's substep costs 2.4x more at peak compression than in the first
    # frames of contact), so at N = 1 every timed region covers the SAME frames of the same trajectory: the data object
    # is rebuilt from the scene, W warm-up frames, then the K timed frames (frames W .. W+K, as in round 1) - five
    # times, median. A sharded run (expensive to rebuild, and the dam barely changes within 5 K frames) continues.
    REPEATS = 5
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    samples, samples_e2e, launches = [], [], 0
    for r in range(REPEATS):
        if runner.sharded is None and r > 0:
            runner.rebuild()
        if r == 0 or runner.sharded is None:
            for _ in range(args.warmup):
                runner.frame()
        l0 = pipe.launch_count()
        samples.append(timed(runner.frame, args.steps))
        launches = pipe.launch_count() - l0
    ms = float(np.median(samples))
    value = n_total * spf * args.steps / (ms * 1e-3)

    # End-to-end through the C ABI with host buffers, over the same frames (same restart rule).
    for r in range(REPEATS):
        if runner.sharded is None:
            runner.rebuild()
        if r == 0 or runner.sharded is None:
            for _ in range(args.warmup if runner.sharded is None else 1):
                frame_e2e()
        samples_e2e.append(timed(frame_e2e, args.steps, finish=pipe.sync))
    ms_e2e = float(np.median(samples_e2e))
    clocks = sampler.stop() if rank == 0 else None  # sampled (100 ms period) over all timed regions
    e2e_value = n_total * spf * args.steps / (ms_e2e * 1e-3)
    nb = len(scene["bodies"])
    h2d = nb * (poses.dtype.itemsize + vels.dtype.itemsize)
    # (a slab copies all its particle slots, spare capacity included: a fixed-size copy needs no host synchronisation)
    d2h = nb * poses.dtype.itemsize + (n_local if runner.sharded is None else runner.data.particle_capacity) * 16
    nblocks, overflow = runner.data.status()

    # Per-kernel durations for the roofline. At N > 1 this runs after both timed regions, on each rank's own slab with
    # the exchanges off (plain substeps); rank 0 reports.
    roof = kernel_roofline(runner, workload, hbm_peak, peak_kind, max(1, min(args.steps, 3)))

    strong = workload == "dam-strong"
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(workload, world, n_total),
                   "particles_total": n_total, "particles_per_gpu": n_total // world, "substeps_per_step": spf,
                   "cell_width": scene["cell_width"], "active_blocks_rank0": nblocks, "grid_overflow": bool(overflow),
                   "parallelism": "1 GPU" if world == 1 else
                   "%d slabs along x (particle migration + node-halo exchange over NVLink peer stores, body "
                   "impulses all-reduced over NCCL, every substep)" % world,
                   "timing": "median of %d regions of %d steps each (CUDA events on the launching stream, max over ranks)%s"
                             % (len(samples), args.steps, "; every region restarts the scene and times frames %d..%d"
                                % (args.warmup, args.warmup + args.steps) if world == 1 else ""),
                   "timed_regions_ms": samples, "timed_regions_e2e_ms": samples_e2e,
                   "l2": "inputs larger than L2 (%.0f MB of particle state per GPU)" % (n_total // world * 220 / 1e6)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
    }
    runner.close()
    runner = None

    extras = args.workload == "auto" and not args.no_extras
    if extras and world == 1:
        also = {}
        try:  # cost profile of the cube along its trajectory (drop, compression, rebound): ms per 10 frames
            tr = Runner(scene, 0, 1, local_rank, stream, shard=False)
            prof = [timed(tr.frame, 10) for _ in range(16)]
            tr.close()
            also["cube_trajectory"] = {"ms_per_10_frames": prof, "substeps": 16 * 10 * spf,
                                       "value_whole_trajectory": n_total * spf * 160 / (sum(prof) * 1e-3), "unit": UNIT,
                                       "note": "160 frames from the initial state; the peak is the cube at maximal compression"}
        except Exception as e:
            also["cube_trajectory"] = {"error": repr(e)}
        for name, wl, warm, frames, blocks in (("column", "column", 0, 100, True), ("mixed", "mixed", 5, 10, False),
                                               ("dam_slab", "dam", 3, 10, False)):
            try:
                also[name] = run_extra(args, name, wl, local_rank, stream, torch, warm, frames, report_blocks=blocks)
            except Exception as e:  # an extra must never cost the main line
                also[name] = {"error": repr(e)}
        out["also"] = also
    if extras and world > 1:
        # like-for-like base of the weak-scaling line: ONE 2M-particle dam slab alone on rank 0's GPU (others idle)
        if rank == 0:
            try:
                out["scaling_base"] = run_extra(args, "dam_slab", "dam", local_rank, stream, torch, 3, 10)
            except Exception as e:
                out["scaling_base"] = {"error": repr(e)}
        dist.barrier()
        # strong scaling: the fixed 16M-particle dam over the same N GPUs
        try:
            s_scene = build_scene(args, world, "dam-strong")
            s_run = Runner(s_scene, rank, world, local_rank, stream, shard=True)
            for _ in range(3):
                s_run.frame()
            s_samples = [timed(s_run.frame, 5) for _ in range(3)]
            s_ms = float(np.median(s_samples))
            s_nb, s_over = s_run.data.status()
            strong_out = {"workload": workload_name("dam-strong", world, s_run.n_total), "particles": s_run.n_total,
                          "value": s_run.n_total * s_run.spf * 5 / (s_ms * 1e-3), "unit": UNIT,
                          "us_per_substep": s_ms * 1e3 / (5 * s_run.spf), "scaling": "strong",
                          "active_blocks_rank0": int(s_nb), "overflow": bool(s_over)}
            s_run.close()
        except Exception as e:
            strong_out = {"error": repr(e)}
        out.setdefault("also", {})["dam_strong"] = strong_out
    if extras and world == 1:
        # ... and the N = 1 point of the strong-scaling series: all 16M particles on this GPU
        try:
            out["also"]["dam_strong"] = run_extra(args, "dam_strong", "dam-strong", local_rank, stream, torch, 2, 3)
            out["also"]["dam_strong"]["scaling"] = "strong"
        except Exception as e:
            out["also"]["dam_strong"] = {"error": repr(e)}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, scene)
    if rank == 0:
        print(json.dumps(out))
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
    scene = build_scene(args, 1, "dam" if workload == "dam-strong" else workload)
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
        "scaling": "strong" if workload == "dam-strong" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
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
    ap.add_argument("--workload", default="auto", choices=["auto", "cube", "dam", "dam-strong", "column", "mixed"],
                    help="auto: cube (configs[1]) at N=1, sand dam (configs[4], 2M particles per GPU, weak) at N>1, each "
                         "with the other configurations measured alongside (`also`, `scaling_base`)")
    ap.add_argument("--no-extras", action="store_true", help="with --workload auto: only the main workload")
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
