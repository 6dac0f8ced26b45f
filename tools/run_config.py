"""Runs one BASELINE config at full size on one GPU and reports throughput + sanity (diagnostic).
    python tools/run_config.py sand4m|dam16m|dam2m|mixed4m|cube1m [frames]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wgsparkl_b200 import scenes  # noqa: E402
from wgsparkl_b200.pipeline import MpmData, MpmPipeline  # noqa: E402

name = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 5
t0 = time.time()
if name == "sand4m":
    scene = scenes.sand_column_3d(100, 400, 100, grid_capacity=131072)
elif name == "dam16m":
    scene = scenes.sand_dam_3d(400, 200, 200)
elif name == "dam2m":  # one GPU's slab of the N-GPU bench workload
    scene = scenes.sand_dam_3d(50, 200, 200, grid_capacity=65536)
elif name == "mixed4m":
    scene = scenes.mixed_coupled_3d(160, 160, 160, grid_capacity=131072, n_dynamic=4)
else:
    scene = scenes.elastic_cube_3d(100, y_offset=-5.0)
n = len(scene["particles"])
print("%s: %d particles, scene built in %.1f s" % (scene["name"], n, time.time() - t0), flush=True)
stream = torch.cuda.Stream()
pipe = MpmPipeline(0, 3)
pipe.set_stream(stream.cuda_stream)
t0 = time.time()
data = MpmData(pipe, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
print("upload %.1f s" % (time.time() - t0), flush=True)
spf = scene["substeps_per_frame"]
for f in range(frames):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        pipe.queue_step(data, spf)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    nb, overflow = data.status()
    print("frame %d: %.2f ms (%.1f us/substep, %.3e particle-substeps/s) blocks %d overflow %s" % (
        f, ms, ms * 1e3 / spf, n * spf / (ms * 1e-3), nb, overflow), flush=True)
pos = data.read_positions()
print("finite:", bool(np.isfinite(pos).all()), "y range", float(pos[:, 1].min()), float(pos[:, 1].max()))
pipe.set_timestamps(True)
pipe.queue_step(data, spf)
t = pipe.timings_ms()
print({k: round(v * 1000 / spf, 1) for k, v in t.items()})
data.close()
pipe.close()
