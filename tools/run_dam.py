"""Diagnostic: one slab's worth of the sand dam (2M particles, the rank-0 slab geometry of the N-GPU bench) on one
GPU, per-pass times.  python tools/run_dam.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wgsparkl_b200 import scenes  # noqa: E402
from wgsparkl_b200.pipeline import MpmData, MpmPipeline  # noqa: E402

scene = scenes.sand_dam_3d(50, 200, 200, grid_capacity=65536)
n = len(scene["particles"])
stream = torch.cuda.Stream()
pipe = MpmPipeline(0, 3)
pipe.set_stream(stream.cuda_stream)
data = MpmData(pipe, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
for f in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        pipe.queue_step(data, 20)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
print("dam slab %d particles: %.1f us/substep, %.3e particle-substeps/s" % (n, ms * 1e3 / 20, n * 20 / (ms * 1e-3)))
pipe.set_timestamps(True)
pipe.queue_step(data, 20)
t = pipe.timings_ms()
print({k: round(v * 1000 / 20, 1) for k, v in t.items()})
data.close()
pipe.close()
