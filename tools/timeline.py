"""Timeline of ONE substep inside the graph replay (who starts / ends when, gaps, overlaps), from %globaltimer stamps
that the kernels take at their first and last CTA while b200mpm_debug_timeline has the recording switched on:
    python tools/timeline.py cube1m [substeps_before]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wgsparkl_b200 import scenes  # noqa: E402
from wgsparkl_b200.pipeline import MpmData, MpmPipeline  # noqa: E402

name = sys.argv[1]
before = int(sys.argv[2]) if len(sys.argv) > 2 else 40
if name == "dam2m":
    scene = scenes.sand_dam_3d(50, 200, 200, grid_capacity=65536)
elif name == "sand4m":
    scene = scenes.sand_column_3d(100, 400, 100, grid_capacity=131072)
else:
    scene = scenes.elastic_cube_3d(100, y_offset=-5.0)
pipe = MpmPipeline(0, 3)
data = MpmData(pipe, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
pipe.queue_step(data, before)
pipe.sync()
rows = []
for rep in range(3):
    data.debug_timeline()  # reset
    pipe.queue_step(data, 3)  # three consecutive graph replays: the middle one is in steady state
    pipe.sync()
    tl = data.debug_timeline()
    rows.append(tl)
tl = rows[-1]
t0 = min(v[0] for v in tl.values() if v)
print("three substeps, first start / last end over all three (us relative to the first kernel):")
for k, v in sorted(tl.items(), key=lambda kv: kv[1][0] if kv[1] else 1 << 62):
    if v:
        print("  %-16s %8.1f .. %8.1f" % (k, (v[0] - t0) / 1e3, (v[1] - t0) / 1e3))
# single substep
data.debug_timeline()
pipe.queue_step(data, 1)
pipe.sync()
tl = data.debug_timeline()
t0 = min(v[0] for v in tl.values() if v)
print("one substep (us relative to its first kernel):")
prev_end = None
for k, v in sorted(tl.items(), key=lambda kv: kv[1][0] if kv[1] else 1 << 62):
    if v:
        print("  %-16s start %7.1f  end %7.1f  (%.1f us)" % (k, (v[0] - t0) / 1e3, (v[1] - t0) / 1e3, (v[1] - v[0]) / 1e3))
data.close()
pipe.close()
