"""Diagnostic: host-side time of every call of one end-to-end frame (bench.py frame_e2e), 1M cube."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from wgsparkl_b200 import scenes  # noqa: E402
from wgsparkl_b200.pipeline import MpmData, MpmPipeline  # noqa: E402

scene = scenes.elastic_cube_3d(100, y_offset=-5.0)
stream = torch.cuda.Stream()
pipe = MpmPipeline(0, 3)
pipe.set_stream(stream.cuda_stream)
data = MpmData(pipe, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
poses, vels = bench.frame_io_arrays(scene)
host = torch.empty((2, data.num_particles, 4), dtype=torch.float32).pin_memory().numpy()
names = ["write_poses", "write_vels", "queue_step", "read_poses", "read_pos_async"]
acc = {k: 0.0 for k in names}
frames = 20
for f in range(frames + 3):
    t = [time.perf_counter()]
    data.write_body_poses(poses); t.append(time.perf_counter())
    data.write_body_vels(vels); t.append(time.perf_counter())
    pipe.queue_step(data, 20); t.append(time.perf_counter())
    data.read_body_poses(); t.append(time.perf_counter())
    data.read_positions_async(host[f & 1]); t.append(time.perf_counter())
    if f >= 3:
        for k, n in enumerate(names):
            acc[n] += t[k + 1] - t[k]
pipe.sync()
print({k: round(v / frames * 1e6, 1) for k, v in acc.items()}, "us per frame; total", round(sum(acc.values()) / frames * 1e6, 1))
# blocking D2H alone
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    data.read_positions(host[0])
print("blocking read_positions: %.1f us" % ((time.perf_counter() - t0) / 5 * 1e6))
data.close()
pipe.close()
