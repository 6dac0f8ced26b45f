"""Diagnostic: how the 1M cube's state evolves over frames (strain distribution, collider-affine particles,
active blocks) and what each pass costs at that point.  python tools/state_stats.py [frames]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wgsparkl_b200 import scenes  # noqa: E402
from wgsparkl_b200.pipeline import MpmData, MpmPipeline  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 30
scene = scenes.elastic_cube_3d(100, y_offset=-5.0)
pipe = MpmPipeline(0, 3)
data = MpmData(pipe, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
spf = scene["substeps_per_frame"]
for f in range(frames + 1):
    if f % 5 == 0:
        p = data.read_particles()
        F = p["def_grad"].reshape(-1, 3, 3).astype(np.float64)
        M = np.einsum("nij,nik->njk", F, F) - np.eye(3)
        tr2 = (M * M).sum(axis=(1, 2))
        aff = p["cdf_affinity"]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.queue_step(data, spf)
        e1.record()
        torch.cuda.synchronize()
        total = e0.elapsed_time(e1) * 1000 / spf
        pipe.set_timestamps(True)
        pipe.queue_step(data, spf)
        t = pipe.timings_ms()
        pipe.set_timestamps(False)
        print("frame %3d: blocks %d  tr(M^2)>0.01: %.2f%%  max|M| %.3f  affinity!=0: %.2f%%  y[%.1f, %.1f]  substep %.1f us: g2p %.1f p2g %.1f sort %.1f" % (
            f, data.status()[0], 100.0 * (tr2 > 0.01).mean(), np.sqrt(tr2.max()), 100.0 * (aff != 0).mean(),
            p["position"][:, 1].min(), p["position"][:, 1].max(), total, t["g2p"] * 1000 / spf, t["p2g"] * 1000 / spf, t["grid sort"] * 1000 / spf))
    else:
        pipe.queue_step(data, spf)
data.close()
pipe.close()
