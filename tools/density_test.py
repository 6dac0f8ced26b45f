"""Diagnostic: per-pass cost vs particles-per-cell density at a fixed particle count (1M), undeformed material.
    python tools/density_test.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wgsparkl_b200 import scenes  # noqa: E402
from wgsparkl_b200.pipeline import MpmData, MpmPipeline  # noqa: E402

for name, (nx, ny, nz), (sx, sy, sz) in (("8 ppc", (100, 100, 100), (0.5, 0.5, 0.5)),
                                        ("16 ppc", (100, 100, 100), (0.5, 0.25, 0.5)),
                                        ("32 ppc", (100, 100, 100), (0.5, 0.125, 0.5)),
                                        ("4 ppc", (100, 100, 100), (0.5, 1.0, 0.5))):
    scene = scenes.elastic_cube_3d(100, y_offset=-5.0, jitter=False)
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    rng = np.random.default_rng(3)
    pos = np.stack([(i.ravel() + 0.5 - nx / 2) * sx, (j.ravel() + 0.5) * sy - 2.3, (k.ravel() + 0.5 - nz / 2) * sz], axis=1)
    pos += rng.uniform(-0.05, 0.05, size=pos.shape) * np.array([sx, sy, sz])
    scene["particles"]["position"][:] = pos.astype(np.float32)
    pipe = MpmPipeline(0, 3)
    data = MpmData(pipe, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    spf = 20
    pipe.queue_step(data, 2)
    pipe.set_timestamps(True)
    pipe.queue_step(data, spf)
    t = pipe.timings_ms()
    pipe.set_timestamps(False)
    print("%-7s blocks %5d  g2p %.1f  p2g %.1f  sort %.1f us" % (name, data.status()[0], t["g2p"] * 1000 / spf, t["p2g"] * 1000 / spf, t["grid sort"] * 1000 / spf))
    data.close()
    pipe.close()
